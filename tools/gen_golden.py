#!/usr/bin/env python3
"""Generate the golden stage-boundary fixtures under tests/golden/ from the CPU oracle.

The reference ships no golden vectors, so these dumps -- produced by the oracle
whose semantics are pinned by tests/test_oracle_kat.py -- are the regression pins for every stage
boundary: both the oracle (CPU tests) and the CUDA path (GPU tests) must reproduce them exactly.
Run from the repo root:  python tools/gen_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import Oracle  # noqa: E402
from quatro_b200 import synth  # noqa: E402
from quatro_b200.capi import default_params, PMC_HEU  # noqa: E402


def result_dict(res):
    d = res.as_dict()
    return {f"res_{k}": np.asarray(v) for k, v in d.items()}


def main():
    o = Oracle()
    o.set_num_threads(1)
    out = ROOT / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    p = default_params()

    # ---- front end + full pipeline on small synthetic scans (16 rings x 360 azimuths) ----
    for seed in (101, 102):
        src, tgt, T = synth.outdoor_pair(seed, rings=16, azimuths=360)
        sv, _ = o.voxelize(src, p.voxel_size, 1)
        tv, _ = o.voxelize(tgt, p.voxel_size, 1)
        sn, sd, ss = o.compute_fpfh(sv, p.normal_radius, p.fpfh_radius, p.fpfh_radius, want_spfh=True)
        tn, td = o.compute_fpfh(tv, p.normal_radius, p.fpfh_radius, p.fpfh_radius)
        corr, n_mutual, _, mutual = o.match(sv, sd, tv, td, p, want_mutual=True)
        res, st = o.register_pair(src, tgt, p)
        np.savez_compressed(out / f"pipeline_seed{seed}.npz", src=src, tgt=tgt, T_gt=T, src_vox=sv, tgt_vox=tv, src_normals=sn,
                            src_desc=sd, src_spfh=ss, tgt_normals=tn, tgt_desc=td, corr=corr, mutual=mutual, status=np.int32(st),
                            **result_dict(res))
        print(f"pipeline_seed{seed}: {len(src)}+{len(tgt)} pts -> {len(sv)}/{len(tv)} voxels, {n_mutual} mutual, L={len(corr)}, "
              f"clique={res.clique_size}, valid={res.valid}")

    # ---- back end on matched pairs ----
    for seed, L, ratio in ((201, 160, 0.35), (202, 700, 0.2)):
        a4, b4, T, inl = synth.matched_pairs(seed, L, inlier_ratio=ratio, noise=0.04)
        adj, deg, ne = o.build_graph(a4, b4, p.noise_bound, p.cbar2)
        clique, kcore, order, max_core = o.max_clique(adj, PMC_HEU)
        res, rm, tm, st = o.solve_pose(a4, b4, clique, p)
        res2, st2, clique2, fin = o.solve_correspondences(a4, b4, p, want_sets=True)
        assert np.array_equal(clique, clique2)
        np.savez_compressed(out / f"solver_seed{seed}.npz", a4=a4, b4=b4, T_gt=T, inlier_mask=inl, adj=adj, degree=deg, n_edges=np.int64(ne),
                            clique=clique, kcore=kcore, kcore_order=order, max_core=np.int32(max_core), rot_mask=rm, trans_mask=tm,
                            final_inliers=fin, status=np.int32(st2), **result_dict(res2))
        print(f"solver_seed{seed}: L={L} E={ne} max_core={max_core} clique={len(clique)} final={len(fin)}")


if __name__ == "__main__":
    main()
