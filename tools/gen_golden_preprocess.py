#!/usr/bin/env python3
"""Golden fixtures of the rows added late in round 2 -- ground removal, range-image sub-cluster rejection, exact clique -- produced by
the CPU oracle (whose semantics are pinned independently by tests/test_preprocess.py and tests/test_oracle_kat.py): regression pins for
the oracle (CPU tests) and the CUDA path (GPU tests).   python tools/gen_golden_preprocess.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import Oracle  # noqa: E402
from quatro_b200 import synth  # noqa: E402
from quatro_b200.capi import default_patchwork_params, default_segment_params, PMC_EXACT, PMC_HEU  # noqa: E402


def graph(seed, n, p, planted=0):
    rng = np.random.default_rng(seed)
    R = rng.uniform(size=(n, n)) < p
    R = np.triu(R, 1); R = R | R.T
    if planted:
        m = rng.choice(n, planted, replace=False)
        R[np.ix_(m, m)] = True
    np.fill_diagonal(R, False)
    adj = np.zeros((n, (n + 31) // 32), np.uint32)
    i, j = np.nonzero(R)
    np.bitwise_or.at(adj, (i, j >> 5), (np.uint32(1) << (j & 31).astype(np.uint32)))
    return adj


def main():
    o = Oracle()
    out = ROOT / "tests" / "golden"
    pp, sp = default_patchwork_params(), default_segment_params()
    pp.num_min_pts = 10          # the constructor default of the reference (patchwork.hpp:57); the yaml's 80 empties most patches of a quarter-density scan
    sp.horizon_scan, sp.ang_res_x = 450, 0.8   # 64 rings x 450 azimuths: a quarter of the columns keeps the fixture small
    src = synth.outdoor_pair(301, rings=64, azimuths=450)[0]
    g, ng, st = o.patchwork(src, pp)
    v, ol = o.segment_cloud(ng, sp)
    np.savez_compressed(out / "preprocess_seed301.npz", scan=src, ground=g, nonground=ng, valid=v, outliers=ol, num_min_pts=np.int32(pp.num_min_pts),
                        horizon_scan=np.int32(sp.horizon_scan), ang_res_x=np.float32(sp.ang_res_x))
    print(f"preprocess_seed301: {len(src)} points -> ground {len(g)}, non-ground {len(ng)} -> valid {len(v)}, outliers {len(ol)}")
    recs = {}
    for name, (seed, n, p, planted, limit) in {"g300": (1, 300, 0.1, 0, 0), "g600": (2, 600, 0.2, 0, 0), "planted": (3, 400, 0.05, 30, 0),
                                                 "cut": (4, 300, 0.7, 0, 3000)}.items():
        adj = graph(seed, n, p, planted)
        c, k, order, mc, fl = o.max_clique_ex(adj, PMC_EXACT, 0.5, limit)
        h = o.max_clique(adj, PMC_HEU)[0]
        recs[f"{name}_adj"] = adj; recs[f"{name}_clique"] = c; recs[f"{name}_heuristic"] = h
        recs[f"{name}_flags"] = np.int32(fl); recs[f"{name}_limit"] = np.int64(limit)
        print(f"exact {name}: n={n} heuristic {len(h)} exact {len(c)} flags {fl}")
    np.savez_compressed(out / "exact_clique.npz", **recs)


if __name__ == "__main__":
    main()
