#!/usr/bin/env python3
"""Independent numerical pins for the [EXT] stages of the oracle (PCL VoxelGrid / NormalEstimation / FPFHEstimation, FLANN 1-NN,
the fp64 TIM mask, pmc core numbers).

This script does NOT import, link or call anything under oracle/ or the CUDA library: every stage is re-derived here from the
published algorithms in float64 numpy / scipy (cKDTree radius search, numpy.linalg.eigh, arctan2) and networkx, so the
fixtures it writes (tests/golden/independent_*.npz) can catch drift of the oracle itself.  tests/test_independent_pins.py
compares the oracle with these fixtures on the statistical tier of SURVEY.md 8(c): exact voxel membership, normals within
1e-3 rad except at degenerate neighbourhoods, >= 99 % of points with max FPFH-bin difference <= 1e-2, mutual nearest
neighbours identical up to float ties, adjacency identical except on the knife edge, core numbers exact.

    python tools/gen_independent_pins.py          # rewrites tests/golden/independent_seed{31,32}.npz
"""
from __future__ import annotations

import sys
from pathlib import Path

import networkx as nx
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from quatro_b200 import synth  # noqa: E402  (the scan generator: host-side input synthesis, not the oracle)

LEAF, RN, RF, BETA = 0.3, 0.5, 0.75, 0.6


sys.path.insert(0, str(ROOT / "tests"))
from independent_ref import voxel_grid, normals_pcl, mutual_nn, tim_graph, fpfh_pcl  # noqa: E402


def main():
    out_dir = ROOT / "tests" / "golden"
    for seed in (31, 32):
        src, tgt, T = synth.outdoor_pair(seed, rings=32, azimuths=900)
        rec = {"T_gt": T}
        desc = {}
        vox = {}
        for name, raw in (("src", src), ("tgt", tgt)):
            keep = raw[raw[:, 3] >= 0][:, :3].astype(np.float64)  # the pipeline drops flagged (ground) points
            cent, idx, uniq = voxel_grid(keep, LEAF)
            nrm, gap = normals_pcl(cent.astype(np.float32).astype(np.float64), RN)
            rec[f"{name}_gap"] = gap
            fp, sp = fpfh_pcl(cent.astype(np.float32).astype(np.float64), nrm, RF)
            rec[f"{name}_vox"] = cent
            rec[f"{name}_normals"] = nrm
            rec[f"{name}_fpfh"] = fp.astype(np.float32)
            desc[name], vox[name] = fp, cent
        corr, mr, mc = mutual_nn(desc["src"], desc["tgt"])
        rec["mutual"] = corr.astype(np.int32)
        rec["margin_row"] = mr
        rec["margin_col"] = mc
        # graph + core numbers on the mutual-NN correspondences (no tuple test: deterministic without the RNG)
        a = vox["src"][corr[:, 0]].astype(np.float32).astype(np.float64)
        b = vox["tgt"][corr[:, 1]].astype(np.float32).astype(np.float64)
        e, knife = tim_graph(a, b, BETA)
        g = nx.from_numpy_array(e)
        core = nx.core_number(g)
        rec["graph_a"] = a.astype(np.float32)
        rec["graph_b"] = b.astype(np.float32)
        rec["adj"] = np.packbits(e, axis=1, bitorder="little")
        rec["knife"] = np.packbits(knife, axis=1, bitorder="little")
        rec["core"] = np.array([core[i] for i in range(len(a))], np.int32)
        clique, _ = nx.max_weight_clique(g, weight=None)
        rec["max_clique_size"] = np.int32(len(clique))
        np.savez_compressed(out_dir / f"independent_seed{seed}.npz", **rec)
        print(f"seed {seed}: vox {len(vox['src'])}/{len(vox['tgt'])}, mutual {len(corr)}, edges {int(e.sum()) // 2}, "
              f"max core {max(core.values())}, exact max clique {len(clique)}")


if __name__ == "__main__":
    main()
