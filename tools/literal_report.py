#!/usr/bin/env python3
"""Measure how often the literal mode of the oracle (libstdc++ std::sort on tied keys, distance-ordered accumulation) changes
correspondences / clique membership / pose w.r.t. the canonical mode, on full-size 64-ring pairs (CPU only).

    python tools/literal_report.py [n_pairs]     -> profiles/r02_literal_vs_canonical.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import Oracle  # noqa: E402
from test_literal_mode import compare_modes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rows = compare_modes(Oracle(), range(n), rings=64, azimuths=1800)
summary = {
    "pairs": n, "scan": "64 rings x 1800 azimuths (street scene), config/params.yaml defaults",
    "voxel_centroids_differing_mean": sum(r["voxel_centroids_differ"] for r in rows) / n,
    "pairs_with_identical_correspondences": sum(r["corr_identical"] for r in rows),
    "mean_correspondence_symmetric_difference": sum(r["corr_symdiff"] for r in rows) / n,
    "pairs_with_identical_clique": sum(r["clique_identical"] for r in rows),
    "pairs_with_equal_clique_size": sum(r["clique_sizes"][0] == r["clique_sizes"][1] for r in rows),
    "both_valid": sum(r["valid"] == (1, 1) for r in rows),
    "max_rot_gap_deg": max(r["rot_deg"] for r in rows), "max_trans_gap_m": max(r["trans_m"] for r in rows),
    "pairs_within_2deg_0.3m": sum(r["rot_deg"] < 2.0 and r["trans_m"] < 0.3 for r in rows),
}
(ROOT / "profiles" / "r02_literal_vs_canonical.json").write_text(json.dumps({"summary": summary, "pairs": rows}, indent=1))
print(json.dumps(summary, indent=1))
