"""The oracle against INDEPENDENT float64 numpy / scipy / networkx restatements of the third-party stages (fixtures written by
tools/gen_independent_pins.py, which never touches oracle/): the statistical parity tier of SURVEY.md 8(c).  These pins do not
replace reference outputs (none exist: the reference cannot be built here) but they catch drift of the oracle itself."""
from pathlib import Path

import numpy as np
import pytest

from quatro_b200 import synth
from quatro_b200.capi import default_params, PMC_HEU

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module", params=[31, 32])
def pin(request):
    seed = request.param
    d = dict(np.load(GOLD / f"independent_seed{seed}.npz"))
    src, tgt, T = synth.outdoor_pair(seed, rings=32, azimuths=900)
    d["src_raw"], d["tgt_raw"] = src, tgt
    assert np.allclose(T, d["T_gt"])
    return d


def test_voxel_centroids_against_numpy(oracle, pin):
    for name in ("src", "tgt"):
        vox, st = oracle.voxelize(pin[f"{name}_raw"], 0.3, 1)
        ref = pin[f"{name}_vox"]
        assert st == 0 and len(vox) == len(ref), "voxel membership differs from the numpy VoxelGrid restatement"
        assert np.abs(vox[:, :3] - ref).max() < 2e-5   # float32 running sums vs float64 means


def test_normals_against_eigh(oracle, pin):
    """PCL 1.8's computeMeanAndCovarianceMatrix accumulates the raw second moments in float32 (single pass, no centring), which
    the oracle restates; against a centred float64 covariance the normal direction therefore degrades with |p|^2 / sigma^2 * 2^-24
    (about 1e-3 rad at 20 m, 1e-2 rad at 60 m for 0.5 m neighbourhoods).  Tiers: well-defined normals (eigen-gap > 0.1) within
    20 m of the sensor agree to 1e-3 rad, everywhere to 2e-2 rad."""
    for name in ("src", "tgt"):
        vox = np.ones((len(pin[f"{name}_vox"]), 4), np.float32)
        vox[:, :3] = pin[f"{name}_vox"]
        n_o, _ = oracle.compute_fpfh(vox, 0.5, 0.75, 0.3)
        n_r, gap = pin[f"{name}_normals"], pin[f"{name}_gap"]
        nan_o, nan_r = np.isnan(n_o[:, 0]), np.isnan(n_r[:, 0])
        assert np.array_equal(nan_o, nan_r), "different sets of points without a normal (< 3 neighbours)"
        ok = ~nan_o & (gap > 0.1)
        dots = (n_o[ok, :3] * n_r[ok, :3]).sum(1)
        ang = np.arccos(np.clip(np.abs(dots), -1, 1))
        rng = np.linalg.norm(vox[ok, :3], axis=1)
        near = rng < 20.0
        assert near.sum() > 500
        assert (ang[near] < 1e-3).mean() >= 0.97, f"{name}: {(ang[near] < 1e-3).mean():.4f} of the near normals within 1e-3 rad of eigh()"
        assert (ang < 1e-3).mean() >= 0.88 and (ang < 2e-2).mean() >= 0.995, (name, (ang < 1e-3).mean(), (ang < 2e-2).mean())
        # orientation: flipped towards the viewpoint (0,0,0); the two may disagree only on the knife edge n . p ~ 0
        view = np.abs((n_r[ok, :3] * vox[ok, :3]).sum(1)) / rng
        assert (dots[(ang < 2e-2) & (view > 1e-3)] > 0).all(), "a normal points away from the viewpoint"
        curv_o, curv_r = n_o[ok, 3], n_r[ok, 3]
        assert np.median(np.abs(curv_o - curv_r)) < 1e-3


def test_fpfh_against_float64_restatement(oracle, pin):
    """SPFH + FPFH of the oracle against the float64 restatement of the published algorithm fed with the SAME (oracle) normals, so
    that only the feature / histogram arithmetic is compared: a bin can flip only when a Darboux feature sits within float32
    rounding of a bin edge (>= 99 % of the points within 1e-2 per bin)."""
    from independent_ref import fpfh_pcl
    vox = np.ones((len(pin["src_vox"]), 4), np.float32)
    vox[:, :3] = pin["src_vox"]
    n_o, d_o = oracle.compute_fpfh(vox, 0.5, 0.75, 0.3)
    d_r, _ = fpfh_pcl(vox[:, :3].astype(np.float64), n_o.astype(np.float64), 0.75)
    diff = np.abs(d_o - d_r).max(1)
    frac = (diff <= 1e-2).mean()
    assert frac >= 0.99, f"FPFH: only {frac:.4f} of the points within 1e-2 per bin of the float64 restatement"
    # histogram invariants of the published algorithm
    sums = d_o.reshape(len(d_o), 3, 11).sum(2)
    live = sums[:, 0] > 0
    assert np.allclose(sums[live], 100.0, atol=1e-2)
    # and with independent (eigh) normals the descriptors stay close in the mean although bins flip
    close = np.abs(d_o - pin["src_fpfh"]).max(1)
    assert np.median(close) < 0.5


def test_mutual_nearest_neighbours_against_numpy(oracle, pin):
    vs = np.ones((len(pin["src_vox"]), 4), np.float32); vs[:, :3] = pin["src_vox"]
    vt = np.ones((len(pin["tgt_vox"]), 4), np.float32); vt[:, :3] = pin["tgt_vox"]
    # identical descriptors on both sides of the comparison: the fixture's float32 FPFH
    p = default_params(); p.use_tuple_test = 0
    corr, n_mutual, st = oracle.match(vs, pin["src_fpfh"], vt, pin["tgt_fpfh"], p)
    ref = pin["mutual"]
    a = {tuple(x) for x in corr.tolist()}
    b = {tuple(x) for x in ref.tolist()}
    # float32 fma chain vs float64 Gram form: only pairs whose best/second-best margin is at rounding level may differ
    sym = a ^ b
    assert len(sym) <= max(4, len(b) // 50), f"{len(sym)} of {len(b)} mutual pairs differ from the float64 brute force"
    for i, j in sym:
        assert pin["margin_row"][i] < 1e-2 or pin["margin_col"][j] < 1e-2, "a differing pair is not a near-tie"


def test_graph_and_core_numbers_against_numpy_networkx(oracle, pin):
    a = np.ones((len(pin["graph_a"]), 4), np.float32); a[:, :3] = pin["graph_a"]
    b = np.ones((len(pin["graph_b"]), 4), np.float32); b[:, :3] = pin["graph_b"]
    L = len(a)
    adj, deg, ne = oracle.build_graph(a, b, 0.3, 1.0)
    got = np.unpackbits(adj.view(np.uint8), axis=1, bitorder="little")[:, :L].astype(bool)
    ref = np.unpackbits(pin["adj"], axis=1, bitorder="little")[:, :L].astype(bool)
    knife = np.unpackbits(pin["knife"], axis=1, bitorder="little")[:, :L].astype(bool)
    assert np.array_equal(got | knife, ref | knife), "adjacency differs from the float64 numpy mask away from the knife edge"
    assert np.array_equal(got, ref)   # and in fact everywhere: both are the literal fp64 expression
    clique, kcore, order, max_core = oracle.max_clique(adj, PMC_HEU)
    assert np.array_equal(kcore - 1, pin["core"]), "core numbers differ from networkx.core_number"
    assert max_core == pin["core"].max()
    # the peel order is a degeneracy ordering: every vertex has at most core(v) neighbours later in the order
    posn = np.empty(L, np.int64); posn[order] = np.arange(L)
    later = (got & (posn[None, :] > posn[:, None])).sum(1)
    assert (later <= pin["core"]).all()
    # the heuristic clique is a clique, bounded by the exact maximum (networkx) and by max_core + 1
    assert got[np.ix_(clique, clique)].sum() == len(clique) * (len(clique) - 1)
    assert len(clique) <= int(pin["max_clique_size"]) <= max_core + 1
