"""Pre-processing before the path: ground removal, include/patchwork.hpp:329-455.

CPU part: the oracle restatement against properties that do not depend on it (the generator's own ground labels, a numpy plane
model) and the reference's documented edge behaviour.  GPU part (-m gpu): qb200_patchwork through the C-ABI is bit-identical to the
oracle -- same points, same order, in both outputs."""
import numpy as np
import pytest

from quatro_b200 import synth
from quatro_b200.capi import default_patchwork_params


def _scene(seed, n_obj=40):
    """Flat ground at z = -1.723 seen from the origin + boxes standing on it: labels known by construction."""
    rng = np.random.default_rng(seed)
    az = rng.uniform(0, 2 * np.pi, 60000)
    r = rng.uniform(3.0, 70.0, 60000)
    g = np.stack([r * np.cos(az), r * np.sin(az), -1.723 + rng.normal(0, 0.02, len(r))], 1)
    objs = []
    for _ in range(n_obj):
        c = rng.uniform(-50, 50, 2)
        if np.hypot(*c) < 5:
            continue
        k = 600
        # walls of a 2 x 2 x 2.5 m box
        side = rng.integers(0, 4, k)
        u, v = rng.uniform(-1, 1, k), rng.uniform(-1.4, 1.1, k)
        x = np.where(side < 2, np.where(side == 0, -1.0, 1.0), u) + c[0]
        y = np.where(side < 2, u, np.where(side == 2, -1.0, 1.0)) + c[1]
        objs.append(np.stack([x, y, v], 1))
    o = np.concatenate(objs)
    pts = np.concatenate([g, o]).astype(np.float32)
    lab = np.concatenate([np.ones(len(g), bool), np.zeros(len(o), bool)])
    perm = rng.permutation(len(pts))
    out = np.ones((len(pts), 4), np.float32)
    out[:, :3] = pts[perm]
    out[:, 3] = np.where(lab[perm], -1.0, 1.0)      # w < 0 marks true ground (travels with the point)
    return out


def test_patchwork_separates_ground_from_objects(oracle):
    pp = default_patchwork_params()
    pts = _scene(1)
    g, ng, st = oracle.patchwork(pts, pp)
    assert st == 0
    r = np.hypot(pts[:, 0], pts[:, 1])
    assert len(g) + len(ng) <= int(((r > pp.min_range) & (r <= pp.max_range)).sum())   # nothing is duplicated
    # ground output is ground; objects end in the non-ground output (walls standing on the plane leave a thin skirt: th_dist)
    assert (g[:, 3] < 0).mean() > 0.97
    wall = ng[ng[:, 3] > 0]
    assert len(wall) > 0.85 * (pts[:, 3] > 0).sum()
    assert (ng[:, 3] < 0).sum() < 0.1 * (pts[:, 3] < 0).sum()
    # generator scans: the ground flag of the synthetic LiDAR agrees with the estimate
    src, _, _ = synth.outdoor_pair(5, rings=32, azimuths=900)
    g, ng, st = oracle.patchwork(src, pp)
    assert (g[:, 3] < 0).mean() > 0.95 and (ng[:, 3] < 0).mean() < 0.02


def test_patchwork_order_and_edge_cases(oracle):
    pp = default_patchwork_params()
    pts = _scene(2, n_obj=10)
    g, ng, _ = oracle.patchwork(pts, pp)
    # every output point is an input point, bit for bit, used once
    key = lambda a: {tuple(x) for x in a.view(np.uint32).reshape(len(a), 4).tolist()}
    assert key(g) | key(ng) <= key(pts) and not (key(g) & key(ng))
    # permutation of the input changes nothing but tie order: ground / non-ground SETS are equal
    perm = np.random.default_rng(0).permutation(len(pts))
    g2, ng2, _ = oracle.patchwork(pts[perm], pp)
    assert key(g2) == key(g) and key(ng2) == key(ng)
    # empty, all-NaN, below the mirror-reflection cut (-1.8 h), fewer than num_min_pts per patch -> nothing comes out
    for bad in (np.zeros((0, 4), np.float32), np.full((100, 4), np.nan, np.float32),
                np.array([[10, 0, -5.0, 1]] * 200, np.float32), pts[:50]):
        g0, n0, st = oracle.patchwork(bad, pp)
        assert st == 0 and len(g0) == 0 and len(n0) == 0
    # a tilted / vertical patch is rejected as a whole: its "ground" part joins the non-ground output first (patchwork.hpp:399-402)
    rng = np.random.default_rng(3)
    wall = np.stack([np.full(400, 6.0), rng.uniform(-1.0, 1.0, 400), rng.uniform(-1.7, 1.0, 400), np.ones(400)], 1).astype(np.float32)
    g3, n3, _ = oracle.patchwork(wall, pp)
    assert len(g3) == 0 and len(n3) == 400


@pytest.mark.gpu
def test_patchwork_gpu_matches_oracle(handle, oracle):
    pp = default_patchwork_params()
    scans = [synth.outdoor_pair(11)[0], synth.outdoor_pair(12, rings=32, azimuths=900)[1], _scene(4)]
    scans[2][::97, 2] = np.nan                       # non-finite points are dropped
    scans[2][5::113, 2] = -0.0
    for pts in scans:
        g_o, n_o, st_o = oracle.patchwork(pts, pp)
        g_g, n_g, st_g = handle.patchwork(pts, pp)
        assert st_g == st_o == 0
        assert g_g.shape == g_o.shape and n_g.shape == n_o.shape
        assert np.array_equal(g_g.view(np.uint32), g_o.view(np.uint32)), "ground output differs (points or order)"
        assert np.array_equal(n_g.view(np.uint32), n_o.view(np.uint32)), "non-ground output differs (points or order)"
    # other parameters: more iterations, global elevation test, looser uprightness
    pp2 = default_patchwork_params()
    pp2.num_iter, pp2.using_global_elevation, pp2.uprightness_thr, pp2.num_min_pts = 5, 1, 0.5, 10
    g_o, n_o, _ = oracle.patchwork(scans[0], pp2)
    g_g, n_g, _ = handle.patchwork(scans[0], pp2)
    assert np.array_equal(g_g.view(np.uint32), g_o.view(np.uint32)) and np.array_equal(n_g.view(np.uint32), n_o.view(np.uint32))
    # edge cases
    for bad in (np.zeros((0, 4), np.float32), np.full((100, 4), np.nan, np.float32), scans[0][:50]):
        g_g, n_g, st = handle.patchwork(bad, pp)
        assert st == 0 and len(g_g) == 0 and len(n_g) == 0
    # the ground-free scan goes through the registration path like a flagged one
    from quatro_b200.capi import default_params
    src, tgt, T = synth.outdoor_pair(13)
    p = default_params()
    p.skip_flagged = 0
    ns, nt = handle.patchwork(src, pp)[1], handle.patchwork(tgt, pp)[1]
    res, st = handle.register_pair(ns, nt, p)
    ref, st_ref = oracle.register_pair(oracle.patchwork(src, pp)[1], oracle.patchwork(tgt, pp)[1], p)
    assert st == st_ref == 0 and res.n_corr == ref.n_corr and res.clique_size == ref.clique_size
    assert np.allclose(res.matrix(), ref.matrix(), atol=1e-9)
    rot, tr = synth.pose_error(res.matrix(), T)
    assert rot < 2.0 and tr < 1.0      # the same pair with the generator's ground flags: 0.99 deg / 0.55 m


# ---- range-image sub-cluster removal (include/imageProjection.hpp:273-294) ------------------------------------------------
def _numpy_segments(pts, sp):
    """Independent restatement: numpy float32 projection, scipy connected components, segment statistics."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    H, W = sp.n_scan, sp.horizon_scan
    p = pts[np.isfinite(pts[:, :3]).all(1)]
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    vert = (np.arctan2(z, np.sqrt(x * x + y * y)).astype(np.float32) * np.float32(180.0)).astype(np.float64) / np.pi
    rf = (vert.astype(np.float32) + np.float32(sp.ang_bottom)) / np.float32(sp.ang_res_y)
    hor = ((np.arctan2(x, y).astype(np.float32) * np.float32(180.0)).astype(np.float64) / np.pi).astype(np.float32)
    qd = (hor.astype(np.float64) - 90.0) / np.float64(np.float32(sp.ang_res_x))
    col = (-(np.sign(qd) * np.floor(np.abs(qd) + 0.5)) + W // 2).astype(np.int64)      # C round(): halves away from zero
    col = np.where(col >= W, col - W, col)
    rng = np.sqrt(x * x + y * y + z * z)
    ok = (rf > -1) & (rf < H) & (col >= 0) & (col < W) & (rng >= 0.1)
    row = np.trunc(rf).astype(np.int64)
    pix = (row * W + col)[ok]
    win = np.full(H * W, -1, np.int64)
    np.maximum.at(win, pix, np.nonzero(ok)[0])
    img = np.full(H * W, np.inf, np.float32)
    img[win >= 0] = rng[win[win >= 0]]
    img = img.reshape(H, W)
    offs = {0: [(-1, 0), (0, 1), (0, -1), (1, 0)], 2: [(-1, -1), (-1, 1), (1, 1), (1, -1)]}[sp.neighbor_mode] if sp.neighbor_mode != 1 else \
        [(-1, 0), (0, 1), (0, -1), (1, 0), (-1, -1), (-1, 1), (1, 1), (1, -1)]
    ax, ay = np.float32(np.float64(sp.ang_res_x) / 180 * np.pi), np.float32(np.float64(sp.ang_res_y) / 180 * np.pi)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    src, dst = [], []
    for di, dj in offs:
        ti, tj = ii + di, (jj + dj) % W
        inb = (ti >= 0) & (ti < H)
        a, b = img[ii[inb], jj[inb]], img[ti[inb], tj[inb]]
        both = np.isfinite(a) & np.isfinite(b)
        d1, d2 = np.maximum(a, b), np.minimum(a, b)
        al = ax if di == 0 else ay
        with np.errstate(invalid="ignore"):
            ang = np.arctan2(d2 * np.sin(al), d1 - d2 * np.cos(al))
        e = both & (ang > sp.segment_theta)
        src.append((ii[inb] * W + jj[inb])[e]); dst.append((ti[inb] * W + tj[inb])[e])
    src, dst = np.concatenate(src), np.concatenate(dst)
    n_comp, lab = connected_components(coo_matrix((np.ones(len(src)), (src, dst)), shape=(H * W, H * W)), directed=False)
    occupied = np.nonzero(win >= 0)[0]
    n_valid = n_out = 0
    order = np.argsort(lab[occupied], kind="stable")
    groups = np.split(occupied[order], np.nonzero(np.diff(lab[occupied][order]))[0] + 1)
    for gpx in groups:
        rows_wo_seed = np.unique(gpx[1:] // W)            # gpx is ascending: gpx[0] is the seed of the row-major sweep
        ok_seg = len(gpx) >= sp.min_pts_for_subclustering or (len(gpx) >= sp.segment_valid_point_num and len(rows_wo_seed) >= sp.segment_valid_line_num)
        if ok_seg: n_valid += len(gpx)
        else: n_out += len(gpx)
    return n_valid, n_out


def test_segment_cloud_against_numpy_scipy(oracle):
    from quatro_b200.capi import default_segment_params
    pp = default_patchwork_params()
    src, _, _ = synth.outdoor_pair(21)
    ng = oracle.patchwork(src, pp)[1]
    # The synthetic LiDAR's rings sit EXACTLY on the row boundaries of the range image (elevation = -25 + k * 26.9 / 63 degrees), where
    # the last ulp of atan2f decides the row: a slightly tilted and shifted sensor frame moves them off the boundaries, so that the
    # comparison with numpy's atan2 (a different last ulp) is about the algorithm, not about rounding.
    a, b = np.radians(0.37), np.radians(-0.23)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    ng = ng.copy()
    ng[:, :3] = (ng[:, :3].astype(np.float64) @ (Rx @ Ry).T + np.array([0.11, -0.05, 0.07])).astype(np.float32)
    for mode in (2, 0, 1):
        sp = default_segment_params()
        sp.neighbor_mode = mode
        v, o = oracle.segment_cloud(ng, sp)
        nv, no = _numpy_segments(ng, sp)
        assert abs(len(v) + len(o) - (nv + no)) <= 3               # same occupied pixels (a point on a pixel boundary may move)
        assert abs(len(v) - nv) <= 0.005 * (nv + no) + 5, (mode, len(v), nv)   # float32 atan2 differs in the last ulp at a few edges
        assert len(v) > 0.5 * (nv + no)
        # outputs are input points (w = 1), each at most once
        allp = {tuple(r) for r in ng[:, :3].view(np.uint32).tolist()}
        got = [tuple(r) for r in np.concatenate([v, o])[:, :3].view(np.uint32).tolist()]
        assert set(got) <= allp and len(set(got)) == len(got)
    # nothing in, nothing out; isolated points are outliers
    sp = default_segment_params()
    v, o = oracle.segment_cloud(np.zeros((0, 4), np.float32), sp)
    assert len(v) == 0 and len(o) == 0
    lone = np.array([[10, 0, 0, 1], [0, 20, -2, 1], [-30, 5, 0.5, 1]], np.float32)
    v, o = oracle.segment_cloud(lone, sp)
    assert len(v) == 0 and len(o) == 3


@pytest.mark.gpu
def test_segment_cloud_gpu_matches_oracle(handle, oracle):
    from quatro_b200.capi import default_segment_params, default_params
    pp = default_patchwork_params()
    clouds = [oracle.patchwork(synth.outdoor_pair(31)[0], pp)[1], synth.outdoor_pair(32, rings=32, azimuths=900)[1], _scene(6)]
    clouds[2][::53, 0] = np.nan
    for pts in clouds:
        for mode in (2, 0, 1):
            sp = default_segment_params()
            sp.neighbor_mode = mode
            v_o, o_o = oracle.segment_cloud(pts, sp)
            v_g, o_g = handle.segment_cloud(pts, sp)
            assert np.array_equal(v_g.view(np.uint32), v_o.view(np.uint32)), f"valid segments differ (mode {mode})"
            assert np.array_equal(o_g.view(np.uint32), o_o.view(np.uint32)), f"outliers differ (mode {mode})"
    sp = default_segment_params()
    sp.n_scan, sp.horizon_scan, sp.ang_res_x, sp.ang_res_y, sp.ang_bottom = 16, 1800, 0.2, 2.0, 15.1     # "VLP-16", imageProjection.hpp:95-102
    v_o, o_o = oracle.segment_cloud(clouds[1], sp)
    v_g, o_g = handle.segment_cloud(clouds[1], sp)
    assert np.array_equal(v_g.view(np.uint32), v_o.view(np.uint32)) and np.array_equal(o_g.view(np.uint32), o_o.view(np.uint32))
    v_g, o_g = handle.segment_cloud(np.zeros((0, 4), np.float32), default_segment_params())
    assert len(v_g) == 0 and len(o_g) == 0
    # the example's pre-processing chain (run_global_registration.cpp:136-162) in front of the registration path
    sp = default_segment_params()
    src, tgt, T = synth.outdoor_pair(33)
    chain_g = lambda c: handle.segment_cloud(handle.patchwork(c, pp)[1], sp)[0]
    chain_o = lambda c: oracle.segment_cloud(oracle.patchwork(c, pp)[1], sp)[0]
    p = default_params()
    p.skip_flagged = 0
    res, st = handle.register_pair(chain_g(src), chain_g(tgt), p)
    ref, st_ref = oracle.register_pair(chain_o(src), chain_o(tgt), p)
    assert st == st_ref and res.n_src_vox == ref.n_src_vox and res.n_corr == ref.n_corr and res.clique_size == ref.clique_size
    assert np.allclose(res.matrix(), ref.matrix(), atol=1e-9)


# ---- independent float64 restatement of Patchwork (numpy: two-pass covariance, LAPACK SVD) ---------------------------------
def _numpy_patchwork(pts, pp):
    """patchwork.hpp:329-455 in float64 numpy, written from the reference text: returns boolean masks (ground, nonground) over pts."""
    P = pts[:, :3].astype(np.float64)
    n = len(P)
    ok = np.isfinite(P).all(1) & ~(P[:, 2] < -1.8 * pp.sensor_height)
    r = np.hypot(P[:, 0], P[:, 1])
    th = np.arctan2(P[:, 1], P[:, 0]); th = np.where(th > 0, th, th + 2 * np.pi)
    ok &= (r <= pp.max_range) & (r > pp.min_range)
    mr = list(pp.min_ranges_each_zone) + [pp.max_range]
    zone = np.clip(np.searchsorted(np.array(mr[1:4]), r, side="right"), 0, 3)
    ground, nonground = np.zeros(n, bool), np.zeros(n, bool)
    concentric = 0
    for k in range(4):
        nr, ns = pp.num_rings_each_zone[k], pp.num_sectors_each_zone[k]
        ring = np.minimum(((r - mr[k]) / ((mr[k + 1] - mr[k]) / nr)).astype(int), nr - 1)
        sec = np.minimum((th / (2 * np.pi / ns)).astype(int), ns - 1)
        for ri in range(nr):
            for si in range(ns):
                idx = np.nonzero(ok & (zone == k) & (ring == ri) & (sec == si))[0]
                if len(idx) <= pp.num_min_pts:
                    continue
                idx = idx[np.argsort(P[idx, 2], kind="stable")]
                z = P[idx, 2]
                init = int((z < pp.adaptive_seed_selection_margin * pp.sensor_height).sum()) if k == 0 else 0
                lpr = z[init:init + pp.num_lpr].mean() if len(z[init:init + pp.num_lpr]) else 0.0
                g = z < lpr + pp.th_seeds
                for _ in range(pp.num_iter):
                    Q = P[idx][g]
                    mean = Q.mean(0)
                    cov = (Q - mean).T @ (Q - mean) / len(Q)
                    U, S, _ = np.linalg.svd(cov)
                    nrm = U[:, 2] if U[2, 2] >= 0 else -U[:, 2]
                    g = P[idx] @ nrm < pp.th_dist + nrm @ mean
                keep = abs(nrm[2]) >= pp.uprightness_thr
                if keep and concentric + ri < pp.num_thresholds:
                    if mean[2] > pp.elevation_thresholds[ri + 2 * k]:
                        keep = pp.flatness_thresholds[ri + 2 * k] > S[2] / S.sum()
                elif keep and pp.using_global_elevation and mean[2] > pp.global_elevation_threshold:
                    keep = False
                if keep:
                    ground[idx[g]] = True; nonground[idx[~g]] = True
                else:
                    nonground[idx] = True
        concentric += nr
    return ground, nonground


def test_patchwork_against_float64_numpy(oracle):
    """The oracle's canonical choices (float single-pass sums in a fixed tree, closed-form eigen solve, n_z >= 0) against a float64
    two-pass covariance + LAPACK SVD restatement: the same points are kept, and the labels agree for >= 99.5 % of them (a point within
    float rounding of the th_dist plane may flip, and with it the next iteration's fit of its patch)."""
    pp = default_patchwork_params()
    for pts in (synth.outdoor_pair(41)[0], _scene(7)):
        pts = pts.copy()
        pts[:, 3] = np.arange(len(pts), dtype=np.float32)          # the 4th channel carries the point id through the outputs
        assert len(pts) < 2 ** 24
        g, ng, _ = oracle.patchwork(pts, pp)
        G, N = _numpy_patchwork(pts, pp)
        got_g, got_n = np.zeros(len(pts), bool), np.zeros(len(pts), bool)
        got_g[g[:, 3].astype(int)] = True; got_n[ng[:, 3].astype(int)] = True
        assert np.array_equal(got_g | got_n, G | N)                 # the same points survive the range / height / patch-size rules
        agree = ((got_g == G) & (got_n == N))[G | N].mean()
        assert agree >= 0.995, agree
