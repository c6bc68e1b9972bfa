"""Independent float64 numpy / scipy restatements of the third-party stages on the hot path (PCL VoxelGrid, NormalEstimation,
computePairFeatures, FPFHEstimation; FLANN exact 1-NN as brute force; the fp64 TIM mask of quatro.hpp:363-385), written from the
published algorithms.  This module never imports oracle/ or the CUDA library: tools/gen_independent_pins.py uses it to write the
independent_*.npz fixtures and tests/test_independent_pins.py uses it to check the oracle."""
import numpy as np
from scipy.spatial import cKDTree


def voxel_grid(pts: np.ndarray, leaf: float):
    """pcl::VoxelGrid::applyFilter: ijk = floor(p / leaf) - floor(min / leaf); linear index i + j*dx + k*dx*dy; centroid per
    occupied voxel, output in ascending linear index.  float32 division like PCL (inverse_leaf_size multiplications), centroids
    in float64."""
    p32 = pts.astype(np.float32)
    inv = np.float32(1.0) / np.float32(leaf)
    mn = np.floor(p32.min(0) * inv).astype(np.int64)
    mx = np.floor(p32.max(0) * inv).astype(np.int64)
    div = mx - mn + 1
    ijk = np.floor(p32 * inv).astype(np.int64) - mn
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    uniq, start, counts = np.unique(idx[order], return_index=True, return_counts=True)
    sums = np.add.reduceat(pts[order].astype(np.float64), start, axis=0)
    return (sums / counts[:, None]), idx, uniq


def normals_pcl(pts: np.ndarray, radius: float):
    """pcl::NormalEstimation: neighbours with d < radius (self included), < 3 -> NaN; covariance about the centroid;
    eigenvector of the smallest eigenvalue; flipped towards the viewpoint (0,0,0); curvature = l0 / (l0+l1+l2)."""
    tree = cKDTree(pts)
    out = np.full((len(pts), 4), np.nan)
    gap = np.zeros(len(pts))   # (l1 - l0) / l2: how well the smallest eigenvector is defined
    nbrs = tree.query_ball_point(pts, radius * (1 + 1e-9))
    for i, nb in enumerate(nbrs):
        nb = [j for j in nb if np.sum((pts[j] - pts[i]) ** 2) < radius * radius]
        if len(nb) < 3:
            continue
        q = pts[nb]
        c = q.mean(0)
        cov = (q - c).T @ (q - c) / len(nb)
        w, v = np.linalg.eigh(cov)
        n = v[:, 0]
        if np.dot(n, -pts[i]) < 0:
            n = -n
        out[i, :3] = n
        s = w.sum()
        out[i, 3] = abs(w[0] / s) if s != 0 else 0.0
        gap[i] = (w[1] - w[0]) / w[2] if w[2] > 0 else 0.0
    return out, gap


def pair_features(p1, n1, p2, n2):
    """pcl::computePairFeatures"""
    d = p2 - p1
    f4 = np.linalg.norm(d)
    if f4 == 0.0:
        return None
    a1 = np.dot(n1, d) / f4
    a2 = np.dot(n2, d) / f4
    # C++ semantics: a comparison with NaN is false (no swap when the other normal is NaN); acos argument clipped against rounding
    def acos_abs(x):
        return np.nan if np.isnan(x) else np.arccos(min(1.0, abs(x)))
    if acos_abs(a1) > acos_abs(a2):
        n1, n2 = n2, n1
        d = -d
        f3 = -a2
    else:
        f3 = a1
    v = np.cross(d, n1)
    vn = np.linalg.norm(v)
    if vn == 0.0:
        return None
    v = v / vn
    w = np.cross(n1, v)
    f2 = np.dot(v, n2)
    f1 = np.arctan2(np.dot(w, n2), np.dot(n1, n2))
    return f1, f2, f3


def fpfh_pcl(pts: np.ndarray, normals: np.ndarray, radius: float):
    """pcl::FPFHEstimation: SPFH of every point over its radius neighbours (bins of 11, increment 100/(k-1)), then the
    1/d^2-weighted sum of the neighbours' SPFHs, each third rescaled to sum 100."""
    tree = cKDTree(pts)
    n = len(pts)
    nbrs = []
    for i, nb in enumerate(tree.query_ball_point(pts, radius * (1 + 1e-9))):
        nb = sorted(j for j in nb if np.sum((pts[j] - pts[i]) ** 2) < radius * radius)
        nbrs.append(nb)
    spfh = np.zeros((n, 33))
    for i, nb in enumerate(nbrs):
        k = len(nb)
        if k < 2:
            continue
        inc = 100.0 / (k - 1)
        for j in nb:
            if j == i:
                continue
            f = pair_features(pts[i], normals[i, :3], pts[j], normals[j, :3])
            if f is None:
                continue
            f1, f2, f3 = f
            # PCL casts the bin to int: NaN becomes INT_MIN on x86 and is clamped to bin 0 -- per feature
            b = tuple(0 if np.isnan(x) else int(np.floor(x)) for x in (11 * (f1 + np.pi) / (2 * np.pi), 11 * (f2 + 1.0) * 0.5, 11 * (f3 + 1.0) * 0.5))
            b = [min(10, max(0, x)) for x in b]
            spfh[i, b[0]] += inc
            spfh[i, 11 + b[1]] += inc
            spfh[i, 22 + b[2]] += inc
    out = np.zeros((n, 33))
    for i, nb in enumerate(nbrs):
        acc = np.zeros(33)
        for j in nb:
            d2 = np.sum((pts[j] - pts[i]) ** 2)
            if d2 == 0.0:
                continue
            acc += spfh[j] / d2
        for t in range(3):
            s = acc[11 * t:11 * t + 11].sum()
            if s != 0:
                acc[11 * t:11 * t + 11] *= 100.0 / s
        out[i] = acc
    return out, spfh


def mutual_nn(da: np.ndarray, db: np.ndarray):
    """feature_matcher.cc:79-180 in its dense form: row/column argmins of the squared-distance matrix (float64, lowest index on
    ties), mutual pairs in ascending source index."""
    D = np.empty((len(da), len(db)))
    for i0 in range(0, len(da), 256):   # direct differences (exact ties stay exact), in row chunks
        blk = da[i0:i0 + 256, None, :] - db[None, :, :]
        D[i0:i0 + 256] = (blk * blk).sum(2)
    r = D.argmin(1)
    c = D.argmin(0)
    i = np.arange(len(da))
    keep = c[r] == i
    # second-best margins: pairs whose decision float32 arithmetic could flip
    Ds = np.sort(D, 1)
    margin_r = Ds[:, 1] - Ds[:, 0]
    Dc = np.sort(D, 0)
    margin_c = Dc[1] - Dc[0]
    return np.stack([i[keep], r[keep]], 1), margin_r, margin_c


def tim_graph(a: np.ndarray, b: np.ndarray, beta: float):
    """quatro.hpp:363-385 in float64 numpy: edge <=> |db/da - 1| <= beta/da and |da/db - 1| <= beta/db."""
    da = np.linalg.norm(a[:, None, :] - a[None, :, :], axis=2)
    db = np.linalg.norm(b[:, None, :] - b[None, :, :], axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        e = (np.abs(db / da - 1) <= beta / da) & (np.abs(da / db - 1) <= beta / db)
    np.fill_diagonal(e, False)
    knife = np.abs(np.abs(da - db) - beta) < 1e-9
    return e, knife
