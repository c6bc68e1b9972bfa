"""ctypes binding of include/quatro_b200.h (one Python method per C entry point).

Used by tests/ and bench.py; mirrors the C-ABI 1:1 so the parity tests read like calls a C/C++
caller (the reference's run_global_registration.cpp) would make.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _build

PMC_EXACT, PMC_HEU, KCORE_HEU, INLIER_NONE = 0, 1, 2, 3
FLAG_CLIQUE_TRUNCATED = 1
COTE_MEDIAN, COTE_WEIGHTED_MEAN = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

STATUS_NAMES = {0: "OK", 1: "DEGENERATE_CLIQUE", 2: "DEGENERATE_INPUT", 3: "CAPACITY_EXCEEDED", -1: "ERR_BAD_ARG",
                -2: "ERR_NO_DEVICE", -3: "ERR_CUDA", -4: "ERR_UNSUPPORTED", -5: "ERR_VOXEL_OVERFLOW"}


class Params(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_float), ("normal_radius", C.c_float), ("fpfh_radius", C.c_float), ("grid_cell", C.c_float),
        ("tuple_scale", C.c_float), ("use_crosscheck", C.c_int32), ("use_tuple_test", C.c_int32),
        ("tuple_trials_per_corr", C.c_int32), ("skip_flagged", C.c_int32), ("reserved0", C.c_int32),
        ("seed", C.c_uint64),
        ("noise_bound", C.c_double), ("cbar2", C.c_double), ("rot_noise_bound", C.c_double),
        ("cote_noise_bound", C.c_double), ("rotation_gnc_factor", C.c_double),
        ("rotation_cost_threshold", C.c_double), ("kcore_heuristic_threshold", C.c_double),
        ("rotation_max_iterations", C.c_int32), ("inlier_selection_mode", C.c_int32), ("cote_mode", C.c_int32),
        ("using_rot_inliers_when_estimating_cote", C.c_int32), ("use_pre_estimated_RyRx", C.c_int32),
        ("max_clique_node_limit", C.c_int32), ("RyRx", C.c_double * 9),
    ]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_batch_slots", C.c_int32), ("max_raw_points", C.c_int32),
                ("max_voxel_points", C.c_int32), ("max_corr", C.c_int32), ("reserved", C.c_int32 * 3)]


class Result(C.Structure):
    _fields_ = [
        ("valid", C.c_int32), ("status", C.c_int32), ("n_src_vox", C.c_int32), ("n_tgt_vox", C.c_int32),
        ("n_mutual", C.c_int32), ("n_corr", C.c_int32), ("max_core", C.c_int32), ("clique_size", C.c_int32),
        ("gnc_iters", C.c_int32), ("n_rot_inliers", C.c_int32), ("n_final_inliers", C.c_int32),
        ("flags", C.c_int32), ("n_edges", C.c_int64), ("cost", C.c_double), ("T", C.c_double * 16),
    ]

    def matrix(self) -> np.ndarray:
        """4x4 pose (row/col indexing as usual); T is stored column-major."""
        return np.array(self.T[:], dtype=np.float64).reshape(4, 4).T.copy()

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "T"}
        d["T"] = self.matrix()
        return d


class PatchworkParams(C.Structure):
    """qb200_patchwork_params (config/patchwork_params.yaml of the reference)."""
    _fields_ = [
        ("sensor_height", C.c_double), ("th_seeds", C.c_double), ("th_dist", C.c_double), ("max_range", C.c_double),
        ("min_range", C.c_double), ("uprightness_thr", C.c_double), ("adaptive_seed_selection_margin", C.c_double),
        ("global_elevation_threshold", C.c_double), ("min_ranges_each_zone", C.c_double * 4),
        ("elevation_thresholds", C.c_double * 8), ("flatness_thresholds", C.c_double * 8),
        ("num_iter", C.c_int32), ("num_lpr", C.c_int32), ("num_min_pts", C.c_int32), ("using_global_elevation", C.c_int32),
        ("num_zones", C.c_int32), ("num_thresholds", C.c_int32), ("num_sectors_each_zone", C.c_int32 * 4),
        ("num_rings_each_zone", C.c_int32 * 4),
    ]


class SegmentParams(C.Structure):
    """qb200_segment_params (per-sensor constants of the ImageProjection constructor)."""
    _fields_ = [("n_scan", C.c_int32), ("horizon_scan", C.c_int32), ("ang_res_x", C.c_float), ("ang_res_y", C.c_float),
                ("ang_bottom", C.c_float), ("segment_theta", C.c_float), ("neighbor_mode", C.c_int32),
                ("min_pts_for_subclustering", C.c_int32), ("segment_valid_point_num", C.c_int32), ("segment_valid_line_num", C.c_int32)]


class Pair(C.Structure):
    _fields_ = [("src", C.c_void_p), ("tgt", C.c_void_p), ("n_src", C.c_int32), ("n_tgt", C.c_int32)]


class CorrSet(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("L", C.c_int32), ("reserved", C.c_int32)]


RESULT_DTYPE = np.dtype([
    ("valid", "<i4"), ("status", "<i4"), ("n_src_vox", "<i4"), ("n_tgt_vox", "<i4"), ("n_mutual", "<i4"),
    ("n_corr", "<i4"), ("max_core", "<i4"), ("clique_size", "<i4"), ("gnc_iters", "<i4"), ("n_rot_inliers", "<i4"),
    ("n_final_inliers", "<i4"), ("flags", "<i4"), ("n_edges", "<i8"), ("cost", "<f8"), ("T", "<f8", (16,)),
])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class QuatroB200Error(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: {STATUS_NAMES.get(code, code)} {detail}")


_LIB: Optional[C.CDLL] = None


def _f32(a, cols=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        assert a.ndim == 2 and a.shape[1] == cols, (a.shape, cols)
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def load_library(build: bool = True) -> C.CDLL:
    """Load libquatro_b200.so (building it with nvcc if stale).  Raises if it cannot be built/loaded."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.build_cuda() if build else _build.CUDA_LIB
    lib = C.CDLL(str(path))
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    P = C.POINTER
    sig = {
        "qb200_default_params": (None, [P(Params)]),
        "qb200_default_config": (None, [P(Config)]),
        "qb200_version": (i32, []),
        "qb200_create": (i32, [P(Config), P(vp)]),
        "qb200_destroy": (None, [vp]),
        "qb200_set_stream": (i32, [vp, vp]),
        "qb200_last_error": (C.c_char_p, [vp]),
        "qb200_launch_count": (i64, [vp]),
        "qb200_voxelize": (i32, [vp, vp, i32, f32, i32, vp, i32, P(i32)]),
        "qb200_default_patchwork_params": (None, [P(PatchworkParams)]),
        "qb200_default_segment_params": (None, [P(SegmentParams)]),
        "qb200_segment_cloud": (i32, [vp, vp, i32, P(SegmentParams), vp, P(i32), vp, P(i32)]),
        "qb200_patchwork": (i32, [vp, vp, i32, P(PatchworkParams), vp, P(i32), vp, P(i32)]),
        "qb200_compute_fpfh": (i32, [vp, vp, i32, f32, f32, f32, vp, vp]),
        "qb200_match": (i32, [vp, vp, i32, vp, vp, i32, vp, P(Params), vp, i32, P(i32), P(i32)]),
        "qb200_build_graph": (i32, [vp, vp, vp, i32, f64, f64, vp, i32, vp, P(i64)]),
        "qb200_max_clique": (i32, [vp, vp, i32, i32, i32, f64, vp, P(i32), vp, vp, P(i32)]),
        "qb200_max_clique_ex": (i32, [vp, vp, i32, i32, i32, f64, i64, vp, P(i32), vp, vp, P(i32), P(i32)]),
        "qb200_solve_pose": (i32, [vp, vp, vp, i32, vp, i32, P(Params), P(Result), vp, vp]),
        "qb200_solve_correspondences": (i32, [vp, vp, vp, i32, P(Params), P(Result)]),
        "qb200_match_and_pack": (i32, [vp, vp, i32, vp, i32, P(Params), vp, vp, vp, i32, P(i32)]),
        "qb200_register_pair": (i32, [vp, vp, i32, vp, i32, P(Params), P(Result)]),
        "qb200_register_batch": (i32, [vp, P(Pair), i32, P(Params), i32, vp]),
        "qb200_get_last_clique": (i32, [vp, vp, i32, P(i32)]),
        "qb200_get_last_final_inliers": (i32, [vp, vp, i32, P(i32)]),
        "qb200_get_last_correspondences": (i32, [vp, vp, vp, vp, i32, P(i32)]),
        "qb200_get_stage_ms": (i32, [vp, vp, i32]),
        "qb200_get_kernel_ms": (i32, [vp, vp, vp, i32]),
        "qb200_debug_tc_distances": (i32, [vp, vp, i32, vp, i32, vp]),
        "qb200_debug_match_stats": (i32, [vp, vp, i32]),
        "qb200_register_batch_enqueue": (i32, [vp, vp, i32, vp, i32, vp]),
        "qb200_register_batch_flush": (i32, [vp]),
        "qb200_debug_tc_profile": (i32, [vp, vp, i32]),
        "qb200_solve_batch": (i32, [vp, vp, i32, vp, i32, vp]),
        "qb200_comm_init_all": (i32, [P(vp), i32]),
        "qb200_register_batch_sharded": (i32, [P(vp), i32, P(Pair), i32, P(Params), i32, vp]),
        "qb200_comm_unique_id": (i32, [vp]),
        "qb200_comm_init_rank": (i32, [vp, i32, i32, vp]),
        "qb200_register_batch_rank": (i32, [vp, P(Pair), i32, P(Params), i32, vp, i32]),
        "qb200_comm_wait": (i32, [vp]),
        "qb200_bind_numa": (i32, [vp]),
        "qb200_debug_match_verify": (i32, [vp, vp, i32]),
        "qb200_get_last_features": (i32, [vp, i32, vp, vp, i32, P(i32)]),
        "qb200_cache_reserve": (i32, [vp, i32]),
        "qb200_cache_scans": (i32, [vp, P(vp), P(i32), P(i32), i32, P(Params), i32]),
        "qb200_register_cached": (i32, [vp, vp, i32, P(Params), vp]),
        "qb200_cache_copy": (i32, [vp, i32, i32]),
        "qb200_cache_read": (i32, [vp, i32, vp, vp, vp, i32, P(i32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "qb200_default_params", "qb200_default_config", "qb200_version", "qb200_create", "qb200_destroy",
    "qb200_set_stream", "qb200_last_error", "qb200_launch_count", "qb200_voxelize", "qb200_compute_fpfh",
    "qb200_default_patchwork_params", "qb200_patchwork", "qb200_default_segment_params", "qb200_segment_cloud", "qb200_match", "qb200_build_graph", "qb200_max_clique", "qb200_max_clique_ex", "qb200_solve_pose", "qb200_solve_correspondences",
    "qb200_match_and_pack", "qb200_register_pair", "qb200_register_batch", "qb200_get_last_clique",
    "qb200_get_last_final_inliers", "qb200_get_last_correspondences", "qb200_get_stage_ms", "qb200_get_kernel_ms",
    "qb200_debug_tc_distances",
    "qb200_debug_match_stats",
    "qb200_register_batch_enqueue",
    "qb200_register_batch_flush",
    "qb200_debug_tc_profile",
    "qb200_solve_batch",
    "qb200_comm_init_all", "qb200_register_batch_sharded", "qb200_comm_unique_id", "qb200_comm_init_rank",
    "qb200_register_batch_rank", "qb200_comm_wait", "qb200_bind_numa",
    "qb200_debug_match_verify", "qb200_get_last_features", "qb200_cache_reserve", "qb200_cache_scans", "qb200_register_cached", "qb200_cache_copy", "qb200_cache_read",
]


def default_params() -> Params:
    """config/params.yaml defaults.  Pure-Python mirror of qb200_default_params (so CPU-only tests can
    build a Params without loading the CUDA library); test_capi checks the two agree."""
    p = Params()
    p.voxel_size, p.normal_radius, p.fpfh_radius, p.grid_cell = 0.3, 0.5, 0.75, 0.0
    p.tuple_scale, p.use_crosscheck, p.use_tuple_test, p.tuple_trials_per_corr = 0.95, 1, 1, 100
    p.skip_flagged, p.seed = 1, 0x5EED
    p.noise_bound, p.cbar2, p.rot_noise_bound, p.cote_noise_bound = 0.3, 1.0, 0.0, 0.3
    p.rotation_gnc_factor, p.rotation_cost_threshold, p.kcore_heuristic_threshold = 1.4, 0.00011, 0.5
    p.rotation_max_iterations, p.inlier_selection_mode, p.cote_mode = 50, PMC_HEU, COTE_MEDIAN
    p.using_rot_inliers_when_estimating_cote, p.use_pre_estimated_RyRx = 0, 0
    for i in range(9):
        p.RyRx[i] = 1.0 if i in (0, 4, 8) else 0.0
    return p


def comm_init_all(handles: Sequence["Handle"]):
    """(A) one process, several devices: ncclCommInitAll over the handles' devices."""
    arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
    st = load_library().qb200_comm_init_all(arr, len(handles))
    if st != 0:
        raise QuatroB200Error(st, "qb200_comm_init_all")


def register_batch_sharded(handles: Sequence["Handle"], pairs: Sequence, params: "Params", kind: int = 0) -> np.ndarray:
    """pairs: (src, tgt) numpy arrays (MEM_HOST) or (src_ptr, n_src, tgt_ptr, n_tgt) device tuples whose pair g lives on the device
    of handles[g % len(handles)].  Returns the records in the order of `pairs`."""
    n = len(pairs)
    arr = (Pair * n)()
    keep = []
    for i, pr in enumerate(pairs):
        if kind == MEM_HOST:
            s, t = _f32(pr[0], 4), _f32(pr[1], 4)
            keep.append((s, t))
            arr[i].src, arr[i].n_src, arr[i].tgt, arr[i].n_tgt = s.ctypes.data, len(s), t.ctypes.data, len(t)
        else:
            arr[i].src, arr[i].n_src, arr[i].tgt, arr[i].n_tgt = pr[0], pr[1], pr[2], pr[3]
    out = np.zeros(n, RESULT_DTYPE)
    hs = (C.c_void_p * len(handles))(*[h.h for h in handles])
    st = load_library().qb200_register_batch_sharded(hs, len(handles), arr, n, C.byref(params), kind, _ptr(out))
    if st != 0:
        raise QuatroB200Error(st, "qb200_register_batch_sharded " + handles[0].last_error())
    return out


def default_patchwork_params() -> PatchworkParams:
    p = PatchworkParams()
    load_library().qb200_default_patchwork_params(C.byref(p))
    return p


def default_segment_params() -> SegmentParams:
    p = SegmentParams()
    load_library().qb200_default_segment_params(C.byref(p))
    return p


def default_config() -> Config:
    c = Config()
    c.device, c.max_batch_slots, c.max_raw_points, c.max_voxel_points, c.max_corr = 0, 64, 131072, 16384, 4096
    return c


class Handle:
    """RAII wrapper of qb200_handle.  Raises QuatroB200Error(ERR_NO_DEVICE) when no GPU is usable."""

    def __init__(self, config: Optional[Config] = None, **kw):
        self.lib = load_library()
        cfg = config or default_config()
        for k, v in kw.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        st = self.lib.qb200_create(C.byref(cfg), C.byref(h))
        if st != 0:
            raise QuatroB200Error(st, "qb200_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.qb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st: int, where: str):
        if st < 0:
            msg = self.lib.qb200_last_error(self.h)
            raise QuatroB200Error(st, where, (msg or b"").decode())
        return st

    def set_stream(self, stream_ptr: int):
        self._check(self.lib.qb200_set_stream(self.h, C.c_void_p(stream_ptr)), "qb200_set_stream")

    def launch_count(self) -> int:
        return int(self.lib.qb200_launch_count(self.h))

    # ---- stages -------------------------------------------------------------------------------
    def voxelize(self, pts4, leaf: float, skip_flagged: int = 1, cap: Optional[int] = None):
        pts4 = _f32(pts4, 4)
        cap = cap or max(1, len(pts4))
        out = np.zeros((cap, 4), np.float32)
        n = C.c_int32(0)
        st = self.lib.qb200_voxelize(self.h, _ptr(pts4), len(pts4), leaf, skip_flagged, _ptr(out), cap, C.byref(n))
        if st != -5:  # ERR_VOXEL_OVERFLOW = PCL's "leaf too small" pass-through: output is the unfiltered input
            self._check(st, "qb200_voxelize")
        return out[: min(n.value, cap)].copy(), st

    def compute_fpfh(self, pts4, normal_radius: float, fpfh_radius: float, grid_cell: float):
        pts4 = _f32(pts4, 4)
        n = len(pts4)
        normals = np.zeros((n, 4), np.float32)
        desc = np.zeros((n, 33), np.float32)
        self._check(self.lib.qb200_compute_fpfh(self.h, _ptr(pts4), n, normal_radius, fpfh_radius, grid_cell, _ptr(normals), _ptr(desc)),
                    "qb200_compute_fpfh")
        return normals, desc

    def match(self, src4, sdesc, tgt4, tdesc, params: Params, cap: Optional[int] = None):
        src4, tgt4, sdesc, tdesc = _f32(src4, 4), _f32(tgt4, 4), _f32(sdesc, 33), _f32(tdesc, 33)
        cap = cap or max(1, min(len(src4), len(tgt4)))
        corr = np.zeros((cap, 2), np.int32)
        n, nm = C.c_int32(0), C.c_int32(0)
        st = self._check(self.lib.qb200_match(self.h, _ptr(src4), len(src4), _ptr(sdesc), _ptr(tgt4), len(tgt4), _ptr(tdesc),
                                              C.byref(params), _ptr(corr), cap, C.byref(n), C.byref(nm)), "qb200_match")
        return corr[: min(n.value, cap)].copy(), nm.value, st

    def build_graph(self, a4, b4, noise_bound: float, cbar2: float, words_per_row: Optional[int] = None):
        a4, b4 = _f32(a4, 4), _f32(b4, 4)
        L = len(a4)
        wpr = words_per_row or (L + 31) // 32
        adj = np.zeros((L, wpr), np.uint32)
        deg = np.zeros(L, np.int32)
        ne = C.c_int64(0)
        self._check(self.lib.qb200_build_graph(self.h, _ptr(a4), _ptr(b4), L, noise_bound, cbar2, _ptr(adj), wpr, _ptr(deg), C.byref(ne)),
                    "qb200_build_graph")
        return adj, deg, ne.value

    def patchwork(self, pts, pp: "PatchworkParams"):
        """qb200_patchwork: (ground (g,4), nonground (m,4), status)."""
        pts = _f32(pts, 4)
        n = len(pts)
        g, ng = np.zeros((max(n, 1), 4), np.float32), np.zeros((max(n, 1), 4), np.float32)
        a, b = C.c_int32(0), C.c_int32(0)
        st = self._check(self.lib.qb200_patchwork(self.h, _ptr(pts), n, C.byref(pp), _ptr(g), C.byref(a), _ptr(ng), C.byref(b)), "qb200_patchwork")
        return g[: a.value].copy(), ng[: b.value].copy(), st

    def segment_cloud(self, pts, sp: "SegmentParams"):
        """qb200_segment_cloud: (valid segments (v,4), outliers (o,4))."""
        pts = _f32(pts, 4)
        npix = sp.n_scan * sp.horizon_scan
        v, o = np.zeros((npix, 4), np.float32), np.zeros((npix, 4), np.float32)
        a, b = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.qb200_segment_cloud(self.h, _ptr(pts), len(pts), C.byref(sp), _ptr(v), C.byref(a), _ptr(o), C.byref(b)),
                    "qb200_segment_cloud")
        return v[: a.value].copy(), o[: b.value].copy()

    def max_clique(self, adj, mode: int = PMC_HEU, kcore_thr: float = 0.5):
        adj = np.ascontiguousarray(adj, np.uint32)
        L, wpr = adj.shape
        clique = np.zeros(max(L, 1), np.int32)
        kcore = np.zeros(max(L, 1), np.int32)
        order = np.zeros(max(L, 1), np.int32)
        n, mc = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.qb200_max_clique(self.h, _ptr(adj), L, wpr, mode, kcore_thr, _ptr(clique), C.byref(n), _ptr(kcore),
                                              _ptr(order), C.byref(mc)), "qb200_max_clique")
        return clique[: n.value].copy(), kcore[:L].copy(), order[:L].copy(), mc.value

    def max_clique_ex(self, adj, mode: int = PMC_EXACT, kcore_thr: float = 0.5, node_limit: int = 0):
        """qb200_max_clique_ex: clique, kcore, order, max_core, flags (FLAG_CLIQUE_TRUNCATED when the node limit stopped the search)."""
        adj = np.ascontiguousarray(adj, np.uint32)
        L, wpr = adj.shape
        clique, kcore, order = (np.zeros(max(L, 1), np.int32) for _ in range(3))
        n, mc, fl = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._check(self.lib.qb200_max_clique_ex(self.h, _ptr(adj), L, wpr, mode, kcore_thr, node_limit, _ptr(clique), C.byref(n),
                                                 _ptr(kcore), _ptr(order), C.byref(mc), C.byref(fl)), "qb200_max_clique_ex")
        return clique[: n.value].copy(), kcore[:L].copy(), order[:L].copy(), mc.value, fl.value

    def solve_pose(self, a4, b4, clique, params: Params):
        a4, b4 = _f32(a4, 4), _f32(b4, 4)
        clique = np.ascontiguousarray(clique, np.int32)
        res = Result()
        rm = np.zeros(max(len(clique), 1), np.uint8)
        tm = np.zeros(max(len(clique), 1), np.uint8)
        st = self._check(self.lib.qb200_solve_pose(self.h, _ptr(a4), _ptr(b4), len(a4), _ptr(clique), len(clique), C.byref(params),
                                                   C.byref(res), _ptr(rm), _ptr(tm)), "qb200_solve_pose")
        return res, rm[: len(clique)], tm[: len(clique)], st

    def solve_correspondences(self, a4, b4, params: Params):
        a4, b4 = _f32(a4, 4), _f32(b4, 4)
        res = Result()
        st = self._check(self.lib.qb200_solve_correspondences(self.h, _ptr(a4), _ptr(b4), len(a4), C.byref(params), C.byref(res)),
                         "qb200_solve_correspondences")
        return res, st

    def match_and_pack(self, src4, tgt4, params: Params, cap: Optional[int] = None):
        src4, tgt4 = _f32(src4, 4), _f32(tgt4, 4)
        cap = cap or max(1, min(len(src4), len(tgt4)))
        corr = np.zeros((cap, 2), np.int32)
        sm = np.zeros((cap, 4), np.float32)
        tm = np.zeros((cap, 4), np.float32)
        n = C.c_int32(0)
        st = self._check(self.lib.qb200_match_and_pack(self.h, _ptr(src4), len(src4), _ptr(tgt4), len(tgt4), C.byref(params), _ptr(corr),
                                                       _ptr(sm), _ptr(tm), cap, C.byref(n)), "qb200_match_and_pack")
        m = min(n.value, cap)
        return corr[:m].copy(), sm[:m].copy(), tm[:m].copy(), st

    def register_pair(self, src4, tgt4, params: Params):
        src4, tgt4 = _f32(src4, 4), _f32(tgt4, 4)
        res = Result()
        st = self._check(self.lib.qb200_register_pair(self.h, _ptr(src4), len(src4), _ptr(tgt4), len(tgt4), C.byref(params), C.byref(res)),
                         "qb200_register_pair")
        return res, st

    def register_batch(self, pairs: Sequence, params: Params, kind: int = MEM_HOST) -> np.ndarray:
        """pairs: sequence of (src, tgt).  MEM_HOST: numpy (n,4) float32 arrays; MEM_DEVICE:
        (src_ptr, n_src, tgt_ptr, n_tgt) tuples of raw device addresses.  Returns a RESULT_DTYPE array."""
        n = len(pairs)
        arr = (Pair * n)()
        keep = []
        for i, pr in enumerate(pairs):
            if kind == MEM_HOST:
                s, t = _f32(pr[0], 4), _f32(pr[1], 4)
                keep.append((s, t))
                arr[i].src, arr[i].n_src, arr[i].tgt, arr[i].n_tgt = s.ctypes.data, len(s), t.ctypes.data, len(t)
            else:
                arr[i].src, arr[i].n_src, arr[i].tgt, arr[i].n_tgt = pr[0], pr[1], pr[2], pr[3]
        out = np.zeros(n, RESULT_DTYPE)
        self._check(self.lib.qb200_register_batch(self.h, arr, n, C.byref(params), kind, _ptr(out)), "qb200_register_batch")
        return out

    def solve_batch(self, sets: Sequence, params: Params, kind: int = MEM_HOST) -> np.ndarray:
        """sets: sequence of (a4, b4) matched point arrays (MEM_HOST: numpy (L,4) float32; MEM_DEVICE: (a_ptr, b_ptr, L)).
        Graph -> clique -> pose for every set; returns a RESULT_DTYPE array."""
        n = len(sets)
        arr = (CorrSet * n)()
        keep = []
        for i, st in enumerate(sets):
            if kind == MEM_HOST:
                a, b = _f32(st[0], 4), _f32(st[1], 4)
                assert len(a) == len(b)
                keep.append((a, b))
                arr[i].a, arr[i].b, arr[i].L = a.ctypes.data, b.ctypes.data, len(a)
            else:
                arr[i].a, arr[i].b, arr[i].L = st[0], st[1], st[2]
        out = np.zeros(n, RESULT_DTYPE)
        self._check(self.lib.qb200_solve_batch(self.h, arr, n, C.byref(params), kind, _ptr(out)), "qb200_solve_batch")
        return out

    def last_features(self, which: int, cap: Optional[int] = None):
        """(normals (n,4), descriptors (n,33)) of the source (0) / target (1) cloud of the last match_and_pack."""
        cap = cap or self.cfg.max_voxel_points
        nrm, desc = np.zeros((cap, 4), np.float32), np.zeros((cap, 33), np.float32)
        n = C.c_int32(0)
        self._check(self.lib.qb200_get_last_features(self.h, which, _ptr(nrm), _ptr(desc), cap, C.byref(n)), "qb200_get_last_features")
        m = min(n.value, cap)
        return nrm[:m].copy(), desc[:m].copy()

    # ---- scan cache ----
    def cache_reserve(self, n_slots: int):
        self._check(self.lib.qb200_cache_reserve(self.h, n_slots), "qb200_cache_reserve")

    def cache_scans(self, scans: Sequence, slot_ids: Sequence[int], params: Params, kind: int = MEM_HOST):
        """scans: (n,4) float32 arrays (MEM_HOST) or (device_ptr, n) tuples (MEM_DEVICE)."""
        n = len(scans)
        keep = [_f32(sc, 4) for sc in scans] if kind == MEM_HOST else None
        ptrs = (C.c_void_p * n)(*([a.ctypes.data for a in keep] if kind == MEM_HOST else [sc[0] for sc in scans]))
        cnts = (C.c_int32 * n)(*([len(a) for a in keep] if kind == MEM_HOST else [sc[1] for sc in scans]))
        ids = (C.c_int32 * n)(*[int(x) for x in slot_ids])
        self._check(self.lib.qb200_cache_scans(self.h, ptrs, cnts, ids, n, C.byref(params), kind), "qb200_cache_scans")

    def register_cached(self, slot_pairs, params: Params) -> np.ndarray:
        sp = np.ascontiguousarray(np.asarray(slot_pairs, np.int32).reshape(-1, 2))
        out = np.zeros(len(sp), RESULT_DTYPE)
        self._check(self.lib.qb200_register_cached(self.h, _ptr(sp), len(sp), C.byref(params), _ptr(out)), "qb200_register_cached")
        return out

    def cache_copy(self, from_slot: int, to_slot: int):
        self._check(self.lib.qb200_cache_copy(self.h, from_slot, to_slot), "qb200_cache_copy")

    def cache_read(self, slot: int, cap: Optional[int] = None):
        """-> (voxel points (n,4), normals (n,4), descriptors (n,33)) of a cached scan."""
        cap = cap or self.cfg.max_voxel_points
        vox, nrm, desc = np.zeros((cap, 4), np.float32), np.zeros((cap, 4), np.float32), np.zeros((cap, 33), np.float32)
        n = C.c_int32(0)
        self._check(self.lib.qb200_cache_read(self.h, slot, _ptr(vox), _ptr(nrm), _ptr(desc), cap, C.byref(n)), "qb200_cache_read")
        m = min(n.value, cap)
        return vox[:m].copy(), nrm[:m].copy(), desc[:m].copy()

    # ---- multi-GPU (comm.cu) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        st = load_library().qb200_comm_unique_id(buf)
        if st != 0:
            raise QuatroB200Error(st, "qb200_comm_unique_id")
        return buf.raw

    def comm_init_rank(self, world: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.qb200_comm_init_rank(self.h, world, rank, buf), "qb200_comm_init_rank")
        self.world, self.rank = world, rank

    def register_batch_rank_raw(self, pair_array, n_local: int, params: Params, kind: int, out_all: np.ndarray, defer: bool = False):
        """out_all: RESULT_DTYPE array of world * n_local records, filled in round-robin global order (index i * world + r)."""
        return self._check(self.lib.qb200_register_batch_rank(self.h, pair_array, n_local, C.byref(params), kind, _ptr(out_all), int(defer)),
                           "qb200_register_batch_rank")

    def bind_numa(self) -> int:
        return int(self.lib.qb200_bind_numa(self.h))

    def comm_wait(self):
        return self._check(self.lib.qb200_comm_wait(self.h), "qb200_comm_wait")

    def register_batch_raw(self, pair_array, n: int, params: Params, kind: int, out: np.ndarray):
        """Zero-overhead variant for bench.py: pre-built (Pair * n) array and RESULT_DTYPE output."""
        return self._check(self.lib.qb200_register_batch(self.h, pair_array, n, C.byref(params), kind, _ptr(out)), "qb200_register_batch")

    def register_batch_enqueue_raw(self, pair_array, n: int, params: Params, kind: int, out: np.ndarray):
        """Pipelined form: queue the batch (pair_array, its scans and `out` must stay alive until register_batch_flush)."""
        return self._check(self.lib.qb200_register_batch_enqueue(self.h, pair_array, n, C.byref(params), kind, _ptr(out)), "qb200_register_batch_enqueue")

    def register_batch_flush(self):
        return self._check(self.lib.qb200_register_batch_flush(self.h), "qb200_register_batch_flush")

    def last_clique(self, cap: int = 1 << 16):
        idx = np.zeros(cap, np.int32)
        n = C.c_int32(0)
        self._check(self.lib.qb200_get_last_clique(self.h, _ptr(idx), cap, C.byref(n)), "qb200_get_last_clique")
        return idx[: n.value].copy()

    def last_final_inliers(self, cap: int = 1 << 16):
        idx = np.zeros(cap, np.int32)
        n = C.c_int32(0)
        self._check(self.lib.qb200_get_last_final_inliers(self.h, _ptr(idx), cap, C.byref(n)), "qb200_get_last_final_inliers")
        return idx[: n.value].copy()

    def last_correspondences(self, cap: int = 1 << 16):
        corr = np.zeros((cap, 2), np.int32)
        sm = np.zeros((cap, 4), np.float32)
        tm = np.zeros((cap, 4), np.float32)
        n = C.c_int32(0)
        self._check(self.lib.qb200_get_last_correspondences(self.h, _ptr(corr), _ptr(sm), _ptr(tm), cap, C.byref(n)),
                    "qb200_get_last_correspondences")
        return corr[: n.value].copy(), sm[: n.value].copy(), tm[: n.value].copy()

    def debug_match_stats(self, reset: bool = True) -> dict:
        out = np.zeros(4, np.uint64)
        self._check(self.lib.qb200_debug_match_stats(self.h, _ptr(out), int(reset)), "qb200_debug_match_stats")
        return {"exact_evals": int(out[0]), "tiles": int(out[1]), "warmups": int(out[2]), "aborted_stripes": int(out[3])}

    def debug_tc_profile(self, reset: bool = True) -> np.ndarray:
        out = np.zeros(24, np.uint64)
        self._check(self.lib.qb200_debug_tc_profile(self.h, _ptr(out), int(reset)), "qb200_debug_tc_profile")
        return out

    def debug_match_verify(self, reset: bool = True) -> dict:
        out = np.zeros(2, np.uint64)
        self._check(self.lib.qb200_debug_match_verify(self.h, _ptr(out), int(reset)), "qb200_debug_match_verify")
        return {"compared": int(out[0]), "mismatches": int(out[1])}

    def debug_tc_distances(self, a33, b33) -> np.ndarray:
        a33, b33 = _f32(a33, 33), _f32(b33, 33)
        out = np.zeros((128, 128), np.float32)
        self._check(self.lib.qb200_debug_tc_distances(self.h, _ptr(a33), len(a33), _ptr(b33), len(b33), _ptr(out)), "qb200_debug_tc_distances")
        return out[: len(a33), : len(b33)].copy()

    def kernel_ms(self):
        """(ms, launches) of the two roofline kernels during the last register_batch:
        index 0 = match_stripe_kernel (K6), 1 = tim_graph_kernel (K8)."""
        ms = np.zeros(2, np.float32)
        calls = np.zeros(2, np.int32)
        self.lib.qb200_get_kernel_ms(self.h, _ptr(ms), _ptr(calls), 2)
        return ms, calls

    def stage_ms(self) -> np.ndarray:
        ms = np.zeros(8, np.float32)
        self.lib.qb200_get_stage_ms(self.h, _ptr(ms), 8)
        return ms
