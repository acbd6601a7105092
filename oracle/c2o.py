"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/liboracle.so (CPU restatement of the reference's cont2contops path) and, when present,
oracle/_ref/libref_knn.so (the reference's own vendored nanoflann).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from contour_context_b200 import ctypes_defs as D

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("c2o_capi.cpp", "c2o_ingest.hpp", "c2o_query.hpp", "c2o_refine.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "c2g_types.h"))
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(_HERE, "_ref", "liboracle_nf.so")
    ref_stale = force or not os.path.exists(ref_so) or any(os.path.getmtime(s) > os.path.getmtime(ref_so) for s in srcs)
    if os.path.exists("/root/reference/thirdparty/nanoflann.hpp") and ref_stale:
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def use_nanoflann_build() -> bool:
    """Switch to oracle/_ref/liboracle_nf.so (the restatement linked against the reference's own nanoflann KD-tree).
    Must be called before the first lib() call; returns False if that build is not available."""
    global _LIB_PATH
    build()
    p = os.path.join(_HERE, "_ref", "liboracle_nf.so")
    if _LIB is None and os.path.exists(p):
        _LIB_PATH = p
        return True
    return False


_LIB_PATH = None


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(_LIB_PATH or os.path.join(_HERE, "liboracle.so"))
        L.c2o_scan_create.restype = C.c_void_p
        L.c2o_scan_create.argtypes = [C.POINTER(D.CmConfig), C.c_int]
        L.c2o_scan_free.argtypes = [C.c_void_p]
        for f in (L.c2o_scan_make_bev, L.c2o_scan_ingest):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.c2o_scan_make_contours.argtypes = [C.c_void_p]
        L.c2o_scan_get_bev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.c2o_scan_n_views.argtypes = [C.c_void_p, C.c_int]
        L.c2o_scan_get_views.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.c2o_scan_get_head.argtypes = [C.c_void_p, C.c_void_p]
        L.c2o_eig2f.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.c2o_ccl8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.c2o_gauss_pdf_f.restype = C.c_float
        L.c2o_gauss_pdf_f.argtypes = [C.c_float, C.c_float, C.c_float]
        L.c2o_exp.restype = C.c_double
        L.c2o_exp.argtypes = [C.c_double]
        L.c2o_atan2f.restype = C.c_float
        L.c2o_atan2f.argtypes = [C.c_float, C.c_float]
        L.c2o_acosf.restype = C.c_float
        L.c2o_acosf.argtypes = [C.c_float]
        L.c2o_db_create.restype = C.c_void_p
        L.c2o_db_create.argtypes = [C.POINTER(D.DbConfig)]
        L.c2o_db_free.argtypes = [C.c_void_p]
        L.c2o_db_add_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        L.c2o_db_push_and_balance.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.c2o_db_n_scans.argtypes = [C.c_void_p]
        L.c2o_db_rebalance_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.c2o_db_layer_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.c2o_db_bucket_tree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.c2o_db_query.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(D.ScoreEnsemble), C.POINTER(D.ScoreEnsemble),
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.c2o_db_layer_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.c2o_run_loop.argtypes = [C.c_void_p, C.POINTER(D.CmConfig), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int, C.c_int, C.POINTER(D.ScoreEnsemble),
                                   C.POINTER(D.ScoreEnsemble), C.c_void_p, C.c_void_p]
        L.c2o_scan_from_descriptor.restype = C.c_void_p
        L.c2o_scan_from_descriptor.argtypes = [C.POINTER(D.CmConfig), C.c_void_p, C.c_void_p]
        L.c2o_std_sort_words.argtypes = [C.c_void_p, C.c_int, C.c_int]
        assert L.c2o_sizeof_scan_head() == D.SCAN_HEAD_DTYPE.itemsize
        assert L.c2o_sizeof_query_result() == D.QUERY_RESULT_DTYPE.itemsize
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's vendored nanoflann behind TreeBucket::knnSearch's call sequence, or None if not built."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libref_knn.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_knn_build.restype = C.c_void_p
        R.ref_knn_build.argtypes = [C.c_void_p, C.c_int]
        R.ref_knn_free.argtypes = [C.c_void_p]
        R.ref_knn_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        _REF = R
    return _REF


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def refine_eval(src: "Scan", tgt: "Scan", T, p):
    """Jet cost + gradient of the GMM-L2 problem (pairs selected at T = (cos, sin, tx, ty)) at p = (x, y, theta).
    Returns (cost, grad[3], n_pairs)."""
    L = lib()
    L.c2o_refine_hook.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    T = np.ascontiguousarray(T, np.float64)
    p = np.ascontiguousarray(p, np.float64)
    out = np.zeros(16, np.float64)
    L.c2o_refine_hook(0, src.h, tgt.h, _ptr(T), _ptr(p), _ptr(out))
    return out[0], out[1:4].copy(), int(out[4])


def refine_solve(src: "Scan", tgt: "Scan", T):
    """ConstellCorrelation::calcCorrelation restated (c2o_refine.hpp) from T = (cos, sin, tx, ty)."""
    L = lib()
    L.c2o_refine_hook.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    T = np.ascontiguousarray(T, np.float64)
    out = np.zeros(16, np.float64)
    L.c2o_refine_hook(1, src.h, tgt.h, _ptr(T), None, _ptr(out))
    return dict(correlation=out[0], x=out[1:4].copy(), initial_cost=out[4], final_cost=out[5], iterations=int(out[6]),
                termination=int(out[7]), n_eval=int(out[8]), n_pairs=int(out[9]), norm=out[10])


class Scan:
    """One ContourManager (oracle side)."""

    def __init__(self, cfg: D.CmConfig, int_id: int = 0, _handle=None):
        self.cfg = cfg
        self.h = _handle if _handle is not None else lib().c2o_scan_create(C.byref(cfg), int_id)

    @classmethod
    def from_descriptor(cls, cfg: D.CmConfig, head: np.ndarray, views: np.ndarray):
        """head: one SCAN_HEAD_DTYPE record; views: the scan's VIEW_CAP-record arena (sorted, level by level)."""
        head = np.ascontiguousarray(head).reshape(1)
        views = np.ascontiguousarray(views)
        return cls(cfg, _handle=lib().c2o_scan_from_descriptor(C.byref(cfg), _ptr(head), _ptr(views)))

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.c2o_scan_free(self.h)
            self.h = None

    def ingest(self, pts: np.ndarray):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 4)
        lib().c2o_scan_ingest(self.h, _ptr(pts), pts.shape[0])
        return self

    def make_bev(self, pts: np.ndarray):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 4)
        lib().c2o_scan_make_bev(self.h, _ptr(pts), pts.shape[0])
        return self

    def make_contours(self):
        lib().c2o_scan_make_contours(self.h)
        return self

    def bev(self):
        n = self.cfg.n_row * self.cfg.n_col
        b, r, c = (np.empty(n, np.float32) for _ in range(3))
        lib().c2o_scan_get_bev(self.h, _ptr(b), _ptr(r), _ptr(c))
        return b, r, c

    def views(self, level: int, presort: bool = False) -> np.ndarray:
        n = lib().c2o_scan_n_views(self.h, level)
        out = np.zeros(n, D.VIEW_DTYPE)
        if n:
            lib().c2o_scan_get_views(self.h, level, int(presort), _ptr(out))
        return out

    def head(self) -> np.ndarray:
        out = np.zeros(1, D.SCAN_HEAD_DTYPE)
        lib().c2o_scan_get_head(self.h, _ptr(out))
        return out[0]


class DB:
    """ContourDB (oracle side)."""

    def __init__(self, cfg: D.DbConfig):
        self.cfg = cfg
        self.h = lib().c2o_db_create(C.byref(cfg))
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.c2o_db_free(self.h)
            self.h = None

    def add_scan(self, scan: Scan, ts: float):
        self._keep.append(scan)
        lib().c2o_db_add_scan(self.h, scan.h, ts)

    def push_and_balance(self, seed: int, ts: float):
        lib().c2o_db_push_and_balance(self.h, seed, ts)

    def n_scans(self):
        return lib().c2o_db_n_scans(self.h)

    def rebalance_stats(self):
        """(rebalancing moves so far, buckets a literal reference would have searched through a stale index afterwards, how many of
        those were the bucket that gave keys away)."""
        m, s, d = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        lib().c2o_db_rebalance_stats(self.h, C.byref(m), C.byref(s), C.byref(d))
        return int(m.value), int(s.value), int(d.value)

    def layer_state(self, ll: int):
        rng = np.zeros(D.NUM_BUCKETS + 1, np.float32)
        ts = np.zeros(D.NUM_BUCKETS, np.int32)
        bs = np.zeros(D.NUM_BUCKETS, np.int32)
        lib().c2o_db_layer_state(self.h, ll, _ptr(rng), _ptr(ts), _ptr(bs))
        return rng, ts, bs

    def bucket_tree(self, ll: int, bucket: int):
        _, ts, _ = self.layer_state(ll)
        n = int(ts[bucket])
        keys = np.zeros((n, D.KEY_DIM), np.float32)
        gidx = np.zeros(n, np.int32)
        seq = np.zeros(n, np.int32)
        if n:
            lib().c2o_db_bucket_tree(self.h, ll, bucket, _ptr(keys), _ptr(gidx), _ptr(seq))
        return keys, gidx, seq

    def query(self, scan: Scan, lb: D.ScoreEnsemble, ub: D.ScoreEnsemble, cap: int = 2048):
        res = np.zeros(1, D.QUERY_RESULT_DTYPE)
        hints = np.zeros(cap, D.HINT_DTYPE)
        scores = np.zeros(cap, D.PAIR_SCORE_DTYPE)
        n = lib().c2o_db_query(self.h, scan.h, C.byref(lb), C.byref(ub), _ptr(res), _ptr(hints), _ptr(scores), cap)
        assert n <= cap
        return res[0], hints[:n], scores[:n]

    def layer_knn(self, ll: int, key: np.ndarray, k: int, max_dist_sq: float):
        key = np.ascontiguousarray(key, np.float32)
        gidx = np.zeros(k, np.int32)
        seq = np.zeros(k, np.int32)
        dist = np.zeros(k, np.float32)
        n = lib().c2o_db_layer_knn(self.h, ll, _ptr(key), k, max_dist_sq, _ptr(gidx), _ptr(seq), _ptr(dist))
        return gidx[:n], seq[:n], dist[:n]


def run_loop(db: DB, cfg: D.CmConfig, pts: np.ndarray, offsets: np.ndarray, first_id: int, ts: np.ndarray,
             do_query: bool, do_add: bool, lb: D.ScoreEnsemble, ub: D.ScoreEnsemble):
    """CPU baseline driver (single thread).  Returns (results, stage seconds[5])."""
    pts = np.ascontiguousarray(pts, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    ts = np.ascontiguousarray(ts, np.float64)
    B = len(offsets) - 1
    res = np.zeros(B, D.QUERY_RESULT_DTYPE)
    t = np.zeros(5, np.float64)
    lib().c2o_run_loop(db.h, C.byref(cfg), _ptr(pts), _ptr(offsets), B, first_id, _ptr(ts), int(do_query), int(do_add),
                       C.byref(lb), C.byref(ub), _ptr(res), _ptr(t))
    return res, t
