"""ctypes binding of include/c2g.h (libc2g.so).  There is NO CPU fallback: if the CUDA library is missing or fails to
load, importing the hot path raises."""
import ctypes as C
import os

import numpy as np

from . import ctypes_defs as D

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ERR_ARG, ERR_CAPACITY, ERR_STATE = -1000, -1001, -1002


class C2gError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "libc2g.so")


def lib():
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise C2gError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the hot path)")
        L = C.CDLL(p)
        vp, ip, ll = C.c_void_p, C.c_int, C.c_longlong
        L.c2g_abi_version.restype = ip
        L.c2g_sizeof.argtypes = [ip]
        L.c2g_create.argtypes = [C.POINTER(D.CmConfig), C.POINTER(D.DbConfig), ip, ip, ip, ll, C.POINTER(vp)]
        L.c2g_destroy.argtypes = [vp]
        L.c2g_set_stream.argtypes = [vp, vp]
        L.c2g_sync.argtypes = [vp]
        L.c2g_ingest.argtypes = [vp, vp, vp, ip, ip, ip, vp]
        L.c2g_ingest_xyz.argtypes = [vp, vp, vp, ip, ip, ip, vp]
        L.c2g_ingest_bev_only.argtypes = [vp, vp, vp, ip, ip]
        L.c2g_get_heads.argtypes = [vp, ip, ip, vp]
        L.c2g_get_views.argtypes = [vp, ip, vp]
        L.c2g_get_bev.argtypes = [vp, ip, vp, vp, vp]
        L.c2g_get_tiles.argtypes = [vp, ip, vp]
        L.c2g_get_bev_compact.argtypes = [vp, ip, ip, vp, vp, vp]
        L.c2g_copy_slots.argtypes = [vp, ip, ip, ip]
        L.c2g_db_set_layer.argtypes = [vp, ip, ip, vp, vp, vp, vp, vp]
        L.c2g_query.argtypes = [vp, ip, ip, C.POINTER(D.ScoreEnsemble), C.POINTER(D.ScoreEnsemble), vp, vp, vp]
        L.c2g_query_async.argtypes = [vp, ip, ip, C.POINTER(D.ScoreEnsemble), C.POINTER(D.ScoreEnsemble)]
        L.c2g_query_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(ll)]
        L.c2g_query_export.argtypes = [vp, ip, vp, vp, vp]
        L.c2g_finish_from_scores.argtypes = [vp, ip, ip, C.POINTER(D.ScoreEnsemble), vp, vp, vp]
        L.c2g_db_add_scans.argtypes = [vp, ip, ip, vp]
        L.c2g_db_push_and_balance.argtypes = [vp, ip, C.c_double]
        L.c2g_db_size.argtypes = [vp]
        L.c2g_online_stage.argtypes = [vp, vp, vp, ip, ip, vp]
        L.c2g_online_stage_xyz.argtypes = [vp, vp, vp, ip, ip, vp]
        L.c2g_online_commit.argtypes = [vp, vp, vp, C.POINTER(D.ScoreEnsemble), C.POINTER(D.ScoreEnsemble), vp]
        L.c2g_online_window.argtypes = [vp, vp, vp, ip, ip, vp, vp, vp, C.POINTER(D.ScoreEnsemble), C.POINTER(D.ScoreEnsemble), vp]
        L.c2g_work_counters.argtypes = [vp, ip, vp]
        L.c2g_online_host_seconds.argtypes = [vp, vp]
        L.c2g_online_runs.restype = ll
        L.c2g_online_runs.argtypes = [vp]
        L.c2g_online_groups.restype = ll
        L.c2g_online_groups.argtypes = [vp]
        L.c2g_db_sync.argtypes = [vp]
        L.c2g_db_layer_state.argtypes = [vp, ip, vp, vp, vp]
        L.c2g_db_bucket_tree.argtypes = [vp, ip, ip, vp, vp, vp]
        L.c2g_hostdb_create.restype = vp
        L.c2g_hostdb_create.argtypes = [ip, C.c_double, C.c_double]
        L.c2g_hostdb_free.argtypes = [vp]
        L.c2g_hostdb_push_key.argtypes = [vp, ip, vp, C.c_double, ip, ip]
        L.c2g_hostdb_balance.argtypes = [vp, ip, C.c_double]
        L.c2g_hostdb_state.argtypes = [vp, ip, vp, vp, vp]
        L.c2g_hostdb_tree.argtypes = [vp, ip, ip, vp, vp, vp]
        L.c2g_debug_clocks.argtypes = [vp, vp]
        L.c2g_scatter_deferred.argtypes = [vp, vp]
        L.c2g_exp_mode.argtypes = [vp]
        L.c2g_selftest_libm.argtypes = [ip, ip, vp, vp]
        L.c2g_launch_count.restype = ll
        L.c2g_query_profile.argtypes = [vp, C.c_int, vp]
        L.c2g_launch_count.argtypes = [vp]
        L.c2g_selftest_stdsort.argtypes = [vp, ip, ip]
        L.c2g_selftest_warpsort.argtypes = [vp, vp, ip, ip]
        sizes = [D.SCAN_HEAD_DTYPE.itemsize, D.VIEW_DTYPE.itemsize, D.BCI_DTYPE.itemsize, D.HINT_DTYPE.itemsize,
                 D.PAIR_SCORE_DTYPE.itemsize, D.QUERY_RESULT_DTYPE.itemsize, C.sizeof(D.CmConfig), C.sizeof(D.DbConfig)]
        for i, s in enumerate(sizes):
            if L.c2g_sizeof(i) != s:
                raise C2gError(f"ABI mismatch: c2g_sizeof({i}) = {L.c2g_sizeof(i)}, binding expects {s}")
        _LIB = L
    return _LIB


def check(rc: int, what: str = "c2g call"):
    if rc != 0:
        raise C2gError(f"{what} failed with code {rc}" + (f" (cudaError {-rc})" if -999 < rc < 0 else ""))


def ptr(a):
    """void* of a numpy array, a torch tensor (host or device) or an int address."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor
