"""CPU-only: the libm restatements the kernels run (csrc/c2g_libm.cuh, compiled here for the host) agree BIT FOR BIT with
this machine's glibc on millions of inputs — std::exp in the variant the host's libm dispatches to (FMA or not),
std::atan2(float, float), std::acos(float), atanf.  These are the calls whose bits reach retrieval keys and BCIs
(include/tools/algos.h:53-56, include/cont2/contour_mng.h:860,1191-1192)."""
import ctypes as C

import numpy as np


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _both(built_lib, oracle, kind, x, out_dtype):
    a = np.zeros(len(x) if kind != 2 else len(x) // 2, out_dtype)
    b = np.zeros_like(a)
    n = len(a)
    assert built_lib.c2g_selftest_libm(kind, n, _p(x), _p(a)) == 0
    oracle.lib().c2o_vec_libm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    oracle.lib().c2o_vec_libm(kind, n, _p(x), _p(b))
    return a, b


def test_exp_variant_matches_host_libm(built_lib, oracle):
    mode = built_lib.c2g_selftest_libm(5, 0, None, None)
    assert mode in (1, 2), "neither glibc exp variant reproduces this host's exp(): device would fall back to libdevice"
    rng = np.random.default_rng(3)
    t = rng.uniform(0, 12, 1_500_000).astype(np.float32)
    x = np.concatenate([(-0.5 * t.astype(np.float64)) * t.astype(np.float64), rng.uniform(-500, 500, 1_000_000),
                        rng.uniform(-1e-3, 1e-3, 200_000), np.array([0.0, -0.0, 1e-300, -1e-300, -745.0, 709.0])])
    a, b = _both(built_lib, oracle, mode - 1, np.ascontiguousarray(x), np.float64)
    assert a.view(np.uint64).tobytes() == b.view(np.uint64).tobytes(), int((a.view(np.uint64) != b.view(np.uint64)).sum())


def test_atan2f_acosf_atanf_match_host_libm(built_lib, oracle):
    rng = np.random.default_rng(4)
    yx = rng.uniform(-150, 150, 4_000_000).astype(np.float32)
    yx[:2000] = rng.choice(np.array([0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 1e30, np.inf, -np.inf], np.float32), 2000)
    a, b = _both(built_lib, oracle, 2, yx, np.float32)
    assert a.view(np.uint32).tobytes() == b.view(np.uint32).tobytes()
    c = np.concatenate([rng.uniform(-1, 1, 2_000_000), np.array([1.0, -1.0, 0.0, 0.5, -0.5, 1e-9])]).astype(np.float32)
    a, b = _both(built_lib, oracle, 3, c, np.float32)
    assert a.view(np.uint32).tobytes() == b.view(np.uint32).tobytes()
    t = np.concatenate([rng.uniform(-50, 50, 1_000_000), rng.uniform(-1, 1, 1_000_000) * 1e-3, 10.0 ** rng.uniform(-8, 9, 200_000)]).astype(np.float32)
    a, b = _both(built_lib, oracle, 4, t, np.float32)
    assert a.view(np.uint32).tobytes() == b.view(np.uint32).tobytes()
