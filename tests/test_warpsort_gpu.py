"""GPU: the warp-cooperative replay of libstdc++ std::sort that the contour kernel runs (csrc/stdsort.cuh: parallel
unguarded-partition emulation + stable rank pass) gives the same permutation as the REAL std::sort (oracle side, compiled
from <algorithm>) for the comparators the reference uses, tie-heavy and adversarial inputs included
(include/cont2/contour_mng.h:596-599 cell_cnt descending, :871-874 bit_pos ascending)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built_lib):
    from contour_context_b200.engine import Engine

    e = Engine(scan_capacity=8, max_batch=2, max_points=1 << 16)
    yield e
    e.close()


def _both(eng, built_lib, oracle, words, desc):
    a, b = words.copy(), words.copy()
    assert built_lib.c2g_selftest_warpsort(eng.h, a.ctypes.data_as(C.c_void_p), len(a), desc) == 0
    oracle.lib().c2o_std_sort_words(b.ctypes.data_as(C.c_void_p), len(b), desc)
    return a, b


@pytest.mark.parametrize("desc", [0, 1])
def test_random_with_ties(eng, built_lib, oracle, desc):
    rng = np.random.default_rng(1)
    for n in list(range(1, 48)) + [63, 64, 65, 70, 100, 127, 128, 129, 257, 700, 2048]:
        for kmax in (1, 2, 3, 8, 50, 60000):
            for _ in range(3):
                keys = rng.integers(0, kmax, n).astype(np.uint32)
                words = (keys << 16) | np.arange(n, dtype=np.uint32)
                a, b = _both(eng, built_lib, oracle, words, desc)
                assert np.array_equal(a, b), (n, kmax)


@pytest.mark.parametrize("desc", [0, 1])
def test_contour_like_area_distributions(eng, built_lib, oracle, desc):
    """cell_cnt of a level: a few large contours and a long tail of 3..8-cell ones (ties everywhere)."""
    rng = np.random.default_rng(2)
    for n in (20, 40, 55, 72, 90, 120, 128):
        for _ in range(40):
            big = rng.integers(9, 260, max(1, n // 6))
            small = rng.integers(3, 9, n - len(big))
            keys = rng.permutation(np.concatenate([big, small])).astype(np.uint32)
            words = (keys << 16) | np.arange(n, dtype=np.uint32)
            a, b = _both(eng, built_lib, oracle, words, desc)
            assert np.array_equal(a, b), n


@pytest.mark.parametrize("desc", [0, 1])
def test_adversarial_patterns(eng, built_lib, oracle, desc):
    """Sorted, reversed, organ-pipe and median-of-3 killer inputs (the depth-limit heapsort fallback)."""
    for n in (17, 33, 100, 128, 500, 1500):
        pats = [np.arange(n), np.arange(n)[::-1], np.minimum(np.arange(n), np.arange(n)[::-1]), np.arange(n) % 7]
        k = n // 2
        killer = np.zeros(n, dtype=np.int64)
        for i in range(k):
            killer[i] = i + 1 if i % 2 == 0 else k + i + (1 if k % 2 == 0 else 0)
        killer[k:] = np.arange(1, n - k + 1) * 2
        pats.append(killer % 65536)
        for p in pats:
            words = ((p.astype(np.uint32) & 0xFFFF) << 16) | np.arange(n, dtype=np.uint32)
            a, b = _both(eng, built_lib, oracle, words, desc)
            assert np.array_equal(a, b), n
