"""CPU-only: the host/device replay of libstdc++ std::sort (csrc/stdsort.cuh) produces the same permutation as the real
std::sort for the comparators the reference uses, including heavy ties (include/cont2/contour_mng.h:596-599,871-874)."""
import ctypes as C

import numpy as np
import pytest


def _both(built_lib, oracle, words, desc):
    a = words.copy()
    b = words.copy()
    assert built_lib.c2g_selftest_stdsort(a.ctypes.data_as(C.c_void_p), len(a), desc) == 0
    oracle.lib().c2o_std_sort_words(b.ctypes.data_as(C.c_void_p), len(b), desc)
    return a, b


@pytest.mark.parametrize("desc", [0, 1])
def test_random_with_ties(built_lib, oracle, desc):
    rng = np.random.default_rng(1)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 257, 700, 2048]:
        for kmax in (1, 2, 3, 8, 50, 60000):
            keys = rng.integers(0, kmax, n).astype(np.uint32)
            words = (keys << 16) | np.arange(n, dtype=np.uint32)
            a, b = _both(built_lib, oracle, words, desc)
            assert np.array_equal(a, b), (n, kmax)


@pytest.mark.parametrize("desc", [0, 1])
def test_adversarial_patterns(built_lib, oracle, desc):
    """Sorted, reversed, organ-pipe and median-of-3 killer inputs reach the heapsort fallback of introsort."""
    for n in (17, 33, 100, 500, 1500):
        pats = [np.arange(n), np.arange(n)[::-1], np.minimum(np.arange(n), np.arange(n)[::-1]), np.arange(n) % 7]
        # median-of-3 killer (Musser)
        k = n // 2
        killer = np.zeros(n, dtype=np.int64)
        for i in range(k):
            killer[i] = i + 1 if i % 2 == 0 else k + i + (1 if k % 2 == 0 else 0)
        killer[k:] = np.arange(1, n - k + 1) * 2
        pats.append(killer % 65536)
        for p in pats:
            words = ((p.astype(np.uint32) & 0xFFFF) << 16) | np.arange(n, dtype=np.uint32)
            a, b = _both(built_lib, oracle, words, desc)
            assert np.array_equal(a, b), n
