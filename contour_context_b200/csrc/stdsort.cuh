// stdsort.cuh — a host/device replay of libstdc++'s std::sort (bits/stl_algo.h: __sort -> __introsort_loop +
// __final_insertion_sort, _S_threshold = 16, median-of-3 pivot, unguarded partition, heapsort fallback at depth 0).
//
// Why it exists: the reference orders contours with an UNSTABLE std::sort on cell_cnt_ (include/cont2/contour_mng.h:
// 596-599), BCI neighbours on bit_pos (:871-874) and potential pairs on orie_diff (:340-342).  Ties are common, and
// the resulting order decides which contour is "seq 0..9" of a level, hence keys, BCIs and every ConstellationPair.
// The only way to reproduce the order on the device is to run the very same algorithm.  tests/test_stdsort.py
// compares this replay (compiled for the host) with the real std::sort on tie-heavy inputs.
//
// Elements are moved by value (T must be trivially copyable); `comp(a, b)` is the strict-weak "a before b" predicate.
#pragma once

#ifndef C2G_HD
#ifdef __CUDACC__
#define C2G_HD __host__ __device__ __forceinline__
#else
#define C2G_HD inline
#endif
#endif

namespace c2g_sort {

template <typename T>
C2G_HD void swp(T &a, T &b) {
  T t = a;
  a = b;
  b = t;
}

template <typename T, typename Cmp>
C2G_HD void unguarded_linear_insert(T *last, Cmp comp) {
  T val = *last;
  T *next = last - 1;
  while (comp(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

template <typename T, typename Cmp>
C2G_HD void insertion_sort(T *first, T *last, Cmp comp) {
  if (first == last) return;
  for (T *i = first + 1; i != last; ++i) {
    if (comp(*i, *first)) {
      T val = *i;
      for (T *p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i + 1)
      *first = val;
    } else {
      unguarded_linear_insert(i, comp);
    }
  }
}

template <typename T, typename Cmp>
C2G_HD void push_heap_(T *first, long hole, long top, T value, Cmp comp) {
  long parent = (hole - 1) / 2;
  while (hole > top && comp(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <typename T, typename Cmp>
C2G_HD void adjust_heap(T *first, long hole, long len, T value, Cmp comp) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (comp(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  push_heap_(first, hole, top, value, comp);
}

template <typename T, typename Cmp>
C2G_HD void heap_sort(T *first, T *last, Cmp comp) {  // __partial_sort(first, last, last): __heap_select + __sort_heap
  long len = last - first;
  if (len >= 2) {  // __make_heap
    long parent = (len - 2) / 2;
    while (true) {
      T value = first[parent];
      adjust_heap(first, parent, len, value, comp);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {  // __sort_heap
    --last;
    T value = *last;  // __pop_heap(first, last, last)
    *last = *first;
    adjust_heap(first, 0L, (long) (last - first), value, comp);
  }
}

template <typename T, typename Cmp>
C2G_HD void move_median_to_first(T *result, T *a, T *b, T *c, Cmp comp) {
  if (comp(*a, *b)) {
    if (comp(*b, *c))
      swp(*result, *b);
    else if (comp(*a, *c))
      swp(*result, *c);
    else
      swp(*result, *a);
  } else if (comp(*a, *c))
    swp(*result, *a);
  else if (comp(*b, *c))
    swp(*result, *c);
  else
    swp(*result, *b);
}

template <typename T, typename Cmp>
C2G_HD T *unguarded_partition(T *first, T *last, T *pivot, Cmp comp) {
  while (true) {
    while (comp(*first, *pivot)) ++first;
    --last;
    while (comp(*pivot, *last)) --last;
    if (!(first < last)) return first;
    swp(*first, *last);
    ++first;
  }
}

// std::sort(first, first + n, comp).  The recursion of __introsort_loop (recurse right, loop left) is unrolled with an
// explicit stack of (last, depth) records: the right part [cut, last) is processed first exactly like the recursive
// call would, then the left part continues with the decremented depth limit.
template <typename T, typename Cmp>
C2G_HD void std_sort(T *first, long n, Cmp comp) {
  if (n <= 0) return;
  T *last = first + n;
  long lg = 0;  // std::__lg(n)
  for (long t = n; t > 1; t >>= 1) ++lg;
  // work stack: ranges still to be processed by __introsort_loop. A range is pushed when we descend into its right
  // part first (recursion) — the left part [lo, cut) is handled after the right part returns.
  struct Frame {
    T *lo, *hi;
    long depth;
  };
  Frame stack[64];
  int sp = 0;
  stack[sp++] = Frame{first, last, lg * 2};
  while (sp > 0) {
    Frame f = stack[--sp];
    T *lo = f.lo, *hi = f.hi;
    long depth = f.depth;
    // __introsort_loop(lo, hi, depth)
    while (hi - lo > 16) {
      if (depth == 0) {
        heap_sort(lo, hi, comp);
        break;
      }
      --depth;
      T *mid = lo + (hi - lo) / 2;
      move_median_to_first(lo, lo + 1, mid, hi - 1, comp);
      T *cut = unguarded_partition(lo + 1, hi, lo, comp);
      // recursive call on [cut, hi) happens BEFORE the loop continues on [lo, cut). Both sub-problems are independent
      // (disjoint ranges), so processing order does not change the result; we push the left part and iterate on the
      // right part to keep the stack shallow like the original recursion.
      stack[sp++] = Frame{lo, cut, depth};
      lo = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    insertion_sort(first, first + 16, comp);
    for (T *i = first + 16; i != last; ++i) unguarded_linear_insert(i, comp);
  } else {
    insertion_sort(first, last, comp);
  }
}

}  // namespace c2g_sort

#ifdef __CUDACC__
// ---- warp-cooperative replay of the same std::sort ----------------------------------------------------------------------------
// One lane replaying libstdc++'s introsort spends ~100 cycles per element move (dependent shared-memory reads and branches); six
// such lanes were the critical path of the contour kernel in round 1.  The algorithm has a data-parallel reading that gives the
// SAME permutation:
//  * __unguarded_partition(first + 1, last, pivot at first): the left scan stops exactly at the elements x with !comp(x, pivot)
//    ("left stoppers", ascending positions L_0 < L_1 < ..), the right scan at the elements with !comp(pivot, x) ("right stoppers",
//    descending positions R_0 > R_1 > .., the pivot's own slot included).  Neither scan ever looks at a swapped slot again, so the
//    k-th swap exchanges L_k and R_k for all k < K = #{k : L_k < R_k}, and the returned cut is min(L_K, R_{K-1}) (L_0 if K = 0).
//    Stopper ranks come from ballots, the swaps are disjoint and run in parallel.  (Checked against the serial loop on 2e5
//    tie-heavy inputs; tests/test_stdsort.py compares the whole replay with the real std::sort.)
//  * __final_insertion_sort is a STABLE sort of whatever arrangement the partition phase left, and no element leaves its final
//    segment of <= 16 slots (or its heap-sorted range), so the final slot of element i is
//    max(i - 15, 0) + #{j in [i - 15, i + 15] : a_j before a_i, or equivalent to it with j < i}.
// Elements are packed words (key << 16 | payload), DESC = larger key first; `a`, `posL`, `posR` are in shared memory (n entries
// each); called by all 32 lanes of a warp.  Ranges of more than 128 elements (cluttered scans only) finish with the serial
// insertion sort on one lane; the depth-limit heapsort fallback is serial too.
namespace c2g_sort {

template <bool DESC>
__device__ __forceinline__ bool w_before(uint32_t x, uint32_t y) {
  return DESC ? (x >> 16) > (y >> 16) : (x >> 16) < (y >> 16);
}

template <bool DESC>
__device__ void warp_std_sort(uint32_t *a, int n, uint16_t *posL, uint16_t *posR, int lane) {
  constexpr unsigned FULLW = 0xFFFFFFFFu;
  const unsigned lt = (1u << lane) - 1u;
  auto cmp = [](uint32_t x, uint32_t y) { return w_before<DESC>(x, y); };
  if (n <= 1) return;
  if (n > 16) {
    int lg = 31 - __clz(n);
    int st_lo[40], st_hi[40], st_d[40];  // at most one frame per level of the depth limit (2 lg n <= 22)
    int sp = 0;
    st_lo[0] = 0;
    st_hi[0] = n;
    st_d[0] = 2 * lg;
    sp = 1;
    while (sp > 0) {
      --sp;
      int lo = st_lo[sp], hi = st_hi[sp], depth = st_d[sp];
      while (hi - lo > 16) {
        if (depth == 0) {
          if (lane == 0) heap_sort(a + lo, a + hi, cmp);
          __syncwarp();
          break;
        }
        --depth;
        if (lane == 0) move_median_to_first(a + lo, a + lo + 1, a + lo + (hi - lo) / 2, a + hi - 1, cmp);
        __syncwarp();
        const uint32_t p = a[lo];
        int nL = 0, nR = 0;
        for (int c0 = lo + 1; c0 < hi; c0 += 32) {  // left stoppers, ascending
          const int i = c0 + lane;
          const bool is = i < hi && !w_before<DESC>(a[i], p);
          const unsigned m = __ballot_sync(FULLW, is);
          if (is) posL[lo + nL + __popc(m & lt)] = (uint16_t) i;
          nL += __popc(m);
        }
        for (int c0 = 0; c0 < hi - lo; c0 += 32) {  // right stoppers, descending (the pivot's slot `lo` is the sentinel)
          const int j = hi - 1 - (c0 + lane);
          const bool is = j >= lo && !w_before<DESC>(p, a[j]);
          const unsigned m = __ballot_sync(FULLW, is);
          if (is) posR[lo + nR + __popc(m & lt)] = (uint16_t) j;
          nR += __popc(m);
        }
        __syncwarp();
        int K = 0;
        const int nm = nL < nR ? nL : nR;
        for (int k0 = 0; k0 < nm; k0 += 32) {  // L_k < R_k holds for a prefix of k
          const int k = k0 + lane;
          const unsigned m = __ballot_sync(FULLW, k < nm && posL[lo + k] < posR[lo + k]);
          K += __popc(m);
          if (m != FULLW) break;
        }
        int cut;
        if (K > 0) {
          cut = posR[lo + K - 1];
          if (K < nL && (int) posL[lo + K] < cut) cut = posL[lo + K];
        } else
          cut = posL[lo];
        for (int k = lane; k < K; k += 32) {
          const int i = posL[lo + k], j = posR[lo + k];
          const uint32_t t = a[i];
          a[i] = a[j];
          a[j] = t;
        }
        __syncwarp();
        st_lo[sp] = lo;  // the left part waits on the stack, the right part is processed next (any order gives the same result)
        st_hi[sp] = cut;
        st_d[sp] = depth;
        ++sp;
        lo = cut;
      }
    }
  }
  if (n > 128) {  // __final_insertion_sort, serial
    if (lane == 0) {
      if (n > 16) {
        insertion_sort(a, a + 16, cmp);
        for (uint32_t *i = a + 16; i != a + n; ++i) unguarded_linear_insert(i, cmp);
      } else
        insertion_sort(a, a + n, cmp);
    }
    __syncwarp();
    return;
  }
  uint32_t v[4];
  int r[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = lane + 32 * q;
    v[q] = 0u;
    r[q] = -1;
    if (i < n) {
      const uint32_t x = a[i];
      const int w0 = i - 15 > 0 ? i - 15 : 0, w1 = i + 16 < n ? i + 16 : n;
      int rank = w0;
      for (int j = w0; j < w1; ++j) {
        const uint32_t y = a[j];
        rank += (w_before<DESC>(y, x) || (!w_before<DESC>(x, y) && j < i)) ? 1 : 0;
      }
      v[q] = x;
      r[q] = rank;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (r[q] >= 0) a[r[q]] = v[q];
  __syncwarp();
}

}  // namespace c2g_sort
#endif  // __CUDACC__
