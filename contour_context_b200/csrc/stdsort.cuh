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
