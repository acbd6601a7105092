// c2g_common.cuh — shared device/host definitions of the B200-native cont2contops hot path.
// Compiled with -fmad=false: the reference is built without FMA contraction (CMakeLists.txt:4,10-11), so every float /
// double multiply-add here must round twice exactly like the x86-64 SSE2 code of the reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/c2g_types.h"

#define C2G_CUDA_TRY(expr)                       \
  do {                                           \
    cudaError_t e__ = (expr);                    \
    if (e__ != cudaSuccess) return -(int) e__;   \
  } while (0)

// Scalar parameters every ingest kernel needs, precomputed on the host in float exactly like the ContourManager
// constructor does (include/cont2/contour_mng.h:478-498).
struct C2gIngestParams {
  c2g_cm_config cfg;
  float x_min_pad, x_max_pad, y_min_pad, y_max_pad;  // x_min_ + 0.01f etc. (contour_mng.h:450-452)
  float half_row_f, half_col_f;                      // float(n_row / 2), float(n_col / 2)
  int half_row, half_col;
  int n_cells;
  int exp_mode;  // which exp() the host's libm implements: 0 unknown (libdevice), 1 glibc, 2 glibc FMA variant (c2g_libm.cuh)
};

// Monotone map float -> uint32 (total order of finite floats and infinities; NaNs never reach it).
__host__ __device__ __forceinline__ uint32_t c2g_orderable(float h) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(h);
#else
  union {
    float f;
    uint32_t u;
  } cv;
  cv.f = h;
  uint32_t u = cv.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float c2g_from_orderable(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union {
    float f;
    uint32_t u;
  } cv;
  cv.u = u;
  return cv.f;
#endif
}

// BEV cell record produced by the scatter kernel: (orderable(height) << 32) | (0xFFFFFFFF - point_index); 0 = empty.
// A max over these keys implements "highest point wins, the earliest point in file order wins ties"
// (strict `<` in include/cont2/contour_mng.h:517).
typedef unsigned long long c2g_cellkey;

// What the scatter kernel hands to the contour kernel per scan of a batch (bev_scatter.cu): bit (c & 31) of word
// r * ceil(n_col / 32) + (c >> 5) of plane l = `bev(r, c) > lv_grads[l]`; fg = the cells of plane 0 in raster order as
// (height, row_f, col_f, 0) with the continuous coordinates of the point that set the height (bev_pixfs_, contour_mng.h:435);
// hdr = (occupied cells = bev_pixfs_.size(), foreground cells).  tiles: the raw 64-bit cells, full-tile variant only.
struct C2gBevOut {
  uint32_t *planes;    // [B][C2G_NLEV][n_row * ceil(n_col / 32)]
  float4 *fg;          // [B][n_cells]
  int2 *hdr;           // [B]
  c2g_cellkey *tiles;  // [B][n_cells] or nullptr
};

// Compact GMM ellipse of one contour view (GMMPair::GMMEllipse, include/cont2/correlation.h:24-33): one 32-byte sector per
// view, written next to the view by the contour kernel, read by the GMM-L2 kernels instead of the 80-byte view record.
// cov = ContourView::getManualCov() (float, column-major), w = cell_cnt, maj = sqrtf(eig_vals[1]) (correlation.h:64,72).
struct __align__(32) c2g_ell {
  float mx, my, c00, c10, c01, c11, w, maj;
};

// `sqrt(x) < y` with IEEE semantics (y >= 0 or NaN) without the square root in all but borderline cases: fl(sqrt(x)) is within
// half an ulp of the real root, so the comparison is decided by x against y^2 whenever they differ by more than a few ulp.
__device__ __forceinline__ bool c2g_sqrt_lt(double x, double y) {
  const double y2 = y * y;
  if (x < y2 * (1.0 - 1e-15)) return true;
  if (x > y2 * (1.0 + 1e-15)) return false;
  return sqrt(x) < y;
}

// Every extern "C" entry point that touches CUDA makes the context's device current for its duration and restores the
// caller's device on exit: a process may own contexts on several GPUs (streams, events and kernel attributes are per device).
struct C2gDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit C2gDeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~C2gDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  C2gDeviceGuard(const C2gDeviceGuard &) = delete;
  C2gDeviceGuard &operator=(const C2gDeviceGuard &) = delete;
};

// cudaFuncSetAttribute is per device: remember which devices of this process already have the attribute (the current device
// is the context's: C2gDeviceGuard)
static inline bool c2g_first_use_on_device(unsigned long long &mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (mask & (1ull << dev)) return false;
  mask |= 1ull << dev;
  return true;
}
