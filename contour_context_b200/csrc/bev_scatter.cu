// bev_scatter.cu — K1: point cloud -> max-height BEV (ContourManager::makeBEV, include/cont2/contour_mng.h:505-556,
// with hashPointToImage :448-463).
//
// B200 mapping: one persistent CTA per SM, one scan per CTA iteration.  The whole 150x150 BEV lives in shared memory as
// 64-bit cell keys (180 KB of the 227 KB carve-out), so the only HBM traffic is the streaming read of the points
// (16 B/point, 128-bit ld.global.nc.L1::no_allocate, UNROLL independent loads in flight per thread) plus one coalesced
// 180 KB write of the finished tile.  The reference's "if (bev < h) bev = h" with first-point-wins ties becomes a 64-bit
// max over (orderable(h) << 32 | ~index); a plain shared-memory read filters the ~3/4 of the points that cannot raise
// their cell any more before the atomic is issued (max is monotone, so a stale read can only cause a redundant atomic,
// never a missed one).
#include "c2g_common.cuh"

namespace {

constexpr int K1_THREADS = 1024;
constexpr int K1_UNROLL = 8;

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// UNIT = true: reso_row == reso_col == 1.0f and the padded square is symmetric (the reference's only shipped setting,
// config/batch_bin_test_config.yaml:33-37 "TODO: reso other than 1.0"): x / 1.0f == x exactly, |x| <= x_max_pad replaces the
// two-sided test (and rejects NaN in the same compare), and the padded bounds already guarantee 0 <= row < n_row,
// 0 <= col < n_col, so only the reference's `row > 0` test remains.  The K1 kernel is issue-bound, not HBM-bound, without
// this diet (profiles/r1_ncu_full_summary.csv: 89 instructions per point, 54 % issue utilisation at 42 % DRAM).
template <bool UNIT>
__device__ __forceinline__ void scatter_point(const float4 pt, const uint32_t idx, const C2gIngestParams &P, c2g_cellkey *tile) {
  const float x = pt.x, y = pt.y;
  int row, col;
  if (UNIT) {
    if (!(fabsf(x) <= P.x_max_pad) || !(fabsf(y) <= P.y_max_pad)) return;
    if (__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq) return;
    row = __float2int_rd(x) + P.half_row;
    col = __float2int_rd(y) + P.half_col;
    if (row <= 0) return;  // `rc.first > 0` (contour_mng.h:515)
  } else {
    // hashPointToImage: reject outside the padded square or inside the blind radius (NaN x/y are dropped)
    if (x < P.x_min_pad || x > P.x_max_pad || y < P.y_min_pad || y > P.y_max_pad) return;
    if (__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq) return;
    if (!(x == x) || !(y == y)) return;
    row = (int) floorf(__fdiv_rn(x, P.cfg.reso_row)) + P.half_row;
    col = (int) floorf(__fdiv_rn(y, P.cfg.reso_col)) + P.half_col;
    if (row <= 0 || row >= P.cfg.n_row || col < 0 || col >= P.cfg.n_col) return;
  }
  const float h = __fadd_rn(P.cfg.lidar_height, pt.z);
  if (!(h > -1000.0f)) return;  // bev_ starts at -1000 and only strictly higher points are stored
  const c2g_cellkey key = ((c2g_cellkey) c2g_orderable(h) << 32) | (c2g_cellkey) (0xFFFFFFFFu - idx);
  c2g_cellkey *cell = tile + row * P.cfg.n_col + col;
  if (key > *(volatile c2g_cellkey *) cell) atomicMax(cell, key);
}

// Branch-free first half of scatter_point<true>: cell and key of a point, key 0 (never wins a max) for a rejected point.
// Same tests in the same float arithmetic; a NaN coordinate fails the first comparison exactly like in the branchy form.
__device__ __forceinline__ void point_key_unit(const float4 pt, const uint32_t idx, const C2gIngestParams &P, int &cell, c2g_cellkey &key) {
  const float x = pt.x, y = pt.y;
  bool ok = (fabsf(x) <= P.x_max_pad) && (fabsf(y) <= P.y_max_pad);
  ok = ok && !(__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq);
  const int row = __float2int_rd(x) + P.half_row, col = __float2int_rd(y) + P.half_col;
  ok = ok && row > 0;  // `rc.first > 0` (contour_mng.h:515)
  const float h = __fadd_rn(P.cfg.lidar_height, pt.z);
  ok = ok && (h > -1000.0f);  // bev_ starts at -1000 and only strictly higher points are stored
  cell = ok ? row * P.cfg.n_col + col : 0;
  key = ok ? (((c2g_cellkey) c2g_orderable(h) << 32) | (c2g_cellkey) (0xFFFFFFFFu - idx)) : 0ull;
}

template <bool UNIT>
__global__ void __launch_bounds__(K1_THREADS, 1)
bev_scatter_kernel(const float4 *__restrict__ pts, const long long *__restrict__ offsets, int B, C2gIngestParams P,
                   c2g_cellkey *__restrict__ tiles_out) {
  extern __shared__ c2g_cellkey tile[];
  const int tid = threadIdx.x;
  const int ncell = P.n_cells;
  for (int c = tid; c < ncell; c += K1_THREADS) tile[c] = 0ull;
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const long long beg = offsets[b];
    const int n = (int) (offsets[b + 1] - beg);
    const float4 *p = pts + beg;
    int i = tid;
    // main loop: UNROLL independent 128-bit loads per thread before any of them is consumed (a register double-buffered
    // software pipeline was measured 5 % slower: the kernel is bound by shared-memory atomics, not by load latency)
    for (; i + (K1_UNROLL - 1) * K1_THREADS < n; i += K1_UNROLL * K1_THREADS) {
      float4 v[K1_UNROLL];
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) v[u] = ld_stream_f4(p + i + u * K1_THREADS);
      if (UNIT) {
        // keys of all UNROLL points first (straight-line code), then all filter reads of the tile, then the atomics: the
        // shared-memory latencies of the eight points overlap instead of adding up behind five branches per point
        int cell[K1_UNROLL];
        c2g_cellkey key[K1_UNROLL], cur[K1_UNROLL];
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u) point_key_unit(v[u], (uint32_t) (i + u * K1_THREADS), P, cell[u], key[u]);
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u) cur[u] = *(volatile c2g_cellkey *) (tile + cell[u]);
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u)
          if (key[u] > cur[u]) atomicMax(tile + cell[u], key[u]);
      } else {
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u) scatter_point<UNIT>(v[u], (uint32_t) (i + u * K1_THREADS), P, tile);
      }
    }
    for (; i < n; i += K1_THREADS) scatter_point<UNIT>(ld_stream_f4(p + i), (uint32_t) i, P, tile);
    __syncthreads();
    // write the finished tile (coalesced 8 B / thread) and reset it for the next scan in the same pass
    c2g_cellkey *out = tiles_out + (size_t) b * ncell;
    for (int c = tid; c < ncell; c += K1_THREADS) {
      out[c] = tile[c];
      tile[c] = 0ull;
    }
    __syncthreads();
  }
}

}  // namespace

// host launcher (called from c2g_api.cu)
int c2g_launch_bev_scatter(const float *pts_dev, const long long *offsets_dev, int B, const C2gIngestParams &P,
                           c2g_cellkey *tiles_dev, int num_sms, cudaStream_t stream) {
  const size_t smem = (size_t) P.n_cells * sizeof(c2g_cellkey);
  static unsigned long long attr_devs = 0ull;
  if (c2g_first_use_on_device(attr_devs)) {
    C2G_CUDA_TRY(cudaFuncSetAttribute(bev_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (C2G_MAX_CELLS * sizeof(c2g_cellkey))));
    C2G_CUDA_TRY(cudaFuncSetAttribute(bev_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (C2G_MAX_CELLS * sizeof(c2g_cellkey))));
  }
  const int grid = B < num_sms ? B : num_sms;
  if (grid <= 0) return 0;
  // the fast path needs: unit resolution, symmetric padded bounds, and bounds that keep floor(x) + n/2 inside the image
  const bool unit = P.cfg.reso_row == 1.0f && P.cfg.reso_col == 1.0f && P.x_min_pad == -P.x_max_pad && P.y_min_pad == -P.y_max_pad &&
                    P.x_max_pad < (float) P.half_row && P.y_max_pad < (float) P.half_col;
  if (unit)
    bev_scatter_kernel<true><<<grid, K1_THREADS, smem, stream>>>((const float4 *) pts_dev, offsets_dev, B, P, tiles_dev);
  else
    bev_scatter_kernel<false><<<grid, K1_THREADS, smem, stream>>>((const float4 *) pts_dev, offsets_dev, B, P, tiles_dev);
  C2G_CUDA_TRY(cudaGetLastError());
  return 0;
}
