// bev_scatter.cu — K1: point cloud -> max-height BEV (ContourManager::makeBEV, include/cont2/contour_mng.h:505-556,
// with hashPointToImage :448-463), handed to the contour kernel as what it consumes: one bit-plane per height level
// (`bev > lv_grads[l]`, the cv::threshold of makeContourRecursiveHelper, src/cont2/contour_mng.cpp:283) and the compact,
// raster-ordered list of the cells above the lowest threshold with their height and the winner's continuous coordinates
// (bev_pixfs_, contour_mng.h:435,528-529) - the only cells contours, moments and keys ever read.
//
// Two kernels: bev_scatter_fast_kernel (below; 32-bit tile + event log, what a batch normally runs) and the general 64-bit
// kernel described here, which takes the scans the fast one defers and serves the full-tile getters.
//
// B200 mapping: one persistent CTA per SM, scans handed out dynamically.  The whole 150x150 BEV lives in shared memory as
// 64-bit cell keys (180 KB of the 227 KB carve-out), so the only HBM traffic is the streaming read of the points
// (16 B/point, 128-bit ld.global.nc.L1::no_allocate, UNROLL independent loads in flight per thread) plus ~40 KB of output
// per scan (round 1 wrote the 180 KB tile and the contour kernel read it back).  The reference's "if (bev < h) bev = h"
// with first-point-wins ties becomes a 64-bit max over (orderable(h) << 32 | ~index); a plain 32-bit shared-memory read of
// the cell's height word filters the points that cannot raise their cell any more before the atomic is issued (max is
// monotone, so a stale read can only cause a redundant atomic, never a missed one).
//
// Measured dead end (round 2): routing the points at or below the lowest threshold (2/3 of a
// scan: the ground) to a 2.8 KB occupancy bitmap with native 32-bit atomicOr instead of the 64-bit compare-and-swap loop.  A
// warp still executes the CAS block of every unrolled point whenever ANY of its lanes needs it (always), so the split only
// adds the second path: 0.87 ms instead of 0.84 ms per 1 184 scans.  The template parameter is kept for that measurement.
#include <cstdlib>

#include "c2g_common.cuh"

namespace {

constexpr int K1_THREADS = 1024;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_UNROLL = 8;
constexpr int K1_STASH = 2048;  // foreground cells staged between the epilogue passes (12 B each)
constexpr unsigned FULLMASK = 0xFFFFFFFFu;
constexpr size_t K1F_SMEM_MAX = 226 * 1024;  // dynamic shared memory the fast kernel opts in to (227 KB per CTA on sm_100, minus its static part)

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// XYZ layout (12 B per point, the separately reported input variant of SURVEY.md 8f-4: the intensity the path never reads is
// dropped on the host): three 32-bit loads per point; consecutive lanes read consecutive points, so the three loads of a warp
// cover the same 384 contiguous bytes and the second and third hit in L1.
template <bool XYZ>
__device__ __forceinline__ float4 ld_point(const float *base, long long i) {
  if (XYZ) {
    const float *q = base + 3 * i;
    return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.0f);
  }
  return ld_stream_f4(reinterpret_cast<const float4 *>(base) + i);
}

// What one point asks of the tile: `key` != 0: raise cell `ix` to `key`; `obit` != 0: mark bit `obit` of occupancy word `ix`
// (a point is one or the other, a rejected point neither).
struct PointOp {
  int ix;
  c2g_cellkey key;
  uint32_t obit;
};

// hashPointToImage + the height of makeBEV.  UNIT = true: reso_row == reso_col == 1.0f and the padded square is symmetric (the
// reference's only shipped setting, config/batch_bin_test_config.yaml:33-37 "TODO: reso other than 1.0"): x / 1.0f == x
// exactly, |x| <= x_max_pad replaces the two-sided test (and rejects NaN in the same compare), and the padded bounds already
// guarantee 0 <= row < n_row, 0 <= col < n_col, so only the reference's `row > 0` test remains (contour_mng.h:515).
// Branch-free: a rejected point yields key == 0 (never wins a max) and obit == 0.
template <bool UNIT, bool LOWFILTER>
__device__ __forceinline__ PointOp point_op(const float4 pt, const uint32_t idx, const C2gIngestParams &P, const float lv_min, const int wpr) {
  const float x = pt.x, y = pt.y;
  bool ok;
  int row, col;
  if (UNIT) {
    ok = (fabsf(x) <= P.x_max_pad) && (fabsf(y) <= P.y_max_pad);
    ok = ok && !(__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq);
    row = __float2int_rd(x) + P.half_row;
    col = __float2int_rd(y) + P.half_col;
    ok = ok && row > 0;  // `rc.first > 0` (contour_mng.h:515)
  } else {
    // reject outside the padded square or inside the blind radius; NaN x / y fail `==` and are dropped
    ok = !(x < P.x_min_pad || x > P.x_max_pad || y < P.y_min_pad || y > P.y_max_pad) && (x == x) && (y == y);
    ok = ok && !(__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq);
    row = ok ? (int) floorf(__fdiv_rn(x, P.cfg.reso_row)) + P.half_row : 0;
    col = ok ? (int) floorf(__fdiv_rn(y, P.cfg.reso_col)) + P.half_col : 0;
    ok = ok && row > 0 && row < P.cfg.n_row && col >= 0 && col < P.cfg.n_col;
  }
  const float h = __fadd_rn(P.cfg.lidar_height, pt.z);
  ok = ok && (h > -1000.0f);  // bev_ starts at -1000 and only strictly higher points are stored (NaN z: never)
  const bool hi = ok && (!LOWFILTER || h > lv_min);
  const bool lo = ok && !hi;
  PointOp op;
  op.ix = hi ? row * P.cfg.n_col + col : (lo ? row * wpr + (col >> 5) : 0);
  op.key = hi ? (((c2g_cellkey) c2g_orderable(h) << 32) | (c2g_cellkey) (0xFFFFFFFFu - idx)) : 0ull;
  op.obit = lo ? (1u << (col & 31)) : 0u;
  return op;
}

template <bool LOWFILTER>
__device__ __forceinline__ void apply_op(const PointOp &op, c2g_cellkey *tile, uint32_t *occ) {
  if (op.key) {
    if (op.key > *(volatile c2g_cellkey *) (tile + op.ix)) atomicMax(tile + op.ix, op.key);
  } else if (LOWFILTER && op.obit) {
    if (op.obit & ~*(volatile uint32_t *) (occ + op.ix)) atomicOr(occ + op.ix, op.obit);
  }
}

template <bool UNIT, bool LOWFILTER, bool XYZ>
__global__ void __launch_bounds__(K1_THREADS, 1)
bev_scatter_kernel(const float *__restrict__ pts, const long long *__restrict__ offsets, int B, C2gIngestParams P, C2gBevOut out,
                   int *__restrict__ work_counter, int variant, const int *__restrict__ scan_list, const int *__restrict__ scan_cnt) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  // list mode: the scans the fast kernel deferred (none, normally: leave before touching the tile)
  if (scan_list) {
    B = *scan_cnt;
    if (B == 0) return;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = P.n_cells, ncol = P.cfg.n_col, nrow = P.cfg.n_row;
  const int wpr = (ncol + 31) >> 5, nwords = nrow * wpr;
  c2g_cellkey *tile = reinterpret_cast<c2g_cellkey *>(k1_smem);                     // [ncell]
  uint32_t *occ = reinterpret_cast<uint32_t *>(tile + ((ncell + 1) & ~1));           // [nwords] occupancy of the low cells
  uint32_t *plane = occ + nwords;                                                    // [NLEV][nwords] staged for a coalesced write
  uint16_t *wpre = reinterpret_cast<uint16_t *>(plane + C2G_NLEV * nwords);          // [nwords] fg cells before each word
  uint16_t *stash_loc = wpre + ((nwords + 3) & ~3);                                  // [K1_STASH] plane word << 5 | bit
  c2g_cellkey *stash_key = reinterpret_cast<c2g_cellkey *>(stash_loc + K1_STASH);    // [K1_STASH]
  __shared__ int s_next, s_warp_occ[K1_WARPS], s_wstash[K1_WARPS], s_nfg, s_iter, s_nstash;
  static_assert((K1_STASH / K1_WARPS & (K1_STASH / K1_WARPS - 1)) == 0, "stash share per warp must be a power of two");
  if (threadIdx.x == 0) s_iter = 0;
  float lv_min = P.cfg.lv_grads[0];
#pragma unroll
  for (int e = 1; e < C2G_NLEV; ++e) lv_min = fminf(lv_min, P.cfg.lv_grads[e]);
  for (int c = tid; c < ncell; c += K1_THREADS) tile[c] = 0ull;
  for (int w = tid; w < nwords; w += K1_THREADS) occ[w] = 0u;
  // scans are handed out dynamically: a CTA that starts late (behind a co-running collective or the previous kernel's tail)
  // must not leave a static share of the batch unprocessed until the end
  while (true) {
    if (tid == 0) s_next = (variant & 2) ? (s_iter++ * (int) gridDim.x + (int) blockIdx.x) : atomicAdd(work_counter, 1);
    __syncthreads();
    if (s_next >= B) break;
    const int b = scan_list ? scan_list[s_next] : s_next;
    if (tid == 0) s_nstash = 0;  // (ordered before its uses by the barrier that ends the main loop)
    const long long beg = offsets[b];
    const int n = (int) (offsets[b + 1] - beg);
    const float *p = pts + (XYZ ? 3 : 4) * beg;
    int i = tid;
    // main loop: UNROLL independent 128-bit loads per thread before any of them is consumed; then the cell / key of all UNROLL
    // points (straight-line code), then all filter reads of the tile, then the atomics: the shared-memory latencies of the
    // eight points overlap instead of adding up behind five branches per point
    for (; i + (K1_UNROLL - 1) * K1_THREADS < n; i += K1_UNROLL * K1_THREADS) {
      float4 v[K1_UNROLL];
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) v[u] = ld_point<XYZ>(p, i + u * K1_THREADS);
      PointOp op[K1_UNROLL];
      uint32_t cur[K1_UNROLL];  // height word of a high point's cell (upper half of the 64-bit key) / occupancy word of a low point
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) op[u] = point_op<UNIT, LOWFILTER>(v[u], (uint32_t) (i + u * K1_THREADS), P, lv_min, wpr);
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) {
        // one unconditional 32-bit read per point (a rejected point reads word 1 of the tile and ignores it)
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tile) + 2 * op[u].ix + 1;
        if (LOWFILTER && op[u].obit) src = occ + op[u].ix;
        cur[u] = *(volatile const uint32_t *) src;
      }
      if (variant & 1) {  // experiment: the 64-bit filter read of round 1
        c2g_cellkey c64[K1_UNROLL];
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u) c64[u] = *(volatile c2g_cellkey *) (tile + op[u].ix);
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u)
          if (op[u].key > c64[u]) atomicMax(tile + op[u].ix, op[u].key);
        continue;
      }
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) {
        // equal heights go to the atomic too: the lower point index wins there.  key == 0 for low / rejected points.
        if (op[u].key && (uint32_t) (op[u].key >> 32) >= cur[u]) atomicMax(tile + op[u].ix, op[u].key);
        if (LOWFILTER && (op[u].obit & ~cur[u])) atomicOr(occ + op[u].ix, op[u].obit);  // obit == 0 for high points
      }
    }
    for (; i < n; i += K1_THREADS) apply_op<LOWFILTER>(point_op<UNIT, LOWFILTER>(ld_point<XYZ>(p, i), (uint32_t) i, P, lv_min, wpr), tile, occ);
    __syncthreads();
    if (variant & 4) {  // experiment: main loop only (outputs are garbage)
      for (int c = tid; c < ncell; c += K1_THREADS) tile[c] = 0ull;
      __syncthreads();
      continue;
    }

    // ---- epilogue 1: occupancy count, reset of the tile, bit-planes.  The cells above the lowest threshold (~5 % of the image)
    // are moved to a small stash (key + plane word + bit; every warp fills its own share, positions from a ballot: no shared counter) so that the gather pass below is a
    // dense, balanced loop over ~1 100 entries instead of a second walk over 22 500 cells with a few long-latency loads per
    // thread; cells that do not fit the stash stay in the tile for a fall-back walk.
    int occ_cnt = 0;
    {
      // one warp per 32-column plane word (lane = column): the plane bits are ballots, no shared atomics; word w advances by
      // K1_WARPS words per step without a division
      const int dr = K1_WARPS / wpr, dw = K1_WARPS - dr * wpr;
      int row = warp / wpr, wi = warp - row * wpr;
      c2g_cellkey *t_out = (!LOWFILTER && out.tiles) ? out.tiles + (size_t) b * ncell : nullptr;
      int wcount = 0;  // stash slots this warp has used (warp-uniform): every warp owns K1_STASH / K1_WARPS slots, no shared counter
      for (int w = warp; w < nwords; w += K1_WARPS) {
        const int col = wi * 32 + lane, c = row * ncol + col;
        const bool in = col < ncol;
        const c2g_cellkey k = in ? tile[c] : 0ull;
        if (t_out && in) t_out[c] = k;
        const float h = k != 0ull ? c2g_from_orderable((uint32_t) (k >> 32)) : -1000.0f;
        const bool isfg = k != 0ull && h > lv_min;
        const unsigned occ_m = __ballot_sync(FULLMASK, k != 0ull);
        const unsigned m = __ballot_sync(FULLMASK, isfg);  // lv_grads increase (checked at c2g_create): plane 0 = all foreground cells
        uint32_t mine = 0;
        if (m) {  // most words hold no cell above the lowest threshold
#pragma unroll
          for (int e = 1; e < C2G_NLEV; ++e) {
            const uint32_t bal = __ballot_sync(FULLMASK, isfg && h > P.cfg.lv_grads[e]);
            if (lane == e) mine = bal;
          }
          if (lane == 0) mine = m;
        }
        if (lane < C2G_NLEV) plane[lane * nwords + w] = mine;
        if (lane == 0) occ_cnt += __popc(occ_m | (LOWFILTER ? occ[w] : 0u));
        bool keep = false;
        if (isfg) {
          const int pos = wcount + __popc(m & ((1u << lane) - 1u));
          if (pos < K1_STASH / K1_WARPS) {
            stash_key[warp * (K1_STASH / K1_WARPS) + pos] = k;
            stash_loc[warp * (K1_STASH / K1_WARPS) + pos] = (uint16_t) ((w << 5) | lane);
          } else
            keep = true;  // stays in the tile for the fall-back walk
        }
        wcount += __popc(m);
        if (k != 0ull && !keep) tile[c] = 0ull;
        row += dr;
        wi += dw;
        if (wi >= wpr) {
          wi -= wpr;
          ++row;
        }
      }
      if (lane == 0) {
        s_wstash[warp] = wcount;
        if (wcount > K1_STASH / K1_WARPS) s_nstash = 1;  // some cells stayed behind
      }
    }
    if (lane == 0) s_warp_occ[warp] = occ_cnt;
    __syncthreads();
    // ---- epilogue 2: exclusive prefix of the per-word foreground counts (raster order = the order moments are accumulated in);
    // lv_grads increase (checked at c2g_create), so plane 0 holds every foreground cell
    if (warp == 0 && !(variant & 16)) {
      int base = 0;
      for (int w0 = 0; w0 < nwords; w0 += 32) {
        const int w = w0 + lane;
        const int cnt = w < nwords ? __popc(plane[w]) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULLMASK, incl, o);
          if (lane >= o) incl += t;
        }
        if (w < nwords) wpre[w] = (uint16_t) (base + incl - cnt);
        base += __shfl_sync(FULLMASK, incl, 31);
      }
      if (lane == 0) s_nfg = base;
    }
    __syncthreads();
    // ---- epilogue 3: foreground records (height + the winner's continuous coordinates: an 8-byte gather from the points just
    // streamed) at their raster rank; coalesced write of the planes
    {
      float4 *fg = out.fg + (size_t) b * ncell;
      auto emit = [&](c2g_cellkey k, int wd, int bit) {
        const float *q = p + (size_t) (XYZ ? 3 : 4) * (0xFFFFFFFFu - (uint32_t) k);
        const float2 xy = make_float2(__ldg(q), __ldg(q + 1));
        float4 rec;
        rec.x = c2g_from_orderable((uint32_t) (k >> 32));
        // pointToContRowCol (contour_mng.h:468-472): x / reso + n_row / 2 - 0.5f, left to right in float
        rec.y = __fsub_rn(__fadd_rn(__fdiv_rn(xy.x, P.cfg.reso_row), P.half_row_f), 0.5f);
        rec.z = __fsub_rn(__fadd_rn(__fdiv_rn(xy.y, P.cfg.reso_col), P.half_col_f), 0.5f);
        rec.w = 0.0f;
        fg[(int) wpre[wd] + __popc(plane[wd] & ((1u << bit) - 1u))] = rec;
      };
      for (int e = tid; e < K1_STASH && !(variant & 8); e += K1_THREADS)
        if ((e & (K1_STASH / K1_WARPS - 1)) < s_wstash[e / (K1_STASH / K1_WARPS)])
          emit(stash_key[e], (int) (stash_loc[e] >> 5), (int) (stash_loc[e] & 31));
      if (s_nstash) {  // cluttered scan: the cells that did not fit a warp's stash share are still in the tile
        const int dr = K1_THREADS / ncol, dc = K1_THREADS - dr * ncol;
        int row = tid / ncol, col = tid - row * ncol;
        for (int c = tid; c < ncell; c += K1_THREADS) {
          const c2g_cellkey k = tile[c];
          if (k != 0ull) {
            tile[c] = 0ull;
            emit(k, row * wpr + (col >> 5), col & 31);
          }
          row += dr;
          col += dc;
          if (col >= ncol) {
            col -= ncol;
            ++row;
          }
        }
      }
      uint32_t *pl_out = out.planes + (size_t) b * C2G_NLEV * nwords;
      for (int j = tid; j < C2G_NLEV * nwords && !(variant & 32); j += K1_THREADS) pl_out[j] = plane[j];
      if (LOWFILTER)
        for (int w = tid; w < nwords; w += K1_THREADS) occ[w] = 0u;
      if (tid == 0) {
        int tot = 0;
        for (int wq = 0; wq < K1_WARPS; ++wq) tot += s_warp_occ[wq];
        out.hdr[b] = make_int2(tot, s_nfg);
      }
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------------------------------------
// The fast scatter kernel: 32-bit tile + event log.
//
// The 64-bit key of the kernel above exists only to settle "which point won the cell" (first point wins ties) together with
// the height, and shared memory has no native 64-bit max: every point that raises its cell (27 % of a scan) runs a
// compare-and-swap loop (47 % of that kernel's stall samples).  But the winner is read for the ~1 100 cells above the lowest
// threshold only (their continuous coordinates; bev_pixfs_).  So the tile keeps the orderable height alone - a native
// ATOMS.MAX.U32 - and a point above the lowest threshold whose atomic reports `old <= mine` (it raised the cell, or tied
// with its current maximum) appends one 8-byte event (height, cell, point index) to a log in shared memory: ~5 events per
// foreground cell (a record-breaking sequence of n values has ~ln n records).  After the scan, the events whose height
// equals the cell's final height are exactly the points that reached the maximum; the smallest index among them is the
// reference's winner.  The tile shrinks to 90 KB, which pays for the log.
//
// A scan that does not fit the scheme - more than 2^17 points (index field), more than K1F_EV_CAP events or K1F_FG_CAP
// foreground cells (cluttered BEVs) - is put on a list and processed by the 64-bit kernel in list mode right after.
constexpr int K1F_EV_CAP = 12288;
constexpr int K1F_FG_CAP = 4096;
constexpr int K1F_MAX_PTS = 1 << 17;

// hashPointToImage + the height of makeBEV (see point_op): cell index and orderable height, 0 = rejected point
template <bool UNIT>
__device__ __forceinline__ void point_cell(const float4 pt, const C2gIngestParams &P, int &ix, uint32_t &oh) {
  const float x = pt.x, y = pt.y;
  bool ok;
  int row, col;
  if (UNIT) {
    ok = (fabsf(x) <= P.x_max_pad) && (fabsf(y) <= P.y_max_pad);
    ok = ok && !(__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq);
    row = __float2int_rd(x) + P.half_row;
    col = __float2int_rd(y) + P.half_col;
    ok = ok && row > 0;  // `rc.first > 0` (contour_mng.h:515)
  } else {
    ok = !(x < P.x_min_pad || x > P.x_max_pad || y < P.y_min_pad || y > P.y_max_pad) && (x == x) && (y == y);
    ok = ok && !(__fadd_rn(__fmul_rn(y, y), __fmul_rn(x, x)) < P.cfg.blind_sq);
    row = ok ? (int) floorf(__fdiv_rn(x, P.cfg.reso_row)) + P.half_row : 0;
    col = ok ? (int) floorf(__fdiv_rn(y, P.cfg.reso_col)) + P.half_col : 0;
    ok = ok && row > 0 && row < P.cfg.n_row && col >= 0 && col < P.cfg.n_col;
  }
  const float h = __fadd_rn(P.cfg.lidar_height, pt.z);
  ok = ok && (h > -1000.0f);  // bev_ starts at -1000 and only strictly higher points are stored (NaN z: never)
  ix = ok ? row * P.cfg.n_col + col : 0;
  oh = ok ? c2g_orderable(h) : 0u;
}

__device__ __forceinline__ unsigned long long k1f_pack(uint32_t oh, int field, uint32_t idx) {
  return ((unsigned long long) oh << 32) | ((unsigned long long) field << 17) | (unsigned long long) idx;
}

template <bool UNIT, bool XYZ>
__global__ void __launch_bounds__(K1_THREADS, 1)
bev_scatter_fast_kernel(const float *__restrict__ pts, const long long *__restrict__ offsets, int B, C2gIngestParams P, C2gBevOut out,
                        int *__restrict__ counters) {  // [0] next scan, [2] number of deferred scans, [3..] their indices
  extern __shared__ __align__(16) unsigned char k1_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = P.n_cells, ncol = P.cfg.n_col, nrow = P.cfg.n_row;
  const int wpr = (ncol + 31) >> 5, nwords = nrow * wpr, ncell4 = (ncell + 3) >> 2;
  uint32_t *htile = reinterpret_cast<uint32_t *>(k1_smem);                                    // [ncell4 * 4] orderable height, 0 = empty
  unsigned long long *events = reinterpret_cast<unsigned long long *>(htile + 4 * ncell4);    // [K1F_EV_CAP]
  uint32_t *plane = reinterpret_cast<uint32_t *>(events + K1F_EV_CAP);                        // [NLEV][nwords] staged for a coalesced write
  uint32_t *win = plane + C2G_NLEV * nwords;                                                  // [K1F_FG_CAP] smallest index that reached the maximum
  uint16_t *wpre = reinterpret_cast<uint16_t *>(win + K1F_FG_CAP);                            // [nwords] fg cells before each word
  __shared__ int s_next, s_warp_occ[K1_WARPS], s_nfg, s_nev;
  float lv_min = P.cfg.lv_grads[0];
#pragma unroll
  for (int e = 1; e < C2G_NLEV; ++e) lv_min = fminf(lv_min, P.cfg.lv_grads[e]);
  const uint32_t thr_o = c2g_orderable(lv_min);  // h > lv_min <=> orderable(h) > thr_o
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int c = tid; c < ncell4; c += K1_THREADS) reinterpret_cast<uint4 *>(htile)[c] = zero4;
  auto defer = [&](int b) {  // thread 0: hand the scan to the 64-bit kernel
    const int slot = atomicAdd(counters + 2, 1);
    counters[3 + slot] = b;
  };
  while (true) {
    if (tid == 0) {
      s_next = atomicAdd(counters, 1);
      s_nev = 0;
    }
    __syncthreads();
    const int b = s_next;
    if (b >= B) break;
    const long long beg = offsets[b];
    const int n = (int) (offsets[b + 1] - beg);
    if (offsets[b + 1] - beg > (long long) K1F_MAX_PTS) {  // the index field of an event has 17 bits
      if (tid == 0) defer(b);
      __syncthreads();  // everybody has read s_next
      continue;
    }
    const float *p = pts + (XYZ ? 3 : 4) * beg;
    int i = tid;
    // main loop: UNROLL independent loads per thread in flight, then cells / heights, then the filter reads (a point that cannot
    // raise its cell any more issues no atomic: max is monotone, a stale read only causes a redundant atomic), then the atomics.
    // The trip count is the same for the lanes of a warp (the log is appended warp-wide).
    for (; (i - lane) + 31 + (K1_UNROLL - 1) * K1_THREADS < n; i += K1_UNROLL * K1_THREADS) {
      float4 v[K1_UNROLL];
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) v[u] = ld_point<XYZ>(p, i + u * K1_THREADS);
      int ix[K1_UNROLL];
      uint32_t oh[K1_UNROLL], cur[K1_UNROLL];
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) point_cell<UNIT>(v[u], P, ix[u], oh[u]);
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) cur[u] = *(volatile const uint32_t *) (htile + ix[u]);
      uint32_t evm = 0u;  // bit u: point u goes to the log
      // all atomics are issued before the first return value is looked at (cur[u] becomes the value the atomic found, or
      // 'nothing to log' for a point that issued none); equal heights go to the atomic too: a tie is an event
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u) cur[u] = (oh[u] != 0u && oh[u] >= cur[u]) ? atomicMax(htile + ix[u], oh[u]) : 0xFFFFFFFFu;
#pragma unroll
      for (int u = 0; u < K1_UNROLL; ++u)
        if (oh[u] > thr_o && cur[u] <= oh[u]) evm |= 1u << u;
      const int c = __popc(evm);
      if (__any_sync(FULLMASK, c != 0)) {
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULLMASK, incl, o);
          if (lane >= o) incl += t;
        }
        int base = 0;
        if (lane == 31) base = atomicAdd(&s_nev, incl);
        base = __shfl_sync(FULLMASK, base, 31);
        int pos = base + incl - c;
#pragma unroll
        for (int u = 0; u < K1_UNROLL; ++u)
          if ((evm >> u) & 1u) {
            if (pos < K1F_EV_CAP) events[pos] = k1f_pack(oh[u], ix[u], (uint32_t) (i + u * K1_THREADS));
            ++pos;
          }
      }
    }
    for (; i < n; i += K1_THREADS) {
      int ix1;
      uint32_t oh1;
      point_cell<UNIT>(ld_point<XYZ>(p, i), P, ix1, oh1);
      if (oh1 != 0u && oh1 >= *(volatile const uint32_t *) (htile + ix1)) {
        const uint32_t old = atomicMax(htile + ix1, oh1);
        if (oh1 > thr_o && old <= oh1) {
          const int pos = atomicAdd(&s_nev, 1);
          if (pos < K1F_EV_CAP) events[pos] = k1f_pack(oh1, ix1, (uint32_t) i);
        }
      }
    }
    __syncthreads();
    const int nev = s_nev;
    if (nev > K1F_EV_CAP) {  // the log overflowed: start over with the 64-bit kernel
      for (int c = tid; c < ncell4; c += K1_THREADS) reinterpret_cast<uint4 *>(htile)[c] = zero4;
      if (tid == 0) defer(b);
      __syncthreads();
      continue;
    }
    // ---- epilogue 1: one warp per 32-column plane word (lane = column): plane bits and the occupied count from ballots
    int occ_cnt = 0;
    {
      const int dr = K1_WARPS / wpr, dw = K1_WARPS - dr * wpr;
      int row = warp / wpr, wi = warp - row * wpr;
      for (int w = warp; w < nwords; w += K1_WARPS) {
        const int col = wi * 32 + lane;
        const uint32_t k = col < ncol ? htile[row * ncol + col] : 0u;
        const bool isfg = k > thr_o;
        const unsigned occ_m = __ballot_sync(FULLMASK, k != 0u);
        const unsigned m = __ballot_sync(FULLMASK, isfg);  // lv_grads increase (checked at c2g_create): plane 0 = all foreground cells
        uint32_t mine = 0;
        if (m) {  // most words hold no cell above the lowest threshold
          const float h = c2g_from_orderable(k);
#pragma unroll
          for (int e = 1; e < C2G_NLEV; ++e) {
            const uint32_t bal = __ballot_sync(FULLMASK, isfg && h > P.cfg.lv_grads[e]);
            if (lane == e) mine = bal;
          }
          if (lane == 0) mine = m;
        }
        if (lane < C2G_NLEV) plane[lane * nwords + w] = mine;
        if (lane == 0) occ_cnt += __popc(occ_m);
        row += dr;
        wi += dw;
        if (wi >= wpr) {
          wi -= wpr;
          ++row;
        }
      }
      for (int j = tid; j < K1F_FG_CAP; j += K1_THREADS) win[j] = 0xFFFFFFFFu;
    }
    if (lane == 0) s_warp_occ[warp] = occ_cnt;
    __syncthreads();
    // ---- epilogue 2: exclusive prefix of the per-word foreground counts (raster order = the order moments are accumulated in)
    if (warp == 0) {
      int base = 0;
      for (int w0 = 0; w0 < nwords; w0 += 32) {
        const int w = w0 + lane;
        const int cnt = w < nwords ? __popc(plane[w]) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULLMASK, incl, o);
          if (lane >= o) incl += t;
        }
        if (w < nwords) wpre[w] = (uint16_t) (base + incl - cnt);
        base += __shfl_sync(FULLMASK, incl, 31);
      }
      if (lane == 0) s_nfg = base;
    }
    __syncthreads();
    if (s_nfg > K1F_FG_CAP) {  // cluttered scan
      for (int c = tid; c < ncell4; c += K1_THREADS) reinterpret_cast<uint4 *>(htile)[c] = zero4;
      if (tid == 0) defer(b);
      __syncthreads();
      continue;
    }
    // ---- epilogue 3a: the events that reached their cell's final height compete for the cell with their point index
    for (int e = tid; e < nev; e += K1_THREADS) {
      const unsigned long long ev = events[e];
      const uint32_t oh = (uint32_t) (ev >> 32), idx = (uint32_t) ev & 0x1FFFFu;
      const int cell = (int) ((uint32_t) ev >> 17);
      unsigned long long keep = 0ull;
      if (htile[cell] == oh) {
        const int row = cell / ncol, col = cell - row * ncol, wd = row * wpr + (col >> 5);
        const int rank = (int) wpre[wd] + __popc(plane[wd] & ((1u << (col & 31)) - 1u));
        atomicMin(win + rank, idx);
        keep = k1f_pack(oh, rank, idx);  // the cell field now holds the raster rank
      }
      events[e] = keep;
    }
    __syncthreads();
    // ---- epilogue 3b: the winners write their record (height + continuous coordinates of the point: an 8-byte gather from the
    // points just streamed) at the cell's raster rank; the tile is cleared, the planes leave with a coalesced write
    {
      float4 *fg = out.fg + (size_t) b * ncell;
      for (int e = tid; e < nev; e += K1_THREADS) {
        const unsigned long long ev = events[e];
        if (ev == 0ull) continue;
        const uint32_t idx = (uint32_t) ev & 0x1FFFFu;
        const int rank = (int) ((uint32_t) ev >> 17);
        if (win[rank] != idx) continue;
        const float *q = p + (size_t) (XYZ ? 3 : 4) * idx;
        const float2 xy = make_float2(__ldg(q), __ldg(q + 1));
        float4 rec;
        rec.x = c2g_from_orderable((uint32_t) (ev >> 32));
        // pointToContRowCol (contour_mng.h:468-472): x / reso + n_row / 2 - 0.5f, left to right in float
        rec.y = __fsub_rn(__fadd_rn(__fdiv_rn(xy.x, P.cfg.reso_row), P.half_row_f), 0.5f);
        rec.z = __fsub_rn(__fadd_rn(__fdiv_rn(xy.y, P.cfg.reso_col), P.half_col_f), 0.5f);
        rec.w = 0.0f;
        fg[rank] = rec;
      }
      for (int c = tid; c < ncell4; c += K1_THREADS) reinterpret_cast<uint4 *>(htile)[c] = zero4;
      uint32_t *pl_out = out.planes + (size_t) b * C2G_NLEV * nwords;
      for (int j = tid; j < C2G_NLEV * nwords; j += K1_THREADS) pl_out[j] = plane[j];
      if (tid == 0) {
        int tot = 0;
        for (int wq = 0; wq < K1_WARPS; ++wq) tot += s_warp_occ[wq];
        out.hdr[b] = make_int2(tot, s_nfg);
      }
    }
    __syncthreads();
  }
}

size_t k1f_smem_bytes(int ncell, int nwords) {
  return (size_t) ((ncell + 3) & ~3) * 4 + (size_t) K1F_EV_CAP * 8 + (size_t) nwords * 4 * C2G_NLEV + (size_t) K1F_FG_CAP * 4 +
         (size_t) ((nwords + 3) & ~3) * 2 + 16;
}

template <bool UNIT, bool XYZ>
int launch_fast(const float *pts, const long long *offsets, int B, const C2gIngestParams &P, const C2gBevOut &out, int *counters, int grid, size_t smem,
                cudaStream_t stream) {
  static unsigned long long attr_devs = 0ull;
  if (c2g_first_use_on_device(attr_devs))
    C2G_CUDA_TRY(cudaFuncSetAttribute(bev_scatter_fast_kernel<UNIT, XYZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) K1F_SMEM_MAX));
  bev_scatter_fast_kernel<UNIT, XYZ><<<grid, K1_THREADS, smem, stream>>>(pts, offsets, B, P, out, counters);
  return 0;
}

size_t k1_smem_bytes(int ncell, int nwords) {
  return (size_t) ((ncell + 1) & ~1) * sizeof(c2g_cellkey) + (size_t) nwords * 4 * (1 + C2G_NLEV) + (size_t) ((nwords + 3) & ~3) * 2 +
         (size_t) K1_STASH * (2 + sizeof(c2g_cellkey)) + 16;
}

template <bool UNIT, bool LOWFILTER, bool XYZ>
int launch_variant(const float *pts, const long long *offsets, int B, const C2gIngestParams &P, const C2gBevOut &out, int *work_counter, int grid,
                   size_t smem, cudaStream_t stream, const int *scan_list = nullptr, const int *scan_cnt = nullptr) {
  static unsigned long long attr_devs = 0ull;
  if (c2g_first_use_on_device(attr_devs))
    C2G_CUDA_TRY(cudaFuncSetAttribute(bev_scatter_kernel<UNIT, LOWFILTER, XYZ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) k1_smem_bytes(C2G_MAX_CELLS, 800)));
  static const int variant = getenv("C2G_K1_VARIANT") ? atoi(getenv("C2G_K1_VARIANT")) : 0;  // measurement hook
  bev_scatter_kernel<UNIT, LOWFILTER, XYZ><<<grid, K1_THREADS, smem, stream>>>(pts, offsets, B, P, out, work_counter, variant, scan_list, scan_cnt);
  return 0;
}

}  // namespace

// host launcher (called from c2g_api.cu).  full_tile != 0: the complete 64-bit tile is written to out.tiles as well (dense-image
// getters); otherwise only the bit-planes, the foreground list and the counts.
int c2g_launch_bev_scatter(const float *pts_dev, const long long *offsets_dev, int B, const C2gIngestParams &P, const C2gBevOut &out,
                           int full_tile, int xyz, int *work_counter, int num_sms, cudaStream_t stream) {
  const int nwords = P.cfg.n_row * ((P.cfg.n_col + 31) / 32);
  const size_t smem = k1_smem_bytes(P.n_cells, nwords);
  const int grid = B < num_sms ? B : num_sms;
  if (grid <= 0) return 0;
  // work_counter: [0] next scan of the running kernel, [1] the same for the list-mode launch, [2] number of deferred scans,
  // [3 .. 3 + B) their indices
  C2G_CUDA_TRY(cudaMemsetAsync(work_counter, 0, 3 * sizeof(int), stream));
  // the fast path needs: unit resolution, symmetric padded bounds, and bounds that keep floor(x) + n/2 inside the image
  const bool unit = P.cfg.reso_row == 1.0f && P.cfg.reso_col == 1.0f && P.x_min_pad == -P.x_max_pad && P.y_min_pad == -P.y_max_pad &&
                    P.x_max_pad < (float) P.half_row && P.y_max_pad < (float) P.half_col;
  static const int lowfilter = getenv("C2G_K1_LOWFILTER") ? atoi(getenv("C2G_K1_LOWFILTER")) : 0;  // measurement hook (header comment)
  static const int force64 = getenv("C2G_K1_64BIT") ? atoi(getenv("C2G_K1_64BIT")) : 0;             // measurement hook: the 64-bit kernel only
  C2gBevOut o = out;
  if (!full_tile) o.tiles = nullptr;  // production: bit-planes, foreground list and counts only
  int rc;
  const size_t smem_fast = k1f_smem_bytes(P.n_cells, nwords);
  if (!full_tile && !force64 && !lowfilter && smem_fast <= K1F_SMEM_MAX) {
    // 32-bit tile + event log; whatever it defers (oversized or cluttered scans) is done by the 64-bit kernel in list mode
#define C2G_K1_FAST(U, X) launch_fast<U, X>(pts_dev, offsets_dev, B, P, o, work_counter, grid, smem_fast, stream)
#define C2G_K1_LIST(U, X) launch_variant<U, false, X>(pts_dev, offsets_dev, B, P, o, work_counter + 1, grid, smem, stream, work_counter + 3, work_counter + 2)
    if (xyz)
      rc = unit ? C2G_K1_FAST(true, true) : C2G_K1_FAST(false, true);
    else
      rc = unit ? C2G_K1_FAST(true, false) : C2G_K1_FAST(false, false);
    if (rc) return rc;
    C2G_CUDA_TRY(cudaGetLastError());
    if (xyz)
      rc = unit ? C2G_K1_LIST(true, true) : C2G_K1_LIST(false, true);
    else
      rc = unit ? C2G_K1_LIST(true, false) : C2G_K1_LIST(false, false);
#undef C2G_K1_FAST
#undef C2G_K1_LIST
    if (rc) return rc;
    C2G_CUDA_TRY(cudaGetLastError());
    return 0;
  }
#define C2G_K1_GO(U, L, X) launch_variant<U, L, X>(pts_dev, offsets_dev, B, P, o, work_counter, grid, smem, stream)
  if (lowfilter && !full_tile && !xyz)
    rc = unit ? C2G_K1_GO(true, true, false) : C2G_K1_GO(false, true, false);
  else if (xyz)
    rc = unit ? C2G_K1_GO(true, false, true) : C2G_K1_GO(false, false, true);
  else
    rc = unit ? C2G_K1_GO(true, false, false) : C2G_K1_GO(false, false, false);
#undef C2G_K1_GO
  if (rc) return rc;
  C2G_CUDA_TRY(cudaGetLastError());
  return 0;
}
