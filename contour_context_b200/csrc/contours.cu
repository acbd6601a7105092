// contours.cu — K2: BEV tile -> multi-level contours -> ContourView statistics -> per-level order -> retrieval keys ->
// BCIs -> per-scan GMM terms.  One persistent CTA per SM, one scan per CTA iteration, everything between the BEV tile and
// the finished descriptor stays in shared memory / L2.
//
// Reference functions restated here (paths relative to the reference repo):
//   ContourManager::makeContourRecursiveHelper   src/cont2/contour_mng.cpp:274-353
//   RunningStatRecorder::runningStatsF           include/cont2/contour.h:74-84
//   ContourView::calcStatVals (+ salience tests) include/cont2/contour.h:142-265
//   ContourManager::makeContoursRecurs           include/cont2/contour_mng.h:588-608 (sort), :693-830 (keys), :848-888 (BCI)
//   GMMPair ctor (scan-only part)                include/cont2/correlation.h:49-82,102-119
//
// How the recursion becomes data-parallel: a level-(L+1) component is a connected component of {bev > lv[L+1]} and is
// contained in exactly one level-L component, so a GLOBAL 8-connected labelling per level finds the same pixel sets as
// the reference's recursive ROI-by-ROI labelling.  What the recursion adds is ORDER: cont_views_[L] is filled in DFS
// order, children of one parent in OpenCV label order, which is the block-raster order of each component's first 2x2
// block with blocks aligned to the parent's bounding-box origin (SURVEY.md §7 hard part 2).  Hence
//   order(level L) = sort by (rank of parent in level L-1, min over pixels of ((r-y0)>>1, (c-x0)>>1)).
// Per-pixel label words pack (rank of the enclosing level-(L-1) component) << 16 | union-find parent pixel, so one native
// 32-bit shared atomicMin implements the union (all words of one tree share the upper half).
//
// Bit-exactness rules: float/double sums that feed views and keys are accumulated in the reference's raster order by a
// single logical accumulator (warp-redundant), never by tree reductions; compiled with -fmad=false.
#include <math_constants.h>

#include "c2g_common.cuh"
#include "stdsort.cuh"
#include "c2g_libm.cuh"

namespace {

// glibc's __exp_data.tab (2 KB, read through L1); see c2g_libm.cuh
__device__ const uint64_t c2g_exp_tab_dev[256] = C2G_EXP_TAB_INIT;

constexpr int K2_THREADS = 1024;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int NC = 2048;        // components (any size) per level
constexpr int NVL = 1024;       // significant components (area >= min_cont_cell_cnt) per level
constexpr int KEY_LIST_CAP = 400;  // cells of one key window that can lie inside the 9.99-cell radius
constexpr int N_ANCH = C2G_NLEV * C2G_MAX_PIV;
constexpr int N_DIVS = 35;
constexpr int WL_CAP = 256;

struct TopView {  // what keys / BCI / GMM need from a sorted view
  float mean0, mean1, eig0, eig1;
  int cnt;
};

struct Smem {
  uint32_t L[C2G_MAX_CELLS];      // label words; reused as scratch after the level loop
  uint8_t msk[C2G_MAX_CELLS + 28];  // bit e: bev > lv_grads[e]
  int c_area[NC], c_minr[NC], c_minc[NC], c_maxr[NC], c_maxc[NC], c_key[NC], c_poi[NC];
  uint16_t c_rank[NC], c_pcid[NC];
  uint16_t sig_slot[NVL], order[NVL];
  uint32_t sig_key[NVL];
  uint8_t px0[NVL], py0[NVL];          // bbox origin of the previous level's components, by rank
  uint32_t sortbuf[C2G_VIEW_CAP];      // (cell_cnt << 16 | presort index), all levels back to back
  int n_views[C2G_NLEV], view_off[C2G_NLEV], layer_cnt[C2G_NLEV];
  TopView top[C2G_NLEV][C2G_MAX_DIST_FIRSTS];
  float divs[N_ANCH][N_DIVS];
  int cnt_point[N_ANCH];
  int ncomp, nsig, status, n_occ;
  double red[K2_WARPS];
  uint32_t t_off[C2G_VIEW_CAP];   // offset of every component's member-cell list in the CTA's global scratch
  uint16_t t_poi[C2G_VIEW_CAP], t_cnt[C2G_VIEW_CAP], torder[C2G_VIEW_CAP];
  int bucket_cnt[16], wq;
};

__device__ __forceinline__ uint32_t uf_find(volatile uint32_t *L, uint32_t c) {
  uint32_t w = L[c];
  uint32_t p = w & 0xFFFFu;
  while (p != c) {
    const uint32_t gp = L[p] & 0xFFFFu;
    if (gp != p) L[c] = (w & 0xFFFF0000u) | gp;  // path halving: any ancestor is a valid parent (links only go down)
    c = p;
    w = L[c];
    p = w & 0xFFFFu;
  }
  return c;
}
// read-only variant for the flatten phase: there every thread stores the final root into its OWN cells, and a path-halving
// store from another thread could overwrite that root with a stale ancestor
__device__ __forceinline__ uint32_t uf_find_ro(const volatile uint32_t *L, uint32_t c) {
  uint32_t p = L[c] & 0xFFFFu;
  while (p != c) {
    c = p;
    p = L[c] & 0xFFFFu;
  }
  return c;
}
__device__ __forceinline__ void uf_union(uint32_t *L, uint32_t a, uint32_t b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) {
      uint32_t t = a;
      a = b;
      b = t;
    }
    const uint32_t hi = ((volatile uint32_t *) L)[a] & 0xFFFF0000u;
    const uint32_t old = atomicMin(&L[a], hi | b);
    if ((old & 0xFFFFu) == a) return;
    a = old & 0xFFFFu;
  }
}
__device__ __forceinline__ int slot_of(const uint32_t *L, int c) {
  uint32_t low = L[c] & 0xFFFFu;
  if (!(low & 0x8000u)) low = L[low] & 0xFFFFu;
  return (int) (low & 0x7FFFu);
}

// ---- Eigen::SelfAdjointEigenSolver<Matrix2f> (Eigen 3.3.7 iterative path), device restatement ----------------------
__device__ __forceinline__ void givens(float p, float q, float &c, float &s) {
  if (q == 0.0f) {
    c = p < 0.0f ? -1.0f : 1.0f;
    s = 0.0f;
  } else if (p == 0.0f) {
    c = 0.0f;
    s = q < 0.0f ? 1.0f : -1.0f;
  } else if (fabsf(p) > fabsf(q)) {
    const float t = q / p;
    float u = sqrtf(1.0f + t * t);
    if (p < 0.0f) u = -u;
    c = 1.0f / u;
    s = -t * c;
  } else {
    const float t = p / q;
    float u = sqrtf(1.0f + t * t);
    if (q < 0.0f) u = -u;
    s = -1.0f / u;
    c = -t * s;
  }
}
__device__ void eig_sym2(float a, float b, float c, float ev[2], float vec[4]) {
  float scale = fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(c));
  if (scale == 0.0f) scale = 1.0f;
  float d0 = a / scale, d1 = c / scale, e = b / scale;
  d1 = d1 + (-1.0f) * (d1 * 0.0f + d1 * 0.0f);  // degenerate Householder rank update of the 2x2 tridiagonalisation
  float q00 = 1.0f, q10 = 0.0f, q01 = 0.0f, q11 = 1.0f;
  bool converged = true;
  for (int iter = 0;; ) {
    if (fabsf(e) <= (fabsf(d0) + fabsf(d1)) * (2.0f * 1.1920928955078125e-07f) || fabsf(e) <= 1.17549435082228750797e-38f) e = 0.0f;
    if (e == 0.0f) break;
    if (++iter > 60) {
      converged = false;
      break;
    }
    const float td = (d0 - d1) * 0.5f;
    float mu = d1;
    if (td == 0.0f) {
      mu -= fabsf(e);
    } else {
      const float e2 = e * e;
      const float at = fabsf(td), ae = fabsf(e);
      float hp, hq;
      if (at > ae) {
        hp = at;
        hq = ae / hp;
      } else {
        hp = ae;
        hq = at / hp;
      }
      const float h = (hp == 0.0f) ? 0.0f : hp * sqrtf(1.0f + hq * hq);
      if (e2 == 0.0f)
        mu -= (e / (td + (td > 0.0f ? 1.0f : -1.0f))) * (e / h);
      else
        mu -= e2 / (td + (td > 0.0f ? h : -h));
    }
    float gc, gs;
    givens(d0 - mu, e, gc, gs);
    const float sdk = gs * d0 + gc * e;
    const float dkp1 = gs * e + gc * d1;
    const float nd0 = gc * (gc * d0 - gs * e) - gs * (gc * e - gs * d1);
    d1 = gs * sdk + gc * dkp1;
    e = gc * sdk - gs * dkp1;
    d0 = nd0;
    const float x0 = q00, y0 = q01, x1 = q10, y1 = q11;
    q00 = gc * x0 + (-gs) * y0;
    q01 = gs * x0 + gc * y0;
    q10 = gc * x1 + (-gs) * y1;
    q11 = gs * x1 + gc * y1;
  }
  if (converged && d1 < d0) {
    float t = d0;
    d0 = d1;
    d1 = t;
    t = q00;
    q00 = q01;
    q01 = t;
    t = q10;
    q10 = q11;
    q11 = t;
  }
  ev[0] = d0 * scale;
  ev[1] = d1 * scale;
  vec[0] = q00;
  vec[1] = q10;
  vec[2] = q01;
  vec[3] = q11;
}

struct Moments {
  int cnt;
  double s0, s1, t00, t01, t11, q0, q1;
  float vol3;
};

__device__ void calc_stat_vals(const Moments &m, const c2g_cm_config &cfg, int level, int poi_r, int poi_c, c2g_view &v) {
  v.level = (int16_t) level;
  v.poi_r = (int16_t) poi_r;
  v.poi_c = (int16_t) poi_c;
  v.cell_cnt = (int16_t) m.cnt;
  const float cntf = (float) m.cnt;
  v.pos_mean[0] = (float) m.s0 / cntf;
  v.pos_mean[1] = (float) m.s1 / cntf;
  v.vol3_mean = m.vol3 / cntf;
  v.com[0] = (float) m.q0 / m.vol3;
  v.com[1] = (float) m.q1 / m.vol3;
  v.eccen = 0.0f;
  for (int i = 0; i < 6; ++i) v.pad_[i] = 0;
  if (m.cnt < cfg.min_cell_cov) {
    const float s2 = 1.0f * cfg.point_sigma * cfg.point_sigma, z2 = 0.0f * cfg.point_sigma * cfg.point_sigma;
    v.pos_cov[0] = s2;
    v.pos_cov[1] = z2;
    v.pos_cov[2] = z2;
    v.pos_cov[3] = s2;
    v.eig_vals[0] = v.eig_vals[1] = cfg.point_sigma;
    v.eig_vecs[0] = 1.0f;
    v.eig_vecs[1] = 0.0f;
    v.eig_vecs[2] = 0.0f;
    v.eig_vecs[3] = 1.0f;
    v.ecc_feat = 0;
    v.com_feat = 0;
  } else {
    const float cm1 = (float) (m.cnt - 1);
    const float m0 = v.pos_mean[0], m1 = v.pos_mean[1];
    const float c00 = ((float) m.t00 - (m0 * m0) * cntf) / cm1;
    const float c01 = ((float) m.t01 - (m0 * m1) * cntf) / cm1;
    const float c11 = ((float) m.t11 - (m1 * m1) * cntf) / cm1;
    v.pos_cov[0] = c00;
    v.pos_cov[1] = c01;
    v.pos_cov[2] = c01;
    v.pos_cov[3] = c11;
    float ev[2], vec[4];
    eig_sym2(c00, c01, c11, ev, vec);
    if (ev[0] < cfg.point_sigma) ev[0] = cfg.point_sigma;
    if (ev[1] < cfg.point_sigma) ev[1] = cfg.point_sigma;
    v.eig_vals[0] = ev[0];
    v.eig_vals[1] = ev[1];
    for (int i = 0; i < 4; ++i) v.eig_vecs[i] = vec[i];
    v.eccen = sqrtf(ev[1] * ev[1] - ev[0] * ev[0]) / ev[1];
    const bool dp = fabsf((ev[0] - ev[1]) / fmaxf(ev[0], ev[1])) > 0.2f;
    v.ecc_feat = (m.cnt > 5 && dp && ev[1] > 2.5f) ? 1 : 0;
    const float dx = v.com[0] - m0, dy = v.com[1] - m1;
    v.com_feat = (sqrtf(dx * dx + dy * dy) > cfg.com_bias_thres) ? 1 : 0;
  }
}

// V * diag(lambda) * V^T in float (ContourView::getManualCov, include/cont2/contour.h:376-378), column-major out
__device__ __forceinline__ void manual_cov(const float ev[2], const float vec[4], float out[4]) {
  const float vd00 = vec[0] * ev[0], vd10 = vec[1] * ev[0], vd01 = vec[2] * ev[1], vd11 = vec[3] * ev[1];
  out[0] = vd00 * vec[0] + vd01 * vec[2];
  out[1] = vd10 * vec[0] + vd11 * vec[2];
  out[2] = vd00 * vec[1] + vd01 * vec[3];
  out[3] = vd10 * vec[1] + vd11 * vec[3];
}

__global__ void __launch_bounds__(K2_THREADS, 1)
contour_kernel(const c2g_cellkey *__restrict__ tiles, const float4 *__restrict__ pts, const long long *__restrict__ offsets,
               int B, C2gIngestParams P, const int *__restrict__ int_ids, int first_slot, float *__restrict__ bev_h,
               float *__restrict__ bev_rf, float *__restrict__ bev_cf, c2g_view *__restrict__ presort_scratch,
               c2g_scan_head *__restrict__ heads, c2g_view *__restrict__ views, c2g_ell *__restrict__ ells,
               uint16_t *__restrict__ cell_lists, int *__restrict__ work_counter, long long *__restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = P.n_cells, ncol = P.cfg.n_col, nrow = P.cfg.n_row;
  const c2g_cm_config &cfg = P.cfg;
  c2g_view *const presort = presort_scratch + (size_t) blockIdx.x * C2G_VIEW_CAP;

  // scans are handed out dynamically (their cost varies 2x with the scene, and a CTA that starts late - e.g. behind a
  // co-running collective - must not leave a static share of the batch unprocessed until the end)
  __shared__ int next_scan;
  while (true) {
    if (tid == 0) next_scan = atomicAdd(work_counter, 1);
    __syncthreads();
    const int b = next_scan;
    if (b >= B) break;
    const size_t cbase = (size_t) b * ncell;
    const float *hg = bev_h + cbase, *rfg = bev_rf + cbase, *cfp = bev_cf + cbase;
    c2g_scan_head *head = heads + (first_slot + b);
    c2g_view *vout = views + (size_t) (first_slot + b) * C2G_VIEW_CAP;
    c2g_ell *eout = ells + (size_t) (first_slot + b) * C2G_VIEW_CAP;

#define C2G_DBG(i) do { if (dbg && blockIdx.x == 0 && tid == 0) dbg[i] = clock64(); } while (0)
    C2G_DBG(0);
    // ---------------- phase A: decode the tile, gather the winner's continuous coordinates -------------------------
    if (tid == 0) {
      S.status = 0;
      S.n_occ = 0;
    }
    __syncthreads();
    {
      const float4 *p = pts + offsets[b];
      int occ = 0;
      constexpr int PA = 6;  // cells per thread per batch: all tile loads, then all gathers, then the stores
      for (int c0 = tid; c0 < ncell; c0 += K2_THREADS * PA) {
        c2g_cellkey k[PA];
        float2 xy[PA];
#pragma unroll
        for (int u = 0; u < PA; ++u) {
          const int c = c0 + u * K2_THREADS;
          k[u] = c < ncell ? tiles[cbase + c] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < PA; ++u) {
          xy[u] = make_float2(0.f, 0.f);
          if (k[u] != 0ull) xy[u] = *reinterpret_cast<const float2 *>(p + (0xFFFFFFFFu - (uint32_t) k[u]));
        }
#pragma unroll
        for (int u = 0; u < PA; ++u) {
          const int c = c0 + u * K2_THREADS;
          if (c >= ncell) continue;
          float h = -1000.0f, rf = -1.0f, cf = -1.0f;
          uint8_t m = 0;
          if (k[u] != 0ull) {
            h = c2g_from_orderable((uint32_t) (k[u] >> 32));
            // pointToContRowCol (contour_mng.h:468-472): x / reso + n_row / 2 - 0.5f, left to right in float
            rf = (xy[u].x / cfg.reso_row + P.half_row_f) - 0.5f;
            cf = (xy[u].y / cfg.reso_col + P.half_col_f) - 0.5f;
#pragma unroll
            for (int e = 0; e < C2G_NLEV; ++e) m |= (h > cfg.lv_grads[e]) ? (1u << e) : 0u;
            occ++;
          }
          bev_h[cbase + c] = h;
          bev_rf[cbase + c] = rf;
          bev_cf[cbase + c] = cf;
          S.msk[c] = m;
          S.L[c] = 0u;
        }
      }
      for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(0xFFFFFFFFu, occ, o);
      if (lane == 0 && occ) atomicAdd(&S.n_occ, occ);  // one same-address shared atomic per warp, not per thread
    }
    __syncthreads();

    C2G_DBG(1);
    // ---------------- phase B: levels ------------------------------------------------------------------------------
    // Every thread owns a contiguous chunk of <= 32 cells and keeps, per level, the bitmask of its foreground cells in a
    // register: the per-level passes below visit set bits only (a few percent of the BEV is above any threshold).
    const int chunk = (ncell + K2_THREADS - 1) / K2_THREADS;
    const int cb0 = tid * chunk, cb1 = min(ncell, cb0 + chunk);
    const int r_first = cb0 / ncol, c_first = cb0 - r_first * ncol;
    uint32_t mk[C2G_NLEV];
    uint32_t rowst = 0;  // cells of the chunk that sit in column 0 (a horizontal run cannot continue across them)
    {
#pragma unroll
      for (int l = 0; l < C2G_NLEV; ++l) mk[l] = 0;
      int col = c_first;
      for (int k = 0; cb0 + k < cb1; ++k) {
        const uint32_t m = S.msk[cb0 + k];
#pragma unroll
        for (int l = 0; l < C2G_NLEV; ++l) mk[l] |= ((m >> l) & 1u) << k;
        if (col == 0) rowst |= 1u << k;
        if (++col == ncol) col = 0;
      }
    }
    uint16_t *const lists = cell_lists + (size_t) blockIdx.x * C2G_NLEV * ncell;  // per-CTA scratch: member cells per component
    int total_views = 0;
    for (int lev = 0; lev < C2G_NLEV; ++lev) {
      const uint8_t bit = (uint8_t) (1u << lev);
      const uint32_t bits = mk[lev];
      const uint32_t starts = bits & (~(bits << 1) | rowst);  // first cell of every horizontal run inside the chunk
      if (tid == 0) {
        S.ncomp = 0;
        S.nsig = 0;
      }
      // B1 label words: upper half keeps the rank of the enclosing component of the previous level, lower half links every
      // cell straight to the first cell of its run
      for (uint32_t bb = bits; bb; bb &= bb - 1) {
        const int k = __ffs(bb) - 1;
        const int rs = 31 - __clz(starts & ((2u << k) - 1u));
        S.L[cb0 + k] = (S.L[cb0 + k] & 0xFFFF0000u) | (uint32_t) (cb0 + rs);
      }
      __syncthreads();
      C2G_DBG(10 + lev * 8 + 0);
      // B2 unions. W: only where a run was cut by the chunk boundary. Row above: N if set (NW/NE then belong to N's run);
      // otherwise NW and NE. A cell whose W neighbour is set skips what W already did (its N/NE are this cell's NW/N).
      for (uint32_t bb = bits; bb; bb &= bb - 1) {
        const int k = __ffs(bb) - 1;
        const int c = cb0 + k;
        int r = r_first, cc = c_first + k;
        while (cc >= ncol) {
          cc -= ncol;
          ++r;
        }
        const bool w_set = cc > 0 && (k > 0 ? ((bits >> (k - 1)) & 1u) : (S.msk[c - 1] & bit));
        if (w_set && k == 0) uf_union(S.L, c, c - 1);
        if (r > 0) {
          const int up = c - ncol;
          const bool n_set = (S.msk[up] & bit) != 0;
          const bool ne_set = cc + 1 < ncol && (S.msk[up + 1] & bit);
          if (!w_set) {
            if (n_set)
              uf_union(S.L, c, up);
            else {
              if (cc > 0 && (S.msk[up - 1] & bit)) uf_union(S.L, c, up - 1);
              if (ne_set) uf_union(S.L, c, up + 1);
            }
          } else if (!n_set && ne_set)
            uf_union(S.L, c, up + 1);
        }
      }
      __syncthreads();
      C2G_DBG(10 + lev * 8 + 1);
      // B3 flatten, phase 1: compressing finds (path halving) on the run starts shorten every chain; no cell is finalised yet,
      // so a halving store can never clobber a final root
      for (uint32_t sb = starts; sb; sb &= sb - 1) (void) uf_find(S.L, cb0 + __ffs(sb) - 1);
      __syncthreads();
      // phase 2: one read-only find per run (now a hop or two), the whole run takes that root
      for (uint32_t sb = starts; sb; sb &= sb - 1) {
        const int k0 = __ffs(sb) - 1;
        const uint32_t root = uf_find_ro(S.L, cb0 + k0);
        const uint32_t run = ((bits >> k0) + 1u == 0u) ? 0xFFFFFFFFu : (((bits >> k0) ^ ((bits >> k0) + 1u)) >> 1);  // low ones of bits>>k0
        uint32_t rb = run;
        const uint32_t nxt = (starts >> k0) & ~1u;  // a row start inside the run splits it
        if (nxt) rb &= (nxt & (0u - nxt)) - 1u;
        for (; rb; rb &= rb - 1) {
          const int c = cb0 + k0 + __ffs(rb) - 1;
          S.L[c] = (S.L[c] & 0xFFFF0000u) | root;
        }
      }
      __syncthreads();
      C2G_DBG(10 + lev * 8 + 2);
      // B4 roots -> table slots. A root is the smallest cell index of its component, hence a run start. One shared-memory
      // atomic per warp (same-address atomics serialise).
      {
        int nroot = 0;
        for (uint32_t sb = starts; sb; sb &= sb - 1) {
          const int c = cb0 + __ffs(sb) - 1;
          if ((S.L[c] & 0xFFFFu) == (uint32_t) c) ++nroot;
        }
        int incl = nroot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= o) incl += t;
        }
        int base = 0;
        if (lane == 31 && incl > 0) base = atomicAdd(&S.ncomp, incl);
        base = __shfl_sync(0xFFFFFFFFu, base, 31);
        int slot_next = base + incl - nroot;
        if (nroot > 0)
          for (uint32_t sb = starts; sb; sb &= sb - 1) {
            const int c = cb0 + __ffs(sb) - 1;
            if ((S.L[c] & 0xFFFFu) != (uint32_t) c) continue;
            int slot = slot_next++;
            if (slot < NC) {
              S.c_area[slot] = 0;
              S.c_minr[slot] = 1 << 20;
              S.c_minc[slot] = 1 << 20;
              S.c_maxr[slot] = -1;
              S.c_maxc[slot] = -1;
              S.c_key[slot] = 1 << 30;
              S.c_poi[slot] = -1;
              S.c_pcid[slot] = (uint16_t) (S.L[c] >> 16);
              S.c_rank[slot] = 0xFFFFu;
            } else {
              slot = 0x7FFF;
              atomicOr(&S.status, 2);
            }
            S.L[c] = (S.L[c] & 0xFFFF0000u) | 0x8000u | (uint32_t) slot;
          }
      }
      __syncthreads();
      C2G_DBG(10 + lev * 8 + 3);
      // B5 per-component area / bbox / first-2x2-block key / last pixel: one flush per horizontal run of the chunk
      for (uint32_t sb = starts; sb; sb &= sb - 1) {
        const int k0 = __ffs(sb) - 1;
        uint32_t run = ((bits >> k0) + 1u == 0u) ? 0xFFFFFFFFu : (((bits >> k0) ^ ((bits >> k0) + 1u)) >> 1);
        const uint32_t nxt = (starts >> k0) & ~1u;
        if (nxt) run &= (nxt & (0u - nxt)) - 1u;
        const int len = __popc(run);
        const int c = cb0 + k0;
        int r = r_first, cc = c_first + k0;
        while (cc >= ncol) {
          cc -= ncol;
          ++r;
        }
        const int cur = slot_of(S.L, c);
        if (cur == 0x7FFF) continue;
        atomicAdd(&S.c_area[cur], len);
        atomicMin(&S.c_minr[cur], r);
        atomicMax(&S.c_maxr[cur], r);
        atomicMin(&S.c_minc[cur], cc);
        atomicMax(&S.c_maxc[cur], cc + len - 1);
        const int pc = S.c_pcid[cur];
        const int x0 = lev ? (int) S.px0[pc & (NVL - 1)] : 0, y0 = lev ? (int) S.py0[pc & (NVL - 1)] : 0;
        atomicMin(&S.c_key[cur], ((r - y0) >> 1) * 128 + ((cc - x0) >> 1));
        atomicMax(&S.c_poi[cur], c + len - 1);
      }
      __syncthreads();
      C2G_DBG(10 + lev * 8 + 4);
      // B6 significant components -> DFS order rank
      const int ncomp = min(S.ncomp, NC);
      for (int s = tid; s < ncomp; s += K2_THREADS)
        if (S.c_area[s] >= cfg.min_cont_cell_cnt) {
          const int i = atomicAdd(&S.nsig, 1);
          if (i < NVL) {
            S.sig_slot[i] = (uint16_t) s;
            S.sig_key[i] = ((uint32_t) S.c_pcid[s] << 16) | (uint32_t) S.c_key[s];
          } else
            atomicOr(&S.status, 2);
        }
      __syncthreads();
      int nsig = min(S.nsig, NVL);
      if (total_views + nsig > C2G_VIEW_CAP) {
        nsig = C2G_VIEW_CAP - total_views;
        if (tid == 0) atomicOr(&S.status, 1);
      }
      for (int i = tid; i < min(S.nsig, NVL); i += K2_THREADS) {
        const uint32_t ki = S.sig_key[i];
        int rank = 0;
        for (int j = 0; j < min(S.nsig, NVL); ++j) rank += (S.sig_key[j] < ki) ? 1 : 0;
        if (rank < nsig) {
          const int sl = S.sig_slot[i];
          S.order[rank] = (uint16_t) sl;
          S.c_rank[sl] = (uint16_t) rank;
          S.sortbuf[total_views + rank] = ((uint32_t) S.c_area[sl] << 16) | (uint32_t) rank;
          S.t_poi[total_views + rank] = (uint16_t) S.c_poi[sl];
        }
      }
      __syncthreads();
      if (tid == 0) {
        S.n_views[lev] = nsig;
        S.view_off[lev] = total_views;
      }
      // list offsets: exclusive prefix of the areas in rank order (one warp; nsig is a few dozen)
      if (warp == 0) {
        int run_off = lev * ncell;
        for (int base = 0; base < nsig; base += 32) {
          const int i = base + lane;
          const int a = i < nsig ? (int) (S.sortbuf[total_views + i] >> 16) : 0;
          int incl = a;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
          }
          if (i < nsig) S.t_off[total_views + i] = (uint32_t) (run_off + incl - a);
          run_off += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
      }
      __syncthreads();
      C2G_DBG(10 + lev * 8 + 5);
      // B7 member cells of every significant component in bbox-raster order -> global scratch list (shared memory only on
      // the read side; the moment accumulation itself is deferred to the balanced task pool after the level loop)
      for (int rk = warp; rk < nsig; rk += K2_WARPS) {
        const int s = S.order[rk];
        const int r0 = S.c_minr[s], c0 = S.c_minc[s], w = S.c_maxc[s] - c0 + 1, hgt = S.c_maxr[s] - r0 + 1;
        uint16_t *wl = lists + S.t_off[total_views + rk];
        int nlist = 0;
        int rr = r0, cc = c0 + lane;  // lane's cell inside the bbox, advanced by 32 cells per step without divisions
        while (cc >= c0 + w) {
          cc -= w;
          ++rr;
        }
        for (int base = 0; base < w * hgt; base += 32) {
          bool member = false;
          int c = 0;
          if (rr < r0 + hgt) {
            c = rr * ncol + cc;
            member = (S.msk[c] & bit) && slot_of(S.L, c) == s;
          }
          const unsigned bal = __ballot_sync(0xFFFFFFFFu, member);
          if (member) wl[nlist + __popc(bal & ((1u << lane) - 1u))] = (uint16_t) c;
          nlist += __popc(bal);
          cc += 32;
          while (cc >= c0 + w) {
            cc -= w;
            ++rr;
          }
        }
      }
      C2G_DBG(10 + lev * 8 + 6);
      // B8 hand the ranks down: upper half of every foreground word = rank of its component (0xFFFF if insignificant)
      for (int i = tid; i < nsig; i += K2_THREADS) {
        const int s = S.order[i];
        S.px0[i] = (uint8_t) S.c_minc[s];
        S.py0[i] = (uint8_t) S.c_minr[s];
      }
      for (uint32_t bb = bits; bb; bb &= bb - 1) {
        const int c = cb0 + __ffs(bb) - 1;
        const int s = slot_of(S.L, c);
        const uint32_t rk = (s == 0x7FFF) ? 0xFFFFu : (uint32_t) S.c_rank[s];
        S.L[c] = (rk << 16) | (S.L[c] & 0xFFFFu);  // the low half (slot / root link) is what concurrent readers use
      }
      total_views += nsig;
      __syncthreads();
    }

    C2G_DBG(2);
    // ---------------- phase C: balanced task pool: per-level std::sort replays + moments / calcStatVals per component -----
    // Tasks are ordered by decreasing size class (floor(log2(area))) so the longest sequential accumulations start first.
    // Big components (> 32 cells) go to the front of the task order, small ones fill it from the back; positions are
    // claimed with one shared atomic per warp and class (same-address shared atomics serialise).
    if (tid == 0) {
      S.wq = 0;
      S.bucket_cnt[0] = 0;             // next free slot at the front
      S.bucket_cnt[1] = total_views;   // one past the last free slot at the back
    }
    __syncthreads();
    for (int v0 = warp * 32; v0 < total_views; v0 += K2_THREADS) {
      const int v = v0 + lane;
      const bool valid = v < total_views;
      const bool big = valid && (S.sortbuf[v] >> 16) > 32u;
      const unsigned mb = __ballot_sync(0xFFFFFFFFu, big), ms = __ballot_sync(0xFFFFFFFFu, valid && !big);
      int fb = 0, bb2 = 0;
      if (lane == 0) {
        if (mb) fb = atomicAdd(&S.bucket_cnt[0], __popc(mb));
        if (ms) bb2 = atomicSub(&S.bucket_cnt[1], __popc(ms));
      }
      fb = __shfl_sync(0xFFFFFFFFu, fb, 0);
      bb2 = __shfl_sync(0xFFFFFFFFu, bb2, 0);
      if (big) S.torder[fb + __popc(mb & ((1u << lane) - 1u))] = (uint16_t) v;
      if (valid && !big) S.torder[bb2 - 1 - __popc(ms & ((1u << lane) - 1u))] = (uint16_t) v;
    }
    __syncthreads();
    // sortbuf words of one level are (area << 16 | rank): sorting them in place is the std::sort of cont_views_[level]
    // (contour_mng.h:596-599); the walk tasks below read the areas from t_cnt copies taken before the sort starts
    for (int v = tid; v < total_views; v += K2_THREADS) S.t_cnt[v] = (uint16_t) (S.sortbuf[v] >> 16);
    __syncthreads();
    if (lane == 0 && warp < C2G_NLEV) {  // one warp per level: the six serial replays run on different schedulers
      uint32_t *first = S.sortbuf + S.view_off[warp];
      c2g_sort::std_sort(first, (long) S.n_views[warp], [](uint32_t a, uint32_t bb) { return (a >> 16) > (bb >> 16); });
      int sum = 0;
      for (int i = 0; i < S.n_views[warp]; ++i) sum += (int) (first[i] >> 16);
      S.layer_cnt[warp] = sum;
    }
    C2G_DBG(58);
    const int n_big = S.bucket_cnt[0];  // torder[0, n_big) = components of more than 32 cells, the rest follow
    // Small components: one THREAD per component walks its own member list (<= 32 cells) with the seven double
    // accumulators of RunningStatRecorder in registers, 32 components per warp instruction. Warps 0..5 are busy with the
    // per-level sorts above, warps 6..31 take the small components.
    if (warp >= C2G_NLEV) {
      for (int i = n_big + (tid - C2G_NLEV * 32); i < total_views; i += K2_THREADS - C2G_NLEV * 32) {
        const int v = S.torder[i];
        const int n = S.t_cnt[v];
        const uint16_t *wl = lists + S.t_off[v];
        double s0 = 0, s1 = 0, t00 = 0, t01 = 0, t11 = 0, q0 = 0, q1 = 0;
        float vol3 = 0.0f;
        for (int j = 0; j < n; ++j) {
          const int cc = wl[j];
          const float hh = hg[cc];
          const double v0 = (double) rfg[cc], v1 = (double) cfp[cc], hd = (double) hh;
          s0 += v0;
          s1 += v1;
          t00 += v0 * v0;
          t01 += v0 * v1;
          t11 += v1 * v1;
          vol3 += hh;
          q0 += hd * v0;
          q1 += hd * v1;
        }
        double *raw = reinterpret_cast<double *>(presort + v);
        raw[0] = s0;
        raw[1] = s1;
        raw[2] = t00;
        raw[3] = t01;
        raw[4] = t11;
        raw[5] = q0;
        raw[6] = q1;
        reinterpret_cast<float *>(raw + 7)[0] = vol3;
      }
    }
    {
      // Big components: one WARP per component from a shared queue (largest size class first). The seven double
      // accumulators live in lanes 0..6: lane k adds a_k * b_k per member cell (s0: v0*1, s1: v1*1, t00: v0*v0, t01: v0*v1,
      // t11: v1*v1, q0: h*v0, q1: h*v1; x*1.0 is exact), so a cell costs the FP64 pipe one DMUL + one DADD per warp.
      // Accumulation order = list order = bbox-raster order.
      const int selA = (lane == 0 || lane == 2 || lane == 3) ? 0 : (lane == 1 || lane == 4) ? 1 : (lane == 5 || lane == 6) ? 2 : 3;
      const int selB = (lane == 2 || lane == 5) ? 0 : (lane == 3 || lane == 4 || lane == 6) ? 1 : 3;
      while (true) {
        int ti = 0;
        if (lane == 0) ti = atomicAdd(&S.wq, 1);
        ti = __shfl_sync(0xFFFFFFFFu, ti, 0);
        if (ti >= n_big) break;
        const int v = S.torder[ti];
        const int n = S.t_cnt[v];
        const uint16_t *wl = lists + S.t_off[v];
        double acc = 0.0;
        float vol3 = 0.0f;
        for (int g = 0; g < n; g += 128) {
          float hv[4], rv[4], cv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = g + u * 32 + lane;
            hv[u] = rv[u] = cv[u] = 0.f;
            if (j < n) {
              const int cc = wl[j];
              hv[u] = hg[cc];
              rv[u] = rfg[cc];
              cv[u] = cfp[cc];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cntu = min(32, n - (g + u * 32));
            for (int src = 0; src < cntu; ++src) {
              // three 32-bit shuffles per cell; every lane converts only the two operands its accumulator needs
              const float hh = __shfl_sync(0xFFFFFFFFu, hv[u], src);
              const float f0 = __shfl_sync(0xFFFFFFFFu, rv[u], src);
              const float f1 = __shfl_sync(0xFFFFFFFFu, cv[u], src);
              const float fa = selA == 0 ? f0 : selA == 1 ? f1 : selA == 2 ? hh : 1.0f;
              const float fb = selB == 0 ? f0 : selB == 1 ? f1 : 1.0f;
              acc += (double) fa * (double) fb;
              vol3 += hh;
            }
          }
        }
        // raw moments go to the component's (still unused) 80-byte presort record: doubles 0..6 from lanes 0..6, the
        // float height sum in word 14; calcStatVals for all components runs afterwards with one thread per component
        double *raw = reinterpret_cast<double *>(presort + v);
        if (lane < 7) raw[lane] = acc;
        if (lane == 7) reinterpret_cast<float *>(raw + 7)[0] = vol3;
      }
    }
    C2G_DBG(59);
    __syncthreads();
    C2G_DBG(60);
    for (int v = tid; v < total_views; v += K2_THREADS) {
      const double *raw = reinterpret_cast<const double *>(presort + v);
      Moments m;
      m.cnt = S.t_cnt[v];
      m.s0 = raw[0];
      m.s1 = raw[1];
      m.t00 = raw[2];
      m.t01 = raw[3];
      m.t11 = raw[4];
      m.q0 = raw[5];
      m.q1 = raw[6];
      m.vol3 = reinterpret_cast<const float *>(raw + 7)[0];
      int lev = 0;
      for (int l = 0; l < C2G_NLEV; ++l)
        if (v >= S.view_off[l] && v < S.view_off[l] + S.n_views[l]) lev = l;
      c2g_view vw;
      const int poi = S.t_poi[v];
      calc_stat_vals(m, cfg, lev, poi / ncol, poi % ncol, vw);
      presort[v] = vw;  // in place: this thread is the only reader and writer of the record
    }
    __syncthreads();
    C2G_DBG(3);
    {
      // copy 80-byte records as 20 x 4-byte words: sorted position j of level l <- presort index (sortbuf & 0xFFFF)
      const uint32_t *src = reinterpret_cast<const uint32_t *>(presort);
      uint32_t *dst = reinterpret_cast<uint32_t *>(vout);
      constexpr int WPV = sizeof(c2g_view) / 4;
      for (int lev = 0; lev < C2G_NLEV; ++lev) {
        const int off = S.view_off[lev], n = S.n_views[lev];
        for (int i = tid; i < n * WPV; i += K2_THREADS) {
          const int j = i / WPV, wd = i - j * WPV;
          const int from = off + (int) (S.sortbuf[off + j] & 0xFFFFu);
          dst[(size_t) (off + j) * WPV + wd] = src[(size_t) from * WPV + wd];
        }
      }
      // compact GMM ellipse next to every view of the levels the GMM-L2 stages use (1..4)
      {
        const int e0 = S.view_off[1], e1 = S.view_off[C2G_NUM_BIN_LAYERS] + S.n_views[C2G_NUM_BIN_LAYERS];
        for (int v = e0 + tid; v < e1; v += K2_THREADS) {
          int lev = 1;
          for (int l = 2; l <= C2G_NUM_BIN_LAYERS; ++l)
            if (v >= S.view_off[l]) lev = l;
          const c2g_view &pv = presort[S.view_off[lev] + (int) (S.sortbuf[v] & 0xFFFFu)];
          float cv[4];
          manual_cov(pv.eig_vals, pv.eig_vecs, cv);
          c2g_ell e;
          e.mx = pv.pos_mean[0];
          e.my = pv.pos_mean[1];
          e.c00 = cv[0];
          e.c10 = cv[1];
          e.c01 = cv[2];
          e.c11 = cv[3];
          e.w = (float) pv.cell_cnt;
          e.maj = sqrtf(pv.eig_vals[1]);
          eout[v] = e;
        }
      }
      for (int i = tid; i < C2G_NLEV * C2G_MAX_DIST_FIRSTS; i += K2_THREADS) {
        const int lev = i / C2G_MAX_DIST_FIRSTS, j = i % C2G_MAX_DIST_FIRSTS;
        TopView t;
        t.cnt = 0;
        t.mean0 = t.mean1 = t.eig0 = t.eig1 = 0.f;
        if (j < S.n_views[lev]) {
          const c2g_view &v = presort[S.view_off[lev] + (S.sortbuf[S.view_off[lev] + j] & 0xFFFFu)];
          t.cnt = v.cell_cnt;
          t.mean0 = v.pos_mean[0];
          t.mean1 = v.pos_mean[1];
          t.eig0 = v.eig_vals[0];
          t.eig1 = v.eig_vals[1];
        }
        S.top[lev][j] = t;
      }
    }
    __syncthreads();

    C2G_DBG(4);
    // ---------------- phase D: retrieval keys (contour_mng.h:693-830) ------------------------------------------------
    // D1: per anchor, ordered list of the window cells that contribute: (dist, higher_cnt). One warp per anchor.
    float *klist_dist = reinterpret_cast<float *>(S.L);                               // [N_ANCH][KEY_LIST_CAP]
    uint8_t *klist_hc = reinterpret_cast<uint8_t *>(klist_dist + N_ANCH * KEY_LIST_CAP);  // [N_ANCH][KEY_LIST_CAP]
    static_assert(N_ANCH * KEY_LIST_CAP * 5 <= sizeof(uint32_t) * C2G_MAX_CELLS, "key lists must fit in the label array");
    const int piv = cfg.piv_firsts;
    const int roi_pad = (int) ceilf(cfg.roi_radius + 1.0f);
    for (int a = warp; a < N_ANCH; a += K2_WARPS) {
      const int ll = a / C2G_MAX_PIV, seq = a % C2G_MAX_PIV;
      int cnt = 0;
      const bool valid = seq < piv && seq < S.n_views[ll] && S.top[ll][seq].cnt >= cfg.min_cont_key_cnt;
      if (valid) {
        const float cx = S.top[ll][seq].mean0, cy = S.top[ll][seq].mean1;
        const int r_cen = (int) cx, c_cen = (int) cy;
        const int r_min = max(0, r_cen - roi_pad), r_max = min(nrow - 1, r_cen + roi_pad);
        const int c_min = max(0, c_cen - roi_pad), c_max = min(ncol - 1, c_cen + roi_pad);
        const int w = c_max - c_min + 1, total = w * (r_max - r_min + 1);
        const double rad = (double) cfg.roi_radius - 1e-2;
        for (int base = 0; base < total; base += 32) {
          const int i = base + lane;
          bool pass = false;
          float dist = 0.f;
          uint8_t hc = 0;
          if (i < total) {
            const int c = (r_min + i / w) * ncol + (c_min + i % w);
            const uint8_t m = S.msk[c];
            if (m & 2u) {  // bev > lv_grads[1]  (cells with bev == lv_grads[1] pass the first test but fail this one)
              const float dx = rfg[c] - cx, dy = cfp[c] - cy;
              dist = sqrtf(dx * dx + dy * dy);
              if ((double) dist < rad) {
                pass = true;
                hc = (uint8_t) __popc((unsigned) (m & 0x3Eu));
              }
            }
          }
          const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
          if (pass) {
            const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
            if (pos < KEY_LIST_CAP) {
              klist_dist[a * KEY_LIST_CAP + pos] = dist;
              klist_hc[a * KEY_LIST_CAP + pos] = hc;
            }
          }
          cnt += __popc(bal);
        }
        if (cnt > KEY_LIST_CAP && lane == 0) atomicOr(&S.status, 4);
      }
      if (lane == 0) S.cnt_point[a] = valid ? cnt : -1;
    }
    __syncthreads();
    C2G_DBG(5);
    // D2: one thread per (anchor, division): sequential float accumulation in raster order
    {
      const float div_len = cfg.roi_radius / (float) ((C2G_KEY_DIM - 3) * 5);
      const double inv_norm_den = sqrt(2 * 3.14159265358979323846 * 1.0f * 1.0f);
      const double rcp_norm_den = 1.0 / inv_norm_den;
      for (int w = tid; w < N_ANCH * N_DIVS; w += K2_THREADS) {
        const int a = w / N_DIVS, d = w - a * N_DIVS;
        const int n = min(S.cnt_point[a], KEY_LIST_CAP);
        float acc = 0.0f;
        if (n > 0) {
          const float x = (float) ((double) ((float) d * div_len) + 0.5 * (double) div_len);
          const float *dl = klist_dist + a * KEY_LIST_CAP;
          const uint8_t *hl = klist_hc + a * KEY_LIST_CAP;
          for (int k = 0; k < n; ++k) {
            const float t = (x - dl[k]) / 1.0f;
            const double q = (-0.5 * (double) t) * (double) t;
            const double e = c2g_exp(q, P.exp_mode, c2g_exp_tab_dev);
            // (float) (e / den): the product with the rounded reciprocal is within 2.5 ulp of the correctly rounded
            // quotient, so both round to the same float unless the product sits within a few ulp of a float rounding
            // midpoint (low 29 mantissa bits == 0x10000000); only then is the real division executed.
            double pq = e * rcp_norm_den;
            const unsigned long long low = (unsigned long long) __double_as_longlong(pq) & 0x1FFFFFFFull;
            if (low - 0x0FFFFFF8ull <= 0x10ull) pq = e / inv_norm_den;
            const float g = (float) pq;
            acc += (float) hl[k] * g;
          }
        }
        S.divs[a][d] = acc;
      }
    }
    __syncthreads();
    for (int w = tid; w < N_ANCH * C2G_KEY_DIM; w += K2_THREADS) {
      const int a = w / C2G_KEY_DIM, kd = w - a * C2G_KEY_DIM;
      const int ll = a / C2G_MAX_PIV, seq = a % C2G_MAX_PIV;
      float val = 0.0f;
      if (S.cnt_point[a] >= 0) {
        const TopView &t = S.top[ll][seq];
        if (kd == 0)
          val = sqrtf(t.eig1 * (float) t.cnt);
        else if (kd == 1)
          val = sqrtf(t.eig0 * (float) t.cnt);
        else if (kd == 2) {
          int accum = 0;
          for (int s2 = 0; s2 <= seq; ++s2) accum += S.top[ll][s2].cnt;
          val = (float) sqrt((double) accum);
        } else {
          const int bn = kd - 3;
          float ring = 0.0f;
          for (int d = 0; d < 5; ++d) ring += S.divs[a][bn * 5 + d];
          const float bin_len = cfg.roi_radius / (float) (C2G_KEY_DIM - 3);
          ring = (float) ((double) ring * ((double) bin_len / sqrt((double) S.cnt_point[a])));
          val = ring;
        }
      }
      head->keys[ll][seq][kd] = val;
    }

    C2G_DBG(6);
    // ---------------- phase E: BCIs (contour_mng.h:848-883), one warp per anchor ------------------------------------
    // candidate t = bl * 10 + j (reference loop order) is evaluated by lane t % 32; kept neighbours are compacted in
    // that order, lane 0 replays std::sort on bit_pos and builds the run boundaries, all lanes write the record.
    {
      c2g_relpt *nei_all = reinterpret_cast<c2g_relpt *>(S.L);                                         // [warps][40]
      uint32_t *ord_all = reinterpret_cast<uint32_t *>(nei_all + K2_WARPS * C2G_MAX_NEI);              // [warps][40]
      uint16_t *seg_all = reinterpret_cast<uint16_t *>(ord_all + K2_WARPS * C2G_MAX_NEI);              // [warps][42]
      c2g_relpt *nei = nei_all + warp * C2G_MAX_NEI;
      uint32_t *ord = ord_all + warp * C2G_MAX_NEI;
      uint16_t *segv = seg_all + warp * (C2G_MAX_NEI + 2);
      for (int a = warp; a < N_ANCH; a += K2_WARPS) {
        const int ll = a / C2G_MAX_PIV, seq = a % C2G_MAX_PIV;
        if (seq >= piv) continue;
        c2g_bci &bci = head->bcis[ll][seq];
        int n = 0;
        if (S.cnt_point[a] >= 0) {
          const TopView an = S.top[ll][seq];
          for (int t0 = 0; t0 < C2G_MAX_NEI; t0 += 32) {
            const int t = t0 + lane;
            bool keep = false;
            c2g_relpt rp;
            rp.level = 0;
            rp.seq = 0;
            rp.bit_pos = 0;
            rp.r = 0.f;
            rp.theta = 0.f;
            if (t < C2G_MAX_NEI) {
              const int bl = t / C2G_MAX_DIST_FIRSTS, j = t % C2G_MAX_DIST_FIRSTS, layer = bl + 1;
              if (j < min(cfg.dist_firsts, S.n_views[layer]) && !(ll == layer && j == seq)) {
                const float vx = S.top[layer][j].mean0 - an.mean0, vy = S.top[layer][j].mean1 - an.mean1;
                const float dist = sqrtf(vx * vx + vy * vy);
                if (!((double) dist > (C2G_BITS_PER_LAYER - 1) * 1.01 + 5.43 - 1e-3 || (double) dist <= 5.43)) {
                  keep = true;
                  const int idx = (int) (fmin(floor(((double) dist - 5.43) / 1.01), C2G_BITS_PER_LAYER - 1.0) + (double) (bl * C2G_BITS_PER_LAYER));
                  rp.level = (int8_t) layer;
                  rp.seq = (int8_t) j;
                  rp.bit_pos = (int16_t) idx;
                  rp.r = dist;
                  rp.theta = c2g_atan2f(vy, vx);
                }
              }
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (keep) {
              const int pos = n + __popc(bal & ((1u << lane) - 1u));
              nei[pos] = rp;
              ord[pos] = ((uint32_t) (uint16_t) rp.bit_pos << 16) | (uint32_t) pos;
            }
            n += __popc(bal);
          }
        }
        __syncwarp();
        int nseg = 0;
        uint64_t bins[4] = {0, 0, 0, 0};
        if (lane == 0) {
          c2g_sort::std_sort(ord, (long) n, [](uint32_t x, uint32_t y) { return (x >> 16) < (y >> 16); });
          if (n > 0) {
            segv[nseg++] = 0;
            for (int p1 = 0; p1 < n; ++p1) {
              const int bp = (int) (ord[p1] >> 16);
              bins[bp >> 6] |= 1ull << (bp & 63);
              if ((ord[segv[nseg - 1]] >> 16) != (ord[p1] >> 16)) segv[nseg++] = (uint16_t) p1;
            }
            segv[nseg++] = (uint16_t) n;
          }
          for (int i = 0; i < 4; ++i) bci.dist_bin[i] = bins[i];
          bci.n_nei = (int16_t) n;
          bci.n_seg = (int16_t) nseg;
          bci.piv_seq = (int8_t) seq;
          bci.level = (int8_t) ll;
          for (int i = 0; i < 6; ++i) bci.pad_[i] = 0;
        }
        nseg = __shfl_sync(0xFFFFFFFFu, nseg, 0);
        __syncwarp();
        for (int i = lane; i < C2G_MAX_NEI; i += 32) {
          c2g_relpt rp;
          rp.level = 0;
          rp.seq = 0;
          rp.bit_pos = 0;
          rp.r = 0.f;
          rp.theta = 0.f;
          if (i < n) rp = nei[ord[i] & 0xFFFFu];
          bci.nei[i] = rp;
        }
        for (int i = lane; i < C2G_MAX_NEI + 2; i += 32) bci.seg[i] = i < nseg ? segv[i] : (uint16_t) 0;
        __syncwarp();
      }
    }
    C2G_DBG(7);
    // ---------------- phase F: scan-only GMM terms (correlation.h:49-82,102-119) -----------------------------------
    __shared__ int n_ell_s[C2G_NUM_BIN_LAYERS];
    if (tid < C2G_NUM_BIN_LAYERS) {
      const int lev = tid + 1;
      const int full = S.layer_cnt[lev];
      int run = 0, k = 0;
      const uint32_t *sb = S.sortbuf + S.view_off[lev];
      for (; k < S.n_views[lev]; ++k) {
        if ((double) run * 1.0 / (double) full >= 0.95) break;
        run += (int) (sb[k] >> 16);
      }
      n_ell_s[tid] = k;
    }
    __syncthreads();
    double acc = 0.0;
    for (int li = 0; li < C2G_NUM_BIN_LAYERS; ++li) {
      const int n = n_ell_s[li];
      const c2g_ell *le = eout + S.view_off[li + 1];
      for (int w = tid; w < n * n; w += K2_THREADS) {
        const int i = w / n, j = w - i * n;
        const c2g_ell A = le[i], Bv = le[j];
        const double c00 = 2.0 * ((double) A.c00 + (double) Bv.c00), c10 = 2.0 * ((double) A.c10 + (double) Bv.c10);
        const double c01 = 2.0 * ((double) A.c01 + (double) Bv.c01), c11 = 2.0 * ((double) A.c11 + (double) Bv.c11);
        const double mx = (double) A.mx - (double) Bv.mx, my = (double) A.my - (double) Bv.my;
        const double det = c00 * c11 - c01 * c10;
        const double invdet = 1.0 / det;
        const double qf = mx * ((c11 * invdet) * mx + (-c01 * invdet) * my) + my * ((-c10 * invdet) * mx + (c00 * invdet) * my);
        acc += (double) A.w * (double) Bv.w / sqrt(det) * c2g_exp(-0.5 * qf, P.exp_mode, c2g_exp_tab_dev);
      }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) S.red[warp] = acc;
    __syncthreads();
    C2G_DBG(8);
    // ---------------- phase G: head ------------------------------------------------------------------------------------
    if (tid == 0) {
      double tot = 0.0;
      for (int i = 0; i < K2_WARPS; ++i) tot += S.red[i];
      head->int_id = int_ids ? int_ids[b] : (first_slot + b);
      head->status = S.status;
      for (int l = 0; l < C2G_NLEV; ++l) {
        head->n_views[l] = S.n_views[l];
        head->view_off[l] = S.view_off[l];
        head->layer_cell_cnt[l] = S.layer_cnt[l];
      }
      for (int l = 0; l < C2G_NUM_BIN_LAYERS; ++l) head->n_ell[l] = n_ell_s[l];
      head->n_occupied = S.n_occ;
      head->pad_ = 0;
      head->gmm_auto_corr = tot;
    }
    __syncthreads();
    C2G_DBG(9);
  }
}

}  // namespace

size_t c2g_contour_smem_bytes() { return sizeof(Smem); }

int c2g_launch_contours(const c2g_cellkey *tiles, const float *pts_dev, const long long *offsets_dev, int B,
                        const C2gIngestParams &P, const int *int_ids_dev, int first_slot, float *bev_h, float *bev_rf,
                        float *bev_cf, c2g_view *presort_scratch, c2g_scan_head *heads, c2g_view *views, c2g_ell *ells,
                        uint16_t *cell_lists, int *work_counter, int num_sms, cudaStream_t stream, long long *dbg) {
  static bool attr_set = false;
  if (!attr_set) {
    C2G_CUDA_TRY(cudaFuncSetAttribute(contour_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(Smem)));
    attr_set = true;
  }
  const int grid = B < num_sms ? B : num_sms;
  if (grid <= 0) return 0;
  C2G_CUDA_TRY(cudaMemsetAsync(work_counter, 0, sizeof(int), stream));
  contour_kernel<<<grid, K2_THREADS, sizeof(Smem), stream>>>(tiles, (const float4 *) pts_dev, offsets_dev, B, P, int_ids_dev,
                                                             first_slot, bev_h, bev_rf, bev_cf, presort_scratch, heads, views, ells, cell_lists,
                                                             work_counter, dbg);
  C2G_CUDA_TRY(cudaGetLastError());
  return 0;
}
