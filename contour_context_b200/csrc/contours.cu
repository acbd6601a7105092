// contours.cu — K2: BEV bit-planes + foreground cell list (bev_scatter.cu) -> multi-level contours -> ContourView statistics -> per-level order -> retrieval keys ->
// BCIs -> per-scan GMM terms.  Persistent CTAs of 512 threads, TWO resident per SM (<= 113 KB of shared memory each), one
// scan per CTA iteration; everything between the BEV tile and the finished descriptor stays in shared memory / L1 / L2.
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
// the reference's recursive ROI-by-ROI labelling, and the six labellings are independent of each other.  What the
// recursion adds is ORDER: cont_views_[L] is filled in DFS order, children of one parent in OpenCV label order, which is
// the block-raster order of each component's first 2x2 block with blocks aligned to the parent's bounding-box origin
// (SURVEY.md §7 hard part 2).  Hence
//   order(level L) = sort by (rank of parent in level L-1, min over pixels of ((r-y0)>>1, (c-x0)>>1)).
//
// Data structures: the thresholded image of every level is a bit-plane (one 32-bit word per 32 columns of a row, built by
// the scatter kernel from its shared-memory tile - round 1 sent the 180 KB tile through HBM and decoded it here); the
// heights and continuous coordinates of the ~6 % of cells above the lowest threshold arrive as a compact raster-ordered list
// (16 B per cell, index = number of plane-0 bits before the cell), so a run of cells is a contiguous slice of it; the union-find runs over horizontal RUNS of set bits, not
// over pixels (a 150 x 150 KITTI BEV has ~1 400 foreground cells but only ~600 runs and ~80 components per level).  One
// warp labels one level (six levels at once, no block barrier inside), runs are numbered in raster order so that
//   * the smallest run id of a component is its first pixel in raster order (the root of min-linking union-find),
//   * the largest run id ends in the component's last pixel (ContourView::poi),
//   * walking a component's runs in id order visits its cells in the reference's accumulation order (bbox-raster order
//     restricted to members == raster order of members).
// Two warps share a level (named barriers between them); per-component statistics are native 32-bit shared atomics on
// four packed words per component.
//
// Bit-exactness rules: float/double sums that feed views and keys are accumulated in the reference's raster order by a
// single logical accumulator, never by tree reductions; compiled with -fmad=false.
#include <math_constants.h>

#include <cstdlib>
#include <type_traits>

#include "c2g_common.cuh"
#include "stdsort.cuh"
#include "c2g_libm.cuh"

namespace {

// glibc's __exp_data.tab (2 KB, read through L1); see c2g_libm.cuh
__device__ const uint64_t c2g_exp_tab_dev[256] = C2G_EXP_TAB_INIT;

constexpr int K2_THREADS = 512;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int K2_CTAS_PER_SM = 2;
constexpr int PLANE_WORDS = 800;  // n_row * ceil(n_col / 32) (checked by make_params): 150 x 5 = 750 for both shipped configs
constexpr int R_POOL = 4096;      // runs of all six levels that fit in shared memory (else: global arena)
constexpr int C_POOL = 960;       // components of all six levels that fit in shared memory (else: global arena)
constexpr int NVL = 1024;         // significant components (area >= min_cont_cell_cnt) per level
constexpr int FG_CAP = 2048;       // foreground cells whose (height, row_f, col_f) are staged in shared memory (else: read from global)
constexpr int KEY_LIST_CAP = 400; // cells of one key window that can lie inside the 9.99-cell radius
constexpr int N_ANCH = C2G_NLEV * C2G_MAX_PIV;
constexpr int N_DIVS = 35;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct TopView {  // what keys / BCI / GMM need from a sorted view
  float mean0, mean1, eig0, eig1;
  int cnt;
};

// One connected component of one level = four 32-bit words, every field that several lanes update concurrently sits where
// one native atomic can maintain it (root / last: first / last run of the component, level-local run ids in raster order):
//   w0 = root << 16 | area                  area: atomicAdd (<= 22 500 cells, never carries into the upper half)
//   w1 = last << 16 | key                   last: atomicMax with the low half 0xFFFF; key (first 2x2 block, relative to the
//                                           parent's bounding-box origin): atomicMin once `last` is final
//   w2 = (255 - minc) << 24 | pcomp         minc: atomicMax with the low bits 0xFFFFFF; pcomp = enclosing component of the
//                                           previous level (0xFFFF = none), written afterwards by one lane
//   w3 = rank << 16 | y0 << 8 | x0          DFS rank (0xFFFF = not a view) and the parent's bounding-box origin
constexpr int CW = 4;

struct Smem {
  uint32_t plane[C2G_NLEV][PLANE_WORDS];  // bit (c & 31) of word r * WPR + (c >> 5): bev(r, c) > lv_grads[level]
  uint16_t fgpre[PLANE_WORDS];            // plane-0 bits before each word = index of the word's first cell in the foreground list
  union {
    uint16_t wpre[C2G_NLEV][PLANE_WORDS];  // number of runs that start before this word (labelling, parent lookup)
    struct {
      uint32_t key[NVL];   // (rank of the parent << 16) | first-2x2-block key
      uint16_t comp[NVL];
    } sig;                 // ranking step
    float fg_h[FG_CAP];    // moments: heights of the foreground cells
  };
  float fg_rf[FG_CAP], fg_cf[FG_CAP];  // continuous coordinates of the foreground cells (moments, key windows)
  union {
    uint32_t run_par[R_POOL];  // union-find parent (run id); after the flatten: root id; finally 0x80000000 | component
    unsigned char bci_scratch[K2_WARPS * 896];
  };
  union {
    uint32_t run_inf[R_POOL];  // row | c0 << 8 | len << 16
    float divs[N_ANCH][N_DIVS];
  };
  uint32_t comp[C_POOL * CW];
  uint32_t sortbuf[C2G_VIEW_CAP];  // (cell_cnt << 16 | presort index), all levels back to back
  uint16_t vcomp[C2G_VIEW_CAP];    // level-local component of every presort view
  uint16_t torder[C2G_VIEW_CAP];   // presort views by decreasing size class (moment tasks)
  int n_views[C2G_NLEV], view_off[C2G_NLEV], layer_cnt[C2G_NLEV];
  int n_runs[C2G_NLEV], run_off[C2G_NLEV], n_comp[C2G_NLEV], comp_off[C2G_NLEV];
  int half_runs[C2G_NLEV][2], half_roots[C2G_NLEV][2], comp_cnt[C2G_NLEV], cls_cnt[16];
  TopView top[C2G_NLEV][C2G_MAX_DIST_FIRSTS];
  int cnt_point[N_ANCH];
  int n_ell[C2G_NUM_BIN_LAYERS];
  int nsig, status, n_occ, n_fg, wq, wq2, next_scan;
  int runs_in_arena, comps_in_arena, arena;  // arena: this CTA's global overflow arena (= blockIdx.x) when in use, -1 = not used
  double red[K2_WARPS];
};
static_assert(sizeof(Smem) <= 113 * 1024, "two CTAs per SM");

// ---- union-find over run ids (links only go from a larger to a smaller id) ------------------------------------------
__device__ __forceinline__ uint32_t uf_find(volatile uint32_t *P, uint32_t c) {
  uint32_t p = P[c];
  while (p != c) {
    const uint32_t gp = P[p];
    if (gp != p) P[c] = gp;  // path halving: any ancestor is a valid parent
    c = p;
    p = gp;
  }
  return c;
}
// read-only variant for the flatten: there every lane stores the final root into its OWN entry, and a path-halving store
// from another lane could overwrite that root with a stale ancestor
__device__ __forceinline__ uint32_t uf_find_ro(const volatile uint32_t *P, uint32_t c) {
  uint32_t p = P[c];
  while (p != c) {
    c = p;
    p = P[c];
  }
  return c;
}
__device__ __forceinline__ void uf_union(uint32_t *P, uint32_t a, uint32_t b) {
  while (true) {
    a = uf_find(P, a);
    b = uf_find(P, b);
    if (a == b) return;
    if (a < b) {
      const uint32_t t = a;
      a = b;
      b = t;
    }
    const uint32_t old = atomicMin(&P[a], b);
    if (old == a) return;
    a = old;
  }
}
// component of a run once the roots carry 0x80000000 | component
__device__ __forceinline__ uint32_t comp_of(const uint32_t *P, uint32_t id) {
  uint32_t v = P[id];
  if (!(v & 0x80000000u)) v = P[v];
  return v & 0x7FFFFFFFu;
}
// the two warps of a level meet at named barrier 1 + level (barrier 0 is __syncthreads)
__device__ __forceinline__ void pair_sync(int level) {
  switch (level) {  // literal ids: ptxas reserves all 16 barriers for a register operand
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
  }
}
// bits of word `wi` of a plane row where a horizontal run starts
__device__ __forceinline__ uint32_t run_starts(const uint32_t *row_words, int wi) {
  const uint32_t bits = row_words[wi];
  const uint32_t prev = wi ? (row_words[wi - 1] >> 31) : 0u;
  return bits & ~((bits << 1) | prev);
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

// per-CTA slice of the global scratch: key-window lists of phase D (distance f32 + higher-level count u8 per listed cell)
constexpr size_t KLIST_BYTES = ((size_t) N_ANCH * KEY_LIST_CAP * 5 + 255) / 256 * 256;

__host__ __device__ inline int arena_level_cap(int n_cells, int n_row) { return n_cells / 2 + n_row; }  // runs (>= components) of one level
__host__ __device__ inline size_t arena_bytes(int n_cells, int n_row) {
  return (size_t) C2G_NLEV * arena_level_cap(n_cells, n_row) * (4 + 4 + 4 * CW);
}

static_assert(K2_WARPS >= 2 * C2G_NLEV, "two warps per level in the labelling phases (the other warps skip them)");

__global__ void __launch_bounds__(K2_THREADS, K2_CTAS_PER_SM)
contour_kernel(const uint32_t *__restrict__ planes_in, const float4 *__restrict__ fg_in, const int2 *__restrict__ hdr_in, int B,
               C2gIngestParams P, const int *__restrict__ int_ids, int first_slot, c2g_view *__restrict__ presort_scratch,
               c2g_scan_head *__restrict__ heads, c2g_view *__restrict__ views, c2g_ell *__restrict__ ells,
               unsigned char *__restrict__ klist_scratch, unsigned char *__restrict__ arenas, int force_arena,
               int *__restrict__ work_counter, long long *__restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = P.n_cells, ncol = P.cfg.n_col, nrow = P.cfg.n_row;
  const int WPR = (ncol + 31) >> 5, nwords = nrow * WPR;
  const c2g_cm_config &cfg = P.cfg;
  c2g_view *const presort = presort_scratch + (size_t) blockIdx.x * C2G_VIEW_CAP;
  float *const klist_dist = reinterpret_cast<float *>(klist_scratch + (size_t) blockIdx.x * KLIST_BYTES);  // [N_ANCH][KEY_LIST_CAP]
  uint8_t *const klist_hc = reinterpret_cast<uint8_t *>(klist_dist + N_ANCH * KEY_LIST_CAP);               // [N_ANCH][KEY_LIST_CAP]
  const int ARL = arena_level_cap(ncell, nrow);
  const bool lvl_warp = warp < 2 * C2G_NLEV;  // warps 0..11 label (two per level); further warps only join the block-wide phases
  const int lev_w = lvl_warp ? warp >> 1 : 0, half = warp & 1;  // level and half of this warp in the labelling phases
  const unsigned lt_mask = (1u << lane) - 1u;

  // scans are handed out dynamically (their cost varies 2x with the scene, and a CTA that starts late - e.g. behind a
  // co-running collective - must not leave a static share of the batch unprocessed until the end)
  while (true) {
    if (tid == 0) S.next_scan = atomicAdd(work_counter, 1);
    __syncthreads();
    const int b = S.next_scan;
    if (b >= B) break;
    const float4 *fg = fg_in + (size_t) b * ncell;  // (height, row_f, col_f, 0) of the foreground cells, raster order
    c2g_scan_head *head = heads + (first_slot + b);
    c2g_view *vout = views + (size_t) (first_slot + b) * C2G_VIEW_CAP;
    c2g_ell *eout = ells + (size_t) (first_slot + b) * C2G_VIEW_CAP;

#define C2G_DBG(i) do { if (dbg && blockIdx.x == 0 && tid == 0) dbg[i] = clock64(); } while (0)
    C2G_DBG(0);
    // ---------------- phase A: the scatter kernel's bit-planes -> shared memory (18 KB, coalesced)
    if (tid == 0) {
      S.status = 0;
      S.n_occ = hdr_in[b].x;
      S.n_fg = hdr_in[b].y;
      S.arena = -1;
      S.runs_in_arena = 0;
      S.comps_in_arena = 0;
    }
    {
      const uint32_t *pl_in = planes_in + (size_t) b * C2G_NLEV * nwords;
#pragma unroll
      for (int e = 0; e < C2G_NLEV; ++e)
        for (int w = tid; w < nwords; w += K2_THREADS) S.plane[e][w] = pl_in[e * nwords + w];
    }
    __syncthreads();
    C2G_DBG(1);
    // the warps that do not label stage the continuous coordinates of the foreground cells (12 B x ~1 100 cells) meanwhile;
    // a scan with more than FG_CAP such cells reads them from global memory instead
    const bool fg_smem = S.n_fg <= FG_CAP;
    if (warp > 2 * C2G_NLEV && fg_smem)
      for (int i = tid - (2 * C2G_NLEV + 1) * 32; i < S.n_fg; i += (K2_WARPS - 2 * C2G_NLEV - 1) * 32) {
        const float4 rec = __ldg(fg + i);
        S.fg_rf[i] = rec.y;
        S.fg_cf[i] = rec.z;
      }
    // ... and one of them indexes the list: exclusive prefix of the plane-0 popcounts
    if (warp == 2 * C2G_NLEV) {
      int base = 0;
      for (int w0 = 0; w0 < nwords; w0 += 32) {
        const int w = w0 + lane;
        const int cnt = w < nwords ? __popc(S.plane[0][w]) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        if (w < nwords) S.fgpre[w] = (uint16_t) (base + incl - cnt);
        base += __shfl_sync(FULL, incl, 31);
      }
    }

    // ---------------- phase B: six independent run-based labellings, two warps per level ------------------------------
    // B1 runs per word -> exclusive prefix (run ids are raster order); each warp scans half of the words
    const int wh = min(nwords, ((nwords / 2 + 31) >> 5) << 5);
    const int wbeg = half ? wh : 0, wend = half ? nwords : wh;
    if (lvl_warp) {
      const uint32_t *pl = S.plane[lev_w];
      int base = 0;
      for (int w0 = wbeg; w0 < wend; w0 += 32) {
        const int w = w0 + lane;
        int cnt = 0;
        if (w < wend) {
          const int row = w / WPR;
          cnt = __popc(run_starts(pl + row * WPR, w - row * WPR));
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        if (w < wend) S.wpre[lev_w][w] = (uint16_t) (base + incl - cnt);
        base += __shfl_sync(FULL, incl, 31);
      }
      if (lane == 0) S.half_runs[lev_w][half] = base;
      pair_sync(lev_w);
      if (half) {
        const int off = S.half_runs[lev_w][0];
        for (int w = wbeg + lane; w < wend; w += 32) S.wpre[lev_w][w] = (uint16_t) (S.wpre[lev_w][w] + off);
        if (lane == 0) S.n_runs[lev_w] = off + base;
      }
    }
    __syncthreads();
    C2G_DBG(10);
    if (tid == 0) {
      int tot = 0;
      for (int l = 0; l < C2G_NLEV; ++l) {
        S.run_off[l] = tot;
        tot += S.n_runs[l];
      }
      if (tot > R_POOL || force_arena) {  // the run tables of this scan live in this CTA's global arena (generic pointers, L2 latency)
        S.arena = blockIdx.x;
        S.runs_in_arena = 1;
        for (int l = 0; l < C2G_NLEV; ++l) S.run_off[l] = l * ARL;
      }
    }
    __syncthreads();

    // B2..B4; GEN = false: tables in shared memory (LDS / ATOMS), GEN = true: generic pointers (global arena)
    auto stage1 = [&](auto gen_tag) {
      constexpr bool GEN = decltype(gen_tag)::value;
      uint32_t *RP = S.run_par, *RI = S.run_inf;
      if (GEN) {
        RP = reinterpret_cast<uint32_t *>(arenas + (size_t) S.arena * arena_bytes(ncell, nrow));
        RI = RP + (size_t) C2G_NLEV * ARL;
      }
      const uint32_t *pl = S.plane[lev_w];
      const uint16_t *wp = S.wpre[lev_w];
      uint32_t *rp = RP + S.run_off[lev_w], *ri = RI + S.run_off[lev_w];
      const int n = S.n_runs[lev_w];
      // B2 run records
      for (int w = wbeg + lane; w < wend; w += 32) {
        const int row = w / WPR, wi = w - row * WPR;
        const uint32_t bits = pl[w];
        uint32_t starts = run_starts(pl + row * WPR, wi);
        uint32_t id = wp[w];
        while (starts) {
          const int bpos = __ffs(starts) - 1;
          starts &= starts - 1;
          const uint32_t t = ~(bits >> bpos);  // the bits shifted in from the top are zeros, i.e. ones of t
          int len = t ? __ffs(t) - 1 : 32;
          if (bpos + len == 32) {
            for (int w2 = wi + 1; w2 < WPR; ++w2) {
              const uint32_t t2 = ~pl[row * WPR + w2];
              const int add = t2 ? __ffs(t2) - 1 : 32;
              len += add;
              if (add < 32) break;
            }
          }
          ri[id] = (uint32_t) row | ((uint32_t) (wi * 32 + bpos) << 8) | ((uint32_t) len << 16);
          rp[id] = id;
          ++id;
        }
      }
      pair_sync(lev_w);
      C2G_DBG(12);
      // B3 unions with the runs of the row above that touch [c0 - 1, c0 + len] (8-connectivity): consecutive run ids.  The runs
      // are taken in raster-order batches of 64 (two warps) and every batch is flattened before the next one starts: links only go
      // to smaller ids, i.e. to earlier batches, whose entries then already name their roots - a find is one or two hops instead of
      // a walk down a chain as long as the component is tall (min-linking without the flatten builds such chains; the unions were
      // 45-90 kcyc of the ~600 kcyc a scan takes).
      for (int id0 = 0; id0 < n; id0 += 64) {
        const int id = id0 + half * 32 + lane;
        if (id < n) {
        const uint32_t inf = ri[id];
        const int row = inf & 255, c0 = (inf >> 8) & 255, len = (inf >> 16) & 255;
        if (row != 0) {
        const int lo = max(c0 - 1, 0), hi = min(c0 + len, ncol - 1);
        const uint32_t *pr = pl + (row - 1) * WPR;
        const uint16_t *wpr = wp + (row - 1) * WPR;
        int first = -1, extra = 0;
        const int w0 = lo >> 5;
        if ((hi >> 5) <= w0 + 1) {
          // common case, branch-free: the window lies inside two adjacent words of the row above
          const uint64_t two = (uint64_t) pr[w0] | ((uint64_t) (w0 + 1 < WPR ? pr[w0 + 1] : 0u) << 32);
          const uint64_t st = two & ~((two << 1) | (uint64_t) (w0 ? pr[w0 - 1] >> 31 : 0u));
          const int sh = lo & 31, wl = hi - lo + 1;
          const uint64_t m = two & ((wl >= 64 ? ~0ull : ((1ull << wl) - 1ull)) << sh);
          if (m) {
            const int pb = __ffsll((long long) m) - 1;
            const uint64_t upto = ~0ull >> (63 - pb);
            first = (int) wpr[w0] + __popcll(st & upto) - 1;  // the run that contains bit pb
            extra = __popcll(st & m & ~upto);                 // runs that start further right inside the window
          }
        } else {
          for (int wi = w0; wi <= (hi >> 5); ++wi) {
            uint32_t m = pr[wi];
            if (wi == w0) m &= FULL << (lo & 31);
            if (wi == (hi >> 5)) m &= FULL >> (31 - (hi & 31));
            if (!m) continue;
            const uint32_t st = run_starts(pr, wi);
            if (first < 0) {
              const int pb = __ffs(m) - 1;
              first = (int) wpr[wi] + __popc(st & (FULL >> (31 - pb))) - 1;
              extra += __popc(st & m & ~(FULL >> (31 - pb)));
            } else
              extra += __popc(st & m);
          }
        }
        if (first >= 0)
          for (int j = 0; j <= extra; ++j) uf_union(rp, (uint32_t) id, (uint32_t) (first + j));
        }
        }
        pair_sync(lev_w);
        if (id < n) rp[id] = uf_find_ro(rp, (uint32_t) id);  // own entry only; other lanes read either the old parent or the root
        pair_sync(lev_w);
      }
      pair_sync(lev_w);
      C2G_DBG(13);
      // B4 flatten: compressing finds first (no entry is final yet), then one read-only find per run
      for (int id = half * 32 + lane; id < n; id += 64) (void) uf_find(rp, (uint32_t) id);
      pair_sync(lev_w);
      int nroot = 0;
      for (int id0 = half * 32; id0 < n; id0 += 64) {
        const int id = id0 + lane;
        bool is_root = false;
        if (id < n) {
          const uint32_t root = uf_find_ro(rp, (uint32_t) id);
          rp[id] = root;
          is_root = root == (uint32_t) id;
        }
        nroot += __popc(__ballot_sync(FULL, is_root));
      }
      if (lane == 0) S.half_roots[lev_w][half] = nroot;
      C2G_DBG(14);
    };
    if (lvl_warp) {
      if (S.runs_in_arena)
        stage1(std::true_type{});
      else
        stage1(std::false_type{});
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int l = 0; l < C2G_NLEV; ++l) {
        S.n_comp[l] = S.half_roots[l][0] + S.half_roots[l][1];
        S.comp_cnt[l] = 0;
        S.comp_off[l] = tot;
        tot += S.n_comp[l];
      }
      if (tot > C_POOL || force_arena) {
        S.arena = blockIdx.x;
        S.comps_in_arena = 1;
        for (int l = 0; l < C2G_NLEV; ++l) S.comp_off[l] = l * ARL;
      }
    }
    __syncthreads();
    C2G_DBG(15);

    int total_views = 0;
    // B5..B9, moments and calcStatVals: everything that reads the run / component tables
    auto stage2 = [&](auto gen_tag) {
      constexpr bool GEN = decltype(gen_tag)::value;
      uint32_t *RP = S.run_par, *RI = S.run_inf, *CWP = S.comp;
      if (GEN) {
        unsigned char *ab = arenas + (size_t) (S.arena < 0 ? 0 : S.arena) * arena_bytes(ncell, nrow);
        if (S.runs_in_arena) {
          RP = reinterpret_cast<uint32_t *>(ab);
          RI = RP + (size_t) C2G_NLEV * ARL;
        }
        if (S.comps_in_arena) CWP = reinterpret_cast<uint32_t *>(ab + (size_t) C2G_NLEV * ARL * 8);
      }
      {
        uint32_t *rp = RP + S.run_off[lev_w], *ri = RI + S.run_off[lev_w];
        uint32_t *cw = CWP + (size_t) S.comp_off[lev_w] * CW;
        const int n = lvl_warp ? S.n_runs[lev_w] : 0, nc = lvl_warp ? S.n_comp[lev_w] : 0;
        // B5 roots -> components (numbered in arrival order; only the DFS rank computed below carries meaning)
        for (int id0 = half * 32; id0 < n; id0 += 64) {
          const int id = id0 + lane;
          const bool is_root = id < n && rp[id] == (uint32_t) id;
          const unsigned bal = __ballot_sync(FULL, is_root);
          int cb = 0;
          if (lane == 0 && bal) cb = atomicAdd(&S.comp_cnt[lev_w], __popc(bal));
          cb = __shfl_sync(FULL, cb, 0);
          if (is_root) {
            const int ci = cb + __popc(bal & lt_mask);
            cw[ci * CW + 0] = (uint32_t) id << 16;
            cw[ci * CW + 1] = ((uint32_t) id << 16) | 0xFFFFu;
            cw[ci * CW + 2] = 0x00FFFFFFu;
            cw[ci * CW + 3] = 0xFFFF0000u;
            rp[id] = 0x80000000u | (uint32_t) ci;
          }
        }
        if (lvl_warp) pair_sync(lev_w);
        C2G_DBG(16);
        // B6 area / last run / leftmost column per component; every run now names its component directly
        for (int id = half * 32 + lane; id < n; id += 64) {
          const uint32_t inf = ri[id];
          const uint32_t c0 = (inf >> 8) & 255, len = (inf >> 16) & 255;
          const uint32_t ci = comp_of(rp, (uint32_t) id);
          atomicAdd(&cw[ci * CW + 0], len);
          atomicMax(&cw[ci * CW + 1], ((uint32_t) id << 16) | 0xFFFFu);
          atomicMax(&cw[ci * CW + 2], ((255u - c0) << 24) | 0x00FFFFFFu);
          rp[id] = 0x80000000u | ci;  // nobody reads a non-root entry but its own lane
        }
        __syncthreads();
        C2G_DBG(2);
        // B7 enclosing component of the previous level (looked up at the first pixel) and its bounding-box origin
        if (lvl_warp && lev_w > 0) {
          const uint32_t *plp = S.plane[lev_w - 1];
          const uint16_t *wpp = S.wpre[lev_w - 1];
          const uint32_t *rpp = RP + S.run_off[lev_w - 1], *rip = RI + S.run_off[lev_w - 1];
          const uint32_t *cwp = CWP + (size_t) S.comp_off[lev_w - 1] * CW;
          for (int ci = half * 32 + lane; ci < nc; ci += 64) {
            const uint32_t inf = ri[cw[ci * CW] >> 16];
            const int row = inf & 255, c0 = (inf >> 8) & 255;
            const int wi = c0 >> 5;
            if (!((plp[row * WPR + wi] >> (c0 & 31)) & 1u)) continue;  // only if lv_grads is not increasing: no parent
            const uint32_t st = run_starts(plp + row * WPR, wi);
            const int rid = (int) wpp[row * WPR + wi] + __popc(st & (FULL >> (31 - (c0 & 31)))) - 1;
            const uint32_t pc = rpp[rid] & 0x7FFFFFFFu;
            const uint32_t y0 = rip[cwp[pc * CW] >> 16] & 255u, x0 = 255u - (cwp[pc * CW + 2] >> 24);
            cw[ci * CW + 2] = (cw[ci * CW + 2] & 0xFF000000u) | pc;
            cw[ci * CW + 3] = 0xFFFF0000u | (y0 << 8) | x0;
          }
        }
        if (lvl_warp) pair_sync(lev_w);
        C2G_DBG(18);
        // B8 first-2x2-block key relative to that origin: min over the runs (a run's minimum is at its first cell)
        for (int id = half * 32 + lane; id < n; id += 64) {
          const uint32_t inf = ri[id];
          const uint32_t ci = rp[id] & 0x7FFFFFFFu;
          const uint32_t w3 = cw[ci * CW + 3];
          const int key = ((((int) (inf & 255) - (int) ((w3 >> 8) & 255)) >> 1) << 7) + (((int) ((inf >> 8) & 255) - (int) (w3 & 255)) >> 1);
          const uint32_t w1 = *(volatile uint32_t *) &cw[ci * CW + 1];
          if ((uint32_t) key < (w1 & 0xFFFFu)) atomicMin(&cw[ci * CW + 1], (w1 & 0xFFFF0000u) | (uint32_t) key);
        }
      }
      __syncthreads();
      C2G_DBG(19);
      // B9 DFS order rank, level by level (the rank of a component needs the rank of its parent)
      for (int lev = 0; lev < C2G_NLEV; ++lev) {
        uint32_t *cw = CWP + (size_t) S.comp_off[lev] * CW;
        const uint32_t *cwp = lev ? CWP + (size_t) S.comp_off[lev - 1] * CW : nullptr;
        const int nc = S.n_comp[lev];
        if (tid == 0) S.nsig = 0;
        __syncthreads();
        for (int ci = tid; ci < nc; ci += K2_THREADS) {
          const uint32_t w0 = cw[ci * CW];
          if ((int) (w0 & 0xFFFFu) >= cfg.min_cont_cell_cnt) {
            const int i = atomicAdd(&S.nsig, 1);
            if (i < NVL) {
              const uint32_t pc = cw[ci * CW + 2] & 0xFFFFu;
              const uint32_t prank = !lev ? 0u : pc == 0xFFFFu ? 0xFFFFu : (cwp[pc * CW + 3] >> 16);
              S.sig.key[i] = (prank << 16) | (cw[ci * CW + 1] & 0xFFFFu);
              S.sig.comp[i] = (uint16_t) ci;
            } else
              atomicOr(&S.status, 2);
          }
        }
        __syncthreads();
        const int nsig_all = min(S.nsig, NVL);
        int nsig = nsig_all;
        if (total_views + nsig > C2G_VIEW_CAP) {
          nsig = C2G_VIEW_CAP - total_views;
          if (tid == 0) atomicOr(&S.status, 1);
        }
        for (int i = tid; i < nsig_all; i += K2_THREADS) {
          const uint32_t ki = S.sig.key[i];
          int rank = 0;
          for (int j = 0; j < nsig_all; ++j) rank += (S.sig.key[j] < ki) ? 1 : 0;
          if (rank < nsig) {
            const int ci = S.sig.comp[i];
            cw[ci * CW + 3] = ((uint32_t) rank << 16) | (cw[ci * CW + 3] & 0xFFFFu);
            S.sortbuf[total_views + rank] = ((cw[ci * CW] & 0xFFFFu) << 16) | (uint32_t) rank;
            S.vcomp[total_views + rank] = (uint16_t) ci;
          }
        }
        if (tid == 0) {
          S.n_views[lev] = nsig;
          S.view_off[lev] = total_views;
        }
        total_views += nsig;
        __syncthreads();
      }
      C2G_DBG(3);

      // ---------------- phase C: per-level std::sort replays + moments / calcStatVals per component ------------------
      // moment tasks by decreasing size class floor(log2(area)): the four components a warp walks together are alike
      if (tid < 16) S.cls_cnt[tid] = 0;
      if (tid == 0) S.wq = S.wq2 = 0;
      __syncthreads();
      for (int v = tid; v < total_views; v += K2_THREADS) atomicAdd(&S.cls_cnt[__clz(S.sortbuf[v] >> 16) - 16], 1);
      __syncthreads();
      if (tid == 0) {
        int run = 0;
        for (int i = 0; i < 16; ++i) {
          const int c = S.cls_cnt[i];
          S.cls_cnt[i] = run;
          run += c;
        }
      }
      __syncthreads();
      for (int v = tid; v < total_views; v += K2_THREADS) S.torder[atomicAdd(&S.cls_cnt[__clz(S.sortbuf[v] >> 16) - 16], 1)] = (uint16_t) v;
      __syncthreads();
      // sortbuf words of one level are (area << 16 | rank): sorting them in place is the std::sort of cont_views_[level]
      // (contour_mng.h:596-599).  One WARP per level replays libstdc++'s introsort cooperatively (stdsort.cuh: same permutation as
      // the serial replay, which was the critical path of this kernel); the ranking scratch is dead and serves as its work space.
      if (half == 0 && lvl_warp) {
        uint32_t *first = S.sortbuf + S.view_off[lev_w];
        const int n = S.n_views[lev_w];
        int sum = 0;
        for (int i = lane; i < n; i += 32) sum += (int) (first[i] >> 16);
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if (lane == 0) S.layer_cnt[lev_w] = sum;
        uint16_t *scr = reinterpret_cast<uint16_t *>(&S.wpre[0][0]);
        static_assert(sizeof(S.wpre) >= 2 * C2G_VIEW_CAP * sizeof(uint16_t), "sort scratch");
        c2g_sort::warp_std_sort<true>(first, n, scr + S.view_off[lev_w], scr + C2G_VIEW_CAP + S.view_off[lev_w], lane);
        C2G_DBG(20);
      }
      __syncthreads();
      if (fg_smem)  // ... and then takes the heights of the foreground cells
        for (int i = tid; i < S.n_fg; i += K2_THREADS) S.fg_h[i] = __ldg(fg + i).x;
      __syncthreads();
      {
        // Moments + calcStatVals.  Every accumulator of RunningStatRecorder is a sequential sum over the component's cells in
        // raster order (bit-exactness: exactly the reference's sequence of double additions; x * 1.0 is exact, the product of two
        // floats is exact in double), so a component offers at most eight-way parallelism: its eight accumulators.  The runs of a
        // component are found by scanning the run ids between its first and last run; a run's cells are a contiguous slice of the
        // foreground list.  Views are taken in decreasing size class (torder):
        //  * the n_big views of >= 64 cells: EIGHT LANES per component, lane k keeps accumulator k (sum of a_k * b_k with
        //    (a, b) = (v0,1) (v1,1) (v0,v0) (v0,v1) (v1,v1) (h,v0) (h,v1), lane 7 the float sum of h), the group tests eight run ids
        //    per step and keeps four loads in flight; four components per warp step;
        //  * the long tail of small views: ONE LANE per component with all accumulators (seven independent chains) - the lanes of a
        //    warp walk components of similar size, so the nested run / cell loops diverge little.  (One lane per component for the
        //    big ones too was measured: the warp holding the 32 largest components ran 5x longer than the rest of the phase.)
        // (height, row_f, col_f) of foreground cell i = ph[i * fs], prf[i * fs], pcf[i * fs]: the shared-memory copies (stride 1) or,
        // for a scan with more than FG_CAP foreground cells, the global records (stride 4 floats)
        const float *gfg = reinterpret_cast<const float *>(fg);
        const int fs = fg_smem ? 1 : 4;
        const float *ph = fg_smem ? S.fg_h : gfg, *prf = fg_smem ? S.fg_rf : gfg + 1, *pcf = fg_smem ? S.fg_cf : gfg + 2;
        auto view_level = [&](int v) {
          int lev = 0;
#pragma unroll
          for (int l = 1; l < C2G_NLEV; ++l)
            if (v >= S.view_off[l]) lev = l;
          return lev;
        };
        const int n_big = min(S.cls_cnt[9], total_views);  // size classes 0..9 = areas >= 64 (cls_cnt holds the class END offsets now)
        const int k8 = lane & 7, g4 = lane >> 3;
        const float *pa = (k8 == 0 || k8 == 2 || k8 == 3) ? prf : (k8 == 1 || k8 == 4) ? pcf : ph;
        const float *pb = (k8 == 2 || k8 == 5) ? prf : pcf;
        const bool b_one = k8 == 0 || k8 == 1 || k8 == 7;
        while (true) {
          int t0 = 0;
          if (lane == 0) t0 = atomicAdd(&S.wq, 4);
          t0 = __shfl_sync(FULL, t0, 0);
          if (t0 >= n_big) break;
          const int t = t0 + g4;
          bool active = t < n_big;
          int v = 0, lev = 0;
          uint32_t ci = 0, rid = 0, last = 0;
          const uint32_t *rp = RP, *ri = RI;
          const uint32_t *cw = CWP;
          if (active) {
            v = S.torder[t];
            lev = view_level(v);
            ci = S.vcomp[v];
            cw = CWP + ((size_t) S.comp_off[lev] + ci) * CW;
            rid = cw[0] >> 16;
            last = cw[1] >> 16;
            rp = RP + S.run_off[lev];
            ri = RI + S.run_off[lev];
          }
          double acc = 0.0;
          float vol3 = 0.0f;
          while (__any_sync(FULL, active)) {
            const uint32_t cand = rid + k8;
            const bool mem = active && cand <= last && (rp[cand] & 0x7FFFFFFFu) == ci;
            const unsigned bal = __ballot_sync(FULL, mem);
            unsigned gb = (bal >> (g4 * 8)) & 0xFFu;
            while (gb) {
              const int tt = __ffs(gb) - 1;
              gb &= gb - 1;
              const uint32_t inf = ri[rid + tt];
              const int row = (int) (inf & 255), c0 = (int) ((inf >> 8) & 255), len = (int) ((inf >> 16) & 255);
              const int wd = row * WPR + (c0 >> 5);
              const int cell0 = (int) S.fgpre[wd] + __popc(S.plane[0][wd] & ((1u << (c0 & 31)) - 1u));
              for (int j0 = 0; j0 < len; j0 += 4) {  // four independent loads in flight; a padding term adds 0 * 1 = +0.0
                float va[4], vb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const bool in = j0 + u < len;
                  va[u] = in ? pa[(cell0 + j0 + u) * fs] : 0.0f;
                  vb[u] = (in && !b_one) ? pb[(cell0 + j0 + u) * fs] : 1.0f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  acc += (double) va[u] * (double) vb[u];
                  vol3 += va[u];
                }
              }
            }
            rid += 8;
            if (rid > last) active = false;
          }
          // lane 0 of the group collects the eight accumulators and finishes the view
          Moments m;
          m.s0 = __shfl_sync(FULL, acc, g4 * 8 + 0);
          m.s1 = __shfl_sync(FULL, acc, g4 * 8 + 1);
          m.t00 = __shfl_sync(FULL, acc, g4 * 8 + 2);
          m.t01 = __shfl_sync(FULL, acc, g4 * 8 + 3);
          m.t11 = __shfl_sync(FULL, acc, g4 * 8 + 4);
          m.q0 = __shfl_sync(FULL, acc, g4 * 8 + 5);
          m.q1 = __shfl_sync(FULL, acc, g4 * 8 + 6);
          m.vol3 = __shfl_sync(FULL, vol3, g4 * 8 + 7);
          if (k8 == 0 && t < n_big) {
            m.cnt = (int) (cw[0] & 0xFFFFu);
            const uint32_t linf = ri[last];
            c2g_view vw;
            calc_stat_vals(m, cfg, lev, (int) (linf & 255), (int) ((linf >> 8) & 255) + (int) ((linf >> 16) & 255) - 1, vw);
            presort[v] = vw;
          }
        }
        __syncwarp();
        while (true) {
          int t0 = 0;
          if (lane == 0) t0 = atomicAdd(&S.wq2, 32);
          t0 = __shfl_sync(FULL, t0, 0) + n_big;
          if (t0 >= total_views) break;
          const int t = t0 + lane;
          const bool have = t < total_views;
          int v = 0, lev = 0;
          uint32_t ci = 0, id = 1, last = 0;
          const uint32_t *rp = RP, *ri = RI, *cw = CWP;
          if (have) {
            v = S.torder[t];
            lev = view_level(v);
            ci = S.vcomp[v];
            cw = CWP + ((size_t) S.comp_off[lev] + ci) * CW;
            id = cw[0] >> 16;  // the first run of a component is one of its own
            last = cw[1] >> 16;
            rp = RP + S.run_off[lev];
            ri = RI + S.run_off[lev];
          }
          Moments m;
          m.cnt = have ? (int) (cw[0] & 0xFFFFu) : 0;
          m.s0 = m.s1 = m.t00 = m.t01 = m.t11 = m.q0 = m.q1 = 0.0;
          m.vol3 = 0.0f;
          // One flat loop instead of the nested run / cell loops: in every step each lane either consumes ONE cell of its current run
          // or moves on to its next member run, so the lanes of the warp advance together whatever the shapes of their components
          // (nested loops made the warp execute the union of all lanes' trip counts: 3.6 active lanes on average).
          int cell = 0, left = 0;  // next foreground-list index / cells left in the current run
          bool busy = have;
          while (__any_sync(FULL, busy)) {
            if (busy && left == 0) {
              while (id <= last && (rp[id] & 0x7FFFFFFFu) != ci) ++id;
              if (id > last)
                busy = false;
              else {
                const uint32_t inf = ri[id];
                const int row = (int) (inf & 255), c0 = (int) ((inf >> 8) & 255);
                const int wd = row * WPR + (c0 >> 5);
                left = (int) ((inf >> 16) & 255);
                cell = (int) S.fgpre[wd] + __popc(S.plane[0][wd] & ((1u << (c0 & 31)) - 1u));
                ++id;
              }
            }
            if (busy) {
              const float h = ph[cell * fs];
              const double v0 = (double) prf[cell * fs], v1 = (double) pcf[cell * fs], hh = (double) h;
              m.s0 += v0;
              m.s1 += v1;
              m.t00 += v0 * v0;
              m.t01 += v0 * v1;
              m.t11 += v1 * v1;
              m.q0 += hh * v0;
              m.q1 += hh * v1;
              m.vol3 += h;
              ++cell;
              --left;
            }
          }
          if (have) {
            const uint32_t linf = ri[last];
            c2g_view vw;
            calc_stat_vals(m, cfg, lev, (int) (linf & 255), (int) ((linf >> 8) & 255) + (int) ((linf >> 16) & 255) - 1, vw);
            presort[v] = vw;
          }
        }
      }
      C2G_DBG(21);
      if (dbg && blockIdx.x == 0 && lane == 0) dbg[32 + warp] = clock64();
      __syncthreads();
      if (dbg && blockIdx.x == 0 && lane == 0) dbg[48 + warp] = clock64();
      C2G_DBG(4);
    };
    if (S.runs_in_arena || S.comps_in_arena)
      stage2(std::true_type{});
    else
      stage2(std::false_type{});
    __syncthreads();
    C2G_DBG(24);
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
      C2G_DBG(22);
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
      C2G_DBG(23);
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

    C2G_DBG(5);
    // ---------------- phase D: retrieval keys (contour_mng.h:693-830) ------------------------------------------------
    // D1: per anchor, ordered list of the window cells that contribute: (dist, higher_cnt). One warp per anchor.
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
        const int nr = r_max - r_min + 1, wbits = c_max - c_min + 1;  // <= 23 x 23 (roi_radius <= 10, checked at c2g_create)
        const double rad = (double) cfg.roi_radius - 1e-2;
        // One lane per window ROW: the row's cells above lv_grads[1] are the set bits of plane 1 inside the column range
        // (cells with bev == lv_grads[1] pass the reference's first test `bev < lv_grads[1] -> skip` but fail its second one,
        // `bev > lv_grads[1]`), so only those ~2-5 cells per row are looked at instead of all 529 window cells.  Raster order
        // of the list = row-major: a warp-wide exclusive scan of the per-row counts places every row's cells.
        const int rr = r_min + lane;
        uint32_t m1 = 0;  // bit i: cell (rr, c_min + i) is above lv_grads[1]
        if (lane < nr) {
          const int w0 = c_min >> 5, sh = c_min & 31;
          const uint32_t *p1 = S.plane[1] + rr * WPR;
          const uint64_t two = (uint64_t) p1[w0] | ((uint64_t) (w0 + 1 < WPR ? p1[w0 + 1] : 0u) << 32);
          m1 = (uint32_t) (two >> sh) & (wbits >= 32 ? FULL : ((1u << wbits) - 1u));
        }
        auto cell_rec = [&](int i) -> float4 {  // foreground record of window column i of this lane's row (.y, .z valid)
          const int cc = c_min + i, wd = rr * WPR + (cc >> 5);
          const int fi = (int) S.fgpre[wd] + __popc(S.plane[0][wd] & ((1u << (cc & 31)) - 1u));
          if (fg_smem) return make_float4(0.f, S.fg_rf[fi], S.fg_cf[fi], 0.f);
          return __ldg(fg + fi);
        };
        uint32_t pm = 0;  // ... and lies inside the radius
        for (uint32_t m = m1; m; m &= m - 1) {
          const int i = __ffs(m) - 1;
          const float4 rec = cell_rec(i);
          const float dx = rec.y - cx, dy = rec.z - cy;
          const float dist = sqrtf(dx * dx + dy * dy);
          if ((double) dist < rad) pm |= 1u << i;
        }
        const int mine = __popc(pm);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        cnt = __shfl_sync(FULL, incl, 31);
        int pos = incl - mine;
        for (uint32_t m = pm; m; m &= m - 1, ++pos) {
          const int i = __ffs(m) - 1;
          if (pos >= KEY_LIST_CAP) break;
          const float4 rec = cell_rec(i);
          const float dx = rec.y - cx, dy = rec.z - cy;
          const int cc = c_min + i, wd = rr * WPR + (cc >> 5), sh = cc & 31;
          int h2 = 1;
#pragma unroll
          for (int l = 2; l < C2G_NLEV; ++l) h2 += (int) ((S.plane[l][wd] >> sh) & 1u);
          klist_dist[a * KEY_LIST_CAP + pos] = sqrtf(dx * dx + dy * dy);
          klist_hc[a * KEY_LIST_CAP + pos] = (uint8_t) h2;
        }
        if (cnt > KEY_LIST_CAP && lane == 0) atomicOr(&S.status, 4);
      }
      if (lane == 0) S.cnt_point[a] = valid ? cnt : -1;
    }
    __syncthreads();
    C2G_DBG(6);
    // D2: one thread per (anchor, division): sequential float accumulation in raster order. The exp evaluations of four
    // consecutive cells are independent straight-line code (the FP64 latency chains overlap); their float terms are then
    // added in order.
    {
      const float div_len = cfg.roi_radius / (float) ((C2G_KEY_DIM - 3) * 5);
      const double inv_norm_den = sqrt(2 * 3.14159265358979323846 * 1.0f * 1.0f);
      const double rcp_norm_den = 1.0 / inv_norm_den;
      // (float) (e / den): the product with the rounded reciprocal is within 2.5 ulp of the correctly rounded quotient, so
      // both round to the same float unless the product sits within a few ulp of a float rounding midpoint (low 29
      // mantissa bits == 0x10000000); only then is the real division executed.
      auto near_mid = [](double pq) { return ((unsigned long long) __double_as_longlong(pq) & 0x1FFFFFFFull) - 0x0FFFFFF8ull <= 0x10ull; };
      auto d2 = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        for (int w = tid; w < N_ANCH * N_DIVS; w += K2_THREADS) {
          const int a = w / N_DIVS, d = w - a * N_DIVS;
          const int n = min(S.cnt_point[a], KEY_LIST_CAP);
          float acc = 0.0f;
          if (n > 0) {
            const float x = (float) ((double) ((float) d * div_len) + 0.5 * (double) div_len);
            const float *dl = klist_dist + a * KEY_LIST_CAP;
            const uint8_t *hl = klist_hc + a * KEY_LIST_CAP;
            int k = 0;
            if (MODE != 0) {
              for (; k + 4 <= n; k += 4) {
                double q[4], pq[4];
                bool special = false, mid = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float t = (x - dl[k + u]) / 1.0f;
                  q[u] = (-0.5 * (double) t) * (double) t;
                  pq[u] = c2g_exp_glibc_main<MODE == 2>(q[u], c2g_exp_tab_dev, &special) * rcp_norm_den;
                  mid |= near_mid(pq[u]);
                }
                if (special || mid) {  // rare: redo the four terms through the full functions
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const double e = c2g_exp(q[u], MODE, c2g_exp_tab_dev);
                    pq[u] = e * rcp_norm_den;
                    if (near_mid(pq[u])) pq[u] = e / inv_norm_den;
                  }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) acc += (float) hl[k + u] * (float) pq[u];
              }
            }
            for (; k < n; ++k) {
              const float t = (x - dl[k]) / 1.0f;
              const double q = (-0.5 * (double) t) * (double) t;
              const double e = c2g_exp(q, MODE, c2g_exp_tab_dev);
              double pq = e * rcp_norm_den;
              if (near_mid(pq)) pq = e / inv_norm_den;
              acc += (float) hl[k] * (float) pq;
            }
          }
          S.divs[a][d] = acc;
        }
      };
      if (P.exp_mode == 2)
        d2(std::integral_constant<int, 2>{});
      else if (P.exp_mode == 1)
        d2(std::integral_constant<int, 1>{});
      else
        d2(std::integral_constant<int, 0>{});
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

    C2G_DBG(7);
    // ---------------- phase E: BCIs (contour_mng.h:848-883), one warp per anchor ------------------------------------
    // candidate t = bl * 10 + j (reference loop order) is evaluated by lane t % 32; kept neighbours are compacted in
    // that order, lane 0 replays std::sort on bit_pos and builds the run boundaries, all lanes write the record.
    {
      unsigned char *scr = S.bci_scratch + warp * 896;  // the run tables are dead: [40] relpt, [40] u32, [42] u16, 2 x [40] u16 sort scratch
      c2g_relpt *nei = reinterpret_cast<c2g_relpt *>(scr);
      uint32_t *ord = reinterpret_cast<uint32_t *>(scr + C2G_MAX_NEI * sizeof(c2g_relpt));
      uint16_t *segv = reinterpret_cast<uint16_t *>(scr + C2G_MAX_NEI * (sizeof(c2g_relpt) + 4));
      uint16_t *sortL = segv + (C2G_MAX_NEI + 2), *sortR = sortL + C2G_MAX_NEI;
      static_assert(C2G_MAX_NEI * (sizeof(c2g_relpt) + 4) + (C2G_MAX_NEI + 2) * 2 + 2 * C2G_MAX_NEI * 2 <= 896, "BCI scratch");
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
            const unsigned bal = __ballot_sync(FULL, keep);
            if (keep) {
              const int pos = n + __popc(bal & ((1u << lane) - 1u));
              nei[pos] = rp;
              ord[pos] = ((uint32_t) (uint16_t) rp.bit_pos << 16) | (uint32_t) pos;
            }
            n += __popc(bal);
          }
        }
        __syncwarp();
        c2g_sort::warp_std_sort<false>(ord, n, sortL, sortR, lane);  // std::sort on bit_pos (contour_mng.h:871-874), the whole warp
        int nseg = 0;
        uint64_t bins[4] = {0, 0, 0, 0};
        if (lane == 0) {
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
        nseg = __shfl_sync(FULL, nseg, 0);
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
    C2G_DBG(8);
    // ---------------- phase F: scan-only GMM terms (correlation.h:49-82,102-119) -----------------------------------
    if (tid < C2G_NUM_BIN_LAYERS) {
      const int lev = tid + 1;
      const int full = S.layer_cnt[lev];
      int run = 0, k = 0;
      const uint32_t *sb = S.sortbuf + S.view_off[lev];
      for (; k < S.n_views[lev]; ++k) {
        if ((double) run * 1.0 / (double) full >= 0.95) break;
        run += (int) (sb[k] >> 16);
      }
      S.n_ell[tid] = k;
    }
    __syncthreads();
    // The pair term is symmetric in (i, j) bit for bit, so only j >= i is evaluated (off-diagonal terms count twice); rows
    // i and n - 1 - i are folded into one line of n + 1 entries to keep the index arithmetic division-only.
    double acc = 0.0;
    for (int li = 0; li < C2G_NUM_BIN_LAYERS; ++li) {
      const int n = S.n_ell[li];
      const c2g_ell *le = eout + S.view_off[li + 1];
      const int R = n >> 1, folded = R * (n + 1), total = folded + (n & 1) * ((n + 1) >> 1);
      for (int w = tid; w < total; w += K2_THREADS) {
        int i, j;
        if (w < folded) {
          const int r = w / (n + 1), c = w - r * (n + 1);
          if (c < n - r) {
            i = r;
            j = r + c;
          } else {
            i = n - 1 - r;
            j = i + (c - (n - r));
          }
        } else {
          i = R;
          j = R + (w - folded);
        }
        const c2g_ell A = le[i], Bv = le[j];
        const double c00 = 2.0 * ((double) A.c00 + (double) Bv.c00), c10 = 2.0 * ((double) A.c10 + (double) Bv.c10);
        const double c01 = 2.0 * ((double) A.c01 + (double) Bv.c01), c11 = 2.0 * ((double) A.c11 + (double) Bv.c11);
        const double mx = (double) A.mx - (double) Bv.mx, my = (double) A.my - (double) Bv.my;
        const double det = c00 * c11 - c01 * c10;
        const double invdet = 1.0 / det;
        const double qf = mx * ((c11 * invdet) * mx + (-c01 * invdet) * my) + my * ((-c10 * invdet) * mx + (c00 * invdet) * my);
        const double term = (double) A.w * (double) Bv.w / sqrt(det) * c2g_exp(-0.5 * qf, P.exp_mode, c2g_exp_tab_dev);
        acc += (i == j) ? term : 2.0 * term;
      }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (lane == 0) S.red[warp] = acc;
    __syncthreads();
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
      for (int l = 0; l < C2G_NUM_BIN_LAYERS; ++l) head->n_ell[l] = S.n_ell[l];
      head->n_occupied = S.n_occ;
      head->pad_ = 0;
      head->gmm_auto_corr = tot;
    }
    __syncthreads();
    C2G_DBG(9);
  }
}


}  // namespace

// test hook: the warp-cooperative std::sort replay on `n` packed words (one warp, shared memory like in the contour kernel)
namespace {
template <bool DESC>
__global__ void warp_sort_selftest_kernel(uint32_t *words, int n) {
  extern __shared__ __align__(16) unsigned char ws_raw[];
  uint32_t *a = reinterpret_cast<uint32_t *>(ws_raw);
  uint16_t *pl = reinterpret_cast<uint16_t *>(a + n), *pr = pl + n;
  for (int i = threadIdx.x; i < n; i += 32) a[i] = words[i];
  __syncwarp();
  c2g_sort::warp_std_sort<DESC>(a, n, pl, pr, threadIdx.x);
  for (int i = threadIdx.x; i < n; i += 32) words[i] = a[i];
}
}  // namespace
int c2g_launch_warp_sort_selftest(uint32_t *words_dev, int n, int desc, cudaStream_t stream) {
  const size_t smem = (size_t) n * 8 + 16;
  if (smem > 200 * 1024) return -1000;
  if (desc) {
    C2G_CUDA_TRY(cudaFuncSetAttribute(warp_sort_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    warp_sort_selftest_kernel<true><<<1, 32, smem, stream>>>(words_dev, n);
  } else {
    C2G_CUDA_TRY(cudaFuncSetAttribute(warp_sort_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    warp_sort_selftest_kernel<false><<<1, 32, smem, stream>>>(words_dev, n);
  }
  C2G_CUDA_TRY(cudaGetLastError());
  return 0;
}

size_t c2g_contour_smem_bytes() { return sizeof(Smem); }
int c2g_contour_max_ctas(int num_sms) { return K2_CTAS_PER_SM * num_sms; }
// layout of the global scratch: [max_ctas][KLIST_BYTES] key-window lists | [max_ctas] overflow arenas (one per resident CTA: a batch of
// cluttered scans must not serialise on a shared pool)
size_t c2g_contour_scratch_bytes(int num_sms, int n_cells, int n_row) {
  return (size_t) c2g_contour_max_ctas(num_sms) * (KLIST_BYTES + arena_bytes(n_cells, n_row));
}

int c2g_launch_contours(const C2gBevOut &bev, int B, const C2gIngestParams &P, const int *int_ids_dev, int first_slot,
                        c2g_view *presort_scratch, c2g_scan_head *heads, c2g_view *views, c2g_ell *ells, unsigned char *k2_scratch,
                        int *work_counter, int num_sms, cudaStream_t stream, long long *dbg) {
  static unsigned long long attr_devs = 0ull;
  if (c2g_first_use_on_device(attr_devs)) {
    C2G_CUDA_TRY(cudaFuncSetAttribute(contour_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(Smem)));
    C2G_CUDA_TRY(cudaFuncSetAttribute(contour_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const int max_ctas = c2g_contour_max_ctas(num_sms);
  const int grid = B < max_ctas ? B : max_ctas;
  if (grid <= 0) return 0;
  unsigned char *klists = k2_scratch;
  unsigned char *arenas = k2_scratch + (size_t) max_ctas * KLIST_BYTES;
  static const int force_arena = getenv("C2G_K2_FORCE_ARENA") ? atoi(getenv("C2G_K2_FORCE_ARENA")) : 0;  // measurement / test hook
  C2G_CUDA_TRY(cudaMemsetAsync(work_counter, 0, sizeof(int), stream));
  contour_kernel<<<grid, K2_THREADS, sizeof(Smem), stream>>>(bev.planes, bev.fg, bev.hdr, B, P, int_ids_dev, first_slot, presort_scratch, heads, views,
                                                             ells, klists, arenas, force_arena, work_counter, dbg);
  C2G_CUDA_TRY(cudaGetLastError());
  return 0;
}

// Dense BEV image of ONE scan of the last batch, on demand (c2g_get_bev: ContourManager::getBevImage + bev_pixfs_): every
// occupied cell gets its height and the winner's continuous coordinates, empty cells the reference's initial values.
// `offsets[b]` is read on the device: the offsets of the last batch live there.
namespace {
__global__ void bev_fill_entry(const c2g_cellkey *tile, const float *pts, const long long *offsets, int b, int fpp, C2gIngestParams P,
                               float *bev_h, float *bev_rf, float *bev_cf) {  // tile: ONE scan (the full-tile scatter variant's output)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_cells) return;
  const float *p = pts + (size_t) fpp * offsets[b];  // fpp: floats per point (4 = KITTI .bin layout, 3 = xyz)
  const c2g_cellkey k = tile[c];
  float h = -1000.0f, rf = -1.0f, cf = -1.0f;
  if (k != 0ull) {
    const float *q = p + (size_t) fpp * (0xFFFFFFFFu - (uint32_t) k);
    const float2 xy = make_float2(q[0], q[1]);
    h = c2g_from_orderable((uint32_t) (k >> 32));
    rf = (xy.x / P.cfg.reso_row + P.half_row_f) - 0.5f;
    cf = (xy.y / P.cfg.reso_col + P.half_col_f) - 0.5f;
  }
  bev_h[c] = h;
  bev_rf[c] = rf;
  bev_cf[c] = cf;
}
}  // namespace
int c2g_launch_bev_fill(const c2g_cellkey *tiles, const float *pts_dev, const long long *offsets_dev, int b, int fpp, const C2gIngestParams &P,
                        float *bev_h, float *bev_rf, float *bev_cf, cudaStream_t stream) {
  bev_fill_entry<<<(P.n_cells + 255) / 256, 256, 0, stream>>>(tiles, pts_dev, offsets_dev, b, fpp, P, bev_h, bev_rf, bev_cf);
  C2G_CUDA_TRY(cudaGetLastError());
  return 0;
}
