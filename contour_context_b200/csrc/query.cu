// query.cu — the database/query half of the hot path on the device.
//
//   knn_kernel     LayerDB::layerKNNSearch + TreeBucket::knnSearch        src/cont2/contour_db.cpp:319-403
//                  (+ dist_ub of ContourDB::queryRangedKNN, include/cont2/contour_db.h:733-749; nanoflann L2 metric order
//                  thirdparty/nanoflann.hpp:428-462) as a flat scan over the device mirror of the KD-tree contents.
//   score_kernel   CandidateManager::checkCandWithHint up to addProposal   include/cont2/contour_db.h:374-437
//                    ContourView::checkSim                                 include/cont2/contour.h:278-329
//                    BCI::checkConstellSim                                 include/cont2/contour_mng.h:288-388
//                    ContourManager::checkConstellCorrespSim               include/cont2/contour_mng.h:1124-1242
//                    ContourManager::getTFFromConstell                     include/cont2/contour_mng.h:1251-1277
//   finish_replay_kernel / finish_corr_kernel / finish_output_kernel
//                  CandidatePoseData::addProposal replay (:286-338), tidyUpCandidates (:494-596) with the GMM-L2 initial
//                  correlation (include/cont2/correlation.h:42-202) and fineOptimize's first ordering (:604-621); the
//                  refinement itself is refine.cu.
//
// With DYNAMIC_THRES=0 (CMakeLists.txt:21) every hint check is independent, so hints are scored in parallel (one warp
// each) and only the small order-dependent proposal merge is replayed sequentially per query scan, in reference order
// (q-level, query seq, ascending key distance).
#include "../../include/c2g.h"
#include "c2g_ctx.cuh"
#include "layer_db_host.h"
#include <algorithm>
#include <chrono>
#include <vector>

#include "stdsort.cuh"
#include "c2g_libm.cuh"

#define C2G_QPROF(ctx, k) do { if ((ctx)->prof_on) cudaEventRecord((ctx)->prof_ev[k], (ctx)->stream); } while (0)

namespace {

constexpr int QK_WARPS = 8;           // warps per CTA in the kNN kernel
constexpr int KNN_BOX_CACHE = 256;    // blocks per bucket whose box distance is kept in shared memory between the two passes
constexpr int SC_WARPS = 4;           // warps per CTA in the score kernel (7.2 KB of scratch per warp)
constexpr int MAX_POT_PAIRS = 400;    // 4 layers x 10 x 10 neighbour pairs
constexpr double C2G_PI = 3.14159265358979323846;

struct LayerDev {
  const float *keys_t;  // [KEY_DIM][cap]
  const int *gidx;
  const signed char *seq;
  const int *orank;     // bucket-major TREE order rank (what the entry's position would be without the kd ordering)
  const float *box_min, *box_max;  // [KEY_DIM][blk_cap] bounding boxes of the 32-key blocks
  int cap, blk_cap;     // row strides of keys_t / box_*: C2G_PHYS_BUCKETS * cap_b, C2G_PHYS_BUCKETS * blkcap_b
  int cap_b, blkcap_b;  // keys of bucket k occupy [p * cap_b, p * cap_b + bucket_cnt[k]), p = phys[k], in blocks of 32 from p * blkcap_b
  int phys[C2G_NUM_BUCKETS];
  int bucket_cnt[C2G_NUM_BUCKETS];
  float ranges[C2G_NUM_BUCKETS + 1];
};
// The part of a LayerDev that changes while a DB grows, per (query scan, layer) of a launch: lets ONE kNN launch serve query
// scans that must see different states of the trees (windowed online loop; see C2gLayerTable for why the older states are
// still intact when the launch runs).
struct KnnVersion {
  int phys[C2G_NUM_BUCKETS];
  int bucket_cnt[C2G_NUM_BUCKETS];
  float ranges[C2G_NUM_BUCKETS + 1];
  int pad_;
};
struct QueryParams {
  LayerDev layer[C2G_NUM_Q_LEVELS_MAX];
  int n_q_levels, q_levels[C2G_NUM_Q_LEVELS_MAX];
  int nnk, piv;
  c2g_sim_config sim;
  c2g_score_ensemble lb;
  int n_row, n_col;
};

__device__ __forceinline__ const c2g_view &view_at(const c2g_scan_head *heads, const c2g_view *views, int slot, int level, int seq) {
  return views[(size_t) slot * C2G_VIEW_CAP + heads[slot].view_off[level] + seq];
}

// ------------------------------------------------------------------------------------------------------------------
// kNN: one warp per query key. Lanes stride over the keys of the visited buckets; the running top-k is kept sorted
// across the warp's registers (slot j lives in lane j % 32, register j / 32) and updated by warp-cooperative insertion.
// ------------------------------------------------------------------------------------------------------------------
// MODE 0: every bucket lives in its primary region and all query scans see the mirror as the launch parameters describe it (a
//         batch against a static DB).  Its bucket addressing is kept in the textual form of rounds 1-2b on purpose: ptxas
//         schedules the kernel 9 % slower (1.10 -> 1.20 ms per 1 184 queries, identical instruction count) as soon as the bucket
//         size is held in a local or the region goes through T.phys;
// MODE 1: same, but buckets may live in their second region; MODE 2: every query scan has its own KnnVersion (online loop).
template <int MODE>
__global__ void __launch_bounds__(QK_WARPS * 32)
knn_kernel(const c2g_scan_head *__restrict__ heads, int first_slot, int q0, int B, QueryParams Q, const KnnVersion *__restrict__ ver,
           c2g_hint *__restrict__ hints, unsigned long long *__restrict__ work) {
  __shared__ float merge_d[QK_WARPS][64];
  __shared__ int merge_i[QK_WARPS][64], merge_o[QK_WARPS][64];
  __shared__ float box_cache[QK_WARPS][KNN_BOX_CACHE];  // box distances of a bucket's first blocks: computed once, used by both passes
  const int lane = threadIdx.x & 31;
  const int wglobal = blockIdx.x * QK_WARPS + (threadIdx.x >> 5);
  const int keys_per_scan = Q.n_q_levels * C2G_MAX_PIV;
  if (wglobal >= B * keys_per_scan) return;
  // layer-major ordering of the work so that the warps of one CTA scan the same table
  const int ll = wglobal / (B * C2G_MAX_PIV);
  const int rem = wglobal - ll * (B * C2G_MAX_PIV);
  const int q = q0 + rem / C2G_MAX_PIV, seq = rem - (rem / C2G_MAX_PIV) * C2G_MAX_PIV;  // queries [q0, q0 + B) of the batch
  const int level = Q.q_levels[ll];
  c2g_hint *out = hints + ((size_t) (q * Q.n_q_levels + ll) * C2G_MAX_PIV + seq) * Q.nnk;
  const float *qk = heads[first_slot + q].keys[level][seq];
  float key[C2G_KEY_DIM];
  float ksum = 0.0f;
#pragma unroll
  for (int d = 0; d < C2G_KEY_DIM; ++d) {
    key[d] = qk[d];
    ksum += key[d];
  }
  float bd[2] = {3.0e38f, 3.0e38f};  // top-k distances (sorted ascending across slots), empty = +big
  int bi[2] = {-1, -1};
  int bo[2] = {0x7FFFFFFF, 0x7FFFFFFF};  // original flat index (bucket-major tree order) of every kept entry: tie order
  int count = 0;
  int n_keys_eval = 0, n_boxes = 0;  // work counters (c2g_work_counters)
  const int K = Q.nnk;  // <= 64
  if (seq < Q.piv && ksum != 0.0f) {  // `q_keys[seq].sum() != 0` (contour_db.h:726); NaN keys search and find nothing
    // dist_ub (contour_db.h:733-749): bounds stored as float, products with double literals evaluated in double
    const float b00 = (float) ((double) key[0] * 0.8), b01 = (float) ((double) key[0] / 0.8);
    const float b10 = (float) ((double) key[1] * 0.8), b11 = (float) ((double) key[1] / 0.8);
    const float b20 = (float) (((double) key[2] * 0.8) * 0.75), b21 = (float) ((double) key[2] / (0.8 * 0.75));
    const float dist_ub = fmaxf((key[0] - b00) * (key[0] - b00), (key[0] - b01) * (key[0] - b01)) +
                          fmaxf((key[1] - b10) * (key[1] - b10), (key[1] - b11) * (key[1] - b11)) +
                          fmaxf((key[2] - b20) * (key[2] - b20), (key[2] - b21) * (key[2] - b21));
    const LayerDev &T = Q.layer[ll];
    // the state of the trees this query scan sees: the launch parameters, or the scan's own entry of the version table
    const KnnVersion *V = MODE == 2 ? ver + ((size_t) (q - q0) * Q.n_q_levels + ll) : nullptr;
    int mid = 0;
    if (MODE == 2) {
      for (int i = 0; i < C2G_NUM_BUCKETS; ++i)
        if (V->ranges[i] <= key[0] && V->ranges[i + 1] > key[0]) {
          mid = i;
          break;
        }
    } else {
      for (int i = 0; i < C2G_NUM_BUCKETS; ++i)
        if (T.ranges[i] <= key[0] && T.ranges[i + 1] > key[0]) {
          mid = i;
          break;
        }
    }
    // visited set of layerKNNSearch's else-if chain: {mid, mid-1, .., 0} and {mid+i : i > mid, mid+i < 6}
    unsigned visit = 0;
    for (int i = 0; i < C2G_NUM_BUCKETS; ++i) {
      if (i == 0)
        visit |= 1u << mid;
      else if (mid - i >= 0)
        visit |= 1u << (mid - i);
      else if (mid + i < C2G_NUM_BUCKETS)
        visit |= 1u << (mid + i);
    }
    // Admission key = (distance, original flat index) in lexicographic order; initial worst = (dist_ub, -1): strict
    // `dist < worst` (nanoflann.hpp:1575), NaN dist_ub admits nothing.  Inside a bucket the mirror is cut into kd-ordered
    // blocks of 32 keys with a 10-D bounding box each (c2g_db_set_layer).  The box distance is accumulated in the same
    // order as the metric, and every term is <= the matching term of any key in the box, so (rounding being monotone)
    // box_dist <= dist holds for the COMPUTED values too: a block is skipped iff box_dist > current K-th distance.
    float thr = dist_ub;
    int thr_o = -1;
    auto box_dist = [&](int blk) -> float {
      float g[C2G_KEY_DIM];
#pragma unroll
      for (int d = 0; d < C2G_KEY_DIM; ++d) {
        const float lo = T.box_min[(size_t) d * T.blk_cap + blk], hi = T.box_max[(size_t) d * T.blk_cap + blk];
        g[d] = fmaxf(fmaxf(lo - key[d], key[d] - hi), 0.0f);
      }
      float r = 0.0f;
      r += ((g[0] * g[0] + g[1] * g[1]) + g[2] * g[2]) + g[3] * g[3];
      r += ((g[4] * g[4] + g[5] * g[5]) + g[6] * g[6]) + g[7] * g[7];
      r += g[8] * g[8];
      r += g[9] * g[9];
      return r;
    };
    auto scan_block = [&](int base, int end) {
      const int i = base + lane;
      const bool in = i < end;
      n_keys_eval += end - base;
      float dist = 3.0e38f;
      int orig = 0x7FFFFFFF;
      if (in) {
        float df[C2G_KEY_DIM];
#pragma unroll
        for (int d = 0; d < C2G_KEY_DIM; ++d) df[d] = key[d] - T.keys_t[(size_t) d * T.cap + i];
        // nanoflann L2_Adaptor::evalMetric: groups of four, then the remainder one by one
        // (stopping after the first group when it alone exceeds the bound for the whole block was measured 15 % slower:
        // the kd blocks are not tight enough in those four dimensions for the vote to pass often)
        float r = 0.0f;
        r += ((df[0] * df[0] + df[1] * df[1]) + df[2] * df[2]) + df[3] * df[3];
        r += ((df[4] * df[4] + df[5] * df[5]) + df[6] * df[6]) + df[7] * df[7];
        r += df[8] * df[8];
        r += df[9] * df[9];
        dist = r;
        orig = T.orank[i];
      }
      const unsigned cand = __ballot_sync(0xFFFFFFFFu, in && (dist < thr || (dist == thr && orig < thr_o)));
      if (cand == 0u) return;
      // Batch merge of the block's admissible keys into the sorted top-64: every (distance, tree rank) pair is unique, so the
      // merged order is a permutation that each element can compute for itself -
      //   kept entry in slot j        -> j + #candidates that precede it
      //   candidate                   -> #kept entries that precede it + #candidates that precede it
      // - from one broadcast round per candidate (the rounds are independent of each other, unlike insertion one by one);
      // the permutation itself goes through 768 bytes of shared memory per warp.  Entries pushed past slot 63 fall off.
      const bool mine = (cand >> lane) & 1u;
      int sh0 = 0, sh1 = 0, before_new = 0, before_kept = 0;
      for (unsigned m = cand; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        const float nd = __shfl_sync(0xFFFFFFFFu, dist, src);
        const int no = __shfl_sync(0xFFFFFFFFu, orig, src);
        const bool k0 = bd[0] < nd || (bd[0] == nd && bo[0] < no);  // my kept entries precede this candidate
        const bool k1 = bd[1] < nd || (bd[1] == nd && bo[1] < no);
        const unsigned le0 = __ballot_sync(0xFFFFFFFFu, k0), le1 = __ballot_sync(0xFFFFFFFFu, k1);
        sh0 += k0 ? 0 : 1;
        sh1 += k1 ? 0 : 1;
        before_new += (mine && (nd < dist || (nd == dist && no < orig))) ? 1 : 0;
        if (lane == src) before_kept = __popc(le0) + __popc(le1);
      }
      float *sd = merge_d[threadIdx.x >> 5];
      int *si = merge_i[threadIdx.x >> 5], *so = merge_o[threadIdx.x >> 5];
      const int p0 = lane + sh0, p1 = lane + 32 + sh1, pn = before_kept + before_new;
      if (p0 < 64) {
        sd[p0] = bd[0];
        si[p0] = bi[0];
        so[p0] = bo[0];
      }
      if (p1 < 64) {
        sd[p1] = bd[1];
        si[p1] = bi[1];
        so[p1] = bo[1];
      }
      if (mine && pn < 64) {
        sd[pn] = dist;
        si[pn] = i;
        so[pn] = orig;
      }
      __syncwarp();
      bd[0] = sd[lane];
      bi[0] = si[lane];
      bo[0] = so[lane];
      bd[1] = sd[lane + 32];
      bi[1] = si[lane + 32];
      bo[1] = so[lane + 32];
      __syncwarp();
      count = min(K, count + __popc(cand));
      if (count == K) {  // worst kept entry = K-th best
        const int ks = K - 1;
        thr = __shfl_sync(0xFFFFFFFFu, ks < 32 ? bd[0] : bd[1], ks & 31);
        thr_o = __shfl_sync(0xFFFFFFFFu, ks < 32 ? bo[0] : bo[1], ks & 31);
      }
    };
    for (int bk = 0; bk < C2G_NUM_BUCKETS; ++bk) {
      if (!((visit >> bk) & 1u)) continue;
      int beg, end, b0, nb;
      if (MODE == 0) {
        beg = bk * T.cap_b, end = beg + T.bucket_cnt[bk];
        if (beg >= end) continue;
        b0 = bk * T.blkcap_b, nb = (T.bucket_cnt[bk] + 31) >> 5;
      } else {
        const int pb = MODE == 2 ? V->phys[bk] : T.phys[bk], cntb = MODE == 2 ? V->bucket_cnt[bk] : T.bucket_cnt[bk];
        beg = pb * T.cap_b, end = beg + cntb;
        if (beg >= end) continue;
        b0 = pb * T.blkcap_b, nb = (cntb + 31) >> 5;
      }
      n_boxes += nb + (nb > KNN_BOX_CACHE ? nb - KNN_BOX_CACHE : 0);
      // pass 1: the block nearest to the query seeds the top-k, so that the sweep below starts with a tight bound
      float best = 3.0e38f;
      int best_j = 0x7FFFFFFF;
      float *bc = box_cache[threadIdx.x >> 5];
      __syncwarp();
      for (int j = lane; j < nb; j += 32) {
        const float bdist = box_dist(b0 + j);
        if (j < KNN_BOX_CACHE) bc[j] = bdist;
        if (bdist < best) {
          best = bdist;
          best_j = j;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int oj = __shfl_xor_sync(0xFFFFFFFFu, best_j, o);
        if (ob < best || (ob == best && oj < best_j)) {
          best = ob;
          best_j = oj;
        }
      }
      const int seed = best_j;  // 0x7FFFFFFF when every box distance is NaN (NaN query key): nothing is admitted anyway
      if (seed < nb && best <= thr) scan_block(beg + seed * 32, min(end, beg + seed * 32 + 32));
      // pass 2: every other block whose box can still hold an admissible key
      for (int j0 = 0; j0 < nb; j0 += 32) {
        const int j = j0 + lane;
        const float bdist = j < nb ? (j < KNN_BOX_CACHE ? bc[j] : box_dist(b0 + j)) : 3.0e38f;  // a lane reads what it wrote itself
        unsigned todo = __ballot_sync(0xFFFFFFFFu, j < nb && j != seed && bdist <= thr);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const float bsrc = __shfl_sync(0xFFFFFFFFu, bdist, src);
          if (!(bsrc <= thr)) continue;  // the bound tightened while earlier blocks of this group were scanned
          const int base = beg + (j0 + src) * 32;
          scan_block(base, min(end, base + 32));
        }
      }
    }
    // slots >= K may hold spill-over entries; they are never emitted
    if (work && lane == 0) {
      atomicAdd(work + 0, (unsigned long long) n_keys_eval);
      atomicAdd(work + 1, (unsigned long long) n_boxes);
    }
  }
  const int level8 = level;
  for (int j = lane; j < Q.nnk; j += 32) {
    c2g_hint h;
    h.q_idx = q;
    h.cand_gidx = -1;
    h.level = (int8_t) level8;
    h.cand_seq = 0;
    h.q_seq = (int8_t) seq;
    h.q_level_idx = (int8_t) ll;
    h.dist_sq = 0.0f;
    const float dj = (j < 32) ? bd[0] : bd[1];
    const int ij = (j < 32) ? bi[0] : bi[1];
    // (j & 31) == lane by construction of the loop
    if (j < count && ij >= 0) {
      h.cand_gidx = Q.layer[ll].gidx[ij];
      h.cand_seq = Q.layer[ll].seq[ij];
      h.dist_sq = dj;
    }
    out[j] = h;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Hint scoring
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool diff_perc_f(float a, float b, float perc) { return fabsf((a - b) / fmaxf(a, b)) > perc; }
__device__ __forceinline__ bool diff_delt_f(float a, float b, float delta) { return fabsf(a - b) > delta; }

__device__ bool check_sim(const c2g_view &s, const c2g_view &t, const c2g_sim_config &th) {
  const float sc = (float) s.cell_cnt, tc = (float) t.cell_cnt;
  if (diff_perc_f(sc, tc, th.tp_cell_cnt) && diff_delt_f(sc, tc, th.ta_cell_cnt)) return false;
  if ((double) fmaxf(s.eig_vals[1], t.eig_vals[1]) > 2.0 && diff_perc_f(sqrtf(s.eig_vals[1]), sqrtf(t.eig_vals[1]), th.tp_eigval)) return false;
  if ((double) fmaxf(s.eig_vals[0], t.eig_vals[0]) > 2.0 && diff_perc_f(sqrtf(s.eig_vals[0]), sqrtf(t.eig_vals[0]), th.tp_eigval)) return false;
  if (max((int) s.cell_cnt, (int) t.cell_cnt) > 15 && diff_delt_f(s.vol3_mean, t.vol3_mean, th.ta_h_bar)) return false;
  const float sx = s.com[0] - s.pos_mean[0], sy = s.com[1] - s.pos_mean[1];
  const float tx = t.com[0] - t.pos_mean[0], ty = t.com[1] - t.pos_mean[1];
  const float r1 = sqrtf(sx * sx + sy * sy), r2 = sqrtf(tx * tx + ty * ty);
  if (diff_delt_f(r1, r2, th.ta_rcom) && diff_perc_f(r1, r2, th.tp_rcom)) return false;
  return true;
}

struct PotPair {
  float orie_diff;
  int8_t seq_src, seq_tgt, level, pad;
};
struct CPairD {
  int8_t level, seq_src, seq_tgt;
};

// What the cascade reads of a ContourView (contour.h:106-119), staged in shared memory per warp
struct ViewLite {
  float mean0, mean1, eig0, eig1, vol3_mean, com0, com1, evx, evy;  // evx/evy = eig_vecs.col(1)
  int16_t cell_cnt;
  uint8_t ecc_feat, valid;
};
__device__ __forceinline__ bool check_sim_lite(const ViewLite &s, const ViewLite &t, const c2g_sim_config &th) {
  const float sc = (float) s.cell_cnt, tc = (float) t.cell_cnt;
  if (diff_perc_f(sc, tc, th.tp_cell_cnt) && diff_delt_f(sc, tc, th.ta_cell_cnt)) return false;
  if ((double) fmaxf(s.eig1, t.eig1) > 2.0 && diff_perc_f(sqrtf(s.eig1), sqrtf(t.eig1), th.tp_eigval)) return false;
  if ((double) fmaxf(s.eig0, t.eig0) > 2.0 && diff_perc_f(sqrtf(s.eig0), sqrtf(t.eig0), th.tp_eigval)) return false;
  if (max((int) s.cell_cnt, (int) t.cell_cnt) > 15 && diff_delt_f(s.vol3_mean, t.vol3_mean, th.ta_h_bar)) return false;
  const float sx = s.com0 - s.mean0, sy = s.com1 - s.mean1;
  const float tx = t.com0 - t.mean0, ty = t.com1 - t.mean1;
  const float r1 = sqrtf(sx * sx + sy * sy), r2 = sqrtf(tx * tx + ty * ty);
  if (diff_delt_f(r1, r2, th.ta_rcom) && diff_perc_f(r1, r2, th.tp_rcom)) return false;
  return true;
}

constexpr int SV_PER_SCAN = C2G_NUM_BIN_LAYERS * C2G_MAX_DIST_FIRSTS;  // levels 1..4 x top-10 contours
struct ScoreScratch {
  PotPair pot[MAX_POT_PAIRS];
  uint32_t ord[MAX_POT_PAIRS];  // sort permutation (indices into pot)
  CPairD c1[MAX_POT_PAIRS + 1];
  CPairD c2[MAX_POT_PAIRS + 1];
  uint8_t drop[MAX_POT_PAIRS + 1];  // orientation filter verdict per entry of c2
  c2g_bci bs, bt;                   // BCI of the candidate (src) and of the query (tgt) anchor
  ViewLite vs[SV_PER_SCAN], vt[SV_PER_SCAN];
};

__device__ __forceinline__ float clamp_ang_f(float ang) {  // clampAng<float>: double arithmetic, stored to float
  return (float) ((double) ang - floor(((double) ang + C2G_PI) / (2 * C2G_PI)) * 2 * C2G_PI);
}
__device__ __forceinline__ void normalized2(float x, float y, float &ox, float &oy) {  // Eigen normalized()
  const float n2 = x * x + y * y;
  if (n2 > 0.0f) {
    const float n = sqrtf(n2);
    ox = x / n;
    oy = y / n;
  } else {
    ox = x;
    oy = y;
  }
}

__device__ __forceinline__ void stage_views(const c2g_scan_head *heads, const c2g_view *views, int slot, ViewLite *dst, int lane) {
  for (int i = lane; i < SV_PER_SCAN; i += 32) {
    const int level = i / C2G_MAX_DIST_FIRSTS + 1, seq = i % C2G_MAX_DIST_FIRSTS;
    ViewLite v;
    v.valid = 0;
    v.mean0 = v.mean1 = v.eig0 = v.eig1 = v.vol3_mean = v.com0 = v.com1 = v.evx = v.evy = 0.f;
    v.cell_cnt = 0;
    v.ecc_feat = 0;
    if (seq < heads[slot].n_views[level]) {
      const c2g_view &g = views[(size_t) slot * C2G_VIEW_CAP + heads[slot].view_off[level] + seq];
      v.valid = 1;
      v.mean0 = g.pos_mean[0];
      v.mean1 = g.pos_mean[1];
      v.eig0 = g.eig_vals[0];
      v.eig1 = g.eig_vals[1];
      v.vol3_mean = g.vol3_mean;
      v.com0 = g.com[0];
      v.com1 = g.com[1];
      v.evx = g.eig_vecs[2];
      v.evy = g.eig_vecs[3];
      v.cell_cnt = g.cell_cnt;
      v.ecc_feat = g.ecc_feat;
    }
    dst[i] = v;
  }
}
__device__ __forceinline__ const ViewLite &lite(const ViewLite *arr, int level, int seq) { return arr[(level - 1) * C2G_MAX_DIST_FIRSTS + seq]; }

// One warp per surviving hint. All lanes stage the two BCIs and the top-10 views of levels 1..4 of both scans into shared
// memory (coalesced / parallel global reads); the order-dependent parts of the cascade then run on lane 0 out of shared
// memory only, the per-pair similarity and orientation tests run one pair per lane.
__device__ void score_hint(const c2g_scan_head *heads, const c2g_view *views, int q_slot, const c2g_hint &hint, const QueryParams &Q,
                           ScoreScratch &sc, c2g_pair_score &rec, int lane) {
  const int cand = hint.cand_gidx, level = hint.level, cseq = hint.cand_seq, qseq = hint.q_seq;
  {
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(&heads[cand].bcis[level][cseq]);
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(&heads[q_slot].bcis[level][qseq]);
    uint32_t *ds = reinterpret_cast<uint32_t *>(&sc.bs), *dt = reinterpret_cast<uint32_t *>(&sc.bt);
    for (int i = lane; i < (int) (sizeof(c2g_bci) / 4); i += 32) {
      ds[i] = s32[i];
      dt[i] = t32[i];
    }
    stage_views(heads, views, cand, sc.vs, lane);
    stage_views(heads, views, q_slot, sc.vt, lane);
  }
  __syncwarp();
  const c2g_bci &src = sc.bs;
  const c2g_bci &tgt = sc.bt;
  int state = 1;  // 1 = continue, <= 0 = final `passed` code
  int npot = 0, n1 = 0;
  if (lane == 0) {
    // (1/4) anchor similarity and (2/4) popcount gate were evaluated by the prefilter; recompute the counts for the record
    int ov1 = 0, ov2 = 0, ov3 = 0;
    for (int i = 0; i < 4; ++i) {
      const unsigned long long s_i = src.dist_bin[i], t_i = tgt.dist_bin[i];
      const unsigned long long sl = (s_i << 1) | (i > 0 ? (src.dist_bin[i - 1] >> 63) : 0ull);
      const unsigned long long sr = (s_i >> 1) | (i < 3 ? (src.dist_bin[i + 1] << 63) : 0ull);
      ov1 += __popcll(s_i & t_i);
      ov2 += __popcll(sl & t_i);
      ov3 += __popcll(sr & t_i);
    }
    rec.constell[0] = ov1 + ov2 + ov3;
    rec.constell[1] = max(ov1, max(ov2, ov3));
    rec.constell[2] = 0;
    if (!check_sim_lite(lite(sc.vs, level, cseq), lite(sc.vt, level, qseq), Q.sim)) {
      state = 0;  // cannot happen for prefilter survivors; keep the record of an anchor failure all-zero
      rec.constell[0] = rec.constell[1] = 0;
    }
    else if (!(rec.constell[0] >= Q.lb.i_ovlp_sum && rec.constell[1] >= Q.lb.i_ovlp_max_one))
      state = -1;
    if (state == 1) {
      const int n_sseg = src.n_seg, n_tseg = tgt.n_seg;
      int p11 = 0, p12;
      for (int p2 = 0; p2 < n_tseg - 1; p2++) {
        const int tb = tgt.nei[tgt.seg[p2]].bit_pos;
        while (p11 < n_sseg - 1 && src.nei[src.seg[p11]].bit_pos < tb - 1) p11++;
        p12 = p11;
        while (p12 < n_sseg - 1 && src.nei[src.seg[p12]].bit_pos <= tb + 1) p12++;
        for (int i = tgt.seg[p2]; i < tgt.seg[p2 + 1]; i++)
          for (int j = src.seg[p11]; j < src.seg[p12]; j++) {
            if (npot < MAX_POT_PAIRS) {
              PotPair pp;
              pp.orie_diff = clamp_ang_f(tgt.nei[i].theta - src.nei[j].theta);
              pp.seq_src = src.nei[j].seq;
              pp.seq_tgt = tgt.nei[i].seq;
              pp.level = src.nei[j].level;
              pp.pad = 0;
              sc.pot[npot] = pp;
              sc.ord[npot] = (uint32_t) npot;
              ++npot;
            }
          }
      }
      // std::sort by orie_diff (contour_mng.h:340-342): sort the index permutation with the same comparator
      const PotPair *pot = sc.pot;
      c2g_sort::std_sort(sc.ord, (long) npot, [pot](uint32_t a, uint32_t b) { return pot[a].orie_diff < pot[b].orie_diff; });
      int longest = 1, longest_beg = 0;
      const float angular_range = (float) (C2G_PI / 16);
      int p1 = 0, p2 = 0;
      const int pot_sz = npot;
      while (p1 < pot_sz) {
        const float dd = sc.pot[sc.ord[p2 % pot_sz]].orie_diff - sc.pot[sc.ord[p1]].orie_diff;
        if ((double) dd + 2 * C2G_PI * (double) (p2 / pot_sz) > (double) angular_range)
          p1++;
        else {
          if (p2 - p1 + 1 > longest) {
            longest = p2 - p1 + 1;
            longest_beg = p1;
          }
          p2++;
        }
      }
      rec.constell[2] = longest;
      if (longest < Q.lb.i_in_ang_rng)
        state = -1;
      else {
        for (int i = longest_beg; i < longest + longest_beg; i++) {
          const PotPair &pp = sc.pot[sc.ord[i % npot]];
          sc.c1[n1].level = pp.level;
          sc.c1[n1].seq_src = pp.seq_src;
          sc.c1[n1].seq_tgt = pp.seq_tgt;
          ++n1;
        }
        sc.c1[n1].level = src.level;
        sc.c1[n1].seq_src = src.piv_seq;
        sc.c1[n1].seq_tgt = tgt.piv_seq;
        ++n1;
      }
    }
  }
  state = __shfl_sync(0xFFFFFFFFu, state, 0);
  if (state != 1) {
    if (lane == 0) rec.passed = state;
    return;
  }
  n1 = __shfl_sync(0xFFFFFFFFu, n1, 0);
  __syncwarp();
  // (3/4) checkConstellCorrespSim step 1: individual similarity, one pair per lane, order-preserving compaction
  int n2 = 0;
  for (int base = 0; base < n1; base += 32) {
    const int i = base + lane;
    bool ok = false;
    CPairD pr;
    pr.level = pr.seq_src = pr.seq_tgt = 0;
    if (i < n1) {
      pr = sc.c1[i];
      ok = check_sim_lite(lite(sc.vs, pr.level, pr.seq_src), lite(sc.vt, pr.level, pr.seq_tgt), Q.sim);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
    if (ok) sc.c2[n2 + __popc(m & ((1u << lane) - 1u))] = pr;
    n2 += __popc(m);
  }
  __syncwarp();
  if (lane == 0) {
    rec.pairwise[0] = n2;
    rec.pairwise[1] = 0;
  }
  if (n2 < Q.lb.i_indiv_sim) {
    if (lane == 0) rec.passed = -2;
    return;
  }
  // step 2.1: the "shaft" (the last qualifying (i, j) among the first <= 10 pairs wins, see SURVEY.md §8a' #6)
  float ssx = 0.f, ssy = 0.f, stx = 0.f, sty = 0.f;
  if (lane == 0) {
    for (int i = 1; i < min(n2, 10); i++)
      for (int j = 0; j < i; j++) {
        const ViewLite &mi = lite(sc.vs, sc.c2[i].level, sc.c2[i].seq_src), &mj = lite(sc.vs, sc.c2[j].level, sc.c2[j].seq_src);
        const float cx = mi.mean0 - mj.mean0, cy = mi.mean1 - mj.mean1;
        if (sqrtf(cx * cx + cy * cy) > sqrtf(ssx * ssx + ssy * ssy)) {
          normalized2(cx, cy, ssx, ssy);
          const ViewLite &ti = lite(sc.vt, sc.c2[i].level, sc.c2[i].seq_tgt), &tj = lite(sc.vt, sc.c2[j].level, sc.c2[j].seq_tgt);
          normalized2(ti.mean0 - tj.mean0, ti.mean1 - tj.mean1, stx, sty);
        }
      }
  }
  ssx = __shfl_sync(0xFFFFFFFFu, ssx, 0);
  ssy = __shfl_sync(0xFFFFFFFFu, ssy, 0);
  stx = __shfl_sync(0xFFFFFFFFu, stx, 0);
  sty = __shfl_sync(0xFFFFFFFFu, sty, 0);
  // step 2.2: orientation verdict per pair (depends on the pair and the fixed shaft only), one pair per lane
  for (int i = lane; i < n2; i += 32) {
    const ViewLite &sc1 = lite(sc.vs, sc.c2[i].level, sc.c2[i].seq_src), &tc1 = lite(sc.vt, sc.c2[i].level, sc.c2[i].seq_tgt);
    uint8_t d = 0;
    if (sc1.ecc_feat && tc1.ecc_feat) {
      const float theta_s = c2g_acosf(ssx * sc1.evx + ssy * sc1.evy);
      const float theta_t = c2g_acosf(stx * tc1.evx + sty * tc1.evy);
      const float pi6 = (float) (C2G_PI / 6);
      d = (diff_delt_f(theta_s, theta_t, pi6) && diff_delt_f((float) (C2G_PI - (double) theta_s), theta_t, pi6)) ? 1 : 0;
    }
    sc.drop[i] = d;
  }
  __syncwarp();
  if (lane == 0) {
    // the reference's swap-with-last removal loop (contour_mng.h:1187-1201), verdicts travel with their entries
    int num_sim = n2;
    for (int i = 0; i < num_sim;) {
      if (sc.drop[i]) {
        const CPairD tmp = sc.c2[i];
        sc.c2[i] = sc.c2[num_sim - 1];
        sc.c2[num_sim - 1] = tmp;
        const uint8_t td = sc.drop[i];
        sc.drop[i] = sc.drop[num_sim - 1];
        sc.drop[num_sim - 1] = td;
        num_sim--;
        continue;
      }
      i++;
    }
    n2 = num_sim;
    rec.pairwise[1] = n2;
    if (n2 < Q.lb.i_orie_sim) {
      rec.passed = -2;
    } else {
      // getTFFromConstell: 2-D Umeyama without scaling, closed form (double), sums in list order
      const double inv_n = 1.0 / (double) n2;
      double sm0 = 0, sm1 = 0, dm0 = 0, dm1 = 0;
      for (int i = 0; i < n2; ++i) {
        const ViewLite &ps = lite(sc.vs, sc.c2[i].level, sc.c2[i].seq_src), &pt = lite(sc.vt, sc.c2[i].level, sc.c2[i].seq_tgt);
        sm0 += (double) ps.mean0;
        sm1 += (double) ps.mean1;
        dm0 += (double) pt.mean0;
        dm1 += (double) pt.mean1;
      }
      sm0 *= inv_n;
      sm1 *= inv_n;
      dm0 *= inv_n;
      dm1 *= inv_n;
      double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
      for (int i = 0; i < n2; ++i) {
        const ViewLite &ps = lite(sc.vs, sc.c2[i].level, sc.c2[i].seq_src), &pt = lite(sc.vt, sc.c2[i].level, sc.c2[i].seq_tgt);
        const double sx = (double) ps.mean0 - sm0, sy = (double) ps.mean1 - sm1;
        const double dx = (double) pt.mean0 - dm0, dy = (double) pt.mean1 - dm1;
        s00 += dx * sx;
        s01 += dx * sy;
        s10 += dy * sx;
        s11 += dy * sy;
      }
      s00 *= inv_n;
      s01 *= inv_n;
      s10 *= inv_n;
      s11 *= inv_n;
      const double ang0 = atan2(s10 - s01, s00 + s11);
      const double c0 = cos(ang0), s0 = sin(ang0);
      const double tx = dm0 - (c0 * sm0 - s0 * sm1);
      const double ty = dm1 - (s0 * sm0 + c0 * sm1);
      const double ang = atan2(s0, c0);
      rec.T[0] = cos(ang);
      rec.T[1] = sin(ang);
      rec.T[2] = tx;
      rec.T[3] = ty;
      rec.passed = 1;
      rec.n_pairs = n2;
      for (int i = 0; i < n2; ++i) {
        const int bit = (sc.c2[i].level - 1) * 100 + sc.c2[i].seq_src * 10 + sc.c2[i].seq_tgt;
        rec.pair_bits[bit >> 6] |= 1ull << (bit & 63);
      }
    }
  }
}

// The same cascade, one THREAD per surviving hint (the warp-per-hint version above spends most of its instructions on
// lane 0: order-dependent list building, the std::sort replay, the two-pointer window, the shaft walk, the removal loop and
// the Umeyama sums are all sequential in the reference).  Per-thread lists live in local memory (4.8 KB per thread);
// descriptors are read straight from the head / view arenas (L1/L2 resident: the hints of one query share their target).
__device__ __forceinline__ const c2g_view &view_of(const c2g_scan_head *heads, const c2g_view *views, int slot, int level, int seq) {
  return views[(size_t) slot * C2G_VIEW_CAP + heads[slot].view_off[level] + seq];
}

__device__ void score_hint_serial(const c2g_scan_head *heads, const c2g_view *views, int q_slot, const c2g_hint &hint, const QueryParams &Q,
                                  c2g_pair_score &rec) {
  const int cand = hint.cand_gidx, level = hint.level, cseq = hint.cand_seq, qseq = hint.q_seq;
  const c2g_bci &src = heads[cand].bcis[level][cseq];
  const c2g_bci &tgt = heads[q_slot].bcis[level][qseq];
  {
    // both 608-byte records are walked below through dependent indices (seg -> nei -> bit_pos): pull their lines into L1 now,
    // with independent requests, instead of paying one L2 round trip per step of those chains
    const char *ps = reinterpret_cast<const char *>(&src), *pt = reinterpret_cast<const char *>(&tgt);
#pragma unroll
    for (int o = 0; o < (int) sizeof(c2g_bci) + 127; o += 128) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(ps + (o < (int) sizeof(c2g_bci) ? o : (int) sizeof(c2g_bci) - 1)));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pt + (o < (int) sizeof(c2g_bci) ? o : (int) sizeof(c2g_bci) - 1)));
    }
  }
  unsigned long long pot[MAX_POT_PAIRS];  // (orie_diff bits << 32) | level << 16 | seq_src << 8 | seq_tgt
  CPairD c2[MAX_POT_PAIRS + 1];
  uint8_t drop[MAX_POT_PAIRS + 1];
  {
    int ov1 = 0, ov2 = 0, ov3 = 0;
    for (int i = 0; i < 4; ++i) {
      const unsigned long long s_i = src.dist_bin[i], t_i = tgt.dist_bin[i];
      const unsigned long long sl = (s_i << 1) | (i > 0 ? (src.dist_bin[i - 1] >> 63) : 0ull);
      const unsigned long long sr = (s_i >> 1) | (i < 3 ? (src.dist_bin[i + 1] << 63) : 0ull);
      ov1 += __popcll(s_i & t_i);
      ov2 += __popcll(sl & t_i);
      ov3 += __popcll(sr & t_i);
    }
    rec.constell[0] = ov1 + ov2 + ov3;
    rec.constell[1] = max(ov1, max(ov2, ov3));
    rec.constell[2] = 0;
  }
  if (!check_sim(view_of(heads, views, cand, level, cseq), view_of(heads, views, q_slot, level, qseq), Q.sim)) {
    rec.constell[0] = rec.constell[1] = 0;  // cannot happen for prefilter survivors; an anchor failure keeps an all-zero record
    rec.passed = 0;
    return;
  }
  if (!(rec.constell[0] >= Q.lb.i_ovlp_sum && rec.constell[1] >= Q.lb.i_ovlp_max_one)) {
    rec.passed = -1;
    return;
  }
  // (2/4) BCI::checkConstellSim (contour_mng.h:309-388): potential pairs, sort by orientation difference, circular window
  int npot = 0;
  {
    const int n_sseg = src.n_seg, n_tseg = tgt.n_seg;
    int p11 = 0, p12;
    for (int p2 = 0; p2 < n_tseg - 1; p2++) {
      const int tb = tgt.nei[tgt.seg[p2]].bit_pos;
      while (p11 < n_sseg - 1 && src.nei[src.seg[p11]].bit_pos < tb - 1) p11++;
      p12 = p11;
      while (p12 < n_sseg - 1 && src.nei[src.seg[p12]].bit_pos <= tb + 1) p12++;
      const int j0 = src.seg[p11], j1 = src.seg[p12];
      for (int i = tgt.seg[p2]; i < tgt.seg[p2 + 1]; i++) {
        const c2g_relpt ti = tgt.nei[i];
        for (int j = j0; j < j1; j++) {
          if (npot < MAX_POT_PAIRS) {
            const c2g_relpt sj = src.nei[j];
            const float od = clamp_ang_f(ti.theta - sj.theta);
            pot[npot++] = ((unsigned long long) __float_as_uint(od) << 32) | ((unsigned long long) (uint8_t) sj.level << 16) |
                          ((unsigned long long) (uint8_t) sj.seq << 8) | (unsigned long long) (uint8_t) ti.seq;
          }
        }
      }
    }
  }
  // std::sort by orie_diff (contour_mng.h:340-342): the records themselves are moved, the comparator reads the key only
  c2g_sort::std_sort(pot, (long) npot, [](unsigned long long a, unsigned long long b) {
    return __uint_as_float((unsigned) (a >> 32)) < __uint_as_float((unsigned) (b >> 32));
  });
  int longest = 1, longest_beg = 0;
  {
    const float angular_range = (float) (C2G_PI / 16);
    int p1 = 0, p2 = 0;
    const int pot_sz = npot;
    while (p1 < pot_sz) {
      const float dd = __uint_as_float((unsigned) (pot[p2 % pot_sz] >> 32)) - __uint_as_float((unsigned) (pot[p1] >> 32));
      if ((double) dd + 2 * C2G_PI * (double) (p2 / pot_sz) > (double) angular_range)
        p1++;
      else {
        if (p2 - p1 + 1 > longest) {
          longest = p2 - p1 + 1;
          longest_beg = p1;
        }
        p2++;
      }
    }
  }
  rec.constell[2] = longest;
  if (longest < Q.lb.i_in_ang_rng) {
    rec.passed = -1;
    return;
  }
  // (3/4) checkConstellCorrespSim step 1 (contour_mng.h:1135-1152): individual similarity of the window's pairs + the anchor
  int n2 = 0;
  for (int i = longest_beg; i <= longest + longest_beg; i++) {
    CPairD pr;
    if (i < longest + longest_beg) {
      const unsigned long long w = pot[i % npot];
      pr.level = (int8_t) ((w >> 16) & 0xFF);
      pr.seq_src = (int8_t) ((w >> 8) & 0xFF);
      pr.seq_tgt = (int8_t) (w & 0xFF);
    } else {
      pr.level = src.level;
      pr.seq_src = src.piv_seq;
      pr.seq_tgt = tgt.piv_seq;
    }
    if (check_sim(view_of(heads, views, cand, pr.level, pr.seq_src), view_of(heads, views, q_slot, pr.level, pr.seq_tgt), Q.sim)) c2[n2++] = pr;
  }
  rec.pairwise[0] = n2;
  rec.pairwise[1] = 0;
  if (n2 < Q.lb.i_indiv_sim) {
    rec.passed = -2;
    return;
  }
  // step 2.1: the "shaft" (the last qualifying (i, j) among the first <= 10 pairs wins, see SURVEY.md §8a' #6)
  float ssx = 0.f, ssy = 0.f, stx = 0.f, sty = 0.f;
  for (int i = 1; i < min(n2, 10); i++)
    for (int j = 0; j < i; j++) {
      const c2g_view &mi = view_of(heads, views, cand, c2[i].level, c2[i].seq_src), &mj = view_of(heads, views, cand, c2[j].level, c2[j].seq_src);
      const float cx = mi.pos_mean[0] - mj.pos_mean[0], cy = mi.pos_mean[1] - mj.pos_mean[1];
      if (sqrtf(cx * cx + cy * cy) > sqrtf(ssx * ssx + ssy * ssy)) {
        normalized2(cx, cy, ssx, ssy);
        const c2g_view &ti = view_of(heads, views, q_slot, c2[i].level, c2[i].seq_tgt), &tj = view_of(heads, views, q_slot, c2[j].level, c2[j].seq_tgt);
        normalized2(ti.pos_mean[0] - tj.pos_mean[0], ti.pos_mean[1] - tj.pos_mean[1], stx, sty);
      }
    }
  // step 2.2: orientation verdict per pair (depends on the pair and the fixed shaft only)
  for (int i = 0; i < n2; ++i) {
    const c2g_view &sc1 = view_of(heads, views, cand, c2[i].level, c2[i].seq_src), &tc1 = view_of(heads, views, q_slot, c2[i].level, c2[i].seq_tgt);
    uint8_t d = 0;
    if (sc1.ecc_feat && tc1.ecc_feat) {
      const float theta_s = c2g_acosf(ssx * sc1.eig_vecs[2] + ssy * sc1.eig_vecs[3]);
      const float theta_t = c2g_acosf(stx * tc1.eig_vecs[2] + sty * tc1.eig_vecs[3]);
      const float pi6 = (float) (C2G_PI / 6);
      d = (diff_delt_f(theta_s, theta_t, pi6) && diff_delt_f((float) (C2G_PI - (double) theta_s), theta_t, pi6)) ? 1 : 0;
    }
    drop[i] = d;
  }
  {
    // the reference's swap-with-last removal loop (contour_mng.h:1187-1201), verdicts travel with their entries
    int num_sim = n2;
    for (int i = 0; i < num_sim;) {
      if (drop[i]) {
        const CPairD tmp = c2[i];
        c2[i] = c2[num_sim - 1];
        c2[num_sim - 1] = tmp;
        const uint8_t td = drop[i];
        drop[i] = drop[num_sim - 1];
        drop[num_sim - 1] = td;
        num_sim--;
        continue;
      }
      i++;
    }
    n2 = num_sim;
  }
  rec.pairwise[1] = n2;
  if (n2 < Q.lb.i_orie_sim) {
    rec.passed = -2;
    return;
  }
  // getTFFromConstell: 2-D Umeyama without scaling, closed form (double), sums in list order
  const double inv_n = 1.0 / (double) n2;
  double sm0 = 0, sm1 = 0, dm0 = 0, dm1 = 0;
  for (int i = 0; i < n2; ++i) {
    const c2g_view &ps = view_of(heads, views, cand, c2[i].level, c2[i].seq_src), &pt = view_of(heads, views, q_slot, c2[i].level, c2[i].seq_tgt);
    sm0 += (double) ps.pos_mean[0];
    sm1 += (double) ps.pos_mean[1];
    dm0 += (double) pt.pos_mean[0];
    dm1 += (double) pt.pos_mean[1];
  }
  sm0 *= inv_n;
  sm1 *= inv_n;
  dm0 *= inv_n;
  dm1 *= inv_n;
  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
  for (int i = 0; i < n2; ++i) {
    const c2g_view &ps = view_of(heads, views, cand, c2[i].level, c2[i].seq_src), &pt = view_of(heads, views, q_slot, c2[i].level, c2[i].seq_tgt);
    const double sx = (double) ps.pos_mean[0] - sm0, sy = (double) ps.pos_mean[1] - sm1;
    const double dx = (double) pt.pos_mean[0] - dm0, dy = (double) pt.pos_mean[1] - dm1;
    s00 += dx * sx;
    s01 += dx * sy;
    s10 += dy * sx;
    s11 += dy * sy;
  }
  s00 *= inv_n;
  s01 *= inv_n;
  s10 *= inv_n;
  s11 *= inv_n;
  const double ang0 = atan2(s10 - s01, s00 + s11);
  const double c0 = cos(ang0), s0 = sin(ang0);
  const double tx = dm0 - (c0 * sm0 - s0 * sm1);
  const double ty = dm1 - (s0 * sm0 + c0 * sm1);
  const double ang = atan2(s0, c0);
  rec.T[0] = cos(ang);
  rec.T[1] = sin(ang);
  rec.T[2] = tx;
  rec.T[3] = ty;
  rec.passed = 1;
  rec.n_pairs = n2;
  for (int i = 0; i < n2; ++i) {
    const int bit = (c2[i].level - 1) * 100 + c2[i].seq_src * 10 + c2[i].seq_tgt;
    rec.pair_bits[bit >> 6] |= 1ull << (bit & 63);
  }
}

// Stage 2 (thread version): one THREAD per surviving hint.
__global__ void __launch_bounds__(128)
score_thread_kernel(const c2g_scan_head *__restrict__ heads, const c2g_view *__restrict__ views, int first_slot, QueryParams Q,
                    const c2g_hint *__restrict__ hints, c2g_pair_score *__restrict__ scores, const int *__restrict__ survivors,
                    const int *__restrict__ n_surv) {
  const int n = *n_surv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int hid = survivors[i];
    const c2g_hint h = hints[hid];
    c2g_pair_score rec;
    rec.constell[0] = rec.constell[1] = rec.constell[2] = 0;
    rec.pairwise[0] = rec.pairwise[1] = 0;
    rec.passed = 0;
    rec.n_pairs = 0;
    rec.pad_ = 0;
    rec.T[0] = rec.T[1] = rec.T[2] = rec.T[3] = 0.0;
    for (int k = 0; k < C2G_PAIR_WORDS; ++k) rec.pair_bits[k] = 0ull;
    rec.pad2_ = 0ull;
    score_hint_serial(heads, views, first_slot + h.q_idx, h, Q, rec);
    scores[hid] = rec;
  }
}

// Stage 1 of the hint cascade, one THREAD per hint slot: the anchor similarity gate (contour_db.h:388) and the 256-bit
// popcount gate of BCI::checkConstellSim (contour_mng.h:291-307) kill ~85 % of the hints with two 80-byte and two 32-byte
// reads each; the record of a dead hint is final here, survivors are queued for the warp-per-hint stage.
__global__ void __launch_bounds__(256)
prefilter_kernel(const c2g_scan_head *__restrict__ heads, const c2g_view *__restrict__ views, int first_slot, long long hid0, long long n_hints,
                 QueryParams Q, const c2g_hint *__restrict__ hints, c2g_pair_score *__restrict__ scores,
                 int *__restrict__ survivors, int *__restrict__ n_surv) {
  const long long hid = hid0 + (long long) blockIdx.x * blockDim.x + threadIdx.x;  // hint slots [hid0, hid0 + n_hints)
  const int lane = threadIdx.x & 31;
  bool survive = false;
  if (hid < hid0 + n_hints) {
    const c2g_hint h = hints[hid];
    c2g_pair_score rec;
    rec.constell[0] = rec.constell[1] = rec.constell[2] = 0;
    rec.pairwise[0] = rec.pairwise[1] = 0;
    rec.passed = 0;
    rec.n_pairs = 0;
    rec.pad_ = 0;
    rec.T[0] = rec.T[1] = rec.T[2] = rec.T[3] = 0.0;
    for (int i = 0; i < C2G_PAIR_WORDS; ++i) rec.pair_bits[i] = 0ull;
    rec.pad2_ = 0ull;
    if (h.cand_gidx >= 0) {
      const int q_slot = first_slot + h.q_idx;
      if (check_sim(view_at(heads, views, h.cand_gidx, h.level, h.cand_seq), view_at(heads, views, q_slot, h.level, h.q_seq), Q.sim)) {
        const c2g_bci &src = heads[h.cand_gidx].bcis[h.level][h.cand_seq];
        const c2g_bci &tgt = heads[q_slot].bcis[h.level][h.q_seq];
        int ov1 = 0, ov2 = 0, ov3 = 0;
        unsigned long long s4[4], t4[4];
        for (int i = 0; i < 4; ++i) {
          s4[i] = src.dist_bin[i];
          t4[i] = tgt.dist_bin[i];
        }
        for (int i = 0; i < 4; ++i) {
          const unsigned long long sl = (s4[i] << 1) | (i > 0 ? (s4[i - 1] >> 63) : 0ull);
          const unsigned long long sr = (s4[i] >> 1) | (i < 3 ? (s4[i + 1] << 63) : 0ull);
          ov1 += __popcll(s4[i] & t4[i]);
          ov2 += __popcll(sl & t4[i]);
          ov3 += __popcll(sr & t4[i]);
        }
        rec.constell[0] = ov1 + ov2 + ov3;
        rec.constell[1] = max(ov1, max(ov2, ov3));
        rec.passed = -1;
        survive = rec.constell[0] >= Q.lb.i_ovlp_sum && rec.constell[1] >= Q.lb.i_ovlp_max_one;
      }
    }
    if (!survive) scores[hid] = rec;  // survivors' records are written by score_kernel
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, survive);
  if (m) {
    int base = 0;
    if (lane == 0) base = atomicAdd(n_surv, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (survive) survivors[base + __popc(m & ((1u << lane) - 1u))] = (int) hid;
  }
}

// Stage 2, one WARP per surviving hint, grid-stride over the survivor list.
__global__ void __launch_bounds__(SC_WARPS * 32)
score_kernel(const c2g_scan_head *__restrict__ heads, const c2g_view *__restrict__ views, int first_slot, QueryParams Q,
             const c2g_hint *__restrict__ hints, c2g_pair_score *__restrict__ scores, const int *__restrict__ survivors,
             const int *__restrict__ n_surv) {
  extern __shared__ __align__(16) unsigned char sc_raw[];
  ScoreScratch *scratch = reinterpret_cast<ScoreScratch *>(sc_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = *n_surv;
  for (int i = blockIdx.x * SC_WARPS + w; i < n; i += gridDim.x * SC_WARPS) {
    const int hid = survivors[i];
    const c2g_hint h = hints[hid];
    c2g_pair_score rec;
    rec.constell[0] = rec.constell[1] = rec.constell[2] = 0;
    rec.pairwise[0] = rec.pairwise[1] = 0;
    rec.passed = 0;
    rec.n_pairs = 0;
    rec.pad_ = 0;
    rec.T[0] = rec.T[1] = rec.T[2] = rec.T[3] = 0.0;
    for (int k = 0; k < C2G_PAIR_WORDS; ++k) rec.pair_bits[k] = 0ull;
    rec.pad2_ = 0ull;
    score_hint(heads, views, first_slot + h.q_idx, h, Q, scratch[w], rec, lane);
    if (lane == 0) scores[hid] = rec;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Proposal replay + tidy-up + GMM-L2 initial correlation, one warp per query scan
// ------------------------------------------------------------------------------------------------------------------
constexpr int FIN_WARPS = 4;  // warps per CTA of the finish kernel: one query scan per CTA, candidates spread over warps

struct Prop {
  double T[4];  // cos, sin, tx, ty
  uint64_t bits[C2G_PAIR_WORDS];
  int vote_cnt;
  float area_perc;
};
struct CandState {
  int gidx, n_prop;
  float corr_init;
  int alive;
  double neg_est_dist;
  Prop prop[C2G_MAX_PROP];
};
// What survives the proposal replay per candidate pose: its selected proposal.  Written by finish_replay_kernel, completed by
// finish_corr_kernel (corr_init, alive), consumed by finish_output_kernel.
struct FinCand {
  int gidx, vote_cnt;
  float area_perc, corr_init;
  double neg_est_dist;
  double T[4];
  int pass;   // passed the area_perc and neg_est_dist gates of tidyUpCandidates: the GMM-L2 initial correlation is due
  int alive;  // passed the correlation gate too
};
struct FinHead {
  int n_before, aft[3], overflow, pad_[3];
};
struct FinishScratch {
  CandState cand[C2G_MAX_CAND];
  int n_cand;
  uint32_t ord[C2G_MAX_CAND];
  float corr[C2G_MAX_CAND];
  uint16_t passlist[C2G_NUM_Q_LEVELS_MAX * C2G_MAX_PIV * 64 + 8];  // hint indices that reached addProposal: one sublist per warp
  int aft[3], overflow;
  int w_aft1[FIN_WARPS], w_aft2[FIN_WARPS], w_npass[FIN_WARPS];
};

__device__ __forceinline__ float cont_perc(const c2g_scan_head *heads, const c2g_view *views, int slot, int level, int seq) {
  // cont_perc_[l][j] = cell_cnt * 1.0f / layer_cell_cnt (contour_mng.h:605-608)
  return (float) view_at(heads, views, slot, level, seq).cell_cnt * 1.0f / (float) heads[slot].layer_cell_cnt[level];
}

// GMM-L2 initial correlation (correlation.h:84-96,125-152,196-202); all lanes of the warp cooperate, every lane returns
// the same value.  src = candidate, tgt = query.  Ellipses come from the compact per-view table the contour kernel wrote
// (one 32-byte sector each).  Two interleaved phases keep the lanes busy: the pre-selection test streams over (source,
// 32 targets) tiles and appends the selected pairs to a small shared-memory queue; whenever 32 pairs are queued their
// (expensive, FP64) terms are evaluated one per lane.
//
// Pre-selection `sqrt(|T mu_s - mu_t|^2) < 3 (sigma_s + sigma_t)`: a float comparison with a 1e-4 relative guard band decides
// all but borderline pairs (float error of both sides is < 1e-5 relative for BEV coordinates), those take the exact path.
__device__ __forceinline__ bool gmm_pair_selected(double qx, double qy, float qxf, float qyf, float amaj, const c2g_ell &b) {
  const float fx = qxf - b.mx, fy = qyf - b.my;
  const float d2 = fx * fx + fy * fy;
  const float y = 3.0f * (amaj + b.maj);
  const float y2 = y * y;
  if (d2 > y2 * 1.0001f) return false;
  if (d2 < y2 * 0.9999f) return true;
  const double ddx = qx - (double) b.mx, ddy = qy - (double) b.my;
  return c2g_sqrt_lt(ddx * ddx + ddy * ddy, 3.0 * (double) (amaj + b.maj));
}

// The warps of the CTA share one pose: warp w takes the source ellipses w, w + n_warps, ... of every level (own queue, own
// partial cost); the partial costs are added in warp order, so the result does not depend on scheduling.
__device__ double gmm_init_corr(const c2g_scan_head *heads, const c2g_ell *ells, int src_slot, int tgt_slot, const double T[4], int lane,
                                int warp, int n_warps, uint32_t *queue /* 64 words of shared memory owned by this warp */,
                                double *partial /* n_warps doubles of shared memory */, unsigned long long *work) {
  const double theta = atan2(T[1], T[0]);
  const double c = cos(theta), s = sin(theta);
  const c2g_ell *se_all = ells + (size_t) src_slot * C2G_VIEW_CAP, *te_all = ells + (size_t) tgt_slot * C2G_VIEW_CAP;
  double cost = 0.0;
  int qn = 0;
  long long n_tests = 0, n_terms = 0;
  auto eval_queued = [&](int cnt) {
    n_terms += cnt;
    if (lane < cnt) {
      const uint32_t pr = queue[lane];
      const c2g_ell a = se_all[pr >> 16], b = te_all[pr & 0xFFFFu];
      const double amx = (double) a.mx, amy = (double) a.my;
      const double a00 = a.c00, a10 = a.c10, a01 = a.c01, a11 = a.c11;
      const double t00 = c * a00 + (-s) * a10, t01 = c * a01 + (-s) * a11;
      const double t10 = s * a00 + c * a10, t11 = s * a01 + c * a11;
      const double ra00 = t00 * c + t01 * (-s), ra01 = t00 * s + t01 * c;
      const double ra10 = t10 * c + t11 * (-s), ra11 = t10 * s + t11 * c;
      const double c00 = 2.0 * (ra00 + (double) b.c00), c10 = 2.0 * (ra10 + (double) b.c10);
      const double c01 = 2.0 * (ra01 + (double) b.c01), c11 = 2.0 * (ra11 + (double) b.c11);
      const double mx = (c * amx + (-s) * amy) + T[2] - (double) b.mx;
      const double my = (s * amx + c * amy) + T[3] - (double) b.my;
      const double det = c00 * c11 - c01 * c10, invdet = 1.0 / det;
      const double qua = -0.5 * (mx * ((c11 * invdet) * mx + (-c01 * invdet) * my) + my * ((-c10 * invdet) * mx + (c00 * invdet) * my));
      cost += -(double) b.w * (double) a.w * 1.0 / sqrt(det) * exp(qua);
    }
  };
  for (int li = 0; li < C2G_NUM_BIN_LAYERS; ++li) {
    const int lev = li + 1;
    const int ns = heads[src_slot].n_ell[li], nt = heads[tgt_slot].n_ell[li];
    const int so = heads[src_slot].view_off[lev], to = heads[tgt_slot].view_off[lev];
    for (int t0 = 0; t0 < nt; t0 += 32) {  // the lane keeps one target ellipse in registers while the sources stream by
      const int ti = t0 + lane;
      c2g_ell b;
      b.mx = b.my = 1.0e30f;
      b.maj = 0.0f;
      if (ti < nt) b = te_all[to + ti];
      n_tests += (long long) min(32, nt - t0) * ((ns - warp + n_warps - 1) / n_warps);
      for (int si = warp; si < ns; si += n_warps) {
        const c2g_ell a = se_all[so + si];  // warp-wide broadcast
        const double amx = (double) a.mx, amy = (double) a.my;
        const double qx = (T[0] * amx + (-T[1]) * amy) + T[2], qy = (T[1] * amx + T[0] * amy) + T[3];
        const bool sel = ti < nt && gmm_pair_selected(qx, qy, (float) qx, (float) qy, a.maj, b);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
        if (m == 0u) continue;
        if (sel) queue[qn + __popc(m & ((1u << lane) - 1u))] = ((uint32_t) (so + si) << 16) | (uint32_t) (to + ti);
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
          eval_queued(32);
          __syncwarp();
          const uint32_t keep = (lane < qn - 32) ? queue[32 + lane] : 0u;
          __syncwarp();
          if (lane < qn - 32) queue[lane] = keep;
          qn -= 32;
          __syncwarp();
        }
      }
    }
  }
  eval_queued(qn);
  for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xFFFFFFFFu, cost, o);
  if (work && lane == 0) {
    atomicAdd(work + 2, (unsigned long long) n_tests);
    atomicAdd(work + 3, (unsigned long long) n_terms);
  }
  if (lane == 0) partial[warp] = cost;
  __syncthreads();
  cost = partial[0];
  for (int w = 1; w < n_warps; ++w) cost += partial[w];
  return -cost / sqrt(heads[src_slot].gmm_auto_corr * heads[tgt_slot].gmm_auto_corr);
}

__device__ void add_proposal(CandState &cs, const double Tp[4], const uint64_t bits[C2G_PAIR_WORDS], int n_pairs) {
  for (int i = 0; i < cs.n_prop; i++) {
    Prop &pr = cs.prop[i];
    // delta_T = T_prop.inverse() * anch.T_delta_
    const double i00 = Tp[0], i01 = Tp[1], i10 = -Tp[1], i11 = Tp[0];  // R^T
    const double itx = -(i00 * Tp[2] + i01 * Tp[3]), ity = -(i10 * Tp[2] + i11 * Tp[3]);
    const double a00 = pr.T[0], a10 = pr.T[1];
    const double d00 = i00 * a00 + i01 * a10, d10 = i10 * a00 + i11 * a10;
    const double dtx = (i00 * pr.T[2] + i01 * pr.T[3]) + itx, dty = (i10 * pr.T[2] + i11 * pr.T[3]) + ity;
    if (sqrt(dtx * dtx + dty * dty) < 2.0 && fabs(atan2(d10, d00)) < 0.3) {
      for (int k = 0; k < C2G_PAIR_WORDS; ++k) pr.bits[k] |= bits[k];
      pr.vote_cnt += n_pairs;
      const int w1 = pr.vote_cnt, w2 = n_pairs;
      const double tbx = (pr.T[2] * w1 + Tp[2] * w2) / (w1 + w2), tby = (pr.T[3] * w1 + Tp[3] * w2) / (w1 + w2);
      const double ang1 = atan2(pr.T[1], pr.T[0]), ang2 = atan2(Tp[1], Tp[0]);
      double diff = ang2 - ang1;
      if (diff < 0) diff += 2 * C2G_PI;
      if (diff > C2G_PI) diff -= 2 * C2G_PI;
      const double ang_bl = diff * w2 / (w1 + w2) + ang1;
      pr.T[0] = cos(ang_bl);
      pr.T[1] = sin(ang_bl);
      pr.T[2] = tbx;
      pr.T[3] = tby;
      return;
    }
  }
  if (cs.n_prop > 3) return;
  Prop &np = cs.prop[cs.n_prop++];
  for (int k = 0; k < 4; ++k) np.T[k] = Tp[k];
  for (int k = 0; k < C2G_PAIR_WORDS; ++k) np.bits[k] = bits[k];
  np.vote_cnt = n_pairs;
  np.area_perc = 0.0f;
}

__global__ void __launch_bounds__(FIN_WARPS * 32)
finish_replay_kernel(const c2g_scan_head *__restrict__ heads, const c2g_view *__restrict__ views, int first_slot, int q0, int B, QueryParams Q,
                     const c2g_hint *__restrict__ hints, const c2g_pair_score *__restrict__ scores, FinHead *__restrict__ fin_head,
                     FinCand *__restrict__ fin_cand) {
  extern __shared__ __align__(16) unsigned char fsm_raw[];
  FinishScratch &F = *reinterpret_cast<FinishScratch *>(fsm_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int q = q0 + blockIdx.x;  // one query scan per CTA
  if (q >= q0 + B) return;
  const int q_slot = first_slot + q;
  const long long per_q = (long long) Q.n_q_levels * C2G_MAX_PIV * Q.nnk;
  const c2g_hint *hq = hints + (size_t) q * per_q;
  const c2g_pair_score *sq = scores + (size_t) q * per_q;
  // the warps scan an equal share of the hint records each (in order); only the few hints that reached addProposal are
  // replayed sequentially afterwards
  {
    const int quarter = (int) ((per_q + FIN_WARPS - 1) / FIN_WARPS);
    const long long beg = (long long) w * quarter, end = beg + quarter < per_q ? beg + quarter : per_q;
    uint16_t *pl = F.passlist + w * quarter;
    int aft1 = 0, aft2 = 0, n_pass = 0;
    for (long long base = beg; base < end; base += 32) {
      const long long i = base + lane;
      int p = 0;
      const bool valid = i < end && hq[i].cand_gidx >= 0;
      if (valid) p = sq[i].passed;
      aft1 += __popc(__ballot_sync(0xFFFFFFFFu, valid && p != 0));
      aft2 += __popc(__ballot_sync(0xFFFFFFFFu, valid && (p == 1 || p == -2)));
      const unsigned pm = __ballot_sync(0xFFFFFFFFu, valid && p == 1);
      if (valid && p == 1) pl[n_pass + __popc(pm & ((1u << lane) - 1u))] = (uint16_t) i;
      n_pass += __popc(pm);
    }
    if (lane == 0) {
      F.w_aft1[w] = aft1;
      F.w_aft2[w] = aft2;
      F.w_npass[w] = n_pass;
    }
  }
  __syncthreads();
  if (w == 0) {
    int overflow = 0;
    const int quarter = (int) ((per_q + FIN_WARPS - 1) / FIN_WARPS);
    if (lane == 0) {
      F.n_cand = 0;
      // replay checkCandWithHint's bookkeeping in reference order: (q-level, query seq, ascending distance)
      int aft1 = 0, aft2 = 0, n_pass = 0;
      for (int ww = 0; ww < FIN_WARPS; ++ww) {
        aft1 += F.w_aft1[ww];
        aft2 += F.w_aft2[ww];
        n_pass += F.w_npass[ww];
      }
      for (int ww = 0; ww < FIN_WARPS; ++ww)
      for (int k0 = 0; k0 < F.w_npass[ww]; ++k0) {
        const int i = F.passlist[ww * quarter + k0];
        const c2g_pair_score &r = sq[i];
        const int gidx = hq[i].cand_gidx;
        int ci = -1;
        for (int k = 0; k < F.n_cand; ++k)
          if (F.cand[k].gidx == gidx) {
            ci = k;
            break;
          }
        if (ci < 0) {
          if (F.n_cand >= C2G_MAX_CAND) {
            overflow = 1;
            continue;
          }
          ci = F.n_cand++;
          F.cand[ci].gidx = gidx;
          F.cand[ci].n_prop = 0;
          F.cand[ci].corr_init = 0.0f;
          F.cand[ci].alive = 0;
          F.cand[ci].neg_est_dist = 0.0;
        }
        add_proposal(F.cand[ci], r.T, r.pair_bits, r.n_pairs);
      }
      F.aft[0] = aft1;
      F.aft[1] = aft2;
      F.aft[2] = n_pass;
      F.overflow = overflow;
    }
  }
  __syncthreads();
  const int n_before = F.n_cand;
  // tidyUpCandidates
  for (int ci = w; ci < n_before; ci += FIN_WARPS) {  // candidate poses are independent: one warp each
    CandState &cs = F.cand[ci];
    int pass = 0;
    // area_perc of every proposal: lanes evaluate the per-pair percentages of 32 map entries at a time, the float sums
    // run in std::map order (ascending (level, seq_src, seq_tgt) == ascending bit index) on all lanes redundantly
    for (int i = 0; i < cs.n_prop; i++) {
      float lev_perc[C2G_NLEV] = {0, 0, 0, 0, 0, 0};
      for (int base = 0; base < C2G_PAIR_BITS; base += 32) {
        const int bit = base + lane;
        const bool set = bit < C2G_PAIR_BITS && ((cs.prop[i].bits[bit >> 6] >> (bit & 63)) & 1ull);
        float pv = 0.0f;
        if (set) {
          const int level = bit / 100 + 1, ss = (bit / 10) % 10, st = bit % 10;
          pv = 0.5f * (cont_perc(heads, views, cs.gidx, level, ss) + cont_perc(heads, views, q_slot, level, st));
        }
        unsigned m = __ballot_sync(0xFFFFFFFFu, set);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float v = __shfl_sync(0xFFFFFFFFu, pv, src);
          lev_perc[(base + src) / 100 + 1] += v;
        }
      }
      const float LW[4] = {0.3f, 0.3f, 0.3f, 0.1f};
      float perc = 0;
      for (int j = 0; j < C2G_NUM_BIN_LAYERS; j++) perc += LW[j] * lev_perc[j + 1];
      if (lane == 0) cs.prop[i].area_perc = perc;
    }
    __syncwarp();
    if (lane == 0) {
      int idx_sel = 0;
      for (int i = 0; i < cs.n_prop; i++)
        if (cs.prop[i].vote_cnt > cs.prop[idx_sel].vote_cnt) idx_sel = i;
      if (idx_sel != 0) {
        const Prop tmp = cs.prop[0];
        cs.prop[0] = cs.prop[idx_sel];
        cs.prop[idx_sel] = tmp;
      }
      pass = 1;
      if (cs.prop[0].area_perc < Q.lb.area_perc) pass = 0;
      if (pass) {
        // getEstSensTF: T_so^-1 * T_delta * T_so with T_so = translate(n_row/2 - 0.5, n_col/2 - 0.5)
        const double ox = Q.n_row / 2 - 0.5, oy = Q.n_col / 2 - 0.5;
        const double *T = cs.prop[0].T;
        const double mx = (T[0] * ox + (-T[1]) * oy) + T[2], my = (T[1] * ox + T[0] * oy) + T[3];  // (T_delta * T_so).translation
        const double ex = (1.0 * mx + 0.0 * my) + (-(1.0 * ox + 0.0 * oy)), ey = (0.0 * mx + 1.0 * my) + (-(0.0 * ox + 1.0 * oy));
        cs.neg_est_dist = -sqrt(ex * ex + ey * ey);
        if (cs.neg_est_dist < (double) Q.lb.neg_est_dist) pass = 0;
      }
    }
    if (lane == 0) {
      FinCand fc;
      fc.gidx = cs.gidx;
      fc.vote_cnt = cs.prop[0].vote_cnt;
      fc.area_perc = cs.prop[0].area_perc;
      fc.corr_init = 0.0f;
      fc.neg_est_dist = cs.neg_est_dist;
      for (int k = 0; k < 4; ++k) fc.T[k] = cs.prop[0].T[k];
      fc.pass = pass;
      fc.alive = 0;
      fin_cand[(size_t) q * C2G_MAX_CAND + ci] = fc;
    }
    __syncwarp();
  }
  if (threadIdx.x == 0) {
    FinHead h;
    h.n_before = n_before;
    h.aft[0] = F.aft[0];
    h.aft[1] = F.aft[1];
    h.aft[2] = F.aft[2];
    h.overflow = F.overflow;
    h.pad_[0] = h.pad_[1] = h.pad_[2] = 0;
    fin_head[q] = h;
  }
}

// tidyUpCandidates' GMM-L2 gate (contour_db.h:553-577): one warp per (query scan, candidate pose)
constexpr int CORR_WARPS = 4;  // warps per pose: the long poses (50 x 50 ellipse pairs per level) set the kernel's makespan
__global__ void __launch_bounds__(CORR_WARPS * 32)
finish_corr_kernel(const c2g_scan_head *__restrict__ heads, const c2g_ell *__restrict__ ells, int first_slot, int q0, int B, float lb_correlation,
                   const FinHead *__restrict__ fin_head, FinCand *__restrict__ fin_cand, unsigned long long *__restrict__ work) {
  __shared__ uint32_t queue[CORR_WARPS][64];
  __shared__ double partial[CORR_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wg = blockIdx.x;  // one pose per CTA: a slot is released as soon as its pose is done (most poses exit at once)
  const int q = q0 + wg / C2G_MAX_CAND, ci = wg % C2G_MAX_CAND;
  if (q >= q0 + B || ci >= fin_head[q].n_before) return;
  FinCand &fc = fin_cand[(size_t) q * C2G_MAX_CAND + ci];
  if (!fc.pass) return;
  const double T[4] = {fc.T[0], fc.T[1], fc.T[2], fc.T[3]};
  const double corr = gmm_init_corr(heads, ells, fc.gidx, first_slot + q, T, lane, warp, CORR_WARPS, queue[warp], partial, work);
  if (threadIdx.x == 0) {
    fc.corr_init = (float) corr;
    fc.alive = (fc.corr_init < lb_correlation) ? 0 : 1;
  }
}

// swap-compaction of the survivors (contour_db.h:580-592), fineOptimize's first std::sort (contour_db.h:616-621; correlation_
// is still 0 for every candidate) and the result records; one warp per query scan.  refine.cu optimises the first
// max_fine_opt records and sorts those by the refined correlation.
__global__ void __launch_bounds__(128)
finish_output_kernel(int q0, int B, const FinHead *__restrict__ fin_head, const FinCand *__restrict__ fin_cand, c2g_query_result *__restrict__ results) {
  __shared__ uint32_t ord_s[4][C2G_MAX_CAND];
  __shared__ int n_s[4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int q = q0 + blockIdx.x * 4 + w;
  if (q >= q0 + B) return;
  const FinHead h = fin_head[q];
  const FinCand *fc = fin_cand + (size_t) q * C2G_MAX_CAND;
  if (lane == 0) {
    uint32_t idx[C2G_MAX_CAND];
    for (int i = 0; i < h.n_before; ++i) idx[i] = (uint32_t) i;
    int p1 = 0, p2 = h.n_before - 1;
    while (p1 <= p2) {
      if (!fc[idx[p1]].alive && fc[idx[p2]].alive) {
        const uint32_t tmp = idx[p1];
        idx[p1] = idx[p2];
        idx[p2] = tmp;
        p1++;
        p2--;
      } else {
        if (fc[idx[p1]].alive) p1++;
        if (!fc[idx[p2]].alive) p2--;
      }
    }
    const int n = p2 + 1;
    uint32_t ord[C2G_MAX_CAND];
    float corr[C2G_MAX_CAND];
    for (int i = 0; i < n; ++i) {
      ord[i] = (uint32_t) i;
      corr[i] = 0.0f;
    }
    const float *cp = corr;
    c2g_sort::std_sort(ord, (long) n, [cp](uint32_t a, uint32_t b) { return cp[a] > cp[b]; });
    for (int i = 0; i < n; ++i) ord_s[w][i] = idx[ord[i]];
    n_s[w] = n;
    c2g_query_result &R = results[q];
    R.n_cand = n;
    R.n_pose_before = h.n_before;
    R.cand_aft_check[0] = h.aft[0];
    R.cand_aft_check[1] = h.aft[1];
    R.cand_aft_check[2] = h.aft[2];
    R.overflow = h.overflow;
    R.best = n > 0 ? 0 : -1;
    R.pad_ = 0;
  }
  __syncwarp();
  const int n = n_s[w];
  static_assert(C2G_MAX_CAND == 32, "one result record per lane");
  c2g_cand c;
  c.cand_gidx = -1;
  c.vote_cnt = 0;
  c.area_perc = 0.f;
  c.corr_init = 0.f;
  c.neg_est_dist = 0.0;
  c.T[0] = c.T[1] = c.T[2] = c.T[3] = 0.0;
  c.corr_fine = 0.f;
  c.fine_iters = -1;
  c.fine_term = 0;
  c.fine_flags = 0;
  if (lane < n) {
    const FinCand &cs = fc[ord_s[w][lane]];
    c.cand_gidx = cs.gidx;
    c.vote_cnt = cs.vote_cnt;
    c.area_perc = cs.area_perc;
    c.corr_init = cs.corr_init;
    c.neg_est_dist = cs.neg_est_dist;
    for (int k = 0; k < 4; ++k) c.T[k] = cs.T[k];
  }
  for (int k = 0; k < 4; ++k) c.T_fine[k] = c.T[k];
  results[q].cand[lane] = c;
}

}  // namespace
int c2g_launch_refine(c2g_ctx *ctx, int first_slot, int q0, int B, cudaStream_t st);  // refine.cu
int c2g_db_sync_mode(c2g_ctx *ctx, int want_kd);
namespace {

int build_query_params(c2g_ctx *ctx, const c2g_score_ensemble *lb, QueryParams &Q) {
  memset(&Q, 0, sizeof(Q));
  Q.n_q_levels = ctx->db.n_q_levels;
  for (int i = 0; i < Q.n_q_levels; ++i) {
    Q.q_levels[i] = ctx->db.q_levels[i];
    const C2gLayerTable &t = ctx->layers[i];
    Q.layer[i].keys_t = t.keys_t;
    Q.layer[i].gidx = t.gidx;
    Q.layer[i].seq = t.seq;
    Q.layer[i].orank = t.orank;
    Q.layer[i].box_min = t.box_min;
    Q.layer[i].box_max = t.box_max;
    Q.layer[i].cap_b = t.cap_b;
    Q.layer[i].blkcap_b = t.blkcap_b;
    Q.layer[i].cap = C2G_PHYS_BUCKETS * t.cap_b;
    Q.layer[i].blk_cap = C2G_PHYS_BUCKETS * t.blkcap_b;
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) Q.layer[i].bucket_cnt[k] = t.bucket_cnt[k];
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) Q.layer[i].phys[k] = t.phys[k];
    for (int k = 0; k <= C2G_NUM_BUCKETS; ++k) Q.layer[i].ranges[k] = t.ranges[k];
  }
  Q.nnk = ctx->db.nnk;
  Q.piv = ctx->P.cfg.piv_firsts;
  Q.sim = ctx->db.cont_sim;
  Q.lb = *lb;
  Q.n_row = ctx->P.cfg.n_row;
  Q.n_col = ctx->P.cfg.n_col;
  return 0;
}

// proposal replay -> GMM-L2 gate -> output -> refinement -> ranking for queries [q0, q0 + B) of the batch, on `st`
int launch_finish(c2g_ctx *ctx, int first_slot, int q0, int B, const QueryParams &Q, const c2g_hint *hints, const c2g_pair_score *scores,
                  cudaStream_t st) {
  static unsigned long long attr_devs = 0ull;
  const size_t smem = sizeof(FinishScratch);
  if (c2g_first_use_on_device(attr_devs)) {
    C2G_CUDA_TRY(cudaFuncSetAttribute(finish_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  }
  FinHead *fh = (FinHead *) ctx->d_fin_head;
  FinCand *fcd = (FinCand *) ctx->d_fin_cand;
  finish_replay_kernel<<<B, FIN_WARPS * 32, smem, st>>>(ctx->d_heads, ctx->d_views, first_slot, q0, B, Q, hints, scores, fh, fcd);
  C2G_CUDA_TRY(cudaGetLastError());
  C2G_QPROF(ctx, 4);
  finish_corr_kernel<<<B * C2G_MAX_CAND, CORR_WARPS * 32, 0, st>>>(ctx->d_heads, ctx->d_ells, first_slot, q0, B, Q.lb.correlation, fh, fcd,
                                                                   ctx->count_work ? ctx->d_work : nullptr);
  C2G_CUDA_TRY(cudaGetLastError());
  C2G_QPROF(ctx, 5);
  finish_output_kernel<<<(B + 3) / 4, 128, 0, st>>>(q0, B, fh, fcd, ctx->d_results);
  C2G_CUDA_TRY(cudaGetLastError());
  C2G_QPROF(ctx, 6);
  ctx->launches += 3;
  return c2g_launch_refine(ctx, first_slot, q0, B, st);
}

}  // namespace

// kd ordering of one bucket, memoised by content: a growing database changes one or two buckets per pushAndBalance
// (contour_db.cpp:63-317), the others keep their blocks
// ---- device mirror of the LayerDB trees: patch records + kernels ------------------------------------------------------------
struct PatchRec {  // one key entering (or moving inside) the mirror
  float k[C2G_KEY_DIM];
  int gidx, pos, orank, seq;
};
struct PatchBlk {  // one 32-key block whose bounding box must be recomputed
  int blk, p0, cnt;
};
static_assert(sizeof(PatchRec) == 56 && sizeof(PatchBlk) == 12, "patch layout");

__global__ void mirror_patch_kernel(const PatchRec *__restrict__ recs, int n, float *__restrict__ keys_t, int stride, int *__restrict__ gidx,
                                    signed char *__restrict__ seq, int *__restrict__ orank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PatchRec r = recs[i];
#pragma unroll
  for (int d = 0; d < C2G_KEY_DIM; ++d) keys_t[(size_t) d * stride + r.pos] = r.k[d];
  gidx[r.pos] = r.gidx;
  seq[r.pos] = (signed char) r.seq;
  orank[r.pos] = r.orank;
}
// one warp per listed block: bounding box of its (<= 32) keys; NaN keys never match anything and fminf / fmaxf ignore them
__global__ void mirror_box_kernel(const PatchBlk *__restrict__ blks, int n, const float *__restrict__ keys_t, int stride,
                                  float *__restrict__ box_min, float *__restrict__ box_max, int blk_stride) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const PatchBlk b = blks[w];
  for (int d = 0; d < C2G_KEY_DIM; ++d) {
    const float first = keys_t[(size_t) d * stride + b.p0];
    const float v = lane < b.cnt ? keys_t[(size_t) d * stride + b.p0 + lane] : first;
    float mn = v, mx = v;
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if (lane == 0) {
      box_min[(size_t) d * blk_stride + b.blk] = mn;
      box_max[(size_t) d * blk_stride + b.blk] = mx;
    }
  }
}

int c2g_query_alloc(c2g_ctx *ctx) {
  ctx->n_hint_slots = (long long) ctx->max_batch * ctx->db.n_q_levels * C2G_MAX_PIV * ctx->db.nnk;
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_hints, sizeof(c2g_hint) * (size_t) ctx->n_hint_slots));
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_scores, sizeof(c2g_pair_score) * (size_t) ctx->n_hint_slots));
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_results, sizeof(c2g_query_result) * (size_t) ctx->max_batch));
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_survivors, sizeof(int) * (size_t) ctx->n_hint_slots));
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_nsurv, sizeof(int) * C2G_QUERY_STREAMS));
  for (int i = 0; i < C2G_QUERY_STREAMS; ++i) {
    C2G_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->qstream[i], cudaStreamNonBlocking));
    C2G_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_qjoin[i], cudaEventDisableTiming));
  }
  C2G_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_qfork, cudaEventDisableTiming));
  C2G_CUDA_TRY(cudaMalloc(&ctx->d_fin_head, sizeof(FinHead) * (size_t) ctx->max_batch));
  C2G_CUDA_TRY(cudaMalloc(&ctx->d_fin_cand, sizeof(FinCand) * (size_t) ctx->max_batch * C2G_MAX_CAND));
  for (int i = 0; i < ctx->db.n_q_levels; ++i) {
    C2gLayerTable &t = ctx->layers[i];
    t.cap_b = ctx->scan_cap * C2G_MAX_PIV;  // any bucket may end up holding every key of the layer
    t.blkcap_b = t.cap_b / 32 + 1;
    const size_t cap = (size_t) C2G_PHYS_BUCKETS * t.cap_b, bcap = (size_t) C2G_PHYS_BUCKETS * t.blkcap_b;
    C2G_CUDA_TRY(cudaMalloc((void **) &t.keys_t, sizeof(float) * C2G_KEY_DIM * cap));
    C2G_CUDA_TRY(cudaMalloc((void **) &t.gidx, sizeof(int) * cap));
    C2G_CUDA_TRY(cudaMalloc((void **) &t.seq, cap));
    C2G_CUDA_TRY(cudaMalloc((void **) &t.orank, sizeof(int) * cap));
    C2G_CUDA_TRY(cudaMalloc((void **) &t.box_min, sizeof(float) * C2G_KEY_DIM * bcap));
    C2G_CUDA_TRY(cudaMalloc((void **) &t.box_max, sizeof(float) * C2G_KEY_DIM * bcap));
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) {
      t.bucket_cnt[k] = 0;
      t.phys[k] = k;
      t.m_n[k] = 0;
      t.m_rv[k] = 0;
      t.m_kd[k] = 0;
      t.m_valid[k] = 0;
    }
    for (int k = 0; k <= C2G_NUM_BUCKETS; ++k) t.ranges[k] = (k == 0) ? -1000.0f : 1000.0f;
  }
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_work, sizeof(unsigned long long) * C2G_WORK_N));
  C2G_CUDA_TRY(cudaMemset(ctx->d_work, 0, sizeof(unsigned long long) * C2G_WORK_N));
  ctx->patch_cap = 16 << 20;
  ctx->patch_off = 0;
  ctx->patch_head = 0;
  C2G_CUDA_TRY(cudaHostAlloc(&ctx->h_patch, ctx->patch_cap, cudaHostAllocDefault));
  C2G_CUDA_TRY(cudaMalloc(&ctx->d_patch, ctx->patch_cap));
  for (int i = 0; i < C2G_PATCH_RING; ++i) {
    ctx->patch_ring[i].used = 0;
    C2G_CUDA_TRY(cudaEventCreateWithFlags(&ctx->patch_ring[i].ev, cudaEventDisableTiming));
  }
  return 0;
}

void c2g_query_free(c2g_ctx *ctx) {
  cudaFree(ctx->d_hints);
  cudaFree(ctx->d_scores);
  cudaFree(ctx->d_results);
  cudaFree(ctx->d_survivors);
  cudaFree(ctx->d_nsurv);
  cudaFree(ctx->d_work);
  for (int i = 0; i < C2G_QUERY_STREAMS; ++i) {
    if (ctx->qstream[i]) cudaStreamDestroy(ctx->qstream[i]);
    if (ctx->ev_qjoin[i]) cudaEventDestroy(ctx->ev_qjoin[i]);
  }
  if (ctx->ev_qfork) cudaEventDestroy(ctx->ev_qfork);
  cudaFree(ctx->d_fin_head);
  cudaFree(ctx->d_fin_cand);
  for (int i = 0; i < C2G_NUM_Q_LEVELS_MAX; ++i) {
    cudaFree(ctx->layers[i].keys_t);
    cudaFree(ctx->layers[i].gidx);
    cudaFree(ctx->layers[i].seq);
    cudaFree(ctx->layers[i].orank);
    cudaFree(ctx->layers[i].box_min);
    cudaFree(ctx->layers[i].box_max);
  }
  if (ctx->h_patch) cudaFreeHost(ctx->h_patch);
  cudaFree(ctx->d_patch);
  for (int i = 0; i < C2G_PATCH_RING; ++i)
    if (ctx->patch_ring[i].ev) cudaEventDestroy(ctx->patch_ring[i].ev);
}

extern "C" {

}  // extern "C"

namespace {

// kd ordering of one bucket for batched queries: the range is split at a multiple of 32 keys along its widest dimension
// (nth_element), recursively, until one 32-key block remains.  order[p] = tree position of the key mirrored at position p.
void kd_order(const C2gKeyRec *tree, int n, std::vector<int> &order) {
  order.resize((size_t) n);
  for (int p = 0; p < n; ++p) order[p] = p;
  std::vector<std::pair<int, int>> stack;
  stack.emplace_back(0, n);
  while (!stack.empty()) {
    const int lo = stack.back().first, hi = stack.back().second;
    stack.pop_back();
    const int m = hi - lo;
    if (m <= 32) continue;
    float mn[C2G_KEY_DIM], mx[C2G_KEY_DIM];
    for (int d = 0; d < C2G_KEY_DIM; ++d) mn[d] = mx[d] = tree[order[lo]].k[d];
    for (int p = lo + 1; p < hi; ++p) {
      const float *kp = tree[order[p]].k;
      for (int d = 0; d < C2G_KEY_DIM; ++d) {
        mn[d] = kp[d] < mn[d] ? kp[d] : mn[d];
        mx[d] = kp[d] > mx[d] ? kp[d] : mx[d];
      }
    }
    int wd = 0;
    float wspan = -1.0f;
    for (int d = 0; d < C2G_KEY_DIM; ++d)
      if (mx[d] - mn[d] > wspan) {
        wspan = mx[d] - mn[d];
        wd = d;
      }
    const int nblk = (m + 31) / 32, left = (nblk / 2) * 32;
    std::nth_element(order.begin() + lo, order.begin() + lo + left, order.begin() + hi, [&](int a, int b2) {
      const float ka = tree[a].k[wd], kb = tree[b2].k[wd];
      return ka < kb || (ka == kb && a < b2);
    });
    stack.emplace_back(lo, lo + left);
    stack.emplace_back(lo + left, hi);
  }
}

// Collects the patch of one layer in host vectors: bucket k of the mirror must end up holding `tree` (n keys).
struct LayerPatch {
  std::vector<PatchRec> recs;
  std::vector<PatchBlk> blks;
};
// what bringing bucket k of the mirror to a tree of n keys (restructure count rv) takes: 0 nothing, 1 an append, 2 a rewrite
int patch_kind(const C2gLayerTable &t, int k, int n, unsigned rv, bool want_kd) {
  const bool grown_only = t.m_valid[k] && t.m_rv[k] == rv && !t.m_kd[k] && n >= t.m_n[k];
  const bool need_kd = want_kd && n > 32;  // a bucket of one block has nothing to order
  if (grown_only && !(need_kd && !t.m_kd[k])) return n == t.m_n[k] ? 0 : 1;
  if (t.m_valid[k] && t.m_rv[k] == rv && n == t.m_n[k] && (t.m_kd[k] || !need_kd)) return 0;  // kd-ordered mirror of an unchanged tree
  return 2;
}
int patch_bucket(C2gLayerTable &t, int k, const C2gKeyRec *tree, int n, unsigned rv, bool want_kd, LayerPatch &out, bool flip = false) {
  const int kind = patch_kind(t, k, n, rv, want_kd);
  if (kind == 0) return 0;
  // flip: a rewrite goes to the bucket's other region, so that a launch enqueued LATER that must still see the old version is not
  // disturbed (the deferred patches of the windowed online loop); everywhere else patches and launches alternate in stream
  // order and the rewrite happens in place
  if (kind == 2 && flip) t.phys[k] = (t.phys[k] + C2G_NUM_BUCKETS) % C2G_PHYS_BUCKETS;
  const int base = t.phys[k] * t.cap_b, bbase = t.phys[k] * t.blkcap_b;
  const int from = kind == 1 ? t.m_n[k] : 0;  // append in tree order
  const bool kd = kind == 2 && want_kd && n > 32;
  std::vector<int> order;
  if (kd) kd_order(tree, n, order);
  for (int p = from; p < n; ++p) {
    const int tp = kd ? order[p] : p;
    PatchRec r;
    for (int d = 0; d < C2G_KEY_DIM; ++d) r.k[d] = tree[tp].k[d];
    r.gidx = tree[tp].gidx;
    r.seq = tree[tp].seq;
    r.pos = base + p;
    r.orank = k * t.cap_b + tp;  // bucket-major tree order, whichever region holds the bucket
    out.recs.push_back(r);
  }
  for (int j = from / 32; j < (n + 31) / 32; ++j) {
    PatchBlk bl;
    bl.blk = bbase + j;
    bl.p0 = base + 32 * j;
    bl.cnt = n - 32 * j < 32 ? n - 32 * j : 32;
    out.blks.push_back(bl);
  }
  t.bucket_cnt[k] = n;
  t.m_n[k] = n;
  t.m_rv[k] = rv;
  t.m_kd[k] = kd ? 1 : 0;
  t.m_valid[k] = 1;
  return kind;
}

// The patch staging ring: a pinned host buffer and its device twin, carved into stretches.  A stretch (and the ring entry that
// describes it) is reused only after the kernels of the patch that last occupied it have run.
int patch_ring_alloc(c2g_ctx *ctx, size_t need, size_t *beg_out) {
  if (need > ctx->patch_cap) {  // a patch larger than the whole ring: grow it (everything in flight is finished first)
    C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    C2G_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    size_t ncap = ctx->patch_cap;
    while (ncap < need) ncap *= 2;
    cudaFreeHost(ctx->h_patch);
    cudaFree(ctx->d_patch);
    ctx->h_patch = ctx->d_patch = nullptr;
    C2G_CUDA_TRY(cudaHostAlloc(&ctx->h_patch, ncap, cudaHostAllocDefault));
    C2G_CUDA_TRY(cudaMalloc(&ctx->d_patch, ncap));
    ctx->patch_cap = ncap;
    ctx->patch_off = 0;
    for (int i = 0; i < C2G_PATCH_RING; ++i) ctx->patch_ring[i].used = 0;
  }
  if (ctx->patch_off + need > ctx->patch_cap) ctx->patch_off = 0;  // wrap
  const size_t beg = ctx->patch_off, end = beg + need;
  for (int i = 0; i < C2G_PATCH_RING; ++i) {
    auto &e = ctx->patch_ring[i];
    if (e.used && (i == ctx->patch_head || (e.beg < end && beg < e.end))) {
      C2G_CUDA_TRY(cudaEventSynchronize(e.ev));
      e.used = 0;
    }
  }
  *beg_out = beg;
  return 0;
}

// marks [beg, beg + need) busy until everything enqueued on the context stream so far has run
int patch_ring_commit(c2g_ctx *ctx, size_t beg, size_t need) {
  auto &e = ctx->patch_ring[ctx->patch_head];
  e.beg = beg;
  e.end = beg + need;
  e.used = 1;
  C2G_CUDA_TRY(cudaEventRecord(e.ev, ctx->stream));
  ctx->patch_head = (ctx->patch_head + 1) % C2G_PATCH_RING;
  ctx->patch_off = beg + need;
  return 0;
}

// the two kernels that apply one layer's uploaded patch (records at dp, blocks at dp + blks_off)
int launch_patch_kernels(c2g_ctx *ctx, C2gLayerTable &t, const char *dp, int n_recs, size_t blks_off, int n_blks) {
  const int stride = C2G_PHYS_BUCKETS * t.cap_b, bstride = C2G_PHYS_BUCKETS * t.blkcap_b;
  if (n_recs > 0)
    mirror_patch_kernel<<<(unsigned) ((n_recs + 255) / 256), 256, 0, ctx->stream>>>((const PatchRec *) dp, n_recs, t.keys_t, stride, t.gidx, t.seq, t.orank);
  if (n_blks > 0)
    mirror_box_kernel<<<(unsigned) ((n_blks * 32 + 255) / 256), 256, 0, ctx->stream>>>((const PatchBlk *) (dp + blks_off), n_blks, t.keys_t, stride, t.box_min,
                                                                                      t.box_max, bstride);
  C2G_CUDA_TRY(cudaGetLastError());
  ctx->launches += 2;
  return 0;
}

// uploads one layer's patch and applies it on the context stream (everything asynchronous)
int apply_patch(c2g_ctx *ctx, C2gLayerTable &t, const LayerPatch &lp) {
  if (lp.recs.empty() && lp.blks.empty()) return 0;
  const size_t rb = lp.recs.size() * sizeof(PatchRec), bb = lp.blks.size() * sizeof(PatchBlk), rb_al = (rb + 255) / 256 * 256;
  const size_t need = (rb_al + bb + 255) / 256 * 256;
  size_t beg = 0;
  int rc = patch_ring_alloc(ctx, need, &beg);
  if (rc) return rc;
  char *hp = (char *) ctx->h_patch + beg;
  const char *dp = (const char *) ctx->d_patch + beg;
  memcpy(hp, lp.recs.data(), rb);
  memcpy(hp + rb_al, lp.blks.data(), bb);
  C2G_CUDA_TRY(cudaMemcpyAsync((void *) dp, hp, rb_al + bb, cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_patch_kernels(ctx, t, dp, (int) lp.recs.size(), rb_al, (int) lp.blks.size());
  if (rc) return rc;
  return patch_ring_commit(ctx, beg, need);
}

// one layer's patch of a window whose upload is deferred: where it sits in the window's staging block
struct DeferredPatch {
  int layer, n_recs, n_blks;
  size_t off, blks_off;  // byte offset of the records in the block; of the blocks relative to the records
};

// would bringing the mirror up to date (tree order) rewrite a bucket that `rewritten` already marks?
bool patches_rewrite_again(c2g_ctx *ctx, const unsigned char (*rewritten)[C2G_NUM_BUCKETS]) {
  for (int ll = 0; ll < ctx->db.n_q_levels; ++ll)
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) {
      const C2gBucket &bk = ctx->hostdb->layers[ll].buckets[k];
      if (rewritten[ll][k] && patch_kind(ctx->layers[ll], k, (int) bk.indexed, bk.restructured, false) == 2) return true;
    }
  return false;
}

// like c2g_db_sync_mode(ctx, 0), but the patches are appended to `block` (host memory) instead of being uploaded; buckets that
// are rewritten (not just appended to) are marked in `rewritten`
int collect_mirror_patches(c2g_ctx *ctx, std::vector<char> &block, std::vector<DeferredPatch> &out, unsigned char (*rewritten)[C2G_NUM_BUCKETS]) {
  for (int ll = 0; ll < ctx->db.n_q_levels; ++ll) {
    const C2gLayerHost &L = ctx->hostdb->layers[ll];
    C2gLayerTable &t = ctx->layers[ll];
    LayerPatch lp;
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) {
      const C2gBucket &bk = L.buckets[k];
      if ((int) bk.tree.size() > t.cap_b) return C2G_ERR_CAPACITY;
      if (patch_bucket(t, k, bk.tree.data(), (int) bk.indexed, bk.restructured, false, lp, true) == 2) rewritten[ll][k] = 1;  // the searchable prefix
    }
    for (int k = 0; k <= C2G_NUM_BUCKETS; ++k) t.ranges[k] = L.ranges[k];
    if (lp.recs.empty() && lp.blks.empty()) continue;
    const size_t rb = lp.recs.size() * sizeof(PatchRec), bb = lp.blks.size() * sizeof(PatchBlk), rb_al = (rb + 255) / 256 * 256;
    const size_t need = (rb_al + bb + 255) / 256 * 256, off = block.size();
    block.resize(off + need);
    memcpy(block.data() + off, lp.recs.data(), rb);
    memcpy(block.data() + off + rb_al, lp.blks.data(), bb);
    out.push_back({ll, (int) lp.recs.size(), (int) lp.blks.size(), off, rb_al});
  }
  return 0;
}

}  // namespace

// Brings the device mirror up to date with the host trees by patching what changed since the last call: keys appended to a
// tree (the common case of the online loop: LayerDB::rebuild pops aged buffer entries to the END of a tree) are appended to
// the bucket's region; a bucket whose tree was permuted or cut by a rebalancing move is rewritten.  want_kd: kd-block the
// buckets (worth it for batched queries over a static DB; the online loop mirrors in tree order).
int c2g_db_sync_mode(c2g_ctx *ctx, int want_kd) {
  if (!ctx) return C2G_ERR_ARG;
  for (int ll = 0; ll < ctx->db.n_q_levels; ++ll) {
    const C2gLayerHost &L = ctx->hostdb->layers[ll];
    C2gLayerTable &t = ctx->layers[ll];
    LayerPatch lp;
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) {
      const C2gBucket &bk = L.buckets[k];
      if ((int) bk.tree.size() > t.cap_b) return C2G_ERR_CAPACITY;
      patch_bucket(t, k, bk.tree.data(), (int) bk.indexed, bk.restructured, want_kd != 0, lp);  // the searchable prefix (C2gBucket::indexed)
    }
    for (int k = 0; k <= C2G_NUM_BUCKETS; ++k) t.ranges[k] = L.ranges[k];
    int rc = apply_patch(ctx, t, lp);
    if (rc) return rc;
  }
  return 0;
}

extern "C" {

int c2g_db_set_layer(c2g_ctx *ctx, int ll, int n, const float *keys_host, const int *gidx_host, const signed char *seq_host,
                     const unsigned char *bucket_host, const float *bucket_ranges_host) {
  if (!ctx || ll < 0 || ll >= ctx->db.n_q_levels || n < 0 || !bucket_ranges_host) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2gLayerTable &t = ctx->layers[ll];
  if (n > 0 && (!keys_host || !gidx_host || !seq_host || !bucket_host)) return C2G_ERR_ARG;
  std::vector<C2gKeyRec> trees[C2G_NUM_BUCKETS];
  for (int i = 0; i < n; ++i) {
    if (bucket_host[i] >= C2G_NUM_BUCKETS) return C2G_ERR_ARG;
    C2gKeyRec r;
    for (int d = 0; d < C2G_KEY_DIM; ++d) r.k[d] = keys_host[(size_t) i * C2G_KEY_DIM + d];
    r.gidx = gidx_host[i];
    r.seq = seq_host[i];
    trees[bucket_host[i]].push_back(r);
  }
  LayerPatch lp;
  for (int k = 0; k < C2G_NUM_BUCKETS; ++k) {
    if ((int) trees[k].size() > t.cap_b) return C2G_ERR_CAPACITY;
    t.m_valid[k] = 0;  // caller-provided contents: always rewritten, and the next c2g_db_sync rewrites them again
    patch_bucket(t, k, trees[k].data(), (int) trees[k].size(), 0u, true, lp);
    t.m_valid[k] = 0;
  }
  for (int k = 0; k <= C2G_NUM_BUCKETS; ++k) t.ranges[k] = bucket_ranges_host[k];
  int rc = apply_patch(ctx, t, lp);
  if (rc) return rc;
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"

static bool thresholds_ok(const c2g_score_ensemble *lb, const c2g_score_ensemble *ub) {
  // CHECKs of the CandidateManager constructor (contour_db.h:365-367)
  return lb->i_ovlp_sum < ub->i_ovlp_sum && lb->i_ovlp_max_one < ub->i_ovlp_max_one && lb->i_in_ang_rng < ub->i_in_ang_rng &&
         lb->i_indiv_sim < ub->i_indiv_sim && lb->i_orie_sim < ub->i_orie_sim && lb->correlation < ub->correlation &&
         lb->area_perc < ub->area_perc && lb->neg_est_dist < ub->neg_est_dist;
}

// kNN of queries [q0, q0 + Bs) of the batch on `st`: against the mirror AS IT IS NOW (the table sizes and bucket boundaries
// travel by value in the launch parameters, so a later patch of the mirror does not affect this launch), or, with `ver`
// (device, [Bs][n_q_levels]), each query scan against its own recorded state of the trees
static int launch_knn(c2g_ctx *ctx, int first_slot, int q0, int Bs, const QueryParams &Q, cudaStream_t st, const KnnVersion *ver = nullptr) {
  const int n_keys = Bs * Q.n_q_levels * C2G_MAX_PIV;
  if (n_keys <= 0) return 0;
  const unsigned grid = (unsigned) ((n_keys + QK_WARPS - 1) / QK_WARPS);
  unsigned long long *work = ctx->count_work ? ctx->d_work : nullptr;
  bool primary = true;  // every bucket in its primary region (always, unless the windowed online loop rewrote a bucket)
  for (int ll = 0; ll < Q.n_q_levels; ++ll)
    for (int k = 0; k < C2G_NUM_BUCKETS; ++k) primary = primary && Q.layer[ll].phys[k] == k;
  if (ver)
    knn_kernel<2><<<grid, QK_WARPS * 32, 0, st>>>(ctx->d_heads, first_slot, q0, Bs, Q, ver, ctx->d_hints, work);
  else if (primary)
    knn_kernel<0><<<grid, QK_WARPS * 32, 0, st>>>(ctx->d_heads, first_slot, q0, Bs, Q, ver, ctx->d_hints, work);
  else
    knn_kernel<1><<<grid, QK_WARPS * 32, 0, st>>>(ctx->d_heads, first_slot, q0, Bs, Q, ver, ctx->d_hints, work);
  C2G_CUDA_TRY(cudaGetLastError());
  ctx->launches += 1;
  return 0;
}

// The kernel chain of ContourDB::queryRangedKNN for the B query scans in slots first_slot.. : (kNN ->) prefilter -> score ->
// proposal replay -> GMM-L2 gate -> output -> refinement -> ranking.  with_knn == 0: the hints are already in ctx->d_hints
// (the windowed online loop fills them with its own versioned launch, c2g_online_commit).
// The batch is cut into sub-batches that run the whole chain on their own streams: most kernels of the chain are
// latency-bound (sequential solver / replay logic) and leave issue slots idle that the kernels of the other sub-batches
// fill.  C2G_QUERY_SPLIT=1 (or an active c2g_query_profile) keeps everything on the context's stream.
static int launch_query_chain(c2g_ctx *ctx, int first_slot, int B, const QueryParams &Q, int with_knn) {
  static const int split_env = getenv("C2G_QUERY_SPLIT") ? atoi(getenv("C2G_QUERY_SPLIT")) : 4;
  int n_sub = ctx->prof_on ? 1 : (split_env < 1 ? 1 : (split_env > C2G_QUERY_STREAMS ? C2G_QUERY_STREAMS : split_env));
  if (B < 2 * n_sub) n_sub = 1;
  C2G_CUDA_TRY(cudaMemsetAsync(ctx->d_nsurv, 0, sizeof(int) * C2G_QUERY_STREAMS, ctx->stream));
  if (n_sub > 1) C2G_CUDA_TRY(cudaEventRecord(ctx->ev_qfork, ctx->stream));
  const long long per_q = (long long) Q.n_q_levels * C2G_MAX_PIV * Q.nnk;
  for (int sb = 0; sb < n_sub; ++sb) {
    const int q0 = (int) ((long long) B * sb / n_sub), q1 = (int) ((long long) B * (sb + 1) / n_sub), Bs = q1 - q0;
    cudaStream_t st = n_sub > 1 ? ctx->qstream[sb] : ctx->stream;
    if (n_sub > 1) C2G_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_qfork, 0));
    C2G_QPROF(ctx, 0);
    if (with_knn) {
      int rc = launch_knn(ctx, first_slot, q0, Bs, Q, st);
      if (rc) return rc;
    }
    C2G_QPROF(ctx, 1);
    const long long hid0 = (long long) q0 * per_q, n_hints = (long long) Bs * per_q;
    int *surv = ctx->d_survivors + hid0, *nsurv = ctx->d_nsurv + sb;
    prefilter_kernel<<<(unsigned) ((n_hints + 255) / 256), 256, 0, st>>>(ctx->d_heads, ctx->d_views, first_slot, hid0, n_hints, Q, ctx->d_hints,
                                                                        ctx->d_scores, surv, nsurv);
    C2G_CUDA_TRY(cudaGetLastError());
    C2G_QPROF(ctx, 2);
    // thread-per-survivor scoring: the survivor count lives on the device, so the grid covers the worst case sparsely and
    // strides (about 15 % of the hints survive the prefilter).  C2G_SCORE_WARP=1 selects the warp-per-survivor variant.
    static const bool warp_variant = getenv("C2G_SCORE_WARP") && atoi(getenv("C2G_SCORE_WARP")) != 0;
    if (!warp_variant) {
      const long long want = (n_hints + 127) / 128;
      const long long cap = (long long) ctx->num_sms * 8;
      score_thread_kernel<<<(unsigned) (want < cap ? want : cap), 128, 0, st>>>(ctx->d_heads, ctx->d_views, first_slot, Q, ctx->d_hints, ctx->d_scores, surv,
                                                                               nsurv);
    } else {
      const long long want = (n_hints + SC_WARPS - 1) / SC_WARPS;
      const long long cap = (long long) ctx->num_sms * 16;
      static unsigned long long sc_attr_devs = 0ull;
      const size_t sc_smem = sizeof(ScoreScratch) * SC_WARPS;
      if (c2g_first_use_on_device(sc_attr_devs)) {
        C2G_CUDA_TRY(cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sc_smem));
      }
      score_kernel<<<(unsigned) (want < cap ? want : cap), SC_WARPS * 32, sc_smem, st>>>(ctx->d_heads, ctx->d_views, first_slot, Q, ctx->d_hints, ctx->d_scores,
                                                                                      surv, nsurv);
    }
    C2G_CUDA_TRY(cudaGetLastError());
    C2G_QPROF(ctx, 3);
    ctx->launches += 2;
    int rc = launch_finish(ctx, first_slot, q0, Bs, Q, ctx->d_hints, ctx->d_scores, st);
    if (rc) return rc;
    if (n_sub > 1) {
      C2G_CUDA_TRY(cudaEventRecord(ctx->ev_qjoin[sb], st));
      C2G_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_qjoin[sb], 0));
    }
  }
  return 0;
}

extern "C" int c2g_query_async(c2g_ctx *ctx, int first_slot, int B, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub) {
  if (!ctx || !lb || !ub || B <= 0 || B > ctx->max_batch || first_slot < 0 || first_slot + B > ctx->scan_cap) return C2G_ERR_ARG;
  if (!thresholds_ok(lb, ub)) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  {
    // the online loop (small B) mirrors the trees in tree order (cheap appends); batched queries kd-block the buckets once
    const int want_kd = B >= 32;
    if (ctx->db_dirty || (want_kd && ctx->db_not_kd)) {
      int rc = c2g_db_sync_mode(ctx, want_kd);
      if (rc) return rc;
      ctx->db_dirty = 0;
      ctx->db_not_kd = want_kd ? 0 : 1;
    }
  }
  QueryParams Q;
  build_query_params(ctx, lb, Q);
  return launch_query_chain(ctx, first_slot, B, Q, 1);
}

// ---- windowed online loop -------------------------------------------------------------------------------------------
// W iterations of BatchBinSpinner::spinOnce's database half (test/batch_bin_test.cpp:179,234,237):
//     queryRangedKNN(scan_i)  ->  addScan(scan_i, ts_i)  ->  pushAndBalance(seed_i, ts_i)
// The loop is serial only in the DB state, and with DYNAMIC_THRES=0 (CMakeLists.txt:21) the only part of a query that
// depends on that state is WHICH keys sit in WHICH tree at that instant (contour_db.cpp:63-317, contour_db.h:102-143): the
// descriptors were ingested in one batch (c2g_online_stage), their keys are on the host, so the LayerDB bookkeeping of all W
// scans is replayed here on the host; the window is cut into RUNS of consecutive scans that see the same trees; each run's kNN
// is launched against the mirror in exactly that state, then the mirror is patched (appended keys, occasionally a rewritten
// bucket) for the next run - everything in stream order, nothing waits for the GPU.  The rest of the chain (hint cascade,
// proposal replay, GMM-L2, refinement) does not read the trees and runs ONCE for the whole window.
int c2g_online_commit_impl(c2g_ctx *ctx, int first_slot, int W, const float *keys_host, const double *ts_host, const int *seeds_host,
                           const c2g_score_ensemble *lb, const c2g_score_ensemble *ub, c2g_query_result *results_host) {
  if (!thresholds_ok(lb, ub)) return C2G_ERR_ARG;
  C2gHostDB &db = *ctx->hostdb;
  if (first_slot != db.n_scans) return C2G_ERR_STATE;  // gidx == all_bevs_.size() at the time of addScan
  auto sync_mirror = [&]() -> int {
    int rc = c2g_db_sync_mode(ctx, 0);  // tree order: appends are patches of a few hundred bytes
    if (rc) return rc;
    ctx->db_dirty = 0;
    ctx->db_not_kd = 1;
    return 0;
  };
  if (ctx->db_dirty) {
    int rc = sync_mirror();
    if (rc) return rc;
  }
  // Pass 1 (host only): the LayerDB bookkeeping of the whole window.  Every change of a tree ends a run of scans that saw the
  // same trees; every scan's view of the trees (bucket boundaries, sizes, regions) goes into a version table, and the mirror
  // patch that follows the run is appended to one block.  Runs are gathered into groups that ONE kNN launch can serve: all
  // patches of a group are applied before its launch, which is safe as long as no bucket is rewritten twice inside the group
  // (appends are invisible to scans with a shorter prefix, one rewrite goes to the bucket's other region: C2gLayerTable).
  struct Group {
    int q0, n, patch_begin, patch_end;
  };
  const int nql = ctx->db.n_q_levels;
  std::vector<Group> groups;
  std::vector<char> block;
  std::vector<DeferredPatch> patches;
  std::vector<KnnVersion> vers((size_t) W * nql);
  unsigned char rewritten[C2G_NUM_Q_LEVELS_MAX][C2G_NUM_BUCKETS] = {};
  QueryParams Q;
  build_query_params(ctx, lb, Q);
  int run_begin = 0, group_begin = 0, group_patch_begin = 0, n_runs = 0;
  const size_t kstride = (size_t) C2G_NLEV * C2G_MAX_PIV * C2G_KEY_DIM;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_mark = now();
  auto lap = [&](int k) {
    const double t = now();
    ctx->online_host_s[k] += t - t_mark;
    t_mark = t;
  };
  auto close_run = [&](int end) {  // scans run_begin..end-1 see the mirror as the tables describe it now
    if (end <= run_begin) return;
    for (int ll = 0; ll < nql; ++ll) {
      const C2gLayerTable &t = ctx->layers[ll];
      KnnVersion v;
      for (int k = 0; k < C2G_NUM_BUCKETS; ++k) {
        v.phys[k] = t.phys[k];
        v.bucket_cnt[k] = t.bucket_cnt[k];
      }
      for (int k = 0; k <= C2G_NUM_BUCKETS; ++k) v.ranges[k] = t.ranges[k];
      v.pad_ = 0;
      for (int q = run_begin; q < end; ++q) vers[(size_t) q * nql + ll] = v;
    }
    run_begin = end;
    ++n_runs;
  };
  auto close_group = [&](int end) {
    groups.push_back({group_begin, end - group_begin, group_patch_begin, (int) patches.size()});
    group_begin = end;
    group_patch_begin = (int) patches.size();
    memset(rewritten, 0, sizeof(rewritten));
  };
  for (int i = 0; i < W; ++i) {
    const unsigned long long v0 = db.tree_version;
    const float *sk = keys_host + (size_t) i * kstride;
    for (int ll = 0; ll < nql; ++ll) {
      const int lev = ctx->db.q_levels[ll];
      for (int seq = 0; seq < ctx->P.cfg.piv_firsts; ++seq)
        c2g_hostdb_push(db, ll, sk + ((size_t) lev * C2G_MAX_PIV + seq) * C2G_KEY_DIM, ts_host[i], first_slot + i, seq);
    }
    db.n_scans++;
    c2g_hostdb_push_and_balance(db, seeds_host[i], ts_host[i]);
    if (db.tree_version != v0) {  // scans run_begin..i saw the trees as mirrored now; scan i + 1 sees the new ones
      lap(0);
      close_run(i + 1);
      if (patches_rewrite_again(ctx, rewritten)) close_group(i + 1);
      int rc = collect_mirror_patches(ctx, block, patches, rewritten);
      if (rc) return rc;
      ctx->db_dirty = 0;
      ctx->db_not_kd = 1;
      lap(2);
    }
  }
  close_run(W);
  close_group(W);
  lap(0);
  // Pass 2: ONE upload of the window's patches and version table, on the copy stream (all host->device traffic of the online
  // loop is issued on that stream in the order it is needed, so a small copy never sits in the copy queue in front of the next
  // window's points waiting for kernels of this window), then per group: its patches, then its kNN launch.
  const size_t ver_off = (block.size() + 255) / 256 * 256, ver_bytes = vers.size() * sizeof(KnnVersion);
  block.resize(ver_off + ver_bytes);
  memcpy(block.data() + ver_off, vers.data(), ver_bytes);
  size_t beg = 0;
  int rc = patch_ring_alloc(ctx, block.size(), &beg);
  if (rc) return rc;
  memcpy((char *) ctx->h_patch + beg, block.data(), block.size());
  const char *dp = (const char *) ctx->d_patch + beg;
  C2G_CUDA_TRY(cudaMemcpyAsync((void *) dp, (char *) ctx->h_patch + beg, block.size(), cudaMemcpyHostToDevice, ctx->copy_stream));
  C2G_CUDA_TRY(cudaEventRecord(ctx->ev_patch_up, ctx->copy_stream));
  C2G_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_patch_up, 0));
  lap(2);
  c2g_trace_mark(ctx, "knn_runs_begin", ctx->stream);
  const KnnVersion *d_vers = (const KnnVersion *) (dp + ver_off);
  for (const Group &g : groups) {
    for (int p = g.patch_begin; p < g.patch_end; ++p) {
      const DeferredPatch &d = patches[p];
      rc = launch_patch_kernels(ctx, ctx->layers[d.layer], dp + d.off, d.n_recs, d.blks_off, d.n_blks);
      if (rc) return rc;
    }
    rc = launch_knn(ctx, first_slot, g.q0, g.n, Q, ctx->stream, d_vers + (size_t) g.q0 * nql);
    if (rc) return rc;
  }
  rc = patch_ring_commit(ctx, beg, block.size());
  if (rc) return rc;
  ctx->online_runs += n_runs;
  ctx->online_groups += (long long) groups.size();
  lap(1);
  c2g_trace_mark(ctx, "chain_begin", ctx->stream);
  rc = launch_query_chain(ctx, first_slot, W, Q, 0);
  if (rc) return rc;
  c2g_trace_mark(ctx, "chain_end", ctx->stream);
  lap(3);
  if (results_host)
    C2G_CUDA_TRY(cudaMemcpyAsync(results_host, ctx->d_results, sizeof(c2g_query_result) * (size_t) W, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

extern "C" {

int c2g_query(c2g_ctx *ctx, int first_slot, int B, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub,
              c2g_query_result *results_host, c2g_hint *hints_host, c2g_pair_score *scores_host) {
  if (!ctx || !results_host) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  int rc = c2g_query_async(ctx, first_slot, B, lb, ub);
  if (rc) return rc;
  const size_t nh = (size_t) B * ctx->db.n_q_levels * C2G_MAX_PIV * ctx->db.nnk;
  C2G_CUDA_TRY(cudaMemcpyAsync(results_host, ctx->d_results, sizeof(c2g_query_result) * (size_t) B, cudaMemcpyDeviceToHost, ctx->stream));
  if (hints_host) C2G_CUDA_TRY(cudaMemcpyAsync(hints_host, ctx->d_hints, sizeof(c2g_hint) * nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (scores_host) C2G_CUDA_TRY(cudaMemcpyAsync(scores_host, ctx->d_scores, sizeof(c2g_pair_score) * nh, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int c2g_query_buffers(c2g_ctx *ctx, void **results_dev, void **hints_dev, void **scores_dev, long long *n_hint_slots) {
  if (!ctx) return C2G_ERR_ARG;
  if (results_dev) *results_dev = ctx->d_results;
  if (hints_dev) *hints_dev = ctx->d_hints;
  if (scores_dev) *scores_dev = ctx->d_scores;
  if (n_hint_slots) *n_hint_slots = ctx->n_hint_slots;
  return 0;
}

int c2g_query_export(c2g_ctx *ctx, int B, void *hints_dst_dev, void *scores_dst_dev, void *results_dst_host) {
  if (!ctx || B <= 0 || B > ctx->max_batch) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  const size_t nh = (size_t) B * ctx->db.n_q_levels * C2G_MAX_PIV * ctx->db.nnk;
  if (hints_dst_dev) C2G_CUDA_TRY(cudaMemcpyAsync(hints_dst_dev, ctx->d_hints, sizeof(c2g_hint) * nh, cudaMemcpyDeviceToDevice, ctx->stream));
  if (scores_dst_dev) C2G_CUDA_TRY(cudaMemcpyAsync(scores_dst_dev, ctx->d_scores, sizeof(c2g_pair_score) * nh, cudaMemcpyDeviceToDevice, ctx->stream));
  if (results_dst_host)
    C2G_CUDA_TRY(cudaMemcpyAsync(results_dst_host, ctx->d_results, sizeof(c2g_query_result) * (size_t) B, cudaMemcpyDefault, ctx->stream));
  return 0;
}

int c2g_finish_from_scores(c2g_ctx *ctx, int first_slot, int B, const c2g_score_ensemble *lb, const void *hints_dev,
                           const void *scores_dev, c2g_query_result *results_host) {
  if (!ctx || !lb || !hints_dev || !scores_dev || B <= 0 || B > ctx->max_batch || first_slot < 0 || first_slot + B > ctx->scan_cap) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  QueryParams Q;
  build_query_params(ctx, lb, Q);
  int rc = launch_finish(ctx, first_slot, 0, B, Q, (const c2g_hint *) hints_dev, (const c2g_pair_score *) scores_dev, ctx->stream);
  if (rc) return rc;
  if (results_host) {
    C2G_CUDA_TRY(cudaMemcpyAsync(results_host, ctx->d_results, sizeof(c2g_query_result) * (size_t) B, cudaMemcpyDeviceToHost, ctx->stream));
    C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

}  // extern "C"
