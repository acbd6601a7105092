// cont2_facade.cpp — ContourManager / ContourDB (reference names and call sequence) implemented on the C-ABI of
// libc2g.so.  Error convention of the reference: no exceptions, CHECK-style abort on failure.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cont2/contour_db.h"

#define C2G_CHECK(expr)                                                              \
  do {                                                                               \
    int rc__ = (expr);                                                               \
    if (rc__ != 0) {                                                                 \
      std::fprintf(stderr, "CHECK failed: %s -> %d (%s:%d)\n", #expr, rc__, __FILE__, __LINE__); \
      std::abort();                                                                  \
    }                                                                                \
  } while (0)

namespace c2g_host {

struct Runtime {
  c2g_ctx *ctx = nullptr;
  c2g_cm_config cm{};
  c2g_db_config db{};
  bool have_cm = false, have_db = false;
  int capacity = 8192;
  int n_staging = 64;          // slots [capacity - n_staging, capacity) hold scans that are not (yet) in the DB
  int max_window = 1;          // C2G_WINDOW
  std::vector<int> free_slots;

  // One context per process: every ContourManager / ContourDB of the process must carry the SAME configuration (the reference
  // keeps a config per object; two different ones in one process - a KITTI and a MulRan manager, say - cannot share the
  // device context, so that is a CHECK failure instead of silently running the second one with the first one's parameters).
  static void checkSame(const void *a, const void *b, size_t n, const char *what) {
    if (std::memcmp(a, b, n) != 0) {
      std::fprintf(stderr, "CHECK failed: a second, different %s reached the GPU runtime (one configuration per process)\n", what);
      std::abort();
    }
  }
  void setCm(const ContourManagerConfig &c) {
    const c2g_cm_config prev = cm;
    fillCm(c);
    if (have_cm) checkSame(&prev, &cm, sizeof(cm), "ContourManagerConfig");
    have_cm = true;
  }
  void fillCm(const ContourManagerConfig &c) {
    std::memset(&cm, 0, sizeof(cm));
    if (c.lv_grads_.size() != C2G_NLEV) {
      std::fprintf(stderr, "CHECK failed: lv_grads_ must have %d levels\n", C2G_NLEV);
      std::abort();
    }
    for (int i = 0; i < C2G_NLEV; ++i) cm.lv_grads[i] = c.lv_grads_[i];
    cm.n_levels = C2G_NLEV;
    cm.reso_row = c.reso_row_;
    cm.reso_col = c.reso_col_;
    cm.n_row = c.n_row_;
    cm.n_col = c.n_col_;
    cm.lidar_height = c.lidar_height_;
    cm.blind_sq = c.blind_sq_;
    cm.min_cont_key_cnt = c.min_cont_key_cnt_;
    cm.min_cont_cell_cnt = c.min_cont_cell_cnt_;
    cm.piv_firsts = c.piv_firsts_;
    cm.dist_firsts = c.dist_firsts_;
    cm.roi_radius = c.roi_radius_;
    ContourViewStatConfig vs;
    cm.min_cell_cov = vs.min_cell_cov;
    cm.point_sigma = vs.point_sigma;
    cm.com_bias_thres = vs.com_bias_thres;
  }
  void setDb(const ContourDBConfig &c) {
    const c2g_db_config prev = db;
    fillDb(c);
    if (have_db) checkSame(&prev, &db, sizeof(db), "ContourDBConfig");
    have_db = true;
  }
  void fillDb(const ContourDBConfig &c) {
    std::memset(&db, 0, sizeof(db));
    db.nnk = c.nnk_;
    db.max_fine_opt = c.max_fine_opt_;
    db.n_q_levels = (int) c.q_levels_.size();
    for (int i = 0; i < db.n_q_levels && i < C2G_NUM_Q_LEVELS_MAX; ++i) db.q_levels[i] = c.q_levels_[i];
    db.cont_sim.ta_cell_cnt = c.cont_sim_cfg_.ta_cell_cnt;
    db.cont_sim.tp_cell_cnt = c.cont_sim_cfg_.tp_cell_cnt;
    db.cont_sim.tp_eigval = c.cont_sim_cfg_.tp_eigval;
    db.cont_sim.ta_h_bar = c.cont_sim_cfg_.ta_h_bar;
    db.cont_sim.ta_rcom = c.cont_sim_cfg_.ta_rcom;
    db.cont_sim.tp_rcom = c.cont_sim_cfg_.tp_rcom;
    db.max_elapse = c.tb_cfg_.max_elapse_;
    db.min_elapse = c.tb_cfg_.min_elapse_;
  }
  void ensure() {
    if (ctx) return;
    if (!have_cm) {
      std::fprintf(stderr, "CHECK failed: no ContourManagerConfig seen before the first GPU call\n");
      std::abort();
    }
    if (!have_db) {  // a ContourManager used before any ContourDB exists: defaults of config/batch_bin_test_config.yaml;
      ContourDBConfig d;  // a ContourDB created later must then carry exactly these (checkSame)
      d.q_levels_ = {1, 2, 3};
      setDb(d);
    }
    if (const char *e = std::getenv("C2G_SCAN_CAPACITY")) capacity = std::atoi(e);
    int dev = 0;
    if (const char *e = std::getenv("C2G_DEVICE")) dev = std::atoi(e);
    // the largest window ContourDB::queryAddBalanceWindow may be given (1 = scan-by-scan calls only); a scan holds at most
    // 1 000 000 floats = 250 000 points (tools/pointcloud_util.h:17)
    if (const char *e = std::getenv("C2G_WINDOW")) max_window = std::atoi(e) > 1 ? std::atoi(e) : 1;
    C2G_CHECK(c2g_create(&cm, &db, dev, capacity, max_window, (long long) max_window * (1 << 18), &ctx));
    for (int s = capacity - 1; s >= capacity - n_staging; --s) free_slots.push_back(s);
  }
  int acquire() {
    ensure();
    if (free_slots.empty()) {
      std::fprintf(stderr, "CHECK failed: more than %d scans alive outside the database\n", n_staging);
      std::abort();
    }
    int s = free_slots.back();
    free_slots.pop_back();
    return s;
  }
  void release(int s) {
    if (s >= capacity - n_staging) free_slots.push_back(s);
  }
  ~Runtime() {
    if (ctx) c2g_destroy(ctx);
  }
};

Runtime &runtime() {
  static Runtime r;
  return r;
}

c2g_ctx *context() {
  runtime().ensure();
  return runtime().ctx;
}

}  // namespace c2g_host

using c2g_host::runtime;

// ---------------------------------------------------------------------------------------------------------------------
ContourManager::ContourManager(const ContourManagerConfig &config, int int_id) : cfg_(config), int_id_(int_id) {
  if (cfg_.n_col_ % 2 != 0 || cfg_.n_row_ % 2 != 0) {  // CHECKs of the reference constructor (contour_mng.h:479-480)
    std::fprintf(stderr, "CHECK failed: n_row_/n_col_ must be even\n");
    std::abort();
  }
  runtime().setCm(config);
  std::memset(&head_, 0, sizeof(head_));
  cont_views_.resize(cfg_.lv_grads_.size());
  layer_keys_.resize(cfg_.lv_grads_.size());
  layer_key_bcis_.resize(cfg_.lv_grads_.size());
}

ContourManager::~ContourManager() {
  if (owns_slot_ && slot_ >= 0) runtime().release(slot_);
}

namespace {
float *g_pinned_buf = nullptr;
size_t g_pinned_floats = 0;
}  // namespace

float *ContourManager::pinnedScanBuffer(size_t n_floats) {
  if (n_floats > g_pinned_floats) {
    if (g_pinned_buf) C2G_CHECK(c2g_host_free(g_pinned_buf));
    void *p = nullptr;
    C2G_CHECK(c2g_host_alloc(&p, n_floats * sizeof(float)));
    g_pinned_buf = static_cast<float *>(p);
    g_pinned_floats = n_floats;
  }
  return g_pinned_buf;
}

void ContourManager::makeBEVFromBin(const float *xyzi, size_t n_points, std::string str_id) {
  if (g_pinned_buf && xyzi >= g_pinned_buf && xyzi + 4 * n_points <= g_pinned_buf + g_pinned_floats) {
    ext_pts_ = xyzi;  // page-locked and owned by the runtime: no host copy
    ext_n_ = n_points;
    pts_.clear();
  } else {
    ext_pts_ = nullptr;
    pts_.assign(xyzi, xyzi + 4 * n_points);
  }
  str_id_ = std::move(str_id);
}

void ContourManager::makeContoursRecurs() {
  const float *src = ext_pts_ ? ext_pts_ : pts_.data();
  const size_t n_points = ext_pts_ ? ext_n_ : pts_.size() / 4;
  if (n_points <= 10) {  // CHECK_GT(ptr_gapc->size(), 10) of makeBEV (contour_mng.h:507)
    std::fprintf(stderr, "CHECK failed: point cloud has <= 10 points\n");
    std::abort();
  }
  auto &rt = runtime();
  if (slot_ < 0) {
    slot_ = rt.acquire();
    owns_slot_ = true;
  }
  const long long offsets[2] = {0, (long long) n_points};
  C2G_CHECK(c2g_ingest(rt.ctx, src, offsets, 1, 0, slot_, &int_id_));
  C2G_CHECK(c2g_get_heads(rt.ctx, slot_, 1, &head_));
  adoptHead();
}

void ContourManager::adoptHead() {
  if (head_.status != 0) {
    std::fprintf(stderr, "CHECK failed: scan %d exceeded a descriptor capacity (status %d)\n", int_id_, head_.status);
    std::abort();
  }
  std::vector<float>().swap(pts_);
  ext_pts_ = nullptr;
  views_loaded_ = false;
  for (size_t ll = 0; ll < cfg_.lv_grads_.size(); ++ll) {
    layer_keys_[ll].clear();
    layer_key_bcis_[ll].clear();
    for (int seq = 0; seq < cfg_.piv_firsts_; ++seq) {
      RetrievalKey k;
      for (int d = 0; d < RET_KEY_DIM; ++d) k[d] = head_.keys[ll][seq][d];
      layer_keys_[ll].push_back(k);
      const c2g_bci &b = head_.bcis[ll][seq];
      BCI bci(b.piv_seq, b.level);
      for (int w = 0; w < C2G_NUM_BIN_LAYERS; ++w)
        for (int bit = 0; bit < 64; ++bit)
          if ((b.dist_bin[w] >> bit) & 1ull) bci.dist_bin_.set(w * 64 + bit, true);
      for (int i = 0; i < b.n_nei; ++i)
        bci.nei_pts_.emplace_back(b.nei[i].level, b.nei[i].seq, b.nei[i].bit_pos, b.nei[i].r, b.nei[i].theta);
      for (int i = 0; i < b.n_seg; ++i) bci.nei_idx_segs_.push_back(b.seg[i]);
      layer_key_bcis_[ll].push_back(bci);
    }
  }
}

void ContourManager::loadViews() const {
  if (views_loaded_ || slot_ < 0) return;
  std::vector<c2g_view> raw(C2G_VIEW_CAP);
  C2G_CHECK(c2g_get_views(runtime().ctx, slot_, raw.data()));
  for (size_t ll = 0; ll < cfg_.lv_grads_.size(); ++ll) {
    cont_views_[ll].clear();
    for (int i = 0; i < head_.n_views[ll]; ++i)
      cont_views_[ll].push_back(std::make_shared<ContourView>(raw[head_.view_off[ll] + i]));
  }
  views_loaded_ = true;
}

std::vector<float> ContourManager::getBevImage() const {
  // only valid right after makeContoursRecurs of this scan (the dense BEV lives in per-batch scratch memory)
  std::vector<float> bev((size_t) cfg_.n_row_ * cfg_.n_col_);
  C2G_CHECK(c2g_get_bev(runtime().ctx, 0, bev.data(), nullptr, nullptr));
  return bev;
}

// ---------------------------------------------------------------------------------------------------------------------
// Host restatements of the reference's public statics (the query path runs the same cascade on the device, csrc/query.cu;
// tests/test_facade_gpu.py checks these against the device's per-hint records).
namespace {
inline float clampAngF(float ang) {  // clampAng<float> (include/tools/algos.h:49-51): double arithmetic, stored to float
  return (float) ((double) ang - std::floor(((double) ang + M_PI) / (2 * M_PI)) * 2 * M_PI);
}
inline void normalize2(float x, float y, float &ox, float &oy) {  // Eigen's normalized(): unchanged when the norm is zero
  const float n2 = x * x + y * y;
  if (n2 > 0.0f) {
    const float n = std::sqrt(n2);
    ox = x / n;
    oy = y / n;
  } else {
    ox = x;
    oy = y;
  }
}
}  // namespace

ScoreConstellSim BCI::checkConstellSim(const BCI &src, const BCI &tgt, const ScoreConstellSim &lb, std::vector<ConstellationPair> &constell_res) {
  const auto and1 = src.dist_bin_ & tgt.dist_bin_, and2 = (src.dist_bin_ << 1) & tgt.dist_bin_, and3 = (src.dist_bin_ >> 1) & tgt.dist_bin_;
  const int ovlp1 = (int) and1.count(), ovlp2 = (int) and2.count(), ovlp3 = (int) and3.count();
  ScoreConstellSim ret;
  ret.i_ovlp_sum = ovlp1 + ovlp2 + ovlp3;
  ret.i_ovlp_max_one = std::max(ovlp1, std::max(ovlp2, ovlp3));
  if (!(ret.i_ovlp_sum >= lb.i_ovlp_sum && ret.i_ovlp_max_one >= lb.i_ovlp_max_one)) return ret;
  // neighbours of the two anchors whose distance bits differ by at most one: two-pointer walk over the bit-position segments
  std::vector<DistSimPair> pot;
  const int n_s = (int) src.nei_idx_segs_.size(), n_t = (int) tgt.nei_idx_segs_.size();
  int p11 = 0;
  for (int p2 = 0; p2 < n_t - 1; p2++) {
    const int tb = tgt.nei_pts_[tgt.nei_idx_segs_[p2]].bit_pos;
    while (p11 < n_s - 1 && src.nei_pts_[src.nei_idx_segs_[p11]].bit_pos < tb - 1) p11++;
    int p12 = p11;
    while (p12 < n_s - 1 && src.nei_pts_[src.nei_idx_segs_[p12]].bit_pos <= tb + 1) p12++;
    for (int i = tgt.nei_idx_segs_[p2]; i < tgt.nei_idx_segs_[p2 + 1]; i++)
      for (int j = src.nei_idx_segs_[p11]; j < src.nei_idx_segs_[p12]; j++)
        pot.emplace_back(src.nei_pts_[j].level, src.nei_pts_[j].seq, tgt.nei_pts_[i].seq, clampAngF(tgt.nei_pts_[i].theta - src.nei_pts_[j].theta));
  }
  std::sort(pot.begin(), pot.end(), [](const DistSimPair &a, const DistSimPair &b) { return a.orie_diff < b.orie_diff; });
  // longest run of orientation differences inside a circular window of pi / 16
  const float angular_range = M_PI / 16;
  int beg = 0, longest = 1, p1 = 0, p2 = 0;
  const int sz = (int) pot.size();
  while (p1 < sz) {
    if (pot[p2 % sz].orie_diff - pot[p1].orie_diff + 2 * M_PI * int(p2 / sz) > angular_range)
      p1++;
    else {
      if (p2 - p1 + 1 > longest) {
        longest = p2 - p1 + 1;
        beg = p1;
      }
      p2++;
    }
  }
  ret.i_in_ang_rng = longest;
  if (longest < lb.i_in_ang_rng) return ret;
  constell_res.clear();
  for (int i = beg; i < longest + beg; i++) constell_res.emplace_back(pot[i % sz].level, pot[i % sz].seq_src, pot[i % sz].seq_tgt);
  constell_res.emplace_back(src.level_, src.piv_seq_, tgt.piv_seq_);  // the anchors are a pair too
  return ret;
}

ScorePairwiseSim ContourManager::checkConstellCorrespSim(const ContourManager &src, const ContourManager &tgt, const std::vector<ConstellationPair> &cstl_in,
                                                         const ScorePairwiseSim &lb, const ContourSimThresConfig &cont_sim,
                                                         std::vector<ConstellationPair> &cstl_out, std::vector<float> &area_perc) {
  ScorePairwiseSim ret;
  cstl_out.clear();
  area_perc.clear();
  for (const auto &pr : cstl_in)  // 1. individual similarity
    if (checkContPairSim(src, tgt, pr, cont_sim)) cstl_out.push_back(pr);
  ret.i_indiv_sim = (int) cstl_out.size();
  if (ret.i_indiv_sim < lb.i_indiv_sim) return ret;
  auto sv = [&](const ConstellationPair &p) -> const ContourView & { return *src.getLevContours(p.level)[p.seq_src]; };
  auto tv = [&](const ConstellationPair &p) -> const ContourView & { return *tgt.getLevContours(p.level)[p.seq_tgt]; };
  // 2.1 the "shaft": the comparison is made against the ALREADY NORMALISED previous shaft (contour_mng.h:1178-1179), so the last
  // qualifying pair among the first <= 10 wins, not the longest (SURVEY.md 8a' #6)
  float ssx = 0.f, ssy = 0.f, stx = 0.f, sty = 0.f;
  const int num_sim0 = (int) cstl_out.size();
  for (int i = 1; i < std::min(num_sim0, 10); i++)
    for (int j = 0; j < i; j++) {
      const float cx = sv(cstl_out[i]).pos_mean_(0) - sv(cstl_out[j]).pos_mean_(0), cy = sv(cstl_out[i]).pos_mean_(1) - sv(cstl_out[j]).pos_mean_(1);
      if (std::sqrt(cx * cx + cy * cy) > std::sqrt(ssx * ssx + ssy * ssy)) {
        normalize2(cx, cy, ssx, ssy);
        normalize2(tv(cstl_out[i]).pos_mean_(0) - tv(cstl_out[j]).pos_mean_(0), tv(cstl_out[i]).pos_mean_(1) - tv(cstl_out[j]).pos_mean_(1), stx, sty);
      }
    }
  // 2.2 drop pairs of two elongated contours whose major axes disagree with the shaft by more than pi / 6 both ways
  int num_sim = num_sim0;
  for (int i = 0; i < num_sim;) {
    const ContourView &sc1 = sv(cstl_out[i]), &tc1 = tv(cstl_out[i]);
    if (sc1.ecc_feat_ && tc1.ecc_feat_) {
      const float theta_s = std::acos(ssx * sc1.eig_vecs_(0, 1) + ssy * sc1.eig_vecs_(1, 1));
      const float theta_t = std::acos(stx * tc1.eig_vecs_(0, 1) + sty * tc1.eig_vecs_(1, 1));
      const float pi6 = M_PI / 6;
      if (std::fabs(theta_s - theta_t) > pi6 && std::fabs((float) (M_PI - theta_s) - theta_t) > pi6) {
        std::swap(cstl_out[i], cstl_out[num_sim - 1]);
        num_sim--;
        continue;
      }
    }
    i++;
  }
  cstl_out.erase(cstl_out.begin() + num_sim, cstl_out.end());
  ret.i_orie_sim = num_sim;
  if (ret.i_orie_sim < lb.i_orie_sim) return ret;
  for (const auto &pr : cstl_out) area_perc.push_back(0.5f * (src.getAreaPerc(pr.level, pr.seq_src) + tgt.getAreaPerc(pr.level, pr.seq_tgt)));
  return ret;
}

// Eigen::umeyama(src, tgt, no scaling) for 2-D point sets, closed form in double (as in csrc/query.cu)
Eigen::Isometry2d ContourManager::tfFromConstell(const ContourManager &src, const ContourManager &tgt, const ConstellationPair *cstl, int n) {
  if (n <= 2) {  // CHECK_GT(num_elem, 2)
    std::fprintf(stderr, "CHECK failed: getTFFromConstell needs more than 2 pairs\n");
    std::abort();
  }
  const double inv_n = 1.0 / (double) n;
  double sm0 = 0, sm1 = 0, dm0 = 0, dm1 = 0;
  for (int i = 0; i < n; ++i) {
    const ContourView &ps = *src.getLevContours(cstl[i].level)[cstl[i].seq_src], &pt = *tgt.getLevContours(cstl[i].level)[cstl[i].seq_tgt];
    sm0 += (double) ps.pos_mean_(0);
    sm1 += (double) ps.pos_mean_(1);
    dm0 += (double) pt.pos_mean_(0);
    dm1 += (double) pt.pos_mean_(1);
  }
  sm0 *= inv_n, sm1 *= inv_n, dm0 *= inv_n, dm1 *= inv_n;
  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
  for (int i = 0; i < n; ++i) {
    const ContourView &ps = *src.getLevContours(cstl[i].level)[cstl[i].seq_src], &pt = *tgt.getLevContours(cstl[i].level)[cstl[i].seq_tgt];
    const double sx = (double) ps.pos_mean_(0) - sm0, sy = (double) ps.pos_mean_(1) - sm1;
    const double dx = (double) pt.pos_mean_(0) - dm0, dy = (double) pt.pos_mean_(1) - dm1;
    s00 += dx * sx, s01 += dx * sy, s10 += dy * sx, s11 += dy * sy;
  }
  const double ang0 = std::atan2(s10 * inv_n - s01 * inv_n, s00 * inv_n + s11 * inv_n);
  const double c0 = std::cos(ang0), s0 = std::sin(ang0);
  Eigen::Isometry2d ret;
  ret.setIdentity();
  ret.rotate(std::atan2(s0, c0));
  ret.pretranslate(V2D(dm0 - (c0 * sm0 - s0 * sm1), dm1 - (s0 * sm0 + c0 * sm1)));
  return ret;
}

// ---------------------------------------------------------------------------------------------------------------------
ContourDB::ContourDB(const ContourDBConfig &config) : cfg_(config) {
  if (cfg_.q_levels_.empty()) {  // CHECK(!cfg_.q_levels_.empty()) (contour_db.h:683)
    std::fprintf(stderr, "CHECK failed: q_levels_ empty\n");
    std::abort();
  }
  runtime().setDb(config);
}

void ContourDB::addScan(const std::shared_ptr<ContourManager> &added, double curr_timestamp) {
  auto &rt = runtime();
  rt.ensure();
  const int gidx = (int) all_bevs_.size();
  if (gidx >= rt.capacity - rt.n_staging) {
    std::fprintf(stderr, "CHECK failed: database capacity (C2G_SCAN_CAPACITY) exhausted\n");
    std::abort();
  }
  // the descriptor moves from its staging slot to slot == gidx (IndexOfKey::gidx, contour_db.h:819)
  C2G_CHECK(c2g_copy_slots(rt.ctx, added->slot_, gidx, 1));
  if (added->owns_slot_) rt.release(added->slot_);
  added->slot_ = gidx;
  added->owns_slot_ = false;
  C2G_CHECK(c2g_db_add_scans(rt.ctx, gidx, 1, &curr_timestamp));
  all_bevs_.emplace_back(added);
}

static c2g_score_ensemble packEnsemble(const CandidateScoreEnsemble &e) {
  c2g_score_ensemble s;
  s.i_ovlp_sum = e.sim_constell.i_ovlp_sum;
  s.i_ovlp_max_one = e.sim_constell.i_ovlp_max_one;
  s.i_in_ang_rng = e.sim_constell.i_in_ang_rng;
  s.i_indiv_sim = e.sim_pair.i_indiv_sim;
  s.i_orie_sim = e.sim_pair.i_orie_sim;
  s.correlation = e.sim_post.correlation;
  s.area_perc = e.sim_post.area_perc;
  s.neg_est_dist = e.sim_post.neg_est_dist;
  return s;
}

// anch_props_[0].correlation_ / T_delta_ of the returned candidate after fineOptimize (contour_db.h:626-627,642-643)
static bool unpackBest(const c2g_query_result &res, int &gidx, double &corr, Eigen::Isometry2d &T) {
  if (res.overflow) C2G_CHECK(C2G_ERR_CAPACITY);  // more candidate poses than C2G_MAX_CAND: the reference keeps them all
  if (!(res.n_cand > 0 && res.best >= 0)) return false;  // ret_size = 1 (contour_db.h:639)
  const c2g_cand &c = res.cand[res.best];
  if (c.fine_flags != 0) C2G_CHECK(C2G_ERR_CAPACITY);
  gidx = c.cand_gidx;
  corr = (double) c.corr_fine;
  T.setIdentity();
  T.rotate(std::atan2(c.T_fine[1], c.T_fine[0]));
  T.pretranslate(V2D(c.T_fine[2], c.T_fine[3]));
  return true;
}

void ContourDB::queryAddBalanceWindow(std::vector<std::shared_ptr<ContourManager>> &scans, const std::vector<double> &ts, const std::vector<int> &seeds,
                                      const CandidateScoreEnsemble &thres_lb, const CandidateScoreEnsemble &thres_ub,
                                      std::vector<WindowResult> &out) {
  auto &rt = runtime();
  rt.ensure();
  const int W = (int) scans.size();
  out.assign((size_t) W, WindowResult());
  if (W == 0) return;
  if (W > rt.max_window || (int) ts.size() != W || (int) seeds.size() != W) {
    std::fprintf(stderr, "CHECK failed: window of %d scans (C2G_WINDOW = %d), %zu timestamps, %zu seeds\n", W, rt.max_window, ts.size(), seeds.size());
    std::abort();
  }
  const int first = (int) all_bevs_.size();
  if (first + W > rt.capacity - rt.n_staging) {
    std::fprintf(stderr, "CHECK failed: database capacity (C2G_SCAN_CAPACITY) exhausted\n");
    std::abort();
  }
  // the points of the window as one buffer: used in place when the scans were read back to back into the pinned buffer
  // (pinnedScanBuffer + makeBEVFromBin), gathered into a staging vector otherwise
  std::vector<long long> offsets((size_t) W + 1, 0);
  std::vector<int> ids((size_t) W);
  bool contiguous = true;
  for (int i = 0; i < W; ++i) {
    ContourManager &cm = *scans[i];
    const size_t n = cm.ext_pts_ ? cm.ext_n_ : cm.pts_.size() / 4;
    if (n <= 10 || cm.slot_ >= 0) {
      std::fprintf(stderr, "CHECK failed: scan %d of the window has <= 10 points or was already processed\n", cm.int_id_);
      std::abort();
    }
    offsets[i + 1] = offsets[i] + (long long) n;
    ids[i] = cm.int_id_;
    contiguous = contiguous && cm.ext_pts_ && (i == 0 || cm.ext_pts_ == scans[i - 1]->ext_pts_ + 4 * scans[i - 1]->ext_n_);
  }
  const float *pts = scans[0]->ext_pts_;
  std::vector<float> gathered;
  if (!contiguous) {
    gathered.resize(4 * (size_t) offsets[W]);
    for (int i = 0; i < W; ++i) {
      const ContourManager &cm = *scans[i];
      std::memcpy(gathered.data() + 4 * offsets[i], cm.ext_pts_ ? cm.ext_pts_ : cm.pts_.data(), sizeof(float) * 4 * (size_t) (offsets[i + 1] - offsets[i]));
    }
    pts = gathered.data();
  }
  const c2g_score_ensemble lb = packEnsemble(thres_lb), ub = packEnsemble(thres_ub);
  std::vector<c2g_query_result> res((size_t) W);
  stp.start();
  C2G_CHECK(c2g_online_window(rt.ctx, pts, offsets.data(), W, 0, ids.data(), ts.data(), seeds.data(), &lb, &ub, res.data()));
  stp.record("window: make bev + query + update (GPU)");
  std::vector<c2g_scan_head> heads((size_t) W);
  C2G_CHECK(c2g_get_heads(rt.ctx, first, W, heads.data()));
  for (int i = 0; i < W; ++i) {
    ContourManager &cm = *scans[i];
    cm.slot_ = first + i;  // IndexOfKey::gidx (contour_db.h:819): the descriptor was ingested straight into its DB slot
    cm.owns_slot_ = false;
    cm.head_ = heads[i];
    cm.adoptHead();
    all_bevs_.emplace_back(scans[i]);
  }
  for (int i = 0; i < W; ++i) {  // candidates are scans added before scan i (possibly earlier scans of this window)
    int gidx = -1;
    out[i].found = unpackBest(res[i], gidx, out[i].corr, out[i].tf);
    if (out[i].found) out[i].cand = all_bevs_[gidx];
  }
}

void ContourDB::pushAndBalance(int seed, double curr_timestamp) { C2G_CHECK(c2g_db_push_and_balance(runtime().ctx, seed, curr_timestamp)); }

void ContourDB::queryRangedKNN(const std::shared_ptr<const ContourManager> &q_ptr, const CandidateScoreEnsemble &thres_lb,
                               const CandidateScoreEnsemble &thres_ub, std::vector<std::shared_ptr<const ContourManager>> &cand_ptrs,
                               std::vector<double> &cand_corr, std::vector<Eigen::Isometry2d> &cand_tf) const {
  cand_ptrs.clear();
  cand_corr.clear();
  cand_tf.clear();
  const c2g_score_ensemble lb = packEnsemble(thres_lb), ub = packEnsemble(thres_ub);
  c2g_query_result res;
  // the reference's three stage timers (contour_db.h:729-787): the device times of the kernels that replace each stage
  // (c2g_query_profile: knn | prefilter + score | proposal replay + GMM-L2 gate + output + refinement + ranking); the host
  // overhead of the call (launches, result read-back) goes to "L2 opt", the stage that ends the query in the reference too
  static bool prof_on = false;
  if (!prof_on) {
    C2G_CHECK(c2g_query_profile(runtime().ctx, 1, nullptr));
    prof_on = true;
  }
  TicToc wall;
  C2G_CHECK(c2g_query(runtime().ctx, q_ptr->deviceSlot(), 1, &lb, &ub, &res, nullptr, nullptr));
  const double t_wall = wall.toc();
  float ms[8];
  C2G_CHECK(c2g_query_profile(runtime().ctx, 1, ms));
  const double t_knn = 1e-3 * ms[0], t_constell = 1e-3 * (ms[1] + ms[2]);
  stp.recordSeconds("KNN search", t_knn);
  stp.recordSeconds("Constell", t_constell);
  stp.recordSeconds("L2 opt", t_wall - t_knn - t_constell > 0 ? t_wall - t_knn - t_constell : 0.0);
  int gidx = -1;
  double corr = 0.0;
  Eigen::Isometry2d T;
  if (unpackBest(res, gidx, corr, T)) {
    cand_ptrs.emplace_back(all_bevs_[gidx]);
    cand_corr.emplace_back(corr);
    cand_tf.emplace_back(T);
  }
}
