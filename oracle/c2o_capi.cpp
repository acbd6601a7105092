// ORACLE — TEST INFRASTRUCTURE ONLY (see c2o_ingest.hpp header).
// Flat C API over the CPU restatement so tests/ and bench.py's cpu_baseline leg can drive it through ctypes.
#include <chrono>

#include "c2o_query.hpp"

using namespace c2o;
typedef std::shared_ptr<Scan> ScanPtr;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void fill_head(const Scan &s, c2g_scan_head *h) {
  std::memset(h, 0, sizeof(*h));
  h->int_id = s.int_id;
  int off = 0;
  for (int l = 0; l < C2G_NLEV; ++l) {
    h->n_views[l] = (int) s.cont_views[l].size();
    h->view_off[l] = off;
    off += h->n_views[l];
    h->layer_cell_cnt[l] = s.layer_cell_cnt[l];
    for (int q = 0; q < s.cfg.piv_firsts && q < C2G_MAX_PIV; ++q) {
      for (int d = 0; d < C2G_KEY_DIM; ++d) h->keys[l][q][d] = s.layer_keys[l][q][d];
      h->bcis[l][q] = s.layer_key_bcis[l][q];
    }
  }
  if (off > C2G_VIEW_CAP) h->status |= 1;
  h->n_occupied = (int) s.bev_pixfs.size();
  GMMScanData g = buildGMMScan(s);
  for (int li = 0; li < 4; ++li) h->n_ell[li] = (int) g.ell[li].size();
  h->gmm_auto_corr = g.auto_corr;
}

extern "C" {

void *c2o_scan_create(const c2g_cm_config *cfg, int int_id) { return new ScanPtr(std::make_shared<Scan>(*cfg, int_id)); }
void c2o_scan_free(void *h) { delete (ScanPtr *) h; }
void c2o_scan_make_bev(void *h, const float *pts, int n) { (*(ScanPtr *) h)->makeBEV(pts, n); }
void c2o_scan_make_contours(void *h) { (*(ScanPtr *) h)->makeContoursRecurs(); }
void c2o_scan_ingest(void *h, const float *pts, int n) {
  (*(ScanPtr *) h)->makeBEV(pts, n);
  (*(ScanPtr *) h)->makeContoursRecurs();
}
// dense images: bev (init -1000), row_f / col_f (-1 where no pillar)
void c2o_scan_get_bev(void *h, float *bev, float *row_f, float *col_f) {
  const Scan &s = **(ScanPtr *) h;
  const size_t n = (size_t) s.cfg.n_row * s.cfg.n_col;
  for (size_t i = 0; i < n; ++i) {
    bev[i] = s.bev[i];
    row_f[i] = -1.0f;
    col_f[i] = -1.0f;
  }
  for (auto &p : s.bev_pixfs) {
    row_f[p.first] = p.second.row_f;
    col_f[p.first] = p.second.col_f;
  }
}
int c2o_scan_n_views(void *h, int level) { return (int) (*(ScanPtr *) h)->cont_views[level].size(); }
void c2o_scan_get_views(void *h, int level, int presort, c2g_view *out) {
  const Scan &s = **(ScanPtr *) h;
  const auto &v = presort ? s.presort_views[level] : s.cont_views[level];
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void c2o_scan_get_head(void *h, c2g_scan_head *out) { fill_head(**(ScanPtr *) h, out); }

// ---- stand-alone helpers exposed for unit tests -----------------------------------------------------------------
void c2o_eig2f(float a, float b, float c, float *evals, float *evecs) { selfAdjointEigen2f(a, b, c, evals, evecs); }
// CCL on an arbitrary mask (h x w uint8) -> labels int32 + stats (n+1) x 5 (left, top, width, height, area); returns n
int c2o_ccl8(const uint8_t *mask, int h, int w, int32_t *labels, int32_t *stats, int stats_cap) {
  std::vector<int> lab;
  std::vector<CCStat> st = connectedComponents8(mask, h, w, lab);
  for (size_t i = 0; i < lab.size(); ++i) labels[i] = lab[i];
  for (size_t i = 0; i < st.size() && (int) i < stats_cap; ++i) {
    stats[i * 5 + 0] = st[i].left;
    stats[i * 5 + 1] = st[i].top;
    stats[i * 5 + 2] = st[i].width;
    stats[i * 5 + 3] = st[i].height;
    stats[i * 5 + 4] = st[i].area;
  }
  return (int) st.size() - 1;
}
float c2o_gauss_pdf_f(float x, float mean, float sd) { return gaussPDF<float>(x, mean, sd); }
double c2o_exp(double x) { return std::exp(x); }
float c2o_atan2f(float y, float x) { return std::atan2(y, x); }
float c2o_acosf(float x) { return std::acos(x); }

// ---- database ----------------------------------------------------------------------------------------------------
void *c2o_db_create(const c2g_db_config *cfg) { return new ContourDB(*cfg); }
void c2o_db_free(void *db) { delete (ContourDB *) db; }
void c2o_db_add_scan(void *db, void *scan, double ts) { ((ContourDB *) db)->addScan(*(ScanPtr *) scan, ts); }
void c2o_db_push_and_balance(void *db, int seed, double ts) { ((ContourDB *) db)->pushAndBalance(seed, ts); }
int c2o_db_n_scans(void *db) { return (int) ((ContourDB *) db)->all_bevs_.size(); }
// rebalancing moves so far / buckets a literal reference would have searched through a stale index afterwards (all layers)
void c2o_db_rebalance_stats(void *db, long long *moves, long long *stale, long long *stale_donor) {
  *moves = *stale = *stale_donor = 0;
  for (const LayerDB &L : ((ContourDB *) db)->layer_db_) {
    *moves += L.n_moves_;
    *stale += L.n_stale_;
    *stale_donor += L.n_stale_donor_;
  }
}

// searchable prefix of every bucket of layer ll: 0 without an index, else the size the index was built over
void c2o_db_indexed(void *db, int ll, int32_t *indexed) {
  const LayerDB &l = ((ContourDB *) db)->layer_db_[ll];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) indexed[i] = l.buckets_[i].has_tree ? (int32_t) l.buckets_[i].indexed_size : 0;
}

// layer state for mirroring / parity: bucket_ranges[7], tree_sizes[6], buffer_sizes[6]
void c2o_db_layer_state(void *db, int ll, float *bucket_ranges, int32_t *tree_sizes, int32_t *buffer_sizes) {
  const LayerDB &l = ((ContourDB *) db)->layer_db_[ll];
  for (int i = 0; i <= C2G_NUM_BUCKETS; ++i) bucket_ranges[i] = l.bucket_ranges_[i];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) {
    tree_sizes[i] = (int) l.buckets_[i].data_tree.size();
    buffer_sizes[i] = (int) l.buckets_[i].buffer.size();
  }
}
// keys + (gidx, level, seq) of one bucket's tree, in tree order
void c2o_db_bucket_tree(void *db, int ll, int bucket, float *keys, int32_t *gidx, int32_t *seq) {
  const TreeBucket &b = ((ContourDB *) db)->layer_db_[ll].buckets_[bucket];
  for (size_t i = 0; i < b.data_tree.size(); ++i) {
    for (int d = 0; d < C2G_KEY_DIM; ++d) keys[i * C2G_KEY_DIM + d] = b.data_tree[i][d];
    gidx[i] = (int32_t) b.gkidx_tree[i].gidx;
    seq[i] = b.gkidx_tree[i].seq;
  }
}

// query one scan; optionally returns the full hint trace (hints in reference order + cascade result per hint)
int c2o_db_query(void *db, void *scan, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub, c2g_query_result *out,
                 c2g_hint *hints, c2g_pair_score *scores, int cap) {
  std::vector<HintTrace> trace;
  ((ContourDB *) db)->queryRangedKNN(*(ScanPtr *) scan, *lb, *ub, *out, &trace);
  int n = (int) trace.size();
  for (int i = 0; i < n && i < cap; ++i) {
    if (hints) hints[i] = trace[i].hint;
    if (scores) scores[i] = trace[i].score;
  }
  return n;
}

// raw kNN of one key against one layer (for kNN parity tests): returns count
int c2o_db_layer_knn(void *db, int ll, const float *key, int k, float max_dist_sq, int32_t *gidx, int32_t *seq, float *dist) {
  Key q;
  for (int d = 0; d < C2G_KEY_DIM; ++d) q[d] = key[d];
  std::vector<std::pair<IndexOfKey, float>> res;
  ((ContourDB *) db)->layer_db_[ll].layerKNNSearch(q, k, max_dist_sq, res);
  for (size_t i = 0; i < res.size(); ++i) {
    gidx[i] = (int32_t) res[i].first.gidx;
    seq[i] = res[i].first.seq;
    dist[i] = res[i].second;
  }
  return (int) res.size();
}

// ---- CPU baseline driver: for each of B scans {ingest; query; (optional) addScan + pushAndBalance}, single thread.
// pts: concatenated scans, offsets[b]..offsets[b+1] in points. t_stage[5]: make bev / KNN search / Constell / L2 opt /
// Update database (seconds, accumulated) — the reference's SequentialTimeProfiler stage names.
void c2o_run_loop(void *db_, const c2g_cm_config *cfg, const float *pts, const int64_t *offsets, int B, int first_id,
                  const double *ts, int do_query, int do_add, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub,
                  c2g_query_result *results, double *t_stage) {
  ContourDB *db = (ContourDB *) db_;
  for (int b = 0; b < B; ++b) {
    double t0 = now_s();
    ScanPtr s = std::make_shared<Scan>(*cfg, first_id + b);
    s->makeBEV(pts + 4 * offsets[b], (int) (offsets[b + 1] - offsets[b]));
    s->makeContoursRecurs();
    std::vector<float>().swap(s->bev);  // clearImage() (test/batch_bin_test.cpp:169)
    t_stage[0] += now_s() - t0;
    if (do_query) {
      c2g_query_result r;
      db->queryRangedKNN(s, *lb, *ub, r, nullptr, &t_stage[1], &t_stage[2], &t_stage[3]);
      if (results) results[b] = r;
    }
    if (do_add) {
      double t1 = now_s();
      db->addScan(s, ts[b]);
      db->pushAndBalance(first_id + b, ts[b]);
      t_stage[4] += now_s() - t1;
    }
  }
}

int c2o_sizeof_scan_head() { return (int) sizeof(c2g_scan_head); }
int c2o_sizeof_query_result() { return (int) sizeof(c2g_query_result); }

}  // extern "C"

// real libstdc++ std::sort on packed (key << 16 | index) words — the reference behaviour tests compare the device-side
// replay (contour_context_b200/csrc/stdsort.cuh) against.
extern "C" void c2o_std_sort_words(uint32_t *words, int n, int desc) {
  if (desc)
    std::sort(words, words + n, [](uint32_t a, uint32_t b) { return (a >> 16) > (b >> 16); });
  else
    std::sort(words, words + n, [](uint32_t a, uint32_t b) { return (a >> 16) < (b >> 16); });
}

// test hook: put n keys straight into bucket 0's tree of q-level ll (gidx = index, seq = 0)
extern "C" void c2o_test_fill_layer(void *db, int ll, const float *keys, int n) {
  TreeBucket &b = ((ContourDB *) db)->layer_db_[ll].buckets_[0];
  for (int i = 0; i < n; ++i) {
    Key k;
    for (int d = 0; d < C2G_KEY_DIM; ++d) k[d] = keys[i * C2G_KEY_DIM + d];
    b.data_tree.push_back(k);
    b.gkidx_tree.push_back(IndexOfKey{(size_t) i, 0, 0});
  }
  b.rebuildTree();
}

// test hook: LayerDB::pushBuffer with a raw key (as ContourDB::addScan would do for one key)
extern "C" void c2o_test_push_key(void *db, int ll, const float *key, double ts, int gidx, int seq) {
  Key k;
  for (int d = 0; d < C2G_KEY_DIM; ++d) k[d] = key[d];
  ContourDB *D = (ContourDB *) db;
  D->layer_db_[ll].pushBuffer(k, ts, IndexOfKey{(size_t) gidx, D->cfg_.q_levels[ll], seq});
}

// Build a Scan from a finished descriptor (head + sorted views) instead of from points: lets bench.py hand the CPU
// baseline the same 5 000-scan database the GPU path built (descriptor parity is what tests/ establish) without spending
// 40 s of single-thread CPU ingest before every baseline run.  The BEV image is not restored (the reference drops it
// after ingest too, test/batch_bin_test.cpp:169).
extern "C" void *c2o_scan_from_descriptor(const c2g_cm_config *cfg, const c2g_scan_head *head, const c2g_view *views) {
  ScanPtr *h = new ScanPtr(std::make_shared<Scan>(*cfg, head->int_id));
  Scan &s = **h;
  std::vector<float>().swap(s.bev);
  for (int l = 0; l < C2G_NLEV; ++l) {
    s.cont_views[l].assign(views + head->view_off[l], views + head->view_off[l] + head->n_views[l]);
    s.layer_cell_cnt[l] = head->layer_cell_cnt[l];
    s.cont_perc[l].clear();
    for (auto &v : s.cont_views[l]) s.cont_perc[l].push_back(v.cell_cnt * 1.0f / s.layer_cell_cnt[l]);
    s.layer_keys[l].clear();
    s.layer_key_bcis[l].clear();
    for (int q = 0; q < cfg->piv_firsts; ++q) {
      std::array<float, C2G_KEY_DIM> k;
      for (int d = 0; d < C2G_KEY_DIM; ++d) k[d] = head->keys[l][q][d];
      s.layer_keys[l].push_back(k);
      s.layer_key_bcis[l].push_back(head->bcis[l][q]);
    }
  }
  return h;
}
extern "C" int c2o_uses_nanoflann() {
#ifdef C2O_USE_NANOFLANN
  return 1;
#else
  return 0;
#endif
}

// vector versions of the host libm calls the reference makes (std::exp(double), std::atan2(float,float), std::acos(float),
// atanf) — the ground truth tests/test_libm.py compares the product's restatements with
extern "C" void c2o_vec_libm(int kind, int n, const void *in, void *out) {
  for (int i = 0; i < n; ++i) {
    switch (kind) {
      case 0:
      case 1: ((double *) out)[i] = std::exp(((const double *) in)[i]); break;
      case 2: ((float *) out)[i] = std::atan2(((const float *) in)[2 * i], ((const float *) in)[2 * i + 1]); break;
      case 3: ((float *) out)[i] = std::acos(((const float *) in)[i]); break;
      case 4: ((float *) out)[i] = std::atan(((const float *) in)[i]); break;
    }
  }
}

// ---- refinement hooks (c2o_refine.hpp) ------------------------------------------------------------------------------
// src/tgt are Scan handles; T = (cos, sin, tx, ty).  mode 0: evaluate the Jet cost at params p (pairs selected at T):
// out = {cost, d/dx, d/dy, d/dtheta, n_pairs}.  mode 1: calcCorrelation from T: out = {correlation, x, y, theta,
// initial_cost, final_cost, iterations, termination, n_eval, n_pairs}.  mode 2: double-precision (non-Jet) normalised
// correlation at p with pairs selected at T (tryProblem).
extern "C" void c2o_refine_hook(int mode, void *src_h, void *tgt_h, const double *T, const double *p, double *out) {
  const Scan &src = **(ScanPtr *) src_h, &tgt = **(ScanPtr *) tgt_h;
  const GMMScanData gs = buildGMMScan(src), gt = buildGMMScan(tgt);
  Iso2 Ti;
  Ti.m00 = T[0];
  Ti.m10 = T[1];
  Ti.m01 = -T[1];
  Ti.m11 = T[0];
  Ti.tx = T[2];
  Ti.ty = T[3];
  if (mode == 0) {
    refine::Problem P = refine::makeProblem(gs, gt, Ti);
    const refine::J3 f = refine::evalJet(P, p);
    out[0] = f.a;
    out[1] = f.v[0];
    out[2] = f.v[1];
    out[3] = f.v[2];
    out[4] = (double) P.pairs.size();
  } else if (mode == 1) {
    const refine::CorrResult r = refine::calcCorrelation(gs, gt, Ti);
    refine::Problem P = refine::makeProblem(gs, gt, Ti);
    out[0] = r.correlation;
    out[1] = r.opt.x[0];
    out[2] = r.opt.x[1];
    out[3] = r.opt.x[2];
    out[4] = r.opt.initial_cost;
    out[5] = r.opt.final_cost;
    out[6] = r.opt.iterations;
    out[7] = r.opt.termination;
    out[8] = r.opt.n_eval;
    out[9] = (double) P.pairs.size();
    out[10] = std::sqrt(gs.auto_corr * gt.auto_corr);
  }
}

// test hook: fill the trees of q-level ll with arbitrary (key, gidx, seq) entries in the given buckets and set the bucket
// ranges (large synthetic key tables for the kNN parity test)
extern "C" void c2o_test_fill_layer2(void *db, int ll, const float *keys, const int32_t *gidx, const int8_t *seq, const uint8_t *bucket,
                                     int n, const float *ranges) {
  LayerDB &L = ((ContourDB *) db)->layer_db_[ll];
  for (int b = 0; b < LayerDB::max_num_backets_; ++b) {
    L.buckets_[b].data_tree.clear();
    L.buckets_[b].gkidx_tree.clear();
  }
  for (int i = 0; i < n; ++i) {
    Key k;
    for (int d = 0; d < C2G_KEY_DIM; ++d) k[d] = keys[i * C2G_KEY_DIM + d];
    TreeBucket &b = L.buckets_[bucket[i]];
    b.data_tree.push_back(k);
    b.gkidx_tree.push_back(IndexOfKey{(size_t) gidx[i], ((ContourDB *) db)->cfg_.q_levels[ll], (int) seq[i]});
  }
  for (int b = 0; b <= LayerDB::max_num_backets_; ++b) L.bucket_ranges_[b] = ranges[b];
  for (int b = 0; b < LayerDB::max_num_backets_; ++b) L.buckets_[b].rebuildTree();
}
