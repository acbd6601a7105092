// ORACLE — TEST INFRASTRUCTURE ONLY (see c2o_ingest.hpp header).
//
// c2o_query.hpp: CPU restatement of the database / query half of the reference's cont2contops path:
//   TreeBucket / LayerDB / ContourDB        include/cont2/contour_db.h:54-217,658-845, src/cont2/contour_db.cpp:63-403
//   BCI::checkConstellSim                   include/cont2/contour_mng.h:288-388
//   ContourManager::checkConstellCorrespSim include/cont2/contour_mng.h:1124-1242
//   ContourManager::getTFFromConstell       include/cont2/contour_mng.h:1251-1277   (Eigen::umeyama -> closed form, see below)
//   CandidateManager                        include/cont2/contour_db.h:264-656
//   GMMPair / ConstellCorrelation           include/cont2/correlation.h:15-202,287-296
//
// Deviations that are documented rather than hidden ("parity unpinned" items):
//  * nanoflann KD-tree search is replaced by an exhaustive scan that evaluates the SAME metric in the SAME summation
//    order (nanoflann.hpp:428-462) and the same result-set semantics (nanoflann.hpp:194-227, contour_db.h:32-52);
//    only the order of exactly-equal distances can differ (tree traversal order vs index order). oracle/_ref builds the
//    reference's vendored nanoflann to cross-check this.
//  * Eigen::umeyama (JacobiSVD inside) is replaced by the closed-form 2-D Kabsch/Umeyama rotation; agreement is at
//    double round-off level, not bit level.
//  * Ceres (correlation.h:206-238) is un-vendored and absent: the L-BFGS refinement inside fineOptimize() is the
//    restatement of Ceres' default line-search minimizer in c2o_refine.hpp (parity unpinned, see its header).
#pragma once

#include <array>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <numeric>

#include "c2o_ingest.hpp"

#ifdef C2O_USE_NANOFLANN
// Built only by `make ref` (oracle/_ref/liboracle_nf.so) when the reference tree is present: the reference's own
// vendored KD-tree (thirdparty/nanoflann.hpp + KDTreeVectorOfVectorsAdaptor.h, included from /root/reference, never
// copied) replaces the exhaustive scan in TreeBucket::knnSearch, so that the timed CPU baseline searches exactly like the
// reference does (leaf size 10, metric_L2, MyKNNResSet, SearchParams(10); contour_db.h:32-52,109-117, contour_db.cpp:381-403).
#include <cstdint>
#include <nanoflann.hpp>
#include <KDTreeVectorOfVectorsAdaptor.h>
#endif

namespace c2o {

using Key = std::array<float, C2G_KEY_DIM>;

#ifdef C2O_USE_NANOFLANN
typedef std::vector<Key> my_vector_of_vectors_t;
typedef KDTreeVectorOfVectorsAdaptor<my_vector_of_vectors_t, float> my_kd_tree_t;
template <typename DD, typename II = size_t, typename CC = size_t>
class MyKNNResSet : public nanoflann::KNNResultSet<DD, II, CC> {
 public:
  explicit MyKNNResSet(CC capacity_) : nanoflann::KNNResultSet<DD, II, CC>(capacity_) {}
  void init(II *indices_, DD *dists_, DD max_dist_metric) {
    this->indices = indices_;
    this->dists = dists_;
    this->count = 0;
    if (this->capacity) this->dists[this->capacity - 1] = max_dist_metric;
  }
};
#endif

inline float keySum(const Key &k) {  // ArrayAsKey::sum (contour_mng.h:74-79)
  float ret(0);
  for (const auto &dat : k) ret += dat;
  return ret;
}

// Eigen::Isometry2d as used by the reference (linear 2x2 + translation)
struct Iso2 {
  double m00 = 1, m01 = 0, m10 = 0, m11 = 1, tx = 0, ty = 0;
  static Iso2 fromAngTrans(double ang, double x, double y) {  // setIdentity(); rotate(ang); pretranslate(t)
    Iso2 r;
    double c = std::cos(ang), s = std::sin(ang);
    r.m00 = c;
    r.m01 = -s;
    r.m10 = s;
    r.m11 = c;
    r.tx = x;
    r.ty = y;
    return r;
  }
  Iso2 inverse() const {  // Transform::inverse(Isometry): R^T, -R^T t
    Iso2 r;
    r.m00 = m00;
    r.m01 = m10;
    r.m10 = m01;
    r.m11 = m11;
    r.tx = -(r.m00 * tx + r.m01 * ty);
    r.ty = -(r.m10 * tx + r.m11 * ty);
    return r;
  }
  Iso2 operator*(const Iso2 &b) const {
    Iso2 r;
    r.m00 = m00 * b.m00 + m01 * b.m10;
    r.m01 = m00 * b.m01 + m01 * b.m11;
    r.m10 = m10 * b.m00 + m11 * b.m10;
    r.m11 = m10 * b.m01 + m11 * b.m11;
    r.tx = (m00 * b.tx + m01 * b.ty) + tx;
    r.ty = (m10 * b.tx + m11 * b.ty) + ty;
    return r;
  }
  void apply(double x, double y, double &ox, double &oy) const {
    ox = (m00 * x + m01 * y) + tx;
    oy = (m10 * x + m11 * y) + ty;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// ConstellationPair (contour_mng.h:221-240)
// ---------------------------------------------------------------------------------------------------------------
struct CPair {
  int8_t level, seq_src, seq_tgt;
  bool operator<(const CPair &a) const {
    return level < a.level || (level == a.level && seq_src < a.seq_src) ||
           (level == a.level && seq_src == a.seq_src && seq_tgt < a.seq_tgt);
  }
};

// BCI::checkConstellSim (contour_mng.h:288-388). ret = {i_ovlp_sum, i_ovlp_max_one, i_in_ang_rng}
inline void checkConstellSim(const c2g_bci &src, const c2g_bci &tgt, const int lb[3], int ret[3],
                             std::vector<CPair> &constell_res) {
  ret[0] = ret[1] = ret[2] = 0;
  // 256-bit ops on 4 words; bit 63 of each 64-bit layer is never set (contour_mng.h:856-861), but do the real
  // 256-bit shift anyway.
  uint64_t s[4], t[4], sl[4], sr[4];
  for (int i = 0; i < 4; ++i) {
    s[i] = src.dist_bin[i];
    t[i] = tgt.dist_bin[i];
  }
  for (int i = 0; i < 4; ++i) {
    sl[i] = (s[i] << 1) | (i > 0 ? (s[i - 1] >> 63) : 0);
    sr[i] = (s[i] >> 1) | (i < 3 ? (s[i + 1] << 63) : 0);
  }
  int ovlp1 = 0, ovlp2 = 0, ovlp3 = 0;
  for (int i = 0; i < 4; ++i) {
    ovlp1 += __builtin_popcountll(s[i] & t[i]);
    ovlp2 += __builtin_popcountll(sl[i] & t[i]);
    ovlp3 += __builtin_popcountll(sr[i] & t[i]);
  }
  int ovlp_sum = ovlp1 + ovlp2 + ovlp3;
  int max_one = std::max(ovlp1, std::max(ovlp2, ovlp3));
  ret[0] = ovlp_sum;
  ret[1] = max_one;
  if (!(ovlp_sum >= lb[0] && max_one >= lb[1])) return;

  struct DistSimPair {
    float orie_diff;
    int8_t seq_src, seq_tgt, level;
  };
  std::vector<DistSimPair> potential_pairs;
  const int n_sseg = src.n_seg, n_tseg = tgt.n_seg;
  int16_t p11 = 0, p12;
  for (int16_t p2 = 0; p2 < n_tseg - 1; p2++) {
    while (p11 < n_sseg - 1 && src.nei[src.seg[p11]].bit_pos < tgt.nei[tgt.seg[p2]].bit_pos - 1) p11++;
    p12 = p11;
    while (p12 < n_sseg - 1 && src.nei[src.seg[p12]].bit_pos <= tgt.nei[tgt.seg[p2]].bit_pos + 1) p12++;
    for (int i = tgt.seg[p2]; i < tgt.seg[p2 + 1]; i++)
      for (int j = src.seg[p11]; j < src.seg[p12]; j++) {
        const c2g_relpt &rp1 = src.nei[j], &rp2 = tgt.nei[i];
        potential_pairs.push_back(DistSimPair{rp2.theta - rp1.theta, rp1.seq, rp2.seq, rp1.level});
      }
  }
  for (auto &x : potential_pairs) clampAng<float>(x.orie_diff);
  std::sort(potential_pairs.begin(), potential_pairs.end(),
            [](const DistSimPair &a, const DistSimPair &b) { return a.orie_diff < b.orie_diff; });

  const float angular_range = M_PI / 16;
  int longest_in_range_beg = 0, longest_in_range = 1, pot_sz = (int) potential_pairs.size(), p1 = 0, p2 = 0;
  while (p1 < pot_sz) {
    if (potential_pairs[p2 % pot_sz].orie_diff - potential_pairs[p1].orie_diff + 2 * M_PI * int(p2 / pot_sz) > angular_range)
      p1++;
    else {
      if (p2 - p1 + 1 > longest_in_range) {
        longest_in_range = p2 - p1 + 1;
        longest_in_range_beg = p1;
      }
      p2++;
    }
  }
  ret[2] = longest_in_range;
  if (longest_in_range < lb[2]) return;
  constell_res.clear();
  for (int i = longest_in_range_beg; i < longest_in_range + longest_in_range_beg; i++)
    constell_res.push_back(CPair{potential_pairs[i % pot_sz].level, potential_pairs[i % pot_sz].seq_src,
                                 potential_pairs[i % pot_sz].seq_tgt});
  constell_res.push_back(CPair{src.level, src.piv_seq, tgt.piv_seq});
}

// ContourManager::checkContPairSim (contour_mng.h:1279-1284)
inline bool checkContPairSim(const Scan &src, const Scan &tgt, const CPair &c, const c2g_sim_config &sim) {
  return checkSim(src.cont_views[c.level][c.seq_src], tgt.cont_views[c.level][c.seq_tgt], sim);
}

// ContourManager::checkConstellCorrespSim (contour_mng.h:1124-1242). ret = {i_indiv_sim, i_orie_sim}
inline void checkConstellCorrespSim(const Scan &src, const Scan &tgt, const std::vector<CPair> &cstl_in, const int lb[2],
                                    const c2g_sim_config &cont_sim, std::vector<CPair> &cstl_out,
                                    std::vector<float> &area_perc, int ret[2]) {
  ret[0] = ret[1] = 0;
  cstl_out.clear();
  area_perc.clear();
  for (auto pr : cstl_in)
    if (checkContPairSim(src, tgt, pr, cont_sim)) cstl_out.push_back(pr);
  ret[0] = (int) cstl_out.size();
  if (ret[0] < lb[0]) return;

  // 2.1 "major axis": the LAST qualifying (i, j) wins because shaft_src is normalised before the comparison
  float shaft_src[2] = {0, 0}, shaft_tgt[2] = {0, 0};
  for (int i = 1; i < std::min((int) cstl_out.size(), 10); i++) {
    for (int j = 0; j < i; j++) {
      const float *mi = src.cont_views[cstl_out[i].level][cstl_out[i].seq_src].pos_mean;
      const float *mj = src.cont_views[cstl_out[j].level][cstl_out[j].seq_src].pos_mean;
      float cx = mi[0] - mj[0], cy = mi[1] - mj[1];
      float cn = std::sqrt(cx * cx + cy * cy);
      float sn = std::sqrt(shaft_src[0] * shaft_src[0] + shaft_src[1] * shaft_src[1]);
      if (cn > sn) {
        // Eigen normalized(): n = squaredNorm(); n > 0 ? v / sqrt(n) : v
        float n2 = cx * cx + cy * cy;
        if (n2 > 0.0f) {
          float nn = std::sqrt(n2);
          shaft_src[0] = cx / nn;
          shaft_src[1] = cy / nn;
        } else {
          shaft_src[0] = cx;
          shaft_src[1] = cy;
        }
        const float *ti = tgt.cont_views[cstl_out[i].level][cstl_out[i].seq_tgt].pos_mean;
        const float *tj = tgt.cont_views[cstl_out[j].level][cstl_out[j].seq_tgt].pos_mean;
        float tx = ti[0] - tj[0], ty = ti[1] - tj[1];
        float t2 = tx * tx + ty * ty;
        if (t2 > 0.0f) {
          float tn = std::sqrt(t2);
          shaft_tgt[0] = tx / tn;
          shaft_tgt[1] = ty / tn;
        } else {
          shaft_tgt[0] = tx;
          shaft_tgt[1] = ty;
        }
      }
    }
  }
  // 2.2
  int num_sim = (int) cstl_out.size();
  for (int i = 0; i < num_sim;) {
    const c2g_view &sc1 = src.cont_views[cstl_out[i].level][cstl_out[i].seq_src];
    const c2g_view &tc1 = tgt.cont_views[cstl_out[i].level][cstl_out[i].seq_tgt];
    if (sc1.ecc_feat && tc1.ecc_feat) {
      // shaft^T * eig_vecs.col(1): 2-term dot product in float; std::acos(float) -> acosf
      float theta_s = std::acos(shaft_src[0] * sc1.eig_vecs[2] + shaft_src[1] * sc1.eig_vecs[3]);
      float theta_t = std::acos(shaft_tgt[0] * tc1.eig_vecs[2] + shaft_tgt[1] * tc1.eig_vecs[3]);
      if (diff_delt<float>(theta_s, theta_t, M_PI / 6) && diff_delt<float>(M_PI - theta_s, theta_t, M_PI / 6)) {
        std::swap(cstl_out[i], cstl_out[num_sim - 1]);
        num_sim--;
        continue;
      }
    }
    i++;
  }
  cstl_out.erase(cstl_out.begin() + num_sim, cstl_out.end());
  ret[1] = (int) cstl_out.size();
  if (ret[1] < lb[1]) return;
  for (const auto &i : cstl_out)
    area_perc.push_back(0.5f * (src.cont_perc[i.level][i.seq_src] + tgt.cont_perc[i.level][i.seq_tgt]));
}

// ContourManager::getTFFromConstell (contour_mng.h:1251-1277): Eigen::umeyama(src, tgt, false) then
// rotate(atan2(T10, T00)), pretranslate(T.block<2,1>(0,2)).  Closed-form 2-D solution (see header).
inline Iso2 getTFFromConstell(const Scan &src, const Scan &tgt, const std::vector<CPair> &cstl) {
  const int n = (int) cstl.size();
  const double one_over_n = 1.0 / (double) n;
  double sm[2] = {0, 0}, dm[2] = {0, 0};
  for (int i = 0; i < n; ++i) {
    const float *ps = src.cont_views[cstl[i].level][cstl[i].seq_src].pos_mean;
    const float *pt = tgt.cont_views[cstl[i].level][cstl[i].seq_tgt].pos_mean;
    sm[0] += (double) ps[0];
    sm[1] += (double) ps[1];
    dm[0] += (double) pt[0];
    dm[1] += (double) pt[1];
  }
  sm[0] *= one_over_n;
  sm[1] *= one_over_n;
  dm[0] *= one_over_n;
  dm[1] *= one_over_n;
  // sigma = (1/n) * dst_demean * src_demean^T
  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
  for (int i = 0; i < n; ++i) {
    const float *ps = src.cont_views[cstl[i].level][cstl[i].seq_src].pos_mean;
    const float *pt = tgt.cont_views[cstl[i].level][cstl[i].seq_tgt].pos_mean;
    double sx = (double) ps[0] - sm[0], sy = (double) ps[1] - sm[1];
    double dx = (double) pt[0] - dm[0], dy = (double) pt[1] - dm[1];
    s00 += dx * sx;
    s01 += dx * sy;
    s10 += dy * sx;
    s11 += dy * sy;
  }
  s00 *= one_over_n;
  s01 *= one_over_n;
  s10 *= one_over_n;
  s11 *= one_over_n;
  // R = argmax tr(R^T sigma) over rotations: angle = atan2(s10 - s01, s00 + s11)
  double ang = std::atan2(s10 - s01, s00 + s11);
  double c = std::cos(ang), s = std::sin(ang);
  // Rt.col(2).head(2) = dst_mean - R * src_mean
  double tx = dm[0] - (c * sm[0] - s * sm[1]);
  double ty = dm[1] - (s * sm[0] + c * sm[1]);
  // ret.rotate(atan2(T10, T00)); ret.pretranslate(t)
  return Iso2::fromAngTrans(std::atan2(s, c), tx, ty);
}

// ---------------------------------------------------------------------------------------------------------------
// GMMPair / ConstellCorrelation::initProblem + tryProblem (correlation.h:23-202), getEstSensTF (:287-296)
// ---------------------------------------------------------------------------------------------------------------
struct GMMEllipse {
  double cov[4];  // column-major
  double mu[2];
  double w;
};

struct GMMScanData {  // per level (levels 1..4) ellipses of one scan + auto correlation
  std::vector<std::vector<GMMEllipse>> ell;
  std::vector<std::vector<float>> max_majax;
  double auto_corr = 0;
};

inline GMMScanData buildGMMScan(const Scan &cm) {
  const int levels[4] = {1, 2, 3, 4};
  const double min_area_perc = 0.95, scale = 2.0;
  GMMScanData g;
  for (int li = 0; li < 4; ++li) {
    const int lev = levels[li];
    int cnt_run = 0, cnt_full = cm.layer_cell_cnt[lev];
    g.ell.emplace_back();
    g.max_majax.emplace_back();
    for (const auto &view : cm.cont_views[lev]) {
      if (cnt_run * 1.0 / cnt_full >= min_area_perc) break;
      float mc[4];
      getManualCov(view, mc);
      GMMEllipse e;
      for (int i = 0; i < 4; ++i) e.cov[i] = (double) mc[i];
      e.mu[0] = (double) view.pos_mean[0];
      e.mu[1] = (double) view.pos_mean[1];
      e.w = double(view.cell_cnt);
      g.ell.back().push_back(e);
      g.max_majax.back().push_back(std::sqrt(view.eig_vals[1]));
      cnt_run += view.cell_cnt;
    }
  }
  for (int li = 0; li < 4; ++li)
    for (size_t i = 0; i < g.ell[li].size(); i++)
      for (size_t j = 0; j < g.ell[li].size(); j++) {
        const GMMEllipse &a = g.ell[li][i], &b = g.ell[li][j];
        double c00 = scale * (a.cov[0] + b.cov[0]), c10 = scale * (a.cov[1] + b.cov[1]);
        double c01 = scale * (a.cov[2] + b.cov[2]), c11 = scale * (a.cov[3] + b.cov[3]);
        double mx = a.mu[0] - b.mu[0], my = a.mu[1] - b.mu[1];
        double det = c00 * c11 - c01 * c10;
        double invdet = 1.0 / det;
        double i00 = c11 * invdet, i01 = -c01 * invdet, i10 = -c10 * invdet, i11 = c00 * invdet;
        double qf = mx * (i00 * mx + i01 * my) + my * (i10 * mx + i11 * my);
        g.auto_corr += a.w * b.w / std::sqrt(det) * std::exp(-0.5 * qf);
      }
  return g;
}

// returns the normalised correlation at T (initProblem(T_init) == tryProblem(T_init) with selection done at T_init)
inline double gmmInitCorrelation(const GMMScanData &src, const GMMScanData &tgt, const Iso2 &T_init) {
  const double scale = 2.0;
  const double px = T_init.tx, py = T_init.ty, theta = std::atan2(T_init.m10, T_init.m00);
  const double c = std::cos(theta), s = std::sin(theta);
  double cost = 0;
  for (int li = 0; li < 4; ++li)
    for (size_t si = 0; si < src.ell[li].size(); si++)
      for (size_t ti = 0; ti < tgt.ell[li].size(); ti++) {
        const GMMEllipse &a = src.ell[li][si], &b = tgt.ell[li][ti];
        double qx, qy;
        T_init.apply(a.mu[0], a.mu[1], qx, qy);
        double dx = qx - b.mu[0], dy = qy - b.mu[1];
        if (!(std::sqrt(dx * dx + dy * dy) < 3.0 * (src.max_majax[li][si] + tgt.max_majax[li][ti]))) continue;
        // operator(): new_cov = scale * (R cov_s R^T + cov_t); new_mu = R mu_s + t - mu_t
        double r00 = c, r01 = -s, r10 = s, r11 = c;
        double a00 = a.cov[0], a10 = a.cov[1], a01 = a.cov[2], a11 = a.cov[3];
        double t00 = r00 * a00 + r01 * a10, t01 = r00 * a01 + r01 * a11;
        double t10 = r10 * a00 + r11 * a10, t11 = r10 * a01 + r11 * a11;
        double ra00 = t00 * r00 + t01 * r01, ra01 = t00 * r10 + t01 * r11;
        double ra10 = t10 * r00 + t11 * r01, ra11 = t10 * r10 + t11 * r11;
        double c00 = scale * (ra00 + b.cov[0]), c10 = scale * (ra10 + b.cov[1]);
        double c01 = scale * (ra01 + b.cov[2]), c11 = scale * (ra11 + b.cov[3]);
        double mx = (r00 * a.mu[0] + r01 * a.mu[1]) + px - b.mu[0];
        double my = (r10 * a.mu[0] + r11 * a.mu[1]) + py - b.mu[1];
        double det = c00 * c11 - c01 * c10;
        double invdet = 1.0 / det;
        double i00 = c11 * invdet, i01 = -c01 * invdet, i10 = -c10 * invdet, i11 = c00 * invdet;
        double qua = -0.5 * (mx * (i00 * mx + i01 * my) + my * (i10 * mx + i11 * my));
        cost += -b.w * a.w * 1.0 / std::sqrt(det) * std::exp(qua);
      }
  return -cost / std::sqrt(src.auto_corr * tgt.auto_corr);
}

// ConstellCorrelation::getEstSensTF (correlation.h:287-296): T_to_tsen^-1 * T_delta * T_so_ssen
inline Iso2 getEstSensTF(const Iso2 &T_delta, const c2g_cm_config &cfg) {
  Iso2 T_so;
  T_so.tx = cfg.n_row / 2 - 0.5;
  T_so.ty = cfg.n_col / 2 - 0.5;
  return T_so.inverse() * T_delta * T_so;
}

}  // namespace c2o
#include "c2o_refine.hpp"
namespace c2o {

// ---------------------------------------------------------------------------------------------------------------
// TreeBucket / LayerDB (contour_db.h:54-217, contour_db.cpp:63-403)
// ---------------------------------------------------------------------------------------------------------------
struct IndexOfKey {
  size_t gidx;
  int level, seq;
};
const float MAX_BUCKET_VAL = 1000.0f;
const float MAX_DIST_SQ = 1e6;

// nanoflann L2_Adaptor::evalMetric summation order (nanoflann.hpp:428-462), a = query
inline float evalMetricL2(const float *a, const float *b) {
  float result = 0.0f;
  int d = 0;
  for (; d + 3 < C2G_KEY_DIM; d += 4) {
    const float diff0 = a[d] - b[d], diff1 = a[d + 1] - b[d + 1], diff2 = a[d + 2] - b[d + 2], diff3 = a[d + 3] - b[d + 3];
    result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
  }
  for (; d < C2G_KEY_DIM; ++d) {
    const float diff0 = a[d] - b[d];
    result += diff0 * diff0;
  }
  return result;
}

struct TreeBucket {
  struct RetrTriplet {
    Key pt;
    double ts;
    IndexOfKey iok;
  };
  double max_elapse, min_elapse;
  float buc_beg, buc_end;
  std::vector<Key> data_tree;
  bool has_tree = false;      // tree_ptr != nullptr
  size_t indexed_size = 0;    // number of points in the built index
#ifdef C2O_USE_NANOFLANN
  std::shared_ptr<my_kd_tree_t> tree_ptr;
#endif
  std::vector<RetrTriplet> buffer;
  std::vector<IndexOfKey> gkidx_tree;

  size_t getTreeSize() const { return data_tree.size(); }
  void pushBuffer(const Key &k, double ts, IndexOfKey iok) { buffer.push_back(RetrTriplet{k, ts, iok}); }
  bool needPopBuffer(double curr_ts) const {
    double ts_overflow = curr_ts - max_elapse;
    if (buffer.empty() || buffer[0].ts > ts_overflow) return false;
    return true;
  }
  void rebuildTree() {
    has_tree = true;
    indexed_size = data_tree.size();
#ifdef C2O_USE_NANOFLANN
    if (data_tree.empty()) {
      tree_ptr.reset();
      has_tree = false;
    } else if (tree_ptr)
      tree_ptr->index->buildIndex();
    else
      tree_ptr = std::make_shared<my_kd_tree_t>(C2G_KEY_DIM, data_tree, 10);
#endif
  }
  void popBufferMax(double curr_ts) {
    double ts_cutoff = curr_ts - min_elapse;
    int gap = 0;
    for (; gap < (int) buffer.size(); gap++)
      if (buffer[gap].ts >= ts_cutoff) break;
    if (gap > 0) {
      for (int i = 0; i < gap; i++) {
        data_tree.push_back(buffer[i].pt);
        gkidx_tree.push_back(buffer[i].iok);
      }
      buffer.erase(buffer.begin(), buffer.begin() + gap);
      rebuildTree();
    }
  }
  // TreeBucket::knnSearch (contour_db.cpp:381-403) with MyKNNResSet (contour_db.h:32-52)
  void knnSearch(int num_res, std::vector<IndexOfKey> &ret_idx, std::vector<float> &out_dist_sq, const Key &q,
                 float max_dist_sq) const {
    ret_idx.clear();
    out_dist_sq.assign(num_res, MAX_DIST_SQ);
    if (!has_tree) return;
    std::vector<size_t> idx(num_res, 0);
#ifdef C2O_USE_NANOFLANN
    {
      MyKNNResSet<float> resultSet(num_res);
      resultSet.init(&idx[0], &out_dist_sq[0], max_dist_sq);
      tree_ptr->index->findNeighbors(resultSet, q.data(), nanoflann::SearchParams(10));
      for (int i = 0; i < num_res; i++) ret_idx.push_back(gkidx_tree[idx[i]]);
      return;
    }
#endif
    size_t count = 0;
    const size_t capacity = (size_t) num_res;
    if (capacity) out_dist_sq[capacity - 1] = max_dist_sq;
    // NOTE: the reference's index was built over data_tree at the last rebuildTree(); rebalancing (LayerDB::rebuild)
    // always ends in popBufferMax -> rebuildTree for both buckets when data moved, so indexed_size == data_tree.size()
    // whenever the two could differ materially; we follow the index (first indexed_size points).
    const size_t npts = std::min(indexed_size, data_tree.size());
    for (size_t p = 0; p < npts; ++p) {
      float dist = evalMetricL2(q.data(), data_tree[p].data());
      if (dist < out_dist_sq[capacity - 1]) {  // nanoflann.hpp:1575 `dist < worst_dist`
        size_t i;
        for (i = count; i > 0; --i) {
          if (out_dist_sq[i - 1] > dist) {
            if (i < capacity) {
              out_dist_sq[i] = out_dist_sq[i - 1];
              idx[i] = idx[i - 1];
            }
          } else
            break;
        }
        if (i < capacity) {
          out_dist_sq[i] = dist;
          idx[i] = p;
        }
        if (count < capacity) count++;
      }
    }
    for (int i = 0; i < num_res; i++) ret_idx.push_back(gkidx_tree.empty() ? IndexOfKey{0, 0, 0} : gkidx_tree[idx[i]]);
  }
};

struct LayerDB {
  static const int min_elem_split_ = 100;
  static constexpr double imba_diff_ratio_ = 0.2;
  static const int max_num_backets_ = C2G_NUM_BUCKETS;
  static const int bucket_chann_ = 0;
  std::vector<TreeBucket> buckets_;
  std::vector<float> bucket_ranges_;
  // how often the literal reference would have been left with a stale index (see the end of rebuild): rebalancing moves so far,
  // and buckets whose data a move changed while their own popBufferMax popped nothing
  long long n_moves_ = 0, n_stale_ = 0, n_stale_donor_ = 0;  // n_stale_donor_: ... of which the bucket that GAVE keys (its tree was permuted and cut)

  explicit LayerDB(double max_elapse, double min_elapse) {
    bucket_ranges_.resize(max_num_backets_ + 1);
    bucket_ranges_.front() = -MAX_BUCKET_VAL;
    bucket_ranges_.back() = MAX_BUCKET_VAL;
    TreeBucket b0;
    b0.max_elapse = max_elapse;
    b0.min_elapse = min_elapse;
    b0.buc_beg = -MAX_BUCKET_VAL;
    b0.buc_end = MAX_BUCKET_VAL;
    buckets_.push_back(b0);
    for (int i = 1; i < max_num_backets_; i++) {
      bucket_ranges_[i] = MAX_BUCKET_VAL;
      TreeBucket b;
      b.max_elapse = max_elapse;
      b.min_elapse = min_elapse;
      b.buc_beg = MAX_BUCKET_VAL;
      b.buc_end = MAX_BUCKET_VAL;
      buckets_.push_back(b);
    }
  }

  void pushBuffer(const Key &layer_key, double ts, IndexOfKey iok) {  // contour_db.h:184-192
    for (int i = 0; i < max_num_backets_; i++) {
      if (bucket_ranges_[i] <= layer_key[bucket_chann_] && layer_key[bucket_chann_] < bucket_ranges_[i + 1]) {
        if (keySum(layer_key) != 0) buckets_[i].pushBuffer(layer_key, ts, iok);
        return;
      }
    }
  }

  void rebuild(int idx_t1, double curr_ts);  // contour_db.cpp:63-317

  // contour_db.cpp:319-379
  void layerKNNSearch(const Key &q_key, const int k_top, const float max_dist_sq,
                      std::vector<std::pair<IndexOfKey, float>> &res_pairs) const {
    int mid_bucket = 0;
    for (int i = 0; i < max_num_backets_; i++) {
      if (bucket_ranges_[i] <= q_key[bucket_chann_] && bucket_ranges_[i + 1] > q_key[bucket_chann_]) {
        mid_bucket = i;
        break;
      }
    }
    float max_dist_sq_run = max_dist_sq;
    res_pairs.clear();
    for (int i = 0; i < max_num_backets_; i++) {
      std::vector<IndexOfKey> tmp_gidx;
      std::vector<float> tmp_dists_sq;
      int which = -1;
      if (i == 0) {
        which = mid_bucket;
      } else if (mid_bucket - i >= 0) {
        float d = q_key[bucket_chann_] - bucket_ranges_[mid_bucket - i + 1];
        if (d * d > max_dist_sq_run) continue;
        which = mid_bucket - i;
      } else if (mid_bucket + i < max_num_backets_) {
        float d = q_key[bucket_chann_] - bucket_ranges_[mid_bucket + i];
        if (d * d > max_dist_sq_run) continue;
        which = mid_bucket + i;
      }
      if (which >= 0) {
        buckets_[which].knnSearch(k_top, tmp_gidx, tmp_dists_sq, q_key, max_dist_sq_run);
        for (int j = 0; j < k_top; j++)
          if (tmp_dists_sq[j] < max_dist_sq_run) {
            if (j < (int) tmp_gidx.size()) res_pairs.emplace_back(tmp_gidx[j], tmp_dists_sq[j]);
          } else
            break;
      }
      std::sort(res_pairs.begin(), res_pairs.end(),
                [](const std::pair<IndexOfKey, float> &a, const std::pair<IndexOfKey, float> &b) { return a.second < b.second; });
      if ((int) res_pairs.size() >= k_top) {
        res_pairs.resize(k_top, res_pairs[0]);
        max_dist_sq_run = res_pairs.back().second;
      }
    }
  }
};

inline void LayerDB::rebuild(int idx_t1, double curr_ts) {
  TreeBucket &tr1 = buckets_[idx_t1], &tr2 = buckets_[idx_t1 + 1];
  bool pb1 = tr1.needPopBuffer(curr_ts), pb2 = tr2.needPopBuffer(curr_ts);
  if (!pb1 && !pb2) return;
  int sz1 = (int) tr1.getTreeSize(), sz2 = (int) tr2.getTreeSize();
  double diff_ratio = 1.0 * std::abs(sz1 - sz2) / std::max(sz1, sz2);  // 0/0 = NaN when both empty (as in the reference)
  if (pb1 && !pb2 && (diff_ratio < imba_diff_ratio_ || std::max(sz1, sz2) < min_elem_split_)) {
    tr1.popBufferMax(curr_ts);
    return;
  }
  if (!pb1 && pb2 && (diff_ratio < imba_diff_ratio_ || std::max(sz1, sz2) < min_elem_split_)) {
    tr2.popBufferMax(curr_ts);
    return;
  }
  if (diff_ratio < 0.5 * imba_diff_ratio_) {
    if (pb1) tr1.popBufferMax(curr_ts);
    if (pb2) tr2.popBufferMax(curr_ts);
    return;
  }
  if (sz1 > sz2) {
    int to_move_max = int((sz1 - sz2 + imba_diff_ratio_ * sz2) / (2 - imba_diff_ratio_));
    int to_move_mid = int((sz1 - sz2) / 2.0);
    int to_move_min = std::max(0, int((sz1 - sz2 - imba_diff_ratio_ * sz1) / (2 - imba_diff_ratio_)));
    std::vector<int> sort_permu(sz1);
    std::iota(sort_permu.begin(), sort_permu.end(), 0);
    std::sort(sort_permu.begin(), sort_permu.end(),
              [&](const int &a, const int &b) { return tr1.data_tree[a][bucket_chann_] < tr1.data_tree[b][bucket_chann_]; });
    int num_to_move = 0;
    float split_val = tr1.buc_end;
    if (tr1.data_tree[sort_permu[sz1 - to_move_mid]][bucket_chann_] != tr1.data_tree[sort_permu[sz1 - to_move_mid - 1]][bucket_chann_]) {
      num_to_move = to_move_mid;
      split_val = tr1.data_tree[sort_permu[sz1 - to_move_mid]][bucket_chann_];
    } else {
      float contagious_val = tr1.data_tree[sort_permu[sz1 - to_move_mid]][bucket_chann_];
      int i = to_move_mid - 1;
      for (; i > to_move_min; i--) {
        if (tr1.data_tree[sort_permu[sz1 - i]][bucket_chann_] != contagious_val) {
          num_to_move = i;
          split_val = tr1.data_tree[sort_permu[sz1 - i]][bucket_chann_];
          break;
        }
      }
      if (num_to_move == 0) {
        i = to_move_mid + 1;
        for (; i < to_move_max; i++) {
          if (tr1.data_tree[sort_permu[sz1 - i]][bucket_chann_] != contagious_val) {
            num_to_move = i - 1;
            split_val = contagious_val;
            break;
          }
        }
      }
    }
    if (num_to_move == 0) {
      tr1.popBufferMax(curr_ts);
      if (pb2) tr2.popBufferMax(curr_ts);
      return;
    }
    for (int i = 0; i < num_to_move; i++) {
      tr2.data_tree.push_back(tr1.data_tree[sort_permu[sz1 - i - 1]]);
      tr2.gkidx_tree.push_back(tr1.gkidx_tree[sort_permu[sz1 - i - 1]]);
    }
    int p_dat = sz1 - 1, p_perm = sz1 - 1;
    for (; p_perm >= sz1 - num_to_move; p_perm--) {
      while (tr1.data_tree[p_dat][bucket_chann_] >= split_val) p_dat--;
      if (sort_permu[p_perm] < p_dat) {
        std::swap(tr1.data_tree[p_dat], tr1.data_tree[sort_permu[p_perm]]);
        std::swap(tr1.gkidx_tree[p_dat], tr1.gkidx_tree[sort_permu[p_perm]]);
        p_dat--;
      }
    }
    tr1.data_tree.resize(p_dat + 1);
    tr1.gkidx_tree.resize(p_dat + 1, tr1.gkidx_tree[0]);
    int p1 = 0, p2 = (int) tr1.buffer.size() - 1, sz_rem;
    while (p1 <= p2) {
      if (tr1.buffer[p1].pt[bucket_chann_] >= split_val && tr1.buffer[p2].pt[bucket_chann_] < split_val) {
        std::swap(tr1.buffer[p1], tr1.buffer[p2]);
        p1++;
        p2--;
      } else {
        if (tr1.buffer[p2].pt[bucket_chann_] >= split_val) p2--;
        if (tr1.buffer[p1].pt[bucket_chann_] < split_val) p1++;
      }
    }
    sz_rem = p2 + 1;
    tr2.buffer.insert(tr2.buffer.end(), tr1.buffer.begin() + sz_rem, tr1.buffer.end());
    tr1.buffer.resize(sz_rem, tr1.buffer.empty() ? TreeBucket::RetrTriplet{} : tr1.buffer[0]);
    tr1.buc_end = tr2.buc_beg = split_val;
    bucket_ranges_[idx_t1 + 1] = split_val;
  } else {
    int to_move_max = int((sz2 - sz1 + imba_diff_ratio_ * sz1) / (2 - imba_diff_ratio_));
    int to_move_mid = int((sz2 - sz1) / 2.0);
    int to_move_min = std::max(0, int((sz2 - sz1 - imba_diff_ratio_ * sz2) / (2 - imba_diff_ratio_)));
    std::vector<int> sort_permu(sz2);
    std::iota(sort_permu.begin(), sort_permu.end(), 0);
    std::sort(sort_permu.begin(), sort_permu.end(),
              [&](const int &a, const int &b) { return tr2.data_tree[a][bucket_chann_] > tr2.data_tree[b][bucket_chann_]; });
    int num_to_move = 0;
    float split_val = tr1.buc_end;
    if (tr2.data_tree[sort_permu[sz2 - to_move_mid]][bucket_chann_] != tr2.data_tree[sort_permu[sz2 - to_move_mid - 1]][bucket_chann_]) {
      num_to_move = to_move_mid;
      split_val = tr2.data_tree[sort_permu[sz2 - to_move_mid - 1]][bucket_chann_];
    } else {
      float contagious_val = tr2.data_tree[sort_permu[sz2 - to_move_mid]][bucket_chann_];
      int i = to_move_mid - 1;
      for (; i > to_move_min; i--) {
        if (tr2.data_tree[sort_permu[sz2 - i]][bucket_chann_] != contagious_val) {
          num_to_move = i;
          split_val = contagious_val;
          break;
        }
      }
      if (num_to_move == 0) {
        i = to_move_mid + 1;
        for (; i < to_move_max; i++) {
          if (tr2.data_tree[sort_permu[sz2 - i]][bucket_chann_] != contagious_val) {
            num_to_move = i - 1;
            split_val = tr2.data_tree[sort_permu[sz2 - i]][bucket_chann_];
            break;
          }
        }
      }
    }
    if (num_to_move == 0) {
      if (pb1) tr1.popBufferMax(curr_ts);
      tr2.popBufferMax(curr_ts);
      return;
    }
    for (int i = 0; i < num_to_move; i++) {
      tr1.data_tree.push_back(tr2.data_tree[sort_permu[sz2 - i - 1]]);
      tr1.gkidx_tree.push_back(tr2.gkidx_tree[sort_permu[sz2 - i - 1]]);
    }
    int p_dat = sz2 - 1, p_perm = sz2 - 1;
    for (; p_perm >= sz2 - num_to_move; p_perm--) {
      while (tr2.data_tree[p_dat][bucket_chann_] < split_val) p_dat--;
      if (sort_permu[p_perm] < p_dat) {
        std::swap(tr2.data_tree[p_dat], tr2.data_tree[sort_permu[p_perm]]);
        std::swap(tr2.gkidx_tree[p_dat], tr2.gkidx_tree[sort_permu[p_perm]]);
        p_dat--;
      }
    }
    tr2.data_tree.resize(p_dat + 1);
    tr2.gkidx_tree.resize(p_dat + 1, tr2.gkidx_tree[0]);
    int p1 = 0, p2 = (int) tr2.buffer.size() - 1, sz_rem;
    while (p1 <= p2) {
      if (tr2.buffer[p1].pt[bucket_chann_] < split_val && tr2.buffer[p2].pt[bucket_chann_] >= split_val) {
        std::swap(tr2.buffer[p1], tr2.buffer[p2]);
        p1++;
        p2--;
      } else {
        if (tr2.buffer[p2].pt[bucket_chann_] < split_val) p2--;
        if (tr2.buffer[p1].pt[bucket_chann_] >= split_val) p1++;
      }
    }
    sz_rem = p2 + 1;
    tr1.buffer.insert(tr1.buffer.end(), tr2.buffer.begin() + sz_rem, tr2.buffer.end());
    tr2.buffer.resize(sz_rem, tr2.buffer.empty() ? TreeBucket::RetrTriplet{} : tr2.buffer[0]);
    tr1.buc_end = tr2.buc_beg = split_val;
    bucket_ranges_[idx_t1 + 1] = split_val;
  }
  std::sort(tr1.buffer.begin(), tr1.buffer.end(), [](const auto &a, const auto &b) { return a.ts < b.ts; });
  std::sort(tr2.buffer.begin(), tr2.buffer.end(), [](const auto &a, const auto &b) { return a.ts < b.ts; });
  // NOTE: rebalancing changed data_tree of both buckets; the reference rebuilds the index only inside popBufferMax (when
  // something is popped).  The RECEIVER's tree only grew at its end: a stale index over its first indexed_size points is
  // well defined (nanoflann reads the vector through a reference) and is kept - the moved keys become searchable at the
  // bucket's next pop, exactly as in the reference.  The DONOR's tree was permuted and cut: a stale index addresses points
  // beyond its end (undefined behaviour in the reference); it is rebuilt here, which is what a run that does not crash
  // observes (n_stale_donor_ counts how often).
  tr1.popBufferMax(curr_ts);
  tr2.popBufferMax(curr_ts);
  n_moves_++;
  n_stale_ += (tr1.indexed_size != tr1.data_tree.size()) + (tr2.indexed_size != tr2.data_tree.size());
  TreeBucket &donor = sz1 > sz2 ? tr1 : tr2;
  if (donor.indexed_size != donor.data_tree.size() || !donor.has_tree) {
    n_stale_donor_++;
    donor.rebuildTree();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CandidateManager (contour_db.h:264-656)
// ---------------------------------------------------------------------------------------------------------------
struct HintTrace {
  c2g_hint hint;
  c2g_pair_score score;
};

struct CandidateManager {
  struct CandidateAnchorProp {
    std::map<CPair, float> constell_;
    Iso2 T_delta_;
    float correlation_ = 0;
    int vote_cnt_ = 0;
    float area_perc_ = 0;
  };
  struct CandidatePoseData {
    std::shared_ptr<const Scan> cm_cand_;
    size_t gidx = 0;
    bool has_corr_est = false;  // corr_est_ != nullptr
    float corr_init = 0;
    double neg_est_dist = 0;
    Iso2 T_init;  // anch_props_[0].T_delta_ before the refinement overwrote it
    int fine_iters = -1, fine_term = 0;
    std::vector<CandidateAnchorProp> anch_props_;
    void addProposal(const Iso2 &T_prop, const std::vector<CPair> &sim_pairs, const std::vector<float> &sim_area_perc) {
      for (size_t i = 0; i < anch_props_.size(); i++) {
        const Iso2 delta_T = T_prop.inverse() * anch_props_[i].T_delta_;
        if (std::sqrt(delta_T.tx * delta_T.tx + delta_T.ty * delta_T.ty) < 2.0 &&
            std::abs(std::atan2(delta_T.m10, delta_T.m00)) < 0.3) {
          for (size_t j = 0; j < sim_pairs.size(); j++) anch_props_[i].constell_.insert({sim_pairs[j], sim_area_perc[j]});
          anch_props_[i].vote_cnt_ += (int) sim_pairs.size();
          int w1 = anch_props_[i].vote_cnt_, w2 = (int) sim_pairs.size();
          double tbx = (anch_props_[i].T_delta_.tx * w1 + T_prop.tx * w2) / (w1 + w2);
          double tby = (anch_props_[i].T_delta_.ty * w1 + T_prop.ty * w2) / (w1 + w2);
          double ang1 = std::atan2(anch_props_[i].T_delta_.m10, anch_props_[i].T_delta_.m00);
          double ang2 = std::atan2(T_prop.m10, T_prop.m00);
          double diff = ang2 - ang1;
          if (diff < 0) diff += 2 * M_PI;
          if (diff > M_PI) diff -= 2 * M_PI;
          double ang_bl = diff * w2 / (w1 + w2) + ang1;
          anch_props_[i].T_delta_ = Iso2::fromAngTrans(ang_bl, tbx, tby);
          return;
        }
      }
      if (anch_props_.size() > 3) return;
      anch_props_.emplace_back();
      anch_props_.back().T_delta_ = T_prop;
      for (size_t j = 0; j < sim_pairs.size(); j++) anch_props_.back().constell_.insert({sim_pairs[j], sim_area_perc[j]});
      anch_props_.back().vote_cnt_ = (int) sim_pairs.size();
    }
  };

  std::shared_ptr<const Scan> cm_tgt_;
  c2g_score_ensemble sim_var_, sim_ub_;
  std::map<int, int> cand_id_pos_pair_;
  std::vector<CandidatePoseData> candidates_;
  int cand_aft_check1 = 0, cand_aft_check2 = 0, cand_aft_check3 = 0;
  std::vector<HintTrace> *trace = nullptr;

  CandidateManager(std::shared_ptr<const Scan> q, const c2g_score_ensemble &lb, const c2g_score_ensemble &ub)
      : cm_tgt_(std::move(q)), sim_var_(lb), sim_ub_(ub) {}

  // checkCandWithHint (contour_db.h:374-488). src = candidate, tgt = query.
  void checkCandWithHint(const std::shared_ptr<const Scan> &cm_cand, size_t gidx, const CPair &anchor_pair,
                         const c2g_sim_config &cont_sim, c2g_pair_score &rec) {
    std::memset(&rec, 0, sizeof(rec));
    int cand_id = cm_cand->int_id;
    bool anchor_sim = checkContPairSim(*cm_cand, *cm_tgt_, anchor_pair, cont_sim);
    if (!anchor_sim) {
      rec.passed = 0;
      return;
    }
    cand_aft_check1++;
    std::vector<CPair> tmp_pairs1;
    const int lbc[3] = {sim_var_.i_ovlp_sum, sim_var_.i_ovlp_max_one, sim_var_.i_in_ang_rng};
    checkConstellSim(cm_cand->layer_key_bcis[anchor_pair.level][anchor_pair.seq_src],
                     cm_tgt_->layer_key_bcis[anchor_pair.level][anchor_pair.seq_tgt], lbc, rec.constell, tmp_pairs1);
    if (rec.constell[2] < sim_var_.i_in_ang_rng) {
      rec.passed = -1;
      return;
    }
    cand_aft_check2++;
    std::vector<CPair> tmp_pairs2;
    std::vector<float> tmp_area_perc;
    const int lbp[2] = {sim_var_.i_indiv_sim, sim_var_.i_orie_sim};
    checkConstellCorrespSim(*cm_cand, *cm_tgt_, tmp_pairs1, lbp, cont_sim, tmp_pairs2, tmp_area_perc, rec.pairwise);
    if (rec.pairwise[1] < sim_var_.i_orie_sim) {
      rec.passed = -2;
      return;
    }
    cand_aft_check3++;
    Iso2 T_pass = getTFFromConstell(*cm_cand, *cm_tgt_, tmp_pairs2);
    rec.passed = 1;
    rec.n_pairs = (int) tmp_pairs2.size();
    rec.T[0] = T_pass.m00;
    rec.T[1] = T_pass.m10;
    rec.T[2] = T_pass.tx;
    rec.T[3] = T_pass.ty;
    for (auto &p : tmp_pairs2) {
      int bit = (p.level - 1) * 100 + p.seq_src * 10 + p.seq_tgt;
      rec.pair_bits[bit >> 6] |= (uint64_t) 1 << (bit & 63);
    }
    auto cand_it = cand_id_pos_pair_.find(cand_id);
    if (cand_it != cand_id_pos_pair_.end()) {
      candidates_[cand_it->second].addProposal(T_pass, tmp_pairs2, tmp_area_perc);
    } else {
      CandidatePoseData new_cand;
      new_cand.cm_cand_ = cm_cand;
      new_cand.gidx = gidx;
      new_cand.addProposal(T_pass, tmp_pairs2, tmp_area_perc);
      cand_id_pos_pair_.insert({cand_id, (int) candidates_.size()});
      candidates_.emplace_back(std::move(new_cand));
    }
  }

  // tidyUpCandidates (contour_db.h:494-596)
  int n_pose_before = 0;
  void tidyUpCandidates(const std::function<const GMMScanData &(const Scan *)> &gmm_of) {
    const int DIST_BIN_LAYERS[4] = {1, 2, 3, 4};
    const float LAYER_AREA_WEIGHTS[4] = {0.3, 0.3, 0.3, 0.1};
    n_pose_before = (int) candidates_.size();
    int cnt_to_rm = 0;
    for (auto &candidate : candidates_) {
      int idx_sel = 0;
      for (int i = 0; i < (int) candidate.anch_props_.size(); i++) {
        std::vector<float> lev_perc(cm_tgt_->cfg.n_levels, 0);
        for (const auto &pr : candidate.anch_props_[i].constell_) lev_perc[pr.first.level] += pr.second;
        float perc = 0;
        for (int j = 0; j < C2G_NUM_BIN_LAYERS; j++) perc += LAYER_AREA_WEIGHTS[j] * lev_perc[DIST_BIN_LAYERS[j]];
        candidate.anch_props_[i].area_perc_ = perc;
        if (candidate.anch_props_[i].vote_cnt_ > candidate.anch_props_[idx_sel].vote_cnt_) idx_sel = i;
      }
      std::swap(candidate.anch_props_[0], candidate.anch_props_[idx_sel]);
      if (candidate.anch_props_[0].area_perc_ < sim_var_.area_perc) {
        cnt_to_rm++;
        continue;
      }
      Iso2 est = getEstSensTF(candidate.anch_props_[0].T_delta_, cm_tgt_->cfg);
      double neg_est_trans_norm2d = -std::sqrt(est.tx * est.tx + est.ty * est.ty);
      candidate.neg_est_dist = neg_est_trans_norm2d;
      if (neg_est_trans_norm2d < sim_var_.neg_est_dist) {
        cnt_to_rm++;
        continue;
      }
      auto corr_score_init =
          (float) gmmInitCorrelation(gmm_of(candidate.cm_cand_.get()), gmm_of(cm_tgt_.get()), candidate.anch_props_[0].T_delta_);
      candidate.corr_init = corr_score_init;
      if (corr_score_init < sim_var_.correlation) {
        cnt_to_rm++;
        continue;
      }
      candidate.has_corr_est = true;
    }
    int p1 = 0, p2 = (int) candidates_.size() - 1;
    while (p1 <= p2) {
      if (!candidates_[p1].has_corr_est && candidates_[p2].has_corr_est) {
        std::swap(candidates_[p1], candidates_[p2]);
        p1++;
        p2--;
      } else {
        if (candidates_[p1].has_corr_est) p1++;
        if (!candidates_[p2].has_corr_est) p2--;
      }
    }
    (void) cnt_to_rm;
    candidates_.erase(candidates_.begin() + p2 + 1, candidates_.end());
  }

  // fineOptimize (contour_db.h:604-648). Returns index of the top-1 in candidates_ order after the sorts, or -1.
  int fineOptimize(int max_fine_opt, const std::function<const GMMScanData &(const Scan *)> &gmm_of) {
    if (candidates_.empty()) return -1;
    for (auto &c : candidates_) c.T_init = c.anch_props_[0].T_delta_;
    std::sort(candidates_.begin(), candidates_.end(), [](const CandidatePoseData &d1, const CandidatePoseData &d2) {
      return d1.anch_props_[0].correlation_ > d2.anch_props_[0].correlation_;
    });
    int pre_sel_size = std::min(max_fine_opt, (int) candidates_.size());
    for (int i = 0; i < pre_sel_size; i++) {
      auto &c = candidates_[i];
      const refine::CorrResult r = refine::calcCorrelation(gmm_of(c.cm_cand_.get()), gmm_of(cm_tgt_.get()), c.anch_props_[0].T_delta_);
      c.anch_props_[0].correlation_ = (float) r.correlation;
      c.anch_props_[0].T_delta_ = r.T;
      c.fine_iters = r.opt.iterations;
      c.fine_term = r.opt.termination;
    }
    std::sort(candidates_.begin(), candidates_.begin() + pre_sel_size, [](const CandidatePoseData &d1, const CandidatePoseData &d2) {
      return d1.anch_props_[0].correlation_ > d2.anch_props_[0].correlation_;
    });
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// ContourDB (contour_db.h:673-845)
// ---------------------------------------------------------------------------------------------------------------
struct ContourDB {
  c2g_db_config cfg_;
  std::vector<LayerDB> layer_db_;
  std::vector<std::shared_ptr<const Scan>> all_bevs_;
  std::map<const Scan *, GMMScanData> gmm_cache_;  // per-scan ellipses + auto-correlation (scan-only quantities)

  explicit ContourDB(const c2g_db_config &c) : cfg_(c) {
    for (int i = 0; i < cfg_.n_q_levels; ++i) layer_db_.emplace_back(cfg_.max_elapse, cfg_.min_elapse);
  }

  // DB scans get their scan-only GMM terms when they are added (read-only afterwards, so concurrent queries are safe);
  // the query scan's own terms are built once per query.  NOTE: the reference rebuilds both inside every initProblem
  // (correlation.h:42-122); caching them makes this CPU baseline FASTER than the reference, never slower.
  const GMMScanData &gmmOfDb(const Scan *s) const { return gmm_cache_.at(s); }

  // queryRangedKNN (contour_db.h:698-811)
  void queryRangedKNN(const std::shared_ptr<const Scan> &q_ptr, const c2g_score_ensemble &thres_lb,
                      const c2g_score_ensemble &thres_ub, c2g_query_result &out, std::vector<HintTrace> *trace,
                      double *t_knn = nullptr, double *t_constell = nullptr, double *t_l2 = nullptr) {
    std::memset(&out, 0, sizeof(out));
    out.best = -1;
    CandidateManager cand_mng(q_ptr, thres_lb, thres_ub);
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int ll = 0; ll < cfg_.n_q_levels; ll++) {
      const int lev = cfg_.q_levels[ll];
      const auto &q_keys = q_ptr->layer_keys[lev];
      for (int seq = 0; seq < (int) q_keys.size(); seq++) {
        if (keySum(q_keys[seq]) != 0) {
          double t0 = now();
          std::vector<std::pair<IndexOfKey, float>> tmp_res;
          float key_bounds[3][2];
          key_bounds[0][0] = q_keys[seq][0] * 0.8;
          key_bounds[0][1] = q_keys[seq][0] / 0.8;
          key_bounds[1][0] = q_keys[seq][1] * 0.8;
          key_bounds[1][1] = q_keys[seq][1] / 0.8;
          key_bounds[2][0] = q_keys[seq][2] * 0.8 * 0.75;
          key_bounds[2][1] = q_keys[seq][2] / (0.8 * 0.75);
          float dist_ub = 1e6;
          dist_ub = std::max((q_keys[seq][0] - key_bounds[0][0]) * (q_keys[seq][0] - key_bounds[0][0]),
                             (q_keys[seq][0] - key_bounds[0][1]) * (q_keys[seq][0] - key_bounds[0][1])) +
                    std::max((q_keys[seq][1] - key_bounds[1][0]) * (q_keys[seq][1] - key_bounds[1][0]),
                             (q_keys[seq][1] - key_bounds[1][1]) * (q_keys[seq][1] - key_bounds[1][1])) +
                    std::max((q_keys[seq][2] - key_bounds[2][0]) * (q_keys[seq][2] - key_bounds[2][0]),
                             (q_keys[seq][2] - key_bounds[2][1]) * (q_keys[seq][2] - key_bounds[2][1]));
          layer_db_[ll].layerKNNSearch(q_keys[seq], cfg_.nnk, dist_ub, tmp_res);
          double t1 = now();
          if (t_knn) *t_knn += t1 - t0;
          for (const auto &sear_res : tmp_res) {
            c2g_pair_score rec;
            CPair ap{(int8_t) lev, (int8_t) sear_res.first.seq, (int8_t) seq};
            cand_mng.checkCandWithHint(all_bevs_[sear_res.first.gidx], sear_res.first.gidx, ap, cfg_.cont_sim, rec);
            if (trace) {
              HintTrace ht;
              ht.hint.q_idx = 0;
              ht.hint.cand_gidx = (int32_t) sear_res.first.gidx;
              ht.hint.level = (int8_t) lev;
              ht.hint.cand_seq = (int8_t) sear_res.first.seq;
              ht.hint.q_seq = (int8_t) seq;
              ht.hint.q_level_idx = (int8_t) ll;
              ht.hint.dist_sq = sear_res.second;
              ht.score = rec;
              trace->push_back(ht);
            }
          }
          if (t_constell) *t_constell += now() - t1;
        }
      }
    }
    double t2 = now();
    std::unique_ptr<GMMScanData> q_gmm;
    const Scan *q_raw = q_ptr.get();
    const std::function<const GMMScanData &(const Scan *)> gmm_of = [this, &q_gmm, q_raw](const Scan *s) -> const GMMScanData & {
      if (s == q_raw) {
        if (!q_gmm) q_gmm.reset(new GMMScanData(buildGMMScan(*s)));
        return *q_gmm;
      }
      return gmmOfDb(s);
    };
    cand_mng.tidyUpCandidates(gmm_of);
    cand_mng.fineOptimize(cfg_.max_fine_opt, gmm_of);
    if (t_l2) *t_l2 += now() - t2;
    out.n_pose_before = cand_mng.n_pose_before;
    out.cand_aft_check[0] = cand_mng.cand_aft_check1;
    out.cand_aft_check[1] = cand_mng.cand_aft_check2;
    out.cand_aft_check[2] = cand_mng.cand_aft_check3;
    out.n_cand = (int) std::min<size_t>(cand_mng.candidates_.size(), C2G_MAX_CAND);
    out.overflow = cand_mng.candidates_.size() > C2G_MAX_CAND;
    for (int i = 0; i < out.n_cand; ++i) {
      const auto &c = cand_mng.candidates_[i];
      out.cand[i].cand_gidx = (int32_t) c.gidx;
      out.cand[i].vote_cnt = c.anch_props_[0].vote_cnt_;
      out.cand[i].area_perc = c.anch_props_[0].area_perc_;
      out.cand[i].corr_init = c.corr_init;
      out.cand[i].neg_est_dist = c.neg_est_dist;
      out.cand[i].T[0] = c.T_init.m00;
      out.cand[i].T[1] = c.T_init.m10;
      out.cand[i].T[2] = c.T_init.tx;
      out.cand[i].T[3] = c.T_init.ty;
      out.cand[i].corr_fine = c.anch_props_[0].correlation_;
      out.cand[i].fine_iters = (int16_t) c.fine_iters;
      out.cand[i].fine_term = (int8_t) c.fine_term;
      out.cand[i].fine_flags = 0;
      out.cand[i].T_fine[0] = c.anch_props_[0].T_delta_.m00;
      out.cand[i].T_fine[1] = c.anch_props_[0].T_delta_.m10;
      out.cand[i].T_fine[2] = c.anch_props_[0].T_delta_.tx;
      out.cand[i].T_fine[3] = c.anch_props_[0].T_delta_.ty;
    }
    out.best = out.n_cand > 0 ? 0 : -1;
  }

  // addScan (contour_db.h:814-824)
  void addScan(const std::shared_ptr<Scan> &added, double curr_timestamp) {
    for (int ll = 0; ll < cfg_.n_q_levels; ll++) {
      int seq = 0;
      for (const auto &permu_key : added->layer_keys[cfg_.q_levels[ll]]) {
        if (keySum(permu_key) != 0)
          layer_db_[ll].pushBuffer(permu_key, curr_timestamp, IndexOfKey{all_bevs_.size(), cfg_.q_levels[ll], seq});
        seq++;
      }
    }
    gmm_cache_.emplace(added.get(), buildGMMScan(*added));
    all_bevs_.emplace_back(added);
  }

  // pushAndBalance (contour_db.h:827-843)
  void pushAndBalance(int seed, double curr_timestamp) {
    int idx_t1 = std::abs(seed) % (2 * (LayerDB::max_num_backets_ - 2));
    if (idx_t1 > (LayerDB::max_num_backets_ - 2)) idx_t1 = 2 * (LayerDB::max_num_backets_ - 2) - idx_t1;
    for (int ll = 0; ll < cfg_.n_q_levels; ll++) layer_db_[ll].rebuild(idx_t1, curr_timestamp);
  }
};

}  // namespace c2o
