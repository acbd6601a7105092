// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped hot path; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call it.
//
// c2o_ingest.hpp: CPU restatement of the per-scan half of the reference's cont2contops path
// (point cloud -> BEV -> recursive multi-level CCL -> ContourView statistics -> sort -> retrieval keys -> BCI).
// Plain C++17, no third-party dependency; float arithmetic must be compiled WITHOUT FMA contraction
// (-ffp-contract=off, no -march) because the reference is built that way (CMakeLists.txt:4,10-11).
//
// PARITY STATUS: "parity unpinned" for the pieces whose arithmetic lives in libraries that are not vendored in the
// reference tree (OpenCV connectedComponentsWithStats label order, Eigen SelfAdjointEigenSolver<Matrix2f>, Eigen
// umeyama, Ceres): the reference ships no executable golden vector for this path.  What pins
// this file instead: (1) CCL label order/stats are cross-checked against the real OpenCV (python cv2) in
// tests/test_oracle_ccl_cv2.py; (2) the eigen-solver restatement is cross-checked against numpy.linalg.eigh;
// (3) the kNN semantics are cross-checked against the reference's own vendored nanoflann (oracle/_ref).
//
// Every function cites the reference file:line it follows (paths relative to the reference repo root).
#pragma once

#include <algorithm>
#include <array>
#include <limits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "../include/c2g_types.h"

namespace c2o {

// ---------------------------------------------------------------------------------------------------------------
// tools/algos.h
// ---------------------------------------------------------------------------------------------------------------
// include/tools/algos.h:12-15
template <typename T>
inline bool diff_perc(const T &num1, const T &num2, const T &perc) {
  return std::abs((num1 - num2) / std::max(num1, num2)) > perc;
}
// include/tools/algos.h:17-20
template <typename T>
inline bool diff_delt(const T &num1, const T &num2, const T &delta) {
  return std::abs(num1 - num2) > delta;
}
// include/tools/algos.h:48-51 (computed in double through M_PI, then stored to T)
template <typename T>
inline void clampAng(T &ang) {
  ang = ang - std::floor((ang + M_PI) / (2 * M_PI)) * 2 * M_PI;
}
// include/tools/algos.h:53-56 (for T=float: exp and sqrt are evaluated in double, result rounded to float)
template <typename T>
inline T gaussPDF(const T &x, const T &mean, const T &sd) {
  return std::exp(-0.5 * ((x - mean) / sd) * ((x - mean) / sd)) / std::sqrt(2 * M_PI * sd * sd);
}

// ---------------------------------------------------------------------------------------------------------------
// Eigen::SelfAdjointEigenSolver<Matrix2f> restated (Eigen 3.3.7: Eigenvalues/SelfAdjointEigenSolver.h compute(),
// Tridiagonalization.h tridiagonalization_inplace, tridiagonal_qr_step, Jacobi/Jacobi.h makeGivens,
// MathFunctions.h hypot_impl).  Used by ContourView::calcStatVals, include/cont2/contour.h:165-172.
// Input: the UPPER triangle (a = m(0,0), b = m(0,1), c = m(1,1)).  Output: eigenvalues ascending, eigenvectors as
// columns in column-major order [v(0,0), v(1,0), v(0,1), v(1,1)].
// ---------------------------------------------------------------------------------------------------------------
struct Givens {
  float c, s;
};
inline Givens makeGivens(float p, float q) {  // JacobiRotation<float>::makeGivens(p, q, 0, false_type)
  Givens g;
  if (q == 0.0f) {
    g.c = p < 0.0f ? -1.0f : 1.0f;
    g.s = 0.0f;
  } else if (p == 0.0f) {
    g.c = 0.0f;
    g.s = q < 0.0f ? 1.0f : -1.0f;
  } else if (std::abs(p) > std::abs(q)) {
    float t = q / p;
    float u = std::sqrt(1.0f + t * t);
    if (p < 0.0f) u = -u;
    g.c = 1.0f / u;
    g.s = -t * g.c;
  } else {
    float t = p / q;
    float u = std::sqrt(1.0f + t * t);
    if (q < 0.0f) u = -u;
    g.s = -1.0f / u;
    g.c = -t * g.s;
  }
  return g;
}
inline float eigen_hypot(float x, float y) {  // numext::hypot -> hypot_impl<float>::run
  float ax = std::abs(x), ay = std::abs(y);
  float p, qp;
  if (ax > ay) {
    p = ax;
    qp = ay / p;
  } else {
    p = ay;
    qp = ax / p;
  }
  if (p == 0.0f) return 0.0f;
  return p * std::sqrt(1.0f + qp * qp);
}

inline void selfAdjointEigen2f(float a, float b, float c, float evals[2], float evecs[4]) {
  // compute(): mat = lower triangle (mirrored from the selfadjoint view), scaled into [-1, 1]
  float scale = std::max(std::max(std::abs(a), std::abs(b)), std::abs(c));
  // mat's strict upper part is zero, so it does not change maxCoeff of cwiseAbs
  if (scale == 0.0f) scale = 1.0f;
  float diag[2] = {a / scale, c / scale};
  float sub = b / scale;
  // tridiagonalization_inplace on a 2x2: the Householder vector has an empty tail => tau = 0, beta = sub, Q = I
  // (Householder.h makeHouseholder: tailSqNorm == 0 branch); the rank update adds -(1*0 + 0*1) = -0 to diag[1].
  {
    float h0 = diag[1] * (0.0f * 1.0f);  // hCoeffs = A22 * (conj(h) * v), h = 0, v = 1
    h0 += (0.0f * -0.5f * (h0 * 1.0f)) * 1.0f;
    diag[1] = diag[1] + (-1.0f) * (1.0f * h0 + h0 * 1.0f);  // selfadjoint rankUpdate(u = v, v = hCoeffs, alpha = -1)
  }
  float q[4] = {1.0f, 0.0f, 0.0f, 1.0f};  // column-major 2x2 identity

  // computeFromTridiagonal_impl, n = 2, maxIterations = 30
  const int n = 2;
  int end = n - 1, start = 0, iter = 0;
  const float considerAsZero = std::numeric_limits<float>::min();
  const float precision = 2.0f * std::numeric_limits<float>::epsilon();
  bool ok = true;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      // internal::isMuchSmallerThan(|sub|, |d_i| + |d_{i+1}|, precision) : |x| <= |y| * prec
      if (std::abs(sub) <= (std::abs(diag[i]) + std::abs(diag[i + 1])) * precision || std::abs(sub) <= considerAsZero)
        sub = 0.0f;
    }
    while (end > 0 && sub == 0.0f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * n) {
      ok = false;
      break;
    }
    start = end - 1;  // n == 2: start = 0
    // tridiagonal_qr_step(diag, subdiag, start=0, end=1, Q, n)
    float td = (diag[end - 1] - diag[end]) * 0.5f;
    float e = sub;
    float mu = diag[end];
    if (td == 0.0f) {
      mu -= std::abs(e);
    } else {
      float e2 = e * e;
      float h = eigen_hypot(td, e);
      if (e2 == 0.0f)
        mu -= (e / (td + (td > 0.0f ? 1.0f : -1.0f))) * (e / h);
      else
        mu -= e2 / (td + (td > 0.0f ? h : -h));
    }
    float x = diag[start] - mu;
    float z = sub;
    {  // k = start = 0 only
      Givens rot = makeGivens(x, z);
      float sdk = rot.s * diag[0] + rot.c * sub;
      float dkp1 = rot.s * sub + rot.c * diag[1];
      diag[0] = rot.c * (rot.c * diag[0] - rot.s * sub) - rot.s * (rot.c * sub - rot.s * diag[1]);
      diag[1] = rot.s * sdk + rot.c * dkp1;
      sub = rot.c * sdk - rot.s * dkp1;
      // q.applyOnTheRight(0, 1, rot): apply_rotation_in_the_plane(col0, col1, rot.transpose()) with (c, -s):
      //   x_i = c*x_i + (-s)*y_i ; y_i = -(-s)*x_i + c*y_i
      for (int i = 0; i < 2; ++i) {
        float xi = q[i], yi = q[2 + i];
        q[i] = rot.c * xi + (-rot.s) * yi;
        q[2 + i] = rot.s * xi + rot.c * yi;
      }
    }
  }
  if (ok) {  // sort ascending (n = 2: swap if diag[1] < diag[0]; minCoeff picks the first minimum)
    if (diag[1] < diag[0]) {
      std::swap(diag[0], diag[1]);
      std::swap(q[0], q[2]);
      std::swap(q[1], q[3]);
    }
  }
  evals[0] = diag[0] * scale;
  evals[1] = diag[1] * scale;
  for (int i = 0; i < 4; ++i) evecs[i] = q[i];
}

// ---------------------------------------------------------------------------------------------------------------
// include/cont2/contour.h
// ---------------------------------------------------------------------------------------------------------------
// RunningStatRecorder (include/cont2/contour.h:48-95); runningStatsF :74-84
struct RunningStat {
  int16_t cell_cnt = 0;
  double pos_sum[2] = {0, 0};
  double pos_tss[4] = {0, 0, 0, 0};  // column-major v v^T
  float vol3 = 0.0f;
  double torq[2] = {0, 0};
  void runningStatsF(float curr_row, float curr_col, float height) {
    cell_cnt += 1;
    double v0 = curr_row, v1 = curr_col;
    pos_sum[0] += v0;
    pos_sum[1] += v1;
    pos_tss[0] += v0 * v0;
    pos_tss[1] += v1 * v0;
    pos_tss[2] += v0 * v1;
    pos_tss[3] += v1 * v1;
    vol3 += height;
    torq[0] += height * v0;  // float * double -> double
    torq[1] += height * v1;
  }
};

// ContourView::calcStatVals (include/cont2/contour.h:142-255), eccentricitySalient :258-260,
// centerOfMassSalient :263-265.
inline void calcStatVals(c2g_view &v, const RunningStat &rec, const c2g_cm_config &cfg) {
  v.cell_cnt = rec.cell_cnt;
  const float cntf = (float) v.cell_cnt;
  v.pos_mean[0] = (float) rec.pos_sum[0] / cntf;
  v.pos_mean[1] = (float) rec.pos_sum[1] / cntf;
  v.vol3_mean = rec.vol3 / cntf;
  v.com[0] = (float) rec.torq[0] / rec.vol3;
  v.com[1] = (float) rec.torq[1] / rec.vol3;
  v.eccen = 0.0f;
  if (v.cell_cnt < cfg.min_cell_cov) {
    const float s2 = 1.0f * cfg.point_sigma * cfg.point_sigma;  // Identity * sigma * sigma
    v.pos_cov[0] = s2;
    v.pos_cov[1] = 0.0f * cfg.point_sigma * cfg.point_sigma;
    v.pos_cov[2] = 0.0f * cfg.point_sigma * cfg.point_sigma;
    v.pos_cov[3] = s2;
    v.eig_vals[0] = cfg.point_sigma;
    v.eig_vals[1] = cfg.point_sigma;
    v.eig_vecs[0] = 1.0f;
    v.eig_vecs[1] = 0.0f;
    v.eig_vecs[2] = 0.0f;
    v.eig_vecs[3] = 1.0f;
    v.ecc_feat = 0;
    v.com_feat = 0;
  } else {
    // pos_cov = (tss.cast<float>() - pos_mean * pos_mean^T * cnt) / (cnt - 1)
    const float cm1 = (float) (v.cell_cnt - 1);
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 2; ++i) {
        float outer = v.pos_mean[i] * v.pos_mean[j];
        v.pos_cov[j * 2 + i] = ((float) rec.pos_tss[j * 2 + i] - outer * cntf) / cm1;
      }
    selfAdjointEigen2f(v.pos_cov[0], v.pos_cov[2], v.pos_cov[3], v.eig_vals, v.eig_vecs);
    if (v.eig_vals[0] < cfg.point_sigma) v.eig_vals[0] = cfg.point_sigma;
    if (v.eig_vals[1] < cfg.point_sigma) v.eig_vals[1] = cfg.point_sigma;
    v.eccen = std::sqrt(v.eig_vals[1] * v.eig_vals[1] - v.eig_vals[0] * v.eig_vals[0]) / v.eig_vals[1];
    v.ecc_feat = (v.cell_cnt > 5 && diff_perc<float>(v.eig_vals[0], v.eig_vals[1], 0.2f) && v.eig_vals[1] > 2.5f) ? 1 : 0;
    const float dx = v.com[0] - v.pos_mean[0], dy = v.com[1] - v.pos_mean[1];
    v.com_feat = (std::sqrt(dx * dx + dy * dy) > cfg.com_bias_thres) ? 1 : 0;
  }
}

// ContourView::checkSim (include/cont2/contour.h:278-329)
inline bool checkSim(const c2g_view &s, const c2g_view &t, const c2g_sim_config &th) {
  if (diff_perc<float>(s.cell_cnt, t.cell_cnt, th.tp_cell_cnt) && diff_delt<float>(s.cell_cnt, t.cell_cnt, th.ta_cell_cnt))
    return false;
  if (std::max(s.eig_vals[1], t.eig_vals[1]) > 2.0 &&
      diff_perc<float>(std::sqrt(s.eig_vals[1]), std::sqrt(t.eig_vals[1]), th.tp_eigval))
    return false;
  if (std::max(s.eig_vals[0], t.eig_vals[0]) > 2.0 &&
      diff_perc<float>(std::sqrt(s.eig_vals[0]), std::sqrt(t.eig_vals[0]), th.tp_eigval))
    return false;
  if (std::max(s.cell_cnt, t.cell_cnt) > 15 && diff_delt<float>(s.vol3_mean, t.vol3_mean, th.ta_h_bar)) return false;
  const float sx = s.com[0] - s.pos_mean[0], sy = s.com[1] - s.pos_mean[1];
  const float tx = t.com[0] - t.pos_mean[0], ty = t.com[1] - t.pos_mean[1];
  const float com_r1 = std::sqrt(sx * sx + sy * sy);
  const float com_r2 = std::sqrt(tx * tx + ty * ty);
  if (diff_delt<float>(com_r1, com_r2, th.ta_rcom) && diff_perc<float>(com_r1, com_r2, th.tp_rcom)) return false;
  return true;
}

// ContourView::getManualCov (include/cont2/contour.h:376-378): V * diag(lambda) * V^T in float, column-major.
inline void getManualCov(const c2g_view &v, float out[4]) {
  // (V * D)(i,k) = V(i,k) * lambda_k ; result(i,j) = sum_k (V D)(i,k) * V(j,k)
  float vd[4];
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 2; ++i) vd[k * 2 + i] = v.eig_vecs[k * 2 + i] * v.eig_vals[k];
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < 2; ++i) out[j * 2 + i] = vd[0 * 2 + i] * v.eig_vecs[0 * 2 + j] + vd[1 * 2 + i] * v.eig_vecs[1 * 2 + j];
}

// ---------------------------------------------------------------------------------------------------------------
// OpenCV connectedComponentsWithStats(mask, labels, stats, centroids, 8, CV_32S) restated for label ORDER and stats.
// OpenCV's 8-connectivity algorithms (BBDT / Spaghetti) scan the image in 2x2 blocks, blocks aligned to the ROI
// origin, and flatten provisional labels in creation order, so final label k belongs to the component whose first
// 2x2 block comes k-th in block-raster order (verified against cv2 in tests/test_oracle_ccl_cv2.py).
// ---------------------------------------------------------------------------------------------------------------
struct CCStat {
  int left, top, width, height, area;
};
// mask: h x w, row-major, nonzero = foreground. labels: h x w int32 out (0 = background). returns stats[1..n]
inline std::vector<CCStat> connectedComponents8(const uint8_t *mask, int h, int w, std::vector<int> &labels) {
  labels.assign((size_t) h * w, 0);
  struct Tmp {
    int key, minr, maxr, minc, maxc, area, seed;
  };
  std::vector<Tmp> comps;
  std::vector<int> stack;
  std::vector<int> tmp_label((size_t) h * w, -1);
  const int bw = (w + 1) / 2;
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < w; ++c) {
      if (!mask[r * w + c] || tmp_label[r * w + c] >= 0) continue;
      const int id = (int) comps.size();
      Tmp t{(r / 2) * bw + (c / 2), r, r, c, c, 0, r * w + c};
      stack.clear();
      stack.push_back(r * w + c);
      tmp_label[r * w + c] = id;
      while (!stack.empty()) {
        int p = stack.back();
        stack.pop_back();
        int pr = p / w, pc = p % w;
        t.area++;
        t.minr = std::min(t.minr, pr);
        t.maxr = std::max(t.maxr, pr);
        t.minc = std::min(t.minc, pc);
        t.maxc = std::max(t.maxc, pc);
        t.key = std::min(t.key, (pr / 2) * bw + (pc / 2));
        for (int dr = -1; dr <= 1; ++dr)
          for (int dc = -1; dc <= 1; ++dc) {
            int nr = pr + dr, nc = pc + dc;
            if (nr < 0 || nr >= h || nc < 0 || nc >= w) continue;
            if (mask[nr * w + nc] && tmp_label[nr * w + nc] < 0) {
              tmp_label[nr * w + nc] = id;
              stack.push_back(nr * w + nc);
            }
          }
      }
      comps.push_back(t);
    }
  std::vector<int> order(comps.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int) i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return comps[a].key < comps[b].key; });
  std::vector<int> final_of(comps.size());
  std::vector<CCStat> stats(comps.size() + 1);
  stats[0] = CCStat{0, 0, w, h, 0};
  for (size_t k = 0; k < order.size(); ++k) {
    const Tmp &t = comps[order[k]];
    final_of[order[k]] = (int) k + 1;
    stats[k + 1] = CCStat{t.minc, t.minr, t.maxc - t.minc + 1, t.maxr - t.minr + 1, t.area};
  }
  for (size_t i = 0; i < labels.size(); ++i)
    if (tmp_label[i] >= 0) labels[i] = final_of[tmp_label[i]];
  return stats;
}

// ---------------------------------------------------------------------------------------------------------------
// ContourManager (include/cont2/contour_mng.h:414-1314, src/cont2/contour_mng.cpp:274-353)
// ---------------------------------------------------------------------------------------------------------------
struct Pixelf {  // include/cont2/contour_mng.h:392-411
  float row_f, col_f, elev;
};

struct Scan {
  c2g_cm_config cfg;
  int int_id = 0;
  float x_max, x_min, y_max, y_min;
  std::vector<float> bev;                           // n_row * n_col, init -1000 (contour_mng.h:488)
  std::vector<std::pair<int, Pixelf>> bev_pixfs;    // sorted by hash (contour_mng.h:435,528-529)
  float max_bin_val = -1e3f, min_bin_val = 1e3f;
  std::vector<std::vector<c2g_view>> cont_views;    // [level][seq] after the sort
  std::vector<std::vector<c2g_view>> presort_views; // [level][dfs order] before the sort (debug / parity)
  std::vector<std::vector<float>> cont_perc;
  std::vector<int> layer_cell_cnt;
  std::vector<std::vector<std::array<float, C2G_KEY_DIM>>> layer_keys;
  std::vector<std::vector<c2g_bci>> layer_key_bcis;

  Scan(const c2g_cm_config &c, int id) : cfg(c), int_id(id) {  // contour_mng.h:478-498
    x_min = -(cfg.n_row / 2) * cfg.reso_row;
    x_max = -x_min;
    y_min = -(cfg.n_col / 2) * cfg.reso_col;
    y_max = -y_min;
    bev.assign((size_t) cfg.n_row * cfg.n_col, -1e3f);
    cont_views.resize(cfg.n_levels);
    presort_views.resize(cfg.n_levels);
    cont_perc.resize(cfg.n_levels);
    layer_cell_cnt.assign(cfg.n_levels, 0);
    layer_keys.resize(cfg.n_levels);
    layer_key_bcis.resize(cfg.n_levels);
  }

  // hashPointToImage (contour_mng.h:448-463)
  std::pair<int, int> hashPointToImage(float px, float py) const {
    std::pair<int, int> res{-1, -1};
    float padding = 1e-2;
    if (px < x_min + padding || px > x_max - padding || py < y_min + padding || py > y_max - padding ||
        (py * py + px * px) < cfg.blind_sq)
      return res;
    // NOTE: for NaN coordinates the reference converts floor(NaN) to int (undefined behaviour; x86 yields INT_MIN and
    // the `rc.first > 0` test drops a NaN x, while a NaN y alone indexes out of bounds).  This restatement drops every
    // point with a NaN x or y.
    if (std::isnan(px) || std::isnan(py)) return res;
    res.first = int(std::floor(px / cfg.reso_row)) + cfg.n_row / 2;
    res.second = int(std::floor(py / cfg.reso_col)) + cfg.n_col / 2;
    return res;
  }

  // makeBEV (contour_mng.h:505-556). pts: n x 4 float (x, y, z, intensity), the KITTI .bin layout
  // (tools/pointcloud_util.h:12-50 keeps x, y, z).
  void makeBEV(const float *pts, int n) {
    std::map<int, Pixelf> tmp_pillars;
    for (int i = 0; i < n; ++i) {
      const float px = pts[4 * i + 0], py = pts[4 * i + 1], pz = pts[4 * i + 2];
      std::pair<int, int> rc = hashPointToImage(px, py);
      if (rc.first > 0) {
        float height = cfg.lidar_height + pz;
        float &cell = bev[(size_t) rc.first * cfg.n_col + rc.second];
        if (cell < height) {
          cell = height;
          // pointToContRowCol (contour_mng.h:468-472): x / reso + n_row / 2 - 0.5f, evaluated left to right in float
          float rf = px / cfg.reso_row + cfg.n_row / 2 - 0.5f;
          float cf = py / cfg.reso_col + cfg.n_col / 2 - 0.5f;
          tmp_pillars[rc.first * cfg.n_col + rc.second] = Pixelf{rf, cf, height};
        }
        max_bin_val = max_bin_val < height ? height : max_bin_val;
        min_bin_val = min_bin_val > height ? height : min_bin_val;
      }
    }
    bev_pixfs.clear();
    bev_pixfs.insert(bev_pixfs.begin(), tmp_pillars.begin(), tmp_pillars.end());
  }

  // search_vec (tools/algos.h:58-68) — binary search on the sorted pillar list
  const Pixelf *searchPix(int hash) const {
    int p1 = 0, p2 = (int) bev_pixfs.size() - 1;
    while (p2 >= p1) {
      int mid = (p1 + p2) / 2;
      if (bev_pixfs[mid].first == hash) return &bev_pixfs[mid].second;
      if (bev_pixfs[mid].first < hash)
        p1 = mid + 1;
      else
        p2 = mid - 1;
    }
    return nullptr;
  }

  // makeContourRecursiveHelper (src/cont2/contour_mng.cpp:274-353).
  // roi = (x, y, w, h) on the BEV; mask = h x w (nonzero = inside the parent component); level-0 call ignores mask.
  void makeContourRecursiveHelper(int rx, int ry, int rw, int rh, const std::vector<uint8_t> &cc_mask, int level) {
    if (level >= cfg.n_levels) return;
    const float h_min = cfg.lv_grads[level];
    std::vector<uint8_t> bin((size_t) rw * rh);
    for (int i = 0; i < rh; ++i)
      for (int j = 0; j < rw; ++j) {
        uint8_t b = bev[(size_t) (ry + i) * cfg.n_col + (rx + j)] > h_min ? 255 : 0;  // cv::threshold BINARY (strict >)
        if (level) b = b & cc_mask[(size_t) i * rw + j];                              // cv::bitwise_and(bin, cc_mask)
        bin[(size_t) i * rw + j] = b;
      }
    std::vector<int> labels;
    std::vector<CCStat> stats = connectedComponents8(bin.data(), rh, rw, labels);
    for (int n = 1; n < (int) stats.size(); ++n) {
      if (stats[n].area < cfg.min_cont_cell_cnt) continue;
      const int gx = stats[n].left + rx, gy = stats[n].top + ry, w = stats[n].width, h = stats[n].height;
      std::vector<uint8_t> mask_n((size_t) w * h);
      for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
          mask_n[(size_t) i * w + j] = labels[(size_t) (stats[n].top + i) * rw + (stats[n].left + j)] == n ? 255 : 0;
      RunningStat rec;
      int poi_r = -1, poi_c = -1;
      for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
          if (mask_n[(size_t) i * w + j]) {
            poi_r = i + gy;
            poi_c = j + gx;
            const Pixelf *px = searchPix(poi_r * cfg.n_col + poi_c);
            rec.runningStatsF(px->row_f, px->col_f, bev[(size_t) poi_r * cfg.n_col + poi_c]);
          }
      c2g_view v;
      std::memset(&v, 0, sizeof(v));
      v.level = (int16_t) level;
      v.poi_r = (int16_t) poi_r;
      v.poi_c = (int16_t) poi_c;
      calcStatVals(v, rec, cfg);
      cont_views[level].push_back(v);
      makeContourRecursiveHelper(gx, gy, w, h, mask_n, level + 1);
    }
  }

  // makeContoursRecurs (contour_mng.h:588-960)
  void makeContoursRecurs() {
    makeContourRecursiveHelper(0, 0, cfg.n_col, cfg.n_row, std::vector<uint8_t>(1, 0), 0);
    presort_views = cont_views;
    for (int ll = 0; ll < cfg.n_levels; ++ll) {
      // contour_mng.h:596-599: the real libstdc++ std::sort (unstable) with the reference comparator
      std::sort(cont_views[ll].begin(), cont_views[ll].end(),
                [](const c2g_view &p1, const c2g_view &p2) { return p1.cell_cnt > p2.cell_cnt; });
      layer_cell_cnt[ll] = 0;
      for (auto &v : cont_views[ll]) layer_cell_cnt[ll] += v.cell_cnt;
      cont_perc[ll].clear();
      for (auto &v : cont_views[ll]) cont_perc[ll].push_back(v.cell_cnt * 1.0f / layer_cell_cnt[ll]);
    }

    const int roi_radius_padded = (int) std::ceil(cfg.roi_radius + 1);
    const int DIST_BIN_LAYERS[4] = {1, 2, 3, 4};
    for (int ll = 0; ll < cfg.n_levels; ++ll) {
      int accumulate_cell_cnt = 0;
      for (int seq = 0; seq < cfg.piv_firsts; ++seq) {
        std::array<float, C2G_KEY_DIM> key;
        key.fill(0.0f);
        c2g_bci bci;
        std::memset(&bci, 0, sizeof(bci));
        bci.piv_seq = (int8_t) seq;
        bci.level = (int8_t) ll;
        std::vector<c2g_relpt> nei_pts;
        std::vector<uint16_t> segs;

        if ((int) cont_views[ll].size() > seq) accumulate_cell_cnt += cont_views[ll][seq].cell_cnt;

        if ((int) cont_views[ll].size() > seq && cont_views[ll][seq].cell_cnt >= cfg.min_cont_key_cnt) {
          const c2g_view &anchor = cont_views[ll][seq];
          const float cen_x = anchor.pos_mean[0], cen_y = anchor.pos_mean[1];
          int r_cen = int(cen_x), c_cen = int(cen_y);
          int r_min = std::max(0, r_cen - roi_radius_padded), r_max = std::min(cfg.n_row - 1, r_cen + roi_radius_padded);
          int c_min = std::max(0, c_cen - roi_radius_padded), c_max = std::min(cfg.n_col - 1, c_cen + roi_radius_padded);

          const int num_bins = C2G_KEY_DIM - 3;
          float bin_len = cfg.roi_radius / num_bins;
          std::vector<float> ring_bins(num_bins, 0);
          const int div_per_bin = 5;
          std::vector<float> discrete_divs(div_per_bin * num_bins, 0);
          float div_len = cfg.roi_radius / (num_bins * div_per_bin);
          int cnt_point = 0;

          for (int rr = r_min; rr <= r_max; rr++) {
            for (int cc = c_min; cc <= c_max; cc++) {
              const float bv = bev[(size_t) rr * cfg.n_col + cc];
              if (bv < cfg.lv_grads[DIST_BIN_LAYERS[0]]) continue;
              const Pixelf *px = searchPix(rr * cfg.n_col + cc);
              const float dx = px->row_f - cen_x, dy = px->col_f - cen_y;
              float dist = std::sqrt(dx * dx + dy * dy);
              if (dist < cfg.roi_radius - 1e-2 && bv > cfg.lv_grads[DIST_BIN_LAYERS[0]]) {
                int higher_cnt = 0;
                for (int ele = DIST_BIN_LAYERS[0]; ele < cfg.n_levels; ele++)
                  if (bv > cfg.lv_grads[ele]) higher_cnt++;
                cnt_point++;
                for (int div_idx = 0; div_idx < num_bins * div_per_bin; div_idx++)
                  discrete_divs[div_idx] += higher_cnt * gaussPDF<float>(div_idx * div_len + 0.5 * div_len, dist, 1.0);
              }
            }
          }
          for (int b = 0; b < num_bins; b++) {
            for (int d = 0; d < div_per_bin; d++) ring_bins[b] += discrete_divs[b * div_per_bin + d];
            ring_bins[b] *= bin_len / std::sqrt(cnt_point);
          }
          key[0] = std::sqrt(anchor.eig_vals[1] * anchor.cell_cnt);
          key[1] = std::sqrt(anchor.eig_vals[0] * anchor.cell_cnt);
          key[2] = std::sqrt(accumulate_cell_cnt);
          for (int nb = 0; nb < num_bins; nb++) key[3 + nb] = ring_bins[nb];

          // BCI (contour_mng.h:848-883)
          for (int bl = 0; bl < C2G_NUM_BIN_LAYERS; bl++) {
            int bit_offset = bl * C2G_BITS_PER_LAYER;
            const auto &lay = cont_views[DIST_BIN_LAYERS[bl]];
            for (int j = 0; j < std::min(cfg.dist_firsts, (int) lay.size()); j++) {
              if (ll != DIST_BIN_LAYERS[bl] || j != seq) {
                const float vx = lay[j].pos_mean[0] - anchor.pos_mean[0];
                const float vy = lay[j].pos_mean[1] - anchor.pos_mean[1];
                float tmp_dist = std::sqrt(vx * vx + vy * vy);
                if (tmp_dist > (C2G_BITS_PER_LAYER - 1) * 1.01 + 5.43 - 1e-3 || tmp_dist <= 5.43) continue;
                float tmp_orie = std::atan2(vy, vx);  // std::atan2(float, float) -> atan2f
                int dist_idx = std::min(std::floor((tmp_dist - 5.43) / 1.01), C2G_BITS_PER_LAYER - 1.0) + bit_offset;
                bci.dist_bin[dist_idx >> 6] |= (uint64_t) 1 << (dist_idx & 63);
                c2g_relpt rp;
                rp.level = (int8_t) DIST_BIN_LAYERS[bl];
                rp.seq = (int8_t) j;
                rp.bit_pos = (int16_t) dist_idx;
                rp.r = tmp_dist;
                rp.theta = tmp_orie;
                nei_pts.push_back(rp);
              }
            }
          }
          if (!nei_pts.empty()) {
            std::sort(nei_pts.begin(), nei_pts.end(),
                      [](const c2g_relpt &p1, const c2g_relpt &p2) { return p1.bit_pos < p2.bit_pos; });
            segs.push_back(0);
            for (int p1 = 0; p1 < (int) nei_pts.size(); p1++)
              if (nei_pts[segs.back()].bit_pos != nei_pts[p1].bit_pos) segs.push_back((uint16_t) p1);
            segs.push_back((uint16_t) nei_pts.size());
          }
        }
        bci.n_nei = (int16_t) nei_pts.size();
        bci.n_seg = (int16_t) segs.size();
        for (size_t i = 0; i < nei_pts.size(); ++i) bci.nei[i] = nei_pts[i];
        for (size_t i = 0; i < segs.size(); ++i) bci.seg[i] = segs[i];
        layer_key_bcis[ll].push_back(bci);
        layer_keys[ll].push_back(key);
      }
    }
  }

  float getAreaPerc(int lev, int seq) const { return cont_perc[lev][seq]; }
};

}  // namespace c2o
