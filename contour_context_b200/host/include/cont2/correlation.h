// cont2/correlation.h (facade) — GMMOptConfig and the static helpers of ConstellCorrelation that the harness uses
// (reference include/cont2/correlation.h:15-20,241-296).  The GMM-L2 correlation and its refinement run in query.cu /
// refine.cu.
#pragma once
#include <cmath>
#include <vector>

#include "cont2/contour_mng.h"

struct GMMOptConfig {
  double min_area_perc_ = 0.95;
  std::vector<int> levels_ = {1, 2, 3, 4};
  double cov_dilate_scale_ = 2.0;
};

// Rigid 3-D pose (what the reference keeps as Eigen::Isometry3d in the evaluator): rotation row-major + translation.
struct C2gPose3 {
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double t[3] = {0, 0, 0};
  // 12 numbers of a row-major 3x4 [R | t]; like evaluator.h:100-103 the rotation passes through a quaternion
  // (Eigen::Quaterniond(Matrix3d) then Transform::rotate(q), i.e. q.toRotationMatrix())
  static C2gPose3 fromRowMajor3x4(const double *v) {
    double m[3][3];
    C2gPose3 P;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) m[r][c] = v[r * 4 + c];
      P.t[r] = v[r * 4 + 3];
    }
    double q[4];  // w x y z
    double tr = m[0][0] + m[1][1] + m[2][2];
    if (tr > 0) {
      tr = std::sqrt(tr + 1.0);
      q[0] = 0.5 * tr;
      tr = 0.5 / tr;
      q[1] = (m[2][1] - m[1][2]) * tr;
      q[2] = (m[0][2] - m[2][0]) * tr;
      q[3] = (m[1][0] - m[0][1]) * tr;
    } else {
      int i = 0;
      if (m[1][1] > m[0][0]) i = 1;
      if (m[2][2] > m[i][i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      tr = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
      q[1 + i] = 0.5 * tr;
      tr = 0.5 / tr;
      q[0] = (m[k][j] - m[j][k]) * tr;
      q[1 + j] = (m[j][i] + m[i][j]) * tr;
      q[1 + k] = (m[k][i] + m[i][k]) * tr;
    }
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    P.R[0][0] = 1 - (tyy + tzz);
    P.R[0][1] = txy - twz;
    P.R[0][2] = txz + twy;
    P.R[1][0] = txy + twz;
    P.R[1][1] = 1 - (txx + tzz);
    P.R[1][2] = tyz - twx;
    P.R[2][0] = txz - twy;
    P.R[2][1] = tyz + twx;
    P.R[2][2] = 1 - (txx + tyy);
    return P;
  }
  // this^-1 * o
  C2gPose3 inverseTimes(const C2gPose3 &o) const {
    C2gPose3 r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) r.R[i][j] = R[0][i] * o.R[0][j] + R[1][i] * o.R[1][j] + R[2][i] * o.R[2][j];
      r.t[i] = R[0][i] * (o.t[0] - t[0]) + R[1][i] * (o.t[1] - t[1]) + R[2][i] * (o.t[2] - t[2]);
    }
    return r;
  }
};

class ConstellCorrelation {
 public:
  // evalMetricEst (reference include/cont2/correlation.h:241-280): error transform T_gt^-1 * T_est between the estimated and the
  // ground-truth pose of the source sensor in the target sensor frame, both projected to SE(2)
  static Eigen::Isometry2d evalMetricEst(const Eigen::Isometry2d &T_delta, const C2gPose3 &gt_src_3d, const C2gPose3 &gt_tgt_3d,
                                         const ContourManagerConfig &bev_config) {
    Eigen::Isometry2d T_est = getEstSensTF(T_delta, bev_config);
    T_est(0, 2) *= bev_config.reso_row_;
    T_est(1, 2) *= bev_config.reso_row_;
    const C2gPose3 T3 = gt_tgt_3d.inverseTimes(gt_src_3d);
    // rotate so that the two z axes align: axis = z0 x z1 (normalised when non-zero), angle = -acos(z0 . z1)
    double ax[3] = {-T3.R[1][2], T3.R[0][2], 0.0};  // (0,0,1) x (z1x, z1y, z1z)
    const double n2 = ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2];
    if (n2 > 0) {
      const double n = std::sqrt(n2);
      ax[0] /= n;
      ax[1] /= n;
      ax[2] /= n;
    }
    const double ang = -std::acos(T3.R[2][2]);
    const double c = std::cos(ang), s = std::sin(ang);
    double D[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) D[i][j] = (1 - c) * ax[i] * ax[j] + (i == j ? c : 0.0);
    D[0][1] += -s * ax[2];
    D[0][2] += s * ax[1];
    D[1][0] += s * ax[2];
    D[1][2] += -s * ax[0];
    D[2][0] += -s * ax[1];
    D[2][1] += s * ax[0];
    const double r00 = D[0][0] * T3.R[0][0] + D[0][1] * T3.R[1][0] + D[0][2] * T3.R[2][0];
    const double r10 = D[1][0] * T3.R[0][0] + D[1][1] * T3.R[1][0] + D[1][2] * T3.R[2][0];
    Eigen::Isometry2d T_gt;
    T_gt.setIdentity();
    T_gt.rotate(std::atan2(r10, r00));
    T_gt.pretranslate(V2D(T3.t[0], T3.t[1]));
    return T_gt.inverse() * T_est;
  }

  static Eigen::Isometry2d getEstSensTF(const Eigen::Isometry2d &T_delta, const ContourManagerConfig &bev_config) {
    Eigen::Isometry2d T_so_ssen = Eigen::Isometry2d::Identity(), T_to_tsen;
    T_so_ssen.pretranslate(V2D(bev_config.n_row_ / 2 - 0.5, bev_config.n_col_ / 2 - 0.5));
    T_to_tsen = T_so_ssen;
    return T_to_tsen.inverse() * T_delta * T_so_ssen;
  }
};
