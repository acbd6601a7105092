// cont2/correlation.h (facade) — GMMOptConfig and the static helpers of ConstellCorrelation that the harness uses
// (reference include/cont2/correlation.h:15-20,287-296).  The GMM-L2 correlation itself runs in query.cu.
#pragma once
#include <vector>

#include "cont2/contour_mng.h"

struct GMMOptConfig {
  double min_area_perc_ = 0.95;
  std::vector<int> levels_ = {1, 2, 3, 4};
  double cov_dilate_scale_ = 2.0;
};

class ConstellCorrelation {
 public:
  static Eigen::Isometry2d getEstSensTF(const Eigen::Isometry2d &T_delta, const ContourManagerConfig &bev_config) {
    Eigen::Isometry2d T_so_ssen = Eigen::Isometry2d::Identity(), T_to_tsen;
    T_so_ssen.pretranslate(V2D(bev_config.n_row_ / 2 - 0.5, bev_config.n_col_ / 2 - 0.5));
    T_to_tsen = T_so_ssen;
    return T_to_tsen.inverse() * T_delta * T_so_ssen;
  }
};
