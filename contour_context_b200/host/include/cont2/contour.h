// cont2/contour.h (facade) — same type and member names as the reference's include/cont2/contour.h for everything the
// cont2_batch_bin_test loop reads.  The statistics are NOT computed here: they are filled from the c2g_view records the
// GPU path produces (contour_context_b200/csrc/contours.cu restates ContourView::calcStatVals, contour.h:142-255).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "cont2/c2g_shims.h"
#include "c2g_types.h"

struct ContourViewStatConfig {  // reference include/cont2/contour.h:32-37
  int16_t min_cell_cov = 4;
  float point_sigma = 1.0;
  float com_bias_thres = 0.5;
};

struct ContourSimThresConfig {  // reference include/cont2/contour.h:40-45
  float ta_cell_cnt = 6, tp_cell_cnt = 0.2;
  float tp_eigval = 0.2;
  float ta_h_bar = 0.3;
  float ta_rcom = 0.4, tp_rcom = 0.25;
};

struct ContourView {  // reference include/cont2/contour.h:97-119
  int16_t level_;
  int16_t poi_[2];
  int16_t cell_cnt_{};
  V2F pos_mean_;
  M2F pos_cov_;
  V2F eig_vals_;
  M2F eig_vecs_;
  float eccen_{};
  float vol3_mean_{};
  V2F com_;
  bool ecc_feat_ = false;
  bool com_feat_ = false;

  ContourView() : level_(0) { poi_[0] = poi_[1] = 0; }
  explicit ContourView(const c2g_view &v) : level_(v.level), cell_cnt_(v.cell_cnt), eccen_(v.eccen), vol3_mean_(v.vol3_mean) {
    poi_[0] = v.poi_r;
    poi_[1] = v.poi_c;
    pos_mean_ = V2F(v.pos_mean[0], v.pos_mean[1]);
    eig_vals_ = V2F(v.eig_vals[0], v.eig_vals[1]);
    com_ = V2F(v.com[0], v.com[1]);
    for (int c = 0; c < 2; ++c)
      for (int r = 0; r < 2; ++r) {
        pos_cov_(r, c) = v.pos_cov[c * 2 + r];
        eig_vecs_(r, c) = v.eig_vecs[c * 2 + r];
      }
    ecc_feat_ = v.ecc_feat != 0;
    com_feat_ = v.com_feat != 0;
  }

  // reference include/cont2/contour.h:278-329 (host restatement; the device runs the same tests in csrc/query.cu:check_sim):
  // cell count (relative AND absolute), both eigenvalue roots when the larger one exceeds 2, mean height for contours of more
  // than 15 cells, centre-of-mass offset radius (absolute AND relative)
  static bool checkSim(const ContourView &cont_src, const ContourView &cont_tgt, const ContourSimThresConfig &simthres) {
    auto diff_perc = [](float a, float b, float perc) { return std::fabs((a - b) / std::max(a, b)) > perc; };
    auto diff_delt = [](float a, float b, float delta) { return std::fabs(a - b) > delta; };
    const float sc = cont_src.cell_cnt_, tc = cont_tgt.cell_cnt_;
    if (diff_perc(sc, tc, simthres.tp_cell_cnt) && diff_delt(sc, tc, simthres.ta_cell_cnt)) return false;
    if (std::max(cont_src.eig_vals_(1), cont_tgt.eig_vals_(1)) > 2.0 &&
        diff_perc(std::sqrt(cont_src.eig_vals_(1)), std::sqrt(cont_tgt.eig_vals_(1)), simthres.tp_eigval))
      return false;
    if (std::max(cont_src.eig_vals_(0), cont_tgt.eig_vals_(0)) > 2.0 &&
        diff_perc(std::sqrt(cont_src.eig_vals_(0)), std::sqrt(cont_tgt.eig_vals_(0)), simthres.tp_eigval))
      return false;
    if (std::max(cont_src.cell_cnt_, cont_tgt.cell_cnt_) > 15 && diff_delt(cont_src.vol3_mean_, cont_tgt.vol3_mean_, simthres.ta_h_bar)) return false;
    const float sx = cont_src.com_(0) - cont_src.pos_mean_(0), sy = cont_src.com_(1) - cont_src.pos_mean_(1);
    const float tx = cont_tgt.com_(0) - cont_tgt.pos_mean_(0), ty = cont_tgt.com_(1) - cont_tgt.pos_mean_(1);
    const float r1 = std::sqrt(sx * sx + sy * sy), r2 = std::sqrt(tx * tx + ty * ty);
    if (diff_delt(r1, r2, simthres.ta_rcom) && diff_perc(r1, r2, simthres.tp_rcom)) return false;
    return true;
  }
};
