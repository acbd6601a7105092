// cont2/contour.h (facade) — same type and member names as the reference's include/cont2/contour.h for everything the
// cont2_batch_bin_test loop reads.  The statistics are NOT computed here: they are filled from the c2g_view records the
// GPU path produces (contour_context_b200/csrc/contours.cu restates ContourView::calcStatVals, contour.h:142-255).
#pragma once
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
};
