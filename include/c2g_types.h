/*
 * c2g_types.h — plain-old-data records shared by the C-ABI (include/c2g.h), the CUDA kernels,
 * the CPU oracle (oracle/) and the Python ctypes bindings.
 *
 * Every struct mirrors a reference type of the cont2contops hot path; the reference file:line each
 * one restates is cited next to it (paths relative to the reference repository root).
 * All structs are fixed-size, trivially copyable and have no implicit padding surprises
 * (sizes are static_assert'ed in C++ translation units).
 */
#ifndef C2G_TYPES_H
#define C2G_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Compile-time shape of the descriptor. The reference hard-codes the same numbers:
 *   RET_KEY_DIM = 10                      include/cont2/contour_mng.h:89
 *   BITS_PER_LAYER = 64                   include/cont2/contour_mng.h:112
 *   DIST_BIN_LAYERS = {1,2,3,4}           include/cont2/contour_mng.h:113
 *   LAYER_AREA_WEIGHTS = {.3,.3,.3,.1}    include/cont2/contour_mng.h:114
 * and both shipped configurations use 6 height levels, piv_firsts_ = 6, dist_firsts_ = 10
 * (config/batch_bin_test_config.yaml:30-46). */
#define C2G_NLEV 6            /* lv_grads_.size() supported by the kernels               */
#define C2G_KEY_DIM 10        /* RET_KEY_DIM                                             */
#define C2G_MAX_PIV 6         /* upper bound of piv_firsts_                              */
#define C2G_MAX_DIST_FIRSTS 10/* upper bound of dist_firsts_                             */
#define C2G_NUM_BIN_LAYERS 4  /* NUM_BIN_KEY_LAYER                                       */
#define C2G_BITS_PER_LAYER 64
#define C2G_MAX_NEI (C2G_NUM_BIN_LAYERS * C2G_MAX_DIST_FIRSTS)
#define C2G_MAX_CELLS 22500   /* n_row * n_col must not exceed this (150 x 150)          */
#define C2G_VIEW_CAP 2048     /* per-scan capacity of the contour-view arena             */
#define C2G_PAIR_BITS 400     /* (level-1)*100 + seq_src*10 + seq_tgt, level in 1..4     */
#define C2G_PAIR_WORDS 7      /* ceil(400 / 64)                                          */
#define C2G_MAX_CAND 32       /* candidate poses kept per query scan                     */
#define C2G_MAX_PROP 4        /* anchor proposals per candidate pose (contour_db.h:326)  */
#define C2G_NUM_Q_LEVELS_MAX 4
#define C2G_NUM_BUCKETS 6     /* LayerDB::max_num_backets_  include/cont2/contour_db.h:162 */

/* ContourManagerConfig (include/cont2/contour_mng.h:92-110) + ContourViewStatConfig
 * (include/cont2/contour.h:32-37). */
typedef struct c2g_cm_config {
  float lv_grads[8];
  int32_t n_levels; /* must be C2G_NLEV */
  float reso_row, reso_col;
  int32_t n_row, n_col;
  float lidar_height;
  float blind_sq;
  int32_t min_cont_key_cnt;
  int32_t min_cont_cell_cnt;
  int32_t piv_firsts;
  int32_t dist_firsts;
  float roi_radius;
  /* ContourViewStatConfig */
  int32_t min_cell_cov;
  float point_sigma;
  float com_bias_thres;
} c2g_cm_config;

/* ContourSimThresConfig (include/cont2/contour.h:40-45). */
typedef struct c2g_sim_config {
  float ta_cell_cnt, tp_cell_cnt;
  float tp_eigval;
  float ta_h_bar;
  float ta_rcom, tp_rcom;
} c2g_sim_config;

/* CandidateScoreEnsemble (include/cont2/contour_db.h:244-250) flattened:
 * ScoreConstellSim{3 int}, ScorePairwiseSim{2 int}, ScorePostProc{3 float}
 * (include/cont2/contour_mng.h:121-219). */
typedef struct c2g_score_ensemble {
  int32_t i_ovlp_sum, i_ovlp_max_one, i_in_ang_rng;
  int32_t i_indiv_sim, i_orie_sim;
  float correlation, area_perc, neg_est_dist;
} c2g_score_ensemble;

/* ContourDBConfig (include/cont2/contour_db.h:658-669) + TreeBucketConfig (:54-57). */
typedef struct c2g_db_config {
  int32_t nnk;
  int32_t max_fine_opt;
  int32_t n_q_levels;
  int32_t q_levels[C2G_NUM_Q_LEVELS_MAX];
  c2g_sim_config cont_sim;
  double max_elapse, min_elapse;
} c2g_db_config;

/* ContourView (include/cont2/contour.h:97-119). eig_vecs/pos_cov are column-major like
 * Eigen's default storage: [m(0,0), m(1,0), m(0,1), m(1,1)]. 80 bytes. */
typedef struct c2g_view {
  int16_t level, poi_r, poi_c, cell_cnt;
  float pos_mean[2];
  float pos_cov[4];
  float eig_vals[2];
  float eig_vecs[4];
  float eccen;
  float vol3_mean;
  float com[2];
  uint8_t ecc_feat, com_feat;
  uint8_t pad_[6];
} c2g_view;

/* BCI::RelativePoint (include/cont2/contour_mng.h:245-258). 12 bytes. */
typedef struct c2g_relpt {
  int8_t level, seq;
  int16_t bit_pos;
  float r, theta;
} c2g_relpt;

/* BCI (include/cont2/contour_mng.h:243-280): 256-bit distance bitset, neighbour list sorted by
 * bit position, run boundaries of equal bit position. 608 bytes. */
typedef struct c2g_bci {
  uint64_t dist_bin[C2G_NUM_BIN_LAYERS];
  c2g_relpt nei[C2G_MAX_NEI];
  uint16_t seg[C2G_MAX_NEI + 2];
  int16_t n_nei, n_seg;
  int8_t piv_seq, level;
  uint8_t pad_[6];
} c2g_bci;

/* Everything ContourManager keeps after clearImage() except the full view lists
 * (include/cont2/contour_mng.h:426-436): per level the number of views, the total cell count,
 * the 6 retrieval keys and the 6 BCIs; plus what the GMM-L2 stage needs that depends on this scan
 * only (number of ellipses up to the 95 % area cut and the auto-correlation,
 * include/cont2/correlation.h:62-77,102-119). Views live in a separate arena of C2G_VIEW_CAP
 * records per scan, level l occupying [view_off[l], view_off[l] + n_views[l]). */
typedef struct c2g_scan_head {
  int32_t int_id;
  int32_t status; /* 0 ok; bit0: view arena overflow; bit1: level list > kernel capacity */
  int32_t n_views[C2G_NLEV];
  int32_t view_off[C2G_NLEV];
  int32_t layer_cell_cnt[C2G_NLEV];
  int32_t n_ell[C2G_NUM_BIN_LAYERS]; /* GMM levels 1..4: #views before the 95 % area break */
  int32_t n_occupied;                /* bev_pixfs_.size() */
  int32_t pad_;
  double gmm_auto_corr;              /* auto_corr of this scan, levels 1..4 */
  float keys[C2G_NLEV][C2G_MAX_PIV][C2G_KEY_DIM];
  c2g_bci bcis[C2G_NLEV][C2G_MAX_PIV];
} c2g_scan_head;

/* One kNN result = one "hint" for CandidateManager::checkCandWithHint
 * (include/cont2/contour_db.h:764-769): candidate scan gidx, ConstellationPair(level, cand_seq,
 * q_seq) and the squared key distance. 16 bytes. */
typedef struct c2g_hint {
  int32_t q_idx;     /* query scan index inside the batch */
  int32_t cand_gidx; /* IndexOfKey::gidx; -1 = empty slot */
  int8_t level, cand_seq, q_seq, q_level_idx;
  float dist_sq;
} c2g_hint;

/* Result of one checkCandWithHint cascade up to (and excluding) addProposal
 * (include/cont2/contour_db.h:374-437). 128 bytes. */
typedef struct c2g_pair_score {
  int32_t constell[3]; /* ScoreConstellSim  */
  int32_t pairwise[2]; /* ScorePairwiseSim  */
  int32_t passed;      /* 1: reached addProposal; gate index that stopped it otherwise (<=0) */
  int32_t n_pairs;     /* tmp_pairs2.size() */
  int32_t pad_;
  double T[4];         /* T_pass as (cos, sin, tx, ty) */
  uint64_t pair_bits[C2G_PAIR_WORDS]; /* set of ConstellationPair in tmp_pairs2 */
  uint64_t pad2_;
} c2g_pair_score;

/* One surviving candidate pose after tidyUpCandidates (include/cont2/contour_db.h:494-596). */
typedef struct c2g_cand {
  int32_t cand_gidx;
  int32_t vote_cnt;
  float area_perc;
  float corr_init;
  double neg_est_dist;
  double T[4]; /* best proposal's T_delta_ as (cos, sin, tx, ty): the constellation estimate, input of the refinement */
  /* ConstellCorrelation::calcCorrelation (include/cont2/correlation.h:206-238), filled for the first
   * min(max_fine_opt, n_cand) entries (CandidateManager::fineOptimize, contour_db.h:604-648); others keep
   * corr_fine = 0, fine_iters = -1 and T_fine = T.  After fineOptimize the reference's anch_props_[0].correlation_ /
   * T_delta_ are exactly (corr_fine, T_fine). */
  float corr_fine;
  int16_t fine_iters; /* line-search iterations performed, -1 = not refined */
  int8_t fine_term;   /* 0 iteration cap, 1 converged, 2 solver failure (parameters stay at T, final cost -1) */
  int8_t fine_flags;  /* bit 0: the pre-selected pair list overflowed the device scratch (result invalid) */
  double T_fine[4];   /* refined (cos, sin, tx, ty) */
} c2g_cand;

/* Per query scan: outcome of the candidate cascade. */
typedef struct c2g_query_result {
  int32_t n_cand;            /* candidates surviving tidyUpCandidates */
  int32_t n_pose_before;     /* candidate poses before tidy-up */
  int32_t cand_aft_check[3]; /* cand_aft_check1..3 (contour_db.h:357-359) */
  int32_t overflow;          /* 1 if more than C2G_MAX_CAND poses were proposed */
  int32_t best;              /* index into cand[] of the pose returned (-1 if none) */
  int32_t pad_;
  c2g_cand cand[C2G_MAX_CAND];
} c2g_query_result;

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
static_assert(sizeof(c2g_view) == 80, "c2g_view layout");
static_assert(sizeof(c2g_relpt) == 12, "c2g_relpt layout");
static_assert(sizeof(c2g_bci) == 608, "c2g_bci layout");
static_assert(sizeof(c2g_hint) == 16, "c2g_hint layout");
static_assert(sizeof(c2g_pair_score) == 128, "c2g_pair_score layout");
static_assert(sizeof(c2g_cand) == 96, "c2g_cand layout");
static_assert(sizeof(c2g_scan_head) % 8 == 0, "c2g_scan_head alignment");
#endif

#endif /* C2G_TYPES_H */
