"""ctypes mirrors of include/c2g_types.h (keep the two files in sync; tests/test_abi.py checks the sizes).

Each class cites the reference type it restates through the header (c2g_types.h carries the file:line).
"""
import ctypes as C

import numpy as np

NLEV = 6
KEY_DIM = 10
MAX_PIV = 6
MAX_DIST_FIRSTS = 10
NUM_BIN_LAYERS = 4
BITS_PER_LAYER = 64
MAX_NEI = NUM_BIN_LAYERS * MAX_DIST_FIRSTS
MAX_CELLS = 22500
VIEW_CAP = 2048
PAIR_WORDS = 7
MAX_CAND = 32
NUM_Q_LEVELS_MAX = 4
NUM_BUCKETS = 6


class CmConfig(C.Structure):
    """ContourManagerConfig + ContourViewStatConfig (c2g_cm_config)."""

    _fields_ = [
        ("lv_grads", C.c_float * 8),
        ("n_levels", C.c_int32),
        ("reso_row", C.c_float),
        ("reso_col", C.c_float),
        ("n_row", C.c_int32),
        ("n_col", C.c_int32),
        ("lidar_height", C.c_float),
        ("blind_sq", C.c_float),
        ("min_cont_key_cnt", C.c_int32),
        ("min_cont_cell_cnt", C.c_int32),
        ("piv_firsts", C.c_int32),
        ("dist_firsts", C.c_int32),
        ("roi_radius", C.c_float),
        ("min_cell_cov", C.c_int32),
        ("point_sigma", C.c_float),
        ("com_bias_thres", C.c_float),
    ]


class SimConfig(C.Structure):
    """ContourSimThresConfig (c2g_sim_config)."""

    _fields_ = [
        ("ta_cell_cnt", C.c_float),
        ("tp_cell_cnt", C.c_float),
        ("tp_eigval", C.c_float),
        ("ta_h_bar", C.c_float),
        ("ta_rcom", C.c_float),
        ("tp_rcom", C.c_float),
    ]


class ScoreEnsemble(C.Structure):
    """CandidateScoreEnsemble (c2g_score_ensemble)."""

    _fields_ = [
        ("i_ovlp_sum", C.c_int32),
        ("i_ovlp_max_one", C.c_int32),
        ("i_in_ang_rng", C.c_int32),
        ("i_indiv_sim", C.c_int32),
        ("i_orie_sim", C.c_int32),
        ("correlation", C.c_float),
        ("area_perc", C.c_float),
        ("neg_est_dist", C.c_float),
    ]


class DbConfig(C.Structure):
    """ContourDBConfig + TreeBucketConfig (c2g_db_config)."""

    _fields_ = [
        ("nnk", C.c_int32),
        ("max_fine_opt", C.c_int32),
        ("n_q_levels", C.c_int32),
        ("q_levels", C.c_int32 * NUM_Q_LEVELS_MAX),
        ("cont_sim", SimConfig),
        ("max_elapse", C.c_double),
        ("min_elapse", C.c_double),
    ]


VIEW_DTYPE = np.dtype(
    [
        ("level", "<i2"),
        ("poi_r", "<i2"),
        ("poi_c", "<i2"),
        ("cell_cnt", "<i2"),
        ("pos_mean", "<f4", (2,)),
        ("pos_cov", "<f4", (4,)),
        ("eig_vals", "<f4", (2,)),
        ("eig_vecs", "<f4", (4,)),
        ("eccen", "<f4"),
        ("vol3_mean", "<f4"),
        ("com", "<f4", (2,)),
        ("ecc_feat", "u1"),
        ("com_feat", "u1"),
        ("pad_", "u1", (6,)),
    ]
)
assert VIEW_DTYPE.itemsize == 80

RELPT_DTYPE = np.dtype([("level", "i1"), ("seq", "i1"), ("bit_pos", "<i2"), ("r", "<f4"), ("theta", "<f4")])
assert RELPT_DTYPE.itemsize == 12

BCI_DTYPE = np.dtype(
    [
        ("dist_bin", "<u8", (NUM_BIN_LAYERS,)),
        ("nei", RELPT_DTYPE, (MAX_NEI,)),
        ("seg", "<u2", (MAX_NEI + 2,)),
        ("n_nei", "<i2"),
        ("n_seg", "<i2"),
        ("piv_seq", "i1"),
        ("level", "i1"),
        ("pad_", "u1", (6,)),
    ]
)
assert BCI_DTYPE.itemsize == 608

SCAN_HEAD_DTYPE = np.dtype(
    [
        ("int_id", "<i4"),
        ("status", "<i4"),
        ("n_views", "<i4", (NLEV,)),
        ("view_off", "<i4", (NLEV,)),
        ("layer_cell_cnt", "<i4", (NLEV,)),
        ("n_ell", "<i4", (NUM_BIN_LAYERS,)),
        ("n_occupied", "<i4"),
        ("pad_", "<i4"),
        ("gmm_auto_corr", "<f8"),
        ("keys", "<f4", (NLEV, MAX_PIV, KEY_DIM)),
        ("bcis", BCI_DTYPE, (NLEV, MAX_PIV)),
    ]
)
assert SCAN_HEAD_DTYPE.itemsize == 23440, SCAN_HEAD_DTYPE.itemsize

HINT_DTYPE = np.dtype(
    [
        ("q_idx", "<i4"),
        ("cand_gidx", "<i4"),
        ("level", "i1"),
        ("cand_seq", "i1"),
        ("q_seq", "i1"),
        ("q_level_idx", "i1"),
        ("dist_sq", "<f4"),
    ]
)
assert HINT_DTYPE.itemsize == 16

PAIR_SCORE_DTYPE = np.dtype(
    [
        ("constell", "<i4", (3,)),
        ("pairwise", "<i4", (2,)),
        ("passed", "<i4"),
        ("n_pairs", "<i4"),
        ("pad_", "<i4"),
        ("T", "<f8", (4,)),
        ("pair_bits", "<u8", (PAIR_WORDS,)),
        ("pad2_", "<u8"),
    ]
)
assert PAIR_SCORE_DTYPE.itemsize == 128

CAND_DTYPE = np.dtype(
    [
        ("cand_gidx", "<i4"),
        ("vote_cnt", "<i4"),
        ("area_perc", "<f4"),
        ("corr_init", "<f4"),
        ("neg_est_dist", "<f8"),
        ("T", "<f8", (4,)),
        ("corr_fine", "<f4"),
        ("fine_iters", "<i2"),
        ("fine_term", "i1"),
        ("fine_flags", "i1"),
        ("T_fine", "<f8", (4,)),
    ]
)
assert CAND_DTYPE.itemsize == 96

QUERY_RESULT_DTYPE = np.dtype(
    [
        ("n_cand", "<i4"),
        ("n_pose_before", "<i4"),
        ("cand_aft_check", "<i4", (3,)),
        ("overflow", "<i4"),
        ("best", "<i4"),
        ("pad_", "<i4"),
        ("cand", CAND_DTYPE, (MAX_CAND,)),
    ]
)
assert QUERY_RESULT_DTYPE.itemsize == 32 + 96 * MAX_CAND


def kitti_cm_config(mulran: bool = False) -> CmConfig:
    """config/batch_bin_test_config.yaml:28-46 (KITTI) or the MulRan lv_grads_ of line 31."""
    cfg = CmConfig()
    grads = [1.0, 2.5, 4.0, 5.5, 7.0, 8.5] if mulran else [1.5, 2.0, 2.5, 3.0, 3.5, 4.0]
    for i, g in enumerate(grads):
        cfg.lv_grads[i] = g
    cfg.n_levels = 6
    cfg.reso_row = cfg.reso_col = 1.0
    cfg.n_row = cfg.n_col = 150
    cfg.lidar_height = 2.0
    cfg.blind_sq = 9.0
    cfg.min_cont_key_cnt = 9
    cfg.min_cont_cell_cnt = 3
    cfg.piv_firsts = 6
    cfg.dist_firsts = 10
    cfg.roi_radius = 10.0
    cfg.min_cell_cov = 4
    cfg.point_sigma = 1.0
    cfg.com_bias_thres = 0.5
    return cfg


def kitti_db_config(mulran: bool = False) -> DbConfig:
    """config/batch_bin_test_config.yaml:6-23."""
    cfg = DbConfig()
    cfg.nnk = 50
    cfg.max_fine_opt = 10
    cfg.n_q_levels = 3
    for i, l in enumerate([1, 2, 3]):
        cfg.q_levels[i] = l
    cfg.cont_sim.ta_cell_cnt = 6.0
    cfg.cont_sim.tp_cell_cnt = 0.2
    cfg.cont_sim.tp_eigval = 0.2
    cfg.cont_sim.ta_h_bar = 0.75 if mulran else 0.3
    cfg.cont_sim.ta_rcom = 0.4
    cfg.cont_sim.tp_rcom = 0.25
    cfg.max_elapse = 25.0
    cfg.min_elapse = 15.0
    return cfg


def kitti_thres():
    """thres_lb_/thres_ub_ of config/batch_bin_test_config.yaml:70-87."""
    lb = ScoreEnsemble(3, 3, 3, 3, 4, 0.3, 0.03, -5.01)
    ub = ScoreEnsemble(6, 6, 6, 6, 6, 0.75, 0.15, -5.0)
    return lb, ub
