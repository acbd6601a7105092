"""Outcome / precision-recall harness of the loop-closure path, host logic only.

Mirrors, with the reference's names and file formats:
  * ContLCDEvaluator           include/eval/evaluator.h:39-425   (gt association, TP/FP/TN/FN bookkeeping, outcome file)
  * ConstellCorrelation::evalMetricEst / getEstSensTF   include/cont2/correlation.h:241-296
  * scripts/pr_mpe.py:71-163   (PR points, max-F1, metric pose error of the true positives)

The C++ twin lives in contour_context_b200/host (eval/evaluator.h + cont2_batch_bin --eval); both write the same file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

TP, FP, TN, FN = 0, 1, 2, 3  # PredictionOutcome::Res (evaluator.h:31-33)


# ---- small SE(2)/SE(3) helpers (Eigen::Isometry semantics) ------------------------------------------------------------
def iso2(theta: float, x: float, y: float) -> np.ndarray:
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[c, -s, x], [s, c, y], [0.0, 0.0, 1.0]])


def iso2_from_cs(T: Sequence[float]) -> np.ndarray:
    """(cos, sin, tx, ty) as stored in c2g_cand.T_fine -> 3x3 homogeneous matrix."""
    return np.array([[T[0], -T[1], T[2]], [T[1], T[0], T[3]], [0.0, 0.0, 1.0]])


def _iso_inv(T: np.ndarray) -> np.ndarray:
    n = T.shape[0] - 1
    R, t = T[:n, :n], T[:n, n]
    out = np.eye(n + 1)
    out[:n, :n] = R.T
    out[:n, n] = -R.T @ t
    return out


def quat_from_rot(m: np.ndarray) -> np.ndarray:
    """Eigen::Quaterniond(Matrix3d) (w, x, y, z): trace branch or largest-diagonal branch, no normalisation."""
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def rot_from_quat(q: np.ndarray) -> np.ndarray:
    """Quaterniond::toRotationMatrix()."""
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx], [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def pose_from_row(vals: Sequence[float]) -> np.ndarray:
    """12 numbers (row-major 3x4) -> 4x4, the rotation passing through a quaternion like evaluator.h:100-103."""
    m = np.asarray(vals, float).reshape(3, 4)
    T = np.eye(4)
    T[:3, :3] = rot_from_quat(quat_from_rot(m[:, :3]))
    T[:3, 3] = m[:, 3]
    return T


def get_est_sens_tf(T_delta: np.ndarray, n_row: int, n_col: int) -> np.ndarray:
    """ConstellCorrelation::getEstSensTF (correlation.h:287-296)."""
    T_so = np.eye(3)
    T_so[0, 2] = n_row // 2 - 0.5
    T_so[1, 2] = n_col // 2 - 0.5
    return _iso_inv(T_so) @ T_delta @ T_so


def eval_metric_est(T_delta: np.ndarray, gt_src_3d: np.ndarray, gt_tgt_3d: np.ndarray, n_row: int, n_col: int, reso: float) -> np.ndarray:
    """ConstellCorrelation::evalMetricEst (correlation.h:241-280): estimate^-1-composed error T_gt^-1 * T_est in SE(2)."""
    T_est = get_est_sens_tf(T_delta, n_row, n_col)
    T_est[:2, 2] *= reso
    T3 = _iso_inv(gt_tgt_3d) @ gt_src_3d
    z0 = np.array([0.0, 0.0, 1.0])
    z1 = T3[:3, 2]
    ax = np.cross(z0, z1)
    nrm2 = float(ax @ ax)
    if nrm2 > 0:
        ax = ax / math.sqrt(nrm2)
    with np.errstate(invalid="ignore"):
        ang = float(np.arccos(np.float64(z0 @ z1)))  # NaN beyond [-1, 1], like std::acos in the reference
    # AngleAxisd(-ang, ax).matrix()
    a = -ang
    c, s = math.cos(a), math.sin(a)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    d_rot = c * np.eye(3) + s * K + (1 - c) * np.outer(ax, ax)
    R_rect = d_rot @ T3[:3, :3]
    T_gt = iso2(math.atan2(R_rect[1, 0], R_rect[0, 0]), T3[0, 3], T3[1, 3])
    return _iso_inv(T_gt) @ T_est


def lookup_nn(q, sorted_vals, tol) -> int:
    """lookupNN (tools/algos.h:12-37): index of the nearest value within tol, or -1."""
    vals = np.asarray(sorted_vals)
    if len(vals) == 0:
        return -1
    i = int(np.searchsorted(vals, q))
    best, bi = None, -1
    for j in (i - 1, i):
        if 0 <= j < len(vals):
            d = abs(vals[j] - q)
            if best is None or d < best:
                best, bi = d, j
    return bi if best is not None and best <= tol else -1


class SimpleRMSE:
    """evaluator.h:12-29 / pr_mpe.py:45-68."""

    def __init__(self):
        self.sum_sqs = 0.0
        self.sum_abs = 0.0
        self.cnt_sqs = 0

    def add_one_err(self, d: Sequence[float]):
        self.cnt_sqs += 1
        tmp = 0.0
        for v in d:
            tmp += v * v
        self.sum_sqs += tmp
        self.sum_abs += math.sqrt(tmp)

    def rmse(self) -> float:
        return math.sqrt(self.sum_sqs / self.cnt_sqs) if self.cnt_sqs else -1.0

    def mean(self) -> float:
        return self.sum_abs / self.cnt_sqs if self.cnt_sqs else -1.0


@dataclass
class PredictionOutcome:
    id_src: int = -1
    id_tgt: int = -1
    tfpn: int = TN
    est_err: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    correlation: float = 0.0


@dataclass
class LaserScanInfo:
    has_gt_positive_lc: bool = False
    sens_pose: Optional[np.ndarray] = None
    seq: int = 0
    ts: float = 0.0
    fpath: str = ""


class ContLCDEvaluator:
    """evaluator.h:39-425.  fpath_pose: '<ts> r00 r01 r02 tx r10 ... tz' per line; fpath_laser: '<ts> <seq> <bin path>'."""

    ts_diff_tol = 10e-3
    min_time_excl = 15.0

    def __init__(self, fpath_pose: str, fpath_laser: str, bar: float):
        self.sim_thres = bar
        gt_tss, gt_poses = [], []
        with open(fpath_pose) as f:
            for line in f:
                vals = line.split()
                if len(vals) < 13:
                    continue
                gt_tss.append(float(vals[0]))
                gt_poses.append(pose_from_row([float(v) for v in vals[1:13]]))
        order = np.argsort(np.asarray(gt_tss), kind="stable")
        gt_tss = [gt_tss[i] for i in order]
        gt_poses = [gt_poses[i] for i in order]
        self.laser_info: List[LaserScanInfo] = []
        self.assigned_seqs: List[int] = []
        with open(fpath_laser) as f:
            for line in f:
                vals = line.split()
                if len(vals) < 3:
                    continue
                ts, seq, path = float(vals[0]), int(vals[1]), vals[2]
                gi = lookup_nn(ts, gt_tss, self.ts_diff_tol)
                if gi < 0:
                    continue
                self.laser_info.append(LaserScanInfo(False, gt_poses[gi], seq, ts, path))
                self.assigned_seqs.append(seq)
        for a, b in zip(self.laser_info[:-1], self.laser_info[1:]):
            assert a.seq < b.seq and a.ts < b.ts, "laser scans must be ordered by seq and time (evaluator.h:182-188)"
        for fast in self.laser_info:  # evaluator.h:194-208
            for slow in self.laser_info:
                if fast.ts < slow.ts + self.min_time_excl:
                    break
                if np.linalg.norm(fast.sens_pose[:3, 3] - slow.sens_pose[:3, 3]) < 5.0:
                    fast.has_gt_positive_lc = True
                    break
        self.p_lidar_curr = -1
        self.tp_trans, self.all_trans = SimpleRMSE(), SimpleRMSE()
        self.tp_rot, self.all_rot = SimpleRMSE(), SimpleRMSE()
        self.pred_records: List[PredictionOutcome] = []

    def load_new_scan(self) -> bool:
        self.p_lidar_curr += 1
        return self.p_lidar_curr < len(self.laser_info)

    def curr_scan_info(self) -> LaserScanInfo:
        return self.laser_info[self.p_lidar_curr]

    def _addr(self, seq: int) -> int:
        a = lookup_nn(seq, self.assigned_seqs, 0)
        assert a >= 0
        return a

    def add_prediction(self, id_tgt: int, est_corr: float, id_src: Optional[int] = None, T_est_delta_2d: Optional[np.ndarray] = None,
                       n_row: int = 150, n_col: int = 150, reso: float = 1.0) -> PredictionOutcome:
        """addPrediction (evaluator.h:305-373).  id_tgt: the query scan; id_src: the retrieved scan (None = negative)."""
        addr_tgt = self._addr(id_tgt)
        rec = PredictionOutcome(id_tgt=id_tgt, correlation=est_corr)
        tgt = self.laser_info[addr_tgt]
        if id_src is not None:
            src = self.laser_info[self._addr(id_src)]
            rec.id_src = id_src
            tf_err = eval_metric_est(T_est_delta_2d, src.sens_pose, tgt.sens_pose, n_row, n_col, reso)
            gt_trans_norm3d = float(np.linalg.norm(src.sens_pose[:3, 3] - tgt.sens_pose[:3, 3]))
            err = [float(tf_err[0, 2]), float(tf_err[1, 2]), math.atan2(tf_err[1, 0], tf_err[0, 0])]
            rec.est_err = err
            if est_corr >= self.sim_thres:
                if tgt.has_gt_positive_lc and gt_trans_norm3d < 5.0:
                    rec.tfpn = TP
                    self.tp_trans.add_one_err(err[:2])
                    self.tp_rot.add_one_err(err[2:])
                else:
                    rec.tfpn = FP
            else:
                rec.tfpn = FN if tgt.has_gt_positive_lc else TN
            self.all_trans.add_one_err(err[:2])
            self.all_rot.add_one_err(err[2:])
        else:
            rec.tfpn = FN if tgt.has_gt_positive_lc else TN
        self.pred_records.append(rec)
        return rec

    def save_prediction_results(self, sav_path: str):
        """savePredictionResults (evaluator.h:377-425): '<tfpn>\\t<tgt>-<src|x>\\t<corr>\\t<ex>\\t<ey>\\t<etheta>\\t<tgt path>\\t<src path>'."""
        with open(sav_path, "w") as f:
            for rec in self.pred_records:
                tgt_path = self.laser_info[self._addr(rec.id_tgt)].fpath
                if rec.id_src < 0:
                    pair, src_path = f"{rec.id_tgt}-x", "x"
                else:
                    pair, src_path = f"{rec.id_tgt}-{rec.id_src}", self.laser_info[self._addr(rec.id_src)].fpath
                f.write("%d\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (rec.tfpn, pair, fmt6(rec.correlation), fmt6(rec.est_err[0]), fmt6(rec.est_err[1]),
                                                        fmt6(rec.est_err[2]), tgt_path[-32:], src_path[-32:]))


def fmt6(v: float) -> str:
    """operator<<(ostream, double) with the default precision of 6 significant digits (%g)."""
    return "%g" % v


# ---- scripts/pr_mpe.py ----------------------------------------------------------------------------------------------------
def read_outcome(path: str):
    """Outcome file -> (tfpn, id_tgt, id_src (-1 = x), corr, err[3]) arrays."""
    tf, it, isr, corr, err = [], [], [], [], []
    with open(path) as f:
        for line in f:
            v = line.split()
            if len(v) < 6:
                continue
            a, b = v[1].split("-")
            tf.append(int(v[0]))
            it.append(int(a))
            isr.append(-1 if b == "x" else int(b))
            corr.append(float(v[2]))
            err.append([float(v[3]), float(v[4]), float(v[5])])
    return np.array(tf), np.array(it), np.array(isr), np.array(corr, float), np.array(err, float).reshape(-1, 3)


def pr_metrics(gt_xyz: np.ndarray, id_tgt: np.ndarray, id_src: np.ndarray, corr: np.ndarray, err: np.ndarray, thres_dist: float = 5.0,
               excl_frames: int = 150) -> dict:
    """get_points_ours2 (scripts/pr_mpe.py:71-163).  gt_xyz[i] = translation of pose line i; ids index pose lines.
    Returns the PR points in outcome order of descending correlation, the max-F1 point and the TP pose errors above its
    similarity threshold."""
    n = gt_xyz.shape[0]
    gt_positive = np.zeros(n)
    # a pose is a ground-truth positive if some pose more than excl_frames earlier lies within thres_dist (pr_mpe.py:83-88)
    for i0 in range(0, n, 512):
        blk = gt_xyz[i0:i0 + 512]
        d = np.linalg.norm(blk[:, None, :] - gt_xyz[None, :, :], axis=2)
        j = np.arange(n)[None, :]
        i = (np.arange(i0, min(n, i0 + 512)))[:, None]
        # query_ball_point returns points with distance <= r
        gt_positive[i0:i0 + 512] = ((d <= thres_dist) & (j < i - excl_frames)).any(axis=1)
    m = len(id_tgt)
    est = np.zeros((m, 4))
    est[:, 0] = corr
    est[:, 3] = id_tgt
    has = id_src >= 0
    dd = np.linalg.norm(gt_xyz[id_tgt[has]] - gt_xyz[id_src[has]], axis=1)
    est[has, 1] = dd < thres_dist
    est[:, 2] = gt_positive[id_tgt]
    orig = est.copy()
    order = (-est[:, 0]).argsort()  # numpy's default (introsort) like the script; ties are broken the same way
    est = est[order]
    tp = np.cumsum(est[:, 1] != 0)
    fp = np.cumsum(est[:, 1] == 0)
    pos_after = np.cumsum((est[:, 2] != 0)[::-1])[::-1]
    fn = np.concatenate([pos_after[1:], [0]])
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = tp / (tp + fn)
        precision = tp / (tp + fp)
    pr_points = np.stack([recall, precision, est[:, 3]], axis=1)
    max_f1, idx = 0.0, -1
    for r, p, k in pr_points:  # get_maxf1_idx (pr_mpe.py:29-42): first strict maximum
        cur = 2 * r * p / (r + p) if (r + p) > 0 else 0
        if max_f1 < cur:
            max_f1, idx = cur, k
    out = dict(pr_points=pr_points, max_f1=float(max_f1), f1_pose_idx=int(idx), gt_positive=gt_positive)
    if idx >= 0:
        sim_thres = float(corr[int(idx)])  # the script indexes the outcome LINES with the pose id (pr_mpe.py:143)
        tr, ro = SimpleRMSE(), SimpleRMSE()
        for i in range(m):
            if corr[i] >= sim_thres and orig[i, 1] == 1 and orig[i, 2] == 1:
                tr.add_one_err([err[i, 0], err[i, 1]])
                ro.add_one_err([err[i, 2]])
        out.update(sim_thres=sim_thres, tp_count=ro.cnt_sqs, rot_mean_deg=ro.mean() / math.pi * 180, rot_rmse_deg=ro.rmse() / math.pi * 180,
                   trans_mean=tr.mean(), trans_rmse=tr.rmse())
    return out
