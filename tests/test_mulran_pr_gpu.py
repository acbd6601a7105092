"""BASELINE.json configs[4] in miniature: the MulRan parameter set (lv_grads_ 1.0 .. 8.5, ta_h_bar 0.75;
config/batch_bin_test_config.yaml:17,31) on a MulRan-shaped synthetic sequence with revisits, online loop (query each scan against
the growing database, add it, balance): the GPU path (windowed C-ABI) and the CPU oracle must produce the same predictions, hence
the same precision-recall curve under the reference's own metric logic (scripts/pr_mpe.py:71-163 restated in eval.pr_metrics,
pinned to the reference's KITTI-08 outcome file by tests/test_eval.py)."""
import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import eval as ev
from contour_context_b200 import synth
from helpers import make_batch

pytestmark = pytest.mark.gpu

N_SCENES, VISITS, N_PTS = 30, 4, 50000


def test_pr_curve_parity_mulran_parameters(built_lib, oracle):
    from contour_context_b200.engine import Engine

    n = N_SCENES * VISITS
    seeds = [700 + i % N_SCENES for i in range(n)]  # the route is driven four times: revisits 30 scans = 90 s later
    visits = [i // N_SCENES for i in range(n)]
    pts, offsets = make_batch(seeds, visits, N_PTS, noise_seed=31)
    ts = 3.0 * np.arange(n)
    cm, dbc = D.kitti_cm_config(True), D.kitti_db_config(True)
    lb, ub = D.kitti_thres()
    # ground truth: scene k sits at x = 1000 k, the sensor pose inside the scene comes from the generator
    gt = np.array([[1000.0 * (s - 700) + synth.sensor_pose(s, v)[0], synth.sensor_pose(s, v)[1], 0.0] for s, v in zip(seeds, visits)])
    eng = Engine(cm_cfg=cm, db_cfg=dbc, scan_capacity=n + 8, max_batch=48, max_points=48 * 65536)
    try:
        g = np.zeros(n, D.QUERY_RESULT_DTYPE)
        for i0 in range(0, n, 48):
            m = min(48, n - i0)
            g[i0:i0 + m] = eng.online_window(pts[offsets[i0]:offsets[i0 + m]], offsets[i0:i0 + m + 1] - offsets[i0], ts[i0:i0 + m],
                                             np.arange(i0, i0 + m), lb, ub, int_ids=np.arange(i0, i0 + m))
    finally:
        eng.close()
    odb = oracle.DB(dbc)
    o = np.zeros(n, D.QUERY_RESULT_DTYPE)
    for i in range(n):
        s = oracle.Scan(cm, i).ingest(pts[offsets[i]:offsets[i + 1]])
        o[i] = odb.query(s, lb, ub)[0]
        odb.add_scan(s, ts[i])
        odb.push_and_balance(i, ts[i])

    def predictions(res):
        src = np.array([int(r["cand"][0]["cand_gidx"]) if r["n_cand"] > 0 else -1 for r in res])
        corr = np.array([float(r["cand"][0]["corr_fine"]) if r["n_cand"] > 0 else 0.0 for r in res])
        return src, corr

    gs, gc = predictions(g)
    os_, oc = predictions(o)
    assert int(g["overflow"].max()) == 0
    assert np.array_equal(gs, os_), "the GPU loop and the CPU oracle pair different scans"
    assert np.abs(gc - oc).max() <= 1e-5
    assert (gs >= 0).sum() >= n // 8, "the revisits must produce loop closures under the MulRan parameters too"
    ids = np.arange(n)
    mg = ev.pr_metrics(gt, ids, gs, gc, np.zeros((n, 3)), excl_frames=10)
    mo = ev.pr_metrics(gt, ids, os_, oc, np.zeros((n, 3)), excl_frames=10)
    # same ranking of the predictions => the same PR curve point by point
    assert np.array_equal(mg["pr_points"][:, 2], mo["pr_points"][:, 2]) or np.abs(np.sort(gc) - np.sort(oc)).max() <= 1e-5
    assert np.allclose(mg["pr_points"][:, :2], mo["pr_points"][:, :2], equal_nan=True)
    assert mg["max_f1"] == mo["max_f1"] and mg["max_f1"] > 0.5
