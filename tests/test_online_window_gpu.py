"""Windowed online loop (BASELINE.json configs[1]: query each scan against the growing DB, then add it) through the C-ABI:
c2g_online_window / c2g_online_stage + c2g_online_commit must return exactly what the scan-by-scan call sequence of
test/batch_bin_test.cpp:179,234,237 (queryRangedKNN -> addScan -> pushAndBalance) returns, and what the CPU oracle returns.

Timestamps are 2 s apart so that, inside every window, keys leave the time-delay buffers (15 s / 25 s gates), trees grow and
rebalancing moves rewrite buckets: the windows are cut into several kNN runs."""
import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import synth
from helpers import make_batch

pytestmark = pytest.mark.gpu

N_DB, N_PTS = 120, 40000


@pytest.fixture(scope="module")
def sequence():
    seeds, visits = synth.db_layout(N_DB, 4, first_scene=700)
    # revisits must come LATER in time than the scans they close a loop with: visit-major order (all scenes once, then again..)
    order = np.argsort(np.asarray(visits), kind="stable")
    seeds, visits = [seeds[i] for i in order], [visits[i] for i in order]
    pts, offsets = make_batch(seeds, visits, N_PTS, noise_seed=21)
    ts = 2.0 * np.arange(N_DB)
    return pts, offsets, ts


def _sequential(pts, offsets, ts):
    from contour_context_b200.engine import Engine

    lb, ub = D.kitti_thres()
    eng = Engine(scan_capacity=N_DB + 8, max_batch=64, max_points=64 * 65536)
    out = np.zeros(N_DB, D.QUERY_RESULT_DTYPE)
    try:
        for i in range(N_DB):
            eng.ingest(pts[offsets[i]:offsets[i + 1]], np.array([0, N_PTS], np.int64), first_slot=i, int_ids=np.array([i]))
            out[i] = eng.query(i, 1, lb, ub)[0]
            eng.db_add_scans(i, 1, [ts[i]])
            eng.db_push_and_balance(i, ts[i])
        state = [eng.db_layer_state(ll) for ll in range(eng.db_cfg.n_q_levels)]
        trees = [[eng.db_bucket_tree(ll, b) for b in range(D.NUM_BUCKETS)] for ll in range(eng.db_cfg.n_q_levels)]
    finally:
        eng.close()
    return out, state, trees


@pytest.fixture(scope="module")
def seq_result(built_lib, sequence):
    return _sequential(*sequence)


def _same_state(eng, state, trees):
    for ll in range(eng.db_cfg.n_q_levels):
        for a, b in zip(eng.db_layer_state(ll), state[ll]):
            assert a.tobytes() == b.tobytes()
        for bk in range(D.NUM_BUCKETS):
            for a, b in zip(eng.db_bucket_tree(ll, bk), trees[ll][bk]):
                assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("window", [1, 7, 37, 64, 120])
def test_window_equals_scan_by_scan(built_lib, sequence, seq_result, window):
    from contour_context_b200.engine import Engine

    pts, offsets, ts = sequence
    ref, state, trees = seq_result
    lb, ub = D.kitti_thres()
    mb = max(64, window)
    eng = Engine(scan_capacity=N_DB + 8, max_batch=mb, max_points=mb * 65536)
    try:
        got = np.zeros(N_DB, D.QUERY_RESULT_DTYPE)
        for i0 in range(0, N_DB, window):
            n = min(window, N_DB - i0)
            got[i0:i0 + n] = eng.online_window(pts[offsets[i0]:offsets[i0 + n]], offsets[i0:i0 + n + 1] - offsets[i0], ts[i0:i0 + n],
                                               np.arange(i0, i0 + n), lb, ub, int_ids=np.arange(i0, i0 + n))
        assert eng.db_size() == N_DB
        bad = [i for i in range(N_DB) if got[i].tobytes() != ref[i].tobytes()]
        assert not bad, f"window {window}: scans {bad[:10]} differ from the scan-by-scan loop"
        _same_state(eng, state, trees)
        if window >= 37:
            assert eng.online_runs() > N_DB // window + 2, "the windows must have been cut into several kNN runs"
            n_win = (N_DB + window - 1) // window
            print(f"window {window}: {eng.online_runs()} runs in {eng.online_groups()} kNN launches over {n_win} windows")
            assert eng.online_groups() < eng.online_runs(), "one launch must serve several runs"
        if window == 120:
            assert eng.online_groups() > 1, "a bucket rewritten twice inside the window must split the launch"
        assert int((ref["n_cand"] > 0).sum()) >= N_DB // 8, "the comparison must cover real loop closures"
        assert int(ref["overflow"].max()) == 0
    finally:
        eng.close()


@pytest.mark.parametrize("xyz", [False, True])
def test_pipelined_stage_commit(built_lib, sequence, seq_result, xyz):
    """stage(k + 1) before commit(k): the next window's copy + ingest are in flight while window k is queried.  xyz: the 12 B /
    point input variant (c2g_online_stage_xyz)."""
    import torch
    from contour_context_b200.engine import Engine

    pts, offsets, ts = sequence
    ref, state, trees = seq_result
    lb, ub = D.kitti_thres()
    W = 24
    eng = Engine(scan_capacity=N_DB + 8, max_batch=W, max_points=W * 65536)
    try:
        host = torch.from_numpy(pts).reshape(-1, 4)
        host = (host[:, :3].contiguous() if xyz else host).pin_memory()
        wins = [(i0, min(W, N_DB - i0)) for i0 in range(0, N_DB, W)]
        outs = [np.zeros(n, D.QUERY_RESULT_DTYPE) for _, n in wins]

        def stage(k):
            i0, n = wins[k]
            eng.online_stage(host[offsets[i0]:offsets[i0 + n]], offsets[i0:i0 + n + 1] - offsets[i0], int_ids=np.arange(i0, i0 + n),
                             on_device=False, xyz=xyz)

        stage(0)
        for k, (i0, n) in enumerate(wins):
            if k + 1 < len(wins):
                stage(k + 1)
            eng.online_commit(ts[i0:i0 + n], np.arange(i0, i0 + n), lb, ub, outs[k])
        eng.sync()
        got = np.concatenate(outs)
        bad = [i for i in range(N_DB) if got[i].tobytes() != ref[i].tobytes()]
        assert not bad, f"scans {bad[:10]} differ from the scan-by-scan loop"
        _same_state(eng, state, trees)
    finally:
        eng.close()


def test_window_equals_oracle_loop(built_lib, oracle, sequence, seq_result):
    """The same growing-DB loop on the CPU oracle (ingest + query + add + balance per scan)."""
    from contour_context_b200 import ctypes_defs as D

    pts, offsets, ts = sequence
    ref, _, _ = seq_result
    lb, ub = D.kitti_thres()
    cfg, dbc = D.kitti_cm_config(), D.kitti_db_config()
    odb = oracle.DB(dbc)
    ores = np.zeros(N_DB, D.QUERY_RESULT_DTYPE)
    for i in range(N_DB):
        s = oracle.Scan(cfg, i).ingest(pts[offsets[i]:offsets[i + 1]])
        ores[i] = odb.query(s, lb, ub)[0]
        odb.add_scan(s, ts[i])
        odb.push_and_balance(i, ts[i])
    n_lc = 0
    for i in range(N_DB):
        g, o = ref[i], ores[i]
        assert g["n_pose_before"] == o["n_pose_before"] and np.array_equal(g["cand_aft_check"], o["cand_aft_check"]), i
        assert g["n_cand"] == o["n_cand"] and g["best"] == o["best"], i
        n = int(g["n_cand"])
        if n:
            n_lc += 1
            gc, oc = g["cand"][:n], o["cand"][:n]
            assert np.array_equal(gc["cand_gidx"], oc["cand_gidx"]) and np.array_equal(gc["vote_cnt"], oc["vote_cnt"]), i
            assert np.abs(gc["corr_init"] - oc["corr_init"]).max() <= 1e-5
            assert np.abs(gc["corr_fine"] - oc["corr_fine"]).max() <= 1e-5
            assert np.array_equal(gc["fine_iters"], oc["fine_iters"])
    assert n_lc >= N_DB // 8
