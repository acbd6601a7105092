"""CPU-only: the product's host-side LayerDB bookkeeping (csrc/layer_db_host.cpp) tracks the oracle's restatement of
LayerDB::pushBuffer / rebuild (src/cont2/contour_db.cpp:63-317) bucket for bucket, key for key, through thousands of
time-gated insertions and rebalancing steps — including runs of identical bucket values ("contagious" splits)."""
import ctypes as C

import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _state_prod(lib, h, ll):
    rng = np.zeros(7, np.float32)
    ts = np.zeros(6, np.int32)
    bs = np.zeros(6, np.int32)
    assert lib.c2g_hostdb_state(h, ll, _p(rng), _p(ts), _p(bs)) == 0
    return rng, ts, bs


@pytest.mark.parametrize("mode", ["continuous", "quantised", "bursty", "sparse"])
def test_rebalance_matches_oracle(built_lib, oracle, mode):
    rng = np.random.default_rng({"continuous": 1, "quantised": 2, "bursty": 3, "sparse": 4}[mode])
    cfg = D.kitti_db_config()
    odb = oracle.DB(cfg)
    olib = oracle.lib()
    olib.c2o_test_push_key.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int]
    h = built_lib.c2g_hostdb_create(cfg.n_q_levels, cfg.max_elapse, cfg.min_elapse)
    n_scans = 1500
    lagging = False  # some bucket held keys its index did not cover yet (received in a move, nothing popped since)
    olib.c2o_db_indexed.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    built_lib.c2g_hostdb_indexed.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    for i in range(n_scans):
        ts = 0.104 * i if mode != "bursty" else 0.104 * i + (30.0 if (i // 200) % 2 else 0.0)
        if mode == "sparse":
            ts = 2.0 * i  # scans far apart: buffers drain completely, so a bucket can receive moved keys with nothing to pop
        for ll in range(cfg.n_q_levels):
            for seq in range(6):
                key = (rng.random(10) * 30 + 1).astype(np.float32)
                if mode == "sparse":
                    key[0] = np.float32(1.0 + 0.03 * i + 6.0 * rng.random())  # a drifting, narrow band of bucket values
                if mode == "quantised":
                    key[0] = np.float32(np.round(key[0] / 2.0) * 2.0)  # long strips of equal bucket values
                if rng.random() < 0.05:
                    key[:] = 0  # all-zero keys are never stored
                olib.c2o_test_push_key(odb.h, ll, _p(key), ts, i, seq)
                assert built_lib.c2g_hostdb_push_key(h, ll, _p(key), ts, i, seq) == 0
        odb.push_and_balance(i, ts)
        assert built_lib.c2g_hostdb_balance(h, i, ts) == 0
        if i % 7 == 0 or i == n_scans - 1 or mode == "sparse":
            for ll in range(cfg.n_q_levels):
                o_rng, o_ts, o_bs = odb.layer_state(ll)
                p_rng, p_ts, p_bs = _state_prod(built_lib, h, ll)
                assert o_rng.tobytes() == p_rng.tobytes(), (mode, i, ll)
                assert np.array_equal(o_ts, p_ts) and np.array_equal(o_bs, p_bs), (mode, i, ll, o_ts, p_ts)
                # the searchable prefix (what the reference's KD index covers) of every bucket
                o_ix, p_ix = np.zeros(6, np.int32), np.zeros(6, np.int32)
                olib.c2o_db_indexed(odb.h, ll, _p(o_ix))
                assert built_lib.c2g_hostdb_indexed(h, ll, _p(p_ix)) == 0
                assert np.array_equal(o_ix, p_ix), (mode, i, ll, o_ix, p_ix)
                lagging |= bool((p_ix < p_ts).any())
    # final: every tree identical in content and order
    split_seen = False
    for ll in range(cfg.n_q_levels):
        _, sizes, _ = odb.layer_state(ll)
        split_seen |= int((sizes > 0).sum()) > 1
        for b in range(6):
            ok, og, osq = odb.bucket_tree(ll, b)
            pk = np.zeros_like(ok)
            pg = np.zeros_like(og)
            ps = np.zeros_like(osq)
            if len(og):
                assert built_lib.c2g_hostdb_tree(h, ll, b, _p(pk), _p(pg), _p(ps)) == 0
            assert ok.tobytes() == pk.tobytes() and np.array_equal(og, pg) and np.array_equal(osq, ps), (mode, ll, b)
    assert split_seen, "the test never exercised a bucket split"
    if mode == "sparse":
        assert lagging, "no bucket ever held received keys that its index did not cover yet: the case is not exercised"
    built_lib.c2g_hostdb_free(h)


def test_trees_only_grow_at_the_end_between_restructures(built_lib):
    """The incremental device mirror (query.cu: c2g_db_sync_mode) appends to a bucket's region as long as the bucket's
    `restructured` counter stands still.  That is only correct if, between two bumps of the counter, the previous tree is a
    PREFIX of the current one — checked here after every pushAndBalance of a 1 200-scan stream with rebalancing moves."""
    rng = np.random.default_rng(7)
    cfg = D.kitti_db_config()
    h = built_lib.c2g_hostdb_create(cfg.n_q_levels, cfg.max_elapse, cfg.min_elapse)
    built_lib.c2g_hostdb_versions.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    prev = {}
    bumps = appends = 0
    for i in range(1200):
        ts = 0.25 * i
        for ll in range(cfg.n_q_levels):
            for seq in range(6):
                key = (rng.random(10) * (10 + 0.02 * i) + 1).astype(np.float32)  # a drifting distribution forces rebalancing
                assert built_lib.c2g_hostdb_push_key(h, ll, _p(key), ts, i, seq) == 0
        assert built_lib.c2g_hostdb_balance(h, i, ts) == 0
        for ll in range(cfg.n_q_levels):
            ver = np.zeros(6, np.uint32)
            assert built_lib.c2g_hostdb_versions(h, ll, _p(ver)) == 0
            _, sizes, _ = _state_prod(built_lib, h, ll)
            for b in range(6):
                n = int(sizes[b])
                keys = np.zeros((max(n, 1), 10), np.float32)
                gidx = np.zeros(max(n, 1), np.int32)
                sq = np.zeros(max(n, 1), np.int32)
                assert built_lib.c2g_hostdb_tree(h, ll, b, _p(keys), _p(gidx), _p(sq)) == 0
                cur = (int(ver[b]), keys[:n].tobytes() + gidx[:n].tobytes() + sq[:n].tobytes(), n, keys[:n].copy(), gidx[:n].copy(), sq[:n].copy())
                if (ll, b) in prev:
                    pv, _, pn, pk, pg, ps = prev[(ll, b)]
                    if pv == cur[0]:
                        assert n >= pn, (i, ll, b)
                        assert pk.tobytes() == cur[3][:pn].tobytes() and np.array_equal(pg, cur[4][:pn]) and np.array_equal(ps, cur[5][:pn]), \
                            f"scan {i} layer {ll} bucket {b}: the tree changed in place without a restructure bump"
                        appends += int(n > pn)
                    else:
                        bumps += 1
                prev[(ll, b)] = cur
    assert bumps > 5 and appends > 100, (bumps, appends)  # both paths of the mirror were exercised
    built_lib.c2g_hostdb_free(h)
