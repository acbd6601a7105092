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


@pytest.mark.parametrize("mode", ["continuous", "quantised", "bursty"])
def test_rebalance_matches_oracle(built_lib, oracle, mode):
    rng = np.random.default_rng({"continuous": 1, "quantised": 2, "bursty": 3}[mode])
    cfg = D.kitti_db_config()
    odb = oracle.DB(cfg)
    olib = oracle.lib()
    olib.c2o_test_push_key.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int]
    h = built_lib.c2g_hostdb_create(cfg.n_q_levels, cfg.max_elapse, cfg.min_elapse)
    n_scans = 1500
    for i in range(n_scans):
        ts = 0.104 * i if mode != "bursty" else 0.104 * i + (30.0 if (i // 200) % 2 else 0.0)
        for ll in range(cfg.n_q_levels):
            for seq in range(6):
                key = (rng.random(10) * 30 + 1).astype(np.float32)
                if mode == "quantised":
                    key[0] = np.float32(np.round(key[0] / 2.0) * 2.0)  # long strips of equal bucket values
                if rng.random() < 0.05:
                    key[:] = 0  # all-zero keys are never stored
                olib.c2o_test_push_key(odb.h, ll, _p(key), ts, i, seq)
                assert built_lib.c2g_hostdb_push_key(h, ll, _p(key), ts, i, seq) == 0
        odb.push_and_balance(i, ts)
        assert built_lib.c2g_hostdb_balance(h, i, ts) == 0
        if i % 97 == 0 or i == n_scans - 1:
            for ll in range(cfg.n_q_levels):
                o_rng, o_ts, o_bs = odb.layer_state(ll)
                p_rng, p_ts, p_bs = _state_prod(built_lib, h, ll)
                assert o_rng.tobytes() == p_rng.tobytes(), (mode, i, ll)
                assert np.array_equal(o_ts, p_ts) and np.array_equal(o_bs, p_bs), (mode, i, ll, o_ts, p_ts)
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
    built_lib.c2g_hostdb_free(h)
