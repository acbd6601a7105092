"""GPU kNN over a LARGE key table (20 000 keys per q-level, 600+ kd blocks, all six buckets populated) against the oracle's
LayerDB::layerKNNSearch (include/cont2/contour_db.h:150-192, src/cont2/contour_db.cpp:319-403) on the same table:
hint identity, order and squared distances bit-exact.  Exercises the box pruning of the blocked mirror (seed block, sweep,
bound tightening), clustered near-duplicates, exact duplicate keys (distance ties) and far random keys."""
import ctypes as C

import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from helpers import make_batch

pytestmark = pytest.mark.gpu

N_KEYS = 20000
N_SCANS = 8


def _dist_ub(k):
    """dist_ub of ContourDB::queryRangedKNN (contour_db.h:733-749): bounds are doubles rounded to float, the rest is float."""
    f32, f64 = np.float32, np.float64
    b = [(f32(f64(k[0]) * 0.8), f32(f64(k[0]) / 0.8)), (f32(f64(k[1]) * 0.8), f32(f64(k[1]) / 0.8)),
         (f32((f64(k[2]) * 0.8) * 0.75), f32(f64(k[2]) / (0.8 * 0.75)))]
    tot = None
    for d in range(3):
        lo, hi = f32(k[d] - b[d][0]), f32(k[d] - b[d][1])
        m = max(f32(lo * lo), f32(hi * hi))
        tot = m if tot is None else f32(tot + m)
    return tot


def test_blocked_knn_matches_layer_search(built_lib, oracle):
    from contour_context_b200.engine import Engine

    eng = Engine(scan_capacity=4096, max_batch=16, max_points=16 * 65536)
    dbc = eng.db_cfg
    pts, offsets = make_batch(list(range(40, 40 + N_SCANS)), [0] * N_SCANS, 60000, noise_seed=5)
    eng.ingest(pts, offsets, first_slot=0, int_ids=np.arange(N_SCANS))
    heads = eng.heads(0, N_SCANS)
    rng = np.random.default_rng(1234)
    odb = oracle.DB(dbc)
    L = oracle.lib()
    L.c2o_test_fill_layer2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    tables = []
    for ll in range(dbc.n_q_levels):
        lev = dbc.q_levels[ll]
        keys, gidx, seq = [], [], []
        for g in range(N_SCANS):
            for s in range(D.MAX_PIV):
                k = heads["keys"][g][lev][s].astype(np.float32)
                if float(k.sum()) == 0.0:
                    continue
                for sigma in (1e-3, 1e-2, 0.05, 0.2):
                    noisy = (k[None, :] * (1.0 + sigma * rng.standard_normal((40, D.KEY_DIM)))).astype(np.float32)
                    keys.append(noisy)
                    gidx += [g] * 40
                    seq += [s] * 40
                    dup = np.repeat(noisy[:2], 15, axis=0)  # exact duplicates: distance ties inside and across blocks
                    keys.append(dup)
                    gidx += [g] * 30
                    seq += [s] * 30
        keys = np.concatenate(keys)
        n_fill = N_KEYS - len(keys)
        assert n_fill > 1000
        lo, hi = keys.min(axis=0), keys.max(axis=0)
        keys = np.concatenate([keys, (lo + (hi - lo) * rng.random((n_fill, D.KEY_DIM))).astype(np.float32)])
        gidx = np.array(gidx + [0] * n_fill, np.int32)
        seq = np.array(seq + [0] * n_fill, np.int8)
        order = rng.permutation(N_KEYS)  # tree order is arbitrary with respect to the key values
        keys, gidx, seq = np.ascontiguousarray(keys[order]), gidx[order], seq[order]
        qs = np.quantile(keys[:, 0], [1 / 6, 2 / 6, 3 / 6, 4 / 6, 5 / 6]).astype(np.float32)
        ranges = np.concatenate([[-1000.0], qs, [1000.0]]).astype(np.float32)
        bucket = np.searchsorted(ranges[1:-1], keys[:, 0], side="right").astype(np.uint8)
        assert len(np.unique(bucket)) == D.NUM_BUCKETS
        eng.db_set_layer(ll, keys, gidx, seq, bucket, ranges)
        L.c2o_test_fill_layer2(odb.h, ll, keys.ctypes.data_as(C.c_void_p), gidx.ctypes.data_as(C.c_void_p),
                               seq.ctypes.data_as(C.c_void_p), bucket.ctypes.data_as(C.c_void_p), N_KEYS,
                               ranges.ctypes.data_as(C.c_void_p))
        tables.append((keys, gidx, seq))
    lb, ub = D.kitti_thres()
    _, hints, _ = eng.query(0, N_SCANS, lb, ub, want_trace=True)
    nnk = dbc.nnk
    n_checked = n_full = 0
    for q in range(N_SCANS):
        for ll in range(dbc.n_q_levels):
            lev = dbc.q_levels[ll]
            for s in range(D.MAX_PIV):
                k = heads["keys"][q][lev][s].astype(np.float32)
                base = ((q * dbc.n_q_levels + ll) * D.MAX_PIV + s) * nnk
                gh = hints[base:base + nnk]
                gh = gh[gh["cand_gidx"] >= 0]
                if float(k.sum()) == 0.0:
                    assert len(gh) == 0
                    continue
                og, osq, od = odb.layer_knn(ll, k, nnk, float(_dist_ub(k)))
                assert len(gh) == len(og), (q, ll, s, len(gh), len(og))
                assert gh["dist_sq"].tobytes() == od.tobytes(), (q, ll, s)
                assert np.array_equal(gh["cand_gidx"], og) and np.array_equal(gh["cand_seq"], osq), (q, ll, s)
                n_checked += 1
                n_full += len(og) == nnk
    assert n_checked >= 100 and n_full >= 100
    eng.close()
