"""CPU-only: what pins the oracle (the reference ships no executable golden vectors for this path):
 * CCL label ORDER and stats against the real OpenCV (cv2.connectedComponentsWithStats, 8-connectivity);
 * the restated Eigen 2x2 self-adjoint eigen solver against numpy.linalg.eigh;
 * kNN result sets against the reference's own vendored nanoflann (oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_ccl_label_order_matches_opencv(oracle):
    cv2 = pytest.importorskip("cv2")
    L = oracle.lib()
    rng = np.random.default_rng(0)
    shapes = [(int(rng.integers(1, 40)), int(rng.integers(1, 40))) for _ in range(300)] + [(150, 150)] * 10 + [(1, 1), (2, 3), (3, 2)]
    for h, w in shapes:
        m = (rng.random((h, w)) < rng.uniform(0.05, 0.8)).astype(np.uint8) * 255
        n_cv, lab_cv, st_cv, _ = cv2.connectedComponentsWithStats(m, connectivity=8, ltype=cv2.CV_32S)
        lab = np.zeros((h, w), np.int32)
        st = np.zeros((h * w + 2, 5), np.int32)
        n = L.c2o_ccl8(_p(m), h, w, _p(lab), _p(st), h * w + 2)
        assert n == n_cv - 1
        assert np.array_equal(lab, lab_cv), (h, w)
        assert np.array_equal(st[1:n + 1], st_cv[1:]), (h, w)


def test_eig2f_against_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(5000):
        A = rng.normal(size=(2, 2)) * rng.uniform(0.05, 60)
        S = (A @ A.T).astype(np.float32)
        ev = np.zeros(2, np.float32)
        vec = np.zeros(4, np.float32)
        L.c2o_eig2f(float(S[0, 0]), float(S[0, 1]), float(S[1, 1]), _p(ev), _p(vec))
        w, _ = np.linalg.eigh(S.astype(np.float64))
        scale = max(abs(w).max(), 1e-30)
        assert ev[0] <= ev[1]
        assert np.abs(ev - w).max() / scale < 2e-6
        V = vec.reshape(2, 2).T
        assert np.abs(S.astype(np.float64) @ V - V * ev[None, :]).max() / scale < 4e-6
        assert abs(np.linalg.det(V.astype(np.float64))) == pytest.approx(1.0, abs=1e-5)
    # diagonal and degenerate inputs
    for a, b, c in [(2.0, 0.0, 1.0), (1.0, 0.0, 1.0), (0.0, 0.0, 0.0), (5.0, 5.0, 5.0)]:
        ev = np.zeros(2, np.float32)
        vec = np.zeros(4, np.float32)
        L.c2o_eig2f(a, b, c, _p(ev), _p(vec))
        w = np.linalg.eigvalsh(np.array([[a, b], [b, c]]))
        assert np.allclose(ev, w, atol=1e-6)


def test_knn_matches_reference_nanoflann(oracle):
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref/libref_knn.so not built (reference tree absent and no prebuilt copy)")
    from contour_context_b200 import ctypes_defs as D

    rng = np.random.default_rng(2)
    for n in (1, 7, 49, 50, 51, 400, 3000):
        keys = (rng.random((n, 10)) * rng.uniform(1, 40)).astype(np.float32)
        tree = R.ref_knn_build(_p(keys), n)
        # oracle DB with a single bucket holding the same keys in the same order
        db = oracle.DB(D.kitti_db_config())
        # push keys through the LayerDB buffer of q-level 0 and pop them into the tree
        lib = oracle.lib()
        lib.c2o_test_fill_layer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.c2o_test_fill_layer(db.h, 0, _p(keys), n)
        for _ in range(20):
            q = (rng.random(10) * 40).astype(np.float32)
            for k, maxd in ((50, 1e6), (50, 30.0), (5, 200.0)):
                ridx = np.zeros(k, np.int64)
                rdist = np.zeros(k, np.float32)
                R.ref_knn_search(tree, _p(q), k, C.c_float(maxd), _p(ridx), _p(rdist))
                valid = rdist < np.float32(maxd)
                gidx, seq, dist = db.layer_knn(0, q, k, maxd)
                assert len(dist) == int(valid.sum()), (n, k, maxd)
                assert dist.tobytes() == rdist[valid].tobytes()
                # identical neighbours wherever distances are unique
                if len(dist):
                    uniq = np.concatenate([[True], np.diff(dist) != 0]) & np.concatenate([np.diff(dist) != 0, [True]])
                    assert np.array_equal(gidx[uniq], ridx[valid][uniq])
        R.ref_knn_free(tree)
