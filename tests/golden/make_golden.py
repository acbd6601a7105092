"""Generates the committed golden fixtures under tests/golden/ (run in the build container, where python-cv2 and
/root/reference are present; the fixtures then travel with the repository).

 * ccl_opencv.npz   — REAL OpenCV outputs (cv2.connectedComponentsWithStats, connectivity 8, the call the reference makes at
                      src/cont2/contour_mng.cpp:298) for a set of masks: pins the label ORDER the DFS contour order depends on.
 * knn_nanoflann.npz — REAL outputs of the reference's vendored nanoflann (oracle/_ref/libref_knn.so, call sequence of
                      TreeBucket::knnSearch, src/cont2/contour_db.cpp:381-403) for random 10-D keys.
 * ingest_oracle.npz — frozen outputs of the CPU oracle for two small synthetic scans (regression pin of the restatement
                      itself; the reference ships no executable vector for this).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    import cv2

    from contour_context_b200 import ctypes_defs as D
    from contour_context_b200 import synth
    from oracle import c2o

    rng = np.random.default_rng(20240925)
    masks, labels, stats = [], [], []
    for (h, w) in [(1, 1), (2, 2), (3, 5), (7, 4), (16, 16), (23, 31), (40, 40), (64, 37)] + [(int(rng.integers(2, 50)), int(rng.integers(2, 50))) for _ in range(24)]:
        m = (rng.random((h, w)) < rng.uniform(0.1, 0.75)).astype(np.uint8) * 255
        n, lab, st, _ = cv2.connectedComponentsWithStats(m, connectivity=8, ltype=cv2.CV_32S)
        masks.append(m)
        labels.append(lab.astype(np.int32))
        stats.append(st.astype(np.int32))
    np.savez_compressed(os.path.join(HERE, "ccl_opencv.npz"), n=len(masks), opencv_version=cv2.__version__,
                        **{f"mask{i}": m for i, m in enumerate(masks)}, **{f"lab{i}": l for i, l in enumerate(labels)},
                        **{f"stat{i}": s for i, s in enumerate(stats)})

    R = c2o.ref_lib()
    assert R is not None, "oracle/_ref/libref_knn.so missing"
    keys = (rng.random((600, 10)) * 25).astype(np.float32)
    queries = (rng.random((12, 10)) * 25).astype(np.float32)
    tree = R.ref_knn_build(keys.ctypes.data_as(C.c_void_p), len(keys))
    idx = np.zeros((len(queries), 2, 50), np.int64)
    dist = np.zeros((len(queries), 2, 50), np.float32)
    maxd = np.array([1e6, 120.0], np.float32)
    for qi, q in enumerate(queries):
        for mi, md in enumerate(maxd):
            R.ref_knn_search(tree, q.ctypes.data_as(C.c_void_p), 50, C.c_float(md), idx[qi, mi].ctypes.data_as(C.c_void_p),
                             dist[qi, mi].ctypes.data_as(C.c_void_p))
    R.ref_knn_free(tree)
    np.savez_compressed(os.path.join(HERE, "knn_nanoflann.npz"), keys=keys, queries=queries, maxd=maxd, idx=idx, dist=dist)

    cfg = D.kitti_cm_config()
    pts = synth.make_scans([7, 7], [0, 2], 20000).numpy()
    out = {"pts": pts}
    for b in range(2):
        s = c2o.Scan(cfg, b).ingest(pts[b])
        h = s.head()
        out[f"keys{b}"] = h["keys"].copy()
        out[f"n_views{b}"] = h["n_views"].copy()
        out[f"cell_cnt{b}"] = h["layer_cell_cnt"].copy()
        out[f"views_l1_{b}"] = s.views(1).view(np.uint8).copy()
        out[f"bci_bits{b}"] = h["bcis"]["dist_bin"].copy()
    np.savez_compressed(os.path.join(HERE, "ingest_oracle.npz"), **out)
    print("golden fixtures written:", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
