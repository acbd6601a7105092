"""GPU: the C++ facade (ContourManager / ContourDB with the reference's names) driven by the ROS-free batch driver
reproduces, scan by scan, what the Python mirror of the same C-ABI computes for the reference's online loop
(test/batch_bin_test.cpp:105-247: make bev -> queryRangedKNN -> addScan -> pushAndBalance)."""
import os
import re
import subprocess

import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batch_driver_matches_python_mirror(built_lib, tmp_path):
    from contour_context_b200.engine import Engine

    exe = os.path.join(ROOT, "contour_context_b200", "host", "cont2_batch_bin")
    assert os.path.exists(exe), "host facade not built (run __graft_entry__.build())"
    # 3 scenes visited 4 times each, interleaved; 30 s between scans so that earlier scans are searchable (25 s gate)
    order = [(s, v) for v in range(4) for s in (40, 41, 42)]
    pts = synth.make_scans([s for s, _ in order], [v for _, v in order], 60000).numpy()
    lines = []
    for i in range(len(order)):
        f = tmp_path / f"{i:06d}.bin"
        pts[i].astype(np.float32).tofile(f)
        lines.append(f"{30.0 * i} {f}")
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(lines) + "\n")
    env = dict(os.environ, C2G_SCAN_CAPACITY="256")
    out = subprocess.run([exe, str(lst)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = {}
    for m in re.finditer(r"LC (\d+) -> (\d+|none)(?: corr ([0-9.]+))?", out.stdout):
        got[int(m.group(1))] = (None, None) if m.group(2) == "none" else (int(m.group(2)), float(m.group(3)))
    assert len(got) == len(order)

    eng = Engine(scan_capacity=64, max_batch=1, max_points=1 << 18)
    lb, ub = D.kitti_thres()
    n_pos = 0
    for i in range(len(order)):
        p = np.ascontiguousarray(pts[i])
        eng.ingest(p, np.array([0, p.shape[0]], np.int64), first_slot=40, int_ids=[i])
        res = eng.query(40, 1, lb, ub)[0]
        if res["n_cand"] > 0:
            n_pos += 1
            assert got[i][0] == int(res["cand"][0]["cand_gidx"]), (i, got[i], res["cand"][0]["cand_gidx"])
            assert abs(got[i][1] - float(res["cand"][0]["corr_fine"])) < 2e-6
        else:
            assert got[i][0] is None, (i, got[i])
        eng.copy_slots(40, i, 1)
        eng.db_add_scans(i, 1, [30.0 * i])
        eng.db_push_and_balance(i, 30.0 * i)
    assert n_pos >= 3, "the revisits should produce loop closures"
    eng.close()
