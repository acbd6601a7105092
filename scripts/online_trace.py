"""Developer aid: timeline of the pipelined windowed loop (C2G_TRACE marks on the copy stream and the kernel stream).
usage: python scripts/online_trace.py [n_windows] [xyz]   (needs a GPU; prints per-window device times in ms)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TRACE = "/tmp/c2g_trace.txt"
os.environ["C2G_TRACE"] = TRACE
if os.path.exists(TRACE):
    os.remove(TRACE)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from contour_context_b200 import ctypes_defs as D  # noqa: E402
from contour_context_b200 import synth  # noqa: E402
from contour_context_b200.engine import Engine  # noqa: E402

g.build_c2g()
NW = int(sys.argv[1]) if len(sys.argv) > 1 else 8
xyz = len(sys.argv) > 2
W, n_pts = 148, 120000
n = NW * W
seeds, visits = synth.db_layout(n, 4, first_scene=0)
order = np.argsort(np.asarray(visits), kind="stable")
seeds, visits = [seeds[i] for i in order], [visits[i] for i in order]
fpp = 3 if xyz else 4
host = torch.empty((n * n_pts, fpp), dtype=torch.float32, pin_memory=True)
for i0 in range(0, n, W):
    blk = synth.make_scans(seeds[i0:i0 + W], visits[i0:i0 + W], n_pts, device="cuda", noise_seed=i0).reshape(-1, 4)
    host[i0 * n_pts:(i0 + W) * n_pts].copy_(blk[:, :fpp])
torch.cuda.synchronize()
offsets = np.arange(W + 1, dtype=np.int64) * n_pts
ts = 0.104 * np.arange(n)
ids = np.arange(n, dtype=np.int32)
lb, ub = D.kitti_thres()
res = torch.empty(n * D.QUERY_RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(D.QUERY_RESULT_DTYPE)
for rep in range(2):
    if os.path.exists(TRACE):
        os.remove(TRACE)
    eng = Engine(device=0, scan_capacity=n + 8, max_batch=W, max_points=W * n_pts)
    stage = lambda k: eng.online_stage(host[k * W * n_pts:(k + 1) * W * n_pts], offsets, int_ids=ids[k * W:(k + 1) * W], on_device=False, xyz=xyz)  # noqa: E731
    stage(0)
    for k in range(NW):
        if k + 1 < NW:
            stage(k + 1)
        eng.online_commit(ts[k * W:(k + 1) * W], ids[k * W:(k + 1) * W], lb, ub, res[k * W:(k + 1) * W])
    eng.sync()
    eng.close()
print(open(TRACE).read())
