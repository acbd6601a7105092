"""Developer timing of the query chain's kernels (CUDA events, c2g_query_profile) on a mid-sized DB: python scripts/quick_query_bench.py [n_db] [Q]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contour_context_b200 import ctypes_defs as D, synth
from contour_context_b200.engine import Engine

n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
N = 120000
eng = Engine(scan_capacity=n_db + Q + 8, max_batch=max(Q, 148), max_points=1024)
lb, ub = D.kitti_thres()
seeds, visits = synth.db_layout(n_db, 4)
for i0 in range(0, n_db, 148):
    n = min(148, n_db - i0)
    pts = synth.make_scans(seeds[i0:i0 + n], visits[i0:i0 + n], N, device="cuda", noise_seed=i0).reshape(-1, 4)
    torch.cuda.synchronize()
    eng.ingest(pts, np.arange(n + 1, dtype=np.int64) * N, first_slot=i0)
    eng.sync()
    for j in range(n):
        eng.db_add_scans(i0 + j, 1, [0.1 * (i0 + j)])
        eng.db_push_and_balance(i0 + j, 0.1 * (i0 + j))
for k in range(16):
    eng.db_push_and_balance(k, 0.1 * n_db + 525.0 + k)
for i0 in range(0, Q, 148):
    n = min(148, Q - i0)
    pts = synth.make_scans([i0 + k for k in range(n)], [4] * n, N, device="cuda", noise_seed=777 + i0).reshape(-1, 4)
    torch.cuda.synchronize()
    eng.ingest(pts, np.arange(n + 1, dtype=np.int64) * N, first_slot=n_db + i0)
eng.sync()
eng.query_profile(True)
acc = {}
for _ in range(4):
    eng.query_async(n_db, Q, lb, ub)
    for k, v in eng.query_profile(True, read=True).items():
        acc.setdefault(k, []).append(v)
eng.query_profile(False)
print({k: round(float(np.mean(v[1:])), 4) for k, v in acc.items()}, "sum", round(sum(float(np.mean(v[1:])) for v in acc.values()), 3))
st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        a.record(st); eng.query_async(n_db, Q, lb, ub); b.record(st)
    st.synchronize(); ts.append(a.elapsed_time(b))
print("query chain (4 sub-batches overlapped): %.3f ms" % float(np.mean(ts[1:])))
