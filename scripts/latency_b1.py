"""Single-scan latency of the C-ABI calls (B = 1, host input): what the online loop pays per scan."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contour_context_b200 import synth, ctypes_defs as D
from contour_context_b200.engine import Engine

N = 120000
eng = Engine(scan_capacity=600, max_batch=8, max_points=8 * 131072)
pts = synth.make_scans(list(range(100, 164)), [0] * 64, N, device="cuda").cpu()
pin = torch.empty((N, 4), dtype=torch.float32, pin_memory=True)
offsets = np.array([0, N], np.int64)
lb, ub = D.kitti_thres()
def t(fn, reps=30):
    ts = []
    for i in range(reps):
        t0 = time.perf_counter(); fn(i); ts.append((time.perf_counter() - t0) * 1e3)
    return np.median(ts[3:]), np.min(ts[3:])
def ing_pinned(i):
    pin.copy_(pts[i % 64]); 
def ing_only(i):
    eng.ingest(pin, offsets, first_slot=500, on_device=False); eng.sync()
def ing_pageable(i):
    eng.ingest(pts[i % 64].numpy(), offsets, first_slot=500, on_device=False); eng.sync()
def heads(i):
    eng.heads(500, 1)
print("host memcpy 1.92 MB into pinned      : median %.3f ms (min %.3f)" % t(ing_pinned))
print("c2g_ingest(pinned, B=1) + sync       : median %.3f ms (min %.3f)" % t(ing_only))
print("c2g_ingest(pageable, B=1) + sync     : median %.3f ms (min %.3f)" % t(ing_pageable))
print("c2g_get_heads(1)                     : median %.3f ms (min %.3f)" % t(heads))
# grow a DB of 400 scans, then time the query of one scan
for i0 in range(0, 400, 8):
    b = synth.make_scans(list(range(1000 + i0 // 4, 1000 + i0 // 4 + 2)) * 4, [0, 0, 1, 1, 2, 2, 3, 3], N, device="cuda").reshape(-1, 4)
    torch.cuda.synchronize()
    eng.ingest(b, np.arange(9, dtype=np.int64) * N, first_slot=i0, on_device=True); eng.sync()
    for j in range(8):
        eng.db_add_scans(i0 + j, 1, np.array([1.0 * (i0 + j)])); eng.db_push_and_balance(i0 + j, 1.0 * (i0 + j))
def q(i):
    eng.ingest(pin, offsets, first_slot=500, on_device=False)
    eng.query(500, 1, lb, ub)
def q_dirty(i):
    eng.db_push_and_balance(i, 1000.0 + i)   # marks the mirror dirty: the next query re-syncs the layer tables
    eng.query(500, 1, lb, ub)
def q_only(i):
    eng.query(500, 1, lb, ub)
print("c2g_query(B=1), clean mirror         : median %.3f ms (min %.3f)" % t(q_only))
print("push_and_balance + c2g_query(B=1)    : median %.3f ms (min %.3f)" % t(q_dirty))
print("c2g_ingest + c2g_query (B=1)         : median %.3f ms (min %.3f)" % t(q))
eng.close()
