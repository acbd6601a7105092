"""Developer timing of the ingest kernels on device-resident input (CUDA events on the context stream)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contour_context_b200 import synth
from contour_context_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = 120000
eng = Engine(scan_capacity=B + 8, max_batch=B, max_points=1024)
chunks = []
for i0 in range(0, B, 74):
    n = min(74, B - i0)
    seeds, visits = synth.db_layout(n, 4, first_scene=i0 // 4)
    chunks.append(synth.make_scans(seeds, visits, N, device="cuda", noise_seed=i0))
pts = torch.cat(chunks).reshape(-1, 4).contiguous()
torch.cuda.synchronize()
offsets = np.arange(B + 1, dtype=np.int64) * N
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(st):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            fn()
            e1.record(st)
        st.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
for _ in range(3):
    eng.ingest(pts, offsets, 0); eng.sync()
k1 = timeit(lambda: eng.ingest_bev_only(pts, offsets))
full = timeit(lambda: eng.ingest(pts, offsets, 0))
bytes_ = B * N * 16
print(f"B={B} K1 bev_scatter: min {k1[0]:.3f} ms median {k1[1]:.3f} ms -> {bytes_/k1[0]/1e6:.1f} GB/s, {B/k1[0]*1e3:.0f} scans/s")
print(f"B={B} K1+K2 ingest : min {full[0]:.3f} ms median {full[1]:.3f} ms -> {bytes_/full[0]/1e6:.1f} GB/s, {B/full[0]*1e3:.0f} scans/s; K2 alone ~{full[0]-k1[0]:.3f} ms")
h = eng.heads(0, B)
print("status", np.unique(h["status"]), "views/level mean", h["n_views"].mean(0))
import ctypes as C
from contour_context_b200 import capi
clk = np.zeros(64, np.int64)
capi.lib().c2g_debug_clocks.argtypes = [C.c_void_p, C.c_void_p]
capi.lib().c2g_debug_clocks(eng.h, clk.ctypes.data_as(C.c_void_p))
names = {0:'start',1:'A done',2:'label+stats',3:'ranks done',4:'sort+moments',5:'calcstat+copy',6:'D1 done',7:'D2+keys done',8:'BCI done',9:'GMM/end'}
t0 = clk[0]
for i in range(10): print(f"  {names[i]:14s} {(clk[i]-t0)/1e3:9.1f} kcyc  (+{(clk[i]-clk[max(i-1,0)])/1e3:.1f})")
fine = {10:'B1 count+barrier',12:'decide+B2 records',13:'B3 unions',14:'B4 flatten',15:'barrier+decide',16:'B5 slots',2:'B6 stats+barrier',18:'B7 parents',19:'B8 keys+barrier'}
prev = clk[1]
for i in [10,12,13,14,15,16,2,18,19]:
    print(f"    {fine[i]:18s} +{(clk[i]-prev)/1e3:.1f} kcyc"); prev = clk[i]
print(f"    after-moments barrier -> lambda exit +{(clk[24]-clk[4])/1e3:.1f}; view copy +{(clk[22]-clk[24])/1e3:.1f}; ellipse table +{(clk[23]-clk[22])/1e3:.1f}; top + barrier +{(clk[5]-clk[23])/1e3:.1f} kcyc")
print(f"    ranks +{(clk[3]-clk[19])/1e3:.1f}; torder+sort(level 0) +{(clk[20]-clk[3])/1e3:.1f}; warp0 moments done +{(clk[21]-clk[20])/1e3:.1f}; barrier +{(clk[4]-clk[21])/1e3:.1f} kcyc")

print("per-warp arrival at the after-moments barrier (kcyc after 'ranks done'):", [round(float(clk[32 + w] - clk[3]) / 1e3, 1) for w in range(16)])
print("per-warp departure from it                                            :", [round(float(clk[48 + w] - clk[3]) / 1e3, 1) for w in range(16)])
print("DBG(4), DBG(24) rel:", round(float(clk[4]-clk[3])/1e3,1), round(float(clk[24]-clk[3])/1e3,1))
