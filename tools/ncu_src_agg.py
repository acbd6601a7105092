"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_src_agg.py dump.csv [top_n] [file:lo-hi,...|-] [kernel substring]   (developer tool; prints samples / instructions per
source line, by file; a dump of several kernels is filtered by the substring of the function name)"""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
cur_file = None
cur_fn = ''
want_fn = sys.argv[4] if len(sys.argv) > 4 else ''
hdr = None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if want_fn and want_fn not in cur_fn: continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "" or hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], r[4:]))
    def f(k):
        try: return float(d.get(k, 0))
        except ValueError: return 0.0
    stalls = {k[6:]: f(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
    rows.append((cur_file, ln, r[1][:90], f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed"), stalls, f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Excessive")))
tot = sum(x[3] for x in rows); toti = sum(x[4] for x in rows)
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
byfile = collections.Counter()
for x in rows: byfile[x[0]] += x[3]
print("by file:", dict(byfile))
print(f"{'file':18s} {'line':>5s} {'samp%':>6s} {'inst%':>6s} {'thr/inst':>8s} top stalls | source")
for x in sorted(rows, key=lambda x: -x[3])[:top]:
    st = sorted(x[6].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{k}:{v:.0f}" for k, v in st if v > 0)
    print(f"{x[0]:18s} {x[1]:5d} {100*x[3]/tot:6.2f} {100*x[4]/toti:6.2f} {x[5]/max(x[4],1):8.1f} {sts:40s} | {x[2].strip()}")
if len(sys.argv) > 3 and sys.argv[3] != '-':  # ranges: file:lo-hi,...
    for spec in sys.argv[3].split(","):
        fn, rg = spec.split(":"); lo, hi = map(int, rg.split("-"))
        s = sum(x[3] for x in rows if x[0] == fn and lo <= x[1] <= hi); i = sum(x[4] for x in rows if x[0] == fn and lo <= x[1] <= hi)
        print(f"{spec:30s} samples {100*s/tot:6.2f}%  inst {100*i/toti:6.2f}%")
