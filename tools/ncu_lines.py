"""Rank CUDA source lines of an .ncu-rep (captured with --import-source on, built with -lineinfo) by warp-stall samples.
usage: python tools/ncu_lines.py <report.ncu-rep> [kernel-substring] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ksel = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, cur_fn, hdr, agg = None, None, None, {}
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or (ksel and ksel not in (cur_fn or "")):
        continue
    if r[2] != "-":
        continue
    try:
        line, smp, inst = int(r[0]), int(r[6] or 0), int(r[7] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur_file, line), [0, 0, r[1].strip()[:105]])
    a[0] += smp
    a[1] += inst
ts, ti = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print(f"samples {ts} instructions {ti}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{a[0] / ts * 100:5.1f}% smp {a[1] / ti * 100:5.1f}% inst {k[0]}:{k[1]:<4} {a[2]}")
