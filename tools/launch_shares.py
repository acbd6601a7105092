"""Per-kernel shares of one device-resident bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(developer tool).  usage: launch_shares.py launches.csv out.csv
A step starts at a bev_scatter_fast_kernel launch that follows the first knn_kernel launch (the DB build before it has no
queries) and ends before the next one."""
import csv
import collections
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
names = [r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0] for r in rows]
us = [float(r[-1]) / 1e3 for r in rows]
first_knn = names.index("knn_kernel")
starts = [i for i, n in enumerate(names) if n == "bev_scatter_fast_kernel" and i > first_knn]
assert len(starts) >= 2, "the list must hold at least one complete step after the first query"
a, b = starts[0], starts[1]
cnt, tot = collections.Counter(), collections.Counter()
for n, t in zip(names[a:b], us[a:b]):
    cnt[n] += 1
    tot[n] += t
total = sum(tot.values())
with open(sys.argv[2], "w") as f:
    f.write(f"# one device-resident step (launches {a}..{b - 1} of {sys.argv[1].split('/')[-1]}): per-launch times under ncu are cold-cache and\n")
    f.write("# serialised (the 4 query sub-batches overlap in a real run): compare SHARES with bench.py's kernel_ms, not absolute times\n")
    f.write("kernel,launches_in_step,mean_us,sum_us,share_of_step\n")
    for n in cnt:
        f.write(f"{n},{cnt[n]},{tot[n] / cnt[n]:.1f},{tot[n]:.1f},{tot[n] / total:.3f}\n")
    f.write(f"# total {total:.1f} us serialised\n")
print(open(sys.argv[2]).read())
