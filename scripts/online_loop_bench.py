"""Online (scan-by-scan) loop through the C++ facade: cont2_batch_bin over N synthetic 120k-point KITTI-format .bin files
(query each scan against the growing DB, then add it: BASELINE.json configs[1] shape).  Prints the driver's stage table and
the per-scan wall time.   usage: python scripts/online_loop_bench.py [n_scans] [tmpdir]"""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contour_context_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    tmp = sys.argv[2] if len(sys.argv) > 2 else "/tmp/c2g_online"
    os.makedirs(tmp, exist_ok=True)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n_scenes = max(1, n // 4)  # visit-major: every scene is revisited after n/4 scans
    seeds = [500 + i % n_scenes for i in range(n)]
    visits = [i // n_scenes for i in range(n)]
    lines = []
    for i0 in range(0, n, 50):
        pts = synth.make_scans(seeds[i0:i0 + 50], visits[i0:i0 + 50], 120000, dev, noise_seed=i0).cpu().numpy()
        for j in range(pts.shape[0]):
            f = os.path.join(tmp, "%06d.bin" % (i0 + j))
            pts[j].astype(np.float32).tofile(f)
            lines.append("%f %s" % (0.104 * (i0 + j) * 10, f))  # 1.04 s apart: every key becomes searchable quickly
    lst = os.path.join(tmp, "list.txt")
    open(lst, "w").write("\n".join(lines) + "\n")
    exe = os.path.join(ROOT, "contour_context_b200", "host", "cont2_batch_bin")
    t0 = time.time()
    out = subprocess.run([exe, lst], capture_output=True, text=True, env=dict(os.environ, C2G_SCAN_CAPACITY=str(n + 64)))
    dt = time.time() - t0
    print(out.stdout[-1800:])
    print("wall: %.1f s for %d scans = %.2f ms / scan (includes process start-up and file reads)" % (dt, n, dt / n * 1e3))
    if out.returncode != 0:
        print(out.stderr[-2000:])
        sys.exit(1)


if __name__ == "__main__":
    main()
