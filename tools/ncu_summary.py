"""Summarise `ncu --page raw --csv` dumps into the few metrics the roofline discussion uses (developer tool).
usage: ncu_summary.py out.csv raw1.csv [raw2.csv ...]"""
import csv, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warp_latency_issue_stalled_barrier.pct",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.pct", "smsp__average_warp_latency_issue_stalled_wait.pct"]
out = csv.writer(open(sys.argv[1], "w"))
out.writerow(["capture", "kernel", "metric", "value", "unit"])
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        for w in WANT:
            if w in idx:
                out.writerow([path.split("/")[-1], name, w, r[idx[w]], units[idx[w]]])
