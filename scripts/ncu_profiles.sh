#!/bin/bash
# Round-2 profile collection (run under gpurun, 1 GPU): ncu --set full captures of the ingest pair and of the query kernels with
# source-level stall samples, plus the launch list of a short bench run.  Outputs under gpurun_out/; summaries are made by
# tools/ncu_summary.py on the CPU box and committed under profiles/.
P=${1:-r2z}
mkdir -p gpurun_out
for k in bev_scatter_fast_kernel contour_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -f -o gpurun_out/${P}_$k python scripts/quick_ingest_bench.py 592 > gpurun_out/${P}_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"knn_kernel|prefilter_kernel|score_thread_kernel|finish_replay_kernel|finish_corr_kernel|refine_kernel" -s 18 -c 6 -f -o gpurun_out/${P}_query python scripts/quick_query_bench.py 5000 592 > gpurun_out/${P}_query.log 2>&1
# launch list of the timed steps: this repo's kernels only (the synthetic-scan generator is torch), DB build skipped (5 000 scans in
# batches: 2 launches each + mirror kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bev_scatter|contour_kernel|knn_kernel|prefilter_kernel|score_thread_kernel|finish_|refine_kernel|rank_kernel|mirror_" -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_launches.log 2>&1
for f in ${P}_bev_scatter_fast_kernel ${P}_contour_kernel ${P}_query; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${f}_src.csv 2>/dev/null
done
ls -la gpurun_out/${P}_* | head -20
