#!/bin/bash
# ncu --set full capture of the two ingest kernels (one launch each, 592-scan batch), source-level stalls included
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$2" -s 4 -c 1 -f -o gpurun_out/$1 python scripts/quick_ingest_bench.py 592 > gpurun_out/$1.log 2>&1
ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$1_src.csv 2>/dev/null
tail -5 gpurun_out/$1.log
