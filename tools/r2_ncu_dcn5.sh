#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dcn_tc5_kernel -s 2 -c 1 -f -o gpurun_out/r2_dcn_tc5 python tools/bench_ops.py --only dcn > gpurun_out/r2_ncu_dcn5.log 2>&1
echo "rc=$?"
timeout 120 ncu -i gpurun_out/r2_dcn_tc5.ncu-rep --page raw --csv > gpurun_out/r2_dcn_tc5_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/r2_dcn_tc5.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_dcn_tc5_sass.csv 2>/dev/null
