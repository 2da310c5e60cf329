#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"afm_tc5_kernel|afm_finish_kernel" -s 4 -c 2 -f -o gpurun_out/r2_afm_tc5 python tools/bench_ops.py --only afm > gpurun_out/r2_ncu_afm.log 2>&1
echo "rc=$?"
timeout 120 ncu -i gpurun_out/r2_afm_tc5.ncu-rep --page raw --csv > gpurun_out/r2_afm_tc5_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/r2_afm_tc5.ncu-rep --page source --csv --print-source sass --kernel-name regex:afm_tc5_kernel > gpurun_out/r2_afm_tc5_sass.csv 2>/dev/null
