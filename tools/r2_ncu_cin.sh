#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cin_tc_layer_kernel -s 4 -c 2 -f -o gpurun_out/r2_cin_tc python tools/cin_profile_driver.py 65536 > gpurun_out/r2_ncu_cin.log 2>&1
echo "cin rc=$?"
timeout 120 ncu -i gpurun_out/r2_cin_tc.ncu-rep --page raw --csv > gpurun_out/r2_cin_tc_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/r2_cin_tc.ncu-rep --page source --csv --print-source sass --kernel-name regex:cin_tc_layer_kernel --launch-skip 1 --launch-count 1 > gpurun_out/r2_cin_tc_l1_sass.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/r2_cin_tc.ncu-rep --page source --csv --print-source sass --kernel-name regex:cin_tc_layer_kernel --launch-skip 0 --launch-count 1 > gpurun_out/r2_cin_tc_l0_sass.csv 2>/dev/null
python tools/cin_profile_driver.py 65536
ls -la gpurun_out/r2_cin_tc*
