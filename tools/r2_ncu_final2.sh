#!/bin/bash
# closing ncu evidence of this session: launch list of the bench command, --set full of the CIN layers and of the gathering dense layer
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2b_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cin_tc_layer_kernel -s 2 -c 2 -f -o gpurun_out/r2b_cin python tools/cin_profile_driver.py 65536 > gpurun_out/r2b_ncu_cin.log 2>&1
echo "cin rc=$?"
timeout 120 ncu -i gpurun_out/r2b_cin.ncu-rep --page raw --csv > gpurun_out/r2b_cin_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2b_cin_raw.csv > gpurun_out/r2b_cin_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cin_tc_layer_kernel -s 3 -c 3 -f -o gpurun_out/r2b_dense python tools/bench_ops.py --only deepfm_generic_mlp400 > gpurun_out/r2b_ncu_dense.log 2>&1
echo "dense rc=$?"
timeout 120 ncu -i gpurun_out/r2b_dense.ncu-rep --page raw --csv > gpurun_out/r2b_dense_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2b_dense_raw.csv > gpurun_out/r2b_dense_summary.txt
grep -E "==|duration|tensor_cycles|dram__bytes_read" gpurun_out/r2b_cin_summary.txt gpurun_out/r2b_dense_summary.txt
TRS_CIN_TRACE=1 timeout 120 python tools/cin_profile_driver.py 65536 2>&1 | grep -E "cin trace" | awk '{k=$4" "$6; if (n[k]++ == 2) print}' > gpurun_out/r2b_cin_trace.txt
TRS_DENSE_TRACE=1 timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>&1 | grep "dense trace" | awk '{k=$4" "$6" "$12" "$14; if (n[k]++ == 3) print}' > gpurun_out/r2b_dense_trace.txt
cat gpurun_out/r2b_cin_trace.txt gpurun_out/r2b_dense_trace.txt
