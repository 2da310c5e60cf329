#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the bench command + --set full of the dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/r2_bench_launches.csv) lines"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deepfm_tc5_kernel -s 6 -c 1 -f -o gpurun_out/r2_deepfm_tc5 $B --no-configs --no-e2e > gpurun_out/r2_ncu_tc5.log 2>&1
echo "tc5 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dcn_tc_kernel|cin_tc_layer_kernel|ffm_interleaved_kernel" -c 6 -f -o gpurun_out/r2_configs $B --no-e2e > gpurun_out/r2_ncu_configs.log 2>&1
echo "configs rc=$?"
for n in r2_deepfm_tc5 r2_configs; do timeout 120 ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep | tail -4
