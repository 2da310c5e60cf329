#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size --clock-control none -c 120 --csv --log-file gpurun_out/r2_mlp400_launches.csv python tools/bench_ops.py --only deepfm_generic_mlp400 > gpurun_out/r2_mlp400.log 2>&1
echo rc=$?
