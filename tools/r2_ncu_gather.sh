#!/bin/bash
# ncu --set full of the gathering dense layer (first cin_tc_layer_kernel of a DeepFM-400 forward)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cin_tc_layer_kernel -s 3 -c 1 -f -o gpurun_out/r2_dense_gather python tools/bench_ops.py --only deepfm_generic_mlp400 > gpurun_out/r2_ncu_gather.log 2>&1
echo "rc=$?"
timeout 120 ncu -i gpurun_out/r2_dense_gather.ncu-rep --page raw --csv > gpurun_out/r2_dense_gather_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/r2_dense_gather.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_dense_gather_sass.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_dense_gather_raw.csv | head -40
