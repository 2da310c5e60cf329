#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -p no:cacheprovider -k "deepfm_wide_mlp" > gpurun_out/r2_tests_f1.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_f1.log)"
for e in "" "TRS_GATHER_NO_MEET=1"; do
echo "== $e"
env $e timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*' | sed 's/"op": "\(.\{12\}\)[^"]*"/\1/' | head -1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cin_tc_layer_kernel -s 3 -c 1 python tools/bench_ops.py --only deepfm_generic_mlp400 2>&1 | grep -E "duration|dram__bytes|hit_rate"
