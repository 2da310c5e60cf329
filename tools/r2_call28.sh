#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -p no:cacheprovider -k "mlp_tensor_core_chain or deepfm_wide_mlp" > gpurun_out/r2_tests_f1.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_f1.log)"
timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*' | sed 's/"op": "\(.\{12\}\)[^"]*"/\1/'
TRS_DENSE_TRACE=1 timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>&1 | grep "dense trace" | awk '{k=$4" "$6" "$12" "$14; if (n[k]++ == 3) print}' 

