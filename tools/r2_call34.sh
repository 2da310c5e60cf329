#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_more.py -x -q -p no:cacheprovider -k "mlp or deepfm or cen or senet or fat or dnn or nfm or fnn" > gpurun_out/r2_tests_y.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_y.log)"
timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*' | sed 's/"op": "\(.\{12\}\)[^"]*"/\1/'
