#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -x -q -p no:cacheprovider -k "cin or xdeepfm or mlp or deepfm_wide" > gpurun_out/r2_tests_cin.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_cin.log)"
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -p no:cacheprovider -k "xdeepfm" > gpurun_out/r2_tests_cin_full.log 2>&1
echo "fullsize rc=$? $(tail -1 gpurun_out/r2_tests_cin_full.log)"
timeout 120 python tools/cin_profile_driver.py 65536 2>&1 | tail -1
TRS_CIN_TRACE=1 timeout 120 python tools/cin_profile_driver.py 65536 2>&1 | grep -E "cin trace" | awk '{k=$4" "$6; if (n[k]++ == 2) print}'
timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*' | sed 's/"op": "\(.\{12\}\)[^"]*"/\1/'
