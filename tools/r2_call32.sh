#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_session.py tests/test_gpu_training.py -x -q -p no:cacheprovider -k "cin or xdeepfm or deepfm or fm" > gpurun_out/r2_tests_x.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_x.log)"
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -p no:cacheprovider -k "xdeepfm" > gpurun_out/r2_tests_x_full.log 2>&1
echo "fullsize rc=$? $(tail -1 gpurun_out/r2_tests_x_full.log)"
timeout 120 python tools/cin_profile_driver.py 65536 2>&1 | tail -1
timeout 200 python tools/bench_ops.py --only xdeepfm 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*'
