#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_more.py tests/test_gpu_fullsize.py tests/test_gpu_modules.py -x -q -p no:cacheprovider -k "cin or xdeepfm or mlp or dense or dnn or senet" > gpurun_out/r2_tests_fold.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_fold.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_fold.log | head
python tools/cin_profile_driver.py 65536
TRS_CIN_NO_FOLD=1 python tools/cin_profile_driver.py 65536
