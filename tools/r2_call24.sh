#!/bin/bash
mkdir -p gpurun_out
TRS_CIN_PAIR=1 TRS_CIN_VERBOSE=1 timeout 90 python tools/cin_profile_driver.py 65536 2>&1 | sort | uniq -c | tail -4
echo "driver rc=$?"
TRS_CIN_PAIR=1 timeout 200 python -m pytest tests/test_gpu_fullsize.py -x -q -p no:cacheprovider -k "xdeepfm" > gpurun_out/r2_tests_pair.log 2>&1
echo "fullsize xdeepfm (pairs) rc=$? $(tail -1 gpurun_out/r2_tests_pair.log)"
timeout 90 python tools/cin_profile_driver.py 65536 2>&1 | tail -1
