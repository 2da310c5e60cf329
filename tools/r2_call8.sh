#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py tests/test_gpu_modules.py tests/test_gpu_fullsize.py -q -p no:cacheprovider > gpurun_out/r2_tests_sess.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_sess.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_sess.log | head -20
