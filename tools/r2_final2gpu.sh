#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_multi.log 2>&1
echo "multi-gpu tests rc=$? $(tail -1 gpurun_out/r2_tests_multi.log)"
bash tools/r2_final8.sh 2
