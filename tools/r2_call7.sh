#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r2_tests_all.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_all.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_all.log | head -20
