#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -p no:cacheprovider -k "afm" > gpurun_out/r2_tests_afm5.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_afm5.log)"
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/r2_tests_afm5.log | head
timeout 120 python tools/bench_ops.py --only afm 2>&1 | grep '"op"' | cut -c1-140
TRS_DISABLE_TC5=1 timeout 120 python tools/bench_ops.py --only afm 2>&1 | grep '"op"' | cut -c1-140
