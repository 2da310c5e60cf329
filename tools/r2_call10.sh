#!/bin/bash
# 2-GPU call: NVLink peer-read probe + the multi-GPU tests (row-sharded DeepFM, table-sharded FFM)
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 ./build/peer_probe > gpurun_out/r2_peer_probe.log 2>&1
echo "probe rc=$?"
cat gpurun_out/r2_peer_probe.log
timeout 900 python -m pytest tests/test_multi_gpu.py -q -p no:cacheprovider > gpurun_out/r2_tests_multi.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_multi.log)"
grep -E "^FAILED|^ERROR|Error|error" gpurun_out/r2_tests_multi.log | head -20
