#!/bin/bash
# round 2, GPU call 3: correctness of the tcgen05 DeepFM kernel, then timings
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -k "packed" > gpurun_out/r2_tests_packed.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_packed.log)"
timeout 150 python tools/r2_deepfm_time.py > gpurun_out/r2_deepfm_time.jsonl 2> gpurun_out/r2_deepfm_time.err
echo "time rc=$?"
tail -30 gpurun_out/r2_deepfm_time.jsonl
tail -5 gpurun_out/r2_deepfm_time.err
