#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -p no:cacheprovider -m gpu -k "ffm or interleaved or field_aware" > gpurun_out/r2_tests_ffm.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_ffm.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_ffm.log | head
python tools/bench_ffm_cols.py --cols 4
python tools/bench_ffm_cols.py --cols 8 --batch 262144
python tools/bench_ffm_cols.py --cols 16 --batch 32768
