#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ffm_blocks.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_blocks.log 2>&1
echo "blocks tests rc=$? $(tail -1 gpurun_out/r2_tests_blocks.log)"
timeout 200 python tools/bench_ffm_blocks_local.py --world 8 2>&1 | tail -2
timeout 200 python tools/bench_ffm_blocks_local.py --world 2 2>&1 | tail -1
