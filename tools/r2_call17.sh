#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_ffm_blocks.py tests/test_gpu_session.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_hint.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_hint.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_hint.log | head
python tools/bench_ops.py --only dcn,cross_layer,xdeepfm,deepfm_generic_mlp400,ffm_model_full 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-78s %9.1f us  %7.2f M samples/s' % (d['op'][:78], d['us'], d['samples_per_s'] / 1e6))"
python tools/r2_deepfm_time.py 2>&1 | grep '"tc5"' | grep 'variant": 1' | cut -c1-200
