#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -k "packed" > gpurun_out/r2_tests_packed.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_packed.log)"
timeout 200 python tools/r2_deepfm_time.py --extra --reps 5 > gpurun_out/r2_deepfm_time2.jsonl 2> gpurun_out/r2_deepfm_time2.err
echo "time rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_deepfm_time2.jsonl'):
    d=json.loads(l)
    print(d['config'][:42], d['kernel'], d['variant'], 'pdl' if d['pdl_overlap'] else 'nopdl', d['us_per_launch_median'])
PY
timeout 120 python tools/r2_tc5_trace.py 1 2560 > gpurun_out/r2_trace_v1_l2.txt 2>&1; echo "trace rc=$?"
