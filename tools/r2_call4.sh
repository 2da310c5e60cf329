#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/r2_tc5_trace.py 0 > gpurun_out/r2_trace_v0.txt 2>&1; echo "trace0 rc=$?"
timeout 120 python tools/r2_tc5_trace.py 1 > gpurun_out/r2_trace_v1.txt 2>&1; echo "trace1 rc=$?"
timeout 400 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -k "packed" > gpurun_out/r2_tests_packed.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_packed.log)"
timeout 150 python tools/r2_deepfm_time.py > gpurun_out/r2_deepfm_time.jsonl 2> gpurun_out/r2_deepfm_time.err
echo "time rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_deepfm_time.jsonl'):
    d=json.loads(l); print(d['config'][:28], d['kernel'], d['variant'], 'pdl' if d['pdl_overlap'] else 'nopdl', d['us_per_launch_median'], d['frac_of_hbm_peak_6547'], '%.1e'%d['normwise_diff_vs_first_kernel'])
PY
head -30 gpurun_out/r2_trace_v0.txt | tail -26
