#!/bin/bash
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo; grep -c avx512f /proc/cpuinfo
timeout 600 python -m pytest tests/test_gpu_session.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_sess.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_sess.log)"
timeout 600 python bench.py --no-configs > gpurun_out/r2_bench_e2e.json 2> gpurun_out/r2_bench_e2e.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_e2e.json'))
print('value', d['value'])
print(json.dumps(d['e2e'], indent=1))
PY
