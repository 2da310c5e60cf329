#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py tests/test_gpu_modules.py -q -p no:cacheprovider > gpurun_out/r2_tests_mod.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_mod.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_mod.log | head -20
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['ms_per_step_min'], d['ms_per_step_max'], 'frac', d['roofline']['frac'], 'of_ceiling', d['roofline']['ceiling']['of_ceiling'])
print('graph', d['graph_replay'])
print('module', d['module_api'])
print('layout_c', d['layout_c'])
print('e2e', d['e2e'])
print('cpu', d['cpu_baseline'])
for k,v in (d['configs'] or {}).items(): print(k, {kk:vv for kk,vv in v.items() if kk!='workload'})
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
