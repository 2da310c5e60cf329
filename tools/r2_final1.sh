#!/bin/bash
# round-2 closing call on one GPU: whole GPU suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/r2_tests_all.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_all.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_all.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$? lines=$(wc -l < gpurun_out/r2_bench_n1.json)"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'of_ceiling', d['roofline']['ceiling']['of_ceiling'])
print('module', d['module_api']['value'], 'layout_c', d['layout_c']['value'], d['layout_c']['frac'])
print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for k,v in (d['configs'] or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), v.get('error'))
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY

