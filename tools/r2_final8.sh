#!/bin/bash
# the driver's scaling command at N = 8 (full bench line incl. configs and the sharded paths)
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$? lines=$(wc -l < gpurun_out/r2_bench_n$N.json)"
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.load(open(f'gpurun_out/r2_bench_n{n}.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'n_gpus', d['n_gpus'], 'e2e', d['e2e']['value'])
for k,v in (d['configs'] or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
for k,v in (d['sharded'] or {}).items(): print('sharded', k, v.get('value'), v.get('ms_per_step'), v.get('nvlink_gbs_per_gpu'), v.get('error'))
print('clocks', d['clocks'])
PY
grep -i "error\|Traceback" gpurun_out/r2_bench_n$N.err | head -5
