#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ffm_blocks.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_blocks.log 2>&1
echo "blocks tests rc=$? $(tail -1 gpurun_out/r2_tests_blocks.log)"
grep -E "^FAILED|^ERROR|Error" gpurun_out/r2_tests_blocks.log | head
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_multi.log 2>&1
echo "multi tests rc=$? $(tail -1 gpurun_out/r2_tests_multi.log)"
grep -E "^FAILED|^ERROR|Error" gpurun_out/r2_tests_multi.log | head -10
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --only-sharded > gpurun_out/r2_sharded_n2.json 2> gpurun_out/r2_sharded_n2.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_sharded_n2.json') if l.startswith('{')][-1])
for k,v in d['sharded'].items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','value','nvlink_gbs_per_gpu','error')})
PY
