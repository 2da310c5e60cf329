#!/bin/bash
# 2-GPU call: block-exchange FFM tests (virtual ranks on one GPU + torchrun x2) and the 2-GPU bench line
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ffm_blocks.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_blocks.log 2>&1
echo "blocks tests rc=$? $(tail -1 gpurun_out/r2_tests_blocks.log)"
grep -E "^FAILED|^ERROR|Error" gpurun_out/r2_tests_blocks.log | head -20
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_multi.log 2>&1
echo "multi tests rc=$? $(tail -1 gpurun_out/r2_tests_multi.log)"
grep -E "^FAILED|^ERROR|Error" gpurun_out/r2_tests_multi.log | head -20
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2_bench_n2.err | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('value', d['value'], 'ms', d['ms_per_step'])
print(json.dumps(d['sharded'], indent=1))
PY
