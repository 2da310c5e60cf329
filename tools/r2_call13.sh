#!/bin/bash
# 8-GPU call: the exchange paths at full size (configs[4] sharded over 8 GPUs, DeepFM row-sharded)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --only-sharded > gpurun_out/r2_sharded_n8.json 2> gpurun_out/r2_sharded_n8.err
echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2_sharded_n8.err | tail -8
cat gpurun_out/r2_sharded_n8.json
