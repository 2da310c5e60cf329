#!/bin/bash
# ncu --set full of the round-2 kernels not captured yet (one GPU): interleaved FFM (single-GPU configs[4] and a 4-column
# shard), the block-exchange kernel with all shards local, and the bench command's launch list with the final build
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffm_interleaved_kernel -s 4 -c 1 -f -o gpurun_out/r2_ffm_inter16 python tools/bench_ffm_cols.py --cols 16 --batch 32768 > gpurun_out/r2_ncu_ffm16.log 2>&1
echo "ffm16 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffm_interleaved_kernel -s 4 -c 1 -f -o gpurun_out/r2_ffm_inter4 python tools/bench_ffm_cols.py --cols 4 > gpurun_out/r2_ncu_ffm4.log 2>&1
echo "ffm4 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffm_blocks_kernel -s 4 -c 1 -f -o gpurun_out/r2_ffm_blocks_local python tools/bench_ffm_blocks_local.py --world 8 > gpurun_out/r2_ncu_blocks.log 2>&1
echo "blocks rc=$?"
for n in r2_ffm_inter16 r2_ffm_inter4 r2_ffm_blocks_local; do timeout 120 ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/r2_bench_launches.csv) lines"
