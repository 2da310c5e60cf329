#!/bin/bash
# Round-1g closing GPU call: whole GPU suite, bench line, ncu launch list of the bench command, per-op table.
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r1g_final_steps.log; }
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r1g_final_tests.log 2>&1
stamp "GPU tests rc=$? $(tail -1 gpurun_out/r1g_final_tests.log)"
timeout 200 python bench.py > gpurun_out/r1g_final_bench_n1.json 2> gpurun_out/r1g_final_bench_n1.err
stamp "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r1g_final_bench_reference.json 2>/dev/null
stamp "reference arm rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_under_ncu.log 2>&1
stamp "launch list rc=$?"
timeout 400 python tools/bench_ops.py --json gpurun_out/r1g_ops_baseline_shapes.json > gpurun_out/r1g_ops.log 2>&1
stamp "bench_ops rc=$?"
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_more.py -q -p no:cacheprovider -k "test_opn_shapes_and_ragged_batches and (vec or num) and (39-16 or 7-8 or 26-32)" > gpurun_out/r1g_memcheck_opn_vec.log 2>&1
stamp "memcheck opn vec/num rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r1g_memcheck_opn_vec.log | tail -1) $(grep -E 'passed|failed' gpurun_out/r1g_memcheck_opn_vec.log | tail -1)"
