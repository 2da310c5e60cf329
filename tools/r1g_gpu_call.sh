#!/bin/bash
# Round-1g GPU call: new bilinear backward tests first, then the whole GPU suite, the bench line, per-op numbers and one
# ncu capture of the DCN kernel.  Every step has its own timeout and writes under gpurun_out/ as it goes.
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r1g_steps.log; }
stamp start
timeout 300 python -m pytest tests/test_gpu_training.py -q -p no:cacheprovider > gpurun_out/r1g_tests_training.log 2>&1
stamp "training tests rc=$? $(tail -1 gpurun_out/r1g_tests_training.log)"
timeout 120 python tools/bench_ops.py --only bilinear_backward --rows-per-field 65536 > gpurun_out/r1g_ops_bilinear_backward.jsonl 2> gpurun_out/r1g_ops_bilinear_backward.err
stamp "bench_ops bilinear_backward rc=$?"
timeout 480 python -m pytest tests -m gpu -q -p no:cacheprovider --ignore=tests/test_gpu_training.py --durations=12 > gpurun_out/r1g_tests_rest.log 2>&1
stamp "other GPU tests rc=$? $(tail -1 gpurun_out/r1g_tests_rest.log)"
timeout 200 python bench.py > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench_n1.err
stamp "bench rc=$?"
timeout 120 python tools/bench_ops.py --only dcn > gpurun_out/r1g_ops_dcn.jsonl 2> gpurun_out/r1g_ops_dcn.err
stamp "bench_ops dcn rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dcn_tc -c 1 -f -o gpurun_out/r1g_dcn_tc python tools/bench_ops.py --only dcn > gpurun_out/r1g_ncu_dcn.log 2>&1
stamp "ncu dcn rc=$?"
timeout 60 ncu -i gpurun_out/r1g_dcn_tc.ncu-rep --page raw --csv > gpurun_out/r1g_dcn_tc_raw.csv 2>/dev/null
stamp "done"
