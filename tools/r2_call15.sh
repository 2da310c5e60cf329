#!/bin/bash
# elect.sync issue of tcgen05.mma: parity + timing of every tcgen05 kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_more.py tests/test_gpu_fullsize.py -x -q -p no:cacheprovider -k "cin or xdeepfm or cross or deepfm or mlp or dense or dnn" > gpurun_out/r2_tests_elect.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_elect.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_tests_elect.log | head
python tools/cin_profile_driver.py 65536
python tools/bench_ops.py --only cross_layer,deepfm_generic_mlp400,xdeepfm,cin_layer 2>&1 | grep -v Warn | tail -12
python tools/r2_deepfm_time.py 2>&1 | tail -6
