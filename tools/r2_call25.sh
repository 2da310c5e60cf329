#!/bin/bash
# fused wide deep branch (gather in layer 1, logit Linear in the last hidden layer's epilogue) + channel-block sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -p no:cacheprovider -k "mlp_tensor_core_chain or deepfm_wide_mlp or cin or xdeepfm or mlp" > gpurun_out/r2_tests_f1.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_f1.log)"
for blk in ${BLOCKS:-0}; do
  echo "== TRS_DENSE_BLOCK=$blk"
  TRS_DENSE_BLOCK=$blk timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*'
done
echo "== TRS_MLP_NO_GATHER=1"
TRS_MLP_NO_GATHER=1 timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*'
