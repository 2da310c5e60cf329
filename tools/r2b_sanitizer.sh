#!/bin/bash
# compute-sanitizer memcheck over the kernels of the closing session (small shapes): gathering dense layer + fused logit
# Linear (split and packed tables), plain dense chain, CIN layers with the rebuilt epilogue, xDeepFM stage 1 on deepfm_fast
mkdir -p gpurun_out
run() {  # name, file, -k selection
  timeout 110 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest $2 -q -p no:cacheprovider -x -k "$3" > gpurun_out/r2b_memcheck_$1.log 2>&1
  echo "memcheck $1 rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r2b_memcheck_$1.log | tail -1) $(grep -E 'passed|failed' gpurun_out/r2b_memcheck_$1.log | tail -1)" | tee -a gpurun_out/r2b_sanitizer_steps.log
}
: > gpurun_out/r2b_sanitizer_steps.log
run wide tests/test_gpu_ops.py "test_deepfm_wide_mlp and (12-32-deep3 or 39-16-deep6) or test_deepfm_wide_mlp_on_packed_table and 12-deep1"
run chain tests/test_gpu_ops.py "test_mlp_tensor_core_chain and dims1"
run cin tests/test_gpu_ops.py "test_cin_tensor_core_wide_layers and (7-16 or 9-32)"
run xdeepfm tests/test_gpu_ops.py "test_fused_model_parity and idx_dtype0-16-6-64-xdeepfm"
