#!/bin/bash
# compute-sanitizer over the kernels added in round 1g (small shapes): memcheck, then racecheck on the shared-memory
# ownership of the two sample-major backward kernels.
mkdir -p gpurun_out
SEL='test_bilinear_backward_kernel and (37-39-16 or 65-5-8) or test_afm_backward_kernel and (37-39-16-16 or 40-5-32-8) or test_sparse_embedding_gradient_coalesced_by_segments and 64-39-16'
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_training.py -q -p no:cacheprovider -k "$SEL" > gpurun_out/r1g_memcheck_training.log 2>&1
echo "memcheck training rc=$?" | tee gpurun_out/r1g_sanitizer_steps.log
SEL2='test_bilinear_written_in_place and 33-5-8 or test_opn_shapes_and_ragged_batches and mat and 7-8-256 or test_mlp_with_a_tall_first_layer and 4104'
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_more.py -q -p no:cacheprovider -k "$SEL2" > gpurun_out/r1g_memcheck_more.log 2>&1
echo "memcheck more rc=$?" | tee -a gpurun_out/r1g_sanitizer_steps.log
SEL3='test_bilinear_backward_kernel and 65-5-8 or test_afm_backward_kernel and 40-5-32-8'
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_training.py -q -p no:cacheprovider -k "$SEL3" > gpurun_out/r1g_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r1g_sanitizer_steps.log
for f in gpurun_out/r1g_memcheck_training.log gpurun_out/r1g_memcheck_more.log gpurun_out/r1g_racecheck.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $f | tail -3; done
