#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (small shapes): block exchange (virtual ranks), row-sharded
# DeepFM, four-producer interleaved FFM, DCN on tcgen05, folded CIN layer 0, tcgen05 DeepFM
mkdir -p gpurun_out
run() {  # name, file, -k selection
  timeout 240 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest $2 -q -p no:cacheprovider -x -k "$3" > gpurun_out/r2_memcheck_$1.log 2>&1
  echo "memcheck $1 rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r2_memcheck_$1.log | tail -1) $(grep -E 'passed|failed' gpurun_out/r2_memcheck_$1.log | tail -1)" | tee -a gpurun_out/r2_sanitizer_steps.log
}
: > gpurun_out/r2_sanitizer_steps.log
run blocks tests/test_ffm_blocks.py "test_virtual_ranks_match_oracle and (13-8-3-517 or 39-16-8-1000) and idx_dtype0 or test_embed_sharded_ffm_virtual_ranks and 16-8 or test_row_sharded_deepfm_virtual_ranks_bit_exact and 3"
run dcn5 tests/test_gpu_ops.py "test_dcn_tcgen05_path and (50-32-2 or 128-16-1)"
run cin tests/test_gpu_ops.py "test_cin_tensor_core_wide_layers and 7-16"
