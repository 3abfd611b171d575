#!/bin/bash
# compute-sanitizer over what was added at the end of r01: multi-tile attention CTAs, packed-row decoder, the all-rows head
# kernels (lr_softmax_rows_bf16, lr_masked_mean_rows_bf16), zero-row gather. usage (GPU box): bash tools/gpu_sanitize3.sh
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -q -m gpu \
    -k "multitile or (test_attention and 577)" > gpurun_out/sanitizer_${tool}_multitile.log 2>&1; echo "$tool multitile exit $?"
  grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_${tool}_multitile.log | tail -4
done
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_engine_gpu.py -q -m gpu \
  -k "packed_valid or attribute_variants or softmax_rows or gather_rows_negative" > gpurun_out/sanitizer_memcheck_packed_attrs.log 2>&1
echo "memcheck packed/attrs exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_memcheck_packed_attrs.log | tail -4
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_engine_gpu.py -q -m gpu \
  -k "softmax_rows or gather_rows_negative" > gpurun_out/sanitizer_racecheck_head_rows.log 2>&1; echo "racecheck head rows exit $?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck_head_rows.log | tail -4
