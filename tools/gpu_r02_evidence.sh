#!/bin/bash
# r02 evidence run (GPU box): full GPU test suite, launch list of one bench step, ncu --set full of every memory-bound
# kernel of the Phi-3.5-V path (SkipCA head / scores, HD gather, embed scatter, im2col, CLIP embed+LN, row gather,
# token plan, preprocessing), compute-sanitizer on the product defaults. Outputs under gpurun_out/.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -s) > gpurun_out/r02_gpu_tests_c.log 2>&1
grep -v "Warning\|warnings.warn" gpurun_out/r02_gpu_tests_c.log | tail -8

# launch list of one device-timed step (2 forwards of 32 samples + preference); -s skips the warm-up step's launches
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(gemm_|attn_|rmsnorm|layernorm|clip_|token_plan|rope_su|hd_gather|embed_scatter|skipca|preference|gather_rows|synth)' \
  --csv --log-file gpurun_out/r02_launches_step.csv python bench.py --profile-run > gpurun_out/r02_ncu_launch.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/r02_ncu_launch.log

# full sections for the memory-bound kernels (a handful of launches each)
K='regex:^(skipca_head_kernel|skipca_scores_kernel|hd_gather_kernel|embed_scatter_kernel|clip_im2col_kernel|clip_embed_ln_kernel|gather_rows_kernel|token_plan_kernel|preference_kernel)'
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" -c 24 -o gpurun_out/r02_membound python bench.py --profile-run > gpurun_out/r02_ncu_membound.log 2>&1
echo "ncu membound exit $?"
timeout 900 ncu --set full --clock-control none -k 'regex:^(resample_u8_kernel|hd_crops_kernel|hd_global_kernel)' -c 12 -o gpurun_out/r02_preprocess \
  python -m pytest tests/test_preprocess_gpu.py -q -m gpu > gpurun_out/r02_ncu_preprocess.log 2>&1
echo "ncu preprocess exit $?"

# compute-sanitizer on the product defaults: multi-tile attention, packed-row decoder, cluster head, new paths
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py tests/test_index_kernels_gpu.py -q -m gpu \
    -k "multitile or (test_attention and 577) or skipca or value_head or hd_gather or embed_scatter or synth" > gpurun_out/r02_sanitizer_${tool}_kernels.log 2>&1
  echo "$tool kernels exit $?"
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_${tool}_kernels.log | tail -3
done
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_engine_gpu.py -q -m gpu \
  -k "packed_valid or attribute_variants or softmax_rows or gather_rows_negative or vision_layer_id or (slim_vs_reference and slim_gpm)" > gpurun_out/r02_sanitizer_memcheck_engine.log 2>&1
echo "memcheck engine exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_memcheck_engine.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_index_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu \
  -k "skipca or value_head or softmax_rows or gather_rows_negative" > gpurun_out/r02_sanitizer_racecheck_head.log 2>&1
echo "racecheck head exit $?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_racecheck_head.log | tail -3
