#!/bin/bash
# compute-sanitizer on the Qwen2.5-VL branch kernels + engine, 2 ncu full captures
mkdir -p gpurun_out
SEL='test_attention_packed_hd96_noncausal or test_attention_segments_hd96 or (test_attention_gqa_hd128 and (4-2-130 or 8-1-257)) or test_gemm_bias_swiglu or test_gemm_bias_rope or test_patch_rows or test_mrope_plan or test_compact_rows or test_qwen_gpu_preprocess or (test_qwen_vs_reference_golden and slim) or test_qwen_last_layer or test_qwen_skipca_without'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_qwen_gpu.py -q -m gpu -k "$SEL" > gpurun_out/sanitizer_memcheck_qwen.log 2>&1; echo "memcheck qwen exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_memcheck_qwen.log | tail -8
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_qwen_gpu.py -q -m gpu -k "(test_attention_segments_hd96 and not 40) or (test_attention_gqa_hd128 and 4-2-130) or test_mrope_plan or test_gemm_bias_swiglu" > gpurun_out/sanitizer_synccheck_qwen.log 2>&1; echo "synccheck qwen exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitizer_synccheck_qwen.log | tail -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_qwen_gpu.py -q -m gpu -k "test_mrope_plan or test_patch_rows or test_compact_rows" > gpurun_out/sanitizer_racecheck_qwen.log 2>&1; echo "racecheck qwen exit $?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitizer_racecheck_qwen.log | tail -5
# ncu full: segment window attention (first window layer = 1st attn launch of the timed step) and the vision qkv+rope GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 33 -c 1 -o gpurun_out/prof_qwen_win_attn -f python tools/bench_qwen.py --profile-run > gpurun_out/ncu_qwen_attn.log 2>&1; echo "ncu win attn exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 200 -c 4 -o gpurun_out/prof_qwen_vit_gemms -f python tools/bench_qwen.py --profile-run > gpurun_out/ncu_qwen_gemm.log 2>&1; echo "ncu vit gemms exit $?"
ls -la gpurun_out/*.ncu-rep
