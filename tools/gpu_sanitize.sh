#!/bin/bash
mkdir -p gpurun_out
SEL='test_rmsnorm or test_layernorm or test_rope or test_token_plan or test_preference or test_clip_front or (test_gemm_plain and (128-256-64 or 300-256-128 or 1000-1024-640)) or test_gemm_epilogues or (test_attention and (130 or 64-2-64))'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_rmsnorm or test_token_plan or (test_gemm_plain and 128-256-64) or (test_attention and 130)" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitizer_synccheck.log | tail -5
