#!/bin/bash
mkdir -p gpurun_out
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "slim_vs_reference_golden and slim_gpm" > gpurun_out/sanitizer_memcheck_engine.log 2>&1; echo "memcheck engine exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_memcheck_engine.log | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_rmsnorm or test_layernorm or test_token_plan or test_clip_front" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitizer_racecheck.log | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_preprocess_gpu.py -q -m gpu > gpurun_out/sanitizer_memcheck_pre.log 2>&1; echo "memcheck preprocess exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_pre.log | tail -3
