#!/bin/bash
# first contact with the GPU: every step in its own process under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gpu_diag.py gemm > gpurun_out/diag_gemm.log 2>&1; echo "diag_gemm exit $?"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not gemm" > gpurun_out/t_kernels_other.log 2>&1; echo "kernels_other exit $?"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" > gpurun_out/t_kernels_gemm.log 2>&1; echo "kernels_gemm exit $?"
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full" > gpurun_out/t_engine.log 2>&1; echo "engine exit $?"
timeout 600 python tools/gpu_diag.py perf > gpurun_out/diag_perf.log 2>&1; echo "diag_perf exit $?"
tail -5 gpurun_out/diag_gemm.log gpurun_out/t_kernels_other.log gpurun_out/t_kernels_gemm.log gpurun_out/t_engine.log
tail -20 gpurun_out/diag_perf.log
