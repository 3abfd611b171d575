#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm_plain and 128-256-64-2" > gpurun_out/t_pair0.log 2>&1; echo "pair smallest exit $?"; tail -5 gpurun_out/t_pair0.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" > gpurun_out/t_pair.log 2>&1; echo "pair all exit $?"; tail -8 gpurun_out/t_pair.log
timeout 300 python tools/gpu_diag.py perf > gpurun_out/diag_perf.log 2>&1; echo "diag exit $?"; grep gemm gpurun_out/diag_perf.log
