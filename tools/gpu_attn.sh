#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -x > gpurun_out/t_attn_tc.log 2>&1; echo "attn tests exit $?"; tail -4 gpurun_out/t_attn_tc.log
timeout 300 python tools/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; echo "diag exit $?"; tail -6 gpurun_out/diag_attn.log
timeout 300 python tools/gpu_diag.py acc 2>&1 | grep attention
timeout 200 python tools/attn_trace.py 2>&1 | grep -E "==|period|segments"
