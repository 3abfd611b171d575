#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > gpurun_out/t_attn_tc.log 2>&1; echo "attn tests exit $?"; tail -4 gpurun_out/t_attn_tc.log
timeout 300 python tools/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; echo "diag exit $?"; grep attention gpurun_out/diag_attn.log
