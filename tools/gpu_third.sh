#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" > gpurun_out/t_kernels_gemm.log 2>&1; echo "kernels_gemm exit $?"; tail -3 gpurun_out/t_kernels_gemm.log
timeout 600 python tools/gpu_diag.py perf > gpurun_out/diag_perf.log 2>&1; echo "diag_perf exit $?"; head -9 gpurun_out/diag_perf.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full" > gpurun_out/t_engine.log 2>&1; echo "engine exit $?"
grep -E "engine-vs|agreement|passed|failed|quirk" gpurun_out/t_engine.log | tail -20
K='regex:^(gemm_|attn_|rmsnorm|layernorm|clip_|token_plan|rope_su|hd_gather|embed_scatter|skipca|preference)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1119 -c 1119 --csv --log-file gpurun_out/launches.csv python bench.py --profile-run > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
