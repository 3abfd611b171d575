#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" > gpurun_out/t_kernels_gemm.log 2>&1; echo "kernels_gemm exit $?"; tail -3 gpurun_out/t_kernels_gemm.log
timeout 600 python tools/gpu_diag.py perf > gpurun_out/diag_perf.log 2>&1; echo "diag_perf exit $?"; cat gpurun_out/diag_perf.log | tail -14
timeout 1500 python -m pytest tests/test_engine_gpu.py -q -m gpu -s > gpurun_out/t_engine.log 2>&1; echo "engine exit $?"
grep -E "engine-vs|agreement|passed|failed" gpurun_out/t_engine.log | tail -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1125 -c 1125 --csv --log-file gpurun_out/launches.csv python bench.py --profile-run > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 95 -c 8 -o gpurun_out/prof_gemm_dec -f python bench.py --profile-run > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 22 -c 2 -o gpurun_out/prof_attn -f python bench.py --profile-run > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out
