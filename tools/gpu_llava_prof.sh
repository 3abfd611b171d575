#!/bin/bash
mkdir -p gpurun_out
K='regex:^(gemm_|attn_|rmsnorm|layernorm|clip_|token_plan|rope_su|anyres|skipca|preference)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 522 -c 522 --csv --log-file gpurun_out/launches_llava.csv python tools/bench_llava.py --profile-run > gpurun_out/ncu_launch_llava.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 30 -c 1 -o gpurun_out/prof_attn_hd128 -f python tools/bench_llava.py --profile-run > gpurun_out/ncu_attn128.log 2>&1; echo "ncu attn exit $?"
