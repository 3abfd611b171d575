#!/bin/bash
mkdir -p gpurun_out
K='regex:^(gemm_|attn_|rmsnorm|layernorm|clip_|token_plan|rope_su|anyres|skipca|preference|patch_rows|mrope|compact_rows|embed_scatter)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 544 -c 544 --csv --log-file gpurun_out/launches_qwen.csv python tools/bench_qwen.py --profile-run > gpurun_out/ncu_launch_qwen.log 2>&1; echo "ncu launches exit $?"
tail -2 gpurun_out/ncu_launch_qwen.log
python tools/launch_summary.py gpurun_out/launches_qwen.csv gpurun_out/launches_qwen.md | head -40
