#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full" > gpurun_out/t_engine.log 2>&1; echo "engine exit $?"
grep -E "engine-vs|agreement|passed|failed|quirk" gpurun_out/t_engine.log | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['achieved'], d['step_roofline'])
except Exception as e: print("parse fail", e)
PY
tail -3 gpurun_out/bench.err
K='regex:^(gemm_|attn_|rmsnorm|layernorm|clip_|token_plan|rope_su|hd_gather|embed_scatter|skipca|preference)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1119 -c 1119 --csv --log-file gpurun_out/launches.csv python bench.py --profile-run > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
