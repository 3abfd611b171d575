#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu exit $?"
grep -E "engine-vs|agreement|passed|failed|quirk|layer |worst|FAILED|Error|padding|17 crops" gpurun_out/t_all_gpu.log | tail -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline']['frac_of_sustained'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print("parse fail", e)
PY
tail -3 gpurun_out/bench.err
NL=$(python -c "import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print(d['gpu_launches']//d['steps'])" 2>/dev/null || echo 1059)
echo "launches per step: $NL"
K='regex:^(gemm_|attn_|rmsnorm|layernorm|clip_|token_plan|rope_su|hd_gather|embed_scatter|skipca|preference|gather_rows)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s $NL -c $NL --csv --log-file gpurun_out/launches.csv python bench.py --profile-run > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python tools/launch_summary.py gpurun_out/launches.csv gpurun_out/launches.md | head -30
