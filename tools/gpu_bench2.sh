#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline']['frac_of_sustained'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print("parse fail", e)
PY
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref arm exit $?"; tail -c 600 gpurun_out/bench_ref.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 100 -c 1 -o gpurun_out/prof_gate_up -f python bench.py --profile-run > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gate_up exit $?"
