#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_qwen_gpu.py tests/test_llava_gpu.py -q -s -k "shortcut or golden or stages" > gpurun_out/t_short.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error|shortcut|last-layer" gpurun_out/t_short.log | tail -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline']['frac_of_sustained'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print("parse fail", e)
PY
tail -3 gpurun_out/bench.err
