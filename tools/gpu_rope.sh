#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "rope" -x > gpurun_out/t_rope.log 2>&1; echo "rope tests exit $?"; tail -12 gpurun_out/t_rope.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full" > gpurun_out/t_engine.log 2>&1; echo "engine exit $?"
grep -E "engine-vs|agreement|passed|failed|quirk|padding|crops" gpurun_out/t_engine.log | tail -14
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline']['frac_of_sustained'])
except Exception as e: print("parse fail", e)
PY
tail -5 gpurun_out/bench.err
