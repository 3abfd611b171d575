#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_llava_gpu.py -q -s > gpurun_out/t_llava.log 2>&1; echo "pytest llava exit $?"
grep -E "passed|failed|FAILED|Error|error|engine|ratio|alone|prob|hidden|projector|inputs_embeds|last_hidden" gpurun_out/t_llava.log | tail -60
timeout 900 python tools/bench_llava.py --steps 3 --warmup 3 > gpurun_out/bench_llava.log 2> gpurun_out/bench_llava.err; echo "bench llava exit $?"
tail -3 gpurun_out/bench_llava.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_llava.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline'])
except Exception as e: print("parse fail", e)
PY
