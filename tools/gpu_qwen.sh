#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_qwen_gpu.py -q -s -x > gpurun_out/t_qwen.log 2>&1; echo "pytest qwen exit $?"
grep -E "passed|failed|FAILED|Error|error|engine|ratio|alone|prob|hidden|vit_|image_embeds|inputs_embeds|last_hidden|assert" gpurun_out/t_qwen.log | tail -70
timeout 900 python tools/bench_qwen.py --steps 3 --warmup 3 > gpurun_out/bench_qwen.log 2> gpurun_out/bench_qwen.err; echo "bench qwen exit $?"
tail -5 gpurun_out/bench_qwen.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_qwen.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d.get('e2e_uint8',{}).get('value'), 'gate_up', d['roofline']['achieved'], d['step_roofline'])
except Exception as e: print("parse fail", e)
PY
