#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_llava_gpu.py -q -x -k "attention_hd128 or token_plan_ex or anyres_embed or value_head or gemm_rope_hd128 or preprocess or slim_bt" > gpurun_out/san_llava_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_llava_memcheck.log | tail -3
timeout 900 $CS --tool synccheck --error-exitcode 9 python -m pytest tests/test_llava_gpu.py -q -x -k "attention_hd128 and 130 or token_plan_ex or anyres_embed" > gpurun_out/san_llava_synccheck.log 2>&1; echo "synccheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_llava_synccheck.log | tail -3
timeout 900 python tools/bench_llava.py --size 13b --steps 2 --warmup 3 > gpurun_out/bench_llava13b.log 2> gpurun_out/bench_llava13b.err; echo "bench 13b exit $?"; tail -2 gpurun_out/bench_llava13b.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_llava13b.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline'])
except Exception as e: print("parse fail", e)
PY
