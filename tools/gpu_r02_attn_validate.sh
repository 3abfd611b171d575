#!/bin/bash
# GPU box: validation of the attention kernel's new product defaults (two-half P.V issue, chunk-level masking, packed
# FFMA2, converged MMA warp for two-tile CTAs): full GPU suite, compute-sanitizer memcheck + synccheck on the attention
# tests of every head_dim / layout, main bench, LLaVA bench.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gpu_tests_f.log 2>&1
grep -v "Warning\|warnings.warn" gpurun_out/r02_gpu_tests_f.log | tail -5
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py tests/test_llava_gpu.py tests/test_qwen_gpu.py -q -m gpu \
    -k "attention or attn" > gpurun_out/r02_sanitizer_${tool}_attention_b.log 2>&1
  echo "$tool attention exit $?"
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_${tool}_attention_b.log | tail -3
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8']['value'], 'gate_up', d['roofline']['achieved'], d['step_roofline']['frac_of_sustained'], 'gpu_ref', d.get('gpu_reference',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
timeout 600 python tools/bench_llava.py > gpurun_out/r02_bench_llava.json 2> gpurun_out/r02_bench_llava.err; echo "llava bench exit $?"; tail -c 1200 gpurun_out/r02_bench_llava.json
