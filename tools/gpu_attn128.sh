#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llava_gpu.py -q -k "attention" > gpurun_out/t_attn128.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t_attn128.log
timeout 300 python tools/attn128_bench.py 2>&1 | tee gpurun_out/attn128_bench.log
