#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu exit $?"
grep -E "engine-vs|agreement|passed|failed|quirk|layer |worst|FAILED|Error" gpurun_out/t_all_gpu.log | tail -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
