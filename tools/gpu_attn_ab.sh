#!/bin/bash
# GPU box: correctness of every attention variant first (a deadlocked variant is killed by `timeout` and dropped, the
# check resumes after it), then the interleaved A/B timing of the variants that passed.
# usage: bash tools/gpu_attn_ab.sh "v1,v2,..." [rounds]
mkdir -p gpurun_out
remaining="$1"; rounds="${2:-7}"; good=""
log=gpurun_out/r02_attn_ab_check.log; : > $log
while [ -n "$remaining" ]; do
  timeout 150 python tools/attn_ab.py --only="$remaining" --check-only > gpurun_out/_chk.txt 2>&1
  cat gpurun_out/_chk.txt >> $log
  next=""; hit=0
  for v in ${remaining//,/ }; do
    if grep -q "^CHECK_OK $v\$" gpurun_out/_chk.txt; then good="$good,$v";
    elif grep -q "^CHECK_BAD $v\$" gpurun_out/_chk.txt; then echo "BAD $v" >> $log;
    elif [ $hit -eq 0 ]; then hit=1; echo "HUNG_OR_CRASHED $v" >> $log;
    else next="$next,$v"; fi
  done
  remaining="${next#,}"
done
good="${good#,}"
echo "variants that passed: $good" | tee -a $log
timeout 400 python tools/attn_ab.py --only="$good" --rounds $rounds > gpurun_out/r02_attn_ab.txt 2>&1
tail -60 gpurun_out/r02_attn_ab.txt
