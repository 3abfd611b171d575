#!/bin/bash
# GPU box: DRAM bytes per launch of the CTA-pair GEMM under different L2 eviction hints (LR_GEMM_L2_HINTS:
# A | W << 2 | C << 4, 0 normal / 1 evict_first / 2 evict_last), one launch per decoder shape under ncu.
mkdir -p gpurun_out
out=gpurun_out/r02_gemm_l2_hints.txt; : > $out
for h in 0 16 18 22 17; do
  echo "== LR_GEMM_L2_HINTS=$h" >> $out
  for shape in "phi gate_up" "phi down" "phi qkv" "phi o"; do
    LR_GEMM_L2_HINTS=$h timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:gemm_pair_kernel --csv python tools/gemm_raster_bench.py --shape "$shape" --once 2>/dev/null \
      | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d[r[-3]]=(r[-1],r[-2])
print('$shape', {k:(v[0]+' '+v[1]) for k,v in d.items()})
" >> $out
  done
done
cat $out
