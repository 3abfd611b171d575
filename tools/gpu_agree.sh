#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tests/decision_agreement_study.py 32 > gpurun_out/decision_agreement.txt 2>&1; echo "agreement exit $?"; tail -16 gpurun_out/decision_agreement.txt
timeout 300 ncu --set full --clock-control none -k regex:"rmsnorm|layernorm1024|rope_su" -c 3 -o gpurun_out/prof_norms -f python tools/norm_only.py > gpurun_out/ncu_norms.log 2>&1; echo "ncu norms exit $?"
