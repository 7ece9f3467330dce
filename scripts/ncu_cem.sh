#!/bin/bash
# ncu --set full capture of one cfg3 predict_action (B=256, N=1024, 3 CEM iterations, 2 components, 400-300 T-mid):
# the state-term kernel (k_mlp2_rows, T-mid mode) and the CEM kernel.  Usage (GPU box, repo root): bash scripts/ncu_cem.sh TAG
# summarise here with: python scripts/summarise_sb_ncu.py TAG cem
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CEM_ONCE=1 timeout 200 python scripts/time_cem_parts.py > $OUT/cem_plain_$TAG.log 2>&1 &&
CEM_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cem|k_mlp2_rows' -c 4 \
  -o $OUT/cem_$TAG -f python scripts/time_cem_parts.py > $OUT/cem_ncu_$TAG.log 2>&1
echo "ncu cem rc=$?"
tail -2 $OUT/cem_ncu_$TAG.log
