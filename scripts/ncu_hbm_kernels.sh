#!/bin/bash
# ncu --set full capture of the memory-bound kernels (K2 / K3 / K6) at the sizes scripts/bench_hbm_kernels.py times them on.
# Usage (GPU box, repo root): bash scripts/ncu_hbm_kernels.sh TAG ; summarise here with scripts/summarise_sb_ncu.py TAG hbm
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
HBM_NCU=1 timeout 300 python scripts/bench_hbm_kernels.py > $OUT/hbm_plain_$TAG.log 2>&1 &&
HBM_NCU=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'^k_|^void k_' -c 60 -o $OUT/hbm_$TAG -f python scripts/bench_hbm_kernels.py > $OUT/hbm_ncu_$TAG.log 2>&1
echo "ncu hbm rc=$?"
tail -3 $OUT/hbm_ncu_$TAG.log
