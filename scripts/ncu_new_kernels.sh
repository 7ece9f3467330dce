#!/bin/bash
# ncu --set full of the round-2 tensor-core kernels outside K1: the training GEMM (k_gemm_tc, cfg4 sizes) and the T-mid
# stack tiles (k_tmid_rows_tc, 4096 x 1024 rows).  Usage on the GPU box: bash scripts/ncu_new_kernels.sh TAG
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python scripts/prof_rows_gemm.py > $OUT/plain_gemm_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -c 12 -f -o $OUT/gemm_tc_$TAG python scripts/prof_rows_gemm.py > $OUT/ncu_gemm_$TAG.log 2>&1
echo "ncu gemm rc=$?"
timeout 200 python scripts/time_tmid.py > $OUT/plain_tmid_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tmid_rows_tc -c 2 -f -o $OUT/tmid_tc_$TAG python scripts/time_tmid.py > $OUT/ncu_tmid_$TAG.log 2>&1
echo "ncu tmid rc=$?"
