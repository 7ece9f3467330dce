#!/bin/bash
# ncu --set full capture of the small-minibatch launches (k_sb_forward, k_sb_update) and the loop glue kernels.
# Usage (GPU box, repo root): bash scripts/ncu_small_batch.sh TAG
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/time_small_batch.py > $OUT/sb_plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sb_|k_env_step_train|k_loop_stage|k_policy_reduce' \
  -c 14 -o $OUT/sb_$TAG -f python scripts/time_small_batch.py > $OUT/sb_ncu_$TAG.log 2>&1
echo "ncu small-batch rc=$?"
tail -3 $OUT/sb_ncu_$TAG.log
