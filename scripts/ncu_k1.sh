#!/bin/bash
# one ncu --set full capture of the K1 kernel on the small driver (run plain first, as the recipe requires)
TAG=${1:-x}
CMD="python scripts/perf_eval.py"
export ONLY=fp16
timeout 200 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_critic_umma -s 2 -c 1 -o gpurun_out/k1_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/plain_$TAG.log
