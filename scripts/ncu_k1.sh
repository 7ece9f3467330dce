#!/bin/bash
# ncu evidence for the headline kernel (one gpurun call, one GPU): launch list of the bench command + one full capture of K1.
# Usage: bash scripts/ncu_k1.sh [tag]      (each ncu pass only after the same command exited 0 without ncu)
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity --sustain-s 0"
timeout 300 $CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_critic_umma_grid3 -s 4 -c 2 -o $OUT/k1_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $OUT/k1_$TAG.ncu-rep --page raw --csv > $OUT/k1_${TAG}_raw.csv 2>/dev/null
ls -la $OUT | grep $TAG
