#!/bin/bash
# K1-split experiments of one gpurun call: knob settings and RLC_UMMA_MICRO decompositions of the fp16c8 kernel.
# Usage: bash scripts/exp_k1.sh <tag> "<VAR=val ...>;<VAR=val ...>;..."   (each ;-separated group is one timed run)
TAG=${1:-exp}
OUT=gpurun_out/exp_$TAG.log
mkdir -p gpurun_out
: > $OUT
IFS=';' read -ra RUNS <<< "$2"
for r in "${RUNS[@]}"; do
  echo "== $r" >> $OUT
  env $r ONLY=${ONLY:-fp16c8} timeout 120 python scripts/perf_eval.py 2>&1 | grep -v "^\[grid3 trace\]\|^TRACE" | tail -${TAILN:-2} >> $OUT
done
if [ -n "$TRACE" ]; then
  env $TRACE ONLY=${ONLY:-fp16c8} RLC_UMMA_PROF=1 RLC_UMMA_TRACE=2 timeout 120 python scripts/perf_eval.py > gpurun_out/trace_$TAG.log 2>&1
  grep "grid3 prof" gpurun_out/trace_$TAG.log | tail -2 >> $OUT
fi
cat $OUT
