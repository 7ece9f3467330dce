#!/bin/bash
# TS-variant tuning sweep: in-kernel cycle accounting + timing for several chunk configurations
export RLC_UMMA_MODE=ts
for cfg in ${CFGS:-2x96 2x80 3x64}; do
  echo "=== RLC_UMMA_TS_CH=$cfg"
  RLC_UMMA_TS_CH=$cfg RLC_UMMA_PROF=1 ONLY=fp16 timeout 120 python scripts/perf_eval.py 2>&1 | grep -m1 "umma prof"
  RLC_UMMA_TS_CH=$cfg timeout 120 python scripts/perf_eval.py 2>&1 | grep "fp16 shared\|per-state\|bf16"
done
echo "=== SS"; RLC_UMMA_MODE=ss ONLY=fp16 timeout 120 python scripts/perf_eval.py 2>&1 | grep "fp16 shared"
