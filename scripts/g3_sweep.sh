#!/bin/bash
# diagnostic decomposition (RLC_UMMA_MICRO bit flags: full barrier protocol, one role's WORK removed) of the split kernels
P=${1:-fp16c8}
for m in 0 1 14 2 4 8 6 10 12; do
  echo "== $P MICRO=$m"
  RLC_UMMA_MICRO=$m ONLY=$P timeout 120 python scripts/perf_eval.py 2>&1 | tail -1
done
ONLY=$P RLC_UMMA_PROF=1 RLC_UMMA_TRACE=1 timeout 120 python scripts/perf_eval.py 2>&1 | tail -10 | head -6
