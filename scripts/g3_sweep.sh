#!/bin/bash
for res in 1 0; do
for kc in 96 80 64; do
  echo "== RESIDENT=$res KC=$kc"
  RLC_G3_RESIDENT=$res RLC_G3_KC=$kc ONLY=fp16x3 timeout 120 python scripts/perf_eval.py 2>&1 | tail -1
done
done
for kc in 96 64; do
echo "== MICRO 2/4/8 KC=$kc"
for m in 2 4 8; do RLC_G3_KC=$kc RLC_UMMA_MICRO=$m ONLY=fp16x3 timeout 120 python scripts/perf_eval.py 2>&1 | tail -1; done
done
