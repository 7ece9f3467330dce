#!/bin/bash
P=${1:-fp16c8}
for res in 1 0; do for kc in 96 64; do
  echo "== $P RESIDENT=$res KC=$kc"
  RLC_G3_KC=$kc RLC_G3_RESIDENT=$res ONLY=$P RLC_UMMA_PROF=1 timeout 120 python scripts/perf_eval.py 2>&1 | tail -2 | head -1 | cut -c1-120
  RLC_G3_KC=$kc RLC_G3_RESIDENT=$res ONLY=$P timeout 120 python scripts/perf_eval.py 2>&1 | tail -1
done; done
