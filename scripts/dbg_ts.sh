export RLC_UMMA_MODE=ts RLC_UMMA_PROF=1 ONLY=fp16
for d in 0 1 2 3; do echo "== dbg $d"; RLC_UMMA_DBG=$d timeout 120 python scripts/perf_eval.py 2>&1 | grep -m1 "umma prof" | sed 's/.*| ep1 wait/ep1 wait/'; done
