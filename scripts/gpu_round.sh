#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + one full capture of K1.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 6000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
if [ "${SECONDARY:-0}" = "1" ]; then
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1
timeout 300 python scripts/bench_updates.py 2>/dev/null | grep "^{" > $OUT/updates_$TAG.jsonl
timeout 300 python scripts/bench_configs.py 2>/dev/null | grep "^{" > $OUT/configs_$TAG.jsonl
timeout 300 python scripts/bench_hbm_kernels.py 2>/dev/null | grep "^{" > $OUT/hbm_kernels_$TAG.jsonl
timeout 300 python scripts/bench_device_loop.py 5000 8 2>/dev/null | grep "^{" > $OUT/device_loop_$TAG.jsonl
fi
if [ "${NCU:-0}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity --sustain-s 0"
  timeout 300 $CMD > $OUT/plain_$TAG.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
  echo "ncu launches rc=$?"
  timeout 300 $CMD > $OUT/plain2_$TAG.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_critic_umma_grid3 -s 4 -c 2 -o $OUT/k1_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full rc=$?"
fi
ls -la $OUT | tail -8
