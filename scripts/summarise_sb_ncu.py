"""gpurun_out/sb_TAG.ncu-rep (scripts/ncu_small_batch.sh) -> profiles/TAG_small_batch_ncu.md: one column per captured launch
of the small-minibatch / device-loop kernels.  Usage: python scripts/summarise_sb_ncu.py TAG
With a second argument "hbm": gpurun_out/hbm_TAG.ncu-rep (scripts/ncu_hbm_kernels.sh) -> profiles/TAG_hbm_kernels_ncu.md."""
import csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
KIND = sys.argv[2] if len(sys.argv) > 2 else "sb"          # sb | hbm | cem
HBM = KIND in ("hbm", "cem")
rep = os.path.join(ROOT, "gpurun_out", f"{KIND}_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio"]
kn = hdr.index("Kernel Name")
# keep the first capture of every distinct (kernel, grid) pair
gi = hdr.index("launch__grid_size") if "launch__grid_size" in hdr else None
seen, keep = set(), []
for r in data:
    key = (r[kn].split("(")[0], r[gi] if gi is not None else "")
    if key not in seen:
        seen.add(key)
        keep.append(r)
if HBM:
    KEYS.insert(12, "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    KEYS.insert(13, "lts__t_sector_hit_rate.pct")
    KEYS.insert(14, "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio")
out = [f"# ncu --set full: small-minibatch fast path and device-loop glue kernels ({tag})\n",
       "Captured by `scripts/ncu_small_batch.sh` from `scripts/time_small_batch.py` (cfg1 shape: B=32, N=62, 200-200); "
       "one column per distinct (kernel, grid size). Durations under ncu are cold-cache and serialised: the warm device "
       "times are the ones in `" + tag + "_small_batch_times.jsonl` (CUDA events over graph replays).\n",
       "| metric | unit | " + " | ".join(f"{r[kn].split('(')[0][:28]} grid {r[gi]}" for r in keep) + " |",
       "|---|---|" + "---|" * len(keep)]
for k in KEYS:
    for i, h in enumerate(hdr):
        if h == k:
            out.append(f"| {h} | {units[i]} | " + " | ".join(r[i] for r in keep) + " |")
if HBM:
    out[0] = f"# ncu --set full: memory-bound kernels K2 / K3 / K6 ({tag})\n"
    out[1] = ("Captured by `scripts/ncu_hbm_kernels.sh` from `HBM_NCU=1 scripts/bench_hbm_kernels.py` (q[4096,1024], cfg4-size "
              "T-mid stack, 1M-transition gather from a 4M-slot ring); one column per distinct (kernel, grid size). Durations "
              "under ncu are cold-cache and serialised: the warm device times are in `" + tag + "_secondary.jsonl` / "
              "`hbm_kernels_" + tag + ".jsonl` (CUDA events over graph replays). dram bytes = measured traffic per launch, "
              "to set beside the algorithmic bytes of the bench lines.\n")
if KIND == "cem":
    out[0] = f"# ncu --set full: cfg3 QT-Opt CEM predict_action ({tag})\n"
    out[1] = ("Captured by `scripts/ncu_cem.sh` (B=256, N=1024, 3 iterations, 2 mixture components, 400-300 T-mid critic): the "
              "state-term kernel (`k_mlp2_rows`, T-mid mode, 8 rows per CTA) and the one-CTA-per-state CEM kernel. Both are "
              "latency-bound (256 / 32 CTAs on 148 SMs); warm device times are in `scripts/time_cem_parts.py`'s output, DESIGN K4.\n")
dst = os.path.join(ROOT, "profiles", f"{tag}_{ {'hbm': 'hbm_kernels', 'cem': 'cem', 'sb': 'small_batch'}[KIND] }_ncu.md")
open(dst, "w").write("\n".join(out) + "\n")
print("wrote", dst, "kernels:", [r[kn].split("(")[0] for r in keep])
