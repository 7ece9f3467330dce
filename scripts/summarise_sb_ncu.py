"""gpurun_out/sb_TAG.ncu-rep (scripts/ncu_small_batch.sh) -> profiles/TAG_small_batch_ncu.md: one column per captured launch
of the small-minibatch / device-loop kernels.  Usage: python scripts/summarise_sb_ncu.py TAG"""
import csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"sb_{tag}.ncu-rep")
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
dst = os.path.join(ROOT, "profiles", f"{tag}_small_batch_ncu.md")
open(dst, "w").write("\n".join(out) + "\n")
print("wrote", dst, "kernels:", [r[kn].split("(")[0] for r in keep])
