"""Turn gpurun_out/{k1_TAG.ncu-rep, launches_TAG.csv, bench_TAG.json} into committed summaries
under profiles/ (run here, no GPU needed):  python scripts/summarise_ncu.py TAG [KERNEL_REGEX]"""
import csv, io, json, os, re, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_uniform.sum", "lts__t_bytes.sum", "sm__cycles_active.avg"]
out = []
rep = os.path.join(G, f"k1_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out.append(f"## ncu --set full, kernel launches in k1_{tag}.ncu-rep (one column per captured launch)\n")
    out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    out.append("|---|---|" + "---|" * (len(rows) - 2))
    kn = hdr.index("Kernel Name")
    out.append("| Kernel Name | | " + " | ".join(r[kn][:48] for r in rows[2:]) + " |")
    traffic = None
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k or h.endswith("." + k):
                out.append(f"| {h} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")
    try:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        def tob(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic = sum(tob(r[ir], units[ir]) + tob(r[iw], units[iw]) for r in rows[2:]) / (len(rows) - 2)
        kname = rows[2][kn]
        m3 = re.search(r"grid3<[^,]*,\s*(?:\(int\))?(\d)", kname)
        prec = ("fp16c8" if m3 and m3.group(1) == "1" else "fp16x3") if "grid3" in kname else "fp16"
        tj = os.path.join(P, "k1_traffic.json")
        cur = json.load(open(tj)) if os.path.exists(tj) else {}
        cur["dram_bytes_per_launch_" + prec] = traffic
        cur["source_" + prec] = f"profiles/{tag}_k1_ncu.md (ncu --set full, {kname[:40]})"
        json.dump(cur, open(tj, "w"), indent=1)
    except Exception as e:
        print("traffic:", e)
launch_csv = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(launch_csv):
    rows = list(csv.reader(open(launch_csv)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
        name = re.sub(r"\(.*", "", r[kn])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f"\n## launch list (ncu --metrics gpu__time_duration.sum --clock-control none) of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity --sustain-s 0`\n")
    out.append("Cold-cache, serialised times: compare SHARES. The command runs warm-up + 2 timed steps, the K1-only loop, the e2e "
               "loops, the update-step loop (k_gemm/k_head_grads/k_colsum/k_adam + repack) and the full drop-in update.\n")
    out.append("| kernel | launches | total us | mean us | share |")
    out.append("|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {v[1] / tot:.3f} |")
    step = {k: v[1] / v[0] for k, v in agg.items()}
    k1 = [v for k, v in step.items() if "k_critic_umma" in k]
    fk = [v for k, v in step.items() if k.startswith("k_fkl") or k.startswith("k_policy_reduce")]
    if k1 and fk:
        pre = [v for k, v in step.items() if "k_grid3_parts" in k or "k_grid_parts" in k]
        lt = [v for k, v in step.items() if "k_grid_logterms" in k]
        rest = (pre[0] if pre else 0.0) + (lt[0] if lt else 0.0)
        out.append(f"\nOne bench *step* = pre-pass ({pre[0] if pre else 0:.1f} us) + K1 ({k1[0]:.1f} us) + grid log-terms "
                   f"({lt[0] if lt else 0:.1f} us) + policy-fused reduction ({fk[0]:.1f} us): K1's share of the step = "
                   f"{k1[0] / (k1[0] + fk[0] + rest):.3f}.")
bj = os.path.join(G, f"bench_{tag}.json")
if os.path.exists(bj) and os.path.getsize(bj):
    out.append(f"\n## bench line of the same build (gpurun_out/bench_{tag}.json)\n\n```json\n" + open(bj).read().strip() + "\n```")
rj = os.path.join(G, f"bench_ref_{tag}.json")
if os.path.exists(rj) and os.path.getsize(rj):
    out.append(f"\n## reference arm (`bench.py --impl reference`)\n\n```json\n" + open(rj).read().strip() + "\n```")
open(os.path.join(P, f"{tag}_k1_ncu.md"), "w").write("\n".join(out) + "\n")
print("wrote", os.path.join(P, f"{tag}_k1_ncu.md"))
