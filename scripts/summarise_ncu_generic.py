"""Raw-page summary of any gpurun_out/<name>.ncu-rep into profiles/<out>.md (run here, no GPU):
python scripts/summarise_ncu_generic.py OUT.md REP1 [REP2 ...]"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    out.append(f"## ncu --set full --clock-control none: {os.path.basename(rep)} (one column per captured launch; cold caches, serialised)\n")
    out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    out.append("|---|---|" + "---|" * (len(rows) - 2))
    out.append("| Kernel Name | | " + " | ".join(r[kn].replace("void ", "")[:40] for r in rows[2:]) + " |")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")
    out.append("")
open(os.path.join(ROOT, "profiles", sys.argv[1]), "w").write("\n".join(out) + "\n")
print("wrote profiles/" + sys.argv[1])
