"""Summarise an .ncu-rep (read here, no GPU): python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--stalls]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("sm__inst_executed.avg.per_cycle_active", "ipc"), ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__issue_active.avg.pct", "issue%"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%")]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")] if "--stalls" in sys.argv else []
for r in rows[2:]:
    out = []
    for k, name in want:
        if k in hdr:
            i = hdr.index(k)
            v = r[i]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                v = v[:40]
            out.append(f"{name}={v}{units[i] if name in ('time','dram_rd','dram_wr') else ''}")
    print("  ".join(out))
    if stall:
        st = sorted(((float(r[hdr.index(h)] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:7]
        print("     stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in st))
