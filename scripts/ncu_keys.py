"""Print the key metrics + stall breakdown of the first kernel in an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:100])
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
    for k in keys:
        if k in d: print(f"  {k} = {d[k]}")
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v or 0) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
    tot = sum(st.values()) or 1
    print("  stalls:", ", ".join(f"{k} {100*v/tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
