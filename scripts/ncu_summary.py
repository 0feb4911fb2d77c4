"""Summarise .ncu-rep captures (ncu --set full) into a small CSV kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/ncu_summary_rNN.csv"""
import csv, subprocess, sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]

w = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(x) for x in WANT if x in hdr]
    if first:
        w.writerow(["report"] + [f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
        first = False
    for r in rows[2:]:
        w.writerow([rep.split("/")[-1]] + [r[i][:90] for i in idx])
