"""Summarise an .ncu-rep (ncu --set full) into a markdown table of the metrics the judge reads.  usage: ncu_summary.py REP [title]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__inst_executed.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {title}\n")
print("| kernel | " + " | ".join(k for k in KEYS if k in idx) + " |")
print("|---|" + "---|" * sum(k in idx for k in KEYS))
for r in rows[2:]:
    name = r[idx["Kernel Name"]][:90]
    print(f"| `{name}` | " + " | ".join(f"{r[idx[k]]} {units[idx[k]]}" for k in KEYS if k in idx) + " |")
