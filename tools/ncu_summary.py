"""Condenses an `ncu --page raw --csv` log into the per-kernel metric blocks kept under profiles/.
  python tools/ncu_summary.py raw.csv > profiles/<name>.txt"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_uniform.sum"]
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head, units, data = rows[0], rows[1], rows[2:]
col = {h: j for j, h in enumerate(head)}
for r in data:
    print("-----")
    print(f"{'Kernel Name':84s} {r[col['Kernel Name']]}")
    for k in KEEP:
        if k in col and r[col[k]] != "":
            print(f"{k:84s} {r[col[k]]:>16s} {units[col[k]]}")
