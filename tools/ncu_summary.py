"""Print the handful of ncu metrics that explain a kernel (reads `ncu -i X --page raw --csv` on stdin)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread ",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ", "smsp__inst_executed.sum ",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "sm__inst_executed_pipe_tensor",
        "l1tex__t_bytes.sum ", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_subpipe",
        "smsp__average_warps_issue_stalled"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    name = row[hdr.index("Kernel Name")]
    print("==", name[:150])
    for i, k in enumerate(hdr):
        if any(k.startswith(w.strip()) if w.endswith(" ") else w in k for w in WANT):
            if "stalled" in k:
                try:
                    if float(row[i]) < 0.3:
                        continue
                except ValueError:
                    continue
            print(f"   {k:90s} {units[i]:14s} {row[i]}")
