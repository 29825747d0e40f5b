"""`ncu -i report.ncu-rep --page raw --csv` -> a compact per-launch table (the columns kept under profiles/).
usage: python tools/ncu_table.py raw.csv out.csv"""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in rows[hi + 2:]:
            if len(r) >= len(hdr):
                w.writerow([r[c] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
