"""Prints the handful of ncu raw-page metrics the roofline discussion uses.  usage: ncu_table.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for w, i in idx:
    print("%-62s %s" % (w, units[i]))
for r in data:
    print(" | ".join(r[i][-24:] for _, i in idx))
