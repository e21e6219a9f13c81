"""Sums an ncu per-launch metric CSV (dram__bytes_read.sum, dram__bytes_write.sum, lts__t_bytes.sum, gpu__time_duration.sum) of
one pass of the hot path into the JSON that bench.py reads for `roofline.traffic`.
usage: sum_traffic.py pass_dram_ncu.csv algorithmic_bytes > r1_traffic.json"""
import collections
import csv
import io
import json
import sys

txt = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(txt))))
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(r["ID"], {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
rd = sum(d["dram__bytes_read.sum"] for d in L.values())
wr = sum(d["dram__bytes_write.sum"] for d in L.values())
fam = collections.Counter()
for d in L.values():
    n = d["name"]
    fam["hb_step (specialised)" if "hb_step_" in n else "fused_kernel" if "fused_kernel" in n else "prodsum_kernel" if "prodsum" in n else n[:40]] += d["gpu__time_duration.sum"]
tot = sum(fam.values())
print(json.dumps({
    "workload": "C2, 2^20 rows, one pass (%d product/sum-out launches)" % len(L),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
    "algorithmic_bytes": int(sys.argv[2]) if len(sys.argv) > 2 else None,
    "l2_bytes": sum(d.get("lts__t_bytes.sum", 0.0) for d in L.values()),
    "sum_kernel_ns_under_ncu": tot,
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none, profiles/r1_pass_dram_ncu.csv",
    "time_share_of_these_launches": {k: round(v / tot, 4) for k, v in fam.items()},
}, indent=1))
