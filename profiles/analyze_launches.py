"""Joins an ncu launch list (gpu__time_duration.sum per launch) of `bench.py --rows R` with the compiled C2
program: per-launch algorithmic bytes, achieved GB/s and time shares.  usage: analyze_launches.py launches.csv R"""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import hbayes_b200 as hb
from hbayes_b200 import synth

path, R = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]
dur = [float(r[-1]) for r in rows]
grid = [r[8] for r in rows]
blk = [r[7] for r in rows]
idx = [i for i, n in enumerate(names) if "prep_rows" in n]
s, e = idx[-2], idx[-1]
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
p = hb.JunctionTree(pnet, device=None).program(observed)
steps = [x for x in p["physical"] if x["kind"] == 0 and not x["shared"]]
L = [i for i in range(s + 1, e) if "prodsum" in names[i] or "fused" in names[i] or "staged" in names[i] or "hb_step_" in names[i]]
assert len(L) == len(steps), (len(L), len(steps))
if any("hb_step_" in n for n in names):
    # final engine: launches are issued by dependency level and the heavy steps run on kernels named after their step
    # (hb_step_<index into the program's step list>); only those can be matched to a step by name
    by_si = {si: x for si, x in enumerate(p["physical"])}
    pairs = [(i, by_si[int(names[i].split("hb_step_")[1].split("(")[0])]) for i in L if "hb_step_" in names[i]]
    L, steps = [a for a, _ in pairs], [b for _, b in pairs]
    print("matched %d specialised launches by name; the templated launches of this pass are not attributed" % len(L))
tot = totb = 0
recs = []
for li, st in zip(L, steps):
    rd = loads = potloads = 0
    terms = 1
    for lv in st["levels"]:
        terms *= lv["d"]
        for i in lv["in"]:
            loads += st["n_out"] * terms
            if i["kind"] == 3:
                rd += min(i["size"], st["n_out"] * terms)
            elif i["kind"] == 1:
                potloads += st["n_out"] * terms
    b = 8 * (rd + st["n_out"]) * R
    k = sum(len(lv["in"]) for lv in st["levels"])
    recs.append((dur[li], b / dur[li], st["n_out"], len(st["levels"]), k, loads, potloads, 1e3 * dur[li] / R / max(loads, 1), grid[li]))
    tot += dur[li]
    totb += b
others = sum(dur[i] for i in range(s, e) if "prodsum" not in names[i] and "fused" not in names[i] and "staged" not in names[i])
print("pass: %d product/sum-out launches %.1f us (%.1f%% of pass), other kernels %.1f us; algorithmic %.0f GB/s (cold-cache, serialised)"
      % (len(L), tot / 1e3, 100 * tot / (tot + others), others / 1e3, totb / tot))
recs.sort(reverse=True)
print("top launches:")
for r in recs[:25]:
    print("  %8.0f ns %6.0f GB/s n_out=%5d levels=%d k=%d loads/row=%5d (pot %5d)  %.3f ps/load/row grid=%s" % r)
tl = sum(r[5] for r in recs)
print("loads per row %d, potential loads %d; mean %.3f ps per load per row" % (tl, sum(r[6] for r in recs), 1e3 * tot / R / tl))
for i in range(s, e):
    if "prodsum" not in names[i] and "fused" not in names[i] and "staged" not in names[i]:
        print("  other:", names[i][-60:], dur[i], "ns")
