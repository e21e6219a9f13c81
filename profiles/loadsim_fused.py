"""Load accounting for the FUSED steps of the C2 / C3 programs (no GPU): loads per row as fused_kernel issues them today,
if every input were loaded once per combination of the summed variables it really depends on (loop-invariant hoisting
inside one output item), and if it were only hoisted out of the adjacent invariant loops.
usage: python profiles/loadsim_fused.py [C2|C3]"""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import hbayes_b200 as hb
from hbayes_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else "C2"
net, observed = synth.config_c3() if which == "C3" else synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
jt = hb.JunctionTree(pnet, device=None)
prog = jt.program(sorted(observed))
dims = net.dims
now = hoist_full = hoist_suffix = 0
rows = []
for idx, s in enumerate(prog["physical"]):
    if s["kind"] != 0 or s["shared"] or len(s["levels"]) < 2: continue
    w = s["w_var"]; U = dims[w] if w >= 0 else 1
    items = (s["n_out"] // U) * ((U + 1) // 2)
    L = s["levels"]
    a = b = c = 0
    terms = 1
    for li, l in enumerate(L):
        terms *= l["d"]
        for x in l["in"]:
            mult = 2 if w in x["scope"] else 1
            a += items * terms * mult
            # distinct combos of summed vars (levels 0..li) the input depends on
            dep = 1
            for m in range(li + 1):
                if x["lstride"][m] != 0: dep *= L[m]["d"]
            b += items * dep * mult
            # suffix hoisting: invariant to the innermost-adjacent outer levels only (contiguous from li-1 downwards)
            suf = L[li]["d"] if x["lstride"][li] != 0 else 1
            m = li - 1
            while m >= 0 and x["lstride"][m] == 0: m -= 1
            for q in range(m + 1): suf *= L[q]["d"]
            c += items * suf * mult
    now += a; hoist_full += b; hoist_suffix += c
    rows.append((a, b, c, idx))
rows.sort(reverse=True)
print(which, "fused loads/row: now", now, "| each input loaded once per combination it depends on", hoist_full, "| hoisted out of adjacent invariant loops only", hoist_suffix)
for r in rows[:8]: print(r)
