"""Load accounting for single-level vs fused steps of the C2 / C3 programs (no GPU), and what a register cache keyed by the
input offset would leave for the single-level steps.   usage: python profiles/loadsim_single.py [C2|C3]"""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import hbayes_b200 as hb
from hbayes_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
net, observed = synth.config_c3() if which == "C3" else synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
jt = hb.JunctionTree(pnet, device=None)
prog = jt.program(sorted(observed))
dims = net.dims
cur_single = dyn_single = cur_fused = 0
prod_single = prod_fused = 0
for s in prog["physical"]:
    if s["kind"] != 0 or s["shared"]: continue
    out_scope, w = s["out_scope"], s["w_var"]
    red = [v for v in out_scope if v != w]
    U = dims[w] if w >= 0 else 1
    nch = (U + 1) // 2
    n_red = int(np.prod([dims[v] for v in red])) if red else 1
    L = s["levels"]
    terms = 1
    if len(L) == 1:
        l = L[0]
        prod_single += s["n_out"] * l["d"]
        for x in l["in"]:
            depw = w in x["scope"]
            per_item = l["d"] * (2 if depw else 1)
            cur_single += n_red * nch * per_item
            pos = [j for j, v in enumerate(red) if v in x["scope"]]
            runs = int(np.prod([dims[red[j]] for j in range(max(pos) + 1)])) if pos else 1
            dyn_single += runs * (nch if depw else 1) * per_item
    else:
        for l in L:
            terms *= l["d"]
            for x in l["in"]:
                depw = w in x["scope"]
                cur_fused += n_red * nch * terms * (2 if depw else 1)
        prod_fused += s["n_out"] * terms
print(which, "single-level: product entries", prod_single, "loads now", cur_single, "with offset cache", dyn_single)
print(which, "fused: product entries", prod_fused, "loads now", cur_fused)
