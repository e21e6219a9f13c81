"""Counts the per-row load instructions (8-byte units) of every per-row step of the C2 program from its JSON description
(no GPU needed): as the kernels issue them today (blocking along w, pairs of states), and what a register cache keyed by
the input offset would leave for the single-level steps, with the current and with the best output order.  Used for the
load-pipe analysis in profiles/README.md and DESIGN.md section 8.   usage: python profiles/loadsim.py"""
import sys, itertools; sys.path.insert(0,'/root/repo')
import numpy as np
import hbayes_b200 as hb
from hbayes_b200 import synth
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
jt = hb.JunctionTree(pnet, device=None)
prog = jt.program(sorted(observed))
dims = net.dims
tot_cur = tot_dyn = tot_opt = 0
rows = []
for i, s in enumerate(prog["physical"]):
    if s["kind"] != 0 or s["shared"]: continue
    out_scope, w = s["out_scope"], s["w_var"]
    red = [v for v in out_scope if v != w]
    U = dims[w] if w >= 0 else 1
    nch = (U + 1) // 2
    L = s["levels"]
    inner = L[-1]
    # outer-level terms multiply the number of innermost evaluations
    outer_terms = int(np.prod([l["d"] for l in L[:-1]])) if len(L) > 1 else 1
    n_red = int(np.prod([dims[v] for v in red])) if red else 1
    def loads(order):
        # current: per item (o', ch): per inner input: d * (2 if dep on w else 1); outer-level inputs: ignore reuse (count as now)
        cur = dyn = 0
        last = {}
        for idx in itertools.product(*[range(dims[v]) for v in order]):
            st = dict(zip(order, idx))
            for ch in range(nch):
                for k, x in enumerate(inner["in"]):
                    depw = w in x["scope"]
                    n = inner["d"] * (2 if depw else 1) * outer_terms
                    cur += n
                    key = tuple(st[v] for v in x["scope"] if v in st) + ((ch,) if depw else ())
                    if outer_terms == 1:
                        if last.get(k) != key:
                            dyn += n
                            last[k] = key
                    else:
                        dyn += n  # fused: inner inputs also move with the outer odometer: no cross-item reuse
        return cur, dyn
    cur, dyn = loads(red)
    best = dyn
    if len(L) == 1 and len(red) <= 6:
        for perm in itertools.permutations(red):
            c, d = loads(list(perm))
            best = min(best, d)
    outer_loads = 0
    t = 1
    for l in L[:-1]:
        t *= l["d"]
        outer_loads += s["n_out"] * t * len(l["in"])
    tot_cur += cur + outer_loads; tot_dyn += dyn + outer_loads; tot_opt += best + outer_loads
    rows.append((cur + outer_loads, dyn + outer_loads, best + outer_loads, i, len(L)))
rows.sort(reverse=True)
print("total cur", tot_cur, "dyn(cur order)", tot_dyn, "dyn(best order)", tot_opt)
for r in rows[:14]: print(r)
