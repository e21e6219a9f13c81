import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
jt = hb.JunctionTree(pnet)
ev = synth.config_c2_evidence(net, observed, 333)
out = jt.posteriors_batch(ev)
want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev, n_threads=8)
assert out.tobytes() == want.tobytes()
g = synth.config_c5(8, 8)
gn = hb.BayesianNetwork.from_tables(g.dims, g.scopes, g.values)
p, r, obs = synth.config_c5_query(g, hb.minFillOrder(gn))
ve = hb.VariableElimination(gn, p, r)
print(ve.posterior_batch(synth.evidence_matrix(g, 70, obs, 5)).sum())
f = hb.factorProduct([hb.Factor([1, 2], [2, 3], np.arange(6.0)), hb.Factor([2], [3], np.ones(3))])
print("sanitizer run ok", f.values.sum())
