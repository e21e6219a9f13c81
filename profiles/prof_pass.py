"""Profiling driver: C2 workload, `rows` evidence rows resident on the device, `passes` passes of the hot path.
usage: prof_pass.py [rows] [passes]      (run plain first, then under ncu; see profiles/README.md)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import hbayes_b200 as hb
from hbayes_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, pnet, device=0)
ev = torch.from_numpy(synth.evidence_matrix(net, rows, observed, 37003)).cuda()
out = torch.empty((rows, pnet.width), dtype=torch.float64, device="cuda")
jt.set_timing(True)
for _ in range(passes):
    jt.posteriors_batch(ev, out)
print(jt.last_stats(), jt.last_timing())
