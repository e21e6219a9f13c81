import os, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import hbayes_b200 as hb
from hbayes_b200 import synth
rows = 1 << 20
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
ev = synth.evidence_matrix(net, rows, observed, 37003)
jt = hb.JunctionTree(pnet); jt.specialise(observed)
stream = torch.cuda.current_stream(); jt.set_stream(stream.cuda_stream)
pev = torch.from_numpy(ev).pin_memory(); pout = torch.empty((rows, pnet.width), dtype=torch.float64, pin_memory=True)
d_out = jt.posteriors_batch(pev.cuda())
for _ in range(2): jt.posteriors_batch(pev, pout)
torch.cuda.synchronize()
ts=[]
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); jt.posteriors_batch(pev, pout); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("HB_ZERO_COPY","1"), "e2e ms", ts, "equal", bool(torch.equal(pout, d_out.cpu())))
