import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
net, observed = synth.config_c2()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
ev = synth.config_c2_evidence(net, observed, rows)
d_ev = torch.from_numpy(ev).cuda()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
jt = hb.JunctionTree(pnet)
t = time.time(); out = jt.posteriors_batch(d_ev); torch.cuda.synchronize(); print("first pass (incl. JIT build) s", time.time() - t)
jt.set_timing(True)
for _ in range(3): jt.posteriors_batch(d_ev, out)
torch.cuda.synchronize()
print("timing", jt.last_timing(), jt.last_stats()["launches"])
want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[:512], n_threads=8)
got = out[:512].cpu().numpy()
print("bit-identical to oracle (512 rows):", got.tobytes() == want.tobytes(), "max abs diff", float(np.abs(got - want).max()))
