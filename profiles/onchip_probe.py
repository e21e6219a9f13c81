"""Probe of the whole-schedule on-chip kernel (hb_tree) on C2: parity against the oracle and the templated kernels, then
device-resident timing.  Usage: python profiles/onchip_probe.py [rows]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
ev = synth.evidence_matrix(net, rows, observed, 37003)
d_ev = torch.from_numpy(ev).cuda()
d_out = torch.empty((rows, pnet.width), dtype=torch.float64, device="cuda")

os.environ["HB_JIT"] = "0"
jt0 = hb.JunctionTree(pnet)
ref = jt0.posteriors_batch(d_ev[:65536].contiguous()).cpu().numpy()
os.environ["HB_JIT"] = "1"
jt = hb.JunctionTree(pnet)
t = time.time()
n = jt.specialise(observed)
print("specialise", n, "kernels in %.2f s" % (time.time() - t), jt.kernel_info(), flush=True)
jt.set_stream(torch.cuda.current_stream().cuda_stream)
got = jt.posteriors_batch(d_ev, d_out)
torch.cuda.synchronize()
g = got[:65536].cpu().numpy()
print("bit-identical to templated kernels on 65536 rows:", g.tobytes() == ref.tobytes(), flush=True)
onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
want = OracleJT(onet).posteriors_batch(ev[:4096], n_threads=os.cpu_count())
print("bit-identical to oracle on 4096 rows:", g[:4096].tobytes() == want.tobytes(), flush=True)
tail = got[-3:].cpu().numpy()
want_tail = OracleJT(onet).posteriors_batch(ev[-3:])
print("tail rows ok:", tail.tobytes() == want_tail.tobytes())
for _ in range(3):
    jt.posteriors_batch(d_ev, d_out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for _ in range(K):
    jt.posteriors_batch(d_ev, d_out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"rows": rows, "ms_per_pass": ms, "Mq_per_s": rows / ms / 1e3, "info": jt.kernel_info(), "stats": jt.last_stats()}))
if os.environ.get("HB_TREE_PROFILE"):
    jt.tree_profile()
    jt.posteriors_batch(d_ev, d_out)
    torch.cuda.synchronize()
    prof = jt.tree_profile()
    n = jt.kernel_info()["levels"] + 2
    tot = float(sum(prof[:n])) or 1.0
    print("phase cycles %:", " ".join("%.1f" % (100.0 * x / tot) for x in prof[:n]), "| cycles per tile:", tot / jt.last_stats()["tiles"])
