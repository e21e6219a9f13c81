"""The drop-in API on all GPUs of the box through ONE handle and ONE call: hb_jt_compile over devices[],
a page-locked evidence matrix of N x 2^20 C2 rows, hb_jt_posteriors_batch shards it over the GPUs (one host thread and one
GPU per slice).  Prints one JSON line: end-to-end queries/s (host buffers in and out), and the same rows on one GPU.
usage: python profiles/multi_gpu_api.py [rows_per_gpu]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import hbayes_b200 as hb
from hbayes_b200 import _lib, synth

rows_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n = _lib.lib().hb_device_count()
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
rows = rows_per_gpu * n
ev = hb.host_alloc((rows, net.n), np.int8)
ev[:] = synth.evidence_matrix(net, rows, observed, 37003)
out = hb.host_alloc((rows, pnet.width), np.float64)
res = {"n_gpus": n, "rows": rows}
for label, devices in (("all_gpus", list(range(n))), ("one_gpu", [0])):
    jt = hb.JunctionTree(pnet, device=devices)
    jt.specialise(observed)
    m = rows if label == "all_gpus" else rows_per_gpu
    jt.posteriors_batch(ev[:m], out[:m])
    t = time.perf_counter()
    reps = 3
    for _ in range(reps):
        jt.posteriors_batch(ev[:m], out[:m])
    dt = (time.perf_counter() - t) / reps
    res[label] = {"rows": m, "ms_per_call": dt * 1e3, "queries_per_s": m / dt, "d2h_GBps": m * pnet.width * 8 / dt / 1e9}
    if label == "all_gpus":
        keep = out[:4096].copy(), out[-4096:].copy()
one = hb.JunctionTree(pnet, device=0)
res["slices_equal_single_gpu"] = bool(np.array_equal(one.posteriors_batch(ev[:4096]), keep[0]) and np.array_equal(one.posteriors_batch(ev[-4096:]), keep[1]))
print(json.dumps(res))
