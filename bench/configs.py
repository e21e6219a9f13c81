#!/usr/bin/env python
"""Secondary measurements (not the driver's contract -- that is /bench.py): the other BASELINE.json configs.

  C1  cancer.net, single evidence set through the stateful reference API (latency) and 2^16-row batches
  C3  500-node random DAG (max cluster 2^20.2), 4096 evidence rows per batch, 125 observed
  C5  20x20 grid via VariableElimination (reference min-fill order), rows per batch given by --c5-rows

Each line: config, rows, ms per pass (CUDA events around the C-ABI call, device-resident buffers), queries/s,
algorithmic GB/s of the product/sum-out kernels, and the oracle port on the host cores for the same rows.
usage: python bench/configs.py [--only C1,C3,C5] [--c5-rows 64] > profiles/r1_configs.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet

THREADS = os.cpu_count() or 1


def time_batch(obj, fn, ev, reps=5):
    d_ev = torch.from_numpy(ev).cuda()
    out = fn(d_ev)
    obj.set_timing(True)
    for _ in range(2):
        fn(d_ev, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(d_ev, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st, tm = obj.last_stats(), obj.last_timing()
    gbs = 8.0 * (st["row_read_per_row"] + st["row_write_per_row"]) * ev.shape[0] / max(tm["prodsum_ms"], 1e-9) / 1e6
    return out, {"rows": int(ev.shape[0]), "ms_per_pass": ms, "queries_per_s": ev.shape[0] / ms * 1e3, "launches": st["launches"],
                 "tiles": st["tiles"], "prodsum_ms": tm["prodsum_ms"], "algorithmic_GBps": gbs,
                 "per_row": {k: st[k] for k in ("row_entries", "prod_per_row", "row_read_per_row", "row_write_per_row")}}


def c1():
    path = os.path.join(ROOT, "tests", "golden", "cancer.net")
    net = hb.importBayesianGraph(path)
    jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, net)
    ojt = OracleJT(OracleNet.load_hugin(path))
    sets = [[], [(0, 0)], [(0, 1)]]
    for ev in sets:  # compile once per observed set
        jt.changeEvidence(ev)
    t = time.perf_counter()
    n = 0
    for _ in range(200):
        for ev in sets:
            jt.changeEvidence(ev)
            for v in range(net.n):
                jt.posterior([v])
            n += 1
    lat = (time.perf_counter() - t) / n
    t = time.perf_counter()
    m = 0
    for _ in range(2000):
        for ev in sets:
            ojt.change_evidence(ev)
            for v in range(net.n):
                ojt.posterior([v])
            m += 1
    olat = (time.perf_counter() - t) / m
    rows = np.full((1 << 16, net.n), -1, np.int8)
    rows[:, 0] = np.arange(1 << 16) % 2
    _, r = time_batch(jt, jt.posteriors_batch, rows)
    r.update(config="C1 cancer.net", single_query_latency_us=lat * 1e6, oracle_single_query_latency_us=olat * 1e6,
             note="latency = changeEvidence + 5 posteriors through the stateful C-ABI calls (one 1-row pass, H2D+D2H+sync)")
    return r


def c3():
    net, observed = synth.config_c3()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.JunctionTree(pnet)
    ev = synth.config_c3_evidence(net, observed, 4096)
    out, r = time_batch(jt, jt.posteriors_batch, ev)
    ojt = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values))
    t = time.perf_counter()
    want = ojt.posteriors_batch(ev[:512], n_threads=THREADS)
    cpu = 512 / (time.perf_counter() - t)
    r.update(config="C3 500-node DAG, max cluster 2^20.2, 125 observed", cpu_port_queries_per_s=cpu, cpu_threads=THREADS,
             bit_identical_first_512=bool(out[:512].cpu().numpy().tobytes() == want.tobytes()))
    return r


def c5(rows):
    g = synth.config_c5(20, 20)
    pnet = hb.BayesianNetwork.from_tables(g.dims, g.scopes, g.values)
    p, rr, observed = synth.config_c5_query(g, hb.minFillOrder(pnet))
    ve = hb.VariableElimination(pnet, p, rr)
    ev = synth.evidence_matrix(g, rows, observed, 400003)
    out, r = time_batch(ve, ve.posterior_batch, ev, reps=2)
    onet = OracleNet.from_arrays(g.dims, g.scopes, g.values)
    k = min(rows, max(2, THREADS))
    t = time.perf_counter()
    want = onet.posterior_marginal_batch(p, rr, ev[:k], n_threads=THREADS)
    cpu = k / (time.perf_counter() - t)
    got = out[:k].cpu().numpy()
    r.update(config="C5 20x20 grid, VE with reference min-fill order, 40 observed", cpu_port_queries_per_s=cpu, cpu_threads=THREADS,
             bit_identical_first_rows=bool(got.tobytes() == want.tobytes()), rows_checked=k)
    return r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="C1,C3,C5")
    ap.add_argument("--c5-rows", type=int, default=64)
    a = ap.parse_args()
    for name in a.only.split(","):
        res = {"C1": c1, "C3": c3, "C5": lambda: c5(a.c5_rows)}[name]()
        print(json.dumps(res), flush=True)
