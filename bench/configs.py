#!/usr/bin/env python
"""Secondary measurements (not the driver's contract -- that is /bench.py): the other BASELINE.json configs.

  C1  cancer.net, single evidence set through the stateful reference API (latency) and 2^16-row batches
  C3  500-node random DAG (max cluster 2^20.2), 4096 evidence rows per batch, 125 observed
  C5  20x20 grid via VariableElimination (reference min-fill order), rows per batch given by --c5-rows
  MPE most probable explanation (Bayes/VariableElimination.hs:101-114) of the 28 unobserved variables of the C2 network
  EM  one learnEM step (Bayes/EM.hs:36-45) on the C2 network: 2^18 samples, 9 observed; then the learnt tables go
      back into the same handle with changeFactor (no recompilation)

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
    try:  # the roofline denominator bench.py uses for its own `roofline` object
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    return out, {"rows": int(ev.shape[0]), "ms_per_pass": ms, "queries_per_s": ev.shape[0] / ms * 1e3, "launches": st["launches"],
                 "tiles": st["tiles"], "prodsum_ms": tm["prodsum_ms"], "algorithmic_GBps": gbs, "hbm_frac": gbs / peak,
                 "per_row": {k: st[k] for k in ("row_entries", "prod_per_row", "row_read_per_row", "row_write_per_row")}}


def c1():
    path = os.path.join(ROOT, "tests", "golden", "cancer.net")
    net = hb.importBayesianGraph(path)
    jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, net)
    ojt = OracleJT(OracleNet.load_hugin(path))
    sets = [[], [(0, 0)], [(0, 1)]]
    for ev in sets:  # compile once per observed set; the on-chip kernel of each is built here, not inside the timed loop
        jt.specialise([v for v, _ in ev])
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
             note="latency = changeEvidence + 5 posteriors through the stateful calls (one launch + one synchronisation; evidence and "
                  "result in mapped pinned memory); single_query_latency_us goes through the Python binding (6 ctypes calls), "
                  "c_abi_single_query_latency_us is the same loop in plain C (bench/latency.c)")
    try:  # the same loop from plain C: the library's own latency, without the Python binding
        import subprocess
        import tempfile

        from hbayes_b200 import _lib

        exe = os.path.join(tempfile.mkdtemp(), "latency")
        libdir = os.path.dirname(_lib.LIB_PATH)
        subprocess.run(["gcc", "-std=c99", "-O2", os.path.join(ROOT, "bench", "latency.c"), "-o", exe, "-L" + libdir, "-lhbayes_cuda",
                        "-Wl,-rpath," + libdir], check=True, capture_output=True)
        r.update(json.loads(subprocess.run([exe, path, "3000"], check=True, capture_output=True, text=True).stdout))
    except Exception as e:
        r["c_abi_latency_error"] = repr(e)[:200]
    return r


def c3():
    net, observed = synth.config_c3()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.JunctionTree(pnet)
    jt.specialise(observed)  # generated kernels built up front (NVRTC or the cubin cache), not inside a timed pass
    ev = synth.config_c3_evidence(net, observed, 4096)
    out, r = time_batch(jt, jt.posteriors_batch, ev)
    r["specialised_kernels"] = jt.specialised_kernels()
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
    ve.specialise(observed)
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


def em(rows=1 << 18):
    import oracle

    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.JunctionTree(pnet)
    ev = synth.config_c2_evidence(net, observed, rows)
    d_ev = torch.from_numpy(ev).cuda()
    learnt = jt.em_step(d_ev)  # compiles the family-query program, sizes the arena
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = 3
    for _ in range(reps):
        learnt = jt.em_step(d_ev)
    ms = (time.perf_counter() - t) / reps * 1e3
    st = jt.last_stats()
    # where the time goes: one chunk, CUDA-event time of the junction-tree pass alone (the rest is the sample-order sum)
    os.environ["HB_EM_CHUNK_ROWS"] = str(rows)
    jt.set_timing(True)
    jt.em_step(d_ev)
    t = time.perf_counter()
    jt.em_step(d_ev)
    one_chunk_ms = (time.perf_counter() - t) * 1e3
    pass_ms = jt.last_timing()["total_ms"]
    del os.environ["HB_EM_CHUNK_ROWS"]
    t = time.perf_counter()
    for v, f in enumerate(learnt):
        perm = [f.variables.index(x) for x in net.scopes[v]]
        own = np.ascontiguousarray(f.values.reshape(f.dims).transpose(perm)).reshape(-1)
        jt.change_factor(hb.Factor(net.scopes[v], [net.dims[x] for x in net.scopes[v]], own))
    change_ms = (time.perf_counter() - t) * 1e3
    k = 2048
    t = time.perf_counter()
    want = oracle.learn_em(net.dims, net.scopes, net.values, ev[:k])
    cpu = k / (time.perf_counter() - t)
    got = hb.JunctionTree(pnet).em_step(ev[:k])
    same = all(g.variables == w[0] and g.values.tobytes() == w[1].tobytes() for g, w in zip(got, want))
    width = sum(int(np.prod([net.dims[x] for x in sc])) + (int(np.prod([net.dims[x] for x in sc[1:]])) if len(sc) > 1 else 0)
                for sc in net.scopes)
    return {"config": "EM step on the C2 network (37 nodes, 9 observed)", "rows": rows, "ms_per_step": ms,
            "samples_per_s": rows / ms * 1e3, "launches": st["launches"], "result_columns_per_sample": width,
            "one_chunk_ms": one_chunk_ms, "one_chunk_tree_pass_ms": pass_ms,
            "change_all_factors_ms": change_ms, "cpu_port_samples_per_s": cpu, "cpu_threads": 1,
            "bit_identical_first_%d" % k: bool(same)}


def greedy_order(net, observed):
    """elimination order of the unobserved variables: greedy smallest resulting table, WITH fill-in edges (the
    reference's minFillOrder adds none and skips isolated vertices, so it is a poor order once p is forced first)"""
    adj = {v: set() for v in range(net.n) if v not in observed}
    for sc in net.scopes:
        fam = [x for x in sc if x in adj]
        for a in fam:
            adj[a].update(b for b in fam if b != a)
    order = []
    while adj:
        v = min(adj, key=lambda x: (np.prod([net.dims[y] for y in adj[x]] + [net.dims[x]]), x))
        nb = adj.pop(v)
        for a in nb:
            adj[a].discard(v)
            adj[a].update(b for b in nb if b != a)
        order.append(v)
    return order


def mpe(rows=1 << 18):
    """most probable explanation of all 28 unobserved variables of the C2 network per evidence row"""
    from oracle import twin

    net, _ = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    # evidence on the 9 non-isolated variables of lowest degree: the reference's algorithm sums the evidence variables out
    # FIRST, i.e. multiplies every table that mentions one into a single product -- C2's own observed set holds hubs
    # of degree 9 and that product alone would be a 5.3 M-entry table per row
    deg = {v: set() for v in range(net.n)}
    for sc in net.scopes:
        for a in sc:
            deg[a].update(b for b in sc if b != a)
    observed = sorted(sorted([v for v in deg if deg[v]], key=lambda v: (len(deg[v]), -v))[:9])
    p = list(observed)
    r = greedy_order(net, observed)
    h = hb.MostProbableExplanation(pnet, p, r)
    ev = synth.evidence_matrix(net, rows, observed, 37004)
    d_ev = torch.from_numpy(ev).cuda()
    states, values = h.explain_batch(d_ev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        h.explain_batch(d_ev, states, values)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = h.last_stats()
    # size-independent property on every row: the value IS the joint probability of (explanation, evidence)
    full = ev.copy()
    full[:, r] = states.cpu().numpy()
    joint = np.ones(rows)
    for v in range(net.n):
        sc = net.scopes[v]
        idx = np.ravel_multi_index([full[:, x].astype(np.int64) for x in sc], [net.dims[x] for x in sc])
        joint *= np.asarray(net.values[v])[idx]
    vals = values.cpu().numpy()
    rel = float(np.max(np.abs(joint - vals) / vals))
    # twin parity on a few rows (pure Python: seconds per row)
    tnet = twin.Network.from_tables(net.dims, net.parents, [list(map(float, v)) for v in net.values])
    same = True
    k = 3
    t = time.perf_counter()
    for row in range(k):
        tv, ti = twin.mpe(tnet, p, r, [(v, int(s)) for v, s in enumerate(ev[row]) if s >= 0])
        want = dict(ti[0])
        same = same and [want[v] for v in r] == full[row, r].tolist() and np.float64(tv).tobytes() == vals[row].tobytes()
    cpu = k / (time.perf_counter() - t)
    return {"config": "MPE of the 28 unobserved variables of the C2 network (9 low-degree variables observed, greedy min-size order)", "rows": rows,
            "ms_per_pass": ms, "explanations_per_s": rows / ms * 1e3, "launches": st["launches"],
            "value_equals_joint_max_rel_err": rel, "twin_rows_checked": k, "bit_identical_to_twin": bool(same),
            "python_twin_explanations_per_s": cpu}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="C1,C3,C5,EM,MPE")
    ap.add_argument("--c5-rows", type=int, default=64)
    a = ap.parse_args()
    for name in a.only.split(","):
        res = {"C1": c1, "C3": c3, "C5": lambda: c5(a.c5_rows), "EM": em, "MPE": mpe}[name]()
        print(json.dumps(res), flush=True)
