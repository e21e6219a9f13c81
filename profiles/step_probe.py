#!/usr/bin/env python
"""Tuning probe for the per-step generated kernels (csrc/jit.cpp `hb_step_<i>`): one configuration per process, the
knobs come from the environment (they are read once per process):

  HB_JIT_INTERLEAVE  0 one contiguous strip of groups per lane | 1 groups dealt round robin for streaming steps | 2 for all
  HB_JIT_UNROLL      groups evaluated together (loads of all of them ahead of the first store)
  HB_JIT_WAVES       CTAs per SM a streaming step is cut into
  HB_JIT_MAX_KERNELS / HB_JIT_BUDGET   how many steps of a program get a kernel of their own
  HB_JIT_BLOCK_U / _V / _MINB / _BODY / _LOOPS / _MIN_JOINT   register blocks over several output variables for re-reading steps
  HB_JIT_DIGIT_ORDER, HB_JIT_SPLIT_LOOP, HB_JIT_STEP_ELIDE, HB_JIT_ZERO_ADD   (profiles/step_block_matrix.sh, csrc/jit.cpp jit_step_shape)
  HB_LANES           streams the steps of a pass overlap on inside the CUDA graph
  HB_JIT=0           precompiled kernels only

  python profiles/step_probe.py C3|C4|C5     -> one JSON line: ms per pass, algorithmic GB/s, sha1 of the result matrix

C4 here = ONE 2^27-entry slice of the 2^30-entry CPT (cutset form of the split clique, 4 rows): what one GPU of the
8-GPU configuration computes per pass, without the exchange."""
import hashlib
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

KNOBS = {k: v for k, v in os.environ.items() if k.startswith("HB_")}


def timed(obj, fn, ev, reps):
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
    return out, ms, st, tm


def main():
    which = sys.argv[1]
    t0 = time.perf_counter()
    if which == "C3":
        net, observed = synth.config_c3()
        pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
        obj = hb.JunctionTree(pnet)
        n_k = obj.specialise(observed)
        ev = synth.config_c3_evidence(net, observed, 4096)
        fn, reps = obj.posteriors_batch, 5
    elif which == "C5":
        g = synth.config_c5(20, 20)
        pnet = hb.BayesianNetwork.from_tables(g.dims, g.scopes, g.values)
        p, rr, observed = synth.config_c5_query(g, hb.minFillOrder(pnet))
        obj = hb.VariableElimination(pnet, p, rr)
        n_k = obj.specialise(observed)
        ev = synth.evidence_matrix(g, 64, observed, 400003)
        fn, reps = obj.posterior_batch, 3
    elif which == "C4":
        net, split = synth.config_c4(29)
        pnet = hb.BayesianNetwork.from_tables_procedural(net.dims, net.scopes, net.values, net.proc_seed)
        combo = [(v, 1 if i == 1 else 0) for i, v in enumerate(split)]
        obj = hb.JunctionTree(pnet, device=0, split=combo)
        observed = sorted(split + [net.n - 1])
        n_k = obj.specialise(observed)
        rng = np.random.RandomState(5)
        ev = np.full((4, net.n), -1, np.int8)
        ev[:, net.n - 1] = rng.randint(0, 2, 4)
        for v, s in combo:
            ev[:, v] = s
        fn, reps = obj.posteriors_batch, 5
    else:
        raise SystemExit("C3 | C4 | C5")
    setup_s = time.perf_counter() - t0
    out, ms, st, tm = timed(obj, fn, ev, reps)
    res = out.cpu().numpy()
    moved = 8.0 * (st["row_read_per_row"] + st["row_write_per_row"]) * ev.shape[0]
    print(json.dumps({"config": which, "knobs": KNOBS, "ms_per_pass": ms, "prodsum_ms": tm["prodsum_ms"], "kernels": n_k, "launches": st["launches"],
                      "algorithmic_GBps": moved / max(tm["prodsum_ms"], 1e-9) / 1e6, "setup_s": setup_s, "rows": int(ev.shape[0]),
                      "sha1": hashlib.sha1(res.tobytes()).hexdigest(), "nan": bool(np.isnan(res).any())}))


if __name__ == "__main__":
    main()
