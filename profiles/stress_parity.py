"""Randomised parity stress (not part of the test-suite: ~10 minutes, mostly oracle time): 400 random networks,
random observed sets and row counts, junction tree and variable elimination, every double compared bit for bit with
the oracle.  Round-1 result on a B200: `stress done, failures: 0` (profiles/r1_stress.log)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
from helpers import random_net, same_bits
bad = 0
for seed in range(1000, 1400):
    net = random_net(seed, n=4 + seed % 23, states=(2, 2 + seed % 5), max_parents=1 + seed % 4, shuffle=seed % 3 == 0)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    try:
        pjt, ojt = hb.JunctionTree(pnet, cmp=seed % 2), OracleJT(onet, cmp=["weighted", "added"][seed % 2])
    except Exception as e:
        print("compile fail", seed, e); bad += 1; continue
    rng = np.random.RandomState(seed)
    observed = sorted(rng.permutation(net.n)[: rng.randint(0, net.n + 1)].tolist())
    rows = [1, 2, 5, 9, 33, 64, 100, 257][seed % 8]
    ev = np.full((rows, net.n), -1, np.int8)
    for v in observed:
        ev[:, v] = rng.randint(0, net.dims[v], rows)
    got, want = pjt.posteriors_batch(ev), ojt.posteriors_batch(ev, n_threads=8)
    if not same_bits(got, want):
        bad += 1
        print("MISMATCH seed", seed, "n", net.n, "obs", observed, "rows", rows, np.nanmax(np.abs(got - want)))
    # VE for one variable
    v = seed % net.n
    order = hb.minFillOrder(pnet)
    p = [x for x in order if x != v] + [x for x in range(net.n) if x not in order and x != v]
    ve = hb.VariableElimination(pnet, p, [v]).posterior_batch(ev)
    wv = onet.posterior_marginal_batch(p, [v], ev, n_threads=8)
    if not same_bits(ve, wv):
        bad += 1
        print("VE MISMATCH seed", seed)
print("stress done, failures:", bad)
