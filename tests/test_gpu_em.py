"""`next` row f1/f2 of the scope table: batched multi-variable posteriors (families) and one EM step
(Bayes/EM.hs:36-66) against the oracle run sample by sample."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
from helpers import G, max_rel_err, random_net, same_bits

pytestmark = pytest.mark.gpu


def oracle_em(net, samples):
    """learnEM restated on the oracle: per sample changeEvidence + posterior family / parents, left-fold sums"""
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    ojt = OracleJT(onet)
    fam = [np.zeros([net.dims[x] for x in net.scopes[v]]) for v in range(net.n)]
    par = [np.zeros([1] + [net.dims[x] for x in net.scopes[v][1:]]) for v in range(net.n)]
    for row in samples:
        ojt.change_evidence([(v, int(s)) for v, s in enumerate(row) if s >= 0])
        for v in range(net.n):
            sc, vals = ojt.posterior(net.scopes[v])
            fam[v] = fam[v] + vals.reshape([net.dims[x] for x in sc]).transpose([sc.index(x) for x in net.scopes[v]])
            if len(net.scopes[v]) > 1:
                psc, pv = ojt.posterior(net.scopes[v][1:])
                par[v] = par[v] + pv.reshape([net.dims[x] for x in psc]).transpose([psc.index(x) for x in net.scopes[v][1:]])[None, ...]
            else:
                par[v] = par[v] + 1.0
    out = []
    for v in range(net.n):
        with np.errstate(divide="ignore", invalid="ignore"):
            out.append(np.where(par[v] == 0.0, 0.0, fam[v] / par[v]).reshape(-1))
    return out


def test_family_posteriors_batch_bitexact():
    net = random_net(77, n=12, states=(2, 3), max_parents=3)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    ojt = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values))
    ev = synth.evidence_matrix(net, 50, [1, 4, 9], 5)
    queries = [net.scopes[v] for v in range(net.n)]
    res = hb.JunctionTree(pnet).queries_batch(queries, ev)
    for r in range(0, 50, 7):
        ojt.change_evidence([(v, int(s)) for v, s in enumerate(ev[r]) if s >= 0])
        for q, (sc, dims, vals) in zip(queries, res):
            osc, ovals = ojt.posterior(q)
            assert sc == osc and same_bits(vals[r], ovals)


@pytest.mark.parametrize("seed", [3, 8])
def test_em_step_matches_the_reference_algorithm(seed):
    net = random_net(seed, n=9, states=(2, 3), max_parents=2)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    observed = [0, 2, 3, 5, 8]  # the other variables are missing in every sample
    samples = synth.evidence_matrix(net, 200, observed, 11 + seed)
    got = hb.learnEM(samples, pnet)
    want = oracle_em(net, samples)
    for v in range(net.n):
        assert max_rel_err(got[v], want[v]) < 1e-12
        d = net.dims[v]
        cols = got[v].reshape(d, -1).sum(0)
        assert np.all((np.abs(cols - 1) < 1e-12) | (cols == 0))  # a learnt CPT column is a distribution (or unseen)
