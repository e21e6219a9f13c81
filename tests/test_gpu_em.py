"""`next` row f1/f2 of the scope table: batched multi-variable posteriors (families) and one EM step
(Bayes/EM.hs:36-66) against the oracle run sample by sample."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
import oracle
from oracle import OracleJT, OracleNet
from helpers import G, max_rel_err, random_net, same_bits

pytestmark = pytest.mark.gpu


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


def check_em(net, samples):
    """learnEM bit for bit: the reference's layout ([parents..., child], cptDivide) and the CPT's own"""
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    want = oracle.learn_em(net.dims, net.scopes, net.values, samples)
    got = hb.learnEM(samples, pnet, reference_layout=True)
    for v in range(net.n):
        assert got[v].variables == want[v][0]
        assert same_bits(got[v].values, want[v][1]), (v, got[v].values, want[v][1])
    own = hb.learnEM(samples, pnet)
    for v in range(net.n):
        sc, vals = want[v]
        dims = [net.dims[x] for x in sc]
        assert same_bits(own[v], vals.reshape(dims).transpose([sc.index(x) for x in net.scopes[v]]).reshape(-1))
    return own


@pytest.mark.parametrize("seed", [3, 8])
def test_em_step_bit_identical_to_the_reference_algorithm(seed):
    net = random_net(seed, n=9, states=(2, 3), max_parents=2)
    observed = [0, 2, 3, 5, 8]  # the other variables are missing in every sample
    samples = synth.evidence_matrix(net, 200, observed, 11 + seed)
    own = check_em(net, samples)
    for v in range(net.n):
        cols = own[v].reshape(net.dims[v], -1).sum(0)
        assert np.all((np.abs(cols - 1) < 1e-12) | (cols == 0))  # a learnt CPT column is a distribution (or unseen)


def test_em_step_mixed_observed_sets_and_chunks(monkeypatch):
    """samples with different missing variables: the programs (and result layouts) differ per observed set, the sum
    still runs sample after sample; HB_EM_CHUNK_ROWS forces several accumulation chunks"""
    net = random_net(5, n=8, states=(2, 3), max_parents=2, shuffle=True)
    rng = np.random.RandomState(4)
    full = np.stack([rng.randint(0, d, size=90) for d in net.dims], axis=1).astype(np.int8)  # ids are not topological
    masks = [[0, 1, 2, 3, 4, 5, 6, 7], [1, 3, 4, 6], [0, 2, 7], []]
    samples = np.full_like(full, -1)
    for r in range(full.shape[0]):
        keep = masks[rng.randint(len(masks))]
        samples[r, keep] = full[r, keep]
    monkeypatch.setenv("HB_EM_CHUNK_ROWS", "16")
    check_em(net, samples)
    monkeypatch.setenv("HB_EM_CHUNK_ROWS", "1")
    check_em(net, samples[:5])


def test_em_step_single_sample_and_device_evidence():
    import torch

    net = random_net(12, n=7, states=(2, 3), max_parents=2)
    samples = synth.evidence_matrix(net, 70, [1, 2, 6], 9)
    check_em(net, samples[:1])  # foldl1 of one sample: no sum at all
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.JunctionTree(pnet)
    a = jt.em_step(samples)
    b = jt.em_step(torch.from_numpy(samples).cuda())
    for fa, fb in zip(a, b):
        assert fa.variables == fb.variables and same_bits(fa.values, fb.values)


def test_em_zero_probability_sample_gives_nan_like_the_reference():
    net = random_net(3, n=5, states=(2, 2), max_parents=1)
    values = [np.array(v, np.float64) for v in net.values]
    values[0][:] = [1.0, 0.0]  # vertex 0 is never in state 1 ...
    net = synth.SynthNet(net.dims, net.parents, [v.tolist() for v in values])
    samples = synth.evidence_matrix(net, 6, [0, 1], 2)
    samples[3, 0] = 1  # ... but one sample says it is: posterior = 0 * (1/0) = NaN, and NaN stays in the sums
    check_em(net, samples)


def test_iterated_em_through_change_factor_increases_likelihood():
    """EM loop on one handle: learnt tables go back in with changeFactor (schedules kept); the data likelihood
    (sum of log P(evidence), from the VE engine) must not decrease"""
    net = random_net(21, n=8, states=(2, 2), max_parents=2)
    truth = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    observed = [0, 2, 3, 5, 7]
    samples = synth.evidence_matrix(net, 400, observed, 1)
    rng = np.random.RandomState(0)
    start = []
    for v in range(net.n):
        t = rng.uniform(0.2, 0.8, size=(net.dims[v], len(net.values[v]) // net.dims[v]))
        start.append((t / t.sum(0)).reshape(-1))
    cur = hb.BayesianNetwork.from_tables(net.dims, net.scopes, start)
    jt = hb.JunctionTree(cur)

    def loglik(pnet):
        # P(e) = norm of the unnormalised VE result; here via chain: prod over observed of posterior-predictive is
        # not needed -- use the joint posterior of the observed variables' first one under growing evidence
        total = 0.0
        j = hb.JunctionTree(pnet)
        for k, v in enumerate(observed):
            ev = samples.copy()
            ev[:, observed[k:]] = -1  # condition on the earlier observed variables only
            post = j.posteriors_batch(ev)
            off = int(sum(net.dims[:v]))
            total += float(np.log(post[np.arange(len(samples)), off + samples[:, v]]).sum())
        return total

    ll = [loglik(cur)]
    for _ in range(3):
        learnt = jt.em_step(samples)
        for v, f in enumerate(learnt):
            perm = [f.variables.index(x) for x in net.scopes[v]]
            own = np.ascontiguousarray(f.values.reshape(f.dims).transpose(perm)).reshape(-1)
            assert jt.change_factor(hb.Factor(net.scopes[v], [net.dims[x] for x in net.scopes[v]], own))
        ll.append(loglik(jt.net))
    assert all(b >= a - 1e-9 for a, b in zip(ll, ll[1:])), ll
    assert ll[-1] > ll[0]
    del truth
