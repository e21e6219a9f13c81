"""`next` row f4: max-product.  `mpe` (Bayes/VariableElimination.hs:79-114, Bayes/Factor/MaxCPT.hs:25-75) through the C ABI
against the pure-Python restatement (oracle/twin.py: mpe), which reproduces the reference's own asia goldens
(tests/test_oracle.py).  Checked: the instantiation list literally (variables, their order, states -- ties go to the LAST
maximal state) and the MAXCPT scalar bit for bit."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import twin
from helpers import G, random_net, same_bits

pytestmark = pytest.mark.gpu
YES, NO = 0, 1


def twin_net(net):
    return twin.Network.from_tables(net.dims, net.parents, [list(map(float, v)) for v in net.values])


def test_asia_mpe_and_map_goldens():
    """Bayes/Test/ReferencePatterns.hs:75-84"""
    path = os.path.join(G, "asia.net")
    pnet, tnet = hb.importBayesianGraph(path), twin.load_hugin(path)
    i = pnet.ids
    x, b, d, a, s, l, t, e = [i[c] for c in "XBDASLTE"]
    named = lambda m: [[(pnet.names[v], st) for v, st in inst] for inst in m]
    m = hb.mpe(pnet, [x, d], [b, a, s, l, t, e], [(x, YES), (d, NO)])
    assert named(m) == [[("E", NO), ("T", NO), ("L", NO), ("S", NO), ("B", NO), ("A", NO)]]
    m = hb.mpe(pnet, [x, d, b, l, t, e], [a, s], [(x, YES), (d, NO)])
    assert named(m) == [[("S", YES), ("A", NO)]]
    for p, r in (([x, d], [b, a, s, l, t, e]), ([x, d, b, l, t, e], [a, s])):
        value, inst = hb.MostProbableExplanation(pnet, p, r).explain([(x, YES), (d, NO)])
        tv, ti = twin.mpe(tnet, p, r, [(x, YES), (d, NO)])
        assert inst == ti and same_bits([value], [tv])


def test_sprinkler_tutorial_queries():
    """Bayes/Examples/Tutorial.hs:404-414 (mpeStandardNetwork): no expected values in the source; twin parity"""
    path = os.path.join(G, "sprinkler.net")
    pnet, tnet = hb.importBayesianGraph(path), twin.load_hugin(path)
    winter, sprinkler, wet, rain, road, roadandrain = range(6)
    ev = [(wet, 1), (road, 1)]
    for p, r in (([wet, road], [winter, sprinkler, rain, roadandrain]), ([wet, road, roadandrain, winter], [sprinkler, rain])):
        value, inst = hb.MostProbableExplanation(pnet, p, r).explain(ev)
        tv, ti = twin.mpe(tnet, p, r, ev)
        assert inst == ti and same_bits([value], [tv])


def split(rng, n, observed, n_sum):
    """p = the observed variables and n_sum others (shuffled), r = the rest (shuffled)"""
    others = [v for v in rng.permutation(n).tolist() if v not in observed]
    p = list(observed) + others[:n_sum]
    rng.shuffle(p)
    return [int(v) for v in p], [int(v) for v in others[n_sum:]]


@pytest.mark.parametrize("seed", range(12))
def test_random_networks_match_the_twin(seed):
    rng = np.random.RandomState(100 + seed)
    net = random_net(seed, n=5 + seed % 5, states=(2, 3), max_parents=2, shuffle=seed % 3 == 0)
    pnet, tnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values), twin_net(net)
    for _ in range(4):
        k = int(rng.randint(0, 3))
        observed = [int(v) for v in rng.permutation(net.n)[:k]]
        p, r = split(rng, net.n, observed, int(rng.randint(0, net.n - k)))
        if not r:
            continue
        ev = [(v, int(rng.randint(net.dims[v]))) for v in observed]
        value, inst = hb.MostProbableExplanation(pnet, p, r).explain(ev)
        tv, ti = twin.mpe(tnet, p, r, ev)
        assert inst == ti, (p, r, ev)
        assert same_bits([value], [tv])


def test_ties_pick_the_last_maximal_state():
    """uniform tables: every maximisation is a tie, and maximumBy keeps the last element (MaxCPT.hs:41)"""
    dims, parents = [3, 2, 4], [[], [0], [0, 1]]
    values = [np.full(3, 0.25), np.full(6, 0.5), np.full(24, 0.25)]  # exact in binary: products tie bit for bit
    pnet = hb.BayesianNetwork.from_tables(dims, [[0], [1, 0], [2, 0, 1]], values)
    tnet = twin.Network.from_tables(dims, parents, [v.tolist() for v in values])
    for r in ([0, 1, 2], [2, 1, 0], [1, 2, 0]):
        value, inst = hb.MostProbableExplanation(pnet, [], r).explain([])
        tv, ti = twin.mpe(tnet, [], r, [])
        assert inst == ti and same_bits([value], [tv])
        assert sorted(inst[0]) == [(0, 2), (1, 1), (2, 3)]


def test_batch_rows_mixed_observed_sets_and_device_buffers():
    import torch

    net = random_net(31, n=9, states=(2, 3), max_parents=2)
    pnet, tnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values), twin_net(net)
    observed = [1, 4, 6]
    p, r = [4, 0, 6, 1, 8], [7, 2, 5, 3]
    h = hb.MostProbableExplanation(pnet, p, r)
    ev = synth.evidence_matrix(net, 300, observed, 5)
    ev[::7, 4] = -1  # some rows observe less
    states, values = h.explain_batch(ev)
    for row in list(range(0, 300, 13)) + [7, 14]:
        asg = [(v, int(s)) for v, s in enumerate(ev[row]) if s >= 0]
        tv, ti = twin.mpe(tnet, p, r, asg)
        want = dict(ti[0])
        assert [want[v] for v in r] == states[row].tolist() and same_bits([values[row]], [tv])
    uniform = np.ascontiguousarray(ev[ev[:, 4] >= 0])
    d_states, d_values = h.explain_batch(torch.from_numpy(uniform).cuda())
    keep = np.flatnonzero(ev[:, 4] >= 0)
    assert np.array_equal(d_states.cpu().numpy(), states[keep]) and same_bits(d_values.cpu().numpy(), values[keep])


def test_errors():
    pnet = hb.importBayesianGraph(os.path.join(G, "cancer.net"))
    h = hb.MostProbableExplanation(pnet, [0, 1], [2, 3, 4])
    with pytest.raises(hb.HbError):
        h.explain([(2, 0)])  # evidence on a maximised variable
    with pytest.raises(hb.HbError):
        hb.MostProbableExplanation(pnet, [0, 1], [2, 2, 3, 4]).explain([])  # a variable twice
    ve = hb.VariableElimination(pnet, [0, 1, 2, 3], [4])
    out = np.zeros(3, np.int8)
    from hbayes_b200._lib import lib

    assert lib().hb_ve_mpe_batch(ve._h, 1, np.full(5, -1, np.int8).ctypes.data, out.ctypes.data, None, 0) != 0  # not an mpe handle
