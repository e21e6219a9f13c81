"""changeFactor (Bayes/Factor.hs:66-81, Bayes/FactorElimination/JTree.hs:107-115), after the reference's own test
Bayes/Test/CompareEliminations.hs:30-62: a tree / network built with a WRONG `wet grass` table, corrected by factor
replacement, must answer like the right network -- here bit for bit, against the oracle on the right network."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
from helpers import G, random_net, same_bits

pytestmark = pytest.mark.gpu

WINTER, SPRINKLER, WET, RAIN, ROAD, ROADANDRAIN = range(6)  # vertex ids of tests/golden/sprinkler.net


def wrong_sprinkler():
    """exampleWithFactorChange (Bayes/Examples.hs:237-253): cpt wet [sprinkler,rain] ~~ [1,1,1,1,1,1,1,1]"""
    right = hb.importBayesianGraph(os.path.join(G, "sprinkler.net"))
    values = [np.asarray(v, np.float64) for v in right.values]
    the_new_factor = hb.Factor(right.scopes[WET], [2, 2, 2], values[WET].copy())
    assert right.scopes[WET] == [WET, SPRINKLER, RAIN]
    assert the_new_factor.values.tolist() == [1, 0.2, 0.1, 0.05, 0, 0.8, 0.9, 0.95]  # CompareEliminations.hs:34
    values[WET] = np.ones(8)
    return right, hb.BayesianNetwork.from_tables(right.dims, right.scopes, values), the_new_factor


def test_junction_tree_factor_change_like_compare_eliminations():
    right, wrong, f = wrong_sprinkler()
    ojt = OracleJT(OracleNet.load_hugin(os.path.join(G, "sprinkler.net")))
    jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, wrong)
    before = jt.posterior([RAIN]).values.copy()
    assert hb.changeFactor(f, jt) is jt
    for ev in ([], [(WET, 1)], [(WET, 1), (SPRINKLER, 1)]):
        hb.changeEvidence(ev, jt)
        ojt.change_evidence(ev)
        for v in range(6):
            assert same_bits(jt.posterior([v]).values, ojt.posterior([v])[1]), (ev, v)
    assert not same_bits(before, jt.posterior([RAIN]).values)
    # the evidence of the handle survives a factor change (JTree.hs:109: NodeValue v (changeFactor f fa) ev)
    hb.changeEvidence([(WET, 1)], jt)
    g = hb.Factor(f.variables, f.dims, np.array([0.5, 0.3, 0.2, 0.1, 0.5, 0.7, 0.8, 0.9]))
    jt.change_factor(g)
    values = [np.asarray(v, np.float64) for v in right.values]
    values[WET] = g.values
    o2 = OracleJT(OracleNet.from_arrays(right.dims, right.scopes, values))
    o2.change_evidence([(WET, 1)])
    for v in range(6):
        assert same_bits(jt.posterior([v]).values, o2.posterior([v])[1])
    # and the batch path sees the new table too (cached programs re-upload their potentials)
    ev = np.full((5, 6), -1, np.int8)
    ev[:, WET] = [0, 1, 1, 0, 1]
    want = o2.posteriors_batch(ev)
    assert same_bits(jt.posteriors_batch(ev), want)


def test_factor_with_other_variable_order_is_not_a_match():
    """isUsingSameVariablesAs compares the variable LISTS (PrivateCPT.hs:135-136): [wet,rain,sprinkler] replaces nothing"""
    right, wrong, f = wrong_sprinkler()
    jt = hb.JunctionTree(wrong)
    before = [jt.posterior([v]).values.copy() for v in range(6)]
    other = hb.Factor([WET, RAIN, SPRINKLER], [2, 2, 2], f.values)
    assert jt.change_factor(other) is False
    assert all(same_bits(b, jt.posterior([v]).values) for v, b in enumerate(before))
    same = wrong.change_factor(other)
    assert same.values == wrong.values


def test_network_and_ve_factor_change():
    right, wrong, f = wrong_sprinkler()
    onet = OracleNet.load_hugin(os.path.join(G, "sprinkler.net"))
    fixed = hb.changeFactor(f, wrong)  # a new network (Factor.hs:80-81 on the graph functor)
    assert fixed is not wrong and wrong.values[WET] == [1.0] * 8
    for q in range(6):
        p = [v for v in range(6) if v != q]
        assert same_bits(hb.priorMarginal(fixed, p, [q]).values, onet.posterior_marginal(p, [q])[1])
    p, r = [WINTER, SPRINKLER, WET, ROAD, ROADANDRAIN], [RAIN]
    ve = hb.VariableElimination(wrong, p, r)
    ve.posterior([(WET, 1)])  # compile + make the program resident before the change
    assert ve.change_factor(f)
    for ev in ([], [(WET, 1)], [(WET, 1), (SPRINKLER, 1)]):
        assert same_bits(ve.posterior(ev).values, onet.posterior_marginal(p, r, ev)[1])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_network_every_table_replaced(seed):
    """start from one random network, replace every CPT by another network's: same answers as building that network"""
    a = random_net(40 + seed, n=10, states=(2, 3), max_parents=3)
    rng = np.random.RandomState(seed)
    new_values = []
    for v in range(a.n):
        t = rng.uniform(0.05, 1.0, size=(a.dims[v], len(a.values[v]) // a.dims[v]))
        new_values.append((t / t.sum(0)).reshape(-1))
    pa = hb.BayesianNetwork.from_tables(a.dims, a.scopes, a.values)
    jt = hb.JunctionTree(pa)
    observed = [1, 4, 7]
    ev = synth.evidence_matrix(a, 40, observed, seed)
    jt.posteriors_batch(ev)  # programs for both observed sets are resident before the change
    jt.posteriors_batch(np.full((3, a.n), -1, np.int8))
    for v in range(a.n):
        assert jt.change_factor(hb.Factor(a.scopes[v], [a.dims[x] for x in a.scopes[v]], new_values[v]))
    ojt = OracleJT(OracleNet.from_arrays(a.dims, a.scopes, new_values))
    assert same_bits(jt.posteriors_batch(ev), ojt.posteriors_batch(ev))
    assert same_bits(jt.posteriors_batch(np.full((3, a.n), -1, np.int8)), ojt.posteriors_batch(np.full((3, a.n), -1, np.int8)))


def test_soft_evidence_tutorial():
    """Bayes/Examples/Tutorial.hs:319-331 (inferencesWithSoftEvidence): a sensor node (softEvidence, BayesianNetwork.hs:86-96)
    observed True, its table swapped with changeFactor for 90 / 50 / 10 % sensors on one tree"""
    a, sensor = 0, 1
    pnet = hb.BayesianNetwork.from_tables([2, 2], [[a], [sensor, a]], [[0.5, 0.5], [1.0, 0.0, 1.0, 0.0]])
    jt = hb.changeEvidence([(sensor, 1)], hb.createJunctionTree(hb.nodeComparisonForTriangulation, pnet))
    for p in (0.9, 0.5, 0.1):
        se = hb.Factor([sensor, a], [2, 2], np.array([p, 1 - p, 1 - p, p]))  # `se` (BayesianNetwork.hs:99-104)
        got = hb.posterior(hb.changeFactor(se, jt), [a]).values
        ojt = OracleJT(OracleNet.from_arrays([2, 2], [[a], [sensor, a]], [[0.5, 0.5], se.values.tolist()]))
        ojt.change_evidence([(sensor, 1)])
        assert same_bits(got, ojt.posterior([a])[1])
        assert np.allclose(got, [1 - p, p], rtol=0, atol=1e-15)
