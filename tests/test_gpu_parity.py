"""GPU parity tests (run on the B200 box): every result that comes out of libhbayes_cuda through the C ABI is
compared with the oracle (the CPU restatement of the reference) on the same inputs.

Bar (BASELINE.json north_star): marginals within 1e-12 relative (1e-15 absolute floor).  The engine keeps
the reference's association order and never fuses multiply-add, so the tests ask for more: BIT-IDENTICAL
doubles (NaN == NaN for zero-probability evidence)."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
from helpers import G, GOLDEN_NETS, max_rel_err, random_evidence_list, random_net, same_bits

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12  # north_star tolerance; asserted in addition to bit equality


def pair(fname=None, net=None):
    if fname:
        return hb.importBayesianGraph(os.path.join(G, fname)), OracleNet.load_hugin(os.path.join(G, fname))
    return hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values), OracleNet.from_arrays(net.dims, net.scopes, net.values)


def assert_same(got, want, what):
    assert max_rel_err(got, want) <= REL_TOL, what
    assert same_bits(got, want), (what, got, want)


def test_cancer_goldens_through_the_reference_api():
    """Bayes/Test/ReferencePatterns.hs:137-161, written the way the reference's test reads."""
    net = hb.importBayesianGraph(os.path.join(G, "cancer.net"))
    jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, net)
    v = net.ids
    gold = {
        (): {"A": [.2, .8], "B": [.32, .68], "C": [.08, .92], "D": [.32, .68], "E": [.616, .384]},
        (("D", 0),): {"A": [.425, .575], "B": [.8, .2], "C": [.2, .8], "E": [.64, .36]},
        (("D", 1),): {"A": [.0941, .9059], "B": [.0941, .9059], "C": [.0235, .9765], "E": [.6047, .3953]},
    }
    for ev, want in gold.items():
        jt2 = hb.changeEvidence([(v[n], s) for n, s in ev], jt)
        for name, vals in want.items():
            got = hb.posterior(jt2, [v[name]])
            assert np.abs(got.values - np.array(vals)).max() < 1e-4
            others = [x for x in range(net.n) if x != v[name]]
            ve = hb.posteriorMarginal(net, others, [v[name]], [(v[n], s) for n, s in ev])
            assert np.abs(ve.values - np.array(vals)).max() < 1e-4


@pytest.mark.parametrize("fname", GOLDEN_NETS)
def test_single_row_api_bitexact(fname):
    pnet, onet = pair(fname)
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    rng = np.random.RandomState(5)
    for trial in range(8):
        ev = random_evidence_list(rng, onet.dims, 3) if trial else []
        pjt.changeEvidence(ev)
        ojt.change_evidence(ev)
        for v in range(onet.n):
            got = pjt.posterior([v])
            assert got.variables == [v]
            assert_same(got.values, ojt.posterior([v])[1], (fname, ev, v))


def test_sprinkler_ve_equals_jt():
    """Bayes/Test/CompareEliminations.hs:67-84 on the product: VE == JT within the reference's 2e-5."""
    pnet, onet = pair("sprinkler.net")
    jt = hb.JunctionTree(pnet)
    for ev in ([], [(2, 1)], [(2, 1), (1, 1)]):
        jt.changeEvidence(ev)
        for v in range(pnet.n):
            others = [x for x in range(pnet.n) if x != v]
            ve = hb.posteriorMarginal(pnet, others, [v], ev)
            assert np.allclose(jt.posterior([v]).values, ve.values, rtol=2e-5, atol=0)
            assert_same(ve.values, onet.posterior_marginal(others, [v], ev)[1], ("ve", ev, v))


def test_zero_probability_evidence_gives_nan_like_the_reference():
    pnet, onet = pair("sprinkler.net")
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    ev = [(3, 0), (4, 1), (5, 1)]  # roadandrain true while rain false: impossible
    pjt.changeEvidence(ev)
    ojt.change_evidence(ev)
    for v in range(onet.n):
        want = ojt.posterior([v])[1]
        assert np.isnan(want).all()
        assert same_bits(pjt.posterior([v]).values, want)


@pytest.mark.parametrize("seed", range(16))
def test_random_nets_batch_bitexact(seed):
    net = random_net(seed, n=4 + seed % 11, states=(2, 4), max_parents=3, shuffle=seed % 2 == 1)
    pnet, onet = pair(net=net)
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    rng = np.random.RandomState(seed)
    observed = sorted(rng.permutation(net.n)[: rng.randint(0, min(5, net.n))].tolist())
    rows = 70 + seed  # not a multiple of 32: ragged last warp
    ev = np.full((rows, net.n), -1, np.int8)
    for v in observed:
        ev[:, v] = rng.randint(0, net.dims[v], rows)
    got = pjt.posteriors_batch(ev)
    assert_same(got, ojt.posteriors_batch(ev), ("batch", seed))


def test_batch_with_mixed_observed_sets_and_empty_batch():
    pnet, onet = pair("asia.net")
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    rng = np.random.RandomState(2)
    ev = np.full((200, onet.n), -1, np.int8)
    for r in range(200):
        for v in rng.permutation(onet.n)[: rng.randint(0, 4)]:
            ev[r, v] = rng.randint(2)
    assert_same(pjt.posteriors_batch(ev), ojt.posteriors_batch(ev), "mixed")
    assert pjt.posteriors_batch(np.zeros((0, onet.n), np.int8)).shape == (0, sum(onet.dims))
    bad = np.full((3, onet.n), -1, np.int8)
    bad[1, 0] = 7  # state out of range
    with pytest.raises(hb.HbError):
        pjt.posteriors_batch(bad)


def test_multi_variable_posterior_and_nothing():
    pnet, onet = pair("asia.net")
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    clusters = pjt.describe()["clusters"]
    for ev in ([], [(0, 1)], [(7, 0), (1, 1)]):
        pjt.changeEvidence(ev)
        ojt.change_evidence(ev)
        for c in clusters:
            for q in (c[:2], c[::-1][:2], c):
                got, want = pjt.posterior(q), ojt.posterior(q)
                assert got.variables == want[0]
                assert_same(got.values, want[1], (ev, q))
    assert pjt.posterior([0, 7]) is None and ojt.posterior([0, 7]) is None  # A and D share no cluster


def test_device_pointer_batch_and_row_tiling():
    import torch

    net, observed = synth.config_c2()
    pnet, onet = pair(net=net)
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    ev = synth.config_c2_evidence(net, observed, 3000)
    want = ojt.posteriors_batch(ev, n_threads=os.cpu_count() or 1)
    d_ev = torch.from_numpy(ev).cuda()
    got = pjt.posteriors_batch(d_ev).cpu().numpy()
    assert_same(got, want, "c2 device pointers")
    stats = pjt.last_stats()
    assert stats["launches"] > 0 and stats["tiles"] == 1
    # force several row tiles with a small workspace: same bits
    pjt.set_workspace(stats["row_entries"] * 8 * 512 + 4096 * 64)
    got2 = pjt.posteriors_batch(ev)
    assert pjt.last_stats()["tiles"] > 1
    assert_same(got2, want, "c2 tiled")


@pytest.mark.parametrize("shape", [(3, 3), (5, 6), (8, 8)])
def test_ve_grid_batch_bitexact(shape):
    g = synth.grid_net(shape[0], shape[1], 21)
    pnet, onet = pair(net=g)
    r = [g.n - 1]
    p = [v for v in hb.minFillOrder(pnet) if v not in r]
    ve = hb.VariableElimination(pnet, p, r)
    rng = np.random.RandomState(3)
    observed = sorted(rng.permutation(g.n - 1)[: max(1, g.n // 10)].tolist())
    ev = synth.evidence_matrix(g, 100, observed, 77)
    got = ve.posterior_batch(ev)
    assert_same(got, onet.posterior_marginal_batch(p, r, ev), ("ve grid", shape))
    prior = hb.priorMarginal(pnet, p, r)
    assert_same(prior.values, onet.posterior_marginal(p, r, [])[1], "prior")


def test_properties_at_scale():
    """Size-independent checks on a batch too large for the oracle: rows sum to one per variable, observed
    variables are one-hot at their state, and permuting the rows permutes the output."""
    net, observed = synth.config_c2()
    pnet, _ = pair(net=net)
    pjt = hb.JunctionTree(pnet)
    rows = 1 << 16
    ev = synth.config_c2_evidence(net, observed, rows)
    out = pjt.posteriors_batch(ev)
    off = np.concatenate([[0], np.cumsum(net.dims)])
    for v in range(net.n):
        blk = out[:, off[v]:off[v + 1]]
        assert np.abs(blk.sum(1) - 1).max() < 1e-12
        if v in observed:
            assert (blk[np.arange(rows), ev[:, v]] > 0.999999).all()
    perm = np.random.RandomState(0).permutation(rows)
    assert same_bits(pjt.posteriors_batch(ev[perm]), out[perm])


def test_asia_goldens_through_the_c_abi():
    """Bayes/Test/ReferencePatterns.hs:119-131 (asia priors, state 0 = yes), abs 1e-4 as the reference asserts"""
    gold = {"A": .01, "S": .5, "T": .0104, "L": .055, "B": .45, "E": .0648, "X": .1103, "D": .4360}
    for fname in ("asia.net", "asia_rev.net"):
        net = hb.importBayesianGraph(os.path.join(G, fname))
        jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, net)
        for name, want in gold.items():
            got = hb.posterior(jt, [net.ids[name]]).values
            assert abs(got[0] - want) < 1e-4 and abs(got[1] - (1 - want)) < 1e-4


@pytest.mark.parametrize("seed", range(6))
def test_ve_equals_jt_on_synthetic_networks(seed):
    """the reference's VE == JT cross-check (Test/CompareEliminations.hs:67-84, rel 2e-5) on random networks,
    both engines on the GPU, at the north-star tolerance"""
    net = random_net(40 + seed, n=8 + 2 * seed, states=(2, 4), max_parents=3)
    pnet, _ = pair(net=net)
    jt = hb.JunctionTree(pnet)
    rng = np.random.RandomState(seed)
    observed = sorted(rng.permutation(net.n)[:3].tolist())
    ev = synth.evidence_matrix(net, 64, observed, 9 + seed)
    post = jt.posteriors_batch(ev)
    off = np.concatenate([[0], np.cumsum(net.dims)])
    order = hb.minFillOrder(pnet)
    for v in range(0, net.n, 3):
        p = [x for x in order if x != v] + [x for x in range(net.n) if x not in order and x != v]
        ve = hb.VariableElimination(pnet, p, [v]).posterior_batch(ev)
        assert max_rel_err(ve, post[:, off[v]:off[v + 1]]) <= 1e-12


def test_disconnected_and_single_variable_networks():
    """empty separators make the messages structural Scalars (the `ps` path of _factorProduct), isolated
    observed variables make whole results the constant 1.0"""
    nets = [
        synth.SynthNet([3], [[]], [[0.2, 0.5, 0.3]]),
        synth.SynthNet([2, 3, 2, 2], [[], [], [1], []], [[0.3, 0.7], [0.2, 0.5, 0.3], [0.1, 0.4, 0.6, 0.9, 0.6, 0.4], [0.5, 0.5]]),
    ]
    for net in nets:
        pnet, onet = pair(net=net)
        pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
        assert pjt.describe() == ojt.describe()
        rng = np.random.RandomState(4)
        for observed in ([], [0], list(range(net.n))[-1:], list(range(net.n))):
            ev = np.full((37, net.n), -1, np.int8)
            for v in observed:
                ev[:, v] = rng.randint(0, net.dims[v], 37)
            assert_same(pjt.posteriors_batch(ev), ojt.posteriors_batch(ev), (net.n, observed))


def test_saved_tree_gives_the_same_posteriors(tmp_path):
    pnet, onet = pair("asia.net")
    jt = hb.JunctionTree(pnet)
    jt.save(tmp_path / "asia.hbjt")
    back = hb.JunctionTree.load(tmp_path / "asia.hbjt", pnet)
    ev = np.full((50, pnet.n), -1, np.int8)
    ev[:, 7] = np.arange(50) % 2
    assert same_bits(back.posteriors_batch(ev), jt.posteriors_batch(ev))
    assert same_bits(back.posterior([3]).values, OracleJT(onet).posterior([3])[1])


def _star(n_children, root_dim, child_dim, seed):
    """one root with many children: the root's bucket holds more tables than the templated kernels take (generic path)"""
    rng = synth.SplitMix64(seed)
    dims = [root_dim] + [child_dim] * n_children
    parents = [[]] + [[0]] * n_children
    values = [synth._random_cpt(rng, dims, v, parents[v]) for v in range(len(dims))]
    return synth.SynthNet(dims, parents, values)


@pytest.mark.parametrize("net", [_star(9, 3, 2, 1), _star(12, 5, 7, 2), synth.random_dag_net(14, 777, states=(5, 7), max_parents=2)],
                         ids=["star9", "star12_dims5x7", "dag_dims5to7"])
def test_many_factor_buckets_and_wide_variables(net):
    """buckets with > 6 tables (prodsum_generic_kernel) and variables with 5-7 states (odd blocking pairs)"""
    pnet, onet = pair(net=net)
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    assert pjt.describe() == ojt.describe()
    rng = np.random.RandomState(6)
    for observed in ([], [1, 3], [0], [2, 5, 6]):
        ev = np.full((130, net.n), -1, np.int8)
        for v in observed:
            ev[:, v] = rng.randint(0, net.dims[v], 130)
        assert_same(pjt.posteriors_batch(ev), ojt.posteriors_batch(ev), observed)
    others = list(range(1, net.n))
    ve = hb.VariableElimination(pnet, others, [0])
    ev = np.full((33, net.n), -1, np.int8)
    ev[:, 2] = rng.randint(0, net.dims[2], 33)
    assert_same(ve.posterior_batch(ev), onet.posterior_marginal_batch(others, [0], ev), "ve star")


@pytest.mark.parametrize("seed", range(20, 44))
def test_random_stress(seed):
    """more random networks, denser evidence, shuffled ids, ragged row counts"""
    net = random_net(seed, n=5 + seed % 17, states=(2, 2 + seed % 4), max_parents=1 + seed % 4, shuffle=seed % 3 == 0)
    pnet, onet = pair(net=net)
    pjt, ojt = hb.JunctionTree(pnet, cmp=seed % 2), OracleJT(onet, cmp=["weighted", "added"][seed % 2])
    rng = np.random.RandomState(seed)
    observed = sorted(rng.permutation(net.n)[: rng.randint(0, net.n)].tolist())
    rows = [1, 2, 3, 7, 63, 65, 200][seed % 7]
    ev = np.full((rows, net.n), -1, np.int8)
    for v in observed:
        ev[:, v] = rng.randint(0, net.dims[v], rows)
    assert_same(pjt.posteriors_batch(ev), ojt.posteriors_batch(ev), ("stress", seed, observed, rows))


def test_plain_c_example_runs(tmp_path):
    import subprocess
    from hbayes_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cancer_query")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", os.path.join(root, "examples", "cancer_query.c"), "-o", exe, "-L" + libdir, "-lhbayes_cuda",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe, os.path.join(G, "cancer.net")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "var 2: 0.425000 0.575000" in r.stdout  # P(A | D = Present), Bayes/Test/ReferencePatterns.hs:146
