"""Pins the parity oracle (oracle/hbayes_oracle.cpp + oracle/twin.py) to the reference's own known answers.

  * golden vectors of Bayes/Test/ReferencePatterns.hs:119-161 (cancer priors/posteriors, asia priors; abs tol 1e-4
    as in :62-65)
  * VE == JT on the sprinkler `example` (Bayes/Test/CompareEliminations.hs:67-84)
  * the structure hand-traced in SURVEY Appendix B
  * C++ oracle == independent pure-Python twin, bit for bit (values) and integer for integer (structure, op traces)
  * brute-force joint enumeration on random networks, incl. children-before-parents ids (moralisation quirk A.6)
"""
import os

import numpy as np
import pytest

import oracle
from oracle import OracleJT, OracleNet, twin
from hbayes_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def jt_all(jt, n):
    return [jt.posterior([v])[1] for v in range(n)]


def test_cancer_structure_appendix_b():
    net = OracleNet.load_hugin(os.path.join(G, "cancer.net"))
    assert net.names == ["D", "E", "A", "C", "B"]
    assert net.scopes == [[0, 3, 4], [1, 3], [2], [3, 2], [4, 2]]
    assert net.values[0] == [.8, .8, .8, .05, .2, .2, .2, .95]
    d = OracleJT(net).describe()
    assert d["elim_order"] == [0, 1, 2, 4, 3]
    assert d["clusters"] == [[2, 3, 4], [0, 3, 4], [1, 3]]
    assert d["sep_vars"] == [[3, 4], [3]] and d["sep_child"] == [1, 2]
    assert d["children"][0] == [2, 1]
    assert d["factors"] == [[4, 3, 2], [0], [1]]


# Bayes/Test/ReferencePatterns.hs:137-161 ; order of variables here: D E A C B
CANCER_GOLD = {
    (): {"A": [.2, .8], "B": [.32, .68], "C": [.08, .92], "D": [.32, .68], "E": [.616, .384]},
    (("D", 0),): {"A": [.425, .575], "B": [.8, .2], "C": [.2, .8], "E": [.64, .36]},
    (("D", 1),): {"A": [.0941, .9059], "B": [.0941, .9059], "C": [.0235, .9765], "E": [.6047, .3953]},
}


@pytest.mark.parametrize("ev", list(CANCER_GOLD))
def test_cancer_goldens(ev):
    net = OracleNet.load_hugin(os.path.join(G, "cancer.net"))
    ids = {n: i for i, n in enumerate(net.names)}
    jt = OracleJT(net)
    jt.change_evidence([(ids[n], s) for n, s in ev])
    for name, want in CANCER_GOLD[ev].items():
        got = jt.posterior([ids[name]])[1]
        assert np.abs(got - np.array(want)).max() < 1e-4, (name, got)
        # VE gives the same
        others = [v for v in range(net.n) if v != ids[name]]
        _, ve = net.posterior_marginal(others, [ids[name]], [(ids[n], s) for n, s in ev])
        assert np.abs(ve - np.array(want)).max() < 1e-4


# Bayes/Test/ReferencePatterns.hs:119-131 (asia priors, state 0 = yes)
ASIA_GOLD = {"A": .01, "S": .5, "T": .0104, "L": .055, "B": .45, "E": .0648, "X": .1103, "D": .4360}


@pytest.mark.parametrize("fname", ["asia.net", "asia_rev.net"])
def test_asia_goldens(fname):
    net = OracleNet.load_hugin(os.path.join(G, fname))
    ids = {n: i for i, n in enumerate(net.names)}
    jt = OracleJT(net)
    for name, want in ASIA_GOLD.items():
        got = jt.posterior([ids[name]])[1]
        assert abs(got[0] - want) < 1e-4 and abs(got[1] - (1 - want)) < 1e-4, (name, got)


def test_sprinkler_ve_equals_jt():
    """Bayes/Test/CompareEliminations.hs:67-84: priors of all variables, rain | wet, rain | wet, sprinkler."""
    net = OracleNet.load_hugin(os.path.join(G, "sprinkler.net"))
    jt = OracleJT(net)
    d = jt.describe()
    assert d["elim_order"] == [0, 2, 4, 5, 1, 3]
    assert d["clusters"] == [[3, 4, 5], [0, 1, 3], [1, 2, 3]]
    for evid in ([], [(2, 1)], [(2, 1), (1, 1)]):
        jt.change_evidence(evid)
        for v in range(net.n):
            others = [x for x in range(net.n) if x != v]
            _, ve = net.posterior_marginal(others, [v], evid)
            got = jt.posterior([v])[1]
            assert np.allclose(got, ve, rtol=2e-5, atol=0), (v, got, ve)
    jt.change_evidence([])
    assert np.allclose(jt.posterior([3])[1], [.48, .52], atol=1e-12)


def same_bits(a, b):
    """bit-identical, except that any NaN equals any NaN (zero-probability evidence)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    nan = np.isnan(a)
    return a.shape == b.shape and (nan == np.isnan(b)).all() and a[~nan].tobytes() == b[~nan].tobytes()


def twin_net(onet):
    return twin.Network.from_tables(onet.dims, [s[1:] for s in onet.scopes], onet.values, onet.names)


def twin_describe(t):
    nodes = sorted(t.factors.keys(), key=lambda c: t.node_vertex[c])
    return [[dv[0] for dv in c] for c in nodes]


NETS = ["cancer.net", "sprinkler.net", "asia.net", "asia_rev.net"]


@pytest.mark.parametrize("fname", NETS)
def test_cpp_equals_twin_bitwise(fname):
    onet = OracleNet.load_hugin(os.path.join(G, fname))
    tnet = twin.load_hugin(os.path.join(G, fname))
    assert [f.vals for f in tnet.cpts] == onet.values
    ojt = OracleJT(onet)
    tjt = twin.create_junction_tree(tnet)
    assert twin_describe(tjt) == ojt.describe()["clusters"]
    rng = np.random.RandomState(7)
    for trial in range(6):
        k = rng.randint(0, min(4, onet.n))
        vs = rng.permutation(onet.n)[:k].tolist()  # caller order is NOT sorted on purpose
        ev = [(int(v), int(rng.randint(onet.dims[v]))) for v in vs]
        ojt.change_evidence(ev)
        twin.change_evidence(tjt, [(tnet.dv(v), s) for v, s in ev])
        for v in range(onet.n):
            a = ojt.posterior([v])[1]
            b = twin.posterior(tjt, [tnet.dv(v)]).vals
            assert same_bits(a, b), (fname, ev, v, a, b)
            others = [x for x in range(onet.n) if x != v]
            _, c = onet.posterior_marginal(others, [v], ev)
            d = twin.posterior_marginal(tnet, [tnet.dv(x) for x in others], [tnet.dv(v)],
                                        [(tnet.dv(x), s) for x, s in ev]).vals
            assert same_bits(c, d)


def random_small_net(seed, shuffle):
    net = synth.random_dag_net(3 + seed % 6, 9000 + seed, states=(2, 3), max_parents=2)
    if shuffle:  # relabel so children may precede parents
        rng = np.random.RandomState(seed)
        perm = rng.permutation(net.n)  # old id -> new id
        inv = np.argsort(perm)
        dims = [net.dims[inv[i]] for i in range(net.n)]
        parents = [[int(perm[p]) for p in net.parents[inv[i]]] for i in range(net.n)]
        values = [net.values[inv[i]] for i in range(net.n)]
        net = synth.SynthNet(dims, parents, values)
    return net


@pytest.mark.parametrize("seed", range(24))
def test_random_nets_vs_bruteforce_and_twin(seed):
    net = random_small_net(seed, shuffle=seed % 2 == 1)
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    tnet = twin.Network.from_tables(net.dims, net.parents, [v.tolist() for v in net.values])
    ojt = OracleJT(onet)
    tjt = twin.create_junction_tree(tnet)
    assert twin_describe(tjt) == ojt.describe()["clusters"]
    rng = np.random.RandomState(seed)
    k = rng.randint(0, 3)
    vs = rng.permutation(net.n)[:k].tolist()
    ev = [(int(v), int(rng.randint(net.dims[v]))) for v in vs]
    ojt.change_evidence(ev)
    twin.change_evidence(tjt, [(tnet.dv(v), s) for v, s in ev])
    for v in range(net.n):
        a = ojt.posterior([v])[1]
        b = np.array(twin.posterior(tjt, [tnet.dv(v)]).vals)
        bf = np.array(twin.brute_force_posterior(tnet, v, dict(ev)))
        assert a.tobytes() == b.tobytes()
        assert np.abs(a - bf).max() < 1e-13
        others = [x for x in range(net.n) if x != v]
        _, c = onet.posterior_marginal(others, [v], ev)
        assert np.abs(c - bf).max() < 1e-13


def test_trace_cancer_d0_matches_appendix_b():
    """The op trace the product's compiled schedule must reproduce (SURVEY Appendix B worked schedule)."""
    net = OracleNet.load_hugin(os.path.join(G, "cancer.net"))
    jt = OracleJT(net)
    tr = jt.change_evidence([(0, 0)], trace=True)
    kinds = [(m["kind"], m["cluster"], m["sep"]) for m in tr if m["kind"] != "posterior"]
    assert kinds == [("up", 1, 1), ("up", 2, 2), ("down", 0, 2), ("down", 0, 1)]
    up1 = tr[0]["steps"]
    assert up1[0]["var"] == 0 and up1[0]["prod"] == [0, 3, 4] and up1[0]["out"] == [3, 4]
    down_c = tr[2]["steps"]  # down {A,C,B} -> sep {C}: eliminate B then A
    assert [s["var"] for s in down_c] == [4, 2, -1]
    assert down_c[0]["prod"] == [4, 2, 3]
    down_cb = tr[3]["steps"]  # message layout follows rule 5: scope [B,C]
    assert down_cb[-1]["out"] == [4, 3]


def test_ve_orders_grid():
    """reference minFill semantics (no fill-in) on a grid starts 0, last, 1, cols, ..."""
    g = synth.grid_net(4, 4, 1)
    onet = OracleNet.from_arrays(g.dims, g.scopes, g.values)
    tnet = twin.Network.from_tables(g.dims, g.parents, [v.tolist() for v in g.values])
    for metric, tw in (("min_fill", twin.min_fill_order), ("min_degree", twin.min_degree_order)):
        assert onet.ve_order(metric) == [dv[0] for dv in tw(tnet)]


def test_factor_algebra_properties():
    """The six QuickCheck properties of Bayes/Factor/CPT.hs:68-98 on the oracle's product / projectOut."""
    rng = np.random.RandomState(3)
    for _ in range(50):
        def rand_factor(lo):
            k = rng.randint(1, 4)
            ids = (lo + rng.permutation(10)[:k]).tolist()
            sc = [(int(i), int(rng.randint(2, 4))) for i in ids]
            return sc, rng.rand(int(np.prod([d for _, d in sc])))
        fa, fb = rand_factor(0), rand_factor(20)
        # product then project the other factor's variables recovers factor * norm(other)
        sc, vals = oracle.factor_product([fa, fb])
        dims = dict(fa[0] + fb[0])
        sc2, out = oracle.factor_project_out([(v, dims[v]) for v in sc], vals, [v for v, _ in fb[0]])
        assert sc2 == [v for v, _ in fa[0]]
        assert np.allclose(out, fa[1] * fb[1].sum(), rtol=1e-12)
        # projection commutes
        full = [(v, dims[v]) for v in sc]
        a1 = oracle.factor_project_out(full, vals, [sc[0]])
        a2 = oracle.factor_project_out([(v, dims[v]) for v in a1[0]], a1[1], [sc[-1]])
        b1 = oracle.factor_project_out(full, vals, [sc[-1]])
        b2 = oracle.factor_project_out([(v, dims[v]) for v in b1[0]], b1[1], [sc[0]])
        assert a2[0] == b2[0] and np.allclose(a2[1], b2[1], rtol=1e-12)
        # project all = norm
        _, allout = oracle.factor_project_out(full, vals, sc)
        assert np.allclose(allout, vals.sum(), rtol=1e-12)


# ---- `next` rows: changeFactor and learnEM restated (oracle/twin.py) ------------------------------


def test_twin_change_factor_like_compare_eliminations():
    """Bayes/Test/CompareEliminations.hs:30-62: a tree built on the network with a wrong `wet` table, corrected with
    changeFactor, answers like the tree of the right network; a factor with another variable ORDER replaces nothing."""
    right = twin.load_hugin(os.path.join(G, "sprinkler.net"))
    wet = 2
    cpts = list(right.cpts)
    cpts[wet] = twin.Factor(right.cpts[wet].vars, [1.0] * 8)
    wrong = twin.Network(right.dims, right.parents, cpts, right.names)
    f = right.cpts[wet]
    assert [dv[0] for dv in f.vars] == [2, 1, 3] and f.vals == [1, 0.2, 0.1, 0.05, 0, 0.8, 0.9, 0.95]
    t = twin.create_junction_tree(wrong)
    ref = twin.create_junction_tree(right)
    swapped = twin.Factor([f.vars[0], f.vars[2], f.vars[1]], f.vals)
    before = [twin.posterior(t, [right.dv(v)]).vals for v in range(6)]
    twin.change_factor_jt(swapped, t)
    assert [twin.posterior(t, [right.dv(v)]).vals for v in range(6)] == before
    twin.change_factor_jt(f, t)
    for ev in ([], [(2, 1)], [(2, 1), (1, 1)]):
        twin.change_evidence(t, [(right.dv(v), s) for v, s in ev])
        twin.change_evidence(ref, [(right.dv(v), s) for v, s in ev])
        for v in range(6):
            assert twin.posterior(t, [right.dv(v)]).vals == twin.posterior(ref, [right.dv(v)]).vals
    assert twin.posterior(t, [right.dv(3)]).vals != before[3]


def test_cpt_divide_and_sum_follow_the_reference_operators():
    """Bayes/Factor/CPT.hs:113-114,144-163: divide _ 0 = 0; tables divide as 1.0 `divide` (b `divide` a) over the
    variable order b ++ (a minus b); a scalar denominator scales by 1.0 / b"""
    a = twin.Factor([(0, 2), (1, 2)], [0.3, 0.0, 0.1, 0.6])  # [child, parent]
    b = twin.Factor([(1, 2)], [0.4, 0.6])
    q = twin.cpt_divide(a, b)
    assert q.vars == [(1, 2), (0, 2)]
    assert q.vals == [1.0 / (0.4 / 0.3), 1.0 / (0.4 / 0.1), 0.0, 1.0 / (0.6 / 0.6)]
    assert twin.cpt_divide(a, twin.Factor.scalar_(4.0)).vals == [(1.0 / 4.0) * x for x in a.vals]
    assert twin.cpt_divide(a, twin.Factor.scalar_(0.0)).is_scalar()
    s = twin.cpt_sum([a, twin.Factor([(1, 2), (0, 2)], [1.0, 2.0, 3.0, 4.0])])
    assert s.vars == a.vars and s.vals == [0.0 + (0.3 + 1.0), 0.0 + (0.0 + 3.0), 0.0 + (0.1 + 2.0), 0.0 + (0.6 + 4.0)]
    assert twin.cpt_sum([twin.Factor.scalar_(2.0), twin.Factor.scalar_(1.0)]).vals == [3.0]


@pytest.mark.parametrize("seed", [3, 6])
def test_learn_em_twin_equals_cpp_oracle_posteriors(seed):
    """the EM fold driven by the pure-Python junction tree and by the C++ oracle's: same bits, same layouts"""
    net = synth.random_dag_net(6, 9000 + seed, states=(2, 3), max_parents=2)
    samples = synth.evidence_matrix(net, 10, [0, 2, 3], 5)
    samples[4, 2] = -1  # one sample with another observed set
    got = oracle.learn_em(net.dims, net.scopes, net.values, samples)
    tnet = twin.Network.from_tables(net.dims, net.parents, [list(map(float, v)) for v in net.values])
    want = twin.learn_em(tnet, [[(v, int(s)) for v, s in enumerate(r) if s >= 0] for r in samples])
    for v, ((sc, vals), f) in enumerate(zip(got, want)):
        assert sc == [dv[0] for dv in f.vars] and vals.tolist() == f.vals
        assert sc[-1] == v and sorted(sc) == sorted(net.scopes[v])  # parents first, child last
        col = vals.reshape(-1, net.dims[v]).sum(1)
        assert np.all((np.abs(col - 1) < 1e-12) | (col == 0))


# ---- `next` row f4: max-product (oracle/twin.py: mpe) -----------------------------------------------

YES, NO = 0, 1


def test_twin_mpe_reproduces_the_reference_goldens_on_asia():
    """Bayes/Test/ReferencePatterns.hs:75-84 (compareMpeAsia): the MPE and a MAP on asia.net under X = Yes, D = No,
    including the ORDER of the instantiation list the reference compares against."""
    net = twin.load_hugin(os.path.join(G, "asia.net"))
    i = {n: k for k, n in enumerate(net.names)}
    x, b, d, a, s, l, t, e = [i[c] for c in "XBDASLTE"]
    named = lambda m: [[(net.names[v], st) for v, st in inst] for inst in m]
    value, m = twin.mpe(net, [x, d], [b, a, s, l, t, e], [(x, YES), (d, NO)])
    assert named(m) == [[("E", NO), ("T", NO), ("L", NO), ("S", NO), ("B", NO), ("A", NO)]]
    value2, m2 = twin.mpe(net, [x, d, b, l, t, e], [a, s], [(x, YES), (d, NO)])
    assert named(m2) == [[("S", YES), ("A", NO)]]  # "a MAP is not always the projection of a MPE"
    # brute force: the MPE value is the largest joint entry consistent with the evidence
    import itertools

    def joint(st):
        return float(np.prod([twin.factor_value(net.cpts[v], [(dv, st[dv[0]]) for dv in net.cpts[v].vars]) for v in range(net.n)]))

    consistent = [st for st in itertools.product(range(2), repeat=net.n) if st[x] == YES and st[d] == NO]
    best = max((joint(st), st) for st in consistent)
    assert abs(best[0] - value) < 1e-15 and all(best[1][v] == stt for v, stt in m[0])


def test_twin_mpe_ties_go_to_the_last_maximal_state():
    """maximumBy keeps the last maximal element (MaxCPT.hs:41): a uniform variable is explained by its LAST state"""
    net = twin.Network.from_tables([3, 2], [[], [0]], [[1 / 3, 1 / 3, 1 / 3], [0.5, 0.5, 0.5, 0.5, 0.5, 0.5]])
    value, m = twin.mpe(net, [], [0, 1], [])
    assert m == [[(1, 1), (0, 2)]] and value == (1 / 3) * 0.5


@pytest.mark.parametrize("seed", range(8))
def test_twin_mpe_and_map_against_brute_force(seed):
    """max-product restatement vs exhaustive enumeration on random networks: the MPE value is the largest joint entry
    consistent with the evidence and the returned instantiation attains it; a MAP (some variables summed out) maximises
    the summed joint"""
    import itertools

    rng = np.random.RandomState(seed)
    net = synth.random_dag_net(4 + seed % 4, 9100 + seed, states=(2, 3), max_parents=2)
    tnet = twin.Network.from_tables(net.dims, net.parents, [list(map(float, v)) for v in net.values])

    def joint(st):
        p = 1.0
        for v in range(net.n):
            sc = net.scopes[v]
            p *= float(np.asarray(net.values[v]).reshape([net.dims[x] for x in sc])[tuple(st[x] for x in sc)])
        return p

    k = int(rng.randint(0, 2))
    observed = [int(v) for v in rng.permutation(net.n)[:k]]
    ev = [(v, int(rng.randint(net.dims[v]))) for v in observed]
    others = [int(v) for v in rng.permutation(net.n) if v not in observed]
    n_sum = int(rng.randint(0, len(others)))
    p, r = observed + others[:n_sum], others[n_sum:]
    value, inst = twin.mpe(tnet, p, r, ev)
    got = dict(inst[0])
    assert sorted(got) == sorted(r)
    best = -1.0
    for states in itertools.product(*[range(net.dims[v]) for v in r]):
        fixed = dict(ev)
        fixed.update(zip(r, states))
        total = 0.0
        free = [v for v in p if v not in observed]
        for rest in itertools.product(*[range(net.dims[v]) for v in free]):
            st = dict(fixed)
            st.update(zip(free, rest))
            total += joint(st)
        if dict(zip(r, states)) == got:
            mine = total
        best = max(best, total)
    assert abs(value - best) <= 1e-12 * best and abs(mine - best) <= 1e-12 * best


def test_learn_em_on_complete_data_is_counting():
    """with every variable observed the family posteriors are indicators, so learnEM (Bayes/EM.hs:36-45) must return the
    empirical conditional frequencies -- in the reference's [parents..., child] layout, zero where a parent
    configuration never occurs (divide _ 0 = 0)"""
    net = synth.random_dag_net(5, 9300, states=(2, 3), max_parents=2)
    samples = synth.evidence_matrix(net, 40, list(range(net.n)), 77)
    learnt = oracle.learn_em(net.dims, net.scopes, net.values, samples)
    for v, (sc, vals) in enumerate(learnt):
        assert sc[-1] == v and sorted(sc[:-1]) == sorted(net.scopes[v][1:])
        table = vals.reshape([net.dims[x] for x in sc])
        for idx in np.ndindex(*table.shape):
            par = np.all([samples[:, x] == s for x, s in zip(sc[:-1], idx[:-1])], axis=0) if len(sc) > 1 else np.ones(len(samples), bool)
            both = par & (samples[:, v] == idx[-1])
            want = 0.0 if both.sum() == 0 else both.sum() / par.sum()
            assert abs(table[idx] - want) <= 4e-16 * max(want, 1.0), (v, idx, table[idx], want)
