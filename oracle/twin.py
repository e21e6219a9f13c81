"""Pure-Python twin of the hbayes hot path -- TEST INFRASTRUCTURE ONLY.

This file is part of the parity ORACLE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
(hbayes_b200/) never does.

It restates, function by function and list operation by list operation, the
reference's Haskell for the path BASELINE.json:north_star names.  It is written
for fidelity, not speed (pure Python loops; IEEE-754 binary64 `*` and `+`, which
is what GHC emits for Double -- no FMA, no reassociation).  The C++ oracle
(oracle/hbayes_oracle.cpp) is the fast restatement; the two are written
independently and integer/bit-compared in tests/test_oracle.py.

Reference files followed (all under /root/reference/):
  Bayes/PrivateTypes.hs:118-128,202,256-268     list-backed Set, DV Ord, indicesForDomain
  Bayes/Factor/PrivateCPT.hs:164-327            norm, scale, product, projectOut, strides
  Bayes/Factor/CPT.hs:44-51,109-141             changeVariableOrder, instance Factor CPT
  Bayes/Factor.hs:90-94,107-108,171-188         factorFromInstantiation, factorDivide
  Bayes/VariableElimination/Buckets.hs          buckets (all)
  Bayes/VariableElimination.hs:56-73,116-236    marginal, prior/posteriorMarginal, orders
  Bayes/FactorElimination/JTree.hs:203-233,288-369,398-491   tree, collect, distribute, evidence, newMessage
  Bayes/FactorElimination.hs:71-111,130-213,223-338,384-427  triangulation, tree build, posterior, moral graph
  Bayes.hs:495-682                               SimpleGraph iteration orders
  Bayes/ImportExport/HuginNet.hs                 .net dialect

Parity pin: tests/test_oracle.py checks this twin against the reference's own
golden vectors (Bayes/Test/ReferencePatterns.hs:119-161 cancer + asia), the
VE == JT test (Bayes/Test/CompareEliminations.hs:67-84) and brute-force joint
enumeration.  No GHC exists in this image, so values beyond the reference's
4-decimal goldens are pinned by this restatement only.
"""
from __future__ import annotations

import itertools
import re
from functools import reduce

# --------------------------------------------------------------------------
# Haskell list helpers (Data.List semantics)
# --------------------------------------------------------------------------


def nub(xs):
    out = []
    for x in xs:
        if x not in out:
            out.append(x)
    return out


def l_union(xs, ys):
    """Data.List.union: xs ++ foldl (flip delete) (nub ys) xs."""
    rest = nub(ys)
    for x in xs:
        if x in rest:
            rest.remove(x)  # delete = first occurrence
    return list(xs) + rest


def l_diff(xs, ys):
    """Data.List.(\\\\): remove the first occurrence of each y."""
    out = list(xs)
    for y in ys:
        if y in out:
            out.remove(y)
    return out


def foldr1(op, xs):
    acc = xs[-1]
    for x in reversed(xs[:-1]):
        acc = op(x, acc)
    return acc


def minimum_by(key, xs):
    """minimumBy (compare `on` key): FIRST minimal element."""
    best = xs[0]
    for x in xs[1:]:
        if key(x) < key(best):
            best = x
    return best


def maximum_by(key, xs):
    """maximumBy (compare `on` key): LAST maximal element."""
    best = xs[0]
    for x in xs[1:]:
        if not (key(best) > key(x)):
            best = x
    return best


# --------------------------------------------------------------------------
# Factors.  DV = (vertex, dim) -- tuple order == derived Ord (PrivateTypes.hs:202)
# --------------------------------------------------------------------------


class Factor:
    """PrivateCPT (PrivateCPT.hs:67-72): Table vars values | Scalar value."""

    __slots__ = ("vars", "vals", "scalar")

    def __init__(self, vars_, vals, scalar=False):
        self.vars = list(vars_)
        self.vals = vals
        self.scalar = scalar

    @staticmethod
    def scalar_(v):
        return Factor([], [v], True)

    def is_scalar(self):
        return self.scalar

    def contains(self, dv):
        return (not self.scalar) and any(v[0] == dv[0] for v in self.vars)

    def __repr__(self):
        return ("Scalar(%r)" % self.vals[0]) if self.scalar else "Table(%r,%r)" % (self.vars, self.vals)


def create_cpt_with_dims(dims, values):
    """createCPTWithDims (PrivateCPT.hs:270-281)."""
    if not dims:
        return Factor.scalar_(values[0])
    p = 1
    for d in dims:
        p *= d[1]
    assert len(values) == p, "incoherent factor size"
    return Factor(dims, list(values))


def index_strides(dvars):
    """indexStrides (PrivateCPT.hs:315-320): scanr (*) 1 (tail dims)."""
    dims = [d[1] for d in dvars]
    out = [1]
    for d in reversed(dims[1:]):
        out.insert(0, out[0] * d)
    return out


def index_strides_for(dr, di):
    """indexStridesFor (PrivateCPT.hs:302-310): strides of dr read with an index over di."""
    ref = dict(zip(dr, index_strides(dr))) if dr else {}
    return [ref.get(dv, 0) for dv in di]


def indices_for_domain(dvars):
    """indicesForDomain (PrivateTypes.hs:256-260): lexicographic, first var slowest."""
    return itertools.product(*[range(d[1]) for d in dvars])


def index_position(strides, idx):
    return sum(s * i for s, i in zip(strides, idx))


def factor_product(l):
    """_factorProduct (Op 1.0 (*)) (PrivateCPT.hs:215-238; CPT.hs:138)."""
    if not l:
        return Factor.scalar_(1.0)
    naked = foldr1(l_union, [f.vars for f in l])
    scalars = [f for f in l if f.is_scalar()]
    cpts = [f for f in l if not f.is_scalar()]
    mul = lambda a, b: a * b
    ps = foldr1(mul, [f.vals[0] for f in scalars]) if scalars else 1.0
    if not cpts:
        return Factor.scalar_(ps)
    strides = [index_strides_for(f.vars, naked) for f in cpts]
    values = []
    for x in indices_for_domain(naked):
        vs = [f.vals[index_position(s, x)] for s, f in zip(strides, cpts)]
        values.append(ps * foldr1(mul, vs))
    return create_cpt_with_dims(naked, values)


def factor_project_out(s, f):
    """factorProjectOut (CPT.hs:140-141) -> cptFactorProjectOutWith (sum . map fst)."""
    if f.is_scalar():
        return f
    d = f.vars
    keep = l_diff(d, s)
    strides = index_strides_for(d, keep + list(s))
    values = []
    for ki in indices_for_domain(keep):
        acc = 0.0  # sum = foldl (+) 0
        for si in indices_for_domain(s):
            acc = acc + f.vals[index_position(strides, ki + si)]
        values.append(acc)
    return create_cpt_with_dims(keep, values)


def factor_norm(f):
    """_factorNorm (PrivateCPT.hs:167-173): left fold from 0 in layout order."""
    acc = 0.0
    if f.is_scalar():
        return f.vals[0]
    for v in f.vals:
        acc = acc + v
    return acc


def factor_scale(s, f):
    """_factorScale (PrivateCPT.hs:180-187): scale s x = s * x."""
    if f.is_scalar():
        return Factor.scalar_(s * f.vals[0])
    return Factor(f.vars, [s * v for v in f.vals])


def factor_divide(f, d):
    """factorDivide (Factor.hs:171-172): scale by 1.0 / d."""
    if d == 0.0:  # Haskell: 1.0 / 0 = Infinity (then 0 * Infinity = NaN, PrivateCPT.hs:180-187)
        inv = float("-inf") if str(d).startswith("-") else float("inf")
    else:
        inv = 1.0 / d
    return factor_scale(inv, f)


def factor_from_instantiation(dv, a):
    """factorFromInstantiation (Factor.hs:90-94)."""
    return create_cpt_with_dims([dv], [1.0 if i == a else 0.0 for i in range(dv[1])])


def factor_value(f, inst):
    """privateFactorValue (PrivateCPT.hs:189-197); inst = [(dv, value)]."""
    if f.is_scalar():
        return f.vals[0]
    strides = index_strides_for(f.vars, [dv for dv, _ in inst])
    return f.vals[index_position(strides, [v for _, v in inst])]


def change_variable_order(old_order, new_order, old_values):
    """changeVariableOrder (CPT.hs:44-51)."""
    old = create_cpt_with_dims(old_order, old_values)
    out = []
    for idx in indices_for_domain(new_order):
        out.append(factor_value(old, list(zip(new_order, idx))))
    return out


# --------------------------------------------------------------------------
# Buckets + marginal (Buckets.hs, VariableElimination.hs:56-73)
# --------------------------------------------------------------------------


class Buckets:
    def __init__(self, order, m):
        self.e = list(order)
        self.m = dict(m)


def create_buckets(s, e, r):
    """createBuckets (Buckets.hs:53-65)."""
    order = list(e) + list(r)
    rf = list(s)
    m = {}
    for dv in order:
        fk = [f for f in rf if f.contains(dv)]
        rf = [f for f in rf if not f.contains(dv)]
        m[dv] = fk  # M.insert
    return Buckets(order, m)


def add_bucket(b, f):
    """addBucket (Buckets.hs:92-98): insertWith' (++) bucket [f] -> prepended."""
    for dv in b.e:
        if f.contains(dv):
            m = dict(b.m)
            m[dv] = [f] + m.get(dv, [])
            return Buckets(b.e, m)
    return b


def update_bucket(dv, f, b):
    """updateBucket (Buckets.hs:74-89)."""
    if f.is_scalar():
        m = dict(b.m)
        m[dv] = [f]
        return Buckets(b.e[1:], m)
    m = dict(b.m)
    m.pop(dv, None)
    return add_bucket(Buckets(b.e[1:], m), f)


def marginalize_one_variable(b, dv, trace=None):
    """marginalizeOneVariable (Buckets.hs:105-111)."""
    fk = b.m[dv]
    p = factor_product(fk)
    f2 = factor_project_out([dv], p)
    if trace is not None:
        trace.append(("elim", dv, [list(f.vars) for f in fk], list(p.vars), list(f2.vars)))
    return update_bucket(dv, f2, b)


def marginal(lf, p, r, assignment, trace=None):
    """marginal (VariableElimination.hs:56-73); assignment = [(dv, value)]."""
    b = create_buckets(lf, p, r)
    for dv, a in assignment:
        b = add_bucket(b, factor_from_instantiation(dv, a))
    for dv in p:
        b = marginalize_one_variable(b, dv, trace)
    rest = []
    for k in sorted(b.m.keys()):  # M.elems: ascending DV
        rest.extend(b.m[k])
    if trace is not None:
        trace.append(("final", [list(f.vars) for f in rest]))
    return factor_product(rest)


# --------------------------------------------------------------------------
# Max-product: MAXCPT and mpe (MaxCPT.hs:25-70; VariableElimination.hs:79-114).
# An element is (value, possible instantiations); an instantiation is a list of (dv, state).
# --------------------------------------------------------------------------


def max_multiply(a, b):
    """instance FactorElement (Double,PossibleInstantiations): multiply (MaxCPT.hs:29-31)."""
    (da, la), (db, lb) = a, b
    if not la:
        return (da * db, lb)
    if not lb:
        return (da * db, la)
    return (da * db, [xa + xb for xa in la for xb in lb])


def convert_to_max_factor(f):
    """convertToMaxFactor (PrivateCPT.hs:89-94)."""
    return Factor(f.vars, [(v, []) for v in f.vals], f.scalar)


def max_factor_product(l):
    """factorProduct of MAXCPT = _factorProduct (Op (Scalar 1.0) (1.0,[]) multiply) (MaxCPT.hs:68)."""
    if not l:
        return Factor.scalar_((1.0, []))
    return _factor_op((1.0, []), max_multiply, l)


def maximization(l):
    """maximization (MaxCPT.hs:38-47); l = [((value, possible), newInst)].  maximumBy keeps the LAST maximal element
    (foldl1: the left one survives only when compare says GT)."""
    if not l:
        return (1.0, [])
    best = l[0]
    for y in l[1:]:
        x = best
        gt = not (x[0][0] < y[0][0]) and not (x[0][0] == y[0][0])  # compare on Double
        best = x if gt else y
    (d, possible), new_inst = best
    if not new_inst:
        return (d, possible)
    if not possible:
        return (d, [new_inst])
    return (d, [new_inst + i for i in possible])


def max_project_out(s, f):
    """factorProjectOut of MAXCPT (MaxCPT.hs:70-71) = cptFactorProjectOutWith maximization (PrivateCPT.hs:243-266)."""
    if f.is_scalar():
        return f
    d = f.vars
    keep = l_diff(d, s)
    strides = index_strides_for(d, keep + list(s))
    values = []
    for ki in indices_for_domain(keep):
        l = []
        for si in indices_for_domain(s):
            l.append((f.vals[index_position(strides, ki + si)], list(zip(s, si))))
        values.append(maximization(l))
    return create_cpt_with_dims(keep, values)


def mpe_marginal(lf, p, r, assignment):
    """mpemarginal (VariableElimination.hs:79-97): sum over p, then maximise over r."""
    b = create_buckets(lf, p, r)
    for dv, a in assignment:
        b = add_bucket(b, factor_from_instantiation(dv, a))
    for dv in p:
        b = marginalize_one_variable(b, dv)
    b = Buckets(b.e, {k: [convert_to_max_factor(f) for f in v] for k, v in b.m.items()})
    for dv in r:  # marginalizeOneVariable on MAXCPT items (Buckets.hs:105-111)
        f2 = max_project_out([dv], max_factor_product(b.m[dv]))
        b = update_bucket(dv, f2, b)
    rest = []
    for k in sorted(b.m.keys()):
        rest.extend(b.m[k])
    return max_factor_product(rest)


def mpe(net, p, r, assignment):
    """mpe (VariableElimination.hs:101-114).  p, r = vertex lists, assignment = [(vertex, state)].
    Returns (value P(mpe, e), instantiations [[(vertex, state)]]) -- the reference returns the instantiations."""
    res = mpe_marginal(list(net.cpts), [net.dv(v) for v in p], [net.dv(v) for v in r], [(net.dv(v), s) for v, s in assignment])
    assert res.is_scalar(), "The final MAXCPT factor should be a scalar one"
    value, insts = res.vals[0]
    return value, [[(dv[0], st) for dv, st in inst] for inst in insts]


# --------------------------------------------------------------------------
# Bayesian network container (what runBN returns: DirectedSG () CPT)
# --------------------------------------------------------------------------


class Network:
    """Vertex ids 0..n-1; cpts[v] has scope [v, parents in `ingoing` order] (Network.hs:69-79)."""

    def __init__(self, dims, parents, cpts, names=None):
        self.dims = list(dims)
        self.parents = [list(p) for p in parents]
        self.cpts = cpts
        self.names = names or ["n%d" % i for i in range(len(dims))]

    @property
    def n(self):
        return len(self.dims)

    def dv(self, v):
        return (v, self.dims[v])

    @staticmethod
    def from_tables(dims, parents, values, names=None):
        cpts = []
        for v in range(len(dims)):
            scope = [(v, dims[v])] + [(p, dims[p]) for p in parents[v]]
            cpts.append(create_cpt_with_dims(scope, values[v]))
        return Network(dims, parents, cpts, names)


# --------------------------------------------------------------------------
# Hugin .net loader (HuginNet.hs; dialect: SURVEY Appendix C)
# --------------------------------------------------------------------------


def _filter_comment(text):
    """filterComment (HuginNet.hs:180-186): '%' to end of line, newline kept."""
    return re.sub(r"%[^\n]*", "", text)


def load_hugin(path):
    text = _filter_comment(open(path).read())
    nodes = []  # (name, dim) in file order -> vertex ids (HuginNet.hs:149-151,172-177)
    pots = []
    pos = 0
    sec = re.compile(r"\s*(net|node\s+([A-Za-z0-9_-]+)|potential\s*([^\n]*))\n\{\n([^}]*)\}", re.S)
    while True:
        m = sec.match(text, pos)
        if not m:
            break
        pos = m.end()
        head, nname, pconds, body = m.group(1), m.group(2), m.group(3), m.group(4)
        if head.startswith("node"):
            states = []
            for line in body.split("\n"):
                sm = re.match(r"\s*states\s*=\s*\((.*)\)\s*;", line)
                if sm:
                    states += re.findall(r'"([^"]+)"', sm.group(1))
            nodes.append((nname, len(states)))
        elif head.startswith("potential"):
            conds = [t for t in re.split(r"[()| ]", pconds) if t]  # splitCPT
            dm = re.search(r"data\s*=\s*([^;]*)", body)
            vals = [t for t in re.split(r"[() \n\t]", dm.group(1)) if t]  # splitValues
            pots.append((conds, [float(v) for v in vals]))
    if text[pos:].strip():
        raise ValueError("unparsed trailing text in %s" % path)
    ids = {name: i for i, (name, _) in enumerate(nodes)}
    dims = [d for _, d in nodes]
    parents = [[] for _ in nodes]
    values = [None] * len(nodes)
    for conds, vals in pots:
        dvs = [(ids[c], dims[ids[c]]) for c in conds]
        dst, cnd = dvs[0], dvs[1:]
        newvalues = change_variable_order(cnd + [dst], dvs, vals)  # addSection :139-147
        parents[dst[0]] = [c[0] for c in cnd]
        values[dst[0]] = newvalues
    return Network.from_tables(dims, parents, values, [n for n, _ in nodes])


# --------------------------------------------------------------------------
# Moral graph / triangulation / cluster tree (FactorElimination.hs)
# --------------------------------------------------------------------------


def moral_graph(net):
    """moralGraph (FactorElimination.hs:384-427) incl. the accumulator quirk.

    Returns the undirected adjacency as dict v -> set(neighbours)."""
    edges = set()  # directed (s, e)
    ingoing = {v: [] for v in range(net.n)}
    for v in range(net.n):
        for p in reversed(net.parents[v]):  # cpt adds edges in reverse; ingoing prepends
            if (p, v) not in edges:
                edges.add((p, v))
                ingoing[v].insert(0, p)
    for v in range(net.n):  # foldlWithVertex' : ascending id, on the accumulator graph
        ps = list(ingoing[v])
        snapshot = set(edges)
        linked = lambda x, y: (x, y) in snapshot or (y, x) in snapshot
        for x in ps:
            for y in ps:
                if x != y and not linked(x, y):
                    if (x, y) not in edges:
                        edges.add((x, y))
                        ingoing[y].insert(0, x)
    adj = {v: set() for v in range(net.n)}
    for (s, e) in edges:
        adj[s].add(e)
        adj[e].add(s)
    return adj


def weighted_edges(net, adj, v):
    """weightedEdges (FactorElimination.hs:80-95): ordered pairs of non-adjacent neighbours."""
    w = lambda x: len(net.cpts[x].vals) if not net.cpts[x].is_scalar() else 1
    nodes = sorted(adj[v])
    tot = 0
    for x in nodes:
        for y in nodes:
            if x != y and y not in adj[x]:
                tot += w(x) * w(y)
    return tot


def number_of_added_edges(net, adj, v):
    """numberOfAddedEdges (FactorElimination.hs:71-78)."""
    nodes = sorted(adj[v])
    return sum(1 for x in nodes for y in nodes if x != y and y not in adj[x])


def triangulate(metric, adj0):
    """triangulate (FactorElimination.hs:130-143): static metric closed over the moral graph."""
    adj = {v: set(n) for v, n in adj0.items()}
    key = {v: metric(v) for v in adj0}
    clusters = []
    order = []
    while adj:
        v = minimum_by(lambda x: key[x], sorted(adj.keys()))
        nb = adj[v]
        for x in nb:
            for y in nb:
                if x != y:
                    adj[x].add(y)
        for x in nb:
            adj[x].discard(v)
        clusters.append(frozenset([v]) | frozenset(nb))
        order.append(v)
        del adj[v]
    return keep_maximal_clusters(clusters), order


def keep_maximal_clusters(l):
    """keepMaximalClusters (FactorElimination.hs:162-186)."""
    prefix = []
    for cur in l:
        hit = None
        for i, s in enumerate(prefix):
            if cur <= s:
                hit = i
                break
        if hit is None:
            prefix.append(cur)
        else:
            r = prefix.pop(hit)
            prefix.append(r)
    return prefix


class JTree:
    """JTree (JTree.hs:89-105) keyed by cluster = tuple of ascending DVs."""

    def __init__(self):
        self.root = None
        self.children = {}  # cluster -> [sep] newest first
        self.parent = {}  # cluster -> sep
        self.sep_parent = {}
        self.sep_child = {}
        self.sep_cluster = {}
        self.node_vertex = {}
        self.factors = {}  # cluster -> [Factor]
        self.evidence = {}  # cluster -> [Factor]
        self.sep_value = {}  # sep -> None | [up, down]
        self.key = 0

    def tree_nodes(self):
        return sorted(self.factors.keys())  # Map.keys: ascending Cluster

    def leaves(self):
        return [c for c in self.tree_nodes() if not self.children.get(c)]


def create_uninitialized_jt(net, cmp="weighted", order=None):
    adj = moral_graph(net)
    if cmp == "weighted":
        metric = lambda v: weighted_edges(net, adj, v)
    elif cmp == "added":
        metric = lambda v: number_of_added_edges(net, adj, v)
    else:
        raise ValueError(cmp)
    if order is not None:
        rank = {v: i for i, v in enumerate(order)}
        metric = lambda v: rank[v]
    vclusters, elim = triangulate(metric, adj)
    clusters = [tuple(sorted((v, net.dims[v]) for v in c)) for c in vclusters]
    # maximumSpanningTree (FactorElimination.hs:253-276)
    t = JTree()
    t.root = clusters[0]
    t.factors[clusters[0]] = []
    t.evidence[clusters[0]] = []
    t.node_vertex[clusters[0]] = 0
    remaining = list(range(1, len(clusters)))
    while remaining:
        poss = []
        for rv in remaining:
            for lv in t.tree_nodes():
                ev = len(vclusters[rv] & vclusters[t.node_vertex[lv]])
                poss.append((rv, lv, ev))
        rf, lf, _ = maximum_by(lambda p: p[2], poss)
        remaining = [x for x in remaining if x != rf]
        found = clusters[rf]
        sepc = tuple(x for x in found if x in lf)  # Set.intersection, ascending
        t.factors[found] = []
        t.evidence[found] = []
        t.node_vertex[found] = rf
        t.key += 1
        ns = t.key
        t.children[lf] = [ns] + t.children.get(lf, [])
        t.sep_child[ns] = found
        t.sep_value[ns] = None
        t.sep_cluster[ns] = sepc
        t.parent[found] = ns
        t.sep_parent[ns] = lf
    t.elim_order = elim
    t.cluster_list = clusters
    return t


def _cluster_size(c):
    p = 1
    for dv in c:
        p *= dv[1]
    return p


def update_tree_values(t, change, factors):
    """updateTreeValues (JTree.hs:424-440)."""
    for f in factors:
        cover = [c for c in t.tree_nodes() if all(v in c for v in f.vars)]
        mc = minimum_by(_cluster_size, cover)
        change(t, mc, f)


def new_message(inputs, f, e, sepc, trace=None):
    """newMessage (JTree.hs:485-491)."""
    allf = f + e + inputs
    keep = list(sepc)
    remove = l_diff(nub([v for x in allf for v in x.vars]), keep)
    return marginal(allf, remove, keep, [], trace)


def collect(t, trace=None):
    """collect (JTree.hs:329-348)."""
    nodes = t.leaves()
    while nodes:
        for h in nodes:  # updateUpSeparator (JTree.hs:288-305)
            seps = t.children.get(h, [])
            if all(t.sep_value[s] is not None for s in seps):
                p = t.parent.get(h)
                if p is not None:
                    incoming = [t.sep_value[s][0] for s in seps]
                    tr = None
                    if trace is not None:
                        tr = []
                        trace.append(("up", h, p, tr))
                    msg = new_message(incoming, t.factors[h], t.evidence[h], t.sep_cluster[p], tr)
                    old = t.sep_value[p]
                    t.sep_value[p] = [msg, old[1] if old else None]
        parents = [t.parent[h] for h in nodes if h in t.parent]
        nodes = sorted(set(t.sep_parent[s] for s in parents))
    return t


def distribute(t, trace=None):
    """distribute (JTree.hs:350-369): pre-order, children newest-first."""

    def go(node):
        ch = t.children.get(node, [])
        for child in ch:  # updateDownSeparator (JTree.hs:307-322)
            below = [t.sep_value[s][0] for s in l_diff(ch, [child])]
            p = t.parent.get(node)
            above = t.sep_value[p][1] if p is not None else None
            incoming = ([above] if above is not None else []) + below
            tr = None
            if trace is not None:
                tr = []
                trace.append(("down", node, child, tr))
            msg = new_message(incoming, t.factors[node], t.evidence[node], t.sep_cluster[child], tr)
            t.sep_value[child][1] = msg
        for child in ch:
            go(t.sep_child[child])

    go(t.root)
    return t


def create_junction_tree(net, cmp="weighted", order=None):
    """createJunctionTree (FactorElimination.hs:298-307)."""
    t = create_uninitialized_jt(net, cmp, order)

    def change(t, c, f):  # setFactors (JTree.hs:443-451): f : oldf
        t.factors[c] = [f] + t.factors[c]

    update_tree_values(t, change, net.cpts)
    return distribute(collect(t))


def change_evidence(t, evidence, trace=None):
    """changeEvidence (JTree.hs:454-466); evidence = [(dv, state)] in the caller's order."""
    for s in t.sep_value:
        t.sep_value[s] = None
    for c in t.evidence:
        t.evidence[c] = []

    def change(t, c, f):
        t.evidence[c] = [f] + t.evidence[c]

    update_tree_values(t, change, [factor_from_instantiation(dv, a) for dv, a in evidence])
    return distribute(collect(t, trace), trace)


def find_cluster_for(t, v):
    """traverseTree (findClusterFor v) (JTree.hs:398-415; FactorElimination.hs:330-338)."""
    state = [None]

    def go(cur):
        if all(x in cur for x in v):
            state[0] = cur  # Stop (Just c): do not descend, siblings still visited
            return
        for s in t.children.get(cur, []):
            go(t.sep_child[s])

    go(t.root)
    return state[0]


def posterior(t, v, trace=None):
    """posterior (FactorElimination.hs:314-327); v = [dv]. Returns None or normalised Factor."""
    c = find_cluster_for(t, v)
    if c is None:
        return None
    p = t.parent.get(c)
    d = t.sep_value[p][1] if p is not None else None
    if d is None:
        d = Factor.scalar_(1.0)
    u = [t.sep_value[s][0] for s in t.children.get(c, [])]
    allf = [d] + u + t.factors[c] + t.evidence[c]
    remove = l_diff(nub([x for f in allf for x in f.vars]), v)
    un = marginal(allf, remove, v, [], trace)
    return factor_divide(un, factor_norm(un))


# --------------------------------------------------------------------------
# changeFactor (Factor.hs:66-81; JTree.hs:107-115) and learnEM (EM.hs:16-66; CPT.hs:144-163)
# --------------------------------------------------------------------------


def change_factor_in_list(f, l):
    """changeFactorInFunctor (Factor.hs:66-71): isUsingSameVariablesAs is LIST equality (PrivateCPT.hs:135-136)."""
    return [f if ((not cf.is_scalar()) and cf.vars == f.vars) else cf for cf in l]


def change_factor_jt(f, t):
    """instance FactorContainer (JTree Cluster) (JTree.hs:107-115): replace in every node, keep the evidence,
    clear the separators, collect + distribute."""
    for c in t.factors:
        t.factors[c] = change_factor_in_list(f, t.factors[c])
    for s in t.sep_value:
        t.sep_value[s] = None
    return distribute(collect(t))


def _factor_op(unit, op, l):
    """_factorProduct (Op _ unit op) (PrivateCPT.hs:215-238) for an arbitrary element operator."""
    naked = foldr1(l_union, [f.vars for f in l])
    scalars = [f for f in l if f.is_scalar()]
    cpts = [f for f in l if not f.is_scalar()]
    ps = foldr1(op, [f.vals[0] for f in scalars]) if scalars else unit
    if not cpts:
        return Factor.scalar_(ps)
    strides = [index_strides_for(f.vars, naked) for f in cpts]
    values = []
    for x in indices_for_domain(naked):
        values.append(op(ps, foldr1(op, [f.vals[index_position(st, x)] for st, f in zip(strides, cpts)])))
    return create_cpt_with_dims(naked, values)


def _divide(a, b):
    """instance FactorElement Double (CPT.hs:113-114): divide _ 0 = 0; divide a b = a / b."""
    if b == 0:
        return 0.0
    return a / b


def cpt_sum(l):
    """cptSum (CPT.hs:162-163)."""
    return _factor_op(0.0, lambda a, b: a + b, l)


def cpt_divide(a, b):
    """cptDivide (CPT.hs:144-160)."""
    if a.is_scalar() and b.is_scalar():
        return Factor.scalar_(_divide(a.vals[0], b.vals[0]))
    if a.is_scalar():
        return factor_scale(a.vals[0], b)
    if b.is_scalar():
        vb = b.vals[0]
        return Factor.scalar_(0.0) if vb == 0.0 else factor_scale(1.0 / vb, a)
    return _factor_op(1.0, _divide, [b, a])


def learn_em(net, samples, posterior_fn=None):
    """learnEM (EM.hs:36-66).  samples: one evidence list [(vertex, state)] per sample.  posterior_fn(evidence, [vertex])
    -> Factor gives `posterior (changeEvidence evidence jt) vars`; default: this twin's own junction tree.
    Returns one Factor per vertex -- with the variable order cptDivide produces."""
    if posterior_fn is None:
        t = create_junction_tree(net)

        def posterior_fn(evidence, vs, _state={"ev": None}):
            if _state["ev"] is not evidence:
                change_evidence(t, [(net.dv(v), s) for v, s in evidence])
                _state["ev"] = evidence
            return posterior(t, [net.dv(v) for v in vs])

    def compute_sample(sample):  # computeSample / computeNode (EM.hs:47-66)
        out = []
        for v in range(net.n):
            f = [dv[0] for dv in net.cpts[v].vars]
            a = posterior_fn(sample, f)
            b = Factor.scalar_(1.0) if len(f) == 1 else posterior_fn(sample, f[1:])
            out.append((a, b))
        return out

    acc = None
    for sample in samples:  # foldl1 sumG (EM.hs:18-27,44)
        cur = compute_sample(sample)
        acc = cur if acc is None else [(cpt_sum([xa, ya]), cpt_sum([xb, yb])) for (xa, xb), (ya, yb) in zip(acc, cur)]
    return [cpt_divide(a, b) for a, b in acc]  # divideG (EM.hs:29-32,45)


# --------------------------------------------------------------------------
# Variable elimination entry points + orders (VariableElimination.hs:116-236)
# --------------------------------------------------------------------------


def posterior_marginal(net, p, r, assignment, trace=None):
    res = marginal(list(net.cpts), p, r, assignment, trace)
    return factor_divide(res, factor_norm(res))


def prior_marginal(net, ea, eb):
    return posterior_marginal(net, ea, eb, [])


def interaction_graph(net):
    """interactionGraph (VariableElimination.hs:150-164): vertices only from factors with >= 2 vars."""
    adj = {}
    for f in net.cpts:
        vs = [v[0] for v in f.vars]
        for x in vs:
            for y in vs:
                if x != y:
                    adj.setdefault(x, set()).add(y)
                    adj.setdefault(y, set()).add(x)
    return adj


def elimination_order_for_metric(net, metric):
    """eliminationOrderForMetric (VariableElimination.hs:210-224): no fill-in edges added."""
    adj = {v: set(n) for v, n in interaction_graph(net).items()}
    l = sorted(adj.keys())
    out = []
    while l:
        best = minimum_by(lambda v: metric(adj, v), l)
        for x in adj[best]:
            adj[x].discard(best)
        del adj[best]
        l.remove(best)
        out.append((best, net.dims[best]))
    return out


def degree_order(net, order):
    """degreeOrder (VariableElimination.hs:186-207): max neighbour count at elimination time, fill-in edges added."""
    adj = {v: set(n) for v, n in interaction_graph(net).items()}
    w = 0
    for dv in order:
        v = dv[0] if isinstance(dv, tuple) else dv
        r = list(adj.get(v, ()))
        w = max(w, len(r))
        for x in r:
            for y in r:
                if x != y:
                    adj[x].add(y)
            adj[x].discard(v)
        adj.pop(v, None)
    return w


def min_degree_order(net):
    return elimination_order_for_metric(net, lambda adj, v: len(adj[v]))


def min_fill_order(net):
    def missing(adj, v):
        r = adj[v]
        return sum(1 for x in r for y in r if x != y and y not in adj[x])

    return elimination_order_for_metric(net, missing)


# --------------------------------------------------------------------------
# Independent check: brute-force joint enumeration (NOT the reference algorithm)
# --------------------------------------------------------------------------


def brute_force_posterior(net, q, evidence):
    """P(q | evidence) by enumerating the joint; evidence = {vertex: state}."""
    acc = [0.0] * net.dims[q]
    for idx in itertools.product(*[range(d) for d in net.dims]):
        if any(idx[v] != s for v, s in evidence.items()):
            continue
        p = 1.0
        for f in net.cpts:
            p *= factor_value(f, [(dv, idx[dv[0]]) for dv in f.vars])
        acc[idx[q]] += p
    z = sum(acc)
    return [a / z for a in acc]
