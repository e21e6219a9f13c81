"""Seeded synthetic networks and evidence for the BASELINE.json configs.

PRNG = SplitMix64 everywhere (scalar for structure/CPTs, vectorised for evidence rows), so the same
seeds give the same bytes on any box.  Networks are plain arrays in the reference's conventions
(vertex id = index, CPT scope = [child, parents...], first variable slowest -- Bayes/Network.hs:69-79,
Bayes/Factor/PrivateCPT.hs:315-320) and can be written as Hugin `.net` text in the dialect
Bayes/ImportExport/HuginNet.hs accepts.
"""
from __future__ import annotations

import numpy as np

_M = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
        return z ^ (z >> 31)

    def uniform(self):
        return (self.next() >> 11) * (1.0 / (1 << 53))

    def randint(self, lo, hi):
        """inclusive bounds"""
        return lo + self.next() % (hi - lo + 1)


def splitmix64_array(seed, n):
    """n consecutive SplitMix64 outputs as uint64 (identical to n calls of SplitMix64(seed).next())."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & _M) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def proc_hash(seed, idx):
    """stateless SplitMix64 of (seed, idx); idx: uint64 array.  Same bits as hbayes_b200/csrc/procedural.hpp."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M) + (np.asarray(idx, np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def proc_cpt_values(seed, d, npar, idx):
    """entries `idx` of a procedural CPT (d child states, npar parent configurations), bit-identical to the C++/CUDA
    proc_cpt_value: u = 0.05 + 0.95 * uniform(hash), value = u(idx) / (u(0, pa) + u(1, pa) + ...)."""
    idx = np.asarray(idx, np.uint64)
    pa = idx % np.uint64(npar)

    def u(i):
        unif = (proc_hash(seed, i) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return 0.05 + 0.95 * unif

    total = np.zeros(idx.shape, np.float64)
    for k in range(d):
        total = total + u(np.uint64(k) * np.uint64(npar) + pa)
    return u(idx) / total


class SynthNet:
    """dims[v], parents[v] (CPT scope = [v] + parents[v]), values[v] in reference layout.  proc_seed[v] != 0 marks a
    procedural CPT (values[v] empty): see proc_cpt_values."""

    def __init__(self, dims, parents, values, names=None, proc_seed=None):
        self.dims = [int(d) for d in dims]
        self.parents = [list(map(int, p)) for p in parents]
        self.values = [np.asarray(v, np.float64) for v in values]
        self.names = names or ["n%d" % i for i in range(len(dims))]
        self.proc_seed = [int(x) for x in proc_seed] if proc_seed is not None else [0] * len(self.dims)

    def materialise(self):
        """explicit copy: procedural CPTs evaluated on the host (small twins only)"""
        vals = []
        for v in range(self.n):
            if self.proc_seed[v]:
                npar = int(np.prod([self.dims[p] for p in self.parents[v]])) if self.parents[v] else 1
                vals.append(proc_cpt_values(self.proc_seed[v], self.dims[v], npar, np.arange(self.dims[v] * npar, dtype=np.uint64)))
            else:
                vals.append(self.values[v])
        return SynthNet(self.dims, self.parents, vals, self.names)

    @property
    def n(self):
        return len(self.dims)

    @property
    def scopes(self):
        return [[v] + self.parents[v] for v in range(self.n)]

    def hugin_text(self):
        """Hugin text: nodes in id order; `data` laid out parents-slowest / child-fastest (HuginNet.hs:139-147)."""
        out = ["net\n{\n\tname = \"synthetic\";\n}\n"]
        for v in range(self.n):
            st = " ".join('"s%d"' % i for i in range(self.dims[v]))
            out.append("node %s\n{\n\tstates = ( %s );\n}\n" % (self.names[v], st))
        for v in range(self.n):
            sc = self.scopes[v]
            shape = [self.dims[x] for x in sc]
            tab = self.values[v].reshape(shape)
            hug = np.moveaxis(tab, 0, -1).reshape(-1)  # [parents..., child]
            head = "potential ( %s | %s )" % (self.names[v], " ".join(self.names[p] for p in self.parents[v]))
            data = " ".join(repr(float(x)) for x in hug)
            out.append("%s\n{\n\tdata = ( %s );\n}\n" % (head, data))
        return "".join(out)

    def write_hugin(self, path):
        with open(path, "w") as f:
            f.write(self.hugin_text())
        return path


def _random_cpt(rng, dims, v, parents):
    """u ~ U(0.05, 1) per state, normalised per parent configuration; reference layout (child slowest)."""
    npar = 1
    for p in parents:
        npar *= dims[p]
    d = dims[v]
    tab = np.empty((d, npar), np.float64)
    for j in range(npar):
        u = [0.05 + 0.95 * rng.uniform() for _ in range(d)]
        s = sum(u)
        for i in range(d):
            tab[i, j] = u[i] / s
    return tab.reshape(-1)


def random_dag_net(n, seed, states=(2, 4), max_parents=3, window=None):
    """Random DAG, ids topological: node v draws k ~ U{0..max_parents} distinct parents from the
    preceding `window` nodes (all predecessors when window is None)."""
    rng = SplitMix64(seed)
    dims = [rng.randint(states[0], states[1]) for _ in range(n)]
    parents = []
    for v in range(n):
        lo = 0 if window is None else max(0, v - window)
        cand = list(range(lo, v))
        k = min(rng.randint(0, max_parents), len(cand))
        ps = []
        for _ in range(k):
            j = rng.randint(0, len(cand) - 1)
            ps.append(cand.pop(j))
        parents.append(ps)
    values = [_random_cpt(rng, dims, v, parents[v]) for v in range(n)]
    return SynthNet(dims, parents, values)


def grid_net(rows, cols, seed):
    """rows x cols binary grid, node (i,j) has parents (i-1,j),(i,j-1); id = cols*i + j (config C5)."""
    rng = SplitMix64(seed)
    n = rows * cols
    dims = [2] * n
    parents = []
    for i in range(rows):
        for j in range(cols):
            ps = []
            if i > 0:
                ps.append(cols * (i - 1) + j)
            if j > 0:
                ps.append(cols * i + j - 1)
            parents.append(ps)
    values = [_random_cpt(rng, dims, v, parents[v]) for v in range(n)]
    return SynthNet(dims, parents, values)


def ancestral_rows(net: SynthNet, n_rows, seed):
    """n_rows ancestral samples as int8 [n_rows][n] (ids must be topological). Vectorised SplitMix64:
    the uniform for (row r, variable v) is output number r*n + v + 1 of the stream."""
    n = net.n
    z = splitmix64_array(seed, n_rows * n).reshape(n_rows, n)
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    rows = np.zeros((n_rows, n), np.int8)
    for v in range(n):
        ps = net.parents[v]
        assert all(p < v for p in ps), "ancestral_rows needs topological ids"
        d = net.dims[v]
        tab = net.values[v].reshape([d] + [net.dims[p] for p in ps])
        idx = tuple(rows[:, p].astype(np.int64) for p in ps)
        probs = tab[(slice(None),) + idx] if ps else np.repeat(tab[:, None], n_rows, 1)  # [d, n_rows]
        cdf = np.cumsum(probs, axis=0)
        s = (u[:, v][None, :] >= cdf[:-1]).sum(axis=0)
        rows[:, v] = s.astype(np.int8)
    return rows


def choose_observed(n, count, seed):
    """`count` distinct variable ids (sorted)."""
    rng = SplitMix64(seed)
    cand = list(range(n))
    out = []
    for _ in range(count):
        out.append(cand.pop(rng.randint(0, len(cand) - 1)))
    return sorted(out)


def evidence_matrix(net: SynthNet, n_rows, observed, seed):
    """int8 [n_rows][n]: ancestral samples (so P(e) > 0) restricted to `observed`; -1 = unobserved."""
    rows = ancestral_rows(net, n_rows, seed)
    ev = np.full_like(rows, -1)
    ev[:, observed] = rows[:, observed]
    return ev


# ---- named configurations of BASELINE.json ---------------------------------------------------


def config_c2():
    """Alarm-sized random DAG: 37 nodes, 2-4 states, <=3 parents (net seed 37001), 9 observed (seed 37002)."""
    net = random_dag_net(37, 37001, states=(2, 4), max_parents=3)
    observed = choose_observed(37, 9, 37002)
    return net, observed


def config_c2_evidence(net, observed, n_rows):
    return evidence_matrix(net, n_rows, observed, 37003)


def config_c5(rows=20, cols=20, seed=400001):
    return grid_net(rows, cols, seed)


def config_c3():
    """500-node random DAG, 2-4 states, <=3 parents from a window of 4 predecessors; net seed 500003 is the
    first of 500001.. whose maximal cluster under the REFERENCE triangulation (static weightedEdges metric,
    SURVEY A.7) lands in [2^19.5, 2^20.5] state combinations (2^20.2); 125 observed variables (seed 500004)."""
    net = random_dag_net(500, 500003, states=(2, 4), max_parents=3, window=4)
    observed = choose_observed(500, 125, 500004)
    return net, observed


def config_c3_evidence(net, observed, n_rows):
    return evidence_matrix(net, n_rows, observed, 500005)


def config_c5_query(net, metric_order):
    """r = [last node], p = reference minFillOrder minus r, 10 % of the other variables observed (seed 400002)."""
    r = [net.n - 1]
    p = [v for v in metric_order if v not in r]
    observed = choose_observed(net.n - 1, max(1, net.n // 10), 400002)
    return p, r, observed


def config_c4(n_parents=29, seed=4300001):
    """BASELINE config C4: one binary child with `n_parents` binary root parents -> a single cluster whose CPT has
    2^(n_parents+1) entries (2^30 = 8 GiB at 29 parents).  The child's CPT is procedural (nobody materialises it on
    the host); the parents' priors are small stored tables.  ids: parents 0..n_parents-1, child n_parents.
    Split variables = the first three parents (outermost strides of the CPT after the child): 8 slices."""
    rng = SplitMix64(seed)
    n = n_parents + 1
    dims = [2] * n
    parents = [[] for _ in range(n_parents)] + [list(range(n_parents))]
    values = [_random_cpt(rng, dims, v, []) for v in range(n_parents)] + [np.zeros(0)]
    proc_seed = [0] * n_parents + [seed + 17]
    return SynthNet(dims, parents, values, proc_seed=proc_seed), [0, 1, 2]
