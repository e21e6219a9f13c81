"""Shared by the parity tests: networks, evidence and bit-comparison helpers."""
import os

import numpy as np

from hbayes_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NETS = ["cancer.net", "sprinkler.net", "asia.net", "asia_rev.net"]


def same_bits(a, b):
    """bit-identical, except that any NaN equals any NaN (zero-probability evidence)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        return False
    nan = np.isnan(a)
    return bool((nan == np.isnan(b)).all() and a[~nan].tobytes() == b[~nan].tobytes())


def max_rel_err(a, b, floor=1e-15):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    ok = ~(np.isnan(a) & np.isnan(b))
    return float((np.abs(a - b)[ok] / np.maximum(np.abs(b)[ok], floor)).max()) if ok.any() else 0.0


def shuffled(net, seed):
    """relabel vertices so that children may precede parents (moralisation quirk A.6 becomes active)"""
    rng = np.random.RandomState(seed)
    perm = rng.permutation(net.n)  # old id -> new id
    inv = np.argsort(perm)
    dims = [net.dims[inv[i]] for i in range(net.n)]
    parents = [[int(perm[p]) for p in net.parents[inv[i]]] for i in range(net.n)]
    values = [net.values[inv[i]] for i in range(net.n)]
    return synth.SynthNet(dims, parents, values)


def random_net(seed, n=None, states=(2, 3), max_parents=2, shuffle=False):
    net = synth.random_dag_net(n or (3 + seed % 8), 9000 + seed, states=states, max_parents=max_parents)
    return shuffled(net, seed) if shuffle else net


def random_evidence_list(rng, dims, max_obs):
    k = int(rng.randint(0, max_obs + 1))
    vs = rng.permutation(len(dims))[:k].tolist()
    return [(int(v), int(rng.randint(dims[v]))) for v in vs]


def oracle_traces(ojt, evidence):
    """{('up'|'down', cluster, sep) | ('posterior', qvar): steps} of the oracle's interpreter for this evidence list"""
    out = {}
    for m in ojt.change_evidence(evidence, trace=True):
        key = ("posterior", m["qvar"]) if m["kind"] == "posterior" else (m["kind"], m["cluster"], m["sep"])
        out[key] = m["steps"]
    return out


def product_traces(prog):
    out = {}
    for g in prog["groups"]:
        key = ("posterior", g["query"][0]) if g["kind"] == "posterior" else (g["kind"], g["cluster"], g["sep"])
        out[key] = g["steps"]
    return out
