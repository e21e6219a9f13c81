"""Whole-schedule on-chip kernel hb_tree (csrc/jit.cpp): changeEvidence + collect + distribute + every
posterior of a tile of evidence rows in ONE launch, per-row tables in shared memory.  Forced for every program that
qualifies (HB_JIT_MIN_WORK=1, HB_JIT_SYNC=1: read once per process -> a subprocess) and compared bit for bit with the oracle
and with the templated kernels (HB_JIT=0) on the same inputs: junction trees with evidence on random networks incl. shuffled
ids, family queries (wide results), variable elimination, zero-probability evidence (NaN rows), ragged batches (1..67 rows),
the single-row stateful API, a split handle (unnormalised results) and C2 itself."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
from helpers import G, random_net, same_bits

out = {}
on_chip = 0
for seed, rows in ((5, 257), (9, 64), (12, 1000), (21, 33), (30, 1)):
    net = random_net(seed, n=12 + seed %% 5, states=(2, 4), max_parents=3, shuffle=seed %% 2 == 1)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    rng = np.random.RandomState(seed)
    observed = sorted(rng.permutation(net.n)[:3].tolist())
    ev = np.full((rows, net.n), -1, np.int8)
    for v in observed:
        ev[:, v] = rng.randint(0, net.dims[v], size=rows)
    jt = hb.JunctionTree(pnet)
    got = jt.posteriors_batch(ev)
    on_chip += jt.kernel_info()["on_chip"]
    out["jt%%d" %% seed] = got
    assert same_bits(got, OracleJT(onet).posteriors_batch(ev)), ("jt", seed)
    fam = jt.queries_batch([net.scopes[v] for v in range(net.n)], ev)
    on_chip += jt.kernel_info()["on_chip"]
    out["fam%%d" %% seed] = np.concatenate([f[2] for f in fam], axis=1)
    order = hb.minFillOrder(pnet)
    r = [order[-1]]
    p = [v for v in order if v not in r]
    if not set(r) & set(observed):
        ve = hb.VariableElimination(pnet, p, r).posterior_batch(ev)
        out["ve%%d" %% seed] = ve
        assert same_bits(ve, onet.posterior_marginal_batch(p, r, ev)), ("ve", seed)
    # every batch length around the tile height: partial last tiles, single rows
    for m in (1, 2, 3, 15, 16, 17, 31, 32, 33, 67):
        if m <= rows:
            assert same_bits(jt.posteriors_batch(ev[:m]), got[:m]), ("ragged", seed, m)
# golden networks through the stateful single-row API, incl. zero-probability evidence (NaN like the reference)
for name in ("cancer.net", "asia.net", "sprinkler.net"):
    pnet = hb.importBayesianGraph(G + "/" + name)
    onet = OracleNet.load_hugin(G + "/" + name)
    jt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    for evidence in ([], [(0, 0)], [(1, 1), (0, 0)]):
        jt.changeEvidence(evidence); ojt.change_evidence(evidence)
        for v in range(pnet.n):
            a, b = jt.posterior([v]).values, ojt.posterior([v])[1]
            assert same_bits(a, b), (name, evidence, v, a, b)
        on_chip += jt.kernel_info()["on_chip"]
net = synth.SynthNet([2, 2], [[], [0]], [np.array([1.0, 0.0]), np.array([0.5, 0.5, 0.5, 0.5])])
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
nan = hb.JunctionTree(pnet).posteriors_batch(np.array([[1, -1], [0, -1]], np.int8))
assert np.isnan(nan[0]).all() and not np.isnan(nan[1]).any()
assert same_bits(nan, OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(np.array([[1, -1], [0, -1]], np.int8)))
out["nan"] = nan
# split handle: results stay unnormalised, the sum over the slices normalised == the unsplit engine (1e-12)
net = random_net(41, n=9, states=(2, 3), max_parents=3)
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
ev = np.full((40, net.n), -1, np.int8)
ev[:, 2] = np.arange(40) %% net.dims[2]
whole = hb.JunctionTree(pnet).posteriors_batch(ev)
total = np.zeros_like(whole)
for s in range(net.dims[5]):
    part = hb.JunctionTree(pnet, split=[(5, s)])
    e = ev.copy(); e[:, 5] = s
    total += part.posteriors_batch(e)
    on_chip += part.kernel_info()["on_chip"]
last = part.normalise_batch(total)
cols = np.r_[0, np.cumsum(net.dims)]
keep = [c for v in range(net.n) if v != 5 for c in range(cols[v], cols[v + 1])]
assert np.abs(last[:, keep] - whole[:, keep]).max() < 1e-12
out["split"] = last
# C2 itself: a prefix against the oracle
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
ev = synth.config_c2_evidence(net, observed, 3000)
jt = hb.JunctionTree(pnet)
n = jt.specialise(observed)
got = jt.posteriors_batch(ev)
info = jt.kernel_info()
on_chip += info["on_chip"]
assert same_bits(got[:512], OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[:512], n_threads=8))
out["c2"] = got
print("on_chip programs ->", on_chip, "c2 info", info, "specialise ->", n)
np.savez(sys.argv[1], **out)
'''


def run(tmp_path, name, env):
    path = str(tmp_path / (name + ".npz"))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, path], capture_output=True, text=True, env=e, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return path, r.stderr + r.stdout


def test_on_chip_kernel_bit_identical_to_oracle_and_to_the_templated_kernels(tmp_path):
    import numpy as np

    chip, log = run(tmp_path, "chip", {"HB_JIT": "1", "HB_JIT_MIN_WORK": "1", "HB_JIT_SYNC": "1", "HB_JIT_VERBOSE": "1",
                                       "HB_CACHE_DIR": str(tmp_path / "cache")})
    assert "on-chip kernel hb_tree" in log, log[-2000:]  # really built and used
    assert int(log.rsplit("on_chip programs ->", 1)[1].split()[0]) >= 15
    assert "'on_chip': 1" in log.rsplit("c2 info", 1)[1]
    plain, plain_log = run(tmp_path, "plain", {"HB_JIT": "0"})
    assert int(plain_log.rsplit("on_chip programs ->", 1)[1].split()[0]) == 0
    a, b = np.load(chip), np.load(plain)
    assert sorted(a.files) == sorted(b.files) and len(a.files) >= 12
    for k in a.files:
        assert a[k].tobytes() == b[k].tobytes(), k


def test_contract_mode_agrees_to_1e_12_and_is_not_the_default():
    """HB_PRECISION_CONTRACT (north star: marginals within 1e-12 relative, 1e-15 absolute floor): the generated kernels fuse
    multiply-add; the default stays bit-exact."""
    import numpy as np

    import hbayes_b200 as hb
    from hbayes_b200 import synth
    from helpers import max_rel_err

    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    ev = synth.config_c2_evidence(net, observed, 20000)
    exact = hb.JunctionTree(pnet).posteriors_batch(ev)  # HB_JIT=0 in this process: precompiled kernels
    os.environ["HB_JIT"] = "1"
    try:
        jt = hb.JunctionTree(pnet)
        assert jt.specialise(observed) > 0
        assert jt.posteriors_batch(ev).tobytes() == exact.tobytes()  # default mode: exact
        jt.set_precision("contract")
        assert jt.specialise(observed) > 0
        fast = jt.posteriors_batch(ev)
        assert jt.kernel_info()["on_chip"] == 1
        jt.set_precision("exact")
        jt.specialise(observed)
        assert jt.posteriors_batch(ev).tobytes() == exact.tobytes()
    finally:
        os.environ["HB_JIT"] = "0"
    assert max_rel_err(fast, exact, floor=1e-15) < 1e-12
    assert fast.tobytes() != exact.tobytes()  # the contracted kernel really ran


PERSIST = r'''
import sys, time
sys.path.insert(0, %(root)r)
import numpy as np
import hbayes_b200 as hb
from hbayes_b200 import synth
net, observed = synth.config_c2()
pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
ev = synth.config_c2_evidence(net, observed, 4096)
path, mode = sys.argv[1], sys.argv[2]
if mode == "save":
    jt = hb.JunctionTree(pnet)
    t = time.perf_counter(); n = jt.specialise(observed); dt = time.perf_counter() - t
    out = jt.posteriors_batch(ev)
    jt.save(path)
else:
    jt = hb.JunctionTree.load(path, pnet)
    jt.posteriors_batch(ev[:2])  # creates the device-resident program (context, potentials) outside the timed call
    t = time.perf_counter(); n = jt.specialise(observed); dt = time.perf_counter() - t
    out = jt.posteriors_batch(ev)
np.save(path + "." + mode + ".npy", out)
print("RESULT", n, "%%.6f" %% dt, jt.kernel_info()["origin"], jt.kernel_info()["on_chip"])
'''


def test_saved_handle_carries_its_compiled_kernels_to_the_next_process(tmp_path):
    """f3 (the counterpart of saving a CALIBRATED tree, Bayes/ImportExport.hs:36-46,52-102): hb_jt_save writes the compiled
    programs' keys and the cubins of their generated kernels; a second process that loads the file runs on them without
    NVRTC (disk cache off here: the file alone carries them) -- specialise in well under 100 ms -- and answers bit for bit."""
    import numpy as np

    path = str(tmp_path / "c2.hbjt")
    env = dict(os.environ, HB_JIT="1", HB_JIT_CACHE="0")
    res = {}
    for mode in ("save", "load"):
        r = subprocess.run([sys.executable, "-c", PERSIST % {"root": ROOT}, path, mode], capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        res[mode] = r.stdout.rsplit("RESULT", 1)[1].split()
    n0, t0, origin0, chip0 = res["save"]
    n1, t1, origin1, chip1 = res["load"]
    assert int(n0) == int(n1) > 0 and chip0 == chip1 == "1"
    assert origin0 == "2"  # first process: compiled by NVRTC
    assert origin1 == "0"  # second process: the cubin came with the file
    assert float(t1) < 0.1 < float(t0)
    assert np.load(path + ".save.npy").tobytes() == np.load(path + ".load.npy").tobytes()
    assert os.path.getsize(path) > 100000  # the cubin is in there


def test_negative_zero_in_a_new_factor_brings_the_zero_adds_back():
    """hb_tree leaves out the `0 +` every reference sum starts with as long as no potential is negative, -0.0 or NaN
    (csrc/jit.cpp jit_zero_add_elidable).  changeFactor with a table that holds -0.0 breaks that premise: the handle drops the
    kernel it had built (the precompiled kernels answer), the next specialisation emits the adds, and every answer is the
    oracle's bit for bit -- +0.0 where the shortcut would have produced -0.0."""
    import numpy as np

    import hbayes_b200 as hb
    from hbayes_b200 import synth
    from oracle import OracleJT, OracleNet

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import same_bits

    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    ev = synth.config_c2_evidence(net, observed, 2048)
    os.environ["HB_JIT"] = "1"
    try:
        jt = hb.JunctionTree(pnet)
        assert jt.specialise(observed) > 0
        before = jt.posteriors_batch(ev)
        assert jt.kernel_info()["on_chip"] == 1
        assert same_bits(before, OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev, n_threads=8))
        # an unobserved variable with parents: its first state gets probability -0.0 under every parent configuration
        v = next(i for i in range(net.n) if i not in observed and len(net.scopes[i]) > 1)
        values = [np.asarray(x, np.float64).copy() for x in net.values]
        per_state = values[v].size // net.dims[v]
        values[v][:per_state] = -0.0
        assert jt.change_factor(hb.Factor(net.scopes[v], [net.dims[x] for x in net.scopes[v]], values[v]))
        want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, values)).posteriors_batch(ev, n_threads=8)
        col = int(sum(net.dims[:v]))
        assert not np.signbit(want[:, col]).any() and (want[:, col] == 0).all()  # the reference's 0 + (-0.0) = +0.0
        got = jt.posteriors_batch(ev)
        assert jt.kernel_info()["on_chip"] == 0  # the kernel without the adds is gone
        assert same_bits(got, want)
        assert jt.specialise(observed) > 0  # rebuilt, now with the adds
        again = jt.posteriors_batch(ev)
        assert jt.kernel_info()["on_chip"] == 1
        assert same_bits(again, want)
        assert "hb_add(zero, " in jt.tree_source(observed) or "hb_add(0.0, " in jt.tree_source(observed)
    finally:
        os.environ["HB_JIT"] = "0"
