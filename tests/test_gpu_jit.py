"""Per-step specialised kernels (csrc/jit.cpp, NVRTC): the library switches to them on the first large batch of a program.
Here they are forced for every program (HB_JIT_MIN_WORK=1, read once per process -> a subprocess) and compared bit for bit
with the oracle AND with the templated kernels (HB_JIT=0) on the same inputs: junction tree with evidence on random
networks incl. shuffled ids, family queries, variable elimination, and a batch with an odd number of rows."""
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
from helpers import random_net, same_bits

out = {}
for seed, rows in ((5, 257), (9, 64), (7, 3), (8, 18), (12, 1000)):  # incl. tiles of fewer than 64 rows
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
    out["jt%%d" %% seed] = got
    assert same_bits(got, OracleJT(onet).posteriors_batch(ev)), ("jt", seed)
    fam = jt.queries_batch([net.scopes[v] for v in range(net.n)], ev)
    out["fam%%d" %% seed] = np.concatenate([f[2] for f in fam], axis=1)
    order = hb.minFillOrder(pnet)
    r = [order[-1]]
    p = [v for v in order if v not in r]
    if not set(r) & set(observed):
        ve = hb.VariableElimination(pnet, p, r).posterior_batch(ev)
        out["ve%%d" %% seed] = ve
        assert same_bits(ve, onet.posterior_marginal_batch(p, r, ev)), ("ve", seed)
# the explicit form: kernels built right after the compilation, before any batch (hb_jt_specialise)
jt2 = hb.JunctionTree(pnet)
n = jt2.specialise(observed)
out["explicit"] = jt2.posteriors_batch(ev)
assert jt2.specialised_kernels() == n
assert same_bits(out["explicit"], got)
print("specialise ->", n)
np.savez(sys.argv[1], **out)
'''


def run(tmp_path, name, env):
    path = str(tmp_path / (name + ".npz"))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, path], capture_output=True, text=True, env=e, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return path, r.stderr + r.stdout


@pytest.mark.parametrize("knobs", [{}, {"HB_JIT_INTERLEAVE": "2", "HB_JIT_UNROLL": "4"}, {"HB_JIT_INTERLEAVE": "0"}, {"HB_JIT_UNROLL": "2", "HB_JIT_MERGE": "0"},
                                   # register blocks over several output variables for EVERY re-reading step (by default only joints of
                                   # >= 8192 entries, i.e. C3 in tests/test_gpu_configs.py), with bodies so small that some keep run-time loops
                                   {"HB_JIT_BLOCK_MIN_JOINT": "1", "HB_JIT_BLOCK_BODY": "12"}],
                         ids=["default", "interleave-all-unroll4", "strips", "unroll2-no-digit-merge", "blocks-everywhere"])
def test_specialised_kernels_bit_identical_to_oracle_and_to_the_templated_kernels(tmp_path, knobs):
    import numpy as np

    # HB_TREE=0: programs this small would otherwise run on the on-chip kernel (tests/test_gpu_onchip.py)
    jit, log = run(tmp_path, "jit", {"HB_JIT": "1", "HB_TREE": "0", "HB_JIT_MIN_WORK": "1", "HB_JIT_SYNC": "1", "HB_JIT_VERBOSE": "1",
                                     "HB_CACHE_DIR": str(tmp_path / "cache"), **knobs})
    assert "specialised kernels" in log, log[-2000:]  # they were really built and used
    assert int(log.rsplit("specialise ->", 1)[1].split()[0]) > 0
    plain, plain_log = run(tmp_path, "plain", {"HB_JIT": "0"})
    assert int(plain_log.rsplit("specialise ->", 1)[1].split()[0]) == 0
    a, b = np.load(jit), np.load(plain)
    assert sorted(a.files) == sorted(b.files) and len(a.files) >= 6
    for k in a.files:
        assert a[k].tobytes() == b[k].tobytes(), k
