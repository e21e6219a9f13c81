"""Split clique with the exchange inside the library: two processes, two GPUs, one
slice of a 2^20-entry clique table each; messages that leave the split clique and posteriors inside it are summed with
ncclAllReduce by libhbayes_cuda itself (hb_jt_compile_split_nccl).  Checked against the unsplit engine on one GPU (which
is bit-identical to the oracle) to 1e-12, on every rank.  The network has clusters on both sides of the split clique,
so there are messages into it (replicated), out of it (reduced) and posteriors that never see the split."""
import os
import sys
import time

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import _lib, synth
from helpers import max_rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split_network(n_parents, seed=7):
    """config C4's child-with-many-parents plus a grandchild below the child and a second child of one parent"""
    net, split = synth.config_c4(n_parents)
    mat = net.materialise()
    rng = np.random.RandomState(seed)

    def cpt(d, npar):
        v = rng.uniform(0.05, 1, size=(d, npar))
        return (v / v.sum(0)).reshape(-1)

    n = mat.n
    dims = list(mat.dims) + [2, 3]
    scopes = [list(s) for s in mat.scopes] + [[n, n - 1], [n + 1, n_parents - 1]]
    values = [np.asarray(v) for v in mat.values] + [cpt(2, 2), cpt(3, 2)]
    return dims, scopes, values, split[:1]  # split along ONE binary parent: two slices, two ranks


def evidence(dims, rows, n_parents):
    rng = np.random.RandomState(11)
    ev = np.full((rows, len(dims)), -1, np.int8)
    ev[:, len(dims) - 2] = rng.randint(0, 2, rows)   # the grandchild
    ev[:, 3] = rng.randint(0, 2, rows)               # an ordinary parent
    return ev


def test_split_program_has_its_exchange_steps():
    """structure (no GPU needed, but kept with the NCCL tests): which tables are all-reduced, which queries gathered"""
    dims, scopes, values, split = split_network(6)
    pnet = hb.BayesianNetwork.from_tables(dims, scopes, values)
    h = hb.JunctionTree(pnet, device=None, split=[(split[0], 1)], nccl=(1, 2, None))
    prog = h.program(sorted([split[0], 3, len(dims) - 2]))
    kinds = {i: g["kind"] for i, g in enumerate(prog["groups"])}
    reduced = [s for s in prog["physical"] if s["kind"] == 0 and s["reduce"]]
    gathered = [s for s in prog["physical"] if s["kind"] == 1 and s["gather"]]
    assert len(gathered) == 1  # the posterior of the split variable itself
    assert {kinds[s["group"]] for s in reduced} >= {"up", "posterior"} or {kinds[s["group"]] for s in reduced} >= {"down", "posterior"}
    plain = hb.JunctionTree(pnet, device=None, split=[(split[0], 1)]).program(sorted([split[0], 3, len(dims) - 2]))
    assert not any(s.get("reduce") for s in plain["physical"] if s["kind"] == 0)  # cutset form: no exchange inside the program


WORKER = r'''
import os, sys, time
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import torch
rank, world, idfile, outfile, n_parents = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
torch.cuda.set_device(rank)
import hbayes_b200 as hb
from test_gpu_split_nccl import split_network, evidence
if rank == 0:
    uid = hb.nccl_unique_id()
    with open(idfile + ".tmp", "wb") as f:
        f.write(uid)
    os.rename(idfile + ".tmp", idfile)
else:
    for _ in range(600):
        if os.path.exists(idfile):
            break
        time.sleep(0.1)
    uid = open(idfile, "rb").read()
dims, scopes, values, split = split_network(n_parents)
pnet = hb.BayesianNetwork.from_tables(dims, scopes, values)
ev = evidence(dims, 24, n_parents)
jt = hb.JunctionTree(pnet, device=rank, split=[(split[0], rank)], nccl=(rank, world, uid))
got = jt.posteriors_batch(ev)                      # host buffers
dev = jt.posteriors_batch(torch.from_numpy(ev).cuda()).cpu().numpy()  # device buffers
assert got.tobytes() == dev.tobytes()
fam = jt.queries_batch([[len(dims) - 3, 0], [2, 5]], ev)  # joint posteriors inside the split clique, one of them WITH the split variable
np.savez(outfile, got=got, fam0=fam[0][2], fam1=fam[1][2], launches=jt.last_stats()["launches"])
'''


@pytest.mark.skipif(_lib.lib().hb_device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n_parents", [8, 19])
def test_two_ranks_sum_their_slices_with_nccl_inside_the_library(tmp_path, n_parents):
    import subprocess

    idfile = str(tmp_path / "nccl.id")
    procs = []
    for rank in range(2):
        out = str(tmp_path / ("rank%d.npz" % rank))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % {"root": ROOT}, str(rank), "2", idfile, out, str(n_parents)],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, so[-2000:] + se[-4000:]
    a, b = np.load(str(tmp_path / "rank0.npz")), np.load(str(tmp_path / "rank1.npz"))
    for k in ("got", "fam0", "fam1"):
        assert a[k].tobytes() == b[k].tobytes(), k  # every rank holds the complete answer
    dims, scopes, values, split = split_network(n_parents)
    pnet = hb.BayesianNetwork.from_tables(dims, scopes, values)
    ev = evidence(dims, 24, n_parents)
    whole = hb.JunctionTree(pnet)  # 2^(n_parents + 1) entries on one GPU
    want = whole.posteriors_batch(ev)
    assert max_rel_err(a["got"], want, floor=1e-15) <= 1e-12
    fam = whole.queries_batch([[len(dims) - 3, 0], [2, 5]], ev)
    assert max_rel_err(a["fam0"], fam[0][2], floor=1e-15) <= 1e-12
    assert max_rel_err(a["fam1"], fam[1][2], floor=1e-15) <= 1e-12
    if n_parents == 8:  # small enough for the oracle: the unsplit engine is the reference's arithmetic bit for bit
        from oracle import OracleJT, OracleNet
        from helpers import same_bits

        assert same_bits(want, OracleJT(OracleNet.from_arrays(dims, scopes, values)).posteriors_batch(ev))
