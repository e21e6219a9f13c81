"""Split clique (BASELINE config C4) on one GPU: the eight slices a box's eight GPUs would own are evaluated one
after the other and summed (the all-reduce of the real run), then normalised.  Checked against the unsplit
engine on the same network (1e-12 relative: the sum over the split variables is reassociated), which in turn is
bit-identical to the oracle; the procedural CPT generated on the device equals the host/Python twin."""
import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
from hbayes_b200.sharded import SplitJunctionTree
from oracle import OracleJT, OracleNet
from helpers import max_rel_err, same_bits

pytestmark = pytest.mark.gpu


def evidence_rows(net, rows, seed):
    rng = np.random.RandomState(seed)
    ev = np.full((rows, net.n), -1, np.int8)
    ev[:, net.n - 1] = rng.randint(0, 2, rows)          # the child is observed in every row
    ev[:, 5] = rng.randint(0, 2, rows)                  # and one ordinary parent
    ev[rows // 2:, 1] = rng.randint(0, 2, rows - rows // 2)  # half of the rows also observe a SPLIT variable
    return ev


@pytest.mark.parametrize("n_parents", [6, 12])
def test_split_equals_unsplit_equals_oracle(n_parents):
    net, split = synth.config_c4(n_parents)
    mat = net.materialise()
    proc = hb.BayesianNetwork.from_tables_procedural(net.dims, net.scopes, net.values, net.proc_seed)
    expl = hb.BayesianNetwork.from_tables(mat.dims, mat.scopes, mat.values)
    assert proc.cpt_values(n_parents).tobytes() == mat.values[n_parents].tobytes()
    ev = evidence_rows(net, 40, 3)
    # rows with different observed sets go in two calls (device-pointer batches need one observed set)
    halves = [ev[:20], ev[20:]]
    want = np.concatenate([OracleJT(OracleNet.from_arrays(mat.dims, mat.scopes, mat.values)).posteriors_batch(h) for h in halves])
    unsplit = np.concatenate([hb.JunctionTree(expl).posteriors_batch(h) for h in halves])
    assert same_bits(unsplit, want)
    from_proc = np.concatenate([hb.JunctionTree(proc).posteriors_batch(h) for h in halves])
    assert same_bits(from_proc, want)  # the device-generated CPT is the same table
    sjt = SplitJunctionTree(proc, split, rank=0, world=1)
    assert len(sjt.slices) == 8
    got = np.concatenate([sjt.posteriors_batch(h).cpu().numpy() for h in halves])
    assert max_rel_err(got, want) <= 1e-12
    # two "ranks" of a world of 2 own four slices each; their partial sums add up to the same result
    parts = [SplitJunctionTree(proc, split, rank=r, world=2).partial_batch(halves[0]) for r in range(2)]
    summed = sjt.slices[0].normalise_batch(parts[0] + parts[1]).cpu().numpy()
    assert max_rel_err(summed, want[:20]) <= 1e-12


def test_few_rows_huge_tables_use_small_tiles():
    """2^21-entry CPT, 3 rows: the row arena is sized for 4 rows, not 64"""
    net, split = synth.config_c4(20)
    proc = hb.BayesianNetwork.from_tables_procedural(net.dims, net.scopes, net.values, net.proc_seed)
    jt = hb.JunctionTree(proc)
    ev = np.full((3, net.n), -1, np.int8)
    ev[:, net.n - 1] = [0, 1, 0]
    out = jt.posteriors_batch(ev)
    assert jt.last_stats()["rows_per_tile"] == 4
    off = np.concatenate([[0], np.cumsum(net.dims)])
    for v in range(net.n):
        assert np.abs(out[:, off[v]:off[v + 1]].sum(1) - 1).max() < 1e-12
    sjt = SplitJunctionTree(proc, split)
    assert max_rel_err(sjt.posteriors_batch(ev).cpu().numpy(), out) <= 1e-12


def test_few_rows_huge_tables_on_generated_kernels():
    """C4's regime (a handful of rows, tables of millions of entries) on the generated step kernels: the groups of a streaming
    step are dealt round robin, so a warp of a 4-row tile still reads contiguous memory (csrc/jit.cpp).  Same bits as the
    precompiled kernels, for the unsplit handle and for one slice of the split clique (scaled one-entry tables)."""
    net, split = synth.config_c4(20)
    proc = hb.BayesianNetwork.from_tables_procedural(net.dims, net.scopes, net.values, net.proc_seed)
    ev = np.full((5, net.n), -1, np.int8)
    ev[:, net.n - 1] = [0, 1, 0, 1, 1]
    jt = hb.JunctionTree(proc)
    before = jt.posteriors_batch(ev)
    assert jt.specialised_kernels() == 0
    n = jt.specialise([net.n - 1])
    after = jt.posteriors_batch(ev)
    if n == 0:
        pytest.skip("NVRTC is not available on this box")
    assert jt.specialised_kernels() == n and jt.last_stats()["rows_per_tile"] == 6
    assert same_bits(after, before)
    combo = [(v, i % 2) for i, v in enumerate(split)]
    ev2 = ev.copy()
    for v, s in combo:
        ev2[:, v] = s
    observed = sorted(split + [net.n - 1])
    sl = hb.JunctionTree(proc, split=combo)
    import torch

    d_ev = torch.from_numpy(ev2).cuda()
    b = sl.posteriors_batch(d_ev).cpu().numpy()
    assert sl.specialise(observed) > 0
    a = sl.posteriors_batch(d_ev).cpu().numpy()
    assert sl.specialised_kernels() > 0 and same_bits(a, b)
