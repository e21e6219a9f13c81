"""GPU parity on the other BASELINE.json configurations (the bench measures C2): C1 cancer.net latency path,
C3 500-node network (max cluster ~2^20), C5 grid networks through VariableElimination with the reference's
min-fill order.  Oracle comparison at sizes the oracle finishes in seconds; size-independent properties at
the full size."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import synth
from oracle import OracleJT, OracleNet
from helpers import max_rel_err, same_bits

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def pair(net):
    return hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values), OracleNet.from_arrays(net.dims, net.scopes, net.values)


def test_c3_500_nodes_bitexact_and_properties():
    net, observed = synth.config_c3()
    pnet, onet = pair(net)
    pjt, ojt = hb.JunctionTree(pnet), OracleJT(onet)
    assert pjt.describe() == ojt.describe()
    sizes = [np.prod([float(net.dims[v]) for v in c]) for c in pjt.describe()["clusters"]]
    assert 2 ** 19.5 <= max(sizes) <= 2 ** 20.5  # BASELINE config: max clique ~2^20 entries
    ev = synth.config_c3_evidence(net, observed, 4096)
    out = pjt.posteriors_batch(ev)
    want = ojt.posteriors_batch(ev[:384], n_threads=THREADS)
    assert max_rel_err(out[:384], want) <= 1e-12
    assert same_bits(out[:384], want)
    off = np.concatenate([[0], np.cumsum(net.dims)])
    for v in range(0, net.n, 7):
        assert np.abs(out[:, off[v]:off[v + 1]].sum(1) - 1).max() < 1e-12
    # priors (no evidence: every table is batch-invariant, computed once at upload)
    prior = pjt.posteriors_batch(np.full((3, net.n), -1, np.int8))
    assert same_bits(prior, ojt.posteriors_batch(np.full((3, net.n), -1, np.int8)))


@pytest.mark.parametrize("shape,rows", [((10, 10), 256), ((12, 12), 64)])
def test_c5_grid_ve_bitexact(shape, rows):
    g = synth.config_c5(shape[0], shape[1])
    pnet, onet = pair(g)
    order = hb.minFillOrder(pnet)
    assert order == onet.ve_order("min_fill")
    p, r, observed = synth.config_c5_query(g, order)
    ve = hb.VariableElimination(pnet, p, r)
    ev = synth.evidence_matrix(g, rows, observed, 400003)
    got = ve.posterior_batch(ev)
    want = onet.posterior_marginal_batch(p, r, ev, n_threads=THREADS)
    assert max_rel_err(got, want) <= 1e-12
    assert same_bits(got, want)


def test_c5_grid_16x16_properties():
    """2^17-entry tables per row, ~12 M product entries per row: too slow for the oracle at this row count;
    checked through normalisation, permutation invariance and agreement with the min-degree order."""
    g = synth.config_c5(16, 16)
    pnet, _ = pair(g)
    p, r, observed = synth.config_c5_query(g, hb.minFillOrder(pnet))
    ve = hb.VariableElimination(pnet, p, r)
    ev = synth.evidence_matrix(g, 96, observed, 400003)
    out = ve.posterior_batch(ev)
    assert np.abs(out.sum(1) - 1).max() < 1e-12 and (out > 0).all()
    perm = np.random.RandomState(1).permutation(96)
    assert same_bits(ve.posterior_batch(ev[perm]), out[perm])
    p2 = [v for v in hb.minDegreeOrder(pnet) if v not in r]
    out2 = hb.VariableElimination(pnet, p2, r).posterior_batch(ev)
    assert max_rel_err(out2, out) < 1e-9  # a different elimination order: same distribution, different rounding


def test_c2_half_million_rows_from_host_buffers_equals_device_path_and_oracle():
    """the host-pointer path whose chunk schedule (10/30/30/20/10 %, D2H overlapping the next chunk's kernels)
    only engages from 2^19 rows is compared with the device-pointer path on ALL rows and with the oracle on a prefix --
    on the precompiled kernels and on the generated on-chip kernel (the one bench.py times)."""
    import torch

    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    rows = (1 << 19) + 77  # ragged: the last chunk and the last tile are partial
    ev = synth.config_c2_evidence(net, observed, rows)
    want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[:2048], n_threads=os.cpu_count() or 1)
    for generated in (False, True):
        os.environ["HB_JIT"] = "1" if generated else "0"
        try:
            jt = hb.JunctionTree(pnet)
            if generated:
                assert jt.specialise(observed) > 0 and jt.kernel_info()["on_chip_eligible"] == 1
            host = jt.posteriors_batch(ev)  # numpy in, numpy out: pageable host memory -> the handle's page-locked bounce ring
            assert bool(jt.kernel_info()["on_chip"]) == generated
            for knob, value in (("HB_COPY_THREADS", "3"), ("HB_BOUNCE", "0")):  # odd slice count; plain cudaMemcpyAsync
                os.environ[knob] = value
                try:
                    again = np.full_like(host, -1.0)
                    jt.posteriors_batch(ev, again)
                finally:
                    del os.environ[knob]
                assert same_bits(again, host)
            dev = jt.posteriors_batch(torch.from_numpy(ev).cuda()).cpu().numpy()
            assert same_bits(host, dev)
            assert same_bits(host[:2048], want)
            assert same_bits(host[-1:], OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[-1:]))
            pinned_ev = torch.from_numpy(ev).pin_memory()
            pinned_out = torch.empty((rows, pnet.width), dtype=torch.float64, pin_memory=True)
            jt.posteriors_batch(pinned_ev, pinned_out)
            assert same_bits(pinned_out.numpy(), host)
        finally:
            os.environ["HB_JIT"] = "0"


def test_c3_generated_step_kernels_bitexact():
    """the per-step generated kernels (hb_step_<i>, what C3 runs on in bench.py) against the oracle, in the
    suite's own process: hb_jt_specialise builds them for C3's program (too large for the on-chip kernel)."""
    net, observed = synth.config_c3()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    ev = synth.config_c3_evidence(net, observed, 192)
    os.environ["HB_JIT"] = "1"
    try:
        jt = hb.JunctionTree(pnet)
        n = jt.specialise(observed)
        got = jt.posteriors_batch(ev)
        info = jt.kernel_info()
    finally:
        os.environ["HB_JIT"] = "0"
    assert n >= 20 and jt.specialised_kernels() == n and not info["on_chip"] and not info["on_chip_eligible"]
    want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[:64], n_threads=os.cpu_count() or 1)
    assert same_bits(got[:64], want)
    plain = hb.JunctionTree(pnet).posteriors_batch(ev)  # HB_JIT=0: the precompiled kernels
    assert same_bits(got, plain)
