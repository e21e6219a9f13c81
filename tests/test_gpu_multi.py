"""The drop-in API on several GPUs of one box: a handle compiled with n_devices > 1
shards the rows of a host-buffer hb_jt_posteriors_batch over its GPUs -- one host thread and one GPU per slice, no
communication -- and answers exactly what one GPU answers.  Page-locked buffers from hb_host_alloc."""
import os

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import _lib, synth
from oracle import OracleJT, OracleNet
from helpers import same_bits

pytestmark = pytest.mark.gpu


def test_page_locked_buffers_from_the_library():
    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    ev = synth.config_c2_evidence(net, observed, 5000)
    pev = hb.host_alloc(ev.shape, np.int8)
    pev[:] = ev
    pout = hb.host_alloc((ev.shape[0], pnet.width), np.float64)
    jt = hb.JunctionTree(pnet)
    assert jt.n_devices == 1
    jt.posteriors_batch(pev, pout)
    assert same_bits(pout, jt.posteriors_batch(ev))
    want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[:256], n_threads=os.cpu_count() or 1)
    assert same_bits(pout[:256], want)


def test_bad_device_lists_are_status_codes():
    net = hb.importBayesianGraph(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cancer.net"))
    for devices in ([0, 0], [0, 99], [-1]):
        with pytest.raises(hb.HbError) as e:
            hb.JunctionTree(net, device=devices)
        assert e.value.code == _lib.HB_E_ARG


@pytest.mark.skipif(_lib.lib().hb_device_count() < 2, reason="needs two GPUs")
def test_rows_sharded_over_the_handles_gpus():
    n_gpus = min(_lib.lib().hb_device_count(), 8)
    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    rows = n_gpus * 6000 + 13
    ev = synth.config_c2_evidence(net, observed, rows)
    ev[7, observed[0]] = -1  # one row with another observed set: its slice falls back to grouping, the others do not
    one = hb.JunctionTree(pnet, device=0).posteriors_batch(ev)
    for generated in (False, True):
        os.environ["HB_JIT"] = "1" if generated else "0"
        try:
            many = hb.JunctionTree(pnet, device=list(range(n_gpus)))
            assert many.n_devices == n_gpus
            if generated:
                assert many.specialise(observed) > 0
            got = many.posteriors_batch(ev)
        finally:
            os.environ["HB_JIT"] = "0"
        assert same_bits(got, one)
    want = OracleJT(OracleNet.from_arrays(net.dims, net.scopes, net.values)).posteriors_batch(ev[-64:], n_threads=os.cpu_count() or 1)
    assert same_bits(got[-64:], want)
    # an error on one GPU's slice comes back as a status code naming the device
    bad = ev.copy()
    bad[-1, observed[1]] = 99
    with pytest.raises(hb.HbError) as e:
        many.posteriors_batch(bad)
    assert e.value.code == _lib.HB_E_ARG and "device" in str(e.value)
