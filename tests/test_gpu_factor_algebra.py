"""The reference's factor-algebra QuickCheck properties (Bayes/Factor/CPT.hs:68-98) as hypothesis tests against the
GPU kernels, through hb_factor_marginalize; plus bit-parity of product / single-variable sum-out with the oracle
(rows a3-a5 of the scope table).  Generators follow `Arbitrary CPT` (CPT.hs:53-62) widened to 2-4 states."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import hbayes_b200 as hb
import oracle
from helpers import same_bits

pytestmark = pytest.mark.gpu
SETTINGS = settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
DIMS = {v: 2 + (v * 7) % 3 for v in range(60)}  # a fixed dimension per vertex id, 2..4


@st.composite
def factors(draw, lo=0, hi=50, max_vars=4):
    k = draw(st.integers(1, max_vars))
    vs = draw(st.lists(st.integers(lo, hi), min_size=k, max_size=k, unique=True))
    size = int(np.prod([DIMS[v] for v in vs]))
    vals = draw(st.lists(st.floats(0.0, 1.0, allow_nan=False), min_size=size, max_size=size))
    return hb.Factor(vs, [DIMS[v] for v in vs], np.array(vals))


def norm(f):
    return float(np.sum(f.values))


def as_oracle(f):
    return [(v, d) for v, d in zip(f.variables, f.dims)], f.values


def iso(a, b, rtol=2e-5):
    """isomorphicFactor (Factor.hs:197-206): same function of the variables, nearlyEqual with relative 2e-5"""
    if sorted(a.variables) != sorted(b.variables):
        return False
    perm = [a.variables.index(v) for v in b.variables]
    ta = a.values.reshape(a.dims).transpose(perm) if a.variables else a.values
    return np.allclose(ta.reshape(-1), b.values, rtol=rtol, atol=1e-300)


@SETTINGS
@given(st.lists(factors(), min_size=1, max_size=4))
def test_product_bitexact_vs_oracle(fs):
    got = hb.factorProduct(fs)
    scope, vals = oracle.factor_product([as_oracle(f) for f in fs])
    assert got.variables == scope and same_bits(got.values, vals)


@SETTINGS
@given(st.lists(factors(0, 8), min_size=1, max_size=4), st.data())
def test_marginalize_one_variable_bitexact_vs_oracle(fs, data):
    v = data.draw(st.sampled_from(sorted({x for f in fs for x in f.variables})))
    got = hb.marginalizeOneVariable(fs, v)
    scope, vals = oracle.factor_product([as_oracle(f) for f in fs])
    dims = {x: d for f in fs for x, d in zip(f.variables, f.dims)}
    oscope, ovals = oracle.factor_project_out([(x, dims[x]) for x in scope], vals, [v])
    assert got.variables == oscope and same_bits(got.values, ovals)


@SETTINGS
@given(factors(), st.floats(0.1, 10.0))
def test_scale_norm(f, s):  # CPT.hs:68-69: norm (s . f) == s * norm f
    scalar = hb.Factor([], [], np.array([s]))
    assert np.isclose(norm(hb.factorProduct([scalar, f])), s * norm(f), rtol=1e-12)


@SETTINGS
@given(factors(0, 20), factors(30, 50))
def test_product_then_project_recovers_the_factor(a, b):  # CPT.hs:71-76 (disjoint scopes)
    p = hb.factorProduct([a, b])
    back = hb.factorProjectOut(b.variables, p)
    want = hb.Factor(a.variables, a.dims, a.values * norm(b))
    assert iso(back, want, rtol=1e-10)


@SETTINGS
@given(factors(0, 12, max_vars=4), st.data())
def test_projection_commutes(f, data):  # CPT.hs:91-98
    if len(f.variables) < 2:
        return
    x, y = data.draw(st.permutations(f.variables))[:2]
    assert iso(hb.factorProjectOut([x, y], f), hb.factorProjectOut([y, x], f), rtol=1e-10)


@SETTINGS
@given(factors())
def test_project_all_is_norm(f):  # CPT.hs:85-89
    r = hb.factorProjectOut(f.variables, f)
    assert r.variables == [] and np.isclose(r.values[0], norm(f), rtol=1e-12)


@SETTINGS
@given(factors(0, 10), factors(5, 15), factors(10, 20))
def test_product_associative(a, b, c):  # CPT.hs:81-83
    left = hb.factorProduct([hb.factorProduct([a, b]), c])
    right = hb.factorProduct([a, hb.factorProduct([b, c])])
    assert iso(left, right, rtol=1e-12)


def test_scalars_only_and_bad_arguments():
    s = hb.factorProduct([hb.Factor([], [], np.array([0.5])), hb.Factor([], [], np.array([0.25]))])
    assert s.variables == [] and s.values[0] == 0.125
    with pytest.raises(hb.HbError):
        hb.marginalizeOneVariable([hb.Factor([1], [2], np.array([0.5, 0.5]))], 3)  # variable in no factor


@SETTINGS
@given(factors(0, 12, max_vars=4), st.data())
def test_project_to_keeps_the_listed_variables(f, data):  # Factor.hs:183-188
    keep = data.draw(st.lists(st.sampled_from(f.variables), unique=True, max_size=len(f.variables))) if f.variables else []
    got = hb.factorProjectTo(keep, f)
    want = hb.factorProjectOut([v for v in f.variables if v not in keep], f)
    assert got.variables == want.variables == [v for v in f.variables if v in keep]
    assert same_bits(got.values, want.values)
