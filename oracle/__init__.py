"""ctypes front-end of the parity oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (hbayes_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/liboracle.so with the committed Makefile (building the checker is not using it)."""
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, i32, i64, dp, ip = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int)
        i64p, i8p, cpp = C.POINTER(C.c_int64), C.POINTER(C.c_int8), C.POINTER(C.c_void_p)
        L.orc_last_error.restype = C.c_char_p
        L.orc_string_free.argtypes = [vp]
        L.orc_net_create.restype = vp
        L.orc_net_create.argtypes = [i32, ip, i64p, ip, i64p, dp]
        L.orc_net_load_hugin.restype = vp
        L.orc_net_load_hugin.argtypes = [C.c_char_p]
        L.orc_net_free.argtypes = [vp]
        L.orc_net_nvars.argtypes = [vp]
        L.orc_net_describe_json.restype = vp
        L.orc_net_describe_json.argtypes = [vp]
        L.orc_jt_create.restype = vp
        L.orc_jt_create.argtypes = [vp, i32]
        L.orc_jt_create_with_order.restype = vp
        L.orc_jt_create_with_order.argtypes = [vp, ip]
        L.orc_jt_free.argtypes = [vp]
        L.orc_jt_change_evidence.argtypes = [vp, i32, ip, ip, cpp]
        L.orc_jt_posterior.restype = i64
        L.orc_jt_posterior.argtypes = [vp, i32, ip, ip, dp, i64]
        L.orc_jt_describe_json.restype = vp
        L.orc_jt_describe_json.argtypes = [vp]
        L.orc_jt_posteriors_batch.argtypes = [vp, i64, i8p, dp, i32]
        L.orc_ve_posterior.restype = i64
        L.orc_ve_posterior.argtypes = [vp, ip, i32, ip, i32, i32, ip, ip, ip, dp, i64, cpp]
        L.orc_ve_posterior_batch.argtypes = [vp, ip, i32, ip, i32, i64, i8p, dp, i64, i32]
        L.orc_ve_order.argtypes = [vp, i32, ip]
        L.orc_factor_product.restype = i64
        L.orc_factor_product.argtypes = [i32, ip, ip, ip, i64p, dp, ip, ip, dp, i64]
        L.orc_factor_project_out.restype = i64
        L.orc_factor_project_out.argtypes = [i32, ip, ip, dp, i32, ip, ip, ip, dp, i64]
        _LIB = L
    return _LIB


def _ia(x):
    a = np.ascontiguousarray(x, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _take_json(ptr):
    s = C.cast(ptr, C.c_char_p).value.decode()
    lib().orc_string_free(ptr)
    return json.loads(s)


class OracleNet:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())
        self.h = handle
        d = _take_json(lib().orc_net_describe_json(self.h))
        self.names, self.dims, self.scopes, self.values = d["names"], d["dims"], d["scopes"], d["values"]

    @staticmethod
    def load_hugin(path):
        return OracleNet(lib().orc_net_load_hugin(str(path).encode()))

    @staticmethod
    def from_arrays(dims, scopes, values):
        dims_a, dims_p = _ia(dims)
        cpt_off = np.zeros(len(dims) + 1, np.int64)
        val_off = np.zeros(len(dims) + 1, np.int64)
        for v in range(len(dims)):
            cpt_off[v + 1] = cpt_off[v] + len(scopes[v])
            val_off[v + 1] = val_off[v] + len(values[v])
        sc_a, sc_p = _ia(np.concatenate([np.asarray(s, np.int32) for s in scopes]))
        vals = np.ascontiguousarray(np.concatenate([np.asarray(v, np.float64) for v in values]))
        h = lib().orc_net_create(len(dims), dims_p, cpt_off.ctypes.data_as(C.POINTER(C.c_int64)), sc_p,
                                 val_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                 vals.ctypes.data_as(C.POINTER(C.c_double)))
        return OracleNet(h)

    @property
    def n(self):
        return len(self.dims)

    def ve_order(self, metric):
        out = np.zeros(self.n, np.int32)
        k = lib().orc_ve_order(self.h, {"min_degree": 0, "min_fill": 1}[metric], out.ctypes.data_as(C.POINTER(C.c_int)))
        return out[:k].tolist()

    def posterior_marginal(self, p, r, evidence=(), trace=False):
        """posteriorMarginal g p r assignment; evidence = [(var, state)] in the caller's order."""
        pa, pp = _ia(p)
        ra, rp = _ia(r)
        ev, evp = _ia([e[0] for e in evidence])
        es, esp = _ia([e[1] for e in evidence])
        size = int(np.prod([self.dims[v] for v in r])) if len(r) else 1
        out = np.zeros(size, np.float64)
        scope = np.zeros(max(len(r), 1), np.int32)
        tj = C.c_void_p()
        k = lib().orc_ve_posterior(self.h, pp, len(p), rp, len(r), len(evidence), evp, esp,
                                   scope.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(C.POINTER(C.c_double)),
                                   size, C.byref(tj) if trace else None)
        if k < 0:
            raise RuntimeError("oracle ve_posterior failed")
        res = (scope[: len(r)].tolist(), out[:k])
        return res + (_take_json(tj),) if trace else res

    def posterior_marginal_batch(self, p, r, evidence, n_threads=1):
        pa, pp = _ia(p)
        ra, rp = _ia(r)
        ev = np.ascontiguousarray(evidence, np.int8)
        width = int(np.prod([self.dims[v] for v in r]))
        out = np.zeros((ev.shape[0], width), np.float64)
        lib().orc_ve_posterior_batch(self.h, pp, len(p), rp, len(r), ev.shape[0], ev.ctypes.data_as(C.POINTER(C.c_int8)),
                                     out.ctypes.data_as(C.POINTER(C.c_double)), width, n_threads)
        return out

    def __del__(self):
        try:
            lib().orc_net_free(self.h)
        except Exception:
            pass


class OracleJT:
    """createJunctionTree / changeEvidence / posterior of the restated reference."""

    def __init__(self, net: OracleNet, cmp="weighted", order=None):
        self.net = net
        if order is not None:
            oa, op = _ia(order)
            self.h = lib().orc_jt_create_with_order(net.h, op)
        else:
            self.h = lib().orc_jt_create(net.h, {"weighted": 0, "added": 1}[cmp])
        if not self.h:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())

    def describe(self):
        return _take_json(lib().orc_jt_describe_json(self.h))

    def change_evidence(self, evidence, trace=False):
        ev, evp = _ia([e[0] for e in evidence])
        es, esp = _ia([e[1] for e in evidence])
        tj = C.c_void_p()
        lib().orc_jt_change_evidence(self.h, len(evidence), evp, esp, C.byref(tj) if trace else None)
        return _take_json(tj) if trace else None

    def posterior(self, qvars):
        qa, qp = _ia(qvars)
        size = int(np.prod([self.net.dims[v] for v in qvars]))
        out = np.zeros(size, np.float64)
        scope = np.zeros(len(qvars), np.int32)
        k = lib().orc_jt_posterior(self.h, len(qvars), qp, scope.ctypes.data_as(C.POINTER(C.c_int)),
                                   out.ctypes.data_as(C.POINTER(C.c_double)), size)
        if k == -1:
            return None
        return scope.tolist(), out[:k]

    def posteriors_batch(self, evidence, n_threads=1):
        ev = np.ascontiguousarray(evidence, np.int8)
        assert ev.ndim == 2 and ev.shape[1] == self.net.n
        out = np.zeros((ev.shape[0], int(sum(self.net.dims))), np.float64)
        lib().orc_jt_posteriors_batch(self.h, ev.shape[0], ev.ctypes.data_as(C.POINTER(C.c_int8)),
                                      out.ctypes.data_as(C.POINTER(C.c_double)), n_threads)
        return out

    def __del__(self):
        try:
            lib().orc_jt_free(self.h)
        except Exception:
            pass


def learn_em(dims, scopes, values, samples):
    """learnEM (Bayes/EM.hs:36-66) with the posteriors of the C++ oracle's junction tree and the twin's cptSum /
    cptDivide.  samples: int8 matrix (-1 = missing).  Returns [(scope, values)] per vertex in the reference's layout."""
    from . import twin

    onet = OracleNet.from_arrays(dims, scopes, values)
    ojt = OracleJT(onet)
    tnet = twin.Network.from_tables(dims, [list(sc[1:]) for sc in scopes], [list(map(float, v)) for v in values])

    def posterior_fn(evidence, vs, _state={"ev": None}):
        if _state["ev"] is not evidence:
            ojt.change_evidence(evidence)
            _state["ev"] = evidence
        sc, vals = ojt.posterior(vs)
        return twin.Factor([(x, dims[x]) for x in sc], [float(x) for x in vals])

    rows = [[(v, int(s)) for v, s in enumerate(row) if s >= 0] for row in np.asarray(samples)]
    learnt = twin.learn_em(tnet, rows, posterior_fn)
    return [([dv[0] for dv in f.vars], np.asarray(f.vals, np.float64)) for f in learnt]


def factor_product(factors):
    """factors = [(scope [(var, dim)...], values)]; returns (scope vars, values)."""
    soff = [0]
    sc, dm, voff, vals = [], [], [0], []
    for scope, v in factors:
        sc += [s[0] for s in scope]
        dm += [s[1] for s in scope]
        soff.append(len(sc))
        vals += list(v)
        voff.append(len(vals))
    dims = {}
    for scope, _ in factors:
        for s in scope:
            dims[s[0]] = s[1]
    size = int(np.prod(list(dims.values()))) if dims else 1
    out = np.zeros(size, np.float64)
    oscope = np.zeros(max(len(dims), 1), np.int32)
    onv = C.c_int()
    sa, sp = _ia(soff)
    ca, cp = _ia(sc)
    da, dp_ = _ia(dm)
    va = np.ascontiguousarray(vals, np.float64)
    vo = np.ascontiguousarray(voff, np.int64)
    k = lib().orc_factor_product(len(factors), sp, cp, dp_, vo.ctypes.data_as(C.POINTER(C.c_int64)),
                                 va.ctypes.data_as(C.POINTER(C.c_double)), oscope.ctypes.data_as(C.POINTER(C.c_int)),
                                 C.byref(onv), out.ctypes.data_as(C.POINTER(C.c_double)), size)
    return oscope[: onv.value].tolist(), out[:k]


def factor_project_out(scope, values, sum_vars):
    ca, cp = _ia([s[0] for s in scope])
    da, dp_ = _ia([s[1] for s in scope])
    sa, sp = _ia(sum_vars)
    va = np.ascontiguousarray(values, np.float64)
    out = np.zeros(len(va), np.float64)
    oscope = np.zeros(max(len(scope), 1), np.int32)
    onv = C.c_int()
    k = lib().orc_factor_project_out(len(scope), cp, dp_, va.ctypes.data_as(C.POINTER(C.c_double)), len(sum_vars), sp,
                                     oscope.ctypes.data_as(C.POINTER(C.c_int)), C.byref(onv),
                                     out.ctypes.data_as(C.POINTER(C.c_double)), len(va))
    return oscope[: onv.value].tolist(), out[:k]
