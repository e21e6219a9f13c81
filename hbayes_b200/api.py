"""Reference-shaped Python front-end over the C ABI (see package docstring for the file:line map)."""
from __future__ import annotations

import ctypes as C
import json
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import (HB_DEVICE_PTRS, HB_E_NO_CLUSTER, HB_HOST_PTRS, HB_STRUCTURE_ONLY, HbError, check, f64p, i32p, i64p, lib)

# the two comparators the reference exports for createJunctionTree (Bayes/FactorElimination.hs:71-111)
nodeComparisonForTriangulation = 0
numberOfAddedEdges = 1


def _i32(xs):
    a = np.ascontiguousarray(xs, dtype=np.int32)
    return a, a.ctypes.data_as(i32p)


def _take_json(ptr):
    s = C.cast(ptr, C.c_char_p).value.decode()
    lib().hb_string_free(ptr)
    return json.loads(s)


class Factor:
    """A discrete factor as the reference's CPT shows it: variable order + flat values, first variable slowest
    (Bayes/Factor/PrivateCPT.hs:67-72)."""

    def __init__(self, variables: Sequence[int], dims: Sequence[int], values: np.ndarray):
        self.variables = list(variables)
        self.dims = list(dims)
        self.values = np.asarray(values, np.float64)

    def __repr__(self):
        return "Factor(%s, %s)" % (self.variables, self.values.tolist())


class BayesianNetwork:
    """What `runBN` of a BNMonad yields (Bayes/BayesianNetwork.hs:211-212): an immutable hb_net handle."""

    def __init__(self, handle):
        self._h = handle
        p = C.c_void_p()
        check(lib().hb_net_describe_json(self._h, C.byref(p)))
        d = _take_json(p)
        self.names, self.dims, self.scopes, self.values = d["names"], d["dims"], d["scopes"], d["values"]
        self.ids = {n: i for i, n in enumerate(self.names)}

    @staticmethod
    def from_tables(dims, scopes, values):
        """CPT v: scope `scopes[v]` = [v, parents...], values in reference layout."""
        n = len(dims)
        da, dp = _i32(dims)
        cpt_off = np.zeros(n + 1, np.int64)
        val_off = np.zeros(n + 1, np.int64)
        for v in range(n):
            cpt_off[v + 1] = cpt_off[v] + len(scopes[v])
            val_off[v + 1] = val_off[v] + len(values[v])
        sa, sp = _i32(np.concatenate([np.asarray(s, np.int32) for s in scopes]))
        vals = np.ascontiguousarray(np.concatenate([np.asarray(v, np.float64) for v in values]))
        h = C.c_void_p()
        check(lib().hb_net_create(n, dp, cpt_off.ctypes.data_as(i64p), sp, val_off.ctypes.data_as(i64p),
                                  vals.ctypes.data_as(f64p), C.byref(h)))
        return BayesianNetwork(h)

    @staticmethod
    def from_tables_procedural(dims, scopes, values, proc_seed):
        """like from_tables; CPT v with proc_seed[v] != 0 is procedural (values[v] must be empty)"""
        n = len(dims)
        da, dp = _i32(dims)
        cpt_off = np.zeros(n + 1, np.int64)
        val_off = np.zeros(n + 1, np.int64)
        for v in range(n):
            cpt_off[v + 1] = cpt_off[v] + len(scopes[v])
            val_off[v + 1] = val_off[v] + len(values[v])
        sa, sp = _i32(np.concatenate([np.asarray(s, np.int32) for s in scopes]))
        vals = np.ascontiguousarray(np.concatenate([np.asarray(v, np.float64) for v in values] + [np.zeros(1)]))
        seeds = np.ascontiguousarray(proc_seed, np.uint64)
        h = C.c_void_p()
        check(lib().hb_net_create_procedural(n, dp, cpt_off.ctypes.data_as(i64p), sp, val_off.ctypes.data_as(i64p),
                                             vals.ctypes.data_as(f64p), seeds.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(h)))
        return BayesianNetwork(h)

    def cpt_values(self, v, offset=0, count=None):
        """entries of CPT v in reference layout (works for procedural CPTs too)"""
        if count is None:
            count = int(np.prod([self.dims[x] for x in self.scopes[v]])) - offset
        out = np.zeros(count, np.float64)
        check(lib().hb_net_cpt_values(self._h, v, offset, count, out.ctypes.data_as(f64p)))
        return out

    def change_factor(self, factor: "Factor") -> "BayesianNetwork":
        """changeFactor f g on a network (Bayes/Factor.hs:66-81): a NEW network in which the CPT whose variable list
        equals the factor's (same variables, same order) carries the factor's values; no match -> an equal network."""
        values = [np.asarray(v, np.float64) for v in self.values]
        for v in range(self.n):
            if list(self.scopes[v]) == list(factor.variables):
                values[v] = np.asarray(factor.values, np.float64).reshape(-1)
        return BayesianNetwork.from_tables(self.dims, self.scopes, values)

    @property
    def n(self):
        return len(self.dims)

    @property
    def width(self):
        return int(sum(self.dims))

    def __del__(self):
        try:
            lib().hb_net_free(self._h)
        except Exception:
            pass


def importBayesianGraph(path) -> BayesianNetwork | None:
    """importBayesianGraph + runBN; returns None when the file does not parse (the reference's `Nothing`)."""
    h = C.c_void_p()
    code = lib().hb_net_load_hugin(str(path).encode(), C.byref(h))
    if code != 0:
        return None
    return BayesianNetwork(h)


def _order(net: BayesianNetwork, metric: int):
    out = np.zeros(net.n, np.int32)
    k = C.c_int32()
    check(lib().hb_ve_order(net._h, metric, out.ctypes.data_as(i32p), C.byref(k)))
    return out[: k.value].tolist()


def degreeOrder(net: BayesianNetwork, order) -> int:
    """degreeOrder (Bayes/VariableElimination.hs:186-207)"""
    oa, op = _i32(list(order))
    w = C.c_int32()
    check(lib().hb_ve_degree_order(net._h, op, len(oa), C.byref(w)))
    return int(w.value)


def minDegreeOrder(net):
    return _order(net, 0)


def minFillOrder(net):
    return _order(net, 1)


def _device_args(device):
    if device is None:  # structure only: parity hooks work, compute calls fail loudly
        return None, HB_STRUCTURE_ONLY, None
    a, p = _i32([device] if isinstance(device, (int, np.integer)) else list(device))  # a list: host batches are sharded over them
    return p, len(a), a


def nccl_unique_id() -> bytes:
    """128 bytes from ncclGetUniqueId: rank 0 creates it, every rank of a split clique passes it to JunctionTree(nccl=...)"""
    buf = C.create_string_buffer(128)
    check(lib().hb_nccl_unique_id(buf))
    return buf.raw


def host_alloc(shape, dtype):
    """a numpy array over page-locked host memory from hb_host_alloc (freed when the array is collected)"""
    import weakref

    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = C.c_void_p()
    check(lib().hb_host_alloc(n, C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib().hb_host_free, p)
    return arr


class _Batched:
    """Shared by JunctionTree and VariableElimination: stream / workspace / stats plumbing."""

    _prefix = ""

    def _fn(self, name):
        return getattr(lib(), self._prefix + name)

    def set_stream(self, cuda_stream: int):
        check(self._fn("set_stream")(self._h, C.c_void_p(cuda_stream)))

    def set_precision(self, mode: str):
        """"exact" (default: bit-identical to the reference) or "contract" (generated kernels may fuse multiply-add: ~1e-15)"""
        check(self._fn("set_precision")(self._h, {"exact": 0, "contract": 1}[mode]))

    def set_workspace(self, nbytes: int):
        check(self._fn("set_workspace")(self._h, int(nbytes)))

    def set_timing(self, on: bool):
        check(self._fn("set_timing")(self._h, 1 if on else 0))

    def last_timing(self):
        """CUDA-event milliseconds of the last batch call: prep / prodsum / output / total (set_timing(True) first)"""
        out = np.zeros(4, np.float64)
        check(self._fn("last_timing")(self._h, out.ctypes.data_as(f64p)))
        return dict(zip(["prep_ms", "prodsum_ms", "output_ms", "total_ms"], out.tolist()))

    def last_stats(self):
        out = np.zeros(8, np.int64)
        check(self._fn("last_stats")(self._h, out.ctypes.data_as(i64p)))
        keys = ["launches", "tiles", "rows_per_tile", "arena_bytes", "row_entries", "prod_per_row", "row_read_per_row",
                "row_write_per_row"]
        return dict(zip(keys, out.tolist()))

    def _run_matrix(self, fn, evidence, out, width):
        """evidence/out: numpy arrays (host) or torch CUDA tensors (device pointers, no copies)."""
        if isinstance(evidence, np.ndarray):
            ev = np.ascontiguousarray(evidence, np.int8)
            assert ev.ndim == 2 and ev.shape[1] == self.net.n
            if out is None:
                out = np.empty((ev.shape[0], width), np.float64)
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (ev.shape[0], width)
            check(fn(self._h, ev.shape[0], ev.ctypes.data, out.ctypes.data, HB_HOST_PTRS))
            return out
        import torch  # device memory + streams only

        assert evidence.dtype == torch.int8 and evidence.is_contiguous() and evidence.shape[1] == self.net.n
        if evidence.is_cuda:
            if out is None:
                out = torch.empty((evidence.shape[0], width), dtype=torch.float64, device=evidence.device)
            assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
            check(fn(self._h, evidence.shape[0], evidence.data_ptr(), out.data_ptr(), HB_DEVICE_PTRS))
        else:  # (pinned) host tensors
            if out is None:
                out = torch.empty((evidence.shape[0], width), dtype=torch.float64, pin_memory=evidence.is_pinned())
            check(fn(self._h, evidence.shape[0], evidence.data_ptr(), out.data_ptr(), HB_HOST_PTRS))
        return out


class JunctionTree(_Batched):
    """createJunctionTree's result.  `device=None` builds the structure only (no GPU needed, no compute possible)."""

    _prefix = "hb_jt_"

    def __init__(self, net: BayesianNetwork, cmp=nodeComparisonForTriangulation, order=None, device=0, split=None, nccl=None):
        self.net = net
        dp, nd, self._keep = _device_args(device)
        h = C.c_void_p()
        if nccl is not None:  # (rank, world, unique id bytes): the slice `split` of a clique split over `world` GPUs, exchange inside the library
            rank, world, uid = nccl
            va, vp_ = _i32([e[0] for e in split])
            sa, sp = _i32([e[1] for e in split])
            buf = C.create_string_buffer(bytes(uid), 128) if uid is not None else None
            check(lib().hb_jt_compile_split_nccl(net._h, int(cmp), HB_STRUCTURE_ONLY if device is None else int(device), len(split), vp_, sp,
                                                 int(rank), int(world), buf, C.byref(h)))
        elif split is not None:  # one slice of a split clique: [(vertex, state)], no calibration
            va, vp_ = _i32([e[0] for e in split])
            sa, sp = _i32([e[1] for e in split])
            check(lib().hb_jt_compile_split(net._h, int(cmp), dp, nd, len(split), vp_, sp, C.byref(h)))
        elif order is not None:
            oa, op = _i32(order)
            check(lib().hb_jt_compile_with_order(net._h, op, dp, nd, C.byref(h)))
        else:
            check(lib().hb_jt_compile(net._h, int(cmp), dp, nd, C.byref(h)))
        self._h = h

    @property
    def n_devices(self):
        return int(lib().hb_jt_n_devices(self._h))

    def describe(self):
        p = C.c_void_p()
        check(lib().hb_jt_describe_json(self._h, C.byref(p)))
        return _take_json(p)

    def save(self, path):
        """network + tree structure to a binary file (cf. Bayes/ImportExport.hs:36-46)"""
        check(lib().hb_jt_save(self._h, str(path).encode()))

    @staticmethod
    def load(path, net: BayesianNetwork, device=0):
        """`net` must be the network the tree was built from (it supplies names / dims to the Python side)"""
        dp, nd, keep = _device_args(device)
        h = C.c_void_p()
        check(lib().hb_jt_load(str(path).encode(), dp, nd, C.byref(h)))
        jt = JunctionTree.__new__(JunctionTree)
        jt.net, jt._keep, jt._h = net, keep, h
        return jt

    def program(self, observed: Iterable[int] = ()):
        """parity hook: operation trace + physical steps compiled for changeEvidence on `observed` (in that order)"""
        oa, op = _i32(list(observed))
        p = C.c_void_p()
        check(lib().hb_jt_program_json(self._h, len(oa), op, C.byref(p)))
        return _take_json(p)

    def specialise(self, observed: Iterable[int] = ()) -> int:
        """build the specialised kernels for this observed list now (seconds) rather than inside the first large batch"""
        oa, op = _i32(list(observed))
        n = C.c_int32()
        check(lib().hb_jt_specialise(self._h, len(oa), op, C.byref(n)))
        return n.value

    def specialised_kernels(self) -> int:
        """steps of the last batch call's program that run on a kernel specialised for them (csrc/jit.cpp)"""
        n = C.c_int32()
        check(lib().hb_jt_specialised_kernels(self._h, C.byref(n)))
        return n.value

    def jit_source(self, observed: Iterable[int] = ()) -> str:
        """CUDA source of the per-step specialised kernels for this observed list (inspection hook, csrc/jit.cpp)"""
        oa, op = _i32(list(observed))
        p = C.c_void_p()
        check(lib().hb_jt_jit_source(self._h, len(oa), op, C.byref(p)))
        s = C.cast(p, C.c_char_p).value.decode()
        lib().hb_string_free(p)
        return s

    def kernel_info(self):
        """what the program of the last batch / specialise call runs on (hb_jt_kernel_info)"""
        out = np.zeros(12, np.int64)
        check(lib().hb_jt_kernel_info(self._h, out.ctypes.data_as(i64p)))
        keys = ["on_chip_eligible", "rows_per_cta", "threads_per_cta", "levels", "smem_entries_per_row", "smem_bytes",
                "fp64_mul_per_row", "fp64_add_per_row", "smem_loads_per_row", "smem_stores_per_row", "on_chip", "origin"]
        return dict(zip(keys, out.tolist()))

    def tree_profile(self):
        """SM cycles per phase of the on-chip kernel since the last call (HB_TREE_PROFILE=1): list of 64 counters"""
        out = np.zeros(64, np.int64)
        check(lib().hb_jt_tree_profile(self._h, out.ctypes.data_as(i64p)))
        return out.tolist()

    def tree_source(self, observed: Iterable[int] = ()) -> str:
        """CUDA source of the whole-schedule on-chip kernel for this observed list (raises HbError when there is none)"""
        oa, op = _i32(list(observed))
        p = C.c_void_p()
        check(lib().hb_jt_tree_source(self._h, len(oa), op, C.byref(p)))
        s = C.cast(p, C.c_char_p).value.decode()
        lib().hb_string_free(p)
        return s

    def changeEvidence(self, evidence: Sequence[tuple]):
        """evidence = [(vertex, state)], order kept (it decides the reference's operation order)."""
        n = len(evidence)
        vs = (C.c_int32 * max(n, 1))(*[int(e[0]) for e in evidence])
        ss = (C.c_int32 * max(n, 1))(*[int(e[1]) for e in evidence])
        check(lib().hb_jt_set_evidence(self._h, n, vs, ss))
        return self

    def posterior(self, variables: Sequence[int]):
        if len(variables) == 1:  # the common case (posterior jt [v]): no numpy round trips on the latency path
            v = int(variables[0])
            if not 0 <= v < self.net.n:
                raise HbError(_lib.HB_E_ARG, "query on unknown vertex")
            size = self.net.dims[v]
            out = (C.c_double * size)()
            q, scope, n = (C.c_int32 * 1)(v), (C.c_int32 * 1)(), C.c_int64()
            code = lib().hb_jt_posterior(self._h, 1, q, scope, out, size, C.byref(n))
            if code == HB_E_NO_CLUSTER:
                return None
            check(code)
            return Factor([v], [size], np.frombuffer(out, np.float64, n.value))
        qa, qp = _i32(list(variables))
        size = int(np.prod([self.net.dims[v] for v in variables]))
        out = np.zeros(size, np.float64)
        scope = np.zeros(len(variables), np.int32)
        n = C.c_int64()
        code = lib().hb_jt_posterior(self._h, len(variables), qp, scope.ctypes.data_as(i32p), out.ctypes.data_as(f64p), size,
                                     C.byref(n))
        if code == HB_E_NO_CLUSTER:
            return None  # the reference's `Nothing` (FactorElimination.hs:319)
        check(code)
        sc = scope.tolist()
        return Factor(sc, [self.net.dims[v] for v in sc], out[: n.value])

    def posteriors_batch(self, evidence, out=None):
        """One query per row: changeEvidence row; posterior [v] for all v.  Returns [n_rows][sum dims]."""
        return self._run_matrix(lib().hb_jt_posteriors_batch, evidence, out, self.net.width)

    def queries_batch(self, queries: Sequence[Sequence[int]], evidence: np.ndarray):
        """batched `posterior jt q` for a list of queries; returns [(Factor-like scope, dims, values [rows][size])]"""
        flat, off = [], [0]
        for q in queries:
            flat += list(q)
            off.append(len(flat))
        fa, fp = _i32(flat)
        oa, op = _i32(off)
        width = C.c_int64()
        check(lib().hb_jt_queries_width(self._h, len(queries), op, fp, C.byref(width)))
        ev = np.ascontiguousarray(evidence, np.int8)
        out = np.empty((ev.shape[0], width.value), np.float64)
        scopes = np.zeros(len(flat), np.int32)
        check(lib().hb_jt_queries_batch(self._h, len(queries), op, fp, ev.shape[0], ev.ctypes.data, out.ctypes.data, HB_HOST_PTRS,
                                        scopes.ctypes.data_as(i32p)))
        res, col = [], 0
        for i, q in enumerate(queries):
            sc = scopes[off[i]:off[i + 1]].tolist()
            dims = [self.net.dims[v] for v in sc]
            size = int(np.prod(dims))
            res.append((sc, dims, out[:, col:col + size]))
            col += size
        return res

    def change_factor(self, factor: "Factor") -> bool:
        """changeFactor f jt (Bayes/FactorElimination/JTree.hs:107-115): new values for the CPT with the factor's variable
        list, evidence kept, tree recalibrated; the compiled schedules stay (a potential upload).  True if a CPT matched."""
        fa, fp = _i32(list(factor.variables))
        vals = np.ascontiguousarray(factor.values, np.float64).reshape(-1)
        hit = C.c_int32()
        check(lib().hb_jt_change_factor(self._h, len(fa), fp, vals.ctypes.data_as(f64p), C.byref(hit)))
        if hit.value:
            self.net = self.net.change_factor(factor)
        return bool(hit.value)

    def em_step(self, samples):
        """learnEM's expectation + maximisation on this tree's network (Bayes/EM.hs:36-66), samples = int8 evidence matrix
        (-1 = missing; rows may have different observed sets) on the host or a CUDA tensor.  Returns, per vertex, the learnt
        table as the reference lays it out: Factor over [parents in the order of their posterior..., child]."""
        if isinstance(samples, np.ndarray):
            ev = np.ascontiguousarray(samples, np.int8)
            ptr, flags, rows = ev.ctypes.data, HB_HOST_PTRS, ev.shape[0]
        else:
            assert samples.is_cuda and samples.is_contiguous() and samples.element_size() == 1
            ptr, flags, rows = samples.data_ptr(), HB_DEVICE_PTRS, samples.shape[0]
        total = sum(int(np.prod([self.net.dims[x] for x in sc])) for sc in self.net.scopes)
        out = np.zeros(total, np.float64)
        scopes = np.zeros(sum(len(sc) for sc in self.net.scopes), np.int32)
        check(lib().hb_jt_em_step(self._h, rows, ptr, out.ctypes.data_as(f64p), scopes.ctypes.data_as(i32p), flags))
        res, at, sat = [], 0, 0
        for sc0 in self.net.scopes:
            sc = scopes[sat:sat + len(sc0)].tolist()
            dims = [self.net.dims[x] for x in sc]
            size = int(np.prod(dims))
            res.append(Factor(sc, dims, out[at:at + size].copy()))
            at, sat = at + size, sat + len(sc0)
        return res

    def set_split(self, fixed: Sequence[tuple]):
        """split clique: this handle owns the slice `fixed` = [(vertex, state)]; batch results become unnormalised"""
        va, vp_ = _i32([e[0] for e in fixed])
        sa, sp = _i32([e[1] for e in fixed])
        check(lib().hb_jt_set_split(self._h, len(fixed), vp_, sp))

    def normalise_batch(self, out):
        """in place: per row and per variable, x * (1.0 / sum x)  (after the slices' results have been summed)"""
        if isinstance(out, np.ndarray):
            assert out.dtype == np.float64 and out.flags.c_contiguous
            check(lib().hb_jt_normalise_batch(self._h, out.shape[0], out.ctypes.data, HB_HOST_PTRS))
        else:
            assert out.is_cuda and out.is_contiguous()
            check(lib().hb_jt_normalise_batch(self._h, out.shape[0], out.data_ptr(), HB_DEVICE_PTRS))
        return out

    def __del__(self):
        try:
            lib().hb_jt_free(self._h)
        except Exception:
            pass


class VariableElimination(_Batched):
    """A (p, r) pair of posteriorMarginal compiled for replay over evidence rows."""

    _prefix = "hb_ve_"

    def __init__(self, net: BayesianNetwork, p: Sequence[int], r: Sequence[int], device=0):
        self.net, self.p, self.r = net, list(p), list(r)
        dp, nd, self._keep = _device_args(device)
        pa, pp = _i32(self.p)
        ra, rp = _i32(self.r)
        h = C.c_void_p()
        check(lib().hb_ve_compile(net._h, pp, len(self.p), rp, len(self.r), dp, nd, C.byref(h)))
        self._h = h
        self.width = int(np.prod([net.dims[v] for v in self.r]))

    def program(self, observed: Iterable[int] = ()):
        oa, op = _i32(list(observed))
        p = C.c_void_p()
        check(lib().hb_ve_program_json(self._h, len(oa), op, C.byref(p)))
        return _take_json(p)

    def posterior(self, assignment: Sequence[tuple] = ()):
        va, vp_ = _i32([e[0] for e in assignment])
        sa, sp = _i32([e[1] for e in assignment])
        out = np.zeros(self.width, np.float64)
        scope = np.zeros(len(self.r), np.int32)
        n = C.c_int64()
        check(lib().hb_ve_posterior(self._h, len(assignment), vp_, sp, scope.ctypes.data_as(i32p), out.ctypes.data_as(f64p),
                                    self.width, C.byref(n)))
        sc = scope.tolist()
        return Factor(sc, [self.net.dims[v] for v in sc], out[: n.value])

    def specialise(self, observed: Iterable[int] = ()) -> int:
        """build the kernels generated for this observed list now rather than inside the first large batch"""
        oa, op = _i32(list(observed))
        n = C.c_int32()
        check(lib().hb_ve_specialise(self._h, len(oa), op, C.byref(n)))
        return n.value

    def posterior_batch(self, evidence, out=None):
        return self._run_matrix(lib().hb_ve_posterior_batch, evidence, out, self.width)

    def change_factor(self, factor: "Factor") -> bool:
        """changeFactor on the network behind this handle (Bayes/Factor.hs:66-81); True if a CPT matched"""
        fa, fp = _i32(list(factor.variables))
        vals = np.ascontiguousarray(factor.values, np.float64).reshape(-1)
        hit = C.c_int32()
        check(lib().hb_ve_change_factor(self._h, len(fa), fp, vals.ctypes.data_as(f64p), C.byref(hit)))
        if hit.value:
            self.net = self.net.change_factor(factor)
        return bool(hit.value)

    def __del__(self):
        try:
            lib().hb_ve_free(self._h)
        except Exception:
            pass


class MostProbableExplanation(_Batched):
    """`mpe g p r` (Bayes/VariableElimination.hs:99-114) compiled for replay: p summed out, r maximised (MAXCPT)."""

    _prefix = "hb_ve_"

    def __init__(self, net: BayesianNetwork, p: Sequence[int], r: Sequence[int], device=0):
        self.net, self.p, self.r = net, list(p), list(r)
        dp, nd, self._keep = _device_args(device)
        pa, pp = _i32(self.p)
        ra, rp = _i32(self.r)
        h = C.c_void_p()
        check(lib().hb_ve_compile_mpe(net._h, pp, len(self.p), rp, len(self.r), dp, nd, C.byref(h)))
        self._h = h

    def program(self, observed: Iterable[int] = ()):
        oa, op = _i32(list(observed))
        p = C.c_void_p()
        check(lib().hb_ve_program_json(self._h, len(oa), op, C.byref(p)))
        return _take_json(p)

    def order(self, observed: Iterable[int] = ()):
        """variable order of the reference's instantiation list for this observed list"""
        oa, op = _i32(list(observed))
        out = np.zeros(len(self.r), np.int32)
        n = C.c_int32()
        check(lib().hb_ve_mpe_order(self._h, len(oa), op, out.ctypes.data_as(i32p), C.byref(n)))
        return out[: n.value].tolist()

    def explain(self, assignment: Sequence[tuple] = ()):
        """-> (P(instantiation, evidence), [[(vertex, state)]]) with the instantiation list as `mpe` returns it"""
        va, vp_ = _i32([e[0] for e in assignment])
        sa, sp = _i32([e[1] for e in assignment])
        states = np.zeros(len(self.r), np.int8)
        value = C.c_double()
        order = np.zeros(len(self.r), np.int32)
        n = C.c_int32()
        check(lib().hb_ve_mpe(self._h, len(assignment), vp_, sp, states.ctypes.data, C.byref(value), order.ctypes.data_as(i32p),
                              C.byref(n)))
        by_var = dict(zip(self.r, states.tolist()))
        inst = [(v, by_var[v]) for v in order[: n.value].tolist()]
        return value.value, ([inst] if inst else [])

    def explain_batch(self, evidence, states=None, values=None):
        """one mpe per evidence row: (int8 [rows][len(r)] states in the order of r, float64 [rows] values).
        NumPy arrays, or CUDA tensors (all rows then share one observed set; results stay on the device)."""
        if isinstance(evidence, np.ndarray):
            ev = np.ascontiguousarray(evidence, np.int8)
            states = np.empty((ev.shape[0], len(self.r)), np.int8) if states is None else states
            values = np.empty(ev.shape[0], np.float64) if values is None else values
            check(lib().hb_ve_mpe_batch(self._h, ev.shape[0], ev.ctypes.data, states.ctypes.data, values.ctypes.data, HB_HOST_PTRS))
            return states, values
        import torch

        assert evidence.is_cuda and evidence.is_contiguous() and evidence.element_size() == 1
        rows = evidence.shape[0]
        states = torch.empty((rows, len(self.r)), dtype=torch.int8, device=evidence.device) if states is None else states
        values = torch.empty(rows, dtype=torch.float64, device=evidence.device) if values is None else values
        check(lib().hb_ve_mpe_batch(self._h, rows, evidence.data_ptr(), states.data_ptr(), values.data_ptr(), HB_DEVICE_PTRS))
        return states, values

    def __del__(self):
        try:
            lib().hb_ve_free(self._h)
        except Exception:
            pass


# ---- the reference's function names -----------------------------------------------------------


def createJunctionTree(cmp, net: BayesianNetwork, device=0) -> JunctionTree:
    return JunctionTree(net, cmp=cmp, device=device)


def changeEvidence(evidence, jt: JunctionTree) -> JunctionTree:
    return jt.changeEvidence(evidence)


def posterior(jt: JunctionTree, variables):
    return jt.posterior(variables)


def posteriorMarginal(net: BayesianNetwork, p, r, assignment, device=0) -> Factor:
    return VariableElimination(net, p, r, device=device).posterior(assignment)


def mpe(net: BayesianNetwork, p, r, assignment, device=0):
    """`mpe g someP someR assignment` (Bayes/VariableElimination.hs:101-114): Most Probable Explanation, or MAP when r is a
    subset.  Returns the reference's [DVISet]: a list with one instantiation [(vertex, state), ...] in its order."""
    return MostProbableExplanation(net, p, r, device=device).explain(assignment)[1]


def priorMarginal(net: BayesianNetwork, ea, eb, device=0) -> Factor:
    return VariableElimination(net, ea, eb, device=device).posterior([])


# ---- class Factor's product / projectOut (Bayes/Factor.hs:113-188) on free-standing factors --------


def _factor_marginalize(factors: Sequence[Factor], sum_var: int, device=0) -> Factor:
    dims = {}
    for f in factors:
        for v, d in zip(f.variables, f.dims):
            dims[v] = d
    n_vars = (max(dims) + 1) if dims else 1
    dim_arr = np.ones(n_vars, np.int32)
    for v, d in dims.items():
        dim_arr[v] = d
    scope_off = np.zeros(len(factors) + 1, np.int32)
    val_off = np.zeros(len(factors) + 1, np.int64)
    for i, f in enumerate(factors):
        scope_off[i + 1] = scope_off[i] + len(f.variables)
        val_off[i + 1] = val_off[i] + len(f.values)
    sa, sp = _i32(np.concatenate([np.asarray(f.variables, np.int32) for f in factors] + [np.zeros(0, np.int32)]))
    vals = np.ascontiguousarray(np.concatenate([f.values for f in factors] + [np.zeros(1)]))
    size = int(np.prod([d for v, d in dims.items() if v != sum_var])) if dims else 1
    out = np.zeros(max(size, 1), np.float64)
    scope = np.zeros(max(len(dims), 1), np.int32)
    ns, nw = C.c_int32(), C.c_int64()
    check(lib().hb_factor_marginalize(n_vars, dim_arr.ctypes.data_as(i32p), len(factors), scope_off.ctypes.data_as(i32p), sp,
                                      val_off.ctypes.data_as(i64p), vals.ctypes.data_as(f64p), sum_var, device,
                                      scope.ctypes.data_as(i32p), C.byref(ns), out.ctypes.data_as(f64p), len(out), C.byref(nw)))
    sc = scope[: ns.value].tolist()
    return Factor(sc, [dims[v] for v in sc], out[: nw.value].copy())


def factorProduct(factors: Sequence[Factor], device=0) -> Factor:
    """factorProduct (Bayes/Factor/CPT.hs:138): scope = foldr1 union, value = ps * (v1 * (v2 * ...))"""
    return _factor_marginalize(factors, -1, device)


def factorProjectOut(variables: Sequence[int], factor: Factor, device=0) -> Factor:
    """factorProjectOut (CPT.hs:140-141).  One variable is the reference operation bit for bit (the only form the
    hot path uses, Buckets.hs:109); several variables are summed one after the other, which reassociates the sum."""
    for v in variables:
        factor = _factor_marginalize([factor], v, device)
    return factor


def factorProjectTo(variables: Sequence[int], factor: Factor, device=0) -> Factor:
    """factorProjectTo s f = factorProjectOut (factorVariables f `difference` s) f  (Bayes/Factor.hs:183-188)"""
    return factorProjectOut([v for v in factor.variables if v not in set(variables)], factor, device)


def marginalizeOneVariable(factors: Sequence[Factor], variable: int, device=0) -> Factor:
    """Buckets.hs:105-111: product of a bucket, then sum out `variable` -- one fused kernel launch"""
    return _factor_marginalize(factors, variable, device)


def changeFactor(factor: Factor, container):
    """`changeFactor f c` of the FactorContainer class (Bayes/Factor.hs:76-81): a network gives a NEW network; a
    JunctionTree / VariableElimination handle is updated in place (and returned), its schedules untouched."""
    if isinstance(container, BayesianNetwork):
        return container.change_factor(factor)
    container.change_factor(factor)
    return container


def learnEM(samples, net: BayesianNetwork, device=0, reference_layout=False):
    """`learnEM samples g` (Bayes/EM.hs:36-45), one expectation/maximisation step, entirely on the GPU: per sample
    changeEvidence + `posterior` of every CPT family and of its parents, summed sample after sample and divided with
    `divide _ 0 = 0` -- bit-identical to the reference's fold.  `samples`: int8 [n_samples][n_vars] evidence matrix
    (-1 = missing; the observed set may differ from row to row).
    Returns the learnt CPT values per vertex in the CPT's own layout [child, parents...] (a pure re-layout of what the
    reference returns), or with reference_layout=True the reference's own Factors, whose variable order is
    [parents in the order of their posterior..., child] (cptDivide, Bayes/Factor/CPT.hs:159-160)."""
    learnt = JunctionTree(net, device=device).em_step(samples)
    if reference_layout:
        return learnt
    out = []
    for v, f in enumerate(learnt):
        perm = [f.variables.index(x) for x in net.scopes[v]]
        out.append(np.ascontiguousarray(f.values.reshape(f.dims).transpose(perm)).reshape(-1))
    return out
