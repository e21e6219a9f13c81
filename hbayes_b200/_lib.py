"""ctypes binding of libhbayes_cuda.so (include/hbayes_cuda.h).  Plumbing only: the product is the library.

The library must have been built in-tree (`make -C hbayes_b200/csrc` or `__graft_entry__.build()`); there is no
fallback of any kind when it is missing.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HB_LIB") or os.path.join(HERE, "libhbayes_cuda.so")  # HB_LIB: A/B builds of the same library

HB_OK, HB_E_ARG, HB_E_NO_CLUSTER, HB_E_CUDA, HB_E_OOM, HB_E_IO, HB_E_UNSUPPORTED, HB_E_NCCL = 0, -1, -2, -3, -4, -5, -6, -7
HB_STRUCTURE_ONLY = -1
HB_HOST_PTRS, HB_DEVICE_PTRS = 0, 1

vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
i32p, i64p, i8p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int8), C.POINTER(C.c_double)
vpp, cpp = C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)

# every symbol include/hbayes_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hb_last_error": (C.c_char_p, []),
    "hb_string_free": (None, [vp]),
    "hb_device_count": (C.c_int, []),
    "hb_host_alloc": (C.c_int, [i64, vpp]),
    "hb_host_free": (None, [vp]),
    "hb_host_register": (C.c_int, [vp, i64]),
    "hb_host_unregister": (C.c_int, [vp]),
    "hb_jt_n_devices": (C.c_int, [vp]),
    "hb_net_create": (C.c_int, [i32, i32p, i64p, i32p, i64p, f64p, vpp]),
    "hb_net_create_procedural": (C.c_int, [i32, i32p, i64p, i32p, i64p, f64p, C.POINTER(C.c_uint64), vpp]),
    "hb_net_cpt_values": (C.c_int, [vp, i32, i64, i64, f64p]),
    "hb_net_load_hugin": (C.c_int, [C.c_char_p, vpp]),
    "hb_net_n_vars": (C.c_int, [vp]),
    "hb_net_dims": (C.c_int, [vp, i32p]),
    "hb_net_describe_json": (C.c_int, [vp, cpp]),
    "hb_net_free": (None, [vp]),
    "hb_ve_order": (C.c_int, [vp, i32, i32p, i32p]),
    "hb_ve_degree_order": (C.c_int, [vp, i32p, i32, i32p]),
    "hb_jt_compile": (C.c_int, [vp, i32, i32p, i32, vpp]),
    "hb_jt_compile_with_order": (C.c_int, [vp, i32p, i32p, i32, vpp]),
    "hb_jt_compile_split": (C.c_int, [vp, i32, i32p, i32, i32, i32p, i32p, vpp]),
    "hb_nccl_unique_id": (C.c_int, [vp]),
    "hb_jt_compile_split_nccl": (C.c_int, [vp, i32, i32, i32, i32p, i32p, i32, i32, vp, vpp]),
    "hb_jt_save": (C.c_int, [vp, C.c_char_p]),
    "hb_jt_load": (C.c_int, [C.c_char_p, i32p, i32, vpp]),
    "hb_jt_describe_json": (C.c_int, [vp, cpp]),
    "hb_jt_program_json": (C.c_int, [vp, i32, i32p, cpp]),
    "hb_jt_set_evidence": (C.c_int, [vp, i32, i32p, i32p]),
    "hb_jt_posterior": (C.c_int, [vp, i32, i32p, i32p, f64p, i64, i64p]),
    "hb_jt_posteriors_batch": (C.c_int, [vp, i64, vp, vp, i32]),
    "hb_jt_queries_batch": (C.c_int, [vp, i32, i32p, i32p, i64, vp, vp, i32, i32p]),
    "hb_jt_queries_width": (C.c_int, [vp, i32, i32p, i32p, i64p]),
    "hb_jt_jit_source": (C.c_int, [vp, i32, i32p, cpp]),
    "hb_jt_tree_source": (C.c_int, [vp, i32, i32p, cpp]),
    "hb_jt_kernel_info": (C.c_int, [vp, i64p]),
    "hb_jt_tree_profile": (C.c_int, [vp, i64p]),
    "hb_jt_specialise": (C.c_int, [vp, i32, i32p, i32p]),
    "hb_jt_specialised_kernels": (C.c_int, [vp, i32p]),
    "hb_jt_change_factor": (C.c_int, [vp, i32, i32p, f64p, i32p]),
    "hb_jt_em_step": (C.c_int, [vp, i64, vp, f64p, i32p, i32]),
    "hb_jt_set_split": (C.c_int, [vp, i32, i32p, i32p]),
    "hb_jt_normalise_batch": (C.c_int, [vp, i64, vp, i32]),
    "hb_jt_set_precision": (C.c_int, [vp, i32]),
    "hb_jt_set_stream": (C.c_int, [vp, vp]),
    "hb_jt_set_workspace": (C.c_int, [vp, i64]),
    "hb_jt_last_stats": (C.c_int, [vp, i64p]),
    "hb_jt_set_timing": (C.c_int, [vp, i32]),
    "hb_jt_last_timing": (C.c_int, [vp, f64p]),
    "hb_jt_free": (None, [vp]),
    "hb_factor_marginalize": (C.c_int, [i32, i32p, i32, i32p, i32p, i64p, f64p, i32, i32, i32p, i32p, f64p, i64, i64p]),
    "hb_ve_compile": (C.c_int, [vp, i32p, i32, i32p, i32, i32p, i32, vpp]),
    "hb_ve_result_scope": (C.c_int, [vp, i32, i32p, i32p, i32p]),
    "hb_ve_program_json": (C.c_int, [vp, i32, i32p, cpp]),
    "hb_ve_posterior": (C.c_int, [vp, i32, i32p, i32p, i32p, f64p, i64, i64p]),
    "hb_ve_posterior_batch": (C.c_int, [vp, i64, vp, vp, i32]),
    "hb_ve_compile_mpe": (C.c_int, [vp, i32p, i32, i32p, i32, i32p, i32, vpp]),
    "hb_ve_mpe": (C.c_int, [vp, i32, i32p, i32p, vp, f64p, i32p, i32p]),
    "hb_ve_mpe_batch": (C.c_int, [vp, i64, vp, vp, vp, i32]),
    "hb_ve_mpe_order": (C.c_int, [vp, i32, i32p, i32p, i32p]),
    "hb_ve_change_factor": (C.c_int, [vp, i32, i32p, f64p, i32p]),
    "hb_ve_specialise": (C.c_int, [vp, i32, i32p, i32p]),
    "hb_ve_set_precision": (C.c_int, [vp, i32]),
    "hb_ve_set_stream": (C.c_int, [vp, vp]),
    "hb_ve_set_workspace": (C.c_int, [vp, i64]),
    "hb_ve_last_stats": (C.c_int, [vp, i64p]),
    "hb_ve_set_timing": (C.c_int, [vp, i32]),
    "hb_ve_last_timing": (C.c_int, [vp, f64p]),
    "hb_ve_free": (None, [vp]),
}

_LIB = None


class HbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libhbayes_cuda error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads the library or raises: a missing build is an error, never a reason to compute elsewhere."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `make -C hbayes_b200/csrc` (or __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(code):
    if code != HB_OK:
        raise HbError(code, lib().hb_last_error().decode())
