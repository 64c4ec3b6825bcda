"""ctypes binding of libhcurl_b200.so (the C ABI in include/hcurl_b200.h).

There is no CPU fallback: if the shared library is missing this module raises
at import of the symbols, and every compute entry point raises HcurlError when
the library reports a failure (e.g. no CUDA device).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HCURL_LIB_PATH") or os.path.join(_HERE, "lib", "libhcurl_b200.so")  # override: diagnostic builds only


class HcurlError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I8 = C.POINTER(C.c_int8)

# name -> (restype, argtypes); mirrors include/hcurl_b200.h one to one
SIGNATURES = {
    "hcurl_version": (C.c_char_p, []),
    "hcurl_last_error": (C.c_char_p, []),
    "hcurl_comm_unique_id": (C.c_int, [_P]),
    "hcurl_ctx_create": (_P, [C.c_int, _P, C.c_int, C.c_int]),
    "hcurl_ctx_destroy": (None, [_P]),
    "hcurl_ctx_device": (C.c_int, [_P]),
    "hcurl_ctx_synchronize": (C.c_int, [_P]),
    "hcurl_ctx_mark": (C.c_int, [_P, C.c_int]),
    "hcurl_ctx_elapsed_ms": (C.c_double, [_P]),
    "hcurl_ctx_launch_count": (C.c_int64, [_P]),
    "hcurl_space_create": (_P, [_P, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, _P, _P, C.c_int32, _P, C.c_int32,
                                C.c_int, _P, _P, _P, _P, _P, _P]),
    "hcurl_space_ext_size": (C.c_int64, [_P]),
    "hcurl_space_destroy": (None, [_P]),
    "hcurl_space_ndof_local": (C.c_int32, [_P]),
    "hcurl_space_ndof_true": (C.c_int32, [_P]),
    "hcurl_space_elem_ndof": (C.c_int, [_P]),
    "hcurl_op_create": (_P, [_P, _P, C.c_int]),
    "hcurl_op_destroy": (None, [_P]),
    "hcurl_op_apply_batch": (C.c_int, [_P, C.c_int, _P, _P, C.c_int]),
    "hcurl_op_set_kernel": (C.c_int, [_P, C.c_int]),
    "hcurl_op_last_kernel": (C.c_int, [_P]),
    "hcurl_op_set_tet_rules": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P]),
    "hcurl_op_qpts_curl": (C.c_int, [_P]),
    "hcurl_op_qpts_mass": (C.c_int, [_P]),
    "hcurl_op_set_coefficients": (C.c_int, [_P, C.c_int, _P, _P, C.c_int]),
    "hcurl_op_set_essential_tdofs": (C.c_int, [_P, _P, C.c_int32]),
    "hcurl_op_apply": (C.c_int, [_P, _P, _P, C.c_int]),
    "hcurl_op_apply_unconstrained": (C.c_int, [_P, _P, _P, C.c_double, C.c_int, C.c_int]),
    "hcurl_op_apply_complex": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int]),
    "hcurl_op_diagonal": (C.c_int, [_P, _P, C.c_int]),
    "hcurl_op_form_rhs": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "hcurl_op_last_kernel_ms": (C.c_double, [_P]),
    "hcurl_op_kernel_stats_reset": (C.c_int, [_P]),
    "hcurl_op_kernel_stats_ms": (C.c_double, [_P, C.POINTER(C.c_int)]),
    "hcurl_pcg_create": (_P, [_P, _P, C.c_int]),
    "hcurl_pcg_destroy": (None, [_P]),
    "hcurl_pcg_set": (C.c_int, [_P, C.c_double, C.c_double, C.c_int, C.c_int]),
    "hcurl_pcg_update": (C.c_int, [_P]),
    "hcurl_pcg_solve": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                  C.POINTER(C.c_int)]),
    "hcurl_pcg_last_loop_ms": (C.c_double, [_P]),
    "hcurl_pcg_last_used_graph": (C.c_int, [_P]),
    "hcurl_pcg_set_chebyshev": (C.c_int, [_P, C.c_int, C.c_int]),
    "hcurl_pcg_lambda_max": (C.c_double, [_P]),
    "hcurl_interp_create": (_P, [_P, C.c_int, C.c_int, C.c_int64, _P, _P, C.c_int32, C.c_int32]),
    "hcurl_interp_create_geom": (_P, [_P, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, C.c_int32, C.c_int32]),
    "hcurl_host_interp_tet_matrix": (C.c_int, [C.c_int, _P]),
    "hcurl_interp_destroy": (None, [_P]),
    "hcurl_interp_apply": (C.c_int, [_P, _P, _P, C.c_int]),
    "hcurl_host_interp_matrix": (C.c_int, [C.c_int, _P]),
    "hcurl_host_interp_encode": (C.c_int, [C.c_int, C.c_int, C.c_int64, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "hcurl_device_alloc": (_P, [_P, C.c_int64]),
    "hcurl_device_free": (None, [_P, _P]),
    "hcurl_copy_to_device": (C.c_int, [_P, _P, _P, C.c_int64]),
    "hcurl_copy_to_host": (C.c_int, [_P, _P, _P, C.c_int64]),
    "hcurl_host_alloc_pinned": (_P, [C.c_int64]),
    "hcurl_host_free_pinned": (None, [_P]),
    "hcurl_dot": (C.c_int, [_P, _P, _P, C.c_int64, C.POINTER(C.c_double)]),
    "hcurl_fp64_peak_tflops": (C.c_double, [_P]),
    "hcurl_host_exodus_open": (_P, [C.c_char_p]),
    "hcurl_host_exodus_close": (None, [_P]),
    "hcurl_host_exodus_counts": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "hcurl_host_exodus_read": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "hcurl_host_ndspace_create_hex": (_P, [C.c_int, C.c_int64, _P]),
    "hcurl_host_ndspace_create_tet": (_P, [C.c_int, C.c_int64, _P]),
    "hcurl_host_ndspace_elem_face_base": (C.c_int, [_P, _P]),
    "hcurl_host_ndspace_destroy": (None, [_P]),
    "hcurl_host_ndspace_nd": (C.c_int, [_P]),
    "hcurl_host_ndspace_ndof": (C.c_int64, [_P]),
    "hcurl_host_ndspace_n_edges": (C.c_int64, [_P]),
    "hcurl_host_ndspace_n_faces": (C.c_int64, [_P]),
    "hcurl_host_ndspace_elem_dofs": (C.c_int, [_P, _P]),
    "hcurl_host_ndspace_boundary_dofs": (C.c_int64, [_P, _P]),
    "hcurl_host_ndspace_face_dofs": (C.c_int64, [_P, C.c_int64, _P, _P]),
    "hcurl_host_hex_p2_pair_codes": (C.c_int, [C.c_int64, _P, _P]),
    "hcurl_host_partition_create": (_P, [C.c_int64, C.c_int, C.c_int64, _P, _P, C.c_int]),
    "hcurl_host_partition_destroy": (None, [_P]),
    "hcurl_host_partition_sizes": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int),
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hcurl_host_partition_copy": (C.c_int, [_P] * 10),
    "hcurl_host_patches_parts": (C.c_int, [_P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hcurl_host_patches_owned": (C.c_int64, [_P, _P, _P]),
    "hcurl_host_modal_tables_p2": (C.c_int, [_P]),
    "hcurl_host_hex_slot_codes": (C.c_int, [C.c_int, C.c_int64, _P, _P]),
    "hcurl_host_hex_dof_nodes": (C.c_int, [C.c_int, _P, _P]),
    "hcurl_host_patches_create": (_P, [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, _P]),
    "hcurl_host_patches_destroy": (None, [_P]),
    "hcurl_host_patches_count": (C.c_int64, [_P]),
    "hcurl_host_patches_boundary_count": (C.c_int64, [_P]),
    "hcurl_host_patches_blob_words": (C.c_int64, [_P]),
    "hcurl_host_patches_max_blob_words": (C.c_int64, [_P]),
    "hcurl_host_patches_max_dofs": (C.c_int64, [_P]),
    "hcurl_host_patches_swizzle": (C.c_int, [_P, _P, _P]),
    "hcurl_host_patches_copy": (C.c_int, [_P, _P, _P, _P]),
}

HEX, TET = 0, 1
PER_ELEM, PER_QPT = 0, 1
PC_NONE, PC_JACOBI, PC_CHEBYSHEV = 0, 1, 2
HOST, DEVICE, DEVICE_EXT = 0, 1, 2


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HcurlError(
                "libhcurl_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().hcurl_last_error().decode()


def check(rc, what):
    if rc != 0:
        raise HcurlError("%s failed: %s" % (what, last_error()))


def check_ptr(p, what):
    if not p:
        raise HcurlError("%s failed: %s" % (what, last_error()))
    return p


def ptr(a):
    """void* of a numpy array, a raw device address (int), or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
