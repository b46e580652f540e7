"""ctypes binding of libexacto_b200.so (include/exacto_b200.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded
this module raises, and every operation of the package fails with it.
"""
from __future__ import annotations

import ctypes
import os

from .error import ExactoError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libexacto_b200.so")

c_u64, c_u32, c_vp, c_sz = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t


class BfvParamsC(ctypes.Structure):
    _fields_ = [
        ("ring_degree", c_u32),
        ("num_ct_moduli", c_u32),
        ("ct_moduli", ctypes.POINTER(c_u64)),
        ("num_aux_moduli", c_u32),
        ("aux_moduli", ctypes.POINTER(c_u64)),
        ("plain_modulus", c_u64),
        ("gadget_base", c_u64),
        ("gadget_digits", c_u32),
    ]


# name -> (restype, argtypes); every symbol include/exacto_b200.h declares.
SYMBOLS = {
    "exb_last_error": (ctypes.c_char_p, []),
    "exb_version": (ctypes.c_char_p, []),
    "exb_context_create": (ctypes.c_int, [ctypes.POINTER(BfvParamsC), ctypes.c_int, ctypes.POINTER(c_vp)]),
    "exb_context_create_ex": (ctypes.c_int, [ctypes.POINTER(BfvParamsC), ctypes.c_int, c_u32, ctypes.POINTER(c_vp)]),
    "exb_context_destroy": (None, [c_vp]),
    "exb_context_set_option": (ctypes.c_int, [c_vp, ctypes.c_char_p, ctypes.c_int64]),
    "exb_ntt_format_id": (ctypes.c_int, [c_vp, c_u32, ctypes.POINTER(c_u64)]),
    "exb_host_alloc": (ctypes.c_int, [c_vp, c_sz, ctypes.POINTER(c_vp)]),
    "exb_host_alloc_ex": (ctypes.c_int, [c_vp, c_sz, c_u32, ctypes.POINTER(c_vp)]),
    "exb_host_free": (ctypes.c_int, [c_vp, c_vp]),
    "exb_host_register": (ctypes.c_int, [c_vp, c_vp, c_sz]),
    "exb_host_unregister": (ctypes.c_int, [c_vp, c_vp]),
    "exb_context_gadget": (ctypes.c_int, [c_vp, ctypes.POINTER(c_u64), ctypes.POINTER(c_u32)]),
    "exb_context_psi": (ctypes.c_int, [c_vp, c_u32, ctypes.POINTER(c_u64)]),
    "exb_launch_count": (ctypes.c_ulonglong, []),
    "exb_profile_enable": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "exb_profile_read": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong)]),
    "exb_device_alloc": (ctypes.c_int, [c_vp, c_sz, ctypes.POINTER(c_vp)]),
    "exb_device_free": (ctypes.c_int, [c_vp, c_vp]),
    "exb_copy_to_device": (ctypes.c_int, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_copy_to_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_synchronize": (ctypes.c_int, [c_vp, c_vp]),
    "exb_ntt_forward": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_sz, c_vp]),
    "exb_ntt_inverse": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_sz, c_vp]),
    "exb_ntt_forward_host": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_sz]),
    "exb_ntt_inverse_host": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_sz]),
    "exb_poly_add": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_poly_sub": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_poly_neg": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_sz, c_vp]),
    "exb_poly_mul": (ctypes.c_int, [c_vp, c_u32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_poly_scalar_mul": (ctypes.c_int, [c_vp, c_u32, c_vp, c_u64, c_vp, c_sz, c_vp]),
    "exb_relin_key_load": (ctypes.c_int, [c_vp, c_vp, c_u32, ctypes.POINTER(c_vp)]),
    "exb_relin_key_load_device": (ctypes.c_int, [c_vp, c_vp, c_u32, c_vp, ctypes.POINTER(c_vp)]),
    "exb_relin_key_destroy": (None, [c_vp]),
    "exb_bfv_mul_and_relin": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_bfv_mul_and_relin_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz]),
    "exb_bfv_add": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_bfv_mul_no_relin": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_bfv_relinearize": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_sz, c_vp]),
    "exb_gadget_decompose": (ctypes.c_int, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "exb_bfv_decrypt": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_sz, c_vp]),
    "exb_bfv_decrypt_host": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_sz]),
    "exb_bfv_apply_automorphism": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp, c_vp, c_sz, c_vp]),
    "exb_bfv_apply_automorphism_host": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp, c_vp, c_sz]),
    "exb_dbfv_mul": (ctypes.c_int, [c_vp, c_u64, c_u32, c_u64, c_vp, c_vp, c_vp, c_vp, c_sz, c_u32, c_u32, c_vp]),
    "exb_dbfv_mul_host": (ctypes.c_int, [c_vp, c_u64, c_u32, c_u64, c_vp, c_vp, c_vp, c_vp, c_sz, c_u32]),
    "exb_dbfv_mul_host_async": (ctypes.c_int, [c_vp, c_u64, c_u32, c_u64, c_vp, c_vp, c_vp, c_vp, c_sz, c_u32, ctypes.POINTER(c_u64)]),
    "exb_bfv_mul_and_relin_host_async": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, ctypes.POINTER(c_u64)]),
    "exb_wait": (ctypes.c_int, [c_vp, c_u64]),
    "exb_dbfv_mul_scatter": (ctypes.c_int, [c_vp, c_u64, c_u32, c_u64, c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), c_u32,
                                            c_sz, c_u32, c_u32, c_vp]),
    "exb_ipc_export": (ctypes.c_int, [c_vp, c_vp, ctypes.c_char_p]),
    "exb_ipc_open": (ctypes.c_int, [c_vp, ctypes.c_char_p, ctypes.POINTER(c_vp)]),
    "exb_ipc_close": (ctypes.c_int, [c_vp, c_vp]),
    "exb_dbfv_small_reps": (ctypes.c_int, [c_u64, c_u32, c_u64, c_vp]),
}

EXB_DBFV_ALL_PRODUCTS = 1
EXB_CTX_REFERENCE_AUX_BASIS = 1
EXB_HOST_WRITE_COMBINED = 1

_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(exacto_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise ExactoError(rc, lib().exb_last_error().decode("utf-8", "replace"))
