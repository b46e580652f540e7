"""CPU ORACLE -- test infrastructure, NOT product code.

ctypes front-end of ``oracle/libexacto_oracle.so`` (built from
``oracle/exacto_oracle.c`` by ``oracle/Makefile``), the plain-C restatement of
the reference's CPU algorithm for the ciphertext-multiplication hot path, plus
``oracle.definition`` (big-int "by definition" oracle) and ``oracle.harness``
(keygen / encrypt / decrypt needed to re-run the reference's decrypt KATs).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  ``exacto_b200`` never does.

Pinning status: see the header of ``oracle/exacto_oracle.h``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libexacto_oracle.so")

EXO_MAX_AUX = 4

ERROR_NAMES = {
    0: "Ok", 1: "InvalidParam", 2: "DimensionMismatch", 3: "ModulusMismatch",
    4: "InvalidRingDegree", 5: "DecryptionError", 6: "DecompositionError",
    7: "LatticeError", 8: "MissingKey", 9: "NotImplemented",
}


class OracleError(Exception):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.kind = ERROR_NAMES.get(code, str(code))
        self.message = message


class _CParams(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_uint32),
        ("num_ct", ctypes.c_uint32),
        ("q", ctypes.c_uint64),
        ("num_aux", ctypes.c_uint32),
        ("aux", ctypes.c_uint64 * EXO_MAX_AUX),
        ("plain_modulus", ctypes.c_uint64),
        ("gadget_base", ctypes.c_uint64),
        ("gadget_digits", ctypes.c_uint32),
    ]


@dataclass(frozen=True)
class OracleParams:
    """The fields of BfvParams the hot path reads (params/mod.rs:11-27)."""
    n: int
    q: int
    aux: tuple = ()
    plain_modulus: int = 257
    gadget_base: int = 1 << 16
    gadget_digits: int = 0          # 0 => compute_gadget_digits (params/mod.rs:126-140)
    num_ct: int = 1
    _c: object = field(default=None, compare=False, repr=False)

    def __post_init__(self):
        g = self.gadget_digits
        if g == 0:
            g, pw = 0, 1
            while pw < self.q:
                pw *= self.gadget_base
                g += 1
            g = max(g, 1)
            object.__setattr__(self, "gadget_digits", g)
        c = _CParams()
        c.n, c.num_ct, c.q, c.num_aux = self.n, self.num_ct, self.q, len(self.aux)
        for i, a in enumerate(self.aux[:EXO_MAX_AUX]):
            c.aux[i] = a
        c.plain_modulus, c.gadget_base, c.gadget_digits = self.plain_modulus, self.gadget_base, g
        object.__setattr__(self, "_c", c)


_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "exacto_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "exacto_oracle.h"))):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u64, u32, p64 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
        pp = ctypes.POINTER(_CParams)
        L.exo_last_error.restype = ctypes.c_char_p
        for name in ("exo_mod_mul", "exo_mod_add", "exo_mod_sub", "exo_mod_pow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64, u64]
        L.exo_mod_neg.restype = u64
        L.exo_mod_neg.argtypes = [u64, u64]
        L.exo_mod_inv.argtypes = [u64, u64, ctypes.POINTER(u64)]
        L.exo_is_prime.argtypes = [u64]
        L.exo_find_psi.argtypes = [u32, u64, ctypes.POINTER(u64)]
        L.exo_ntt_tables.argtypes = [u32, u64, p64, p64, ctypes.POINTER(u64)]
        L.exo_ntt_fwd.argtypes = [u32, u64, p64]
        L.exo_ntt_inv.argtypes = [u32, u64, p64]
        L.exo_ntt_fwd_batch.argtypes = [u32, u64, p64, ctypes.c_size_t, ctypes.c_int]
        L.exo_ntt_inv_batch.argtypes = [u32, u64, p64, ctypes.c_size_t, ctypes.c_int]
        L.exo_poly_mul_naive.argtypes = [u32, u64, p64, p64, p64]
        L.exo_poly_mul_naive.restype = None
        L.exo_gadget_decompose.argtypes = [u32, u64, p64, u64, u32, p64]
        L.exo_gadget_decompose.restype = None
        L.exo_bfv_add.argtypes = [pp, p64, p64, p64]
        L.exo_bfv_add.restype = None
        L.exo_bfv_mul_no_relin.argtypes = [pp, p64, p64, p64]
        L.exo_bfv_mul_and_relin.argtypes = [pp, p64, p64, p64, p64]
        L.exo_bfv_mul_and_relin_batch.argtypes = [pp, p64, p64, p64, p64, ctypes.c_size_t, ctypes.c_int]
        L.exo_apply_automorphism.argtypes = [u32, u64, p64, u64, p64]
        L.exo_apply_automorphism.restype = None
        L.exo_bfv_apply_automorphism_batch.argtypes = [pp, p64, p64, u64, p64, ctypes.c_size_t, ctypes.c_int]
        L.exo_relinearize.argtypes = [pp, p64, p64, p64]
        L.exo_small_reps.argtypes = [u64, u32, u64, p64]
        L.exo_small_reps.restype = None
        L.exo_dbfv_mul.argtypes = [pp, u64, u32, u64, p64, p64, p64, p64, ctypes.c_int]
        L.exo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise OracleError(rc, lib().exo_last_error().decode())


def _u64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def max_threads() -> int:
    return lib().exo_max_threads()


# ---- ring/modular.rs -------------------------------------------------------
def mod_mul(a, b, m): return int(lib().exo_mod_mul(a, b, m))
def mod_add(a, b, m): return int(lib().exo_mod_add(a, b, m))
def mod_sub(a, b, m): return int(lib().exo_mod_sub(a, b, m))
def mod_neg(a, m): return int(lib().exo_mod_neg(a, m))
def mod_pow(a, e, m): return int(lib().exo_mod_pow(a, e, m))


def mod_inv(a, m):
    out = ctypes.c_uint64()
    return int(out.value) if lib().exo_mod_inv(a, m, ctypes.byref(out)) else None


def is_prime(m): return bool(lib().exo_is_prime(m))


# ---- ring/ntt.rs -----------------------------------------------------------
def find_psi(n: int, q: int) -> int:
    out = ctypes.c_uint64()
    _check(lib().exo_find_psi(n, q, ctypes.byref(out)))
    return int(out.value)


def ntt_tables(n: int, q: int):
    a, b, ninv = np.zeros(n, np.uint64), np.zeros(n, np.uint64), ctypes.c_uint64()
    _check(lib().exo_ntt_tables(n, q, _ptr(a), _ptr(b), ctypes.byref(ninv)))
    return a, b, int(ninv.value)


def ntt_fwd(a, q: int, threads: int = 1) -> np.ndarray:
    """Forward negacyclic NTT over the last axis (natural -> bit-reversed)."""
    out = _u64(a).copy()
    n = out.shape[-1]
    _check(lib().exo_ntt_fwd_batch(n, q, _ptr(out), out.size // n, threads))
    return out


def ntt_inv(a, q: int, threads: int = 1) -> np.ndarray:
    out = _u64(a).copy()
    n = out.shape[-1]
    _check(lib().exo_ntt_inv_batch(n, q, _ptr(out), out.size // n, threads))
    return out


def poly_mul_naive(a, b, q: int) -> np.ndarray:
    a, b = _u64(a), _u64(b)
    out = np.zeros_like(a)
    lib().exo_poly_mul_naive(a.shape[-1], q, _ptr(a), _ptr(b), _ptr(out))
    return out


# ---- bfv/ ------------------------------------------------------------------
def gadget_decompose(coeffs, q: int, base: int, num_digits: int) -> np.ndarray:
    c = _u64(coeffs)
    out = np.zeros((num_digits, c.shape[-1]), np.uint64)
    lib().exo_gadget_decompose(c.shape[-1], q, _ptr(c), base, num_digits, _ptr(out))
    return out


def bfv_add(p: OracleParams, a, b) -> np.ndarray:
    a, b = _u64(a, (2, p.n)), _u64(b, (2, p.n))
    out = np.zeros_like(a)
    lib().exo_bfv_add(ctypes.byref(p._c), _ptr(a), _ptr(b), _ptr(out))
    return out


def bfv_mul_no_relin(p: OracleParams, ct1, ct2) -> np.ndarray:
    ct1, ct2 = _u64(ct1, (2, p.n)), _u64(ct2, (2, p.n))
    out = np.zeros((3, p.n), np.uint64)
    _check(lib().exo_bfv_mul_no_relin(ctypes.byref(p._c), _ptr(ct1), _ptr(ct2), _ptr(out)))
    return out


def bfv_mul_and_relin(p: OracleParams, ct1, ct2, rlk, threads: int = 1) -> np.ndarray:
    """ct [..,2,n] NTT domain, rlk [G,2,n] -> [..,2,n]."""
    ct1, ct2 = _u64(ct1), _u64(ct2)
    rlk = _u64(rlk, (p.gadget_digits, 2, p.n))
    assert ct1.shape == ct2.shape and ct1.shape[-2:] == (2, p.n)
    out = np.zeros_like(ct1)
    batch = ct1.size // (2 * p.n)
    _check(lib().exo_bfv_mul_and_relin_batch(ctypes.byref(p._c), _ptr(ct1), _ptr(ct2), _ptr(rlk),
                                             _ptr(out), batch, threads))
    return out


def relinearize(p: OracleParams, ct3, rlk) -> np.ndarray:
    """bfv/keyswitch.rs:59-101: ct3 [..,3,n] NTT domain -> [..,2,n]."""
    ct3 = _u64(ct3)
    rlk = _u64(rlk, (p.gadget_digits, 2, p.n))
    flat = ct3.reshape(-1, 3, p.n)
    out = np.zeros((flat.shape[0], 2, p.n), np.uint64)
    for i in range(flat.shape[0]):
        _check(lib().exo_relinearize(ctypes.byref(p._c), _ptr(np.ascontiguousarray(flat[i])), _ptr(rlk), _ptr(out[i])))
    return out.reshape(ct3.shape[:-2] + (2, p.n))


# ---- Galois automorphism + key switch ----------------------------------------
def apply_automorphism(coeffs, q: int, k: int) -> np.ndarray:
    """bfv/keygen.rs:218-239 on one coefficient-domain poly."""
    a = _u64(coeffs)
    out = np.zeros_like(a)
    lib().exo_apply_automorphism(a.shape[-1], q, _ptr(a), k, _ptr(out))
    return out


def bfv_apply_automorphism(p: OracleParams, ct, gk, k: int, threads: int = 1) -> np.ndarray:
    """bfv/eval.rs:512-561.  ct [..,2,n] NTT domain, gk [G,2,n] -> [..,2,n]."""
    ct = _u64(ct)
    gk = _u64(gk, (p.gadget_digits, 2, p.n))
    assert ct.shape[-2:] == (2, p.n)
    out = np.zeros_like(ct)
    _check(lib().exo_bfv_apply_automorphism_batch(ctypes.byref(p._c), _ptr(ct), _ptr(gk), k, _ptr(out),
                                                  ct.size // (2 * p.n), threads))
    return out


# ---- dbfv/ -----------------------------------------------------------------
def small_reps(base: int, d: int, plain_modulus: int) -> np.ndarray:
    out = np.zeros((max(d - 1, 0), d), np.int64)
    if d > 1:
        lib().exo_small_reps(base, d, plain_modulus, _ptr(out))
    return out


def dbfv_mul(p: OracleParams, base: int, d: int, dbfv_plain_modulus: int, ct1, ct2, rlk,
             threads: int = 1) -> np.ndarray:
    """ct [d,2,n] NTT domain -> [d,2,n] (dbfv/eval.rs:82-149)."""
    ct1, ct2 = _u64(ct1, (d, 2, p.n)), _u64(ct2, (d, 2, p.n))
    rlk = _u64(rlk, (p.gadget_digits, 2, p.n))
    out = np.zeros_like(ct1)
    _check(lib().exo_dbfv_mul(ctypes.byref(p._c), base, d, dbfv_plain_modulus, _ptr(ct1), _ptr(ct2),
                              _ptr(rlk), _ptr(out), threads))
    return out
