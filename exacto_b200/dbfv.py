"""dbfv/: DbfvCiphertext and dbfv_mul / dbfv_add / dbfv_sub / dbfv_neg with the reference's
signatures, guards and metadata rules (dbfv/eval.rs:11-149, dbfv/ciphertext.rs:10-34);
the d*d per-digit BFV multiplications, the per-k accumulation and reduction::reduce
(dbfv/reduction.rs:15-60) run as one batched launch sequence on the GPU.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _native
from .bfv import (BfvCiphertext, GaloisKey, RelinKey, bfv_add, bfv_apply_automorphism_batch, bfv_neg,
                  bfv_sub, relinearize)
from .error import DimensionMismatch, InvalidParam, NotImplementedErr
from .params import DbfvParams
from .ring import _ptr, _u64


class DbfvCiphertext:
    """dbfv/ciphertext.rs:10-22."""

    def __init__(self, limbs: List[BfvCiphertext], degree: int, mul_depth: int, params: DbfvParams):
        self.limbs = limbs
        self.degree = degree
        self.mul_depth = mul_depth
        self.params = params

    def num_limbs(self) -> int:
        return len(self.limbs)

    def needs_reduction(self) -> bool:
        return self.degree > self.params.num_digits

    @staticmethod
    def from_array(arr, params: DbfvParams, degree: Optional[int] = None, mul_depth: int = 0) -> "DbfvCiphertext":
        """arr: [limbs][2][n] NTT-domain residues."""
        arr = _u64(arr)
        limbs = [BfvCiphertext.from_array(arr[i], params.bfv_params) for i in range(arr.shape[0])]
        return DbfvCiphertext(limbs, params.num_digits if degree is None else degree, mul_depth, params)

    def to_array(self) -> np.ndarray:
        return np.stack([l.to_array() for l in self.limbs])


def _limbwise(ct1, ct2, fn):
    if ct1.num_limbs() != ct2.num_limbs():
        raise DimensionMismatch(ct1.num_limbs(), ct2.num_limbs())
    limbs = [fn(a, b) for a, b in zip(ct1.limbs, ct2.limbs)]
    return DbfvCiphertext(limbs, max(ct1.degree, ct2.degree), max(ct1.mul_depth, ct2.mul_depth), ct1.params)


def dbfv_add(ct1: DbfvCiphertext, ct2: DbfvCiphertext) -> DbfvCiphertext:
    """dbfv/eval.rs:11-33."""
    return _limbwise(ct1, ct2, bfv_add)


def dbfv_sub(ct1: DbfvCiphertext, ct2: DbfvCiphertext) -> DbfvCiphertext:
    """dbfv/eval.rs:36-58."""
    return _limbwise(ct1, ct2, bfv_sub)


def dbfv_neg(ct: DbfvCiphertext) -> DbfvCiphertext:
    """dbfv/eval.rs:61-70."""
    return DbfvCiphertext([bfv_neg(l) for l in ct.limbs], ct.degree, ct.mul_depth, ct.params)


def dbfv_mul(ct1: DbfvCiphertext, ct2: DbfvCiphertext, rlk: RelinKey, *, all_products: bool = False) -> DbfvCiphertext:
    """dbfv/eval.rs:82-149.  ``all_products=True`` also computes the products whose output
    limb reduce() discards (what the reference does); the result is bit-identical."""
    params = ct1.params
    d = params.num_digits
    if ct1.num_limbs() != d or ct2.num_limbs() != d:                     # :90-94
        raise InvalidParam("multiplication requires d-limb ciphertexts")
    next_depth = max(ct1.mul_depth, ct2.mul_depth) + 1
    if next_depth > 1:                                                   # :96-102
        raise NotImplementedErr(
            "chained dBFV multiplication requires ciphertext-level lattice reduction (paper §4.6.2)")
    for limb in list(ct1.limbs) + list(ct2.limbs):                       # bfv/eval.rs:93-97 per product
        if len(limb.c) != 2:
            raise InvalidParam("multiplication requires degree-1 ciphertexts")
    out = dbfv_mul_batch(params, ct1.to_array()[None], ct2.to_array()[None], rlk, all_products=all_products)
    return DbfvCiphertext.from_array(out[0], params, degree=d, mul_depth=next_depth)   # :138-146, reduction.rs:54-59


def dbfv_mul_batch(params: DbfvParams, ct1: np.ndarray, ct2: np.ndarray, rlk: RelinKey, *,
                   all_products: bool = False, device: Optional[int] = None) -> np.ndarray:
    """Batched host-buffer form: ct [B][d][2][n] -> [B][d][2][n] (exb_dbfv_mul_host)."""
    ct1, ct2 = _u64(ct1), _u64(ct2)
    n, d = params.bfv_params.ring_degree, params.num_digits
    Lq = params.bfv_params.ct_basis.num_moduli()          # multi-prime limbs are [2][L][n]
    if ct1.shape != ct2.shape or ct1.shape[1:] != ((d, 2, n) if Lq == 1 else (d, 2, Lq, n)):
        raise InvalidParam("multiplication requires d-limb ciphertexts")
    ctx = params.bfv_params.context(device)
    out = np.empty_like(ct1)
    flags = _native.EXB_DBFV_ALL_PRODUCTS if all_products else 0
    _native.check(_native.lib().exb_dbfv_mul_host(ctx.handle, params.base, d, params.plain_modulus, _ptr(ct1),
                                                  _ptr(ct2), rlk.native(ctx), _ptr(out), ct1.shape[0], flags))
    return out


def dbfv_relinearize(ct: DbfvCiphertext, rlk: RelinKey) -> DbfvCiphertext:
    """dbfv/keyswitch.rs:9-23: BFV relinearisation of every limb; metadata unchanged."""
    return DbfvCiphertext([relinearize(l, rlk) for l in ct.limbs], ct.degree, ct.mul_depth, ct.params)


def dbfv_apply_automorphism(ct: DbfvCiphertext, gk: GaloisKey) -> DbfvCiphertext:
    """dbfv/advanced.rs:15-30: the BFV automorphism + key switch on every limb (one batched launch);
    degree and mul_depth carry over."""
    for limb in ct.limbs:
        if len(limb.c) != 2:                                             # bfv/eval.rs:516-520
            raise InvalidParam("automorphism requires degree-1 ciphertext")
    out = bfv_apply_automorphism_batch(ct.params.bfv_params, ct.to_array(), gk)
    return DbfvCiphertext.from_array(out, ct.params, degree=ct.degree, mul_depth=ct.mul_depth)


def small_reps(base: int, d: int, plain_modulus: int) -> np.ndarray:
    """SmallReps::compute_simple (dbfv/lattice.rs:104-122)."""
    out = np.zeros((max(d - 1, 0), d), np.int64)
    if d > 1:
        _native.check(_native.lib().exb_dbfv_small_reps(base, d, plain_modulus, out.ctypes.data))
    return out
