"""bootstrap/: the two multiplication entry points north_star names, with the reference's
signatures and rlk-selection logic (bootstrap/bfv_host.rs:242-288).

The GPU content of these functions is ``dbfv_mul`` (this package).  The bootstrap *body*
(``dbfv_bootstrap`` -> ``bfv_bootstrap``: modulus switch, coefficient extraction, digit
extraction polynomial, bootstrap/bfv_host.rs:134-236) is outside the hot-path scope of this
round: callers supply it as ``bsk.bootstrap(ct)``; without one the
functions fail loudly after the multiplication instead of silently skipping the refresh.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import math

import numpy as np

from .bfv import (BfvCiphertext, RelinKey, bfv_add, bfv_mul_and_relin, bfv_scalar_mul, trivial_encrypt)
from .dbfv import DbfvCiphertext, dbfv_mul
from .error import InvalidParam, NotImplementedErr
from .params import BfvParams


class BootstrapKey:
    """bootstrap/bfv_host.rs:19-41 reduced to what the chain logic reads: the boot parameter set,
    its relinearisation key, and a refresh callable standing in for dbfv_bootstrap."""

    def __init__(self, boot_params: BfvParams, boot_rlk: RelinKey,
                 bootstrap: Optional[Callable[[DbfvCiphertext], DbfvCiphertext]] = None):
        self.boot_params = boot_params
        self.boot_rlk = boot_rlk
        self.bootstrap = bootstrap


def dbfv_bootstrap(ct: DbfvCiphertext, bsk: BootstrapKey) -> DbfvCiphertext:
    """bootstrap/bfv_host.rs:212-236 (body supplied by the caller, see module docstring)."""
    if bsk.bootstrap is None:
        raise NotImplementedErr("dbfv_bootstrap body is outside the device library's hot-path scope; "
                                "pass BootstrapKey(bootstrap=...)")
    out = bsk.bootstrap(ct)
    out.mul_depth = 0                                                    # :233
    return out


def dbfv_mul_then_bootstrap(ct1: DbfvCiphertext, ct2: DbfvCiphertext, rlk: RelinKey, bsk: BootstrapKey) -> DbfvCiphertext:
    """bootstrap/bfv_host.rs:242-250."""
    return dbfv_bootstrap(dbfv_mul(ct1, ct2, rlk), bsk)


def _same_bfv(a: BfvParams, b: BfvParams) -> bool:
    return (a.plain_modulus == b.plain_modulus and a.ring_degree == b.ring_degree
            and a.ct_basis.moduli == b.ct_basis.moduli)


def dbfv_mul_chain_then_bootstrap(cts: Sequence[DbfvCiphertext], rlk: RelinKey, bsk: BootstrapKey) -> DbfvCiphertext:
    """bootstrap/bfv_host.rs:258-288: fold with per-step rlk selection by parameter equality."""
    if len(cts) == 0:
        raise InvalidParam("dbfv_mul_chain_then_bootstrap requires at least one ciphertext")
    acc = cts[0]
    for ct in cts[1:]:
        use_boot_rlk = _same_bfv(acc.params.bfv_params, bsk.boot_params)          # :271-274
        rhs = ct if _same_bfv(acc.params.bfv_params, ct.params.bfv_params) else dbfv_bootstrap(ct, bsk)   # :276-283
        acc = dbfv_mul_then_bootstrap(acc, rhs, bsk.boot_rlk if use_boot_rlk else rlk, bsk)
    return acc


# ---- Paterson-Stockmeyer polynomial evaluation ---------------------------------
def _ps_plan(num_coeffs: int):
    d = max(num_coeffs - 1, 0)
    k = max(int(math.ceil(math.sqrt(d + 1.0))), 2)                      # digit_extract.rs:112
    return d, k, (d + k) // k


def eval_poly_homomorphic(ct_x: BfvCiphertext, poly_coeffs, rlk: RelinKey) -> BfvCiphertext:
    """bootstrap/digit_extract.rs:100-157: f(x) on an encrypted x with baby steps x^1..x^k (multiplication
    tree :119-125) and a giant-step Horner recursion (:151-154); every product is bfv_mul_and_relin."""
    params = ct_x.params
    coeffs = [int(c) for c in poly_coeffs]
    d, k, num_groups = _ps_plan(len(coeffs))
    if d == 0:
        return trivial_encrypt(coeffs[0], params)
    baby = [trivial_encrypt(1, params), ct_x]
    for i in range(2, k + 1):
        half = i // 2
        baby.append(bfv_mul_and_relin(baby[half], baby[i - half], rlk))
    groups = []
    for i in range(num_groups):
        g = trivial_encrypt(0, params)
        for j in range(k):
            idx = i * k + j
            if idx >= len(coeffs):
                break
            if coeffs[idx] == 0:
                continue
            g = bfv_add(g, bfv_scalar_mul(baby[j], coeffs[idx]))
        groups.append(g)
    result = groups.pop()
    while groups:
        g = groups.pop()
        result = bfv_add(bfv_mul_and_relin(result, baby[k], rlk), g)
    return result


def eval_poly_homomorphic_batch(params: BfvParams, ct_x, poly_coeffs, rlk: RelinKey):
    """The same evaluation on a device-resident batch ct_x [B, 2, n] (torch.int64 CUDA tensor): every baby /
    giant step is ONE batched bfv_mul_and_relin launch sequence over the B independent ciphertexts, scalar
    multiplications and additions are point-wise kernels (NTT of a constant polynomial is the constant
    vector, so `poly_scalar_mul` equals the reference's bfv_plain_mul by a constant bit for bit)."""
    import torch
    from . import batch
    coeffs = [int(c) for c in poly_coeffs]
    q, p = params.ct_basis.moduli[0], params.plain_modulus
    delta = q // p

    def trivial(m):
        t = torch.zeros_like(ct_x)
        v = (m % p) * delta % q
        t[:, 0, :] = v - (1 << 64) if v >= (1 << 63) else v
        return t

    d, k, num_groups = _ps_plan(len(coeffs))
    if d == 0:
        return trivial(coeffs[0])
    baby = [trivial(1), ct_x]
    for i in range(2, k + 1):
        half = i // 2
        baby.append(batch.bfv_mul_and_relin(params, baby[half], baby[i - half], rlk))
    groups = []
    for i in range(num_groups):
        g = trivial(0)
        for j in range(k):
            idx = i * k + j
            if idx >= len(coeffs):
                break
            if coeffs[idx] == 0:
                continue
            term = batch.poly_scalar_mul(params, 0, baby[j], coeffs[idx] % p)
            g = batch.poly_add(params, 0, g, term)
        groups.append(g)
    result = groups.pop()
    while groups:
        g = groups.pop()
        result = batch.poly_add(params, 0, batch.bfv_mul_and_relin(params, result, baby[k], rlk), g)
    return result
