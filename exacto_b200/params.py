"""Parameter sets: the host mirror of params/mod.rs and params/presets.rs.

``BfvParamsBuilder().ring_degree(..)...build()`` keeps the reference's builder surface
(params/mod.rs:39-124).  A ``BfvParams`` owns (lazily, per CUDA device) the native
context of libexacto_b200.so that holds the NTT plans and HPS constants on the GPU.
"""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _native
from .error import ExactoError, InvalidParam, InvalidRingDegree


def _is_prime(m: int) -> bool:
    if m < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if m == p:
            return True
        if m % p == 0:
            return False
    d, s = m - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, m)
        if x in (1, m - 1):
            continue
        for _ in range(s - 1):
            x = x * x % m
            if x == m - 1:
                break
        else:
            return False
    return True


def make_plan(n: int, modulus: int) -> None:
    """Validation half of ring/ntt.rs:19-29 (the plan itself lives on the GPU)."""
    if n < 2 or n & (n - 1):
        raise InvalidRingDegree(n)
    if not _is_prime(modulus) or (modulus - 1) % (2 * n) != 0:
        raise InvalidParam(f"cannot create NTT plan for n={n}, q={modulus} (need prime q ≡ 1 mod {2 * n})")


class RnsBasis:
    """ring/rns.rs:21-63: the moduli of an RNS basis (plans live in the native context)."""

    def __init__(self, moduli: List[int], ring_degree: int):
        for q in moduli:
            make_plan(ring_degree, q)
        self.moduli = [int(q) for q in moduli]
        self.ring_degree = ring_degree
        self.barrett_ks = [(1 << 64) // q for q in self.moduli]          # :40-42
        self.q_star_inv = []                                             # :45-54
        for i, qi in enumerate(self.moduli):
            prod = 1
            for j, qj in enumerate(self.moduli):
                if i != j:
                    prod = prod * (qj % qi) % qi
            self.q_star_inv.append(pow(prod, -1, qi))

    def num_moduli(self) -> int:
        return len(self.moduli)

    def __eq__(self, other):
        return isinstance(other, RnsBasis) and self.moduli == other.moduli and self.ring_degree == other.ring_degree


def compute_gadget_digits(ct_moduli: List[int], base: int) -> int:
    """params/mod.rs:126-140."""
    q_big = 1
    for q in ct_moduli:
        q_big *= q
    pw, digits = 1, 0
    while pw < q_big:
        pw *= base
        digits += 1
    return max(digits, 1)


class _NativeContext:
    """Owns one exb_context (and the device relin keys created from it)."""

    def __init__(self, params: "BfvParams", device: int):
        L = _native.lib()
        ct = (ctypes.c_uint64 * len(params.ct_basis.moduli))(*params.ct_basis.moduli)
        aux_list = params.aux_basis.moduli if params.aux_basis is not None else []
        aux = (ctypes.c_uint64 * max(len(aux_list), 1))(*aux_list)
        c = _native.BfvParamsC()
        c.ring_degree = params.ring_degree
        c.num_ct_moduli, c.ct_moduli = len(params.ct_basis.moduli), ct
        c.num_aux_moduli, c.aux_moduli = len(aux_list), aux
        c.plain_modulus, c.gadget_base, c.gadget_digits = params.plain_modulus, params.gadget_base, params.gadget_digits
        h = ctypes.c_void_p()
        _native.check(L.exb_context_create_ex(ctypes.byref(c), device, params.context_flags, ctypes.byref(h)))
        self.handle, self.device, self._L = h, device, L

    def set_option(self, name: str, value: int) -> None:
        """exb_context_set_option: tuning knobs (device_chunk_bytes, host_chunk_products, tensor_per_product,
        relin_narrow)."""
        _native.check(self._L.exb_context_set_option(self.handle, name.encode(), int(value)))

    def ntt_format_id(self, index: int = 0) -> int:
        """Identifier of the NTT-domain word format this context produces (exb_ntt_format_id)."""
        out = ctypes.c_uint64()
        _native.check(self._L.exb_ntt_format_id(self.handle, index, ctypes.byref(out)))
        return out.value

    def __del__(self):
        try:
            if self.handle:
                self._L.exb_context_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass
class BfvParams:
    """params/mod.rs:11-27."""
    ring_degree: int
    plain_modulus: int
    ct_basis: RnsBasis
    aux_basis: Optional[RnsBasis]
    sigma: float
    gadget_base: int
    gadget_digits: int
    context_flags: int = field(default=0, repr=False, compare=False)   # exb_context_create_ex flags (set before first use)
    _ctx: Dict[int, _NativeContext] = field(default_factory=dict, repr=False, compare=False)
    _lock: threading.Lock = field(default_factory=threading.Lock, repr=False, compare=False)

    def context(self, device: Optional[int] = None) -> _NativeContext:
        """The native context on `device` (default: current torch device, else 0)."""
        if device is None:
            device = default_device()
        with self._lock:
            ctx = self._ctx.get(device)
            if ctx is None:
                ctx = self._ctx[device] = _NativeContext(self, device)
            return ctx

    def modulus(self, index: int) -> int:
        """Native modulus index: 0 = q_0, 1..A = aux primes, then the remaining ciphertext primes q_1.."""
        if index == 0:
            return self.ct_basis.moduli[0]
        num_aux = len(self.aux_basis.moduli) if self.aux_basis is not None else 0
        if index <= num_aux:
            return self.aux_basis.moduli[index - 1]
        return self.ct_basis.moduli[index - num_aux]

    def ct_index(self, l: int) -> int:
        """Native modulus index of ciphertext prime q_l."""
        num_aux = len(self.aux_basis.moduli) if self.aux_basis is not None else 0
        return 0 if l == 0 else num_aux + l


_default_device = [None]


def set_default_device(device: int) -> None:
    _default_device[0] = device


def default_device() -> int:
    if _default_device[0] is not None:
        return _default_device[0]
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


class BfvParamsBuilder:
    """params/mod.rs:29-124 (same defaults: n=4096, p=65537, sigma=3.2, gadget auto 2^16)."""

    def __init__(self):
        self._n, self._p = 4096, 65537
        self._ct: List[int] = []
        self._aux: List[int] = []
        self._sigma, self._gb = 3.2, 0

    def ring_degree(self, n): self._n = n; return self
    def plain_modulus(self, p): self._p = p; return self
    def ct_moduli(self, m): self._ct = list(m); return self
    def aux_moduli(self, m): self._aux = list(m); return self
    def sigma(self, s): self._sigma = s; return self
    def gadget_base(self, b): self._gb = b; return self

    def build(self) -> BfvParams:
        n = self._n
        if n < 2 or n & (n - 1):
            raise InvalidRingDegree(n)
        if not self._ct:
            raise InvalidParam("must specify at least one ciphertext modulus")
        if self._p < 2:
            raise InvalidParam("plaintext modulus must be >= 2")
        ct_basis = RnsBasis(self._ct, n)
        aux_basis = RnsBasis(self._aux, n) if self._aux else None
        base = self._gb if self._gb else 1 << 16
        digits = max(compute_gadget_digits(self._ct, base), 1)
        return BfvParams(n, self._p, ct_basis, aux_basis, self._sigma, base, digits)


@dataclass
class DbfvParams:
    """params/mod.rs:143-192."""
    bfv_params: BfvParams
    base: int
    num_digits: int
    plain_modulus: int      # 0 == 2^64

    @staticmethod
    def new(bfv_params: BfvParams, base: int, num_digits: int, plain_modulus: int) -> "DbfvParams":
        if base < 2:
            raise InvalidParam("base must be >= 2")
        if num_digits < 1:
            raise InvalidParam("num_digits must be >= 1")
        bd = min(base ** num_digits, (1 << 128) - 1)
        p128 = (1 << 64) if plain_modulus == 0 else plain_modulus
        if bd < p128:
            raise InvalidParam(f"base^digits = {bd} < plain_modulus = {p128}")
        return DbfvParams(bfv_params, base, num_digits, plain_modulus)


# ---- presets (params/presets.rs) ----------------------------------------------------
def compact_bfv() -> BfvParams:                                   # :24-35
    return (BfvParamsBuilder().ring_degree(1024).plain_modulus(257)
            .ct_moduli([1099509805057]).aux_moduli([562949953443841]).sigma(3.2).build())


def small_bfv() -> BfvParams:                                     # :39-51 (no aux basis: mul is refused)
    return (BfvParamsBuilder().ring_degree(4096).plain_modulus(65537)
            .ct_moduli([576460752308273153]).sigma(3.2).build())


def u64_dbfv() -> DbfvParams:                                     # :61-75
    bfv = (BfvParamsBuilder().ring_degree(4096).plain_modulus(1040407)
           .ct_moduli([1152921504606830593])
           .aux_moduli([18014398509998081, 36028797018972161])
           .gadget_base(256).sigma(3.2).build())
    return DbfvParams.new(bfv, 256, 8, 0)


def compact_dbfv() -> DbfvParams:                                 # :86-98
    bfv = (BfvParamsBuilder().ring_degree(1024).plain_modulus(929)
           .ct_moduli([1099509805057]).aux_moduli([562949953443841]).sigma(3.2).build())
    return DbfvParams.new(bfv, 16, 2, 256)


def cfg3_prime_dbfv() -> DbfvParams:
    """BASELINE config 3 made runnable: README custom n=4096 / q=576460752308273153 / p=65537
    (README.md:109-119) plus the u64 profile's aux primes, wrapped as dBFV p=65536, b=256, d=2.
    Without the aux primes the reference itself returns NotImplemented (SURVEY finding 4)."""
    bfv = (BfvParamsBuilder().ring_degree(4096).plain_modulus(65537)
           .ct_moduli([576460752308273153])
           .aux_moduli([18014398509998081, 36028797018972161]).sigma(3.2).build())
    return DbfvParams.new(bfv, 256, 2, 65536)
