"""ring/: CoeffPoly, NttPoly, RnsPoly -- host mirror of the reference's types whose
arithmetic runs on the GPU through the C ABI (no CPU fallback).

Reference: ring/poly.rs:6-147 (CoeffPoly), ring/ntt.rs:11-139 (NttPoly, make_plan),
ring/rns.rs:14-217 (RnsPoly).  Data lives in numpy uint64 arrays like the reference's
``Vec<u64>``; each operation copies to the device, runs the kernel and copies back.
The batched, device-resident entry points used for throughput are in ``exacto_b200.batch``.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native
from .error import DimensionMismatch, ExactoError, ModulusMismatch
from .params import BfvParams, BfvParamsBuilder, RnsBasis, make_plan as _validate_plan


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


@dataclass(frozen=True)
class Plan:
    """Stand-in for ``Arc<concrete_ntt::prime64::Plan>``: a native context + modulus index."""
    params: BfvParams
    index: int

    def modulus(self) -> int:
        return self.params.modulus(self.index)

    @property
    def ring_degree(self) -> int:
        return self.params.ring_degree


_plan_cache = {}


def make_plan(n: int, modulus: int) -> Plan:
    """ring/ntt.rs:19-29."""
    _validate_plan(n, modulus)
    key = (n, modulus)
    if key not in _plan_cache:
        params = BfvParamsBuilder().ring_degree(n).plain_modulus(2).ct_moduli([modulus]).build()
        _plan_cache[key] = Plan(params, 0)
    return _plan_cache[key]


def plans_of(params: BfvParams) -> List[Plan]:
    count = 1 + (len(params.aux_basis.moduli) if params.aux_basis is not None else 0)
    return [Plan(params, i) for i in range(count)]


class _Scratch:
    """Device scratch for one host-array operation."""

    def __init__(self, plan: Plan):
        self.ctx = plan.params.context()
        self.L = _native.lib()
        self.ptrs = []

    def up(self, a: np.ndarray):
        p = ctypes.c_void_p()
        _native.check(self.L.exb_device_alloc(self.ctx.handle, a.nbytes, ctypes.byref(p)))
        self.ptrs.append(p)
        _native.check(self.L.exb_copy_to_device(self.ctx.handle, p, _ptr(a), a.nbytes, None))
        return p

    def down(self, p, shape) -> np.ndarray:
        out = np.empty(shape, np.uint64)
        _native.check(self.L.exb_copy_to_host(self.ctx.handle, _ptr(out), p, out.nbytes, None))
        _native.check(self.L.exb_synchronize(self.ctx.handle, None))
        return out

    def close(self):
        for p in self.ptrs:
            self.L.exb_device_free(self.ctx.handle, p)
        self.ptrs = []


def _binary(plan: Plan, name: str, a: np.ndarray, b: Optional[np.ndarray], scalar: Optional[int] = None):
    s = _Scratch(plan)
    try:
        da = s.up(a)
        fn = getattr(s.L, name)
        if name == "exb_poly_neg":
            _native.check(fn(s.ctx.handle, plan.index, da, da, a.size, None))
        elif name == "exb_poly_scalar_mul":
            _native.check(fn(s.ctx.handle, plan.index, da, scalar, da, a.size, None))
        else:
            db = s.up(b)
            _native.check(fn(s.ctx.handle, plan.index, da, db, da, a.size, None))
        return s.down(da, a.shape)
    finally:
        s.close()


class CoeffPoly:
    """ring/poly.rs:6-9: coefficients in [0, modulus)."""

    def __init__(self, coeffs, modulus: int):
        self.coeffs = _u64(coeffs)
        self.modulus = int(modulus)

    @staticmethod
    def zero(n: int, modulus: int) -> "CoeffPoly":
        return CoeffPoly(np.zeros(n, np.uint64), modulus)

    @staticmethod
    def from_coeffs(coeffs, modulus: int) -> "CoeffPoly":         # :20-24 (reduces mod q)
        return CoeffPoly(_u64(coeffs) % np.uint64(modulus), modulus)

    def __len__(self):
        return self.coeffs.shape[0]

    def _check(self, other):
        if len(self) != len(other):
            raise DimensionMismatch(len(self), len(other))
        if self.modulus != other.modulus:
            raise ModulusMismatch()

    def add(self, other): self._check(other); return CoeffPoly(_binary(make_plan(len(self), self.modulus), "exb_poly_add", self.coeffs, other.coeffs), self.modulus)
    def sub(self, other): self._check(other); return CoeffPoly(_binary(make_plan(len(self), self.modulus), "exb_poly_sub", self.coeffs, other.coeffs), self.modulus)
    def neg(self): return CoeffPoly(_binary(make_plan(len(self), self.modulus), "exb_poly_neg", self.coeffs, None), self.modulus)
    def scalar_mul(self, s: int): return CoeffPoly(_binary(make_plan(len(self), self.modulus), "exb_poly_scalar_mul", self.coeffs, None, int(s) % (1 << 64)), self.modulus)

    def is_zero(self) -> bool:
        return not self.coeffs.any()

    def centered_coeffs(self) -> np.ndarray:                      # :138-147
        half = self.modulus // 2
        c = self.coeffs.astype(object)
        return np.array([int(v) - self.modulus if int(v) > half else int(v) for v in c], dtype=np.int64)

    def __eq__(self, other):
        return isinstance(other, CoeffPoly) and self.modulus == other.modulus and np.array_equal(self.coeffs, other.coeffs)


class NttPoly:
    """ring/ntt.rs:11-15: evaluations + modulus + plan.  Eval order: natural -> bit-reversed."""

    def __init__(self, evals, modulus: int, plan: Plan):
        self.evals = _u64(evals)
        self.modulus = int(modulus)
        self.plan = plan

    @staticmethod
    def zero(n: int, modulus: int, plan: Plan) -> "NttPoly":
        return NttPoly(np.zeros(n, np.uint64), modulus, plan)

    @staticmethod
    def from_coeff_poly(poly: CoeffPoly, plan: Plan) -> "NttPoly":     # :42-55
        if poly.modulus != plan.modulus():
            raise ModulusMismatch()
        ctx = plan.params.context()
        out = np.empty_like(poly.coeffs)
        _native.check(_native.lib().exb_ntt_forward_host(ctx.handle, plan.index, _ptr(poly.coeffs), _ptr(out), 1))
        return NttPoly(out, poly.modulus, plan)

    def to_coeff_poly(self) -> CoeffPoly:                              # :58-67
        ctx = self.plan.params.context()
        out = np.empty_like(self.evals)
        _native.check(_native.lib().exb_ntt_inverse_host(ctx.handle, self.plan.index, _ptr(self.evals), _ptr(out), 1))
        return CoeffPoly(out, self.modulus)

    def __len__(self):
        return self.evals.shape[0]

    def _check(self, other):
        if len(self) != len(other) or self.modulus != other.modulus:
            raise ModulusMismatch()                                    # :76-78

    def add(self, o): self._check(o); return NttPoly(_binary(self.plan, "exb_poly_add", self.evals, o.evals), self.modulus, self.plan)
    def sub(self, o): self._check(o); return NttPoly(_binary(self.plan, "exb_poly_sub", self.evals, o.evals), self.modulus, self.plan)
    def neg(self): return NttPoly(_binary(self.plan, "exb_poly_neg", self.evals, None), self.modulus, self.plan)
    def mul(self, o): self._check(o); return NttPoly(_binary(self.plan, "exb_poly_mul", self.evals, o.evals), self.modulus, self.plan)
    def scalar_mul(self, s: int): return NttPoly(_binary(self.plan, "exb_poly_scalar_mul", self.evals, None, int(s) % (1 << 64)), self.modulus, self.plan)

    def is_zero(self) -> bool:
        return not self.evals.any()

    def __eq__(self, other):
        return isinstance(other, NttPoly) and self.modulus == other.modulus and np.array_equal(self.evals, other.evals)


class RnsPoly:
    """ring/rns.rs:14-17: one NttPoly per RNS prime."""

    def __init__(self, components: List[NttPoly], ring_degree: int):
        self.components = components
        self.ring_degree = ring_degree

    @staticmethod
    def zero(params: BfvParams) -> "RnsPoly":
        comps = []
        for l in range(params.ct_basis.num_moduli()):
            plan = Plan(params, params.ct_index(l))
            comps.append(NttPoly.zero(params.ring_degree, plan.modulus(), plan))
        return RnsPoly(comps, params.ring_degree)

    @staticmethod
    def from_coeff_poly(poly: CoeffPoly, params: BfvParams) -> "RnsPoly":     # :84-105
        if len(poly) != params.ring_degree:
            raise DimensionMismatch(params.ring_degree, len(poly))
        comps = []
        for l in range(params.ct_basis.num_moduli()):
            plan = Plan(params, params.ct_index(l))
            comps.append(NttPoly.from_coeff_poly(CoeffPoly.from_coeffs(poly.coeffs, plan.modulus()), plan))
        return RnsPoly(comps, params.ring_degree)

    def to_coeff_poly(self) -> CoeffPoly:                                      # :114-151
        coeff = [c.to_coeff_poly() for c in self.components]
        if len(coeff) == 1:                                                    # :130-132
            return coeff[0]
        # u128 CRT of the reference, truncated like `val as u64` / `big_q as u64` (:135-150)
        moduli = [c.modulus for c in coeff]
        m128, m64 = (1 << 128) - 1, (1 << 64) - 1
        big_q = 1
        for q in moduli:
            big_q = (big_q * q) & m128
        inv = []
        for i, qi in enumerate(moduli):                                        # RnsBasis::new :46-55
            prod = 1
            for j, qj in enumerate(moduli):
                if i != j:
                    prod = prod * (qj % qi) % qi
            inv.append(pow(prod, -1, qi))
        out = np.zeros(self.ring_degree, np.uint64)
        for j in range(self.ring_degree):
            val = 0
            for i, qi in enumerate(moduli):
                t = int(coeff[i].coeffs[j]) * inv[i] % qi
                val = ((val + t * (big_q // qi)) & m128) % big_q
            out[j] = val & m64
        return CoeffPoly(out, big_q & m64)

    def num_components(self) -> int:
        return len(self.components)

    def _zip(self, other, op):
        if len(self.components) != len(other.components):
            raise DimensionMismatch(len(self.components), len(other.components))
        return RnsPoly([getattr(a, op)(b) for a, b in zip(self.components, other.components)], self.ring_degree)

    def add(self, o): return self._zip(o, "add")
    def sub(self, o): return self._zip(o, "sub")
    def mul(self, o): return self._zip(o, "mul")
    def neg(self): return RnsPoly([c.neg() for c in self.components], self.ring_degree)
    def scalar_mul(self, s: int): return RnsPoly([c.scalar_mul(s) for c in self.components], self.ring_degree)
