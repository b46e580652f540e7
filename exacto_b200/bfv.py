"""bfv/: BfvCiphertext, RelinKey and the evaluation functions of bfv/eval.rs that sit
on the ciphertext-multiplication hot path, with the reference's signatures and error
behaviour; the arithmetic runs in libexacto_b200.so on the GPU.

    bfv_mul_and_relin(ct1, ct2, rlk)   bfv/eval.rs:73-82
    bfv_mul_no_relin(ct1, ct2)         bfv/eval.rs:89-108
    relinearize(ct, rlk)               bfv/keyswitch.rs:59-101
    gadget_decompose(poly, base, G)    bfv/keyswitch.rs:11-52
    bfv_add / bfv_sub / bfv_neg        bfv/eval.rs:14-62
    bfv_apply_automorphism / bfv_trace bfv/eval.rs:512-588 (Galois automorphism + key switch)
    bfv_inner_product                  bfv/eval.rs:593-606
    relinearize passthrough rules      bfv/keyswitch.rs:59-70
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from . import _native
from .error import ExactoError, InvalidParam
from .params import BfvParams
from .ring import NttPoly, Plan, RnsPoly, _ptr, _u64


class BfvCiphertext:
    """bfv/mod.rs:19-24: ``c`` is a list of RnsPoly (2, or 3 before relinearisation)."""

    def __init__(self, c: List[RnsPoly], params: BfvParams):
        self.c = c
        self.params = params

    def degree(self) -> int:
        return len(self.c) - 1

    # -- array views used by the batched entry points --------------------------------
    @staticmethod
    def from_array(arr, params: BfvParams) -> "BfvCiphertext":
        """arr: [k][n] NTT-domain residues mod q_0, or [k][L][n] for a multi-prime ciphertext modulus."""
        arr = _u64(arr)
        L = params.ct_basis.num_moduli()
        if L == 1 and arr.ndim == 2:
            arr = arr[:, None, :]
        plans = [Plan(params, params.ct_index(l)) for l in range(L)]
        return BfvCiphertext([RnsPoly([NttPoly(arr[i, l].copy(), plans[l].modulus(), plans[l]) for l in range(L)],
                                      params.ring_degree) for i in range(arr.shape[0])], params)

    def to_array(self) -> np.ndarray:
        """[k][n] for a single ciphertext prime, [k][L][n] otherwise."""
        if len(self.c[0].components) == 1:
            return np.stack([ci.components[0].evals for ci in self.c])
        return np.stack([np.stack([comp.evals for comp in ci.components]) for ci in self.c])


class RelinKey:
    """bfv/keygen.rs:39-45: ``keys[g] = (rlk0_g, rlk1_g)`` in the NTT domain.  The device copy
    (Montgomery form) is created on first use and cached per device."""

    def __init__(self, keys, params: BfvParams):
        L = params.ct_basis.num_moduli()
        if isinstance(keys, np.ndarray):
            self.array = _u64(keys)                                    # [G][2][n]  ([G][2][L][n] for L > 1 primes)
        elif L == 1:
            self.array = np.stack([np.stack([k0.components[0].evals, k1.components[0].evals])
                                   for (k0, k1) in keys]) if len(keys) else np.zeros((0, 2, params.ring_degree), np.uint64)
        else:
            self.array = np.stack([np.stack([np.stack([c.evals for c in k.components]) for k in pair]) for pair in keys]) \
                if len(keys) else np.zeros((0, 2, L, params.ring_degree), np.uint64)
        self.params = params
        self._native = {}

    @property
    def keys(self):
        n, L = self.params.ring_degree, self.params.ct_basis.num_moduli()
        plans = [Plan(self.params, self.params.ct_index(l)) for l in range(L)]
        arr = self.array if self.array.ndim == 4 else self.array[:, :, None, :]
        return [tuple(RnsPoly([NttPoly(arr[g, c, l], plans[l].modulus(), plans[l]) for l in range(L)], n) for c in range(2))
                for g in range(arr.shape[0])]

    def native(self, ctx) -> ctypes.c_void_p:
        key = id(ctx)
        if key not in self._native:
            h = ctypes.c_void_p()
            _native.check(_native.lib().exb_relin_key_load(ctx.handle, _ptr(self.array), self.array.shape[0],
                                                           ctypes.byref(h)))
            self._native[key] = (h, ctx)
        return self._native[key][0]

    def __del__(self):
        try:
            for h, _ctx in self._native.values():
                _native.lib().exb_relin_key_destroy(h)
        except Exception:
            pass


class GaloisKey(RelinKey):
    """bfv/keygen.rs:47-54: key-switch key from s(X^element) to s(X); same [G][2][n] layout as a
    RelinKey, plus the Galois element."""

    def __init__(self, keys, element: int, params: BfvParams):
        super().__init__(keys, params)
        self.element = int(element)


def _zip_op(ct1: BfvCiphertext, ct2: BfvCiphertext, op: str, lone_rhs) -> BfvCiphertext:
    c = []
    for i in range(max(len(ct1.c), len(ct2.c))):
        a = ct1.c[i] if i < len(ct1.c) else None
        b = ct2.c[i] if i < len(ct2.c) else None
        if a is not None and b is not None:
            c.append(getattr(a, op)(b))
        elif a is not None:
            c.append(a)
        else:
            c.append(lone_rhs(b))
    return BfvCiphertext(c, ct1.params)


def bfv_add(ct1: BfvCiphertext, ct2: BfvCiphertext) -> BfvCiphertext:
    """bfv/eval.rs:14-31 (ragged lengths allowed)."""
    return _zip_op(ct1, ct2, "add", lambda b: b)


def bfv_sub(ct1: BfvCiphertext, ct2: BfvCiphertext) -> BfvCiphertext:
    """bfv/eval.rs:34-51."""
    return _zip_op(ct1, ct2, "sub", lambda b: b.neg())


def bfv_neg(ct: BfvCiphertext) -> BfvCiphertext:
    """bfv/eval.rs:54-60."""
    return BfvCiphertext([ci.neg() for ci in ct.c], ct.params)


def _check_degree1(ct1: BfvCiphertext, ct2: BfvCiphertext):
    if len(ct1.c) != 2 or len(ct2.c) != 2:                              # bfv/eval.rs:93-97
        raise InvalidParam("multiplication requires degree-1 ciphertexts")


def bfv_mul_and_relin(ct1: BfvCiphertext, ct2: BfvCiphertext, rlk: RelinKey) -> BfvCiphertext:
    """bfv/eval.rs:73-82.  Borrows the inputs, returns a fresh ciphertext sharing ct1.params."""
    _check_degree1(ct1, ct2)
    params = ct1.params
    out = bfv_mul_and_relin_batch(params, ct1.to_array()[None], ct2.to_array()[None], rlk)
    return BfvCiphertext.from_array(out[0], params)


def bfv_mul_and_relin_batch(params: BfvParams, ct1: np.ndarray, ct2: np.ndarray, rlk: RelinKey,
                            device: Optional[int] = None) -> np.ndarray:
    """Batched host-buffer form: ct [B][2][n] -> [B][2][n] (exb_bfv_mul_and_relin_host)."""
    ct1, ct2 = _u64(ct1), _u64(ct2)
    n, L = params.ring_degree, params.ct_basis.num_moduli()
    if ct1.shape != ct2.shape or ct1.shape[1:] != ((2, n) if L == 1 else (2, L, n)):
        raise InvalidParam("multiplication requires degree-1 ciphertexts")
    ctx = params.context(device)
    out = np.empty_like(ct1)
    _native.check(_native.lib().exb_bfv_mul_and_relin_host(ctx.handle, _ptr(ct1), _ptr(ct2), rlk.native(ctx),
                                                           _ptr(out), ct1.shape[0]))
    return out


def _dev_call(params: BfvParams, fn, *arrays, out_shape, extra=()):
    """Device round trip for the stand-alone halves: host arrays -> HBM -> kernel sequence -> host array."""
    import torch  # noqa: F401  (device memory plumbing only)
    from . import batch
    devs = [batch.to_device(a) for a in arrays]
    out = batch.to_device(np.zeros(out_shape, np.uint64))
    ctx = params.context(devs[0].device.index)
    _native.check(fn(ctx, devs, out))
    return batch.to_host(out)


def bfv_mul_no_relin(ct1: BfvCiphertext, ct2: BfvCiphertext) -> BfvCiphertext:
    """bfv/eval.rs:89-108: the degree-2 product (c0 d0, c0 d1 + c1 d0, c1 d1), scaled and rounded."""
    _check_degree1(ct1, ct2)
    params = ct1.params
    out = bfv_mul_no_relin_batch(params, ct1.to_array()[None], ct2.to_array()[None])
    return BfvCiphertext.from_array(out[0], params)


def bfv_mul_no_relin_batch(params: BfvParams, ct1: np.ndarray, ct2: np.ndarray) -> np.ndarray:
    """ct [B][2][n] x2 -> [B][3][n] (exb_bfv_mul_no_relin)."""
    ct1, ct2 = _u64(ct1), _u64(ct2)
    n, Lq = params.ring_degree, params.ct_basis.num_moduli()
    tail = (n,) if Lq == 1 else (Lq, n)                   # multi-prime ciphertexts are [B][k][L][n]
    if ct1.shape != ct2.shape or ct1.shape[1:] != (2,) + tail:
        raise InvalidParam("multiplication requires degree-1 ciphertexts")
    L = _native.lib()
    return _dev_call(params, lambda ctx, d, o: L.exb_bfv_mul_no_relin(ctx.handle, d[0].data_ptr(), d[1].data_ptr(), o.data_ptr(),
                                                                      ct1.shape[0], None),
                     ct1, ct2, out_shape=(ct1.shape[0], 3) + tail)


def relinearize(ct: BfvCiphertext, rlk: RelinKey) -> BfvCiphertext:
    """bfv/keyswitch.rs:59-101: degree-2 -> degree-1 with the relinearisation key; fewer than three components
    are returned unchanged (:63-65), more than three are rejected (:66-70)."""
    if len(ct.c) < 3:
        return ct
    if len(ct.c) > 3:
        raise InvalidParam("relinearization only supports degree-2 ciphertexts")
    out = relinearize_batch(ct.params, ct.to_array()[None], rlk)
    return BfvCiphertext.from_array(out[0], ct.params)


def relinearize_batch(params: BfvParams, ct3: np.ndarray, rlk: RelinKey) -> np.ndarray:
    """ct [B][3][n] -> [B][2][n] (exb_bfv_relinearize)."""
    ct3 = _u64(ct3)
    n, Lq = params.ring_degree, params.ct_basis.num_moduli()
    tail = (n,) if Lq == 1 else (Lq, n)
    if ct3.shape[1:] != (3,) + tail:
        raise InvalidParam("relinearization only supports degree-2 ciphertexts")
    L = _native.lib()
    return _dev_call(params, lambda ctx, d, o: L.exb_bfv_relinearize(ctx.handle, d[0].data_ptr(), 3, rlk.native(ctx), o.data_ptr(),
                                                                     ct3.shape[0], None),
                     ct3, out_shape=(ct3.shape[0], 2) + tail)


def gadget_decompose(poly, params: BfvParams) -> list:
    """bfv/keyswitch.rs:11-52 with the context's gadget base and digit count: balanced base-B digits of the centred
    coefficients, each digit polynomial stored mod q."""
    from .ring import CoeffPoly
    q, n, G = params.ct_basis.moduli[0], params.ring_degree, params.gadget_digits
    coeffs = _u64(poly.coeffs)[None]
    L = _native.lib()
    out = _dev_call(params, lambda ctx, d, o: L.exb_gadget_decompose(ctx.handle, d[0].data_ptr(), o.data_ptr(), 1, None),
                    coeffs, out_shape=(1, G, n))
    return [CoeffPoly(out[0, g], q) for g in range(G)]


# ---- Galois automorphism + key switch --------------------------------------------------------------
def bfv_apply_automorphism(ct: BfvCiphertext, gk: GaloisKey) -> BfvCiphertext:
    """bfv/eval.rs:512-561: sigma_k on (c0, c1), then key-switch c1 back to s with the Galois key."""
    if len(ct.c) != 2:                                                   # :516-520
        raise InvalidParam("automorphism requires degree-1 ciphertext")
    out = bfv_apply_automorphism_batch(ct.params, ct.to_array()[None], gk)
    return BfvCiphertext.from_array(out[0], ct.params)


def bfv_apply_automorphism_batch(params: BfvParams, ct: np.ndarray, gk: GaloisKey,
                                 device: Optional[int] = None) -> np.ndarray:
    """Batched host-buffer form: ct [B][2][n] -> [B][2][n] (exb_bfv_apply_automorphism_host)."""
    ct = _u64(ct)
    if ct.ndim != 3 or ct.shape[1:] != (2, params.ring_degree):
        raise InvalidParam("automorphism requires degree-1 ciphertext")
    ctx = params.context(device)
    out = np.empty_like(ct)
    _native.check(_native.lib().exb_bfv_apply_automorphism_host(ctx.handle, _ptr(ct), gk.element, gk.native(ctx),
                                                                _ptr(out), ct.shape[0]))
    return out


def bfv_trace(ct: BfvCiphertext, galois_elements: Sequence[int], galois_keys) -> BfvCiphertext:
    """bfv/eval.rs:573-588: result <- result + sigma_k(result) for each k in order."""
    result = ct
    for k in galois_elements:
        gk = galois_keys.get(k)
        if gk is None:
            raise InvalidParam(f"missing Galois key for element {k}")
        result = bfv_add(result, bfv_apply_automorphism(result, gk))
    return result


def bfv_inner_product(cts: Sequence[BfvCiphertext], pts) -> BfvCiphertext:
    """bfv/eval.rs:593-606: sum_i pt_i * ct_i."""
    if len(cts) == 0 or len(cts) != len(pts):
        raise InvalidParam("mismatched ct/pt lengths")
    acc = bfv_plain_mul(cts[0], pts[0])
    for ct, pt in zip(cts[1:], pts[1:]):
        acc = bfv_add(acc, bfv_plain_mul(ct, pt))
    return acc


# ---- plaintext operations used by the polynomial evaluator (bootstrap/digit_extract.rs) -------------
def scale_plaintext(plaintext, params: BfvParams) -> RnsPoly:
    """bfv/encrypt.rs:181-229 for a single ciphertext prime: NTT(Delta * m mod q), Delta = floor(q / p)."""
    from .ring import CoeffPoly
    q = params.ct_basis.moduli[0]
    delta = q // params.plain_modulus
    coeffs = np.array([(int(m) % q) * delta % q for m in plaintext.coeffs], dtype=np.uint64)
    return RnsPoly.from_coeff_poly(CoeffPoly(coeffs, q), params)


def trivial_encrypt(m: int, params: BfvParams) -> BfvCiphertext:
    """bootstrap/digit_extract.rs:161-177: ct = (Delta * m, 0), zero noise."""
    from .ring import CoeffPoly
    pt = np.zeros(params.ring_degree, np.uint64)
    pt[0] = m % params.plain_modulus
    return BfvCiphertext([scale_plaintext(CoeffPoly(pt, params.plain_modulus), params), RnsPoly.zero(params)], params)


def bfv_plain_mul(ct: BfvCiphertext, plaintext) -> BfvCiphertext:
    """bfv/eval.rs:468-486: every component times NTT(plaintext) (no Delta scaling)."""
    pt = RnsPoly.from_coeff_poly(plaintext, ct.params)
    return BfvCiphertext([ci.mul(pt) for ci in ct.c], ct.params)


def bfv_plain_add(ct: BfvCiphertext, plaintext) -> BfvCiphertext:
    """bfv/eval.rs:489-504: c0 += Delta * m."""
    c = list(ct.c)
    c[0] = c[0].add(scale_plaintext(plaintext, ct.params))
    return BfvCiphertext(c, ct.params)


def bfv_scalar_mul(ct: BfvCiphertext, scalar: int) -> BfvCiphertext:
    """bootstrap/digit_extract.rs:192-197: multiply by the constant plaintext `scalar mod p`."""
    from .ring import CoeffPoly
    pt = np.zeros(ct.params.ring_degree, np.uint64)
    pt[0] = scalar % ct.params.plain_modulus
    return bfv_plain_mul(ct, CoeffPoly(pt, ct.params.plain_modulus))
