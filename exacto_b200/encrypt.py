"""bfv/encrypt.rs, bfv/encoding.rs, dbfv/encrypt.rs, dbfv/decrypt.rs: the deterministic halves of
encryption and decryption around the ciphertext-multiplication path, so whole pipelines stay on
the GPU.  Sampling stays with the caller: the reference draws (a, e) from ChaCha20 + a Gaussian CDT
(sampling/*.rs) whose exact stream is outside this path, so the encrypt functions here take the
sampled polynomials explicitly and are bit-exact functions of them.

    SecretKey                          bfv/keygen.rs:13-17
    encode_scalar / decode_scalar      bfv/encoding.rs:7-24
    encrypt_sk_with_samples            bfv/encrypt.rs:79-106 given (a, e)
    decrypt                            bfv/encrypt.rs:111-178 (decrypt_kernel)
    dbfv_encrypt_sk_with_samples, dbfv_encrypt_poly_sk_with_samples   dbfv/encrypt.rs:73-118
    dbfv_decrypt / dbfv_decrypt_poly   dbfv/decrypt.rs:20-79
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import _native
from .bfv import BfvCiphertext, scale_plaintext
from .dbfv import DbfvCiphertext
from .error import InvalidParam
from .params import BfvParams, DbfvParams
from .ring import CoeffPoly, RnsPoly, _ptr, _u64


class SecretKey:
    """bfv/keygen.rs:13-17: ``poly`` is s in RNS-NTT form."""

    def __init__(self, poly: RnsPoly, params: BfvParams):
        self.poly = poly
        self.params = params

    @staticmethod
    def from_coeffs(coeffs, params: BfvParams) -> "SecretKey":
        """Ternary (or any) coefficients already reduced mod q, as gen_secret_key stores them (:64-80)."""
        q = params.ct_basis.moduli[0]
        return SecretKey(RnsPoly.from_coeff_poly(CoeffPoly(_u64(coeffs), q), params), params)

    @staticmethod
    def from_ntt(evals, params: BfvParams) -> "SecretKey":
        ct = BfvCiphertext.from_array(_u64(evals)[None], params)
        return SecretKey(ct.c[0], params)

    def ntt_array(self) -> np.ndarray:
        """[n] for a single ciphertext prime, [L][n] otherwise."""
        if len(self.poly.components) == 1:
            return self.poly.components[0].evals
        return np.stack([c.evals for c in self.poly.components])


def encode_scalar(m: int, params: BfvParams) -> CoeffPoly:
    """bfv/encoding.rs:7-19."""
    if m >= params.plain_modulus:
        raise InvalidParam(f"plaintext {m} >= plain_modulus {params.plain_modulus}")
    coeffs = np.zeros(params.ring_degree, np.uint64)
    coeffs[0] = m
    return CoeffPoly(coeffs, params.plain_modulus)


def decode_scalar(poly: CoeffPoly) -> int:
    """bfv/encoding.rs:22-24."""
    return int(poly.coeffs[0])


def encrypt_sk_with_samples(plaintext: CoeffPoly, sk: SecretKey, params: BfvParams, a: CoeffPoly,
                            e: CoeffPoly) -> BfvCiphertext:
    """bfv/encrypt.rs:79-106 with the caller's samples: ct = (-a s + e + Delta m, a); ``a`` uniform and
    ``e`` Gaussian coefficient polynomials mod q (what sample_uniform_poly / sample_gaussian_poly return)."""
    delta_m = scale_plaintext(plaintext, params)
    a_ntt = RnsPoly.from_coeff_poly(a, params)
    e_ntt = RnsPoly.from_coeff_poly(e, params)
    c0 = a_ntt.mul(sk.poly).neg().add(e_ntt).add(delta_m)
    return BfvCiphertext([c0, a_ntt], params)


def encrypt_sk_with_sampler(plaintext: CoeffPoly, sk: SecretKey, params: BfvParams, sampler) -> BfvCiphertext:
    """bfv/encrypt.rs:79-106 drawing (a, e) from ``sampler`` in the reference's order (see keygen.py)."""
    q, n = params.ct_basis.moduli[0], params.ring_degree
    a = CoeffPoly(sampler.uniform(n, q), q)
    e = CoeffPoly(sampler.gaussian(n, q, params.sigma), q)
    return encrypt_sk_with_samples(plaintext, sk, params, a, e)


def encrypt_pk_with_sampler(plaintext: CoeffPoly, pk, params: BfvParams, sampler) -> BfvCiphertext:
    """bfv/encrypt.rs:29-64: ct = (pk0 u + e1 + Delta m, pk1 u + e2), u binary; sampling order u, e1, e2."""
    q, n = params.ct_basis.moduli[0], params.ring_degree
    delta_m = scale_plaintext(plaintext, params)
    u = RnsPoly.from_coeff_poly(CoeffPoly(sampler.binary(n, q), q), params)
    e1 = RnsPoly.from_coeff_poly(CoeffPoly(sampler.gaussian(n, q, params.sigma), q), params)
    e2 = RnsPoly.from_coeff_poly(CoeffPoly(sampler.gaussian(n, q, params.sigma), q), params)
    return BfvCiphertext([pk.pk0.mul(u).add(e1).add(delta_m), pk.pk1.mul(u).add(e2)], params)


def decrypt(ct: BfvCiphertext, sk: SecretKey) -> CoeffPoly:
    """bfv/encrypt.rs:111-178: m = round(p (c0 + c1 s + c2 s^2 + ...) / q) mod p, any ciphertext degree."""
    params = ct.params
    out = decrypt_batch(params, ct.to_array()[None], sk)
    return CoeffPoly(out[0], params.plain_modulus)


def decrypt_batch(params: BfvParams, ct: np.ndarray, sk: SecretKey, device: Optional[int] = None) -> np.ndarray:
    """Batched host-buffer form: ct [B][k][n] -> plaintext coefficients [B][n] (exb_bfv_decrypt_host)."""
    ct = _u64(ct)
    Lq = params.ct_basis.num_moduli()                    # multi-prime ciphertexts are [batch][components][L][n]
    tail = (params.ring_degree,) if Lq == 1 else (Lq, params.ring_degree)
    if ct.ndim != 2 + len(tail) or ct.shape[2:] != tail or ct.shape[1] < 1:
        raise InvalidParam("decrypt expects [batch][components][n]")
    ctx = params.context(device)
    out = np.empty((ct.shape[0], params.ring_degree), np.uint64)
    s = np.ascontiguousarray(sk.ntt_array())
    _native.check(_native.lib().exb_bfv_decrypt_host(ctx.handle, _ptr(ct), ct.shape[1], _ptr(s), _ptr(out), ct.shape[0]))
    return out


# ---- dBFV ------------------------------------------------------------------------------------------
def digit_decompose(value: int, base: int, num_digits: int) -> List[int]:
    """dbfv/decomposition.rs:8-16."""
    out = []
    for _ in range(num_digits):
        out.append(value % base)
        value //= base
    return out


def digit_recompose_signed(digits: Sequence[int], base: int, modulus: int, bfv_plain_mod: int) -> int:
    """dbfv/decomposition.rs:45-68 (modulus 0 = 2^64)."""
    half_t = bfv_plain_mod // 2
    result, power = 0, 1
    for dg in digits:
        dg = int(dg)
        result += (dg - bfv_plain_mod if dg > half_t else dg) * power
        power *= base
    return result % (1 << 64) if modulus == 0 else result % modulus


def _encrypt_digit_polys(digit_polys, sk: SecretKey, params: DbfvParams, a_samples, e_samples) -> DbfvCiphertext:
    """dbfv/encrypt.rs encrypt_sk_digit_polys: one BFV encryption per digit polynomial."""
    bfv = params.bfv_params
    if len(a_samples) != len(digit_polys) or len(e_samples) != len(digit_polys):
        raise InvalidParam("need one (a, e) sample pair per digit")
    limbs = [encrypt_sk_with_samples(CoeffPoly(dp, bfv.plain_modulus), sk, bfv, a, e)
             for dp, a, e in zip(digit_polys, a_samples, e_samples)]
    return DbfvCiphertext(limbs, params.num_digits, 0, params)


def dbfv_encrypt_sk_with_samples(plaintext: int, sk: SecretKey, params: DbfvParams, a_samples, e_samples) -> DbfvCiphertext:
    """dbfv/encrypt.rs:73-82 + digit_decompose_scalar (:106-118)."""
    reduced = plaintext % (1 << 64) if params.plain_modulus == 0 else plaintext % params.plain_modulus
    n = params.bfv_params.ring_degree
    polys = []
    for dg in digit_decompose(reduced, params.base, params.num_digits):
        c = np.zeros(n, np.uint64)
        c[0] = dg
        polys.append(c)
    return _encrypt_digit_polys(polys, sk, params, a_samples, e_samples)


def dbfv_encrypt_poly_sk_with_samples(plaintext: CoeffPoly, sk: SecretKey, params: DbfvParams, a_samples,
                                      e_samples) -> DbfvCiphertext:
    """dbfv/encrypt.rs:94-104: coefficient-wise base-b digits of a Z_p[X]/(X^n+1) plaintext."""
    if params.plain_modulus == 0:
        raise InvalidParam("polynomial dBFV plaintext requires finite plain_modulus (plain_modulus=0 is scalar-only)")
    n = params.bfv_params.ring_degree
    if len(plaintext) != n:
        raise InvalidParam("plaintext length must equal the ring degree")
    polys = np.zeros((params.num_digits, n), np.uint64)
    for i, c in enumerate(plaintext.coeffs):
        for k, dg in enumerate(digit_decompose(int(c) % params.plain_modulus, params.base, params.num_digits)):
            polys[k, i] = dg
    return _encrypt_digit_polys(list(polys), sk, params, a_samples, e_samples)


def _digit_polys(value_or_poly, params: DbfvParams, poly: bool):
    n = params.bfv_params.ring_degree
    if not poly:
        reduced = value_or_poly % (1 << 64) if params.plain_modulus == 0 else value_or_poly % params.plain_modulus
        out = []
        for dg in digit_decompose(reduced, params.base, params.num_digits):
            c = np.zeros(n, np.uint64)
            c[0] = dg
            out.append(c)
        return out
    if params.plain_modulus == 0:
        raise InvalidParam("polynomial dBFV plaintext requires finite plain_modulus (plain_modulus=0 is scalar-only)")
    if len(value_or_poly) != n:
        raise InvalidParam("plaintext length must equal the ring degree")
    polys = np.zeros((params.num_digits, n), np.uint64)
    for i, c in enumerate(value_or_poly.coeffs):
        for k, dg in enumerate(digit_decompose(int(c) % params.plain_modulus, params.base, params.num_digits)):
            polys[k, i] = dg
    return list(polys)


def dbfv_encrypt_sk_with_sampler(plaintext, sk: SecretKey, params: DbfvParams, sampler, poly: bool = False) -> DbfvCiphertext:
    """dbfv/encrypt.rs:73-104 (scalar, or a Z_p[X]/(X^n+1) CoeffPoly with poly=True): one BFV encryption per digit."""
    bfv = params.bfv_params
    limbs = [encrypt_sk_with_sampler(CoeffPoly(dp, bfv.plain_modulus), sk, bfv, sampler)
             for dp in _digit_polys(plaintext, params, poly)]
    return DbfvCiphertext(limbs, params.num_digits, 0, params)


def dbfv_encrypt_with_sampler(plaintext, pk, params: DbfvParams, sampler, poly: bool = False) -> DbfvCiphertext:
    """dbfv/encrypt.rs:17-60: public-key encryption of every digit polynomial."""
    bfv = params.bfv_params
    limbs = [encrypt_pk_with_sampler(CoeffPoly(dp, bfv.plain_modulus), pk, bfv, sampler)
             for dp in _digit_polys(plaintext, params, poly)]
    return DbfvCiphertext(limbs, params.num_digits, 0, params)


def _decrypt_limbs(ct: DbfvCiphertext, sk: SecretKey, count: int) -> np.ndarray:
    limbs = ct.limbs[:count]
    if all(len(l.c) == len(limbs[0].c) for l in limbs):          # one batched launch
        return decrypt_batch(ct.params.bfv_params, np.stack([l.to_array() for l in limbs]), sk)
    return np.stack([decrypt(l, sk).coeffs for l in limbs])


def dbfv_decrypt(ct: DbfvCiphertext, sk: SecretKey) -> int:
    """dbfv/decrypt.rs:20-45."""
    params = ct.params
    if params.plain_modulus != 0:
        return decode_scalar(dbfv_decrypt_poly(ct, sk))
    digits = [int(v) for v in _decrypt_limbs(ct, sk, ct.num_limbs())[:, 0]]
    use = min(params.num_digits, len(digits))
    return digit_recompose_signed(digits[:use], params.base, params.plain_modulus, params.bfv_params.plain_modulus)


def dbfv_decrypt_poly(ct: DbfvCiphertext, sk: SecretKey) -> CoeffPoly:
    """dbfv/decrypt.rs:51-79."""
    params = ct.params
    if params.plain_modulus == 0:
        raise InvalidParam("polynomial dBFV decrypt requires finite plain_modulus (plain_modulus=0 is scalar-only)")
    t = params.bfv_params.plain_modulus
    use = min(params.num_digits, ct.num_limbs())
    polys = _decrypt_limbs(ct, sk, use)
    n = params.bfv_params.ring_degree
    out = np.zeros(n, np.uint64)
    for i in range(n):
        out[i] = digit_recompose_signed([polys[k, i] for k in range(use)], params.base, params.plain_modulus, t)
    return CoeffPoly(out, params.plain_modulus)
