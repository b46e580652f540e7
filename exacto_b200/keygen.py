"""bfv/keygen.rs: key generation around the multiplication path.  Sampling stays with the caller
(the reference's ChaCha20 / Gaussian-CDT stream, sampling/*.rs, is outside this path): every function
takes a ``sampler`` with

    sampler.ternary(n, q)          -> coefficients in {0, 1, q-1}     (sampling/uniform.rs:29-46)
    sampler.uniform(n, q)          -> coefficients uniform in [0, q)  (sampling/uniform.rs:5-26)
    sampler.gaussian(n, q, sigma)  -> centred Gaussian mod q          (sampling/gaussian.rs:15-70)
    sampler.binary(n, q)           -> coefficients in {0, 1}          (sampling/uniform.rs, encrypt_pk only)

and calls it in the reference's order, so the keys are bit-exact functions of the sampled
polynomials; all transforms and products run on the GPU.

    gen_secret_key_with_sampler   bfv/keygen.rs:64-80
    gen_public_key_with_sampler   bfv/keygen.rs:89-121
    gen_relin_key_with_sampler    bfv/keygen.rs:123-162
    gen_galois_key_with_sampler   bfv/keygen.rs:170-210
    apply_automorphism            bfv/keygen.rs:218-239 (coefficient polynomial, host)
"""
from __future__ import annotations

import numpy as np

from .bfv import GaloisKey, RelinKey
from .encrypt import SecretKey
from .params import BfvParams
from .ring import CoeffPoly, RnsPoly


def apply_automorphism(poly: CoeffPoly, k: int) -> CoeffPoly:
    """X^i -> X^(i k) mod (X^n + 1): signed scatter with mod-q accumulation (zero coefficients skipped)."""
    n, q = len(poly), poly.modulus
    out = [0] * n
    for i, c in enumerate(poly.coeffs):
        c = int(c)
        if c == 0:
            continue
        e = (i * k) % (2 * n)
        if e < n:
            out[e] = (out[e] + c) % q
        else:
            out[e - n] = (out[e - n] - c) % q
    return CoeffPoly(np.array(out, dtype=np.uint64), q)


def gen_secret_key_with_sampler(params: BfvParams, sampler) -> SecretKey:
    q = params.ct_basis.moduli[0]
    return SecretKey.from_coeffs(sampler.ternary(params.ring_degree, q), params)


class PublicKey:
    """bfv/keygen.rs:30-34: (pk0, pk1) = (-(a s + e), a)."""

    def __init__(self, pk0: RnsPoly, pk1: RnsPoly, params: BfvParams):
        self.pk0, self.pk1, self.params = pk0, pk1, params


def gen_public_key_with_sampler(sk: SecretKey, sampler) -> PublicKey:
    params = sk.params
    n, q = params.ring_degree, params.ct_basis.moduli[0]
    a = RnsPoly.from_coeff_poly(CoeffPoly(sampler.uniform(n, q), q), params)
    e = RnsPoly.from_coeff_poly(CoeffPoly(sampler.gaussian(n, q, params.sigma), q), params)
    return PublicKey(a.mul(sk.poly).add(e).neg(), a, params)


def _key_switch_key(sk: SecretKey, target: RnsPoly, sampler):
    """keys[i] = (-(a_i s + e_i) + base^i * target, a_i): the loop shared by :138-155 and :186-203."""
    params = sk.params
    n, q = params.ring_degree, params.ct_basis.moduli[0]
    keys, gadget_t = [], target
    for i in range(params.gadget_digits):
        a = RnsPoly.from_coeff_poly(CoeffPoly(sampler.uniform(n, q), q), params)
        e = RnsPoly.from_coeff_poly(CoeffPoly(sampler.gaussian(n, q, params.sigma), q), params)
        k0 = a.mul(sk.poly).add(e).neg().add(gadget_t)
        keys.append((k0, a))
        if i + 1 < params.gadget_digits:
            gadget_t = gadget_t.scalar_mul(params.gadget_base)
    return keys


def gen_relin_key_with_sampler(sk: SecretKey, sampler) -> RelinKey:
    return RelinKey(_key_switch_key(sk, sk.poly.mul(sk.poly), sampler), sk.params)


def gen_galois_key_with_sampler(sk: SecretKey, element: int, sampler) -> GaloisKey:
    params = sk.params
    s_auto = apply_automorphism(sk.poly.to_coeff_poly(), element)
    return GaloisKey(_key_switch_key(sk, RnsPoly.from_coeff_poly(s_auto, params), sampler), element, params)
