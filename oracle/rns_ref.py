"""CPU ORACLE -- test infrastructure, NOT product code.

Big-integer restatement of the reference's MULTI-PRIME ciphertext multiplication (ct_basis with L > 1 primes):

    bfv_mul_generic_rns            bfv/eval.rs:113-147   exact CRT -> BigInt negacyclic tensor -> round(p t / Q)
    reconstruct_centered_bigint    bfv/eval.rs:719-760
    scale_tensor_component_bigint  bfv/eval.rs:818-831
    centered_bigint_to_rns         bfv/eval.rs:762-790
    relinearize                    bfv/keyswitch.rs:59-101, with the L > 1 branch of RnsPoly::to_coeff_poly
                                   (ring/rns.rs:114-151: u128 CRT, coefficient AND modulus truncated to u64)
    gadget_decompose               bfv/keyswitch.rs:11-52 (on that truncated modulus)
    RnsPoly::from_coeff_poly       ring/rns.rs:84-105 (c % q_l per prime)
    dbfv_mul                       dbfv/eval.rs:82-149 (d^2 products, per-k sums, reduce with all-zero reps)
    keygen / encrypt / decrypt     bfv/keygen.rs:64-162, bfv/encrypt.rs:79-229 (for the reference's decrypt KAT
                                   bfv/eval.rs:903-927)

Polynomials are numpy uint64 arrays [L][n] of NTT-domain residues (one row per ciphertext prime), transforms via the
C oracle's per-prime NTT.  Python integers play num-bigint's role.  Pinned against the reference's own multi-prime
test (n = 16, Q = 65537 * 1099509805057, products 21 / 200 / 0 decrypt correctly) in tests/test_oracle.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

import oracle as O

U64 = (1 << 64) - 1
U128 = (1 << 128) - 1


@dataclass(frozen=True)
class RnsParams:
    n: int
    moduli: Tuple[int, ...]            # ct_basis.moduli
    plain_modulus: int
    gadget_base: int
    gadget_digits: int = 0             # 0 = compute_gadget_digits (params/mod.rs:126-140)

    @property
    def Q(self) -> int:
        q = 1
        for m in self.moduli:
            q *= m
        return q

    @property
    def G(self) -> int:
        if self.gadget_digits:
            return self.gadget_digits
        g, pw = 0, 1
        while pw < self.Q:
            pw *= self.gadget_base
            g += 1
        return max(g, 1)


def to_coeff_components(P: RnsParams, poly: np.ndarray) -> np.ndarray:
    return np.stack([O.ntt_inv(poly[l], P.moduli[l]) for l in range(len(P.moduli))])


def from_coeffs_mod_each(P: RnsParams, coeffs: Sequence[int]) -> np.ndarray:
    """RnsPoly::from_coeff_poly (ring/rns.rs:84-105) for non-negative integer coefficients."""
    return np.stack([O.ntt_fwd(np.array([c % q for c in coeffs], dtype=np.uint64), q) for q in P.moduli])


def reconstruct_centered(P: RnsParams, poly: np.ndarray) -> List[int]:
    """bfv/eval.rs:719-760."""
    comps = to_coeff_components(P, poly)
    Q = P.Q
    terms = []
    for qi in P.moduli:
        q_star = Q // qi
        terms.append(q_star * pow(q_star % qi, -1, qi))
    half = Q >> 1
    out = []
    for j in range(P.n):
        x = sum(t * int(comps[i][j]) for i, t in enumerate(terms)) % Q
        out.append(x - Q if x > half else x)
    return out


def negacyclic_mul(a: Sequence[int], b: Sequence[int], n: int) -> List[int]:
    """poly_mul_bigint (bfv/eval.rs:792-810): exact negacyclic product of signed integer polynomials
    (Kronecker substitution: the same integers as the reference's O(n^2) loop)."""
    bound = max(1, max(abs(x) for x in a)) * max(1, max(abs(x) for x in b)) * n
    k = bound.bit_length() + 2
    A = sum(int(x) << (k * i) for i, x in enumerate(a))
    B = sum(int(x) << (k * i) for i, x in enumerate(b))
    prod = A * B
    full = []
    mask, half = (1 << k) - 1, 1 << (k - 1)
    for _ in range(2 * n - 1):
        c = prod & mask
        if c >= half:
            c -= 1 << k
        full.append(c)
        prod = (prod - c) >> k
    assert prod == 0
    res = full[:n]
    for i in range(n, 2 * n - 1):
        res[i - n] -= full[i]
    return res


def scale_component(P: RnsParams, t: Sequence[int]) -> List[int]:
    """scale_tensor_component_bigint (bfv/eval.rs:818-831): truncating division like BigInt's."""
    Q, p = P.Q, P.plain_modulus
    half = Q >> 1
    return [-(((-p * x) + half) // Q) if p * x < 0 else (p * x + half) // Q for x in t]


def centered_to_rns(P: RnsParams, coeffs: Sequence[int]) -> np.ndarray:
    """centered_bigint_to_rns (bfv/eval.rs:762-790): c % q_l made non-negative, then NTT."""
    return np.stack([O.ntt_fwd(np.array([c % q for c in coeffs], dtype=np.uint64), q) for q in P.moduli])


def bfv_mul_no_relin(P: RnsParams, ct1: np.ndarray, ct2: np.ndarray) -> np.ndarray:
    """bfv_mul_generic_rns (bfv/eval.rs:113-147).  ct [2][L][n] -> [3][L][n]."""
    n = P.n
    c0, c1 = reconstruct_centered(P, ct1[0]), reconstruct_centered(P, ct1[1])
    d0, d1 = reconstruct_centered(P, ct2[0]), reconstruct_centered(P, ct2[1])
    t0 = negacyclic_mul(c0, d0, n)
    t1 = [x + y for x, y in zip(negacyclic_mul(c0, d1, n), negacyclic_mul(c1, d0, n))]
    t2 = negacyclic_mul(c1, d1, n)
    return np.stack([centered_to_rns(P, scale_component(P, t)) for t in (t0, t1, t2)])


def to_coeff_poly_truncated(P: RnsParams, poly: np.ndarray) -> Tuple[List[int], int]:
    """RnsPoly::to_coeff_poly for L > 1 (ring/rns.rs:114-151): u128 CRT, then `val as u64` and
    `big_q as u64`.  Returns (coefficients, modulus) exactly as the reference's CoeffPoly holds them."""
    comps = to_coeff_components(P, poly)
    L = len(P.moduli)
    if L == 1:
        return [int(x) for x in comps[0]], P.moduli[0]
    big_q = 1
    for q in P.moduli:
        big_q = (big_q * q) & U128                      # u128 product (wraps like release-mode Rust)
    q_star_inv = []
    for i, qi in enumerate(P.moduli):                   # RnsBasis::new (ring/rns.rs:46-55)
        prod = 1
        for j, qj in enumerate(P.moduli):
            if i != j:
                prod = prod * (qj % qi) % qi
        q_star_inv.append(pow(prod, -1, qi))
    coeffs = []
    for j in range(P.n):
        val = 0
        for i, qi in enumerate(P.moduli):
            t = int(comps[i][j]) * q_star_inv[i] % qi
            q_star = big_q // qi
            val = ((val + t * q_star) & U128) % big_q
        coeffs.append(val & U64)
    return coeffs, big_q & U64


def _trunc_rem(a: int, b: int) -> int:
    r = abs(a) % abs(b)
    return -r if a < 0 else r


def _trunc_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return -q if (a < 0) != (b < 0) else q


def gadget_decompose(coeffs: Sequence[int], q: int, base: int, num_digits: int) -> List[List[int]]:
    """bfv/keyswitch.rs:11-52 with i128 semantics (truncating % and /)."""
    half_base, half_q = base // 2, q // 2
    digits = [[0] * len(coeffs) for _ in range(num_digits)]
    for pos, c in enumerate(coeffs):
        remaining = c - q if c > half_q else c
        for d in range(num_digits):
            rem = _trunc_rem(remaining, base)
            if rem < -half_base:
                rem += base
            elif rem >= half_base:
                rem -= base
            digits[d][pos] = (_trunc_rem(rem, q) + q) % q
            remaining = _trunc_div(remaining - rem, base)
    return digits


def relinearize(P: RnsParams, ct3: np.ndarray, rlk: np.ndarray) -> np.ndarray:
    """bfv/keyswitch.rs:59-101.  ct3 [3][L][n], rlk [G][2][L][n] -> [2][L][n]."""
    coeffs, modulus = to_coeff_poly_truncated(P, ct3[2])
    digits = gadget_decompose(coeffs, modulus, P.gadget_base, P.G)
    out = ct3[:2].copy()
    for g, dig in enumerate(digits[:rlk.shape[0]]):
        dr = from_coeffs_mod_each(P, dig)
        for l, q in enumerate(P.moduli):
            for c in range(2):
                prod = np.array([int(x) * int(y) % q for x, y in zip(dr[l], rlk[g, c, l])], dtype=np.uint64)
                out[c, l] = (out[c, l] + prod) % np.uint64(q)
    return out


def bfv_mul_and_relin(P: RnsParams, ct1, ct2, rlk) -> np.ndarray:
    """bfv/eval.rs:73-82."""
    return relinearize(P, bfv_mul_no_relin(P, ct1, ct2), rlk)


def dbfv_mul(P: RnsParams, d: int, ct1, ct2, rlk) -> np.ndarray:
    """dbfv/eval.rs:82-149 for p = b^d (all-zero small representatives: limbs k >= d are dropped).
    ct [d][2][L][n] -> [d][2][L][n]."""
    out = np.zeros_like(ct1)
    for i in range(d):
        for j in range(d - i):
            prod = bfv_mul_and_relin(P, ct1[i], ct2[j], rlk)
            for l, q in enumerate(P.moduli):
                out[i + j, :, l] = (out[i + j, :, l] + prod[:, l]) % np.uint64(q)
    return out


# ---- harness for the decrypt KAT (bfv/eval.rs:903-927) --------------------------------------------------
def _mul(P, a, b):
    return np.stack([np.array([int(x) * int(y) % q for x, y in zip(a[l], b[l])], dtype=np.uint64)
                     for l, q in enumerate(P.moduli)])


def _add(P, a, b):
    return np.stack([(a[l] + b[l]) % np.uint64(q) for l, q in enumerate(P.moduli)])


def _neg(P, a):
    return np.stack([(np.uint64(q) - a[l]) % np.uint64(q) for l, q in enumerate(P.moduli)])


def gen_secret_key(P: RnsParams, rng) -> np.ndarray:
    """bfv/keygen.rs:64-80: ternary coefficients stored mod q_0, then reduced per prime (the reference's
    from_coeff_poly maps q_0 - 1 to (q_0 - 1) % q_l, not to -1: restated as is)."""
    from oracle import harness as H
    s = H.sample_ternary(P.n, P.moduli[0], rng)
    return from_coeffs_mod_each(P, [int(x) for x in s])


def gen_relin_key(P: RnsParams, s: np.ndarray, rng, sigma: float = 3.2) -> np.ndarray:
    """bfv/keygen.rs:123-162 -> [G][2][L][n]."""
    from oracle import harness as H
    s_sq = _mul(P, s, s)
    gadget = s_sq
    keys = []
    for i in range(P.G):
        a = from_coeffs_mod_each(P, [int(x) for x in H.sample_uniform(P.n, P.moduli[0], rng)])
        e = from_coeffs_mod_each(P, [int(x) for x in H.sample_gaussian(P.n, P.moduli[0], sigma, rng)])
        rlk0 = _add(P, _neg(P, _add(P, _mul(P, a, s), e)), gadget)
        keys.append(np.stack([rlk0, a]))
        if i + 1 < P.G:
            gadget = np.stack([np.array([int(x) * (P.gadget_base % q) % q for x in gadget[l]], dtype=np.uint64)
                               for l, q in enumerate(P.moduli)])
    return np.stack(keys)


def encrypt_sk(P: RnsParams, pt: Sequence[int], s: np.ndarray, rng, sigma: float = 3.2) -> np.ndarray:
    """bfv/encrypt.rs:79-106 + scale_plaintext :181-229 -> [2][L][n]."""
    from oracle import harness as H
    delta = P.Q // P.plain_modulus
    dm = np.stack([O.ntt_fwd(np.array([(int(m) % q) * (delta % q) % q for m in pt], dtype=np.uint64), q) for q in P.moduli])
    a = from_coeffs_mod_each(P, [int(x) for x in H.sample_uniform(P.n, P.moduli[0], rng)])
    e = from_coeffs_mod_each(P, [int(x) for x in H.sample_gaussian(P.n, P.moduli[0], sigma, rng)])
    c0 = _add(P, _add(P, _neg(P, _mul(P, a, s)), e), dm)
    return np.stack([c0, a])


def decrypt(P: RnsParams, ct: np.ndarray, s: np.ndarray) -> List[int]:
    """bfv/encrypt.rs:111-178."""
    phase, sp = ct[0], s
    for i in range(1, ct.shape[0]):
        phase = _add(P, phase, _mul(P, ct[i], sp))
        if i < ct.shape[0] - 1:
            sp = _mul(P, sp, s)
    comps = to_coeff_components(P, phase)
    Q, p = P.Q, P.plain_modulus
    terms = [(Q // qi) * pow((Q // qi) % qi, -1, qi) for qi in P.moduli]
    half = Q >> 1
    out = []
    for j in range(P.n):
        x = sum(t * int(comps[i][j]) for i, t in enumerate(terms)) % Q
        out.append(((x * p + half) // Q) % p)
    return out
