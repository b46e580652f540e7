"""ORACLE (test infrastructure, NOT product code): restatement of the reference's BFV bootstrap
pipeline on numpy arrays, built from the C oracle's ciphertext operations.

    compute_rounding_poly / lagrange_interpolate   bootstrap/digit_extract.rs:19-91
    trivial_encrypt_poly                           bootstrap/digit_extract.rs:180-189
    bfv_plain_mul / bfv_monomial_mul               bfv/eval.rs:468-486, :613-652
    required_trace_elements, naive/shifted trace,
    extract_coefficient, coeffs_to_slots,
    slots_to_coeffs, gen_trace_galois_keys         bootstrap/coeffs_to_slots.rs:21-200
    create_boot_sk, gen_bootstrap_key              bootstrap/bfv_host.rs:49-117, :291-325
    bfv_bootstrap, dbfv_bootstrap,
    dbfv_mul_then_bootstrap, ..._chain_...         bootstrap/bfv_host.rs:134-288

Ciphertexts are [k][n] uint64 arrays in the NTT domain (single ciphertext prime); keys are
[G][2][n].  Pinned by the reference's own bootstrap tests (bootstrap/bfv_host.rs:388-453 and
:455-560), restated in tests/test_oracle.py on the reference's toy parameter sets (n = 16).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from . import (OracleParams, apply_automorphism, bfv_add, bfv_apply_automorphism, bfv_mul_and_relin, dbfv_mul,
               ntt_fwd, ntt_inv)
from . import harness as H


# ---- digit_extract.rs ------------------------------------------------------------------------------
def lagrange_interpolate(values, p: int) -> List[int]:                       # :37-91
    n = len(values)
    if n == 0:
        return []
    if n == 1:
        return [values[0] % p]
    result = [0] * n
    for j in range(n):
        if values[j] % p == 0:
            continue
        num = [0] * n
        num[0] = 1
        deg = 0
        for k in range(n):
            if k == j:
                continue
            neg_k = (-k) % p
            new = [0] * n
            for d in range(deg + 1):
                if d + 1 < n:
                    new[d + 1] = (new[d + 1] + num[d]) % p
                new[d] = (new[d] + num[d] * neg_k) % p
            num = new
            deg += 1
        denom = 1
        for k in range(n):
            if k != j:
                denom = denom * ((j - k) % p) % p
        scale = values[j] % p * pow(denom, -1, p) % p
        for d in range(n):
            result[d] = (result[d] + num[d] * scale) % p
    return result


def compute_rounding_poly(t_orig: int, q_prime: int, t_boot: int) -> List[int]:   # :19-29
    values = [((t_orig * (x % q_prime) + q_prime // 2) // q_prime) % t_orig for x in range(t_boot)]
    return lagrange_interpolate(values, t_boot)


def scale_plaintext(p: OracleParams, coeffs) -> np.ndarray:                  # bfv/encrypt.rs:181-229 (L = 1)
    delta = p.q // p.plain_modulus
    return ntt_fwd(np.array([(int(m) % p.q) * delta % p.q for m in coeffs], dtype=np.uint64), p.q)


def trivial_encrypt_poly(p: OracleParams, coeffs) -> np.ndarray:             # :180-189
    return np.stack([scale_plaintext(p, coeffs), np.zeros(p.n, np.uint64)])


def bfv_plain_mul(p: OracleParams, ct, pt_coeffs) -> np.ndarray:             # bfv/eval.rs:468-486
    pt = ntt_fwd(np.array([int(c) % p.q for c in pt_coeffs], dtype=np.uint64), p.q)
    return np.stack([H._mul(c, pt, p.q) for c in ct])


def bfv_monomial_mul(p: OracleParams, ct, j: int) -> np.ndarray:             # bfv/eval.rs:613-652
    n, q = p.n, p.q
    j %= 2 * n
    if j == 0:
        return np.array(ct, dtype=np.uint64)
    out = []
    for c in ct:
        coeffs = ntt_inv(c, q)
        res = [0] * n
        for i, v in enumerate(coeffs):
            v = int(v)
            if v == 0:
                continue
            idx = (i + j) % (2 * n)
            if idx < n:
                res[idx] = (res[idx] + v) % q
            else:
                res[idx - n] = (res[idx - n] - v) % q
        out.append(ntt_fwd(np.array(res, dtype=np.uint64), q))
    return np.stack(out)


# ---- coeffs_to_slots.rs ------------------------------------------------------------------------------
def required_trace_elements(n: int) -> List[int]:                            # :167-181
    if n <= 32 or n & (n - 1):
        return list(range(3, 2 * n, 2))
    elems, step = [], n
    while step >= 2:
        elems.append(step + 1)
        step >>= 1
    return elems


def gen_trace_galois_keys(p: OracleParams, s_ntt, rng) -> Dict[int, np.ndarray]:   # :184-195
    return {k: H.gen_galois_key(p, s_ntt, k, rng) for k in required_trace_elements(p.n)}


def _trace(p: OracleParams, ct, gks) -> np.ndarray:                          # :55-95
    n = p.n
    if n <= 32 or n & (n - 1):                                               # naive_trace: sum sigma_k(ct)
        result = np.array(ct, dtype=np.uint64)
        for k in range(3, 2 * n, 2):
            if k not in gks:
                raise KeyError(f"missing Galois key for element {k}")
            result = bfv_add(p, result, bfv_apply_automorphism(p, ct, gks[k], k))
        return result
    result = np.array(ct, dtype=np.uint64)                                   # relative-trace chain
    for k in required_trace_elements(n):
        if k not in gks:
            raise KeyError(f"missing Galois key for element {k}")
        result = bfv_add(p, result, bfv_apply_automorphism(p, result, gks[k], k))
    return result


def extract_coefficient(p: OracleParams, ct, j: int, gks) -> np.ndarray:     # :21-50
    n, t = p.n, p.plain_modulus
    shifted = np.array(ct, dtype=np.uint64) if j == 0 else bfv_monomial_mul(p, ct, 2 * n - j)
    result = _trace(p, shifted, gks)
    scale = [0] * n
    scale[0] = pow(n % t, -1, t)
    return bfv_plain_mul(p, result, scale)


def coeffs_to_slots(p: OracleParams, ct, gks) -> List[np.ndarray]:           # :103-116
    return [extract_coefficient(p, ct, j, gks) for j in range(p.n)]


def slots_to_coeffs(p: OracleParams, slots) -> np.ndarray:                   # :122-142
    assert len(slots) == p.n
    result = np.array(slots[0], dtype=np.uint64)
    for j in range(1, p.n):
        result = bfv_add(p, result, bfv_monomial_mul(p, slots[j], j))
    return result


# ---- bfv_host.rs -------------------------------------------------------------------------------------
@dataclass
class BootstrapKeyRef:
    bsk: np.ndarray                      # [2][n] under boot params
    boot: OracleParams
    boot_rlk: np.ndarray                 # [G][2][n]
    galois_keys: Dict[int, np.ndarray] = field(default_factory=dict)
    rounding_poly: List[int] = field(default_factory=list)
    t_orig: int = 0
    q_prime: int = 0


def _center_to(coeffs, q_from: int, m_to: int) -> np.ndarray:
    """c in [0, q_from) read as centred, reduced to [0, m_to) (create_boot_sk :297-311, gen_bootstrap_key :72-90)."""
    out = []
    for c in coeffs:
        c = int(c)
        if c == 0:
            out.append(0)
        elif c <= q_from // 2:
            out.append(c % m_to)
        else:
            out.append(m_to - ((q_from - c) % m_to))
    return np.array(out, dtype=np.uint64)


def create_boot_sk(orig: OracleParams, boot: OracleParams, s_ntt) -> np.ndarray:   # :291-325
    return ntt_fwd(_center_to(ntt_inv(s_ntt, orig.q), orig.q, boot.q), boot.q)


def gen_bootstrap_key(orig: OracleParams, boot: OracleParams, s_ntt, q_prime: int, t_orig: int, rng) -> BootstrapKeyRef:
    """:49-117 (sampling order: bsk encryption, relin key, Galois keys)."""
    assert boot.n == orig.n, "boot params must have same ring degree"
    s_pt = _center_to(ntt_inv(s_ntt, orig.q), orig.q, boot.plain_modulus)
    boot_sk = create_boot_sk(orig, boot, s_ntt)
    bsk = H.encrypt_sk(boot, s_pt, boot_sk, rng)
    boot_rlk = H.gen_relin_key(boot, boot_sk, rng)
    gks = gen_trace_galois_keys(boot, boot_sk, rng)
    return BootstrapKeyRef(bsk, boot, boot_rlk, gks, compute_rounding_poly(t_orig, q_prime, boot.plain_modulus),
                           t_orig, q_prime)


def bfv_bootstrap(orig: OracleParams, ct, bk: BootstrapKeyRef) -> np.ndarray:     # :134-209
    q, qp, boot = orig.q, bk.q_prime, bk.boot
    assert np.asarray(ct).shape[0] == 2, "bootstrap requires degree-1 ciphertext"
    t_boot = boot.plain_modulus
    c0 = ntt_inv(ct[0], q)
    c1 = ntt_inv(ct[1], q)
    sw = lambda v: ((qp * int(v) + q // 2) // q) % qp
    c0p = [sw(v) % t_boot for v in c0]
    c1p = [sw(v) % t_boot for v in c1]
    ct_phase = bfv_add(boot, trivial_encrypt_poly(boot, c0p), bfv_plain_mul(boot, bk.bsk, c1p))
    if not np.asarray(c1).any():                                            # trivial fast path :181-186
        return H.eval_poly_homomorphic(boot, ct_phase, bk.rounding_poly, bk.boot_rlk)
    slots = coeffs_to_slots(boot, ct_phase, bk.galois_keys)
    rounded = [H.eval_poly_homomorphic(boot, s, bk.rounding_poly, bk.boot_rlk) for s in slots]
    return slots_to_coeffs(boot, rounded)


def dbfv_bootstrap(setup: H.DbfvSetup, ct, bk: BootstrapKeyRef):              # :212-236
    limbs = np.stack([bfv_bootstrap(setup.bfv, limb, bk) for limb in ct])
    return H.DbfvSetup(bk.boot, setup.base, setup.d, setup.plain_modulus), limbs


def dbfv_mul_then_bootstrap(setup: H.DbfvSetup, ct1, ct2, rlk, bk: BootstrapKeyRef):   # :242-250
    prod = dbfv_mul(setup.bfv, setup.base, setup.d, setup.plain_modulus, ct1, ct2, rlk)
    return dbfv_bootstrap(setup, prod, bk)


def _same_bfv(a: OracleParams, b: OracleParams) -> bool:
    return a.plain_modulus == b.plain_modulus and a.n == b.n and a.q == b.q


def dbfv_mul_chain_then_bootstrap(cts, rlk, bk: BootstrapKeyRef):                 # :258-288
    """cts: list of (DbfvSetup, array)."""
    assert len(cts) > 0
    acc_setup, acc = cts[0]
    for setup, ct in cts[1:]:
        use_boot = _same_bfv(acc_setup.bfv, bk.boot)
        if not _same_bfv(acc_setup.bfv, setup.bfv):
            setup, ct = dbfv_bootstrap(setup, ct, bk)
        acc_setup, acc = dbfv_mul_then_bootstrap(acc_setup, acc, ct, bk.boot_rlk if use_boot else rlk, bk)
    return acc_setup, acc
