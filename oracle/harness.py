"""CPU ORACLE (test infrastructure) -- input construction and decryption.

Restates, on numpy + Python ints + the oracle NTT, the pieces of the reference
that make *valid inputs* for the hot path and check its outputs semantically
(the reference's tests assert decrypted plaintexts, never ciphertext words):

* params/presets.rs:24-98           -> ``compact_bfv`` / ``compact_dbfv`` / ``u64_dbfv`` / ``cfg3_prime``
* sampling/uniform.rs, gaussian.rs  -> samplers (same distributions, our RNG stream;
  reproducing ChaCha20 + libm ``exp`` bit-for-bit is neither needed nor promised)
* bfv/keygen.rs:64-162              -> ``gen_secret_key`` / ``gen_relin_key``
* bfv/encrypt.rs:79-229             -> ``encrypt_sk`` / ``decrypt``
* dbfv/encrypt.rs:73-118,204-229    -> ``dbfv_encrypt_sk`` / ``dbfv_encrypt_poly_sk``
* dbfv/decrypt.rs:20-79, dbfv/decomposition.rs:45-68 -> ``dbfv_decrypt`` / ``dbfv_decrypt_poly``

Citations are into /root/reference/src/.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import OracleParams, ntt_fwd, ntt_inv


@dataclass(frozen=True)
class DbfvSetup:
    bfv: OracleParams
    base: int
    d: int
    plain_modulus: int      # 0 == 2^64


# ---- presets (params/presets.rs) --------------------------------------------
def compact_bfv() -> OracleParams:                       # :24-35
    return OracleParams(n=1024, q=1099509805057, aux=(562949953443841,), plain_modulus=257,
                        gadget_base=1 << 16)


def compact_dbfv() -> DbfvSetup:                         # :86-98
    bfv = OracleParams(n=1024, q=1099509805057, aux=(562949953443841,), plain_modulus=929,
                       gadget_base=1 << 16)
    return DbfvSetup(bfv, 16, 2, 256)


def u64_dbfv() -> DbfvSetup:                             # :61-75
    bfv = OracleParams(n=4096, q=1152921504606830593,
                       aux=(18014398509998081, 36028797018972161), plain_modulus=1040407,
                       gadget_base=256)
    return DbfvSetup(bfv, 256, 8, 0)


def cfg3_prime() -> DbfvSetup:
    """BASELINE config 3 made runnable: README custom params (README.md:109-119) plus the
    u64 profile's aux primes -- the reference itself errors without them (SURVEY finding 4)."""
    bfv = OracleParams(n=4096, q=576460752308273153,
                       aux=(18014398509998081, 36028797018972161), plain_modulus=65537,
                       gadget_base=1 << 16)
    return DbfvSetup(bfv, 256, 2, 65536)


def toy(n: int = 16, q: int = 1099509805057, aux=(562949953443841,), p: int = 257,
        gadget_base: int = 1 << 16) -> OracleParams:
    return OracleParams(n=n, q=q, aux=tuple(aux), plain_modulus=p, gadget_base=gadget_base)


# ---- samplers (sampling/) -----------------------------------------------------
def sample_uniform(n, q, rng) -> np.ndarray:             # uniform.rs:5-26
    return rng.integers(0, q, size=n, dtype=np.uint64)


def sample_ternary(n, q, rng) -> np.ndarray:             # uniform.rs:29-46
    v = rng.integers(0, 3, size=n)
    return np.where(v == 0, np.uint64(q - 1), np.where(v == 1, np.uint64(0), np.uint64(1))).astype(np.uint64)


def sample_gaussian(n, q, sigma, rng) -> np.ndarray:     # gaussian.rs:15-70 (CDT, tail 6 sigma)
    tail = int(np.ceil(6.0 * sigma))
    xs = np.arange(-tail, tail + 1)
    w = np.exp(-(xs.astype(np.float64) ** 2) / (2.0 * sigma * sigma))
    s = rng.choice(xs, size=n, p=w / w.sum())
    return np.array([int(v) % q for v in s], dtype=np.uint64)


# ---- pointwise helpers (Python ints; 60-bit moduli overflow numpy) -----------
def _mul(a, b, q):
    return np.array((a.astype(object) * b.astype(object)) % q, dtype=np.uint64)


def _add(a, b, q):
    return np.array((a.astype(object) + b.astype(object)) % q, dtype=np.uint64)


def _neg(a, q):
    return np.array((-a.astype(object)) % q, dtype=np.uint64)


# ---- keys (bfv/keygen.rs) ------------------------------------------------------
def gen_secret_key(p: OracleParams, rng) -> np.ndarray:              # :64-80
    return ntt_fwd(sample_ternary(p.n, p.q, rng), p.q)


def gen_relin_key(p: OracleParams, s_ntt: np.ndarray, rng, sigma: float = 3.2) -> np.ndarray:   # :123-162
    q = p.q
    gadget_s_sq = _mul(s_ntt, s_ntt, q)
    keys = np.zeros((p.gadget_digits, 2, p.n), np.uint64)
    for g in range(p.gadget_digits):
        a = ntt_fwd(sample_uniform(p.n, q, rng), q)
        e = ntt_fwd(sample_gaussian(p.n, q, sigma, rng), q)
        keys[g, 0] = _add(_neg(_add(_mul(a, s_ntt, q), e, q), q), gadget_s_sq, q)
        keys[g, 1] = a
        if g + 1 < p.gadget_digits:
            gadget_s_sq = _mul(gadget_s_sq, np.full(p.n, p.gadget_base % q, np.uint64), q)
    return keys


def gen_galois_key(p: OracleParams, s_ntt: np.ndarray, element: int, rng, sigma: float = 3.2) -> np.ndarray:
    """bfv/keygen.rs:170-210: key-switch key from s(X^element) to s(X), [G][2][n] NTT domain."""
    from . import apply_automorphism
    q = p.q
    gadget_s_auto = ntt_fwd(apply_automorphism(ntt_inv(s_ntt, q), q, element), q)
    keys = np.zeros((p.gadget_digits, 2, p.n), np.uint64)
    for g in range(p.gadget_digits):
        a = ntt_fwd(sample_uniform(p.n, q, rng), q)
        e = ntt_fwd(sample_gaussian(p.n, q, sigma, rng), q)
        keys[g, 0] = _add(_neg(_add(_mul(a, s_ntt, q), e, q), q), gadget_s_auto, q)
        keys[g, 1] = a
        if g + 1 < p.gadget_digits:
            gadget_s_auto = _mul(gadget_s_auto, np.full(p.n, p.gadget_base % q, np.uint64), q)
    return keys


# ---- BFV encrypt / decrypt (bfv/encrypt.rs) ------------------------------------
def encrypt_sk(p: OracleParams, pt_coeffs, s_ntt, rng, sigma: float = 3.2) -> np.ndarray:      # :79-106
    q = p.q
    delta = q // p.plain_modulus                                     # :181-229
    m = np.array([(int(v) % q) * delta % q for v in pt_coeffs], dtype=np.uint64)
    delta_m = ntt_fwd(m, q)
    a = ntt_fwd(sample_uniform(p.n, q, rng), q)
    e = ntt_fwd(sample_gaussian(p.n, q, sigma, rng), q)
    c0 = _add(_add(_neg(_mul(a, s_ntt, q), q), e, q), delta_m, q)
    return np.stack([c0, a])


def encode_scalar(p: OracleParams, m: int) -> np.ndarray:            # bfv/encoding.rs:7-19
    assert m < p.plain_modulus
    pt = np.zeros(p.n, np.uint64)
    pt[0] = m
    return pt


def decrypt(p: OracleParams, ct, s_ntt) -> np.ndarray:               # :111-178 (single prime)
    q, t = p.q, p.plain_modulus
    ct = np.asarray(ct, dtype=np.uint64)
    phase = ct[0].copy()
    s_pow = s_ntt
    for i in range(1, ct.shape[0]):
        phase = _add(phase, _mul(ct[i], s_pow, q), q)
        if i < ct.shape[0] - 1:
            s_pow = _mul(s_pow, s_ntt, q)
    x = ntt_inv(phase, q)
    return np.array([((int(v) * t + q // 2) // q) % t for v in x], dtype=np.uint64)


def phase_noise_inf(p: OracleParams, ct, s_ntt, expected_pt) -> int:
    """max |phase - Delta*m| centered -- diagnostic like dbfv/eval.rs:455-519."""
    q, t = p.q, p.plain_modulus
    phase = _add(ct[0], _mul(ct[1], s_ntt, q), q)
    x = ntt_inv(phase, q)
    delta = q // t
    worst = 0
    for v, m in zip(x, expected_pt):
        e = (int(v) - delta * int(m)) % q
        worst = max(worst, min(e, q - e))
    return worst


# ---- dBFV (dbfv/encrypt.rs, dbfv/decrypt.rs, dbfv/decomposition.rs) ---------------
def digit_decompose(value: int, base: int, d: int):                  # decomposition.rs:8-16
    out = []
    for _ in range(d):
        out.append(value % base)
        value //= base
    return out


def dbfv_encrypt_sk(setup: DbfvSetup, value: int, s_ntt, rng) -> np.ndarray:   # encrypt.rs:73-118
    reduced = value % (1 << 64) if setup.plain_modulus == 0 else value % setup.plain_modulus
    limbs = []
    for dig in digit_decompose(reduced, setup.base, setup.d):
        limbs.append(encrypt_sk(setup.bfv, encode_scalar(setup.bfv, dig), s_ntt, rng))
    return np.stack(limbs)


def dbfv_encrypt_poly_sk(setup: DbfvSetup, coeffs, s_ntt, rng) -> np.ndarray:  # encrypt.rs:83-104
    assert setup.plain_modulus != 0
    n = setup.bfv.n
    digs = np.zeros((setup.d, n), np.uint64)
    for i, c in enumerate(coeffs):
        for k, dg in enumerate(digit_decompose(int(c) % setup.plain_modulus, setup.base, setup.d)):
            digs[k, i] = dg
    return np.stack([encrypt_sk(setup.bfv, digs[k], s_ntt, rng) for k in range(setup.d)])


def digit_recompose_signed(digits, base, modulus, bfv_plain_mod) -> int:       # decomposition.rs:45-68
    half_t = bfv_plain_mod // 2
    result, power = 0, 1
    for dg in digits:
        dg = int(dg)
        result += (dg - bfv_plain_mod if dg > half_t else dg) * power
        power *= base
    return result % (1 << 64) if modulus == 0 else result % modulus


def dbfv_decrypt(setup: DbfvSetup, ct, s_ntt) -> int:                # decrypt.rs:20-45
    if setup.plain_modulus != 0:
        return int(dbfv_decrypt_poly(setup, ct, s_ntt)[0])
    digits = [int(decrypt(setup.bfv, limb, s_ntt)[0]) for limb in ct]
    return digit_recompose_signed(digits[:setup.d], setup.base, 0, setup.bfv.plain_modulus)


def dbfv_decrypt_poly(setup: DbfvSetup, ct, s_ntt) -> np.ndarray:    # decrypt.rs:51-79
    t = setup.bfv.plain_modulus
    polys = [decrypt(setup.bfv, limb, s_ntt) for limb in ct[:setup.d]]
    n = setup.bfv.n
    out = np.zeros(n, np.uint64)
    for i in range(n):
        out[i] = digit_recompose_signed([pl[i] for pl in polys], setup.base, setup.plain_modulus, t)
    return out


# ---- Paterson-Stockmeyer evaluation (bootstrap/digit_extract.rs:100-197) on the oracle ------------------
def trivial_encrypt(p: OracleParams, m: int) -> np.ndarray:          # :161-177
    pt = np.zeros(p.n, np.uint64)
    pt[0] = (m % p.plain_modulus) * (p.q // p.plain_modulus) % p.q
    return np.stack([ntt_fwd(pt, p.q), np.zeros(p.n, np.uint64)])


def bfv_scalar_mul(p: OracleParams, ct, scalar: int) -> np.ndarray:  # :192-197 + bfv/eval.rs:468-486
    pt = np.zeros(p.n, np.uint64)
    pt[0] = scalar % p.plain_modulus
    pt_ntt = ntt_fwd(pt, p.q)
    return np.stack([_mul(c, pt_ntt, p.q) for c in ct])


def eval_poly_homomorphic(p: OracleParams, ct_x, poly_coeffs, rlk) -> np.ndarray:   # :100-157
    import math
    from . import bfv_add, bfv_mul_and_relin
    coeffs = [int(c) for c in poly_coeffs]
    d = max(len(coeffs) - 1, 0)
    if d == 0:
        return trivial_encrypt(p, coeffs[0])
    k = max(int(math.ceil(math.sqrt(d + 1.0))), 2)
    baby = [trivial_encrypt(p, 1), np.asarray(ct_x, dtype=np.uint64)]
    for i in range(2, k + 1):
        half = i // 2
        baby.append(bfv_mul_and_relin(p, baby[half], baby[i - half], rlk))
    groups = []
    for i in range((d + k) // k):
        g = trivial_encrypt(p, 0)
        for j in range(k):
            idx = i * k + j
            if idx >= len(coeffs):
                break
            if coeffs[idx] == 0:
                continue
            g = bfv_add(p, g, bfv_scalar_mul(p, baby[j], coeffs[idx]))
        groups.append(g)
    result = groups.pop()
    while groups:
        g = groups.pop()
        result = bfv_add(p, bfv_mul_and_relin(p, result, baby[k], rlk), g)
    return result
