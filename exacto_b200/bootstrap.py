"""bootstrap/: the multiplication entry points north_star names (bootstrap/bfv_host.rs:242-288) and the
bootstrap body they call, as a sequence of GPU primitives of this package:

    bfv_bootstrap                bootstrap/bfv_host.rs:134-209   modulus switch (the decrypt kernel on a context with
                                 plaintext modulus q'), re-encryption of the phase (point-wise kernels), CoeffsToSlots (automorphism + key-switch kernel,
                                 batched over the n shifted copies), rounding polynomial (batched
                                 Paterson-Stockmeyer over all slots: bfv_mul_and_relin), SlotsToCoeffs
    dbfv_bootstrap               bootstrap/bfv_host.rs:212-236   all limbs in one batch; params swap, mul_depth = 0
    coeffs_to_slots.rs           required_trace_elements, extract_coefficient, coeffs_to_slots, slots_to_coeffs
    digit_extract.rs             compute_rounding_poly, lagrange_interpolate, trivial_encrypt_poly,
                                 eval_poly_homomorphic
    gen_bootstrap_key_with_sampler / create_boot_sk   bootstrap/bfv_host.rs:49-117, :291-325

The reference exercises this body at toy scale only (n = 16, bootstrap/bfv_host.rs:345-560); the same code
runs at any ring degree the library supports.  X^j multiplications are done in the NTT domain (a point-wise
product with NTT(+-X^(j mod n))), which is the same element of Z_q[X]/(X^n+1) as the reference's coefficient
rotation (bfv/eval.rs:613-652), hence the same canonical words.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import math

import numpy as np

from .bfv import (BfvCiphertext, GaloisKey, RelinKey, bfv_add, bfv_apply_automorphism, bfv_mul_and_relin, bfv_plain_mul,
                  bfv_scalar_mul, scale_plaintext, trivial_encrypt)
from .dbfv import DbfvCiphertext, dbfv_mul
from .error import InvalidParam, NotImplementedErr
from .params import BfvParams, DbfvParams
from .ring import CoeffPoly, RnsPoly


class BootstrapKey:
    """bootstrap/bfv_host.rs:19-41.  ``bootstrap`` (optional) replaces the built-in dbfv_bootstrap body with a
    caller-supplied refresh (e.g. a remote bootstrapping service); without key material and without it,
    dbfv_bootstrap fails loudly."""

    def __init__(self, boot_params: BfvParams, boot_rlk: RelinKey,
                 bootstrap: Optional[Callable[[DbfvCiphertext], DbfvCiphertext]] = None, *,
                 bsk: Optional[BfvCiphertext] = None, galois_keys: Optional[Dict[int, GaloisKey]] = None,
                 rounding_poly: Optional[Sequence[int]] = None, t_orig: int = 0, q_prime: int = 0):
        self.boot_params = boot_params
        self.boot_rlk = boot_rlk
        self.bootstrap = bootstrap
        self.bsk = bsk
        self.galois_keys = galois_keys or {}
        self.rounding_poly = list(rounding_poly) if rounding_poly is not None else None
        self.t_orig = t_orig
        self.q_prime = q_prime


# ---- digit_extract.rs -----------------------------------------------------------------------------------
def lagrange_interpolate(values: Sequence[int], p: int) -> List[int]:
    """digit_extract.rs:37-91: coefficients of the interpolant through (i, values[i]), p prime."""
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


def compute_rounding_poly(t_orig: int, q_prime: int, t_boot: int) -> List[int]:
    """digit_extract.rs:19-29: x -> round(t_orig * (x mod q') / q') mod t_orig on Z_{t_boot}."""
    values = [((t_orig * (x % q_prime) + q_prime // 2) // q_prime) % t_orig for x in range(t_boot)]
    return lagrange_interpolate(values, t_boot)


def trivial_encrypt_poly(poly: CoeffPoly, params: BfvParams) -> BfvCiphertext:
    """digit_extract.rs:180-189: (Delta * m(X), 0)."""
    return BfvCiphertext([scale_plaintext(poly, params), RnsPoly.zero(params)], params)


# ---- bfv/eval.rs:613-652 and coeffs_to_slots.rs -----------------------------------------------------------
def _monomial(j: int, params: BfvParams) -> RnsPoly:
    """NTT of X^j mod (X^n + 1) over q (j reduced mod 2n; X^n = -1)."""
    n, q = params.ring_degree, params.ct_basis.moduli[0]
    j %= 2 * n
    c = np.zeros(n, np.uint64)
    c[j % n] = 1 if j < n else q - 1
    return RnsPoly.from_coeff_poly(CoeffPoly(c, q), params)


def bfv_monomial_mul(ct: BfvCiphertext, j: int) -> BfvCiphertext:
    """bfv/eval.rs:613-634: every component times X^j."""
    n = ct.params.ring_degree
    if j % (2 * n) == 0:
        return ct
    mono = _monomial(j, ct.params)
    return BfvCiphertext([ci.mul(mono) for ci in ct.c], ct.params)


def required_trace_elements(n: int) -> List[int]:
    """coeffs_to_slots.rs:167-181."""
    if n <= 32 or n & (n - 1):
        return list(range(3, 2 * n, 2))
    elems, step = [], n
    while step >= 2:
        elems.append(step + 1)
        step >>= 1
    return elems


def _gk(galois_keys, k: int) -> GaloisKey:
    gk = galois_keys.get(k)
    if gk is None:
        raise InvalidParam(f"missing Galois key for element {k}")
    return gk


def _trace(ct: BfvCiphertext, galois_keys) -> BfvCiphertext:
    """coeffs_to_slots.rs:55-95 (naive sum for n <= 32, relative-trace chain otherwise)."""
    n = ct.params.ring_degree
    result = ct
    if n <= 32 or n & (n - 1):
        for k in range(3, 2 * n, 2):
            result = bfv_add(result, bfv_apply_automorphism(ct, _gk(galois_keys, k)))
        return result
    for k in required_trace_elements(n):
        result = bfv_add(result, bfv_apply_automorphism(result, _gk(galois_keys, k)))
    return result


def extract_coefficient(ct: BfvCiphertext, j: int, galois_keys) -> BfvCiphertext:
    """coeffs_to_slots.rs:21-50: X^-j, trace, n^-1."""
    params = ct.params
    n, t = params.ring_degree, params.plain_modulus
    shifted = ct if j == 0 else bfv_monomial_mul(ct, 2 * n - j)
    result = _trace(shifted, galois_keys)
    try:
        n_inv = pow(n % t, -1, t)
    except ValueError:
        raise InvalidParam("n not invertible mod t")
    return bfv_scalar_mul(result, n_inv)


def coeffs_to_slots(ct: BfvCiphertext, galois_keys) -> List[BfvCiphertext]:
    """coeffs_to_slots.rs:103-116."""
    return [extract_coefficient(ct, j, galois_keys) for j in range(ct.params.ring_degree)]


def slots_to_coeffs(slots: Sequence[BfvCiphertext]) -> BfvCiphertext:
    """coeffs_to_slots.rs:122-142: sum_j X^j * slot_j."""
    if len(slots) == 0:
        raise InvalidParam("empty slots")
    n = slots[0].params.ring_degree
    if len(slots) != n:
        raise InvalidParam(f"expected {n} slots, got {len(slots)}")
    result = slots[0]
    for j in range(1, n):
        result = bfv_add(result, bfv_monomial_mul(slots[j], j))
    return result


# ---- bfv_host.rs ------------------------------------------------------------------------------------------
def _center_to(coeffs, q_from: int, m_to: int) -> np.ndarray:
    """c in [0, q_from) read as centred and reduced mod m_to (bfv_host.rs:72-90, :297-311).  Like the reference, a
    negative value that is a multiple of m_to comes out as m_to itself, not 0: create_boot_sk reduces it in
    RnsPoly::from_coeff_poly (c % q), and the plaintext copy goes through scale_plaintext's `m % q_i`
    (bfv/encrypt.rs:193) exactly as upstream.  Ternary keys never reach that branch."""
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


def create_boot_sk(sk, boot_params: BfvParams):
    """bfv_host.rs:291-325: the same polynomial s in the boot NTT domain."""
    from .encrypt import SecretKey
    q = sk.params.ct_basis.moduli[0]
    qb = boot_params.ct_basis.moduli[0]
    return SecretKey.from_coeffs(_center_to(sk.poly.to_coeff_poly().coeffs, q, qb), boot_params)


def gen_bootstrap_key_with_sampler(sk, boot_params: BfvParams, q_prime: int, t_orig: int, sampler) -> BootstrapKey:
    """bfv_host.rs:49-117 with the caller's sampler (see keygen.py); sampling order: bsk, relin key, Galois keys."""
    from .encrypt import encrypt_sk_with_samples
    from .keygen import gen_galois_key_with_sampler, gen_relin_key_with_sampler
    n = sk.params.ring_degree
    if boot_params.ring_degree != n:
        raise InvalidParam("boot params must have same ring degree")
    q, qb, tb = sk.params.ct_basis.moduli[0], boot_params.ct_basis.moduli[0], boot_params.plain_modulus
    s_pt = CoeffPoly(_center_to(sk.poly.to_coeff_poly().coeffs, q, tb), tb)
    boot_sk = create_boot_sk(sk, boot_params)
    a = CoeffPoly(sampler.uniform(n, qb), qb)
    e = CoeffPoly(sampler.gaussian(n, qb, boot_params.sigma), qb)
    bsk = encrypt_sk_with_samples(s_pt, boot_sk, boot_params, a, e)
    boot_rlk = gen_relin_key_with_sampler(boot_sk, sampler)
    gks = {k: gen_galois_key_with_sampler(boot_sk, k, sampler) for k in required_trace_elements(n)}
    return BootstrapKey(boot_params, boot_rlk, bsk=bsk, galois_keys=gks,
                        rounding_poly=compute_rounding_poly(t_orig, q_prime, tb), t_orig=t_orig, q_prime=q_prime)


_MODSWITCH_PARAMS: Dict[tuple, BfvParams] = {}


def _modswitch_params(n: int, q: int, q_prime: int) -> BfvParams:
    """Context (n, q, plaintext modulus q') whose decrypt kernel performs the modulus switch q -> q'."""
    key = (n, q, q_prime)
    if key not in _MODSWITCH_PARAMS:
        from .params import BfvParamsBuilder
        _MODSWITCH_PARAMS[key] = BfvParamsBuilder().ring_degree(n).plain_modulus(q_prime).ct_moduli([q]).build()
    return _MODSWITCH_PARAMS[key]


def _bootstrap_batch(orig: BfvParams, cts: np.ndarray, bsk: BootstrapKey) -> np.ndarray:
    """bfv_bootstrap on a batch [B][2][n] of degree-1 ciphertexts under ``orig`` -> [B][2][n] under the boot
    parameters; every step runs on the GPU."""
    import torch
    from . import batch
    boot = bsk.boot_params
    n, q, qp = orig.ring_degree, orig.ct_basis.moduli[0], bsk.q_prime
    qb, tb = boot.ct_basis.moduli[0], boot.plain_modulus
    B = cts.shape[0]
    dev = batch.to_device(cts)
    # modulus switch q -> q' (:147-171): c' = floor((q' c + floor(q/2)) / q) mod q' on the coefficients of c0 and c1.
    # That is exactly the scale-and-round step of decrypt with plaintext modulus q' on a one-component
    # "ciphertext", so the decrypt kernel does it (INTT + Shoup scale) on a context (n, q, p = q').
    ms = _modswitch_params(n, q, qp)
    zero_key = torch.zeros(n, dtype=torch.int64, device=dev.device)
    sw = batch.bfv_decrypt(ms, dev.reshape(B * 2, 1, n), zero_key).reshape(B, 2, n)      # values in [0, q')
    if tb < qp:
        sw = torch.remainder(sw, tb)                                                       # `% t_boot` :163-171
    trivial = [not cts[b, 1].any() for b in range(B)]                            # :179 (c1 = 0 iff its NTT is 0)
    delta = qb // tb
    c0n = batch.ntt_forward(boot, 0, batch.poly_scalar_mul(boot, 0, sw[:, 0].contiguous(), delta))   # scale_plaintext
    c1n = batch.ntt_forward(boot, 0, sw[:, 1].contiguous())
    bsk_dev = batch.to_device(np.broadcast_to(bsk.bsk.to_array(), (B, 2, n)).copy())
    c1n2 = torch.stack([c1n, c1n], dim=1).contiguous()
    phase = batch.poly_mul(boot, 0, bsk_dev, c1n2)                               # bfv_plain_mul(bsk, c1')  :174
    phase[:, 0] = batch.poly_add(boot, 0, phase[:, 0].contiguous(), c0n)         # + trivial_encrypt_poly(c0') :173-175
    out = torch.empty_like(phase)
    t_idx = [b for b in range(B) if trivial[b]]
    r_idx = [b for b in range(B) if not trivial[b]]
    if t_idx:                                                                    # fast path :181-186
        out[t_idx] = eval_poly_homomorphic_batch(boot, phase[t_idx].contiguous(), bsk.rounding_poly, bsk.boot_rlk)
    if r_idx:                                                                    # full ring path :188-206
        R = len(r_idx)
        ph = phase[r_idx].contiguous()                                           # [R][2][n]
        # monomial tables are built once, uploaded once ([n][n]) and broadcast on the device
        mono_in = batch.to_device(np.stack([_monomial(2 * n - j, boot).components[0].evals for j in range(n)]))    # X^-j
        mono_out = batch.to_device(np.stack([_monomial(j, boot).components[0].evals for j in range(n)]))           # X^j
        def times(x, mono):                                                      # x [R][n][2][n] * mono[j]
            return batch.poly_mul(boot, 0, x, mono[None, :, None, :].expand(R, n, 2, n).contiguous())
        shifted = times(ph[:, None].expand(R, n, 2, n).contiguous(), mono_in)
        flat = shifted.reshape(R * n, 2, n)
        gks = bsk.galois_keys
        if n <= 32 or n & (n - 1):                                               # naive_trace
            acc = flat.clone()
            for k in range(3, 2 * n, 2):
                acc = batch.poly_add(boot, 0, acc, batch.bfv_apply_automorphism(boot, flat, _gk(gks, k)))
        else:                                                                    # relative-trace chain
            acc = flat
            for k in required_trace_elements(n):
                acc = batch.poly_add(boot, 0, acc, batch.bfv_apply_automorphism(boot, acc, _gk(gks, k)))
        try:
            n_inv = pow(n % tb, -1, tb)
        except ValueError:
            raise InvalidParam("n not invertible mod t")
        slots = batch.poly_scalar_mul(boot, 0, acc, n_inv)
        rounded = eval_poly_homomorphic_batch(boot, slots, bsk.rounding_poly, bsk.boot_rlk)
        packed = times(rounded.reshape(R, n, 2, n), mono_out)
        res = packed[:, 0].contiguous()
        for j in range(1, n):
            res = batch.poly_add(boot, 0, res, packed[:, j].contiguous())
        out[r_idx] = res
    return batch.to_host(out)


def bfv_bootstrap(ct: BfvCiphertext, bsk: BootstrapKey) -> BfvCiphertext:
    """bootstrap/bfv_host.rs:134-209."""
    if len(ct.c) != 2:
        raise InvalidParam("bootstrap requires degree-1 ciphertext")
    if bsk.bsk is None or bsk.rounding_poly is None:
        raise NotImplementedErr("BootstrapKey carries no key material (bsk / rounding_poly)")
    out = _bootstrap_batch(ct.params, ct.to_array()[None], bsk)
    return BfvCiphertext.from_array(out[0], bsk.boot_params)


def dbfv_bootstrap(ct: DbfvCiphertext, bsk: BootstrapKey) -> DbfvCiphertext:
    """bootstrap/bfv_host.rs:212-236: every limb refreshed (one batch), dBFV metadata kept, BFV params swapped
    to the boot set, mul_depth restarted."""
    if bsk.bootstrap is not None:
        out = bsk.bootstrap(ct)
        out.mul_depth = 0                                                    # :233
        return out
    if bsk.bsk is None or bsk.rounding_poly is None:
        raise NotImplementedErr("dbfv_bootstrap needs a BootstrapKey with key material (bsk, galois_keys, "
                                "rounding_poly) or a bootstrap= callable")
    refreshed = DbfvParams.new(bsk.boot_params, ct.params.base, ct.params.num_digits, ct.params.plain_modulus)   # :218-223
    for limb in ct.limbs:
        if len(limb.c) != 2:
            raise InvalidParam("bootstrap requires degree-1 ciphertext")
    out = _bootstrap_batch(ct.params.bfv_params, ct.to_array(), bsk)
    return DbfvCiphertext.from_array(out, refreshed, degree=ct.degree, mul_depth=0)


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
