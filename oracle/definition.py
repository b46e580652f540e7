"""CPU ORACLE (test infrastructure) -- the "by definition" big-int oracle.

Independent of ``exacto_oracle.c``: no NTT, no RNS, no HPS.  It works in the
coefficient domain on Python integers and restates what the reference's hot
path *means*:

* the degree-2 tensor of the centered inputs over Z[X]/(X^n+1);
* ``round_nearest(p * t / q) mod q`` (bfv/eval.rs:689-703 ``scale_tensor_component``;
  equal to the HPS result while |m| < P/2 -- SURVEY.md finding 2);
* balanced gadget digits (bfv/keyswitch.rs:11-52) and the key-switch inner
  product (bfv/keyswitch.rs:83-95) with schoolbook negacyclic products;
* the dBFV convolution + degree reduction (dbfv/eval.rs:109-146,
  dbfv/reduction.rs:28-52).

Pure-Python loops: for small n only.  Citations are into /root/reference/src/.
"""
from __future__ import annotations

from typing import List, Sequence


def center(c: int, m: int) -> int:
    """Strict centering c > floor(m/2) => c - m (bfv/eval.rs:232,304)."""
    return c - m if c > m // 2 else c


def negacyclic_mul_int(a: Sequence[int], b: Sequence[int]) -> List[int]:
    n = len(a)
    r = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            k = i + j
            if k < n:
                r[k] += x * y
            else:
                r[k - n] -= x * y
    return r


def negacyclic_mul_mod(a, b, q):
    return [x % q for x in negacyclic_mul_int([int(v) for v in a], [int(v) for v in b])]


def round_div(num: int, q: int) -> int:
    """Sign-symmetric rounding helper (bfv/eval.rs:324-328)."""
    return (num + q // 2) // q if num >= 0 else -((-num + q // 2) // q)


def trunc_divmod(a: int, b: int):
    """Rust i128 `/` and `%` truncate toward zero."""
    qd = abs(a) // abs(b)
    if (a < 0) != (b < 0):
        qd = -qd
    return qd, a - qd * b


def gadget_decompose(c_mod_q: int, q: int, base: int, num_digits: int) -> List[int]:
    """Balanced digits of one coefficient, each mod q (bfv/keyswitch.rs:24-44)."""
    remaining = center(int(c_mod_q), q)
    half = base // 2
    out = []
    for _ in range(num_digits):
        _, rem = trunc_divmod(remaining, base)
        if rem < -half:
            rem += base
        elif rem >= half:
            rem -= base
        out.append(rem % q)
        remaining, _ = trunc_divmod(remaining - rem, base)
    return out


def bfv_mul_no_relin_coeff(ct1, ct2, q: int, p: int):
    """Coefficient-domain inputs [2][n] -> [3][n] coefficient-domain outputs."""
    c0, c1 = ([center(int(v), q) for v in row] for row in ct1)
    d0, d1 = ([center(int(v), q) for v in row] for row in ct2)
    t0 = negacyclic_mul_int(c0, d0)
    t1 = [x + y for x, y in zip(negacyclic_mul_int(c0, d1), negacyclic_mul_int(c1, d0))]
    t2 = negacyclic_mul_int(c1, d1)
    return [[round_div(p * x, q) % q for x in t] for t in (t0, t1, t2)]


def relinearize_coeff(c3, rlk, q: int, base: int, num_digits: int):
    """c3 [3][n], rlk [G][2][n], all coefficient domain -> [2][n]."""
    n = len(c3[0])
    digits = [[0] * n for _ in range(num_digits)]
    for pos in range(n):
        for g, dig in enumerate(gadget_decompose(c3[2][pos], q, base, num_digits)):
            digits[g][pos] = dig
    out0, out1 = [int(v) for v in c3[0]], [int(v) for v in c3[1]]
    for g in range(num_digits):
        # centered digits keep the integers small; the result mod q is the same
        dg = [center(v, q) for v in digits[g]]
        p0 = negacyclic_mul_int(dg, [int(v) for v in rlk[g][0]])
        p1 = negacyclic_mul_int(dg, [int(v) for v in rlk[g][1]])
        out0 = [(x + y) % q for x, y in zip(out0, p0)]
        out1 = [(x + y) % q for x, y in zip(out1, p1)]
    return [out0, out1]


def bfv_mul_and_relin_coeff(ct1, ct2, rlk, q: int, p: int, base: int, num_digits: int):
    return relinearize_coeff(bfv_mul_no_relin_coeff(ct1, ct2, q, p), rlk, q, base, num_digits)


def small_reps(base: int, d: int, plain_modulus: int):
    """SmallReps::compute_simple (dbfv/lattice.rs:104-122)."""
    reps = []
    for j in range(d, 2 * d - 1):
        val = pow(base, j, 1 << 64) if plain_modulus == 0 else pow(base, j, plain_modulus)
        digs = []
        for _ in range(d):
            digs.append(val % base)
            val //= base
        reps.append(digs)
    return reps


def dbfv_mul_coeff(ct1, ct2, rlk, q: int, p: int, gbase: int, gdigits: int, base: int, d: int,
                   dbfv_plain_modulus: int):
    """ct [d][2][n] coefficient domain -> [d][2][n] coefficient domain."""
    n = len(ct1[0][0])
    limbs = [None] * (2 * d - 1)
    for i in range(d):
        for j in range(d):
            prod = bfv_mul_and_relin_coeff(ct1[i], ct2[j], rlk, q, p, gbase, gdigits)
            k = i + j
            if limbs[k] is None:
                limbs[k] = prod
            else:
                limbs[k] = [[(x + y) % q for x, y in zip(a, b)] for a, b in zip(limbs[k], prod)]
    out = [[list(c) for c in limbs[i]] for i in range(d)]
    reps = small_reps(base, d, dbfv_plain_modulus)
    for j in range(d, 2 * d - 1):
        rep = reps[j - d]
        for i in range(d):
            s = rep[i]
            if s == 0:
                continue
            for c in range(2):
                out[i][c] = [(x + (s % q) * y) % q for x, y in zip(out[i][c], limbs[j][c])]
    assert all(len(c) == n for l in out for c in l)
    return out
