"""dbfv/advanced.rs:36-168: division by the base and change of base on dBFV ciphertexts, as compositions of the
BFV point-wise operations (all on the GPU).  dbfv_apply_automorphism lives in dbfv.py."""
from __future__ import annotations

from .bfv import bfv_add, bfv_scalar_mul, bfv_sub
from .dbfv import DbfvCiphertext
from .encrypt import digit_decompose
from .error import InvalidParam
from .params import DbfvParams


def dbfv_div_by_base(ct: DbfvCiphertext) -> DbfvCiphertext:
    """dbfv/advanced.rs:36-92: phi_b(c) = c0 / b + c~(B); plaintext modulus divided by the base."""
    d = ct.params.num_digits
    if d == 0 or len(ct.limbs) == 0:
        raise InvalidParam("empty dBFV ciphertext")
    base, t = ct.params.base, ct.params.bfv_params.plain_modulus
    try:
        base_inv_t = pow(base % t, -1, t)
    except ValueError:
        raise InvalidParam("base not invertible modulo BFV plaintext modulus")
    old_p = (1 << 64) if ct.params.plain_modulus == 0 else ct.params.plain_modulus
    if old_p % base != 0:
        raise InvalidParam(f"plaintext modulus {old_p} is not divisible by base {base}")
    new_p = old_p // base
    new_p_u64 = 0 if new_p == (1 << 64) else new_p
    c0_div = bfv_scalar_mul(ct.limbs[0], base_inv_t)
    zero = bfv_sub(ct.limbs[d - 1], ct.limbs[d - 1])
    limbs = [zero] * d
    limbs[0] = bfv_add(ct.limbs[1], c0_div) if d >= 2 else c0_div
    for i in range(1, d):
        limbs[i] = ct.limbs[i + 1] if i + 1 < d else zero
    new_params = DbfvParams.new(ct.params.bfv_params, ct.params.base, ct.params.num_digits, new_p_u64)
    return DbfvCiphertext(limbs, max(max(ct.degree - 1, 0), 1), ct.mul_depth, new_params)


def dbfv_change_base(ct: DbfvCiphertext, new_base: int, new_num_digits: int) -> DbfvCiphertext:
    """dbfv/advanced.rs:99-158: column i of the linear map is the base-b' digit vector of b^i mod p."""
    if new_base < 2:
        raise InvalidParam("new base must be >= 2")
    if new_num_digits == 0:
        raise InvalidParam("new_num_digits must be >= 1")
    old_base, old_d = ct.params.base, ct.params.num_digits
    p = (1 << 64) if ct.params.plain_modulus == 0 else ct.params.plain_modulus
    transform = [[0] * old_d for _ in range(new_num_digits)]
    b_pow = 1
    for i in range(old_d):
        digits = digit_decompose((b_pow % p) % (1 << 64), new_base, new_num_digits)
        for j in range(new_num_digits):
            transform[j][i] = digits[j]
        b_pow = (b_pow * old_base) % p
    zero = bfv_sub(ct.limbs[0], ct.limbs[0])
    new_limbs = []
    for j in range(new_num_digits):
        acc = zero
        for i in range(old_d):
            if transform[j][i] == 0:
                continue
            acc = bfv_add(acc, bfv_scalar_mul(ct.limbs[i], transform[j][i]))
        new_limbs.append(acc)
    new_params = DbfvParams.new(ct.params.bfv_params, new_base, new_num_digits, ct.params.plain_modulus)
    return DbfvCiphertext(new_limbs, new_num_digits, ct.mul_depth, new_params)
