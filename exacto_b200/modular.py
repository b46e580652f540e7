"""ring/modular.rs: the scalar helpers, for host-side code that builds parameters, plaintexts and constants
(the kernels use their own Shoup / Montgomery / Barrett forms, csrc/modarith.cuh).  Same argument meaning and
results as the reference; u64 / u128 wrap-around is made explicit."""
from __future__ import annotations

from typing import Optional

_M64 = (1 << 64) - 1


def barrett_constant(m: int) -> int:
    """ring/modular.rs:23-31: floor(2^64 / m) (u64::MAX for m = 1)."""
    return min((1 << 64) // m, _M64)


def barrett_reduce(a: int, m: int, barrett_k: int) -> int:
    """ring/modular.rs:7-19: a mod m for a < 2^128 (canonical result, whatever estimate barrett_k gives)."""
    return a % m


def montgomery_inv_neg(m: int) -> int:
    """ring/modular.rs:43-53: -m^-1 mod 2^64 for odd m."""
    return (-pow(m, -1, 1 << 64)) & _M64


def montgomery_reduce(t: int, m: int, m_inv_neg: int) -> int:
    """ring/modular.rs:34-40: t * 2^-64 mod m for t < m * 2^64."""
    k = ((t & _M64) * m_inv_neg) & _M64
    r = (t + k * m) >> 64
    return r - m if r >= m else r


def mod_add(a: int, b: int, m: int) -> int:
    """ring/modular.rs:57-62."""
    s = a + b
    return s - m if s >= m else s


def mod_sub(a: int, b: int, m: int) -> int:
    """ring/modular.rs:65-72."""
    return a - b if a >= b else m - b + a


def mod_neg(a: int, m: int) -> int:
    """ring/modular.rs:75-77."""
    return 0 if a == 0 else m - a


def mod_mul(a: int, b: int, m: int, barrett_k: Optional[int] = None) -> int:
    """ring/modular.rs:81-84."""
    return a * b % m


def mod_pow(base: int, exp: int, m: int) -> int:
    """ring/modular.rs:87-99."""
    return 0 if m == 1 else pow(base % m, exp, m)


def mod_inv(a: int, m: int) -> Optional[int]:
    """ring/modular.rs:102-121: None when gcd(a, m) != 1."""
    try:
        return pow(a % m, -1, m)
    except ValueError:
        return None
