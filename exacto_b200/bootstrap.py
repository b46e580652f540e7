"""bootstrap/: the two multiplication entry points north_star names, with the reference's
signatures and rlk-selection logic (bootstrap/bfv_host.rs:242-288).

The GPU content of these functions is ``dbfv_mul`` (this package).  The bootstrap *body*
(``dbfv_bootstrap`` -> ``bfv_bootstrap``: modulus switch, coefficient extraction, digit
extraction polynomial, bootstrap/bfv_host.rs:134-236) is outside the hot-path scope of this
round: callers supply it as ``bsk.bootstrap(ct)``; without one the
functions fail loudly after the multiplication instead of silently skipping the refresh.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from .bfv import RelinKey
from .dbfv import DbfvCiphertext, dbfv_mul
from .error import InvalidParam, NotImplementedErr
from .params import BfvParams


class BootstrapKey:
    """bootstrap/bfv_host.rs:19-41 reduced to what the chain logic reads: the boot parameter set,
    its relinearisation key, and a refresh callable standing in for dbfv_bootstrap."""

    def __init__(self, boot_params: BfvParams, boot_rlk: RelinKey,
                 bootstrap: Optional[Callable[[DbfvCiphertext], DbfvCiphertext]] = None):
        self.boot_params = boot_params
        self.boot_rlk = boot_rlk
        self.bootstrap = bootstrap


def dbfv_bootstrap(ct: DbfvCiphertext, bsk: BootstrapKey) -> DbfvCiphertext:
    """bootstrap/bfv_host.rs:212-236 (body supplied by the caller, see module docstring)."""
    if bsk.bootstrap is None:
        raise NotImplementedErr("dbfv_bootstrap body is outside the device library's hot-path scope; "
                                "pass BootstrapKey(bootstrap=...)")
    out = bsk.bootstrap(ct)
    out.mul_depth = 0                                                    # :233
    return out


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
