"""Page-locked host memory and asynchronous host-buffer calls (exb_host_alloc / exb_*_host_async / exb_wait).

The reference owns plain ``Vec<u64>`` (bfv/mod.rs:19-24) and its calls are CPU-synchronous; on the GPU the
host-buffer entry points run at PCIe speed only from page-locked memory, and a stream of calls only reaches
max(PCIe, kernels) when the next call's upload overlaps this call's kernels and download.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _native
from .bfv import RelinKey
from .error import InvalidParam
from .params import DbfvParams


class PinnedArray:
    """A page-locked uint64 array (exb_host_alloc); ``.array`` is a numpy view.  Freed on close()/GC."""

    def __init__(self, ctx, shape: Tuple[int, ...], write_combined: bool = False):
        self._ctx, self._L = ctx, _native.lib()
        nbytes = int(np.prod(shape)) * 8
        p = ctypes.c_void_p()
        flags = _native.EXB_HOST_WRITE_COMBINED if write_combined else 0     # inputs the CPU only writes
        _native.check(self._L.exb_host_alloc_ex(ctx.handle, nbytes, flags, ctypes.byref(p)))
        self._ptr = p
        buf = (ctypes.c_uint64 * (max(nbytes, 8) // 8)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.uint64, count=int(np.prod(shape))).reshape(shape)

    def close(self) -> None:
        if self._ptr is not None:
            self.array = None
            self._L.exb_host_free(self._ctx.handle, self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(params: DbfvParams, shape: Tuple[int, ...], device: Optional[int] = None) -> PinnedArray:
    return PinnedArray(params.bfv_params.context(device), shape)


class PendingMul:
    """An asynchronous dbfv_mul in flight; ``wait()`` returns the output array once it has landed."""

    def __init__(self, ctx, ticket: int, out: np.ndarray, keep):
        self._ctx, self._ticket, self._out, self._keep = ctx, ticket, out, keep

    def wait(self) -> np.ndarray:
        if self._ticket is not None:
            _native.check(_native.lib().exb_wait(self._ctx.handle, self._ticket))
            self._ticket, self._keep = None, None
        return self._out


def dbfv_mul_batch_async(params: DbfvParams, ct1: np.ndarray, ct2: np.ndarray, rlk: RelinKey, out: np.ndarray, *,
                         all_products: bool = False, device: Optional[int] = None) -> PendingMul:
    """exb_dbfv_mul_host_async: ct [B][d][2][n] -> ``out`` (same shape, caller-owned, ideally page-locked).
    Inputs and ``out`` must stay alive and untouched until ``wait()``."""
    n, d = params.bfv_params.ring_degree, params.num_digits
    for a in (ct1, ct2, out):
        if a.dtype != np.uint64 or not a.flags.c_contiguous:
            raise InvalidParam("asynchronous calls need contiguous uint64 arrays (no implicit copies)")
    if ct1.shape != ct2.shape or ct1.shape != out.shape or ct1.shape[1:] != (d, 2, n):
        raise InvalidParam("multiplication requires d-limb ciphertexts")
    ctx = params.bfv_params.context(device)
    flags = _native.EXB_DBFV_ALL_PRODUCTS if all_products else 0
    ticket = ctypes.c_uint64()
    _native.check(_native.lib().exb_dbfv_mul_host_async(ctx.handle, params.base, d, params.plain_modulus,
                                                        ct1.ctypes.data, ct2.ctypes.data, rlk.native(ctx),
                                                        out.ctypes.data, ct1.shape[0], flags, ctypes.byref(ticket)))
    return PendingMul(ctx, ticket.value, out, (ct1, ct2, rlk))
