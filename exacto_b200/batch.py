"""Device-resident batched entry points (throughput path).

Ciphertext batches live in HBM as ``torch.int64`` CUDA tensors holding the u64 residues
bit-for-bit (torch is plumbing: memory + streams); the work is done by the C ABI
(``exb_dbfv_mul``, ``exb_bfv_mul_and_relin``, ``exb_ntt_*``) on the current torch stream.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _native
from .bfv import RelinKey
from .error import InvalidParam
from .params import BfvParams, DbfvParams


def to_device(arr: np.ndarray, device=None) -> torch.Tensor:
    """numpy uint64 -> CUDA int64 tensor with the same bits."""
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint64).view(np.int64))
    return t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)


def _check(t: torch.Tensor, shape_tail, what: str):
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise InvalidParam(f"{what}: need a contiguous CUDA int64 tensor")
    if tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise InvalidParam(f"{what}: trailing shape {tuple(t.shape)} != (..., {shape_tail})")


def _ct_tail(params: BfvParams):
    """Trailing shape of one ciphertext component: (n,) for a single ciphertext prime, (L, n) otherwise."""
    L = params.ct_basis.num_moduli()
    return (params.ring_degree,) if L == 1 else (L, params.ring_degree)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def ntt_forward(params: BfvParams, index: int, polys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NttPoly::from_coeff_poly (ring/ntt.rs:42-55) over [..., n]."""
    n = params.ring_degree
    _check(polys, (n,), "ntt_forward")
    out = torch.empty_like(polys) if out is None else out
    ctx = params.context(polys.device.index)
    _native.check(_native.lib().exb_ntt_forward(ctx.handle, index, polys.data_ptr(), out.data_ptr(),
                                                polys.numel() // n, _stream(polys)))
    return out


def ntt_inverse(params: BfvParams, index: int, polys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NttPoly::to_coeff_poly (ring/ntt.rs:58-67) over [..., n]."""
    n = params.ring_degree
    _check(polys, (n,), "ntt_inverse")
    out = torch.empty_like(polys) if out is None else out
    ctx = params.context(polys.device.index)
    _native.check(_native.lib().exb_ntt_inverse(ctx.handle, index, polys.data_ptr(), out.data_ptr(),
                                                polys.numel() // n, _stream(polys)))
    return out


def bfv_mul_and_relin(params: BfvParams, ct1: torch.Tensor, ct2: torch.Tensor, rlk: RelinKey,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bfv/eval.rs:73-82 over [B, 2, n]."""
    tail = (2,) + _ct_tail(params)
    _check(ct1, tail, "ct1"); _check(ct2, tail, "ct2")
    if ct1.shape != ct2.shape:
        raise InvalidParam("ct1/ct2 shape mismatch")
    out = torch.empty_like(ct1) if out is None else out
    ctx = params.context(ct1.device.index)
    _native.check(_native.lib().exb_bfv_mul_and_relin(ctx.handle, ct1.data_ptr(), ct2.data_ptr(), rlk.native(ctx),
                                                      out.data_ptr(), ct1.numel() // int(np.prod(tail)), _stream(ct1)))
    return out


def dbfv_mul(params: DbfvParams, ct1: torch.Tensor, ct2: torch.Tensor, rlk: RelinKey,
             out: Optional[torch.Tensor] = None, *, all_products: bool = False, limb_mask: int = 0) -> torch.Tensor:
    """dbfv/eval.rs:82-149 over [B, d, 2, n].  ``limb_mask`` selects output limbs (multi-GPU k-sharding)."""
    d = params.num_digits
    tail = (d, 2) + _ct_tail(params.bfv_params)
    _check(ct1, tail, "ct1"); _check(ct2, tail, "ct2")
    if ct1.shape != ct2.shape:
        raise InvalidParam("ct1/ct2 shape mismatch")
    out = torch.empty_like(ct1) if out is None else out
    ctx = params.bfv_params.context(ct1.device.index)
    flags = _native.EXB_DBFV_ALL_PRODUCTS if all_products else 0
    _native.check(_native.lib().exb_dbfv_mul(ctx.handle, params.base, d, params.plain_modulus, ct1.data_ptr(),
                                             ct2.data_ptr(), rlk.native(ctx), out.data_ptr(),
                                             ct1.numel() // int(np.prod(tail)), flags, limb_mask, _stream(ct1)))
    return out


def bfv_mul_no_relin(params: BfvParams, ct1: torch.Tensor, ct2: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bfv/eval.rs:89-108 over [B, 2, n] -> [B, 3, n]."""
    tail = _ct_tail(params)
    _check(ct1, (2,) + tail, "ct1"); _check(ct2, (2,) + tail, "ct2")
    if ct1.shape != ct2.shape or ct1.dim() != 2 + len(tail):
        raise InvalidParam("ct1/ct2: need matching [batch, 2, n]")
    out = torch.empty((ct1.shape[0], 3) + tail, dtype=torch.int64, device=ct1.device) if out is None else out
    ctx = params.context(ct1.device.index)
    _native.check(_native.lib().exb_bfv_mul_no_relin(ctx.handle, ct1.data_ptr(), ct2.data_ptr(), out.data_ptr(), ct1.shape[0],
                                                     _stream(ct1)))
    return out


def relinearize(params: BfvParams, ct3: torch.Tensor, rlk: RelinKey, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bfv/keyswitch.rs:59-101 over [B, 3, n] -> [B, 2, n]."""
    tail = _ct_tail(params)
    _check(ct3, (3,) + tail, "ct3")
    if ct3.dim() != 2 + len(tail):
        raise InvalidParam("ct3: need [batch, 3, n]")
    out = torch.empty((ct3.shape[0], 2) + tail, dtype=torch.int64, device=ct3.device) if out is None else out
    ctx = params.context(ct3.device.index)
    _native.check(_native.lib().exb_bfv_relinearize(ctx.handle, ct3.data_ptr(), 3, rlk.native(ctx), out.data_ptr(), ct3.shape[0],
                                                    _stream(ct3)))
    return out


def bfv_apply_automorphism(params: BfvParams, ct: torch.Tensor, gk, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bfv/eval.rs:512-561 over [B, 2, n] (or [B, d, 2, n]: every limb, dbfv/advanced.rs:15-30); ``gk`` is a
    GaloisKey.  ``out`` must not alias ``ct``."""
    n = params.ring_degree
    _check(ct, (2, n), "ct")
    out = torch.empty_like(ct) if out is None else out
    ctx = params.context(ct.device.index)
    _native.check(_native.lib().exb_bfv_apply_automorphism(ctx.handle, ct.data_ptr(), gk.element, gk.native(ctx),
                                                           out.data_ptr(), ct.numel() // (2 * n), _stream(ct)))
    return out


def bfv_decrypt(params: BfvParams, ct: torch.Tensor, sk_ntt: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bfv/encrypt.rs:111-178 over [B, k, n] ciphertexts with the secret key [n] in the NTT domain -> [B, n]
    plaintext coefficients mod p."""
    n = params.ring_degree
    _check(ct, (n,), "ct"); _check(sk_ntt, (n,), "sk_ntt")
    if ct.dim() != 3:
        raise InvalidParam("ct: need [batch, components, n]")
    out = torch.empty((ct.shape[0], n), dtype=torch.int64, device=ct.device) if out is None else out
    ctx = params.context(ct.device.index)
    _native.check(_native.lib().exb_bfv_decrypt(ctx.handle, ct.data_ptr(), ct.shape[1], sk_ntt.data_ptr(), out.data_ptr(),
                                                ct.shape[0], _stream(ct)))
    return out


def _poly_binary(name: str, params: BfvParams, index: int, a: torch.Tensor, b: Optional[torch.Tensor],
                 out: Optional[torch.Tensor], scalar: int = 0) -> torch.Tensor:
    _check(a, (params.ring_degree,), name)
    out = torch.empty_like(a) if out is None else out
    ctx = params.context(a.device.index)
    L = _native.lib()
    if name == "exb_poly_neg":
        rc = L.exb_poly_neg(ctx.handle, index, a.data_ptr(), out.data_ptr(), a.numel(), _stream(a))
    elif name == "exb_poly_scalar_mul":
        rc = L.exb_poly_scalar_mul(ctx.handle, index, a.data_ptr(), scalar % (1 << 64), out.data_ptr(), a.numel(), _stream(a))
    else:
        _check(b, (params.ring_degree,), name)
        if a.shape != b.shape:
            raise InvalidParam("operand shape mismatch")
        rc = getattr(L, name)(ctx.handle, index, a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream(a))
    _native.check(rc)
    return out


def poly_add(params, index, a, b, out=None):
    """NttPoly::add (ring/ntt.rs:75-89) over [..., n]."""
    return _poly_binary("exb_poly_add", params, index, a, b, out)


def poly_sub(params, index, a, b, out=None):
    """NttPoly::sub (ring/ntt.rs:92-105)."""
    return _poly_binary("exb_poly_sub", params, index, a, b, out)


def poly_neg(params, index, a, out=None):
    """NttPoly::neg (ring/ntt.rs:108-113)."""
    return _poly_binary("exb_poly_neg", params, index, a, None, out)


def poly_mul(params, index, a, b, out=None):
    """NttPoly::mul (ring/ntt.rs:119-129)."""
    return _poly_binary("exb_poly_mul", params, index, a, b, out)


def poly_scalar_mul(params, index, a, scalar: int, out=None):
    """NttPoly::scalar_mul (ring/ntt.rs:132-139)."""
    return _poly_binary("exb_poly_scalar_mul", params, index, a, None, out, scalar)


def launch_count() -> int:
    return int(_native.lib().exb_launch_count())
