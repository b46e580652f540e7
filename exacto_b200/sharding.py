"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for plumbing).

Two levels (SURVEY.md section 8 e):
  * batch-parallel: independent ciphertext pairs are split across ranks, the relinearisation
    key is replicated, and there is NO data-path collective (``pair_range``);
  * within one dbfv_mul: ranks own disjoint *output limbs k* (products with equal i+j stay on
    one rank so the per-k accumulation is local), then one all-gather of the d output limbs
    (``limb_masks`` + ``gather_limbs``).  This is the only place the path has a real exchange.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def pair_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `batch` pairs owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def limb_masks(d: int, world: int) -> List[int]:
    """Assign output limbs k in [0, d) to ranks, balancing the number of live products
    (limb k costs k+1 products when p = b^d).  Returns one bitmask per rank; ranks beyond d
    get 0 (= nothing to do)."""
    load = [0] * world
    masks = [0] * world
    for k in sorted(range(d), key=lambda k: -(k + 1)):          # longest-processing-time first
        r = min(range(world), key=lambda r: (load[r], r))
        masks[r] |= 1 << k
        load[r] += k + 1
    return masks


def limb_owner(masks: List[int], d: int) -> List[int]:
    """owner[k] = the rank whose mask holds output limb k."""
    return [next(r for r, m in enumerate(masks) if (m >> k) & 1) for k in range(d)]


def gather_limbs(out: torch.Tensor, masks: List[int], group=None) -> torch.Tensor:
    """Collective gather of k-sharded results (the NCCL / gloo transport; the NVLink peer-store transport is
    ``KShard``).  ``out`` is [B, d, 2, n]; rank r has filled the limbs in masks[r]; on return every rank holds
    the complete tensor IN PLACE.  Only owned limbs travel: limb k is broadcast by its owner as one contiguous
    [B, 2, n] slab, so every rank receives exactly the limbs it does not own -- (N-1)/N of a ciphertext when the
    limbs divide evenly -- instead of N whole tensors."""
    world = dist.get_world_size(group)
    if world == 1:
        return out
    rank = dist.get_rank(group)
    d = out.shape[1]
    owner = limb_owner(masks, d)
    works, slabs = [], []
    for k in range(d):
        slab = out[:, k].contiguous()                       # [B, 2, n]
        src = owner[k] if group is None else dist.get_global_rank(group, owner[k])
        works.append(dist.broadcast(slab, src=src, group=group, async_op=True))
        slabs.append(slab)
    for k, (w, slab) in enumerate(zip(works, slabs)):
        w.wait()
        if owner[k] != rank:
            out[:, k].copy_(slab)
    return out


def gather_wire_bytes_per_rank(masks: List[int], d: int, rank: int, limb_bytes: int) -> int:
    """Bytes ``rank`` receives per ciphertext in gather_limbs / KShard: the limbs it does not own."""
    return sum(limb_bytes for k in range(d) if not (masks[rank] >> k) & 1)


class _DevBuf:
    """A cudaMalloc allocation (exb_device_alloc) exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, owner):
        self.ptr, self.shape, self._owner = ptr, tuple(shape), owner
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<i8", "data": (ptr, False), "version": 2}


class KShard:
    """k-sharded dbfv_mul over the GPUs of one box (one process per GPU, ``torch.distributed`` initialised).

    Every rank holds the same pairs and owns the output limbs of ``masks[rank]``.  Each rank's output buffer is a
    cudaMalloc allocation shared with the other ranks through CUDA IPC; ``exb_dbfv_mul_scatter`` copies every
    finished limb into all N output buffers over NVLink peer memory (copy engines by default; the relinearisation
    kernel's epilogue can do the stores itself, measured slower).  One tiny all-reduce per call orders the ranks; two output buffers alternate so
    a rank may still read call i's result while call i+1 is being written.  ``mul`` returns the complete
    [pairs, d, 2, n] tensor (a view of the current buffer, valid until the call after next)."""

    def __init__(self, params, pairs: int, device, group=None):
        import ctypes
        from . import _native
        self._ctypes, self._native, self._L = ctypes, _native, _native.lib()
        self.params, self.group, self.device = params, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n, d = params.bfv_params.ring_degree, params.num_digits
        self.shape = (pairs, d, 2, n)
        self.masks = limb_masks(d, self.world)
        self.ctx = params.bfv_params.context(device.index)
        self._flag = torch.zeros(1, device=device)
        nbytes = pairs * d * 2 * n * 8
        self._own, self._peers, self._views, self._peer_arr = [], [], [], []
        for _ in range(2):
            p = ctypes.c_void_p()
            _native.check(self._L.exb_device_alloc(self.ctx.handle, nbytes, ctypes.byref(p)))
            handle = ctypes.create_string_buffer(64)
            _native.check(self._L.exb_ipc_export(self.ctx.handle, p, handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            peers = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                q = ctypes.c_void_p()
                _native.check(self._L.exb_ipc_open(self.ctx.handle, h, ctypes.byref(q)))
                peers.append(q)
            self._own.append(p)
            self._peers.append(peers)
            self._peer_arr.append((ctypes.c_void_p * max(len(peers), 1))(*[q.value for q in peers]))
            self._views.append(torch.as_tensor(_DevBuf(p.value, self.shape, self), device=device))
        self._i = 0
        self.transport = ("peer DMA over NVLink (one strided cudaMemcpy2DAsync per peer and run of owned limbs, CUDA IPC "
                          "mappings, per-peer streams joined to the compute stream) + 1 all-reduce barrier")
        dist.barrier(group=group)

    def wire_bytes_per_pair_per_rank(self) -> int:
        n, d = self.shape[3], self.shape[1]
        return gather_wire_bytes_per_rank(self.masks, d, self.rank, 2 * n * 8)

    def mul(self, ct1: torch.Tensor, ct2: torch.Tensor, rlk, *, all_products: bool = False) -> torch.Tensor:
        assert tuple(ct1.shape) == self.shape and tuple(ct2.shape) == self.shape
        b = self._i & 1
        self._i += 1
        params, d = self.params, self.shape[1]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.masks[self.rank]:
            self._native.check(self._L.exb_dbfv_mul_scatter(
                self.ctx.handle, params.base, d, params.plain_modulus, ct1.data_ptr(), ct2.data_ptr(),
                rlk.native(self.ctx), self._own[b], self._peer_arr[b], len(self._peers[b]), self.shape[0],
                self._native.EXB_DBFV_ALL_PRODUCTS if all_products else 0, self.masks[self.rank], stream))
        dist.all_reduce(self._flag, group=self.group)        # every rank's stores have landed before anyone reads
        return self._views[b]

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._views = []
        for peers in self._peers:
            for q in peers:
                self._L.exb_ipc_close(self.ctx.handle, q)
        dist.barrier(group=self.group)
        for p in self._own:
            self._L.exb_device_free(self.ctx.handle, p)
        self._own, self._peers = [], []
