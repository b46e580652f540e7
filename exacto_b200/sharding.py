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


def gather_limbs(out: torch.Tensor, masks: List[int], group=None) -> torch.Tensor:
    """All-gather of k-sharded results.  ``out`` is [B, d, 2, n]; rank r has filled the limbs in
    masks[r].  Every rank returns the complete tensor.  One collective per batch."""
    world = dist.get_world_size(group)
    if world == 1:
        return out
    parts = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(parts, out.contiguous(), group=group)
    full = out.clone()
    for r, m in enumerate(masks):
        for k in range(out.shape[1]):
            if (m >> k) & 1:
                full[:, k] = parts[r][:, k]
    return full
