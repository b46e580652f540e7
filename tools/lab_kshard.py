"""Lab: where a k-sharded step spends its time (run under torchrun)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import exacto_b200 as E
from exacto_b200 import batch, _native
from exacto_b200.sharding import KShard
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
params = E.u64_dbfv(); P = params.bfv_params; q = P.modulus(0)
rng = np.random.default_rng(1)
kp = 148
a = batch.to_device(rng.integers(0, q, (kp, 8, 2, 4096), dtype=np.uint64), dev)
b = batch.to_device(rng.integers(0, q, (kp, 8, 2, 4096), dtype=np.uint64), dev)
rlk = E.RelinKey(rng.integers(0, q, (8, 2, 4096), dtype=np.uint64), P)
ks = KShard(params, kp, dev)
out = torch.empty_like(a)
L = _native.lib(); ctx = ks.ctx
def timeit(fn, steps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
mask = ks.masks[rank]
t_full = timeit(lambda: batch.dbfv_mul(params, a, b, rlk, out=out))
t_mask = timeit(lambda: batch.dbfv_mul(params, a, b, rlk, out=out, limb_mask=mask))
flag = torch.zeros(1, device=dev)
t_ar = timeit(lambda: dist.all_reduce(flag))
t_mask_ar = timeit(lambda: (batch.dbfv_mul(params, a, b, rlk, out=out, limb_mask=mask), dist.all_reduce(flag)))
t_ks = timeit(lambda: ks.mul(a, b, rlk))
ctx.set_option("kshard_kernel_stores", 1)
t_ks2 = timeit(lambda: ks.mul(a, b, rlk))
ctx.set_option("kshard_kernel_stores", 0)
_native.check(L.exb_profile_enable(ctx.handle, 1))
for _ in range(10): batch.dbfv_mul(params, a, b, rlk, out=out, limb_mask=mask)
torch.cuda.synchronize()
ms = (ctypes.c_double * 5)(); n = (ctypes.c_ulonglong * 5)()
_native.check(L.exb_profile_read(ctx.handle, ms, n))
stages = [round(ms[i] / max(n[i], 1), 3) for i in range(4)]
for r in range(world):
    if r == rank:
        print(f"rank {rank} mask {mask:#x}: full {t_full:.3f} ms | own limbs {t_mask:.3f} | all_reduce alone {t_ar:.3f} | own+all_reduce {t_mask_ar:.3f} | kshard dma {t_ks:.3f} kernel-stores {t_ks2:.3f} | stages lift/t01/t32/relin {stages}", flush=True)
    dist.barrier()
ks.close(); dist.destroy_process_group()
