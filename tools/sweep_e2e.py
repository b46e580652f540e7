"""Lab: e2e (host-buffer, async pipelined) throughput vs pipeline depth and chunk size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import exacto_b200 as E
from exacto_b200 import hostmem
import bench
params = E.u64_dbfv(); P = params.bfv_params; q = P.modulus(0)
ctx = P.context(0)
pairs = 1480
shape = (pairs, 8, 2, 4096)
WC = bool(int(os.environ.get("WC", "0")))
h1, h2 = hostmem.PinnedArray(ctx, shape, WC), hostmem.PinnedArray(ctx, shape, WC)
rng = np.random.default_rng(1)
tmp = np.empty(shape, np.uint64)
bench.fill_uniform(rng, tmp, q); h1.array[:] = tmp
bench.fill_uniform(rng, tmp, q); h2.array[:] = tmp
outs = [hostmem.PinnedArray(ctx, shape), hostmem.PinnedArray(ctx, shape)]
rlk = E.RelinKey(rng.integers(0, q, (8, 2, 4096), dtype=np.uint64), P)
def run(steps):
    pend = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(steps):
        nxt = hostmem.dbfv_mul_batch_async(params, h1.array, h2.array, rlk, outs[s & 1].array)
        if pend is not None: pend.wait()
        pend = nxt
    pend.wait()
    return pairs * steps / (time.perf_counter() - t0)
for slots in (4,):
    for chunk in (3996,):
        ctx.set_option("host_slots", slots); ctx.set_option("host_chunk_products", chunk)
        run(2)
        print(f"slots={slots} chunk_products={chunk}: {run(8):.0f} dbfv_mul/s", flush=True)
