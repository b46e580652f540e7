import sys, numpy as np, torch, statistics
sys.path.insert(0, "/root/repo")
import exacto_b200 as E
from exacto_b200 import batch
rng = np.random.default_rng(1)
dp = E.u64_dbfv(); P = dp.bfv_params; q = P.modulus(0)
rlk = E.RelinKey(rng.integers(0, q, (8, 2, 4096), dtype=np.uint64), P)
res = []
for B in (16, 20, 24, 32, 40, 48, 64, 96):
    a = batch.to_device(rng.integers(0, q, (B, 8, 2, 4096), dtype=np.uint64)); b = batch.to_device(rng.integers(0, q, (B, 8, 2, 4096), dtype=np.uint64))
    o = torch.empty_like(a)
    for _ in range(3): batch.dbfv_mul(dp, a, b, rlk, out=o)
    ts = []
    for _ in range(11):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.dbfv_mul(dp, a, b, rlk, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res.append((B, round(statistics.median(ts), 3)))
print(res)
