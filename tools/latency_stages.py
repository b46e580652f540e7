import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
import exacto_b200 as E
from exacto_b200 import batch, _native
dp = E.u64_dbfv(); P = dp.bfv_params; q = P.modulus(0)
rng = np.random.default_rng(1)
rlk = E.RelinKey(rng.integers(0, q, (8, 2, 4096), dtype=np.uint64), P)
ctx = P.context(0); L = _native.lib()
for B in (1, 2, 4, 8, 16):
    a = batch.to_device(rng.integers(0, q, (B, 8, 2, 4096), dtype=np.uint64)); b = batch.to_device(rng.integers(0, q, (B, 8, 2, 4096), dtype=np.uint64))
    for _ in range(3): batch.dbfv_mul(dp, a, b, rlk)
    torch.cuda.synchronize()
    L.exb_profile_enable(ctx.handle, 1)
    for _ in range(20): batch.dbfv_mul(dp, a, b, rlk)
    ms = (ctypes.c_double * 5)(); n = (ctypes.c_ulonglong * 5)()
    L.exb_profile_read(ctx.handle, ms, n); L.exb_profile_enable(ctx.handle, 0)
    print(B, [round(ms[i] / max(n[i], 1) * 1000, 1) for i in range(4)], 'us: lift, tensor(01), tensor c2, relin')
