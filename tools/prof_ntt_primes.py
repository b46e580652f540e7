import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import exacto_b200 as E
from exacto_b200 import batch
P = E.u64_dbfv().bfv_params
rng = np.random.default_rng(1)
for idx in (1, 0):
    q = P.modulus(idx)
    x = batch.to_device(rng.integers(0, q, (8192, 4096), dtype=np.uint64)); y = torch.empty_like(x)
    for _ in range(3):
        batch.ntt_forward(P, idx, x, out=y); batch.ntt_inverse(P, idx, y, out=x)
torch.cuda.synchronize()
