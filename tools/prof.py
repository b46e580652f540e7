"""Small, fixed workloads for ncu captures (never used for bench numbers).

    python tools/prof.py ntt  [count]     # batched forward+inverse NTT, n=4096, prime q
    python tools/prof.py mul  [pairs]     # u64-profile dbfv_mul, device resident
    python tools/prof.py all  [pairs]     # dbfv_mul twice, then automorphism+key switch and decrypt twice each
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import exacto_b200 as E
from exacto_b200 import batch

what = sys.argv[1] if len(sys.argv) > 1 else "ntt"
params = E.u64_dbfv()
P = params.bfv_params
q = P.modulus(0)
rng = np.random.default_rng(1)
if what == "ntt":
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    x = batch.to_device(rng.integers(0, q, (count, 4096), dtype=np.uint64))
    y = torch.empty_like(x)
    for _ in range(3):
        batch.ntt_forward(P, 0, x, out=y)
        batch.ntt_inverse(P, 0, y, out=x)
else:
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    a = batch.to_device(rng.integers(0, q, (pairs, 8, 2, 4096), dtype=np.uint64))
    b = batch.to_device(rng.integers(0, q, (pairs, 8, 2, 4096), dtype=np.uint64))
    rlk = E.RelinKey(rng.integers(0, q, (8, 2, 4096), dtype=np.uint64), P)
    for _ in range(2):
        batch.dbfv_mul(params, a, b, rlk)
    if what == "all":
        ct = batch.to_device(rng.integers(0, q, (2048, 2, 4096), dtype=np.uint64))
        gk = E.GaloisKey(rng.integers(0, q, (8, 2, 4096), dtype=np.uint64), 3, P)
        sk = batch.to_device(rng.integers(0, q, 4096, dtype=np.uint64))
        for _ in range(2):
            batch.bfv_apply_automorphism(P, ct, gk)
            batch.bfv_decrypt(P, ct, sk)
torch.cuda.synchronize()
print("done", what)
