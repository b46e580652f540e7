"""Smallest end-to-end case for compute-sanitizer (tools): one u64-profile BFV mul + one compact dBFV mul + NTTs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import exacto_b200 as E
from exacto_b200 import batch
rng = np.random.default_rng(3)
for params in (E.u64_dbfv(), E.compact_dbfv()):
    P = params.bfv_params
    q, n, d = P.modulus(0), P.ring_degree, params.num_digits
    a = batch.to_device(rng.integers(0, q, (1, d, 2, n), dtype=np.uint64))
    b = batch.to_device(rng.integers(0, q, (1, d, 2, n), dtype=np.uint64))
    rlk = E.RelinKey(rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64), P)
    out = batch.dbfv_mul(params, a[:, :2].contiguous() if False else a, b, rlk)
    x = batch.ntt_forward(P, 0, a.reshape(-1, n))
    y = batch.ntt_inverse(P, 0, x)
    torch.cuda.synchronize()
    assert torch.equal(y, a.reshape(-1, n))
print("sanitize case ok")
