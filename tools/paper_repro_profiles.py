"""The three profiles of the reference's src/bin/paper_repro.rs (n = 4096, p = 2^64 scalar dBFV), device resident:
one dbfv_mul alone (what the reference times, 20-trial mean) and a batch of 64.  Prints one JSON object."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import exacto_b200 as E
from exacto_b200 import batch

PUBLISHED_MS = {"d=4, b=2^16": 8.881, "d=8, b=2^8": 31.395, "d=16, b=2^4": 160.679}   # reports/paper_reproduction.md:7-9
rng = np.random.default_rng(1)
out = {}
for name, base, d, p, gb in [("d=4, b=2^16", 1 << 16, 4, 34_359_738_367, 256), ("d=8, b=2^8", 256, 8, 1_040_407, 256),
                             ("d=16, b=2^4", 16, 16, 12_289, 16)]:
    bfv = (E.BfvParamsBuilder().ring_degree(4096).plain_modulus(p).ct_moduli([1152921504606830593])
           .aux_moduli([18014398509998081, 36028797018972161]).gadget_base(gb).build())
    dp = E.DbfvParams.new(bfv, base, d, 0)
    q, G = bfv.modulus(0), bfv.gadget_digits
    rlk = E.RelinKey(rng.integers(0, q, (G, 2, 4096), dtype=np.uint64), bfv)
    row = {"published_reference_ms": PUBLISHED_MS[name]}
    for B in (1, 64):
        a = batch.to_device(rng.integers(0, q, (B, d, 2, 4096), dtype=np.uint64))
        b = batch.to_device(rng.integers(0, q, (B, d, 2, 4096), dtype=np.uint64))
        o = torch.empty_like(a)
        for _ in range(3):
            batch.dbfv_mul(dp, a, b, rlk, out=o)
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); batch.dbfv_mul(dp, a, b, rlk, out=o); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        row[f"batch{B}"] = {"ms": ms, "dbfv_mul_per_s": B / (ms * 1e-3)}
    out[name] = row
print(json.dumps(out))
