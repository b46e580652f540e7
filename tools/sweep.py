"""Batch-size sweep of the ct-mul entry points over the BASELINE configs, device resident,
CUDA events, median of reps.  Prints one JSON object (goes to profiles/, not a bench line).

    python tools/sweep.py [reps]
"""
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

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
rng = np.random.default_rng(5)


def timed(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


out = {}
cfgs = [("cfg1 compact_bfv: bfv_mul_and_relin n=1024", E.compact_bfv(), None),
        ("cfg2 compact_dbfv: dbfv_mul n=1024 d=2", None, E.compact_dbfv()),
        ("cfg3' README+aux: dbfv_mul n=4096 d=2", None, E.cfg3_prime_dbfv()),
        ("cfg4 u64 profile: dbfv_mul n=4096 d=8", None, E.u64_dbfv()),
        ("cfg4 u64 profile: bfv_mul_and_relin n=4096", E.u64_dbfv().bfv_params, None)]
for name, bfv, dbfv in cfgs:
    P = bfv if bfv is not None else dbfv.bfv_params
    q, n, G = P.modulus(0), P.ring_degree, P.gadget_digits
    rlk = E.RelinKey(rng.integers(0, q, (G, 2, n), dtype=np.uint64), P)
    rows = {}
    for B in (1, 8, 64, 512):
        shape = (B, 2, n) if dbfv is None else (B, dbfv.num_digits, 2, n)
        a = batch.to_device(rng.integers(0, q, shape, dtype=np.uint64))
        b = batch.to_device(rng.integers(0, q, shape, dtype=np.uint64))
        o = torch.empty_like(a)
        if dbfv is None:
            ms = timed(lambda: batch.bfv_mul_and_relin(P, a, b, rlk, out=o))
        else:
            ms = timed(lambda: batch.dbfv_mul(dbfv, a, b, rlk, out=o))
        rows[str(B)] = {"ms": ms, "per_s": B / (ms * 1e-3)}
        del a, b, o
    out[name] = rows
# batched NTT at the three SURVEY 8(d) batch sizes (inputs 128 MiB .. 2 GiB, all > L2), u64-profile primes
P = E.u64_dbfv().bfv_params
ntt = {}
for count in (4096, 16384, 65536):
    for idx in range(3):
        q = P.modulus(idx)
        x = batch.to_device(rng.integers(0, q, (count, 4096), dtype=np.uint64))
        y = torch.empty_like(x)
        for nm, fn in (("fwd", batch.ntt_forward), ("inv", batch.ntt_inverse)):
            ms = timed(lambda: fn(P, idx, x, out=y))
            ntt[f"{nm}_prime{idx}_x{count}"] = {"ms": ms, "ntt_per_s": count / (ms * 1e-3),
                                                 "gbs": count * 65536 / (ms * 1e-3) / 1e9}
        del x, y
out["ntt_n4096"] = ntt
print(json.dumps(out))
