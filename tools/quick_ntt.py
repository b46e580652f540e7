"""Quick NTT throughput probe (tools): python tools/quick_ntt.py [count]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import exacto_b200 as E
from exacto_b200 import batch
P = E.u64_dbfv().bfv_params
count = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
for idx in range(3):
    q = P.modulus(idx)
    x = batch.to_device(rng.integers(0, q, (count, 4096), dtype=np.uint64)); y = torch.empty_like(x)
    for name, fn in (("fwd", batch.ntt_forward), ("inv", batch.ntt_inverse)):
        for _ in range(3): fn(P, idx, x, out=y)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(P, idx, x, out=y); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        print(f"prime{idx} {name}: {count / ms / 1e3:.2f} M NTT/s  ({count * 65536 / ms / 1e6 / 6543.4 * 100:.1f}% of 6543 GB/s)")
