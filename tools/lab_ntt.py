"""Lab: NTT kernel variants (needs a library built with -DEXB_LAB; see tools/lab_build.sh).
EXB_NTT_DBG=1 copy only, 2 transforms without global traffic, 3 twiddle loads from 16 hot entries."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import exacto_b200 as E
from exacto_b200 import batch
P = E.u64_dbfv().bfv_params
count = 16384
rng = np.random.default_rng(1)
for idx in (0, 1):
    q = P.modulus(idx)
    x = batch.to_device(rng.integers(0, q, (count, 4096), dtype=np.uint64)); y = torch.empty_like(x)
    for name, fn in (("fwd", batch.ntt_forward), ("inv", batch.ntt_inverse)):
        for _ in range(3): fn(P, idx, x, out=y)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(P, idx, x, out=y); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        print(f"DBG={os.environ.get('EXB_NTT_DBG','0')} NB={os.environ.get('EXB_NTT_NB','3')} prime{idx} {name}: {count/ms/1e3:.2f} M NTT/s  ({count*65536/ms/1e6/6543.4*100:.1f}% of HBM)")
