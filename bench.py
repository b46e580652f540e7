#!/usr/bin/env python
"""bench.py -- dBFV u64-profile ciphertext multiplications/s (and batched NTTs/s at n=4096).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm

A "step" is one dbfv_mul (dbfv/eval.rs:82-149) over a batch of independent synthetic ciphertext
pairs of the paper_repro u64 profile (BASELINE.json configs[3]: n=4096, q=1152921504606830593,
aux 18014398509998081 & 36028797018972161, BFV p=1040407, B=256 -> G=8; dBFV p=2^64, b=256, d=8).
Rank r of an N-GPU run owns its own `--pairs` pairs (weak scaling, no data-path collective).

`value`  : dbfv_mul/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the host-buffer C-ABI call (exb_dbfv_mul_host): pinned host
           inputs -> H2D -> kernels -> D2H inside the timed region.
`roofline`: the dominant kernel of the step (tensor+scale), algorithmic bytes / measured duration.
`ntt`    : batched forward / inverse NTT throughput at n=4096 with its HBM roofline fraction.
`cpu_baseline`: oracle/exacto_oracle.c (C restatement of the reference's CPU schedule, OpenMP over
           the d^2 products like rayon) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PUBLISHED_MS_PER_DBFV_MUL = 31.395      # reports/paper_reproduction.md:9 = BASELINE.md section 1, target profile row
PUBLISHED_DBFV_MUL_PER_S = 1000.0 / PUBLISHED_MS_PER_DBFV_MUL   # 31.9 dbfv_mul/s: unstated CPU, all cores (rayon)
WORKLOAD = ("paper_repro u64 profile: dbfv_mul, n=4096, q=1152921504606830593 (60 bit), aux "
            "18014398509998081 & 36028797018972161, BFV p=1040407, gadget B=256 G=8, dBFV p=2^64 b=256 d=8")
N, D, A, G = 4096, 8, 2, 8
# dram__bytes_read.sum + dram__bytes_write.sum of one tensor01_kernel launch at the default workload (148 pairs),
# from profiles/r01_ncu_fused_kernels_final.json (ncu --set full on tools/prof.py mul 148)
TENSOR01_DRAM_BYTES = 398925568 + 72187648


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle takes an explicit count)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth(seed: int, pairs: int, q: int):
    rng = np.random.default_rng(seed)
    ct1 = rng.integers(0, q, (pairs, D, 2, N), dtype=np.uint64)
    ct2 = rng.integers(0, q, (pairs, D, 2, N), dtype=np.uint64)
    return ct1, ct2


class CpuWork:
    """One synthetic u64-profile pair for the oracle port (paper_repro.rs:144-164 times fixed inputs too)."""

    def __init__(self, seed: int = 1337 + D):
        import oracle as O
        from oracle import harness as H
        self.O, self.S = O, H.u64_dbfv()
        ct1, ct2 = synth(seed, 1, self.S.bfv.q)
        self.ct1, self.ct2 = ct1[0], ct2[0]
        self.rlk = np.random.default_rng(seed + 1).integers(0, self.S.bfv.q, (G, 2, N), dtype=np.uint64)
        self.run(1, 1)                                   # builds the NTT plans outside any timed region

    def run(self, trials: int, threads: int) -> float:
        """Seconds for `trials` dbfv_mul calls on `threads` host threads."""
        S = self.S
        t0 = time.perf_counter()
        for _ in range(trials):
            self.O.dbfv_mul(S.bfv, S.base, S.d, S.plain_modulus, self.ct1, self.ct2, self.rlk, threads=threads)
        return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path.  The Rust crate cannot be built
    in this image (no cargo/rustc), so this is the oracle port (oracle/exacto_oracle.c): same op
    sequence, u128 % reduction, OpenMP over the d^2 products like rayon, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = 2                                   # dbfv_muls per step (bounded sample of the workload)
    work = CpuWork()
    for _ in range(args.warmup):
        work.run(sample, threads)
    elapsed = 0.0
    for _ in range(args.steps):
        elapsed += work.run(sample, threads)
    value = sample * args.steps / elapsed if elapsed > 0 else 0.0
    ms_per_step = elapsed / max(args.steps, 1) * 1e3
    line = {
        "impl": "reference", "metric": "dbfv_mul_per_s", "value": value, "unit": "dbfv_mul/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": value / PUBLISHED_DBFV_MUL_PER_S, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} dbfv_mul per step on one synthetic pair"},
        "cpu_baseline": {"value": value, "unit": "dbfv_mul/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} dbfv_mul per step x {args.steps} steps, all 64 products (reference schedule)"},
        "e2e": {"value": value, "unit": "dbfv_mul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "published_anchor_ms_per_dbfv_mul": PUBLISHED_MS_PER_DBFV_MUL,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so the pinned staging buffers of the e2e leg are
    first-touched on the GPU's NUMA node (matters when several ranks share the host).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import exacto_b200 as E
    from exacto_b200 import _native, batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    params = E.u64_dbfv()
    P = params.bfv_params
    q = P.modulus(0)
    pairs = args.pairs
    ct1_h, ct2_h = synth(0xE8AC70 + rank, pairs, q)
    rlk_arr = np.random.default_rng(0x51AB).integers(0, q, (G, 2, N), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, P)
    ct1, ct2 = batch.to_device(ct1_h, dev), batch.to_device(ct2_h, dev)
    out = torch.empty_like(ct1)
    ctx = P.context(local)
    L = _native.lib()

    if args.kshard and world > 1:
        # strong-scaling mode: every rank holds the SAME pairs, computes only its output limbs (products with
        # equal i+j stay on one rank), then one NCCL all-gather of the limbs -- the path's only exchange step
        from exacto_b200.sharding import limb_masks
        ct1_h, ct2_h = synth(0xE8AC70, pairs, q)
        ct1, ct2 = batch.to_device(ct1_h, dev), batch.to_device(ct2_h, dev)
        out = torch.zeros_like(ct1)
        masks = limb_masks(D, world)
        my_limbs = [k for k in range(D) if (masks[rank] >> k) & 1]
        owner = [next(r for r in range(world) if (masks[r] >> k) & 1) for k in range(D)]
        gathered = [torch.empty_like(out) for _ in range(world)]

        def step():
            if masks[rank]:
                batch.dbfv_mul(params, ct1, ct2, rlk, out=out, limb_mask=masks[rank])
            dist.all_gather(gathered, out)
            for k in range(D):
                if owner[k] != rank:
                    out[:, k].copy_(gathered[owner[k]][:, k])
    else:
        args.kshard = False

        def step():
            batch.dbfv_mul(params, ct1, ct2, rlk, out=out, all_products=args.all_products)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    _native.check(L.exb_profile_enable(ctx.handle, 1))
    launches0 = batch.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = batch.launch_count() - launches0
    stage_ms = (ctypes.c_double * 5)()
    stage_n = (ctypes.c_ulonglong * 5)()
    _native.check(L.exb_profile_read(ctx.handle, stage_ms, stage_n))
    _native.check(L.exb_profile_enable(ctx.handle, 0))
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = pairs * (1 if args.kshard else world) * args.steps / (elapsed_ms * 1e-3)
    if args.kshard:       # verify the gathered result against a full local computation (outside the timed region)
        full = batch.dbfv_mul(params, ct1, ct2, rlk)
        torch.cuda.synchronize()
        assert torch.equal(full, out), "k-sharded + all-gather result differs from the local dbfv_mul"

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H per step) -----------------------
    e2e_pairs = min(pairs, args.e2e_pairs)
    h1 = torch.from_numpy(ct1_h[:e2e_pairs].view(np.int64)).pin_memory()
    h2 = torch.from_numpy(ct2_h[:e2e_pairs].view(np.int64)).pin_memory()
    ho = torch.empty_like(h1).pin_memory()
    flags = _native.EXB_DBFV_ALL_PRODUCTS if args.all_products else 0

    def e2e_step():
        _native.check(L.exb_dbfv_mul_host(ctx.handle, params.base, D, params.plain_modulus, h1.data_ptr(),
                                          h2.data_ptr(), rlk.native(ctx), ho.data_ptr(), e2e_pairs, flags))

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = e2e_pairs * world * args.steps / e2e_s
    checksum = int(ho[0, 0, 0, :4].sum().item())          # D2H result actually read on the host
    ct_bytes = D * 2 * N * 8

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        n_products = 64 if args.all_products else 36
        n_limbs = 15 if args.all_products else 8
        small = os.environ.get("EXB_AUX_BASIS") != "reference"
        per_limb = small and stage_n[2] > 0          # tensor01_kernel (comps 0/1 per output limb) + per-product comp 2
        # algorithmic bytes per pair of the dominant kernel: unique inputs (both operands in
        # every base it reads) + its outputs
        if not small:
            in_bytes = 2 * D * 2 * (1 + A) * N * 8                      # q + two 64-bit aux bases
        else:
            in_bytes = 2 * D * 2 * N * 8 + 2 * D * 2 * 3 * N * 4        # q (u64) + three 27-bit internal primes (u32)
        if per_limb:
            kernel_name = ("tensor01_kernel (components 0/1 per output limb: per product point-wise tensor mod q + INTT + "
                           "rounding term; per limb the summed 27-bit point-wise tensors, 3 INTT32 and the exact m recombination)")
            tensor_bytes = in_bytes + n_limbs * 2 * N * 8
        else:
            kernel_name = ("tensor32_kernel" if small else "tensor_kernel") + " (per product and component: point-wise tensor, INTTs, hps_scale, gadget digits)"
            tensor_bytes = in_bytes + n_products * (2 * N * 8 + G * N * 2)
        t_ms = stage_ms[1] / max(stage_n[1], 1)
        achieved = tensor_bytes * pairs / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        stages = {nm: {"ms_per_launch": stage_ms[i] / max(stage_n[i], 1), "launches": int(stage_n[i])}
                  for i, nm in enumerate(["lift", "tensor01_per_limb" if per_limb else "tensor_scale", "tensor_c2_per_product",
                                          "relin", "reduce"]) if stage_n[i]}
        total_stage = sum(stage_ms[i] for i in range(5)) or 1.0
        line = {
            "metric": "dbfv_mul_per_s", "value": value, "unit": "dbfv_mul/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong" if args.kshard else "weak",
            "vs_baseline": value / PUBLISHED_DBFV_MUL_PER_S, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": pairs,
                       "parallelism": (f"output limbs sharded x{world} (limb_masks), one NCCL all_gather per step, same {pairs} pairs on every rank"
                                       if args.kshard else f"pairs sharded x{world}, no collective"),
                       "products_per_dbfv_mul": n_products,
                       "dead_products": "computed (reference schedule)" if args.all_products else
                       "skipped (28 of 64 products feed limbs k>=d that reduce() discards; output bit-identical)",
                       "l2": f"inputs+outputs {3 * pairs * ct_bytes / 2**20:.0f} MiB + workspace > 126 MB L2 per step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "dbfv_mul/s", "h2d_bytes_per_step": 2 * e2e_pairs * ct_bytes,
                    "d2h_bytes_per_step": e2e_pairs * ct_bytes, "pairs_per_step": e2e_pairs, "result_checksum": checksum},
            "gpu_launches": int(launches),
            "bfv_mul_and_relin_equiv_per_s": value * 64,
            "roofline": {"kernel": kernel_name,
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": (TENSOR01_DRAM_BYTES if (pairs == 148 and per_limb and not args.all_products) else None),
                         "algorithmic_bytes": tensor_bytes * pairs,
                         "peak_source": peak_src, "share_of_step": stage_ms[1] / total_stage,
                         "note": "integer-pipe bound kernel: see DESIGN.md; HBM fraction is reported, not the target",
                         # what actually bounds it (ncu --set full of this kernel at this workload,
                         # profiles/r01_ncu_fused_kernels_final.json): the integer-multiply pipe
                         "binding_pipe": ({"metric": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                           "busy_frac": 0.523, "issue_active_frac": 0.502} if per_limb else None)},
            "stages": stages,
            "published_anchor_ms_per_dbfv_mul": PUBLISHED_MS_PER_DBFV_MUL,
        }
        if not args.no_ntt:
            line["ntt"] = bench_ntt(torch, batch, P, peak, peak_src, args)
            line["widened"] = bench_widened(torch, batch, P, peak, args)
            # one dbfv_mul alone (what src/bin/paper_repro.rs times on the CPU: 31.395 ms published)
            lat = []
            a1, b1, o1 = ct1[:1].contiguous(), ct2[:1].contiguous(), torch.empty_like(ct1[:1])
            for i in range(args.ntt_reps + 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); batch.dbfv_mul(params, a1, b1, rlk, out=o1); e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    lat.append(e0.elapsed_time(e1))
            line["single_call"] = {"dbfv_mul_ms": statistics.median(lat), "published_reference_ms": PUBLISHED_MS_PER_DBFV_MUL}
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            work = CpuWork()
            per = work.run(args.cpu_trials, threads) / args.cpu_trials
            per1 = work.run(2, 1) / 2
            line["cpu_baseline"] = {"value": 1.0 / per, "unit": "dbfv_mul/s", "cores": threads, "kind": "port",
                                    "sample": f"mean of {args.cpu_trials} dbfv_mul on one synthetic pair, OpenMP over the 64 products",
                                    "one_thread_value": 1.0 / per1,
                                    "note": "C restatement of the reference schedule (scalar NTT, u128 %); the Rust crate cannot be built here"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def bench_ntt(torch, batch, P, peak, peak_src, args):
    """Batched NTT/INTT at n=4096: `count` polynomials (> L2), CUDA events, median of reps."""
    count = args.ntt_count
    res = {"count": count, "bytes_per_ntt": 2 * 8 * N}
    rng = np.random.default_rng(0xE8AC70)
    for idx in range(1 + A):
        q = P.modulus(idx)
        x = batch.to_device(rng.integers(0, q, (count, N), dtype=np.uint64))
        y = torch.empty_like(x)
        for name, fn in (("fwd", batch.ntt_forward), ("inv", batch.ntt_inverse)):
            for _ in range(3):
                fn(P, idx, x, out=y)
            times = []
            for _ in range(args.ntt_reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(P, idx, x, out=y); e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = statistics.median(times)
            per_s = count / (ms * 1e-3)
            gbs = per_s * 2 * 8 * N / 1e9
            res[f"{name}_prime{idx}"] = {"ntt_per_s": per_s, "ms": ms, "achieved_gbs": gbs, "frac": gbs / peak}
        del x, y
    best = max(v["frac"] for k, v in res.items() if isinstance(v, dict))
    worst = min(v["frac"] for k, v in res.items() if isinstance(v, dict))
    res["roofline"] = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                       "frac_min": worst, "frac_max": best, "target_frac": 0.60}
    return res


def bench_widened(torch, batch, P, peak, args):
    """The SURVEY 8(f) rows that got their own kernels: Galois automorphism + key switch (bfv/eval.rs:512-561) and
    decrypt (bfv/encrypt.rs:111-178), batched at n=4096 on the u64 profile's BFV parameters (inputs > L2)."""
    import exacto_b200 as E
    rng = np.random.default_rng(0xA070)
    q = P.modulus(0)
    count = 2048
    ct = batch.to_device(rng.integers(0, q, (count, 2, N), dtype=np.uint64))
    gk = E.GaloisKey(rng.integers(0, q, (G, 2, N), dtype=np.uint64), 3, P)
    sk = batch.to_device(rng.integers(0, q, N, dtype=np.uint64))
    out = torch.empty_like(ct)
    dec = torch.empty((count, N), dtype=torch.int64, device=ct.device)
    res = {"count": count}
    for name, fn, nbytes in (("automorphism_keyswitch", lambda: batch.bfv_apply_automorphism(P, ct, gk, out=out), 4 * 8 * N),
                             ("decrypt", lambda: batch.bfv_decrypt(P, ct, sk, out=dec), 3 * 8 * N)):
        for _ in range(3):
            fn()
        times = []
        for _ in range(args.ntt_reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        res[name] = {"per_s": count / (ms * 1e-3), "ms": ms, "algorithmic_bytes_each": nbytes,
                     "hbm_frac": count * nbytes / (ms * 1e-3) / 1e9 / peak}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="exacto_b200", choices=["exacto_b200", "reference"])
    ap.add_argument("--pairs", type=int, default=148, help="ciphertext pairs per GPU per step")
    ap.add_argument("--e2e-pairs", type=int, default=148)
    ap.add_argument("--all-products", action="store_true", help="compute all 64 products like the reference")
    ap.add_argument("--ntt-count", type=int, default=16384)
    ap.add_argument("--ntt-reps", type=int, default=10)
    ap.add_argument("--cpu-trials", type=int, default=20)
    ap.add_argument("--kshard", action="store_true",
                    help="N>1: shard ONE batch by output limb + NCCL all-gather (strong scaling) instead of sharding pairs")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
