"""CPU tier: host mirror logic (params, guards, metadata), the C ABI surface and the
multi-process sharding logic.  No GPU compute is attempted here."""
import os
import re

import numpy as np
import pytest

import exacto_b200 as E
from common import H, O, ROOT


# ---- params/mod.rs ------------------------------------------------------------------------
def test_presets_match_reference_numbers():
    c = E.compact_bfv()
    assert (c.ring_degree, c.plain_modulus, c.gadget_base, c.gadget_digits) == (1024, 257, 1 << 16, 3)
    u = E.u64_dbfv()
    assert (u.base, u.num_digits, u.plain_modulus) == (256, 8, 0)
    assert (u.bfv_params.gadget_base, u.bfv_params.gadget_digits, u.bfv_params.plain_modulus) == (256, 8, 1040407)
    d = E.compact_dbfv()
    assert (d.base, d.num_digits, d.plain_modulus, d.bfv_params.plain_modulus) == (16, 2, 256, 929)
    assert E.small_bfv().aux_basis is None
    assert E.cfg3_prime_dbfv().bfv_params.gadget_digits == 4


def test_builder_validation():
    with pytest.raises(E.ExactoError) as e:
        E.BfvParamsBuilder().ring_degree(1000).ct_moduli([65537]).build()
    assert e.value.kind == "InvalidRingDegree"
    with pytest.raises(E.ExactoError, match="at least one ciphertext modulus"):
        E.BfvParamsBuilder().build()
    with pytest.raises(E.ExactoError, match="plaintext modulus must be >= 2"):
        E.BfvParamsBuilder().ring_degree(16).ct_moduli([65537]).plain_modulus(1).build()
    # PRIMES_4096 of params/presets.rs:9-13 are not prime: plan creation must refuse
    with pytest.raises(E.ExactoError, match="cannot create NTT plan"):
        E.BfvParamsBuilder().ring_degree(4096).ct_moduli([0xFFFFFFFFFFE00001]).build()
    with pytest.raises(E.ExactoError, match="base\\^digits"):
        E.DbfvParams.new(E.compact_bfv(), 16, 1, 256)
    with pytest.raises(E.ExactoError, match="base must be >= 2"):
        E.DbfvParams.new(E.compact_bfv(), 1, 8, 256)


def test_compute_gadget_digits():
    assert E.compute_gadget_digits([1099509805057], 1 << 16) == 3
    assert E.compute_gadget_digits([1152921504606830593], 256) == 8
    assert E.compute_gadget_digits([65537, 1099509805057], 8) == 19      # bfv/eval.rs:903-911 params
    assert E.compute_gadget_digits([65537], 1 << 16) == 2


# ---- guards of dbfv_mul / bfv_mul_and_relin fire before any device work ------------------
def _fake_dbfv(params, limbs=None, depth=0, comps=2):
    n = params.bfv_params.ring_degree
    arr = np.zeros((params.num_digits if limbs is None else limbs, comps, n), np.uint64)
    return E.DbfvCiphertext.from_array(arr, params, mul_depth=depth)


def test_dbfv_mul_guards():
    p = E.compact_dbfv()
    rlk = E.RelinKey(np.zeros((3, 2, 1024), np.uint64), p.bfv_params)
    with pytest.raises(E.ExactoError, match="multiplication requires d-limb ciphertexts") as e:
        E.dbfv_mul(_fake_dbfv(p, limbs=3), _fake_dbfv(p), rlk)
    assert e.value.kind == "InvalidParam"
    with pytest.raises(E.ExactoError, match="chained dBFV multiplication requires ciphertext-level lattice reduction") as e:
        E.dbfv_mul(_fake_dbfv(p, depth=1), _fake_dbfv(p), rlk)        # dbfv/eval.rs:292-313
    assert e.value.kind == "NotImplemented"
    with pytest.raises(E.ExactoError, match="multiplication requires degree-1 ciphertexts"):
        E.dbfv_mul(_fake_dbfv(p, comps=3), _fake_dbfv(p), rlk)
    with pytest.raises(E.ExactoError, match="multiplication requires degree-1 ciphertexts"):
        E.bfv_mul_and_relin(E.BfvCiphertext.from_array(np.zeros((3, 1024), np.uint64), p.bfv_params),
                            E.BfvCiphertext.from_array(np.zeros((2, 1024), np.uint64), p.bfv_params), rlk)
    with pytest.raises(E.ExactoError) as e:
        E.dbfv_add(_fake_dbfv(p, limbs=1), _fake_dbfv(p))
    assert e.value.kind == "DimensionMismatch"


def test_chain_requires_nonempty():
    p = E.compact_dbfv()
    rlk = E.RelinKey(np.zeros((3, 2, 1024), np.uint64), p.bfv_params)
    with pytest.raises(E.ExactoError, match="requires at least one ciphertext"):
        E.dbfv_mul_chain_then_bootstrap([], rlk, E.BootstrapKey(p.bfv_params, rlk))
    one = _fake_dbfv(p)
    assert E.dbfv_mul_chain_then_bootstrap([one], rlk, E.BootstrapKey(p.bfv_params, rlk)) is one


# ---- the C ABI library loads here and exports every symbol the header declares -----------
def test_abi_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "exacto_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(exb_[a-z0-9_]+)\s*\(", header))
    from exacto_b200 import _native
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    for name in declared:
        assert hasattr(native_lib, name)
    assert b"sm_100a" in native_lib.exb_version()


def test_small_reps_through_abi(native_lib):
    assert E.small_reps(16, 2, 250).tolist() == [[6, 0]]
    assert not E.small_reps(256, 8, 0).any()
    assert np.array_equal(E.small_reps(3, 3, 20), O.small_reps(3, 3, 20))


def test_no_cpu_fallback_without_gpu(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.ExactoError) as e:
        E.compact_bfv().context(0)
    assert e.value.kind == "Cuda"


def test_product_never_imports_oracle():
    """The product must not import, include, link or load the oracle or the host emulator."""
    pkg = os.path.join(ROOT, "exacto_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, re.M), f
                assert not re.search(r"#\s*include\s*[<\"][^>\"]*(oracle|host_emul)", src), f
                assert "libexacto_oracle" not in src and "libexb_emul" not in src, f


# ---- sharding (pure functions + a world_size-2 gloo run) ---------------------------------------
def test_pair_range_and_limb_masks():
    from exacto_b200.sharding import limb_masks, pair_range
    for batch in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            spans = [pair_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    for d in (1, 2, 8, 16):
        for world in (1, 2, 4, 8):
            masks = limb_masks(d, world)
            assert sum(masks) == (1 << d) - 1 and all(a & b == 0 for i, a in enumerate(masks) for b in masks[i + 1:])
    loads = [sum(k + 1 for k in range(8) if (m >> k) & 1) for m in limb_masks(8, 2)]
    assert loads == [18, 18]


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from exacto_b200.sharding import gather_limbs, gather_wire_bytes_per_rank, limb_masks, limb_owner, pair_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    P, base, d, pm = H.toy(16), 16, 2, 256
    rng = np.random.default_rng(99)
    ct1 = rng.integers(0, P.q, (4, d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (4, d, 2, P.n), dtype=np.uint64)
    rlk = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    full = np.stack([O.dbfv_mul(P, base, d, pm, a, b, rlk) for a, b in zip(ct1, ct2)])
    # (1) batch-parallel: every rank computes its pair slice, no collective on the data path
    lo, hi = pair_range(4, rank, world)
    mine = np.stack([O.dbfv_mul(P, base, d, pm, a, b, rlk) for a, b in zip(ct1[lo:hi], ct2[lo:hi])])
    assert np.array_equal(mine, full[lo:hi])
    # (2) k-sharded: this rank fills only its limbs, one all-gather completes the result
    masks = limb_masks(d, world)
    part = np.zeros_like(full)
    for k in range(d):
        if (masks[rank] >> k) & 1:
            part[:, k] = full[:, k]
    got = gather_limbs(torch.from_numpy(part.view(np.int64)), masks).numpy().view(np.uint64)
    assert np.array_equal(got, full)
    # only owned limbs travel: each rank receives exactly the limbs it does not own
    limb_bytes = 2 * P.n * 8
    assert sum(gather_wire_bytes_per_rank(masks, d, r, limb_bytes) for r in range(world)) == d * limb_bytes * (world - 1)
    assert sorted(limb_owner(masks, d)) == sorted(r for r in range(world) for k in range(d) if (masks[r] >> k) & 1)
    # u64 profile (d = 8) split over 3 ranks with uneven limb counts
    m3 = limb_masks(8, world)
    t8 = torch.arange(3 * 8 * 2 * 4, dtype=torch.int64).reshape(3, 8, 2, 4)
    mine8 = torch.zeros_like(t8)
    for k in range(8):
        if (m3[rank] >> k) & 1:
            mine8[:, k] = t8[:, k]
    assert torch.equal(gather_limbs(mine8, m3), t8)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_gloo_world2_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_cpp_host_mirror_builds_and_links(native_lib):
    """include/exacto_b200.hpp compiles against the C ABI and the driver links the in-tree library."""
    import subprocess
    import __graft_entry__ as g
    exe = g.build_cpp_driver()
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 2 and "usage" in res.stderr


def test_modular_helpers_reference_kats():
    """ring/modular.rs:127-205 restated on exacto_b200.modular (host scalars), plus agreement with the oracle."""
    from exacto_b200 import modular as M
    from common import O
    m = 65537
    bk = M.barrett_constant(m)
    assert [M.barrett_reduce(a, m, bk) for a in (0, 1, m, m + 1, 123456789)] == [0, 1, 0, 1, 123456789 % m]
    assert M.mod_mul(1234, 5678, m, bk) == 1234 * 5678 % m and M.mod_mul(0, 5678, m, bk) == 0 and M.mod_mul(1, 5678, m, bk) == 5678
    assert M.mod_add(100, 200, m) == 300 and M.mod_add(m - 1, 2, m) == 1
    assert M.mod_sub(200, 100, m) == 100 and M.mod_sub(100, 200, m) == m - 100
    assert M.mod_neg(0, m) == 0 and M.mod_neg(1, m) == m - 1 and M.mod_add(100, M.mod_neg(100, m), m) == 0
    assert M.mod_pow(2, 10, m) == 1024 and M.mod_pow(2, 16, m) == 65536 and M.mod_pow(3, 0, m) == 1
    assert M.mod_mul(12345, M.mod_inv(12345, m), m, bk) == 1 and M.mod_inv(2, 4) is None
    mi = M.montgomery_inv_neg(m)
    assert (m * mi + 1) % (1 << 64) == 0 and M.montgomery_reduce(12345 * m, m, mi) == 0
    q = 1152921504606830593
    for a, b in [(q - 1, q - 2), (12345678901234567, 98765432109876543), (0, 7)]:
        assert M.mod_mul(a, b, q) == O.mod_mul(a, b, q) and M.mod_add(a, b, q) == O.mod_add(a, b, q)
        assert M.mod_sub(a, b, q) == O.mod_sub(a, b, q)
    assert M.montgomery_reduce((q - 5) * (q - 7), q, M.montgomery_inv_neg(q)) == (q - 5) * (q - 7) * pow(1 << 64, -1, q) % q
