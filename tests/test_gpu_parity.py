"""GPU tier: parity of the CUDA path (through the C ABI of libexacto_b200.so) with the oracle,
the golden fixtures and the reference's own known-answer tests.  Integer arithmetic: the
bar is bit-exact.  Citations are into /root/reference/src/."""
import numpy as np
import pytest

import exacto_b200 as E
from common import CASES, H, O, digest, golden, golden_inputs, to_dbfv_params, to_params

pytestmark = pytest.mark.gpu

import os as _os
TESTS_DIR = _os.path.dirname(_os.path.abspath(__file__))
ROOT_DIR = _os.path.dirname(TESTS_DIR)

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(native_lib):
    assert torch.cuda.is_available(), "the gpu tier needs a CUDA device"


# ---- ring/ntt.rs:170-212 + per-prime parity ------------------------------------------------------
def test_reference_ntt_tests_n16():
    n, q = 16, 65537
    plan = E.make_plan(n, q)
    a = E.CoeffPoly.from_coeffs(np.array([1, 2, 3, 4] + [0] * 12), q)
    b = E.CoeffPoly.from_coeffs(np.array([5, 6, 7, 8] + [0] * 12), q)
    na, nb = E.NttPoly.from_coeff_poly(a, plan), E.NttPoly.from_coeff_poly(b, plan)
    assert na.to_coeff_poly() == a                                                   # :170-178
    assert np.array_equal(na.mul(nb).to_coeff_poly().coeffs, O.poly_mul_naive(a.coeffs, b.coeffs, q))   # :181-195
    assert np.array_equal(na.add(nb).to_coeff_poly().coeffs, (a.coeffs + b.coeffs) % np.uint64(q))      # :198-212
    assert np.array_equal(na.evals, O.ntt_fwd(a.coeffs, q))
    # X^3 * X^3 = -X^2 (ring/poly.rs:195-202) through the NTT at n=16
    x3 = E.CoeffPoly.from_coeffs(np.eye(16, dtype=np.uint64)[3], q)
    sq = E.NttPoly.from_coeff_poly(x3, plan)
    assert np.array_equal(sq.mul(sq).to_coeff_poly().coeffs, np.eye(16, dtype=np.uint64)[6])
    x15 = E.NttPoly.from_coeff_poly(E.CoeffPoly.from_coeffs(np.eye(16, dtype=np.uint64)[15], q), plan)
    want = np.zeros(16, np.uint64); want[14] = q - 1
    assert np.array_equal(x15.mul(x15).to_coeff_poly().coeffs, want)


@pytest.mark.parametrize("n,q", [(1024, 1099509805057), (1024, 562949953443841), (4096, 1152921504606830593),
                                 (4096, 18014398509998081), (4096, 36028797018972161), (4096, 576460752308273153),
                                 (2048, 1152921504606830593), (8192, 1152921504606830593), (2, 65537)])
def test_ntt_vs_oracle(n, q):
    from exacto_b200 import batch
    plan = E.make_plan(n, q)
    rng = np.random.default_rng(n + q % 1000)
    x = rng.integers(0, q, (5, n), dtype=np.uint64)
    x[0, :2] = [0, q - 1]
    x[1] = q - 1
    x[2] = 0
    dx = batch.to_device(x)
    fy = batch.ntt_forward(plan.params, 0, dx)
    assert np.array_equal(batch.to_host(fy), O.ntt_fwd(x, q))
    assert np.array_equal(batch.to_host(batch.ntt_inverse(plan.params, 0, fy)), x)
    assert batch.to_host(batch.ntt_forward(plan.params, 0, dx, out=dx)).tobytes() == batch.to_host(fy).tobytes()   # in place
    g = golden()
    key = f"ntt/{n}_{q}"
    if key in g:
        ramp = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(12345)) % np.uint64(q)
        assert digest(batch.to_host(batch.ntt_forward(plan.params, 0, batch.to_device(ramp)))) == str(g[key])


@pytest.mark.parametrize("count", [1, 3, 297, 899, 2 * 296 * 3 + 5])
def test_ntt_tma_pipeline_many_polys(count):
    """The TMA transform kernel's three-image ring (mbarrier phases, store drain before reload) over several
    iterations per CTA, ragged tails and in place, word for word against the oracle; the cp.async kernel (option
    "ntt_cp_async") must give the same words."""
    from exacto_b200 import batch
    P = E.u64_dbfv().bfv_params
    rng = np.random.default_rng(count)
    for idx in (0, 1):                                                   # LAZY 1 (60-bit q) and LAZY 2 (54-bit aux prime)
        q = P.modulus(idx)
        x = rng.integers(0, q, (count, 4096), dtype=np.uint64)
        x[0, :3] = [0, q - 1, 1]
        want = O.ntt_fwd(x, q, threads=O.max_threads())
        dx = batch.to_device(x)
        fy = batch.ntt_forward(P, idx, dx)
        assert np.array_equal(batch.to_host(fy), want)
        assert np.array_equal(batch.to_host(batch.ntt_inverse(P, idx, fy)), x)
        batch.ntt_forward(P, idx, dx, out=dx)                            # in place
        assert np.array_equal(batch.to_host(dx), want)
        ctx = P.context()
        ctx.set_option("ntt_cp_async", 1)
        try:
            assert np.array_equal(batch.to_host(batch.ntt_inverse(P, idx, dx)), x)
            assert np.array_equal(batch.to_host(batch.ntt_forward(P, idx, batch.to_device(x))), want)
        finally:
            ctx.set_option("ntt_cp_async", 0)


def test_ntt_linearity_full_size():
    """Size-independent property at the bench size: NTT(a + b) = NTT(a) + NTT(b), round trip."""
    from exacto_b200 import batch
    P = E.u64_dbfv().bfv_params
    for idx in range(3):
        q = P.modulus(idx)
        rng = np.random.default_rng(idx)
        a = rng.integers(0, q, (512, 4096), dtype=np.uint64); b = rng.integers(0, q, (512, 4096), dtype=np.uint64)
        s = np.array((a.astype(object) + b.astype(object)) % q, dtype=np.uint64)
        fa, fb, fs = (batch.to_host(batch.ntt_forward(P, idx, batch.to_device(v))) for v in (a, b, s))
        assert np.array_equal(np.array((fa.astype(object) + fb.astype(object)) % q, dtype=np.uint64), fs)
        assert np.array_equal(batch.to_host(batch.ntt_inverse(P, idx, batch.to_device(fs))), s)


def test_ring_types_ops():
    q, n = 1099509805057, 1024
    plan = E.make_plan(n, q)
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, q, n, dtype=np.uint64), rng.integers(0, q, n, dtype=np.uint64)
    A, B = E.NttPoly(a, q, plan), E.NttPoly(b, q, plan)
    ao, bo = a.astype(object), b.astype(object)
    assert np.array_equal(A.add(B).evals, np.array((ao + bo) % q, dtype=np.uint64))
    assert np.array_equal(A.sub(B).evals, np.array((ao - bo) % q, dtype=np.uint64))
    assert np.array_equal(A.neg().evals, np.array((-ao) % q, dtype=np.uint64))
    assert np.array_equal(A.mul(B).evals, np.array((ao * bo) % q, dtype=np.uint64))
    assert np.array_equal(A.scalar_mul(2 ** 63 + 5).evals, np.array((ao * ((2 ** 63 + 5) % q)) % q, dtype=np.uint64))
    with pytest.raises(E.ExactoError) as e:
        A.add(E.NttPoly(a[:16], 65537, E.make_plan(16, 65537)))
    assert e.value.kind == "ModulusMismatch"
    cp = E.CoeffPoly(a, q)
    assert cp.add(E.CoeffPoly(b, q)) == E.CoeffPoly(np.array((ao + bo) % q, dtype=np.uint64), q)
    r = E.RnsPoly.from_coeff_poly(cp, plan.params)
    assert r.to_coeff_poly() == cp and r.mul(r).num_components() == 1


# ---- golden fixtures through bfv_mul_and_relin / dbfv_mul ----------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
def test_golden_cases(name):
    P, base, d, pm, seed, full = CASES[name]
    g = golden()
    ct1, ct2, rlk_arr = golden_inputs(P, d, seed)
    assert digest(ct1, ct2, rlk_arr) == str(g[f"{name}/in_sha256"])
    params = to_dbfv_params(P, base, d, pm)
    rlk = E.RelinKey(rlk_arr, params.bfv_params)
    out = E.dbfv_mul(E.DbfvCiphertext.from_array(ct1, params), E.DbfvCiphertext.from_array(ct2, params), rlk)
    assert (out.degree, out.mul_depth, out.num_limbs()) == (d, 1, d)               # dbfv/eval.rs:138-146
    assert digest(out.to_array()) == str(g[f"{name}/dbfv_sha256"])
    one = E.bfv_mul_and_relin(E.BfvCiphertext.from_array(ct1[0], params.bfv_params),
                              E.BfvCiphertext.from_array(ct2[0], params.bfv_params), rlk)
    assert len(one.c) == 2 and digest(one.to_array()) == str(g[f"{name}/bfv_sha256"])
    if full:
        assert np.array_equal(out.to_array(), g[f"{name}/dbfv_out"])


@pytest.mark.parametrize("name,batch_n", [("compact_dbfv", 5), ("cfg3p_dbfv", 3), ("u64_dbfv", 2), ("n64_a2_rep", 7)])
def test_batched_vs_oracle_host_and_device(name, batch_n):
    from exacto_b200 import batch
    P, base, d, pm, seed, _ = CASES[name]
    rng = np.random.default_rng(seed + 7)
    ct1 = rng.integers(0, P.q, (batch_n, d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (batch_n, d, 2, P.n), dtype=np.uint64)
    rlk_arr = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    want = np.stack([O.dbfv_mul(P, base, d, pm, a, b, rlk_arr, threads=O.max_threads()) for a, b in zip(ct1, ct2)])
    params = to_dbfv_params(P, base, d, pm)
    rlk = E.RelinKey(rlk_arr, params.bfv_params)
    assert np.array_equal(E.dbfv_mul_batch(params, ct1, ct2, rlk), want)                        # host buffers
    assert np.array_equal(E.dbfv_mul_batch(params, ct1, ct2, rlk, all_products=True), want)     # reference's d^2 schedule
    d1, d2 = batch.to_device(ct1), batch.to_device(ct2)
    assert np.array_equal(batch.to_host(batch.dbfv_mul(params, d1, d2, rlk)), want)             # device resident
    # k-sharded: two "ranks" own disjoint limbs
    from exacto_b200.sharding import limb_masks
    acc = torch.zeros_like(d1)
    for m in limb_masks(d, 2):
        if m:
            batch.dbfv_mul(params, d1, d2, rlk, out=acc, limb_mask=m)
    assert np.array_equal(batch.to_host(acc), want)
    bw = O.bfv_mul_and_relin(P, ct1[:, 0], ct2[:, 0], rlk_arr, threads=O.max_threads())
    assert np.array_equal(E.bfv_mul_and_relin_batch(params.bfv_params, ct1[:, 0], ct2[:, 0], rlk), bw)
    assert np.array_equal(batch.to_host(batch.bfv_mul_and_relin(params.bfv_params, batch.to_device(ct1[:, 0]),
                                                                 batch.to_device(ct2[:, 0]), rlk)), bw)


def test_edge_patterns_cfg4():
    """0, 1, floor(q/2), floor(q/2)+1, q-1 patterns (the centring / rounding / wrap branches)."""
    P = H.u64_dbfv().bfv
    q, n = P.q, P.n
    rng = np.random.default_rng(77)
    rlk_arr = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    pats = [np.zeros(n, np.uint64), np.ones(n, np.uint64), np.full(n, q // 2, np.uint64),
            np.full(n, q // 2 + 1, np.uint64), np.full(n, q - 1, np.uint64)]
    ct1 = np.stack([np.stack([O.ntt_fwd(pats[i], q), O.ntt_fwd(pats[(i + 1) % 5], q)]) for i in range(5)])
    ct2 = np.stack([np.stack([O.ntt_fwd(pats[(i + 2) % 5], q), O.ntt_fwd(pats[(i + 3) % 5], q)]) for i in range(5)])
    params = to_params(P)
    got = E.bfv_mul_and_relin_batch(params, ct1, ct2, E.RelinKey(rlk_arr, params))
    assert np.array_equal(got, O.bfv_mul_and_relin(P, ct1, ct2, rlk_arr, threads=O.max_threads()))
    zero = np.zeros((1, 2, n), np.uint64)
    assert not E.bfv_mul_and_relin_batch(params, zero, ct2[:1], E.RelinKey(rlk_arr, params)).any()
    assert E.bfv_mul_and_relin_batch(params, ct1[:0], ct2[:0], E.RelinKey(rlk_arr, params)).shape == (0, 2, n)   # empty batch


# ---- decrypt KATs of the reference with valid keys ----------------------------------------------------------
def test_compact_bfv_kats():
    """bfv/eval.rs:883-900 (3 * 7 = 21) + add/sub (:845-881)."""
    P, params = H.compact_bfv(), E.compact_bfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(P, rng); rlk = E.RelinKey(H.gen_relin_key(P, s, rng), params)
    enc = lambda m: E.BfvCiphertext.from_array(H.encrypt_sk(P, H.encode_scalar(P, m), s, rng), params)
    assert int(H.decrypt(P, E.bfv_mul_and_relin(enc(3), enc(7), rlk).to_array(), s)[0]) == 21
    assert int(H.decrypt(P, E.bfv_add(enc(10), enc(20)).to_array(), s)[0]) == 30
    assert int(H.decrypt(P, E.bfv_sub(enc(50), enc(20)).to_array(), s)[0]) == 30
    assert int(H.decrypt(P, E.bfv_neg(enc(5)).to_array(), s)[0]) == 257 - 5


def test_compact_dbfv_kats():
    """dbfv/eval.rs:224-290: 3*7, (3+X)(2+X), 15*15 / 10*20 / 12*12 mod 256, add, sub."""
    S, params = H.compact_dbfv(), E.compact_dbfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(S.bfv, rng); rlk = E.RelinKey(H.gen_relin_key(S.bfv, s, rng), params.bfv_params)
    enc = lambda v: E.DbfvCiphertext.from_array(H.dbfv_encrypt_sk(S, v, s, rng), params)
    dec = lambda ct: H.dbfv_decrypt(S, ct.to_array(), s)
    assert dec(E.dbfv_mul(enc(3), enc(7), rlk)) == 21
    for a, b in [(15, 15), (10, 20), (12, 12)]:
        assert dec(E.dbfv_mul(enc(a), enc(b), rlk)) == a * b % 256
    assert dec(E.dbfv_add(enc(10), enc(20))) == 30 and dec(E.dbfv_sub(enc(50), enc(20))) == 30
    pa = np.zeros(S.bfv.n, np.uint64); pb = np.zeros(S.bfv.n, np.uint64)
    pa[:2] = [3, 1]; pb[:2] = [2, 1]
    prod = E.dbfv_mul(E.DbfvCiphertext.from_array(H.dbfv_encrypt_poly_sk(S, pa, s, rng), params),
                      E.DbfvCiphertext.from_array(H.dbfv_encrypt_poly_sk(S, pb, s, rng), params), rlk)
    assert H.dbfv_decrypt_poly(S, prod.to_array(), s)[:3].tolist() == [6, 5, 1]
    with pytest.raises(E.ExactoError, match="chained dBFV multiplication requires ciphertext-level lattice reduction"):
        E.dbfv_mul(prod, enc(5), rlk)                                                 # dbfv/eval.rs:292-313


def test_u64_profile_kats():
    """dbfv/eval.rs:345-382 (BFV 3*7, 0*5, 10*20, 100*100) and :521-564 (dBFV 3*7, 1000*2000)."""
    S, params = H.u64_dbfv(), E.u64_dbfv()
    rng = np.random.default_rng(101)
    s = H.gen_secret_key(S.bfv, rng); rlk = E.RelinKey(H.gen_relin_key(S.bfv, s, rng), params.bfv_params)
    encb = lambda m: E.BfvCiphertext.from_array(H.encrypt_sk(S.bfv, H.encode_scalar(S.bfv, m), s, rng), params.bfv_params)
    for a, b in [(3, 7), (0, 5), (10, 20), (100, 100)]:
        assert int(H.decrypt(S.bfv, E.bfv_mul_and_relin(encb(a), encb(b), rlk).to_array(), s)[0]) == a * b
    enc = lambda v: E.DbfvCiphertext.from_array(H.dbfv_encrypt_sk(S, v, s, rng), params)
    assert H.dbfv_decrypt(S, E.dbfv_mul(enc(3), enc(7), rlk).to_array(), s) == 21
    assert H.dbfv_decrypt(S, E.dbfv_mul(enc(1000), enc(2000), rlk).to_array(), s) == 2_000_000
    assert H.dbfv_decrypt(S, E.dbfv_mul(enc(2 ** 32 + 3), enc(2 ** 32 + 5), rlk).to_array(), s) == ((2 ** 32 + 3) * (2 ** 32 + 5)) % 2 ** 64


def test_commutativity_full_size():
    """Size-independent property at the bench size: the integer tensor is symmetric, so
    dbfv_mul(a, b) == dbfv_mul(b, a) bit for bit; zero times anything is zero."""
    from exacto_b200 import batch
    params = E.u64_dbfv()
    P = params.bfv_params
    q, n, d = P.modulus(0), 4096, 8
    rng = np.random.default_rng(5)
    a = batch.to_device(rng.integers(0, q, (16, d, 2, n), dtype=np.uint64))
    b = batch.to_device(rng.integers(0, q, (16, d, 2, n), dtype=np.uint64))
    rlk = E.RelinKey(rng.integers(0, q, (8, 2, n), dtype=np.uint64), P)
    ab, ba = batch.dbfv_mul(params, a, b, rlk), batch.dbfv_mul(params, b, a, rlk)
    assert torch.equal(ab, ba) and ab.abs().sum().item() != 0
    assert not batch.dbfv_mul(params, torch.zeros_like(a), b, rlk).any()
    assert torch.equal(batch.dbfv_mul(params, a, b, rlk, all_products=True), ab)


# ---- error pins through the native context (dbfv/eval.rs:385-453) ---------------------------------------------
def test_misaligned_device_buffers_are_refused():
    """The n = 4096 kernels use 256-bit accesses: a device pointer that is not 32-byte aligned is an
    EXB_INVALID_PARAM, not a fault; a 32-byte-aligned view (any whole row offset) works."""
    import torch
    from exacto_b200 import batch
    dp = E.u64_dbfv()
    P = dp.bfv_params
    q, n, d = P.modulus(0), 4096, 8
    rng = np.random.default_rng(5)
    rlk = E.RelinKey(rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64), P)
    flat = batch.to_device(rng.integers(0, q, 2 * d * 2 * n + 4, dtype=np.uint64))
    good = flat[: d * 2 * n].view(1, d, 2, n)
    odd = flat[1: 1 + d * 2 * n].view(1, d, 2, n)                 # 8 bytes past a 32-byte boundary
    shifted = flat[4: 4 + d * 2 * n].view(1, d, 2, n)             # 32 bytes past: fine
    with pytest.raises(E.ExactoError, match="32-byte aligned"):
        batch.dbfv_mul(dp, odd, good, rlk)
    a = batch.dbfv_mul(dp, shifted, good, rlk)
    b = batch.dbfv_mul(dp, shifted.clone(), good, rlk)
    assert torch.equal(a, b)


def test_native_error_pins():
    n = 4096
    z = np.zeros((1, 2, n), np.uint64)
    p1 = (E.BfvParamsBuilder().ring_degree(n).plain_modulus(1040407).ct_moduli([18014398509506561])
          .aux_moduli([36028797018972161]).gadget_base(256).build())
    with pytest.raises(E.ExactoError, match="single aux prime too small") as e:
        E.bfv_mul_and_relin_batch(p1, z, z, E.RelinKey(np.zeros((p1.gadget_digits, 2, n), np.uint64), p1))
    assert e.value.kind == "InvalidParam"
    p0 = (E.BfvParamsBuilder().ring_degree(n).plain_modulus(1040407).ct_moduli([18014398509506561])
          .gadget_base(256).build())
    with pytest.raises(E.ExactoError, match="schoolbook BFV multiplication can overflow i128") as e:
        E.bfv_mul_and_relin_batch(p0, z, z, E.RelinKey(np.zeros((p0.gadget_digits, 2, n), np.uint64), p0))
    assert e.value.kind == "NotImplemented"
    small = E.small_bfv()     # BASELINE config 3 as literally written: no aux basis -> the reference errors too
    with pytest.raises(E.ExactoError, match="schoolbook BFV multiplication can overflow i128"):
        E.bfv_mul_and_relin_batch(small, z, z, E.RelinKey(np.zeros((small.gadget_digits, 2, n), np.uint64), small))
    p3 = (E.BfvParamsBuilder().ring_degree(n).plain_modulus(257).ct_moduli([1152921504606830593])
          .aux_moduli([18014398509998081, 36028797018972161, 576460752308273153]).build())
    with pytest.raises(E.ExactoError, match="HPS scaling supports 1 or 2 aux primes, got 3"):
        E.bfv_mul_and_relin_batch(p3, z, z, E.RelinKey(np.zeros((p3.gadget_digits, 2, n), np.uint64), p3))


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_both_aux_basis_paths():
    """cfg 3' and cfg 4 through both tensor paths: the internal 27-bit auxiliary basis (default when it is
    provably result-identical) and the reference's own aux primes (EXB_CTX_REFERENCE_AUX_BASIS)."""
    g = golden()
    for name in ("cfg3p_dbfv", "u64_dbfv"):
        P, base, d, pm, seed, _ = CASES[name]
        ct1, ct2, rlk_arr = golden_inputs(P, d, seed)
        for mode in ("reference", "internal"):
            params = to_dbfv_params(P, base, d, pm)            # fresh params -> fresh native context
            params.bfv_params.context_flags = 1 if mode == "reference" else 0
            rlk = E.RelinKey(rlk_arr, params.bfv_params)
            out = E.dbfv_mul_batch(params, ct1[None], ct2[None], rlk)
            assert digest(out[0]) == str(g[f"{name}/dbfv_sha256"]), (name, mode)


def test_eval_poly_homomorphic_paterson_stockmeyer():
    """bootstrap/digit_extract.rs:100-157 as a sequence of hot-path multiplications;
    single-ciphertext API and device-resident batch, bit-exact vs the oracle restatement + decrypt KAT."""
    from exacto_b200 import batch
    P = O.OracleParams(n=4096, q=1152921504606830593, aux=(18014398509998081, 36028797018972161),
                       plain_modulus=17, gadget_base=256)       # p = 17: noise budget for depth 3
    params = to_params(P)
    rng = np.random.default_rng(7)
    s = H.gen_secret_key(P, rng)
    rlk_arr = H.gen_relin_key(P, s, rng)
    rlk = E.RelinKey(rlk_arr, params)
    coeffs = [1, 2, 3, 1, 0, 7]                                   # degree 5 -> k = 3, two baby + one giant step
    xs = [5, 11, 16]
    cts = np.stack([H.encrypt_sk(P, H.encode_scalar(P, x), s, rng) for x in xs])
    want = np.stack([H.eval_poly_homomorphic(P, ct, coeffs, rlk_arr) for ct in cts])
    one = E.eval_poly_homomorphic(E.BfvCiphertext.from_array(cts[0], params), coeffs, rlk)
    assert np.array_equal(one.to_array(), want[0])
    got = batch.to_host(E.eval_poly_homomorphic_batch(params, batch.to_device(cts), coeffs, rlk))
    assert np.array_equal(got, want)
    for x, ct in zip(xs, got):
        assert int(H.decrypt(P, ct, s)[0]) == sum(c * x ** i for i, c in enumerate(coeffs)) % 17
    assert np.array_equal(E.trivial_encrypt(100, params).to_array(), H.trivial_encrypt(P, 100))
    assert np.array_equal(E.eval_poly_homomorphic(E.BfvCiphertext.from_array(cts[0], params), [42], rlk).to_array(),
                          H.trivial_encrypt(P, 42))


@pytest.mark.parametrize("preset", ["compact_dbfv", "u64_dbfv"])
def test_cpp_host_mirror(preset, tmp_path):
    """The C++ host mirror (include/exacto_b200.hpp: same names / guards as the reference) end to end:
    dbfv_mul, bfv_mul_and_relin, dbfv_add, NTT round trip, dbfv_apply_automorphism, bfv_trace and the error pins,
    bit-compared with the oracle."""
    import subprocess
    import __graft_entry__ as g
    exe = g.build_cpp_driver()
    S = getattr(H, preset)()
    P = S.bfv
    rng = np.random.default_rng(31)
    ct1 = rng.integers(0, P.q, (S.d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (S.d, 2, P.n), dtype=np.uint64)
    rlk = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([ct1.ravel(), ct2.ravel(), rlk.ravel()]).tofile(fin)
    res = subprocess.run([exe, preset, str(fin), str(fout)], capture_output=True, text=True)
    assert res.returncode == 0 and "guards ok" in res.stdout, res.stdout + res.stderr
    out = np.fromfile(fout, dtype=np.uint64)
    lim = S.d * 2 * P.n
    assert np.array_equal(out[:lim].reshape(S.d, 2, P.n),
                          O.dbfv_mul(P, S.base, S.d, S.plain_modulus, ct1, ct2, rlk, threads=O.max_threads()))
    assert np.array_equal(out[lim:lim + 2 * P.n].reshape(2, P.n), O.bfv_mul_and_relin(P, ct1[0], ct2[0], rlk))
    want_sum = np.array((ct1.astype(object) + ct2.astype(object)) % P.q, dtype=np.uint64)
    assert np.array_equal(out[lim + 2 * P.n:2 * lim + 2 * P.n].reshape(S.d, 2, P.n), want_sum)
    off = 2 * lim + 2 * P.n
    assert np.array_equal(out[off:off + P.n], O.ntt_inv(ct1[0, 0], P.q))
    rot = O.bfv_apply_automorphism(P, ct1, rlk, 3, threads=O.max_threads())
    assert np.array_equal(out[off + P.n:off + P.n + lim].reshape(S.d, 2, P.n), rot)
    assert np.array_equal(out[off + P.n + lim:off + P.n + lim + 2 * P.n].reshape(2, P.n), O.bfv_add(P, ct1[0], rot[0]))
    assert np.array_equal(out[off + P.n + lim + 2 * P.n:], H.decrypt(P, ct1[0], rlk[0, 1]))


def test_device_api_chunk_loop():
    """exb_dbfv_mul splits large batches into workspace-bounded chunks; force tiny chunks
    (exb_context_set_option "device_chunk_bytes") and compare with the oracle; same for the host pipeline."""
    from exacto_b200 import batch
    P, base, d, pm, seed, _ = CASES["n64_a2_rep"]
    rng = np.random.default_rng(3)
    ct1 = rng.integers(0, P.q, (7, d, 2, P.n), dtype=np.uint64); ct2 = rng.integers(0, P.q, (7, d, 2, P.n), dtype=np.uint64)
    rk = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    params = to_dbfv_params(P, base, d, pm)
    ctx = params.bfv_params.context()
    ctx.set_option("device_chunk_bytes", 2 * 40000)                 # ~2 pairs per chunk at n=64
    ctx.set_option("host_chunk_products", 2 * d * d)                # 2 pairs per host chunk
    rlk = E.RelinKey(rk, params.bfv_params)
    want = np.stack([O.dbfv_mul(P, base, d, pm, a, b, rk) for a, b in zip(ct1, ct2)])
    got = batch.to_host(batch.dbfv_mul(params, batch.to_device(ct1), batch.to_device(ct2), rlk))
    assert np.array_equal(got, want)
    assert np.array_equal(E.dbfv_mul_batch(params, ct1, ct2, rlk), want)
    with pytest.raises(E.ExactoError):
        ctx.set_option("no_such_option", 1)


@pytest.mark.parametrize("n", [16, 2048, 8192])
def test_generic_ring_degrees(n):
    """Ring degrees other than 4096 take the generic shared-memory path (any n <= 8192)."""
    # n = 8192 needs aux primes = 1 mod 16384 (the u64-profile pair is only = 1 mod 8192)
    aux = (36028797019389953, 36028797019488257) if n == 8192 else (18014398509998081, 36028797018972161)
    P = O.OracleParams(n=n, q=1152921504606830593, aux=aux, plain_modulus=1040407, gadget_base=256)
    rng = np.random.default_rng(n)
    ct1 = rng.integers(0, P.q, (2, 2, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (2, 2, 2, n), dtype=np.uint64)
    rlk_arr = rng.integers(0, P.q, (P.gadget_digits, 2, n), dtype=np.uint64)
    params = to_dbfv_params(P, 256, 2, 65536)
    got = E.dbfv_mul_batch(params, ct1, ct2, E.RelinKey(rlk_arr, params.bfv_params))
    want = np.stack([O.dbfv_mul(P, 256, 2, 65536, a, b, rlk_arr, threads=O.max_threads()) for a, b in zip(ct1, ct2)])
    assert np.array_equal(got, want)


# ---- Galois automorphism + key switch ------------------------------------------------------
def test_automorphism_kats():
    """bfv/eval.rs:929-976 (sigma_3 keeps the scalar 10; 1 + 2X -> 1 + 2X^3) and
    dbfv/advanced.rs:182-193 (dBFV value 42 through sigma_3), with valid keys."""
    P, params = H.compact_bfv(), E.compact_bfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(P, rng)
    gk = E.GaloisKey(H.gen_galois_key(P, s, 3, rng), 3, params)
    ct = E.BfvCiphertext.from_array(H.encrypt_sk(P, H.encode_scalar(P, 10), s, rng), params)
    assert int(H.decrypt(P, E.bfv_apply_automorphism(ct, gk).to_array(), s)[0]) == 10
    pt = np.zeros(P.n, np.uint64); pt[:2] = [1, 2]
    ct = E.BfvCiphertext.from_array(H.encrypt_sk(P, pt, s, rng), params)
    dec = H.decrypt(P, E.bfv_apply_automorphism(ct, gk).to_array(), s)
    assert dec[:4].tolist() == [1, 0, 0, 2] and np.count_nonzero(dec) == 2
    # trace over {3}: m + sigma_3(m) = 2 + 2X + 2X^3 (bfv/eval.rs:573-588)
    dec = H.decrypt(P, E.bfv_trace(ct, [3], {3: gk}).to_array(), s)
    assert dec[:4].tolist() == [2, 2, 0, 2]
    with pytest.raises(E.ExactoError, match="missing Galois key for element 5"):
        E.bfv_trace(ct, [5], {3: gk})
    with pytest.raises(E.ExactoError, match="automorphism requires degree-1 ciphertext"):
        E.bfv_apply_automorphism(E.BfvCiphertext(ct.c + ct.c[:1], params), gk)
    with pytest.raises(E.ExactoError, match="Galois element must be odd"):
        E.bfv_apply_automorphism(ct, E.GaloisKey(gk.array, 4, params))
    # inner product 3*ct_a + 5*ct_b (bfv/eval.rs:593-606)
    enc = lambda m: E.BfvCiphertext.from_array(H.encrypt_sk(P, H.encode_scalar(P, m), s, rng), params)
    const = lambda v: E.CoeffPoly.from_coeffs(np.array([v] + [0] * (P.n - 1)), P.plain_modulus)
    assert int(H.decrypt(P, E.bfv_inner_product([enc(2), enc(7)], [const(3), const(5)]).to_array(), s)[0]) == 41
    with pytest.raises(E.ExactoError, match="mismatched ct/pt lengths"):
        E.bfv_inner_product([enc(2)], [])

    S, dparams = H.compact_dbfv(), E.compact_dbfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(S.bfv, rng)
    gk = E.GaloisKey(H.gen_galois_key(S.bfv, s, 3, rng), 3, dparams.bfv_params)
    dct = E.DbfvCiphertext.from_array(H.dbfv_encrypt_sk(S, 42, s, rng), dparams)
    auto = E.dbfv_apply_automorphism(dct, gk)
    assert H.dbfv_decrypt(S, auto.to_array(), s) == 42 and (auto.degree, auto.mul_depth) == (dct.degree, dct.mul_depth)


@pytest.mark.parametrize("preset,elements", [("compact", [3, 5, 2047]), ("u64", [3, 4097, 8191]), ("cfg3", [5]),
                                             ("n16", [3, 31]), ("n8192", [3, 16383])])
def test_automorphism_vs_oracle(preset, elements):
    """Word-for-word parity of exb_bfv_apply_automorphism (host and device entry points) with the
    literal restatement of bfv/eval.rs:512-561 on uniform, edge and zero ciphertexts."""
    from exacto_b200 import batch
    P = {"compact": H.compact_bfv(), "u64": H.u64_dbfv().bfv, "cfg3": H.cfg3_prime().bfv,
         "n16": O.OracleParams(n=16, q=1152921504606830593, aux=(18014398509998081,), plain_modulus=17, gadget_base=10),
         "n8192": O.OracleParams(n=8192, q=1152921504606830593, aux=(36028797019389953, 36028797019488257),
                                 plain_modulus=1040407, gadget_base=256)}[preset]
    params = to_params(P)
    q, n = P.q, P.n
    rng = np.random.default_rng(n + 1)
    B = 9
    ct = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    edge = np.zeros(n, np.uint64)
    edge[:6] = [0, 1, q - 1, q // 2, q // 2 + 1, 2]
    ct[1, 0] = O.ntt_fwd(edge, q); ct[1, 1] = O.ntt_fwd(edge[::-1].copy(), q)
    ct[2] = 0
    ct[3] = O.ntt_fwd(np.full((2, n), q - 1, np.uint64), q)
    ct[4] = O.ntt_fwd(np.full((2, n), q // 2 + 1, np.uint64), q)
    karr = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    for k in elements:
        gk = E.GaloisKey(karr, k, params)
        want = O.bfv_apply_automorphism(P, ct, karr, k, threads=O.max_threads())
        assert np.array_equal(E.bfv_apply_automorphism_batch(params, ct, gk), want), (preset, k, "host")
        dev = batch.bfv_apply_automorphism(params, batch.to_device(ct), gk)
        assert np.array_equal(batch.to_host(dev), want), (preset, k, "device")
    # sigma_a o sigma_b = sigma_ab on ciphertext components when the key switch is the identity gadget:
    # with an all-zero key only c0' = NTT(sigma(INTT c0)) survives -> compose and compare
    zero = E.GaloisKey(np.zeros_like(karr), 3, params)
    a = E.bfv_apply_automorphism_batch(params, ct, zero)
    ab = E.bfv_apply_automorphism_batch(params, a, E.GaloisKey(np.zeros_like(karr), 5, params))
    direct = E.bfv_apply_automorphism_batch(params, ct, E.GaloisKey(np.zeros_like(karr), 15, params))
    assert np.array_equal(ab[:, 0], direct[:, 0]) and not ab[:, 1].any()
    with pytest.raises(E.ExactoError, match="must not alias"):
        d = batch.to_device(ct)
        batch.bfv_apply_automorphism(params, d, zero, out=d)


# ---- keygen / encrypt / decrypt around the path ---------------------------------------------
class NpSampler:
    """The harness samplers behind exacto_b200's sampler protocol (same rng consumption order as the oracle)."""

    def __init__(self, rng):
        self.rng = rng

    def ternary(self, n, q): return H.sample_ternary(n, q, self.rng)
    def uniform(self, n, q): return H.sample_uniform(n, q, self.rng)
    def gaussian(self, n, q, sigma): return H.sample_gaussian(n, q, sigma, self.rng)


@pytest.mark.parametrize("preset", ["compact", "u64"])
def test_keygen_encrypt_decrypt_parity(preset):
    """Keys and ciphertexts are bit-exact functions of the sampled polynomials (bfv/keygen.rs:64-210,
    bfv/encrypt.rs:79-106); decrypt (bfv/encrypt.rs:111-178) matches the big-int restatement for degree-1 and
    degree-2 ciphertexts, host and device entry points; dBFV encrypt/decrypt round trip (dbfv/decrypt.rs tests)."""
    from exacto_b200 import batch
    S = H.compact_dbfv() if preset == "compact" else H.u64_dbfv()
    P = S.bfv
    dparams = to_dbfv_params(P, S.base, S.d, S.plain_modulus)
    params = dparams.bfv_params
    r1, r2 = np.random.default_rng(9), np.random.default_rng(9)
    s_ref = H.gen_secret_key(P, r1)
    rlk_ref = H.gen_relin_key(P, s_ref, r1)
    gk_ref = H.gen_galois_key(P, s_ref, 5, r1)
    smp = NpSampler(r2)
    sk = E.gen_secret_key_with_sampler(params, smp)
    rlk = E.gen_relin_key_with_sampler(sk, smp)
    gk = E.gen_galois_key_with_sampler(sk, 5, smp)
    assert np.array_equal(sk.ntt_array(), s_ref)
    assert np.array_equal(rlk.array, rlk_ref) and np.array_equal(gk.array, gk_ref) and gk.element == 5
    pt = r1.integers(0, P.plain_modulus, P.n, dtype=np.uint64); r2.integers(0, P.plain_modulus, P.n, dtype=np.uint64)
    ct_ref = H.encrypt_sk(P, pt, s_ref, r1)
    a = E.CoeffPoly(smp.uniform(P.n, P.q), P.q); e = E.CoeffPoly(smp.gaussian(P.n, P.q, 3.2), P.q)
    ct = E.encrypt_sk_with_samples(E.CoeffPoly(pt, P.plain_modulus), sk, params, a, e)
    assert np.array_equal(ct.to_array(), ct_ref)
    assert np.array_equal(E.decrypt(ct, sk).coeffs, pt)
    prod3 = O.bfv_mul_no_relin(P, ct_ref, ct_ref)                                  # degree 2
    rnd = r1.integers(0, P.q, (3, 3, P.n), dtype=np.uint64)
    rnd[0] = prod3
    want = np.stack([H.decrypt(P, c, s_ref) for c in rnd])
    assert np.array_equal(E.decrypt_batch(params, rnd, sk), want)
    dev = batch.bfv_decrypt(params, batch.to_device(rnd), batch.to_device(s_ref))
    assert np.array_equal(batch.to_host(dev), want)
    for value in [0, 1, 42, 255]:                                                   # dbfv/decrypt.rs:96-108
        smp2 = [(E.CoeffPoly(smp.uniform(P.n, P.q), P.q), E.CoeffPoly(smp.gaussian(P.n, P.q, 3.2), P.q)) for _ in range(S.d)]
        dct = E.dbfv_encrypt_sk_with_samples(value, sk, dparams, [x[0] for x in smp2], [x[1] for x in smp2])
        assert E.dbfv_decrypt(dct, sk) == value
    with pytest.raises(E.ExactoError, match="plaintext 1000000000 >= plain_modulus"):
        E.encode_scalar(1000000000, params)


NO_AUX_GPU = {
    "boot_orig": (O.OracleParams(n=16, q=65537, aux=(), plain_modulus=5), None),
    "boot_scheme": (O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=29, gadget_base=8), None),
    "boot_dbfv": (O.OracleParams(n=16, q=65537, aux=(), plain_modulus=97, gadget_base=8), (4, 2, 16)),
    "n1024_q40": (O.OracleParams(n=1024, q=1099509805057, aux=(), plain_modulus=257), (16, 2, 256)),
    "n4096_q50": (O.OracleParams(n=4096, q=1125899906826241, aux=(), plain_modulus=257, gadget_base=256), None),
}


@pytest.mark.parametrize("name", list(NO_AUX_GPU))
def test_no_aux_params_match_schoolbook(name):
    """Parameter sets without an auxiliary basis: the reference's exact i128 schoolbook branch
    (bfv/eval.rs:415-464) vs the device's HPS pipeline on an internal auxiliary basis -- word for word."""
    P, dbfv = NO_AUX_GPU[name]
    params = to_params(P)
    q, n = P.q, P.n
    rng = np.random.default_rng(len(name))
    rlk_arr = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, params)
    ct1 = rng.integers(0, q, (3, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, q, (3, 2, n), dtype=np.uint64)
    edge = np.full(n, q // 2, np.uint64); edge[::2] = q // 2 + 1
    ct1[0] = O.ntt_fwd(np.stack([edge, edge[::-1].copy()]), q)
    ct2[0] = O.ntt_fwd(np.stack([edge, edge]), q)
    assert np.array_equal(E.bfv_mul_and_relin_batch(params, ct1, ct2, rlk),
                          O.bfv_mul_and_relin(P, ct1, ct2, rlk_arr, threads=O.max_threads()))
    if dbfv:
        b, d, pm = dbfv
        dp = to_dbfv_params(P, b, d, pm)
        a = rng.integers(0, q, (2, d, 2, n), dtype=np.uint64); c = rng.integers(0, q, (2, d, 2, n), dtype=np.uint64)
        want = np.stack([O.dbfv_mul(P, b, d, pm, x, y, rlk_arr, threads=O.max_threads()) for x, y in zip(a, c)])
        assert np.array_equal(E.dbfv_mul_batch(dp, a, c, E.RelinKey(rlk_arr, dp.bfv_params)), want)
    with pytest.raises(E.ExactoError, match="schoolbook BFV multiplication can overflow i128"):   # README config 3
        p3 = to_params(O.OracleParams(n=4096, q=576460752308273153, aux=(), plain_modulus=65537))
        z = np.zeros((1, 2, 4096), np.uint64)
        E.bfv_mul_and_relin_batch(p3, z, z, E.RelinKey(np.zeros((4, 2, 4096), np.uint64), p3))


# ---- bootstrap (bootstrap/bfv_host.rs:134-288, tests :345-560): SURVEY 8(f)1, BASELINE config 5 ----------------
def _product_bootstrap_key(bk, boot_params):
    return E.BootstrapKey(boot_params, E.RelinKey(bk.boot_rlk, boot_params),
                          bsk=E.BfvCiphertext.from_array(bk.bsk, boot_params),
                          galois_keys={k: E.GaloisKey(v, k, boot_params) for k, v in bk.galois_keys.items()},
                          rounding_poly=bk.rounding_poly, t_orig=bk.t_orig, q_prime=bk.q_prime)


def test_bootstrap_single_and_ring():
    """bfv_host.rs:398-453 on the reference's toy parameter sets: trivial ciphertexts m = 0..4 decode to m after
    the refresh; a real encryption takes the full ring path (CoeffsToSlots, rounding polynomial per slot,
    SlotsToCoeffs) and must equal the oracle pipeline word for word; on a parameter set that fits the noise
    budget the ring path must also equal the clear-text model of the bootstrap for all n coefficients."""
    from oracle import bootstrap_ref as B
    from test_oracle import _boot_params, expected_ring_bootstrap
    orig, boot, qp = _boot_params()
    op, bp = to_params(orig), to_params(boot)
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(orig, rng)
    bk = B.gen_bootstrap_key(orig, boot, s, qp, orig.plain_modulus, rng)
    pk = _product_bootstrap_key(bk, bp)
    sk = E.SecretKey.from_ntt(s, op)
    boot_sk = E.create_boot_sk(sk, bp)
    assert np.array_equal(boot_sk.ntt_array(), B.create_boot_sk(orig, boot, s))
    assert E.compute_rounding_poly(5, 25, 29) == bk.rounding_poly
    for m in range(5):
        out = E.bfv_bootstrap(E.trivial_encrypt(m, op), pk)
        assert out.params is bp and E.decode_scalar(E.decrypt(out, boot_sk)) % 5 == m
        assert np.array_equal(out.to_array(), B.bfv_bootstrap(orig, H.trivial_encrypt(orig, m), bk))
    ct = H.encrypt_sk(orig, H.encode_scalar(orig, 3), s, rng)
    got = E.bfv_bootstrap(E.BfvCiphertext.from_array(ct, op), pk)
    assert np.array_equal(got.to_array(), B.bfv_bootstrap(orig, ct, bk))
    with pytest.raises(E.ExactoError, match="bootstrap requires degree-1 ciphertext"):
        E.bfv_bootstrap(E.BfvCiphertext.from_array(np.zeros((3, 16), np.uint64), op), pk)
    # coefficient extraction / packing on their own (coeffs_to_slots.rs tests)
    pt = rng.integers(0, boot.plain_modulus, boot.n, dtype=np.uint64)
    cte = E.BfvCiphertext.from_array(H.encrypt_sk(boot, pt, boot_sk.ntt_array(), rng), bp)
    slots = E.coeffs_to_slots(cte, pk.galois_keys)
    assert [E.decode_scalar(E.decrypt(sl, boot_sk)) for sl in slots] == [int(v) for v in pt]
    assert np.array_equal(E.decrypt(E.slots_to_coeffs(slots), boot_sk).coeffs, pt)
    assert np.array_equal(slots[5].to_array(), B.extract_coefficient(boot, cte.to_array(), 5, bk.galois_keys))

    orig2 = O.OracleParams(n=16, q=65537, aux=(), plain_modulus=2)
    boot2 = O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=5, gadget_base=8)
    op2, bp2 = to_params(orig2), to_params(boot2)
    r1, r2 = np.random.default_rng(3), np.random.default_rng(3)
    s2 = H.gen_secret_key(orig2, r1)
    sk2 = E.gen_secret_key_with_sampler(op2, NpSampler(r2))
    bk2 = B.gen_bootstrap_key(orig2, boot2, s2, 4, 2, r1)
    pk2 = E.gen_bootstrap_key_with_sampler(sk2, bp2, 4, 2, NpSampler(r2))          # same stream -> same key material
    assert np.array_equal(pk2.bsk.to_array(), bk2.bsk) and np.array_equal(pk2.boot_rlk.array, bk2.boot_rlk)
    assert sorted(pk2.galois_keys) == sorted(bk2.galois_keys)
    assert all(np.array_equal(pk2.galois_keys[k].array, bk2.galois_keys[k]) for k in bk2.galois_keys)
    ct2 = H.encrypt_sk(orig2, r1.integers(0, 2, 16, dtype=np.uint64), s2, r1)
    out2 = E.bfv_bootstrap(E.BfvCiphertext.from_array(ct2, op2), pk2)
    assert np.array_equal(out2.to_array(), B.bfv_bootstrap(orig2, ct2, bk2))
    assert np.array_equal(E.decrypt(out2, E.create_boot_sk(sk2, bp2)).coeffs, expected_ring_bootstrap(orig2, boot2, 4, ct2, s2))


def test_dbfv_mul_then_bootstrap_and_chain():
    """bfv_host.rs:455-560: dbfv_mul_then_bootstrap resets mul_depth and swaps in the boot parameter set, the
    refreshed ciphertext multiplies again under the boot key material, the chain helper folds three inputs;
    every ciphertext equals the oracle pipeline's word for word."""
    from oracle import bootstrap_ref as B
    from test_oracle import _dbfv_boot_params
    S, boot, qp = _dbfv_boot_params()
    dp = to_dbfv_params(S.bfv, S.base, S.d, S.plain_modulus)
    bp = to_params(boot)
    rng = np.random.default_rng(777)
    s = H.gen_secret_key(S.bfv, rng)
    rlk_arr = H.gen_relin_key(S.bfv, s, rng)
    bk = B.gen_bootstrap_key(S.bfv, boot, s, qp, S.bfv.plain_modulus, rng)
    pk = _product_bootstrap_key(bk, bp)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    boot_sk = E.create_boot_sk(E.SecretKey.from_ntt(s, dp.bfv_params), bp)
    pa = np.zeros(16, np.uint64); pa[:2] = [3, 1]
    pb = np.zeros(16, np.uint64); pb[0] = 2
    ca, cb = H.dbfv_encrypt_poly_sk(S, pa, s, rng), H.dbfv_encrypt_poly_sk(S, pb, s, rng)
    refreshed = E.dbfv_mul_then_bootstrap(E.DbfvCiphertext.from_array(ca, dp), E.DbfvCiphertext.from_array(cb, dp), rlk, pk)
    S2, want = B.dbfv_mul_then_bootstrap(S, ca, cb, rlk_arr, bk)
    assert refreshed.mul_depth == 0 and refreshed.params.bfv_params.plain_modulus == 257 and refreshed.degree == 2
    assert np.array_equal(refreshed.to_array(), want)
    c3 = H.dbfv_encrypt_poly_sk(S2, pb, boot_sk.ntt_array(), rng)
    nxt = E.dbfv_mul(refreshed, E.DbfvCiphertext.from_array(c3, refreshed.params), pk.boot_rlk)
    assert nxt.mul_depth == 1
    assert np.array_equal(nxt.to_array(), O.dbfv_mul(S2.bfv, S2.base, S2.d, S2.plain_modulus, want, c3, bk.boot_rlk))
    dec = E.dbfv_decrypt_poly(nxt, boot_sk)
    assert dec.modulus == 16 and len(dec) == 16 and isinstance(E.dbfv_decrypt(nxt, boot_sk), int)
    mk = lambda m: H.dbfv_encrypt_poly_sk(S, np.array([m % 16] + [0] * 15, np.uint64), s, rng)
    arrs = [mk(3), mk(2), mk(5)]
    chained = E.dbfv_mul_chain_then_bootstrap([E.DbfvCiphertext.from_array(a, dp) for a in arrs], rlk, pk)
    S3, want_chain = B.dbfv_mul_chain_then_bootstrap([(S, a) for a in arrs], rlk_arr, bk)
    assert chained.mul_depth == 0 and chained.params.bfv_params.plain_modulus == 257
    assert np.array_equal(chained.to_array(), want_chain)
    assert len(E.dbfv_decrypt_poly(chained, boot_sk)) == 16


def test_cpp_dbfv_mul_then_bootstrap_and_chain(tmp_path):
    """north_star's dbfv_mul_then_bootstrap / dbfv_mul_chain_then_bootstrap at the COMPILED boundary
    (include/exacto_b200.hpp, bootstrap/bfv_host.rs:242-288 with the rlk selection of :271-284): the C++ host mirror
    drives the GPU through the C ABI on the reference's toy parameter sets; every ciphertext equals the oracle
    pipeline's word for word."""
    import subprocess
    import __graft_entry__ as g
    from oracle import bootstrap_ref as B
    from test_oracle import _dbfv_boot_params
    exe = g.build_cpp_driver(name="bootstrap_driver")
    S, boot, qp = _dbfv_boot_params()
    orig = S.bfv
    rng = np.random.default_rng(777)
    s = H.gen_secret_key(orig, rng)
    rlk_arr = H.gen_relin_key(orig, s, rng)
    bk = B.gen_bootstrap_key(orig, boot, s, qp, orig.plain_modulus, rng)
    mk = lambda m: H.dbfv_encrypt_poly_sk(S, np.array([m % 16] + [0] * 15, np.uint64), s, rng)
    cts = [mk(3), mk(2), mk(5)]
    words = [orig.n, orig.q, orig.plain_modulus, orig.gadget_base, boot.q, boot.plain_modulus, boot.gadget_base,
             qp, S.base, S.d, S.plain_modulus, len(bk.rounding_poly), len(bk.galois_keys), len(cts)]
    parts = [np.array(words, np.uint64), np.array(bk.rounding_poly, np.uint64), rlk_arr.ravel(), bk.bsk.ravel(), bk.boot_rlk.ravel()]
    for k in sorted(bk.galois_keys):
        parts += [np.array([k], np.uint64), bk.galois_keys[k].ravel()]
    parts += [c.ravel() for c in cts]
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate(parts).astype(np.uint64).tofile(fin)
    res = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True)
    assert res.returncode == 0 and "metadata ok" in res.stdout, res.stdout + res.stderr
    out = np.fromfile(fout, dtype=np.uint64).reshape(3, S.d, 2, orig.n)
    S2, want = B.dbfv_mul_then_bootstrap(S, cts[0], cts[1], rlk_arr, bk)
    assert np.array_equal(out[0], want)
    _, want_chain = B.dbfv_mul_chain_then_bootstrap([(S, a) for a in cts], rlk_arr, bk)
    assert np.array_equal(out[1], want_chain)
    _, third = B.dbfv_bootstrap(S, cts[2], bk)
    assert np.array_equal(out[2], O.dbfv_mul(S2.bfv, S2.base, S2.d, S2.plain_modulus, want, third, bk.boot_rlk))


def test_shared_context_from_host_threads_and_streams():
    """The reference's functions are re-entrant (no global state).  One context shared by host threads
    (ctypes drops the GIL) and by device-resident calls on different CUDA streams must still give the oracle's
    words: entry points serialise on the context and a workspace slot waits for its previous stream."""
    import threading
    from exacto_b200 import batch
    S = H.cfg3_prime()
    P = S.bfv
    dp = to_dbfv_params(P, S.base, S.d, S.plain_modulus)
    rng = np.random.default_rng(99)
    rlk_arr = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    ins = [(rng.integers(0, P.q, (3, S.d, 2, P.n), dtype=np.uint64), rng.integers(0, P.q, (3, S.d, 2, P.n), dtype=np.uint64))
           for _ in range(4)]
    want = [np.stack([O.dbfv_mul(P, S.base, S.d, S.plain_modulus, x, y, rlk_arr, threads=O.max_threads()) for x, y in zip(a, b)])
            for a, b in ins]
    rlk.native(dp.bfv_params.context())                    # upload once before the threads start
    got, errs = [None] * 4, []

    def work(i):
        try:
            for _ in range(3):
                got[i] = E.dbfv_mul_batch(dp, ins[i][0], ins[i][1], rlk)
        except Exception as exc:                           # pragma: no cover
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs, errs
    assert all(np.array_equal(g, w) for g, w in zip(got, want))
    streams = [torch.cuda.Stream() for _ in range(2)]
    outs = []
    for i, st in enumerate(streams * 2):
        with torch.cuda.stream(st):
            a, b = batch.to_device(ins[i][0]), batch.to_device(ins[i][1])
            outs.append(batch.dbfv_mul(dp, a, b, rlk))
    torch.cuda.synchronize()
    assert all(np.array_equal(batch.to_host(o), w) for o, w in zip(outs, want))


def test_async_host_api_pinned_memory_and_slot_pool():
    """exb_*_host_async / exb_wait: several calls in flight on page-locked buffers (exb_host_alloc), waited out of
    order; a ticket is waited once.  Device-resident calls from more host threads and streams than the context has
    workspace slots (slot reuse is ordered by the slot's own event).  exb_host_register pins a numpy array in place."""
    import ctypes, threading
    from exacto_b200 import _native, batch, hostmem
    S = H.cfg3_prime()
    P = S.bfv
    dp = to_dbfv_params(P, S.base, S.d, S.plain_modulus)
    ctx = dp.bfv_params.context()
    rng = np.random.default_rng(7)
    rlk_arr = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    shape = (5, S.d, 2, P.n)
    calls = []
    for i in range(6):
        a, b, o = (hostmem.pinned_empty(dp, shape) for _ in range(3))
        a.array[:] = rng.integers(0, P.q, shape, dtype=np.uint64)
        b.array[:] = rng.integers(0, P.q, shape, dtype=np.uint64)
        calls.append((a, b, o))
    want = [np.stack([O.dbfv_mul(P, S.base, S.d, S.plain_modulus, x, y, rlk_arr, threads=O.max_threads())
                      for x, y in zip(a.array, b.array)]) for a, b, _ in calls]
    ctx.set_option("host_chunk_products", 2 * S.d * S.d)             # several chunks per call
    pend = [hostmem.dbfv_mul_batch_async(dp, a.array, b.array, rlk, o.array) for a, b, o in calls]
    for i in reversed(range(6)):
        assert np.array_equal(pend[i].wait(), want[i]), i
    L = _native.lib()
    t = ctypes.c_uint64()
    a, b, o = calls[0]
    _native.check(L.exb_dbfv_mul_host_async(ctx.handle, dp.base, S.d, dp.plain_modulus, a.array.ctypes.data,
                                            b.array.ctypes.data, rlk.native(ctx), o.array.ctypes.data, 5, 0, ctypes.byref(t)))
    assert t.value != 0
    _native.check(L.exb_wait(ctx.handle, t.value))
    assert L.exb_wait(ctx.handle, t.value) == 1                      # EXB_INVALID_PARAM: already waited
    assert b"ticket" in L.exb_last_error()
    # pin an ordinary numpy array in place
    x = np.ascontiguousarray(rng.integers(0, P.q, shape, dtype=np.uint64))
    _native.check(L.exb_host_register(ctx.handle, x.ctypes.data, x.nbytes))
    got = E.dbfv_mul_batch(dp, x, calls[1][1].array, rlk)
    _native.check(L.exb_host_unregister(ctx.handle, x.ctypes.data))
    assert np.array_equal(got[0], O.dbfv_mul(P, S.base, S.d, S.plain_modulus, x[0], calls[1][1].array[0], rlk_arr))
    # 6 host threads x own stream x 3 calls each on one context (4 workspace slots)
    dev_in = [(batch.to_device(a.array), batch.to_device(b.array)) for a, b, _ in calls]
    torch.cuda.synchronize()
    res, errs = [None] * 6, []

    def work(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    res[i] = batch.dbfv_mul(dp, dev_in[i][0], dev_in[i][1], rlk)
            st.synchronize()
        except Exception as exc:                           # pragma: no cover
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [t_.start() for t_ in ts]; [t_.join() for t_ in ts]
    assert not errs, errs
    torch.cuda.synchronize()
    assert all(np.array_equal(batch.to_host(r), w) for r, w in zip(res, want))
    # the NTT-domain format id is a function of (n, q, psi, ordering)
    assert ctx.ntt_format_id(0) == to_dbfv_params(P, S.base, S.d, S.plain_modulus).bfv_params.context().ntt_format_id(0)
    assert ctx.ntt_format_id(0) != ctx.ntt_format_id(1)
    assert ctx.ntt_format_id(0) >> 56 == 1


@pytest.mark.parametrize("pairs,kernel_stores", [(2, 0), (30, 0), (2, 1), (30, 1)])
def test_dbfv_mul_scatter_peer_stores(pairs, kernel_stores):
    """exb_dbfv_mul_scatter (k-sharded dbfv_mul): two 'ranks' on one GPU own disjoint output limbs and deliver their
    finished limbs into both output buffers -- with the copy engines (default) or from the relin epilogue (narrow and
    wide relin paths); the union is the oracle's dbfv_mul, and limbs a rank does not own stay untouched in its
    peers' buffers until their owner writes."""
    import ctypes
    from exacto_b200 import _native, batch
    from exacto_b200.sharding import limb_masks
    S = H.u64_dbfv()
    P = S.bfv
    dp = E.u64_dbfv()
    ctx = dp.bfv_params.context()
    L = _native.lib()
    rng = np.random.default_rng(77 + pairs)
    ct1 = rng.integers(0, P.q, (pairs, S.d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (pairs, S.d, 2, P.n), dtype=np.uint64)
    rlk_arr = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    want = np.stack([O.dbfv_mul(P, S.base, S.d, S.plain_modulus, a, b, rlk_arr, threads=O.max_threads())
                     for a, b in zip(ct1[:3], ct2[:3])])
    a, b = batch.to_device(ct1), batch.to_device(ct2)
    masks = limb_masks(S.d, 2)
    outs = [torch.full_like(a, -1), torch.full_like(a, -1)]
    st = torch.cuda.current_stream().cuda_stream
    ctx.set_option("kshard_kernel_stores", kernel_stores)
    for r in range(2):
        peer = (ctypes.c_void_p * 1)(outs[1 - r].data_ptr())
        _native.check(L.exb_dbfv_mul_scatter(ctx.handle, dp.base, S.d, dp.plain_modulus, a.data_ptr(), b.data_ptr(),
                                             rlk.native(ctx), outs[r].data_ptr(), peer, 1, pairs, 0, masks[r], st))
        if r == 0:
            torch.cuda.synchronize()
            h0 = batch.to_host(outs[1])
            for k in range(S.d):
                if (masks[0] >> k) & 1:
                    assert np.array_equal(h0[:3, k], want[:, k])
                else:
                    assert np.all(h0[:, k] == np.uint64(0xFFFFFFFFFFFFFFFF))
    torch.cuda.synchronize()
    ctx.set_option("kshard_kernel_stores", 0)
    assert np.array_equal(batch.to_host(outs[0])[:3], want) and torch.equal(outs[0], outs[1])
    # p != b^d (non-zero small representatives) is refused: the general reduction needs limbs of other ranks
    C = CASES["n64_a2_rep"]
    pr = to_dbfv_params(C[0], C[1], C[2], C[3])
    z = torch.zeros((1, C[2], 2, C[0].n), dtype=torch.int64, device="cuda")
    rk = E.RelinKey(np.zeros((C[0].gadget_digits, 2, C[0].n), np.uint64), pr.bfv_params)
    c2 = pr.bfv_params.context()
    rc = L.exb_dbfv_mul_scatter(c2.handle, pr.base, C[2], pr.plain_modulus, z.data_ptr(), z.data_ptr(), rk.native(c2),
                                z.data_ptr(), None, 0, 1, 0, 1, st)
    assert rc == 9 and b"small representatives" in L.exb_last_error()


def _rns_params(P):
    return E.BfvParamsBuilder().ring_degree(P.n).plain_modulus(P.plain_modulus).ct_moduli(list(P.moduli)).gadget_base(P.gadget_base).build()


@pytest.mark.parametrize("name", ["ref_n16", "n64_two60", "n32_three40_d2"])
def test_multi_prime_ct_modulus(name):
    """Multi-prime ciphertext modulus on the GPU: bfv_mul_generic_rns (bfv/eval.rs:113-147) and the
    L > 1 relinearize, device- and host-buffer entry points, word for word against oracle/rns_ref.py."""
    from exacto_b200 import batch
    from common import RNS_CASES, R, rns_inputs
    P, d, base, pm = RNS_CASES[name]
    bp = _rns_params(P)
    assert bp.gadget_digits == P.G
    pairs, L, n = 3, len(P.moduli), P.n
    ct1, ct2, rlk_arr = rns_inputs(P, d, pairs, 11)
    rlk = E.RelinKey(rlk_arr, bp)
    want = np.stack([R.dbfv_mul(P, d, a, b, rlk_arr) for a, b in zip(ct1, ct2)])
    if d == 1:
        a, b = batch.to_device(ct1[:, 0]), batch.to_device(ct2[:, 0])
        want3 = np.stack([R.bfv_mul_no_relin(P, x[0], y[0]) for x, y in zip(ct1, ct2)])
        got3 = batch.bfv_mul_no_relin(bp, a, b)
        assert np.array_equal(batch.to_host(got3), want3)
        assert np.array_equal(batch.to_host(batch.relinearize(bp, got3, rlk)), want[:, 0])
        assert np.array_equal(batch.to_host(batch.bfv_mul_and_relin(bp, a, b, rlk)), want[:, 0])
        assert np.array_equal(E.bfv_mul_and_relin_batch(bp, ct1[:, 0], ct2[:, 0], rlk), want[:, 0])      # host buffers
        # object API: RnsPoly with one NttPoly per prime
        o = E.bfv_mul_and_relin(E.BfvCiphertext.from_array(ct1[0, 0], bp), E.BfvCiphertext.from_array(ct2[0, 0], bp), rlk)
        assert len(o.c) == 2 and len(o.c[0].components) == L and np.array_equal(o.to_array(), want[0, 0])
        o3 = E.bfv_mul_no_relin(E.BfvCiphertext.from_array(ct1[1, 0], bp), E.BfvCiphertext.from_array(ct2[1, 0], bp))
        assert np.array_equal(E.relinearize(o3, rlk).to_array(), want[1, 0])
    dp = E.DbfvParams.new(bp, base, d, pm)
    got = batch.dbfv_mul(dp, batch.to_device(ct1), batch.to_device(ct2), rlk)
    assert np.array_equal(batch.to_host(got), want)
    assert np.array_equal(E.dbfv_mul_batch(dp, ct1, ct2, rlk), want)
    # ring layer on every ciphertext prime: from_coeff_poly / to_coeff_poly (ring/rns.rs:84-151)
    rng = np.random.default_rng(3)
    coeffs = rng.integers(0, min(P.moduli), n, dtype=np.uint64)
    rp = E.RnsPoly.from_coeff_poly(E.CoeffPoly(coeffs, P.moduli[0]), bp)
    assert all(np.array_equal(c.evals, O.ntt_fwd(coeffs % np.uint64(q), q)) for c, q in zip(rp.components, P.moduli))
    back = rp.to_coeff_poly()
    tc, tm = R.to_coeff_poly_truncated(P, np.stack([c.evals for c in rp.components]))
    assert back.modulus == tm and [int(x) for x in back.coeffs] == tc


def test_multi_prime_reference_kat_and_full_ring_degree():
    """The reference's own multi-prime test (bfv/eval.rs:903-927: 3*7, 10*20, 0*5 at n = 16, Q = 65537 * 1099509805057)
    through the GPU, and one product at n = 4096 with two 60-bit ciphertext primes against the big-integer oracle."""
    from exacto_b200 import batch
    from common import R
    P = R.RnsParams(16, (65537, 1099509805057), 257, 8)
    bp = _rns_params(P)
    rng = np.random.default_rng(1234)
    s = R.gen_secret_key(P, rng)
    rlk_arr = R.gen_relin_key(P, s, rng)
    rlk = E.RelinKey(rlk_arr, bp)
    for a, b, want in [(3, 7, 21), (10, 20, 200), (0, 5, 0)]:
        c1 = R.encrypt_sk(P, [a] + [0] * 15, s, rng)
        c2 = R.encrypt_sk(P, [b] + [0] * 15, s, rng)
        prod = E.bfv_mul_and_relin(E.BfvCiphertext.from_array(c1, bp), E.BfvCiphertext.from_array(c2, bp), rlk)
        assert R.decrypt(P, prod.to_array(), s) == [want] + [0] * 15
        assert np.array_equal(prod.to_array(), R.bfv_mul_and_relin(P, c1, c2, rlk_arr))
        # decrypt on the device too (bfv/encrypt.rs:111-178 with its BigUint CRT), degree 1 and degree 2
        sk = E.SecretKey(E.BfvCiphertext.from_array(s[None], bp).c[0], bp)
        assert [int(x) for x in E.decrypt(prod, sk).coeffs] == [want] + [0] * 15
        prod3 = E.bfv_mul_no_relin(E.BfvCiphertext.from_array(c1, bp), E.BfvCiphertext.from_array(c2, bp))
        assert [int(x) for x in E.decrypt(prod3, sk).coeffs] == [want] + [0] * 15
    P2 = R.RnsParams(4096, (1152921504606830593, 576460752308273153), 65537, 1 << 16)
    bp2 = _rns_params(P2)
    c1 = np.stack([np.stack([rng.integers(0, q, 4096, dtype=np.uint64) for q in P2.moduli]) for _ in range(2)])
    c2 = np.stack([np.stack([rng.integers(0, q, 4096, dtype=np.uint64) for q in P2.moduli]) for _ in range(2)])
    got3 = batch.to_host(batch.bfv_mul_no_relin(bp2, batch.to_device(c1[None]), batch.to_device(c2[None])))[0]
    assert np.array_equal(got3, R.bfv_mul_no_relin(P2, c1, c2))
    s2 = np.stack([rng.integers(0, q, 4096, dtype=np.uint64) for q in P2.moduli])
    sk2 = E.SecretKey(E.BfvCiphertext.from_array(s2[None], bp2).c[0], bp2)
    assert [int(x) for x in E.decrypt_batch(bp2, got3[None], sk2)[0]] == R.decrypt(P2, got3, s2)      # random phase: every digit path
    # refused sets report why (the single-prime pins are in test_native_error_pins)
    big = E.BfvParamsBuilder().ring_degree(32).plain_modulus(257).ct_moduli(
        [1152921504606830593, 576460752308273153, 1099509805057]).build()
    z = np.zeros((1, 2, 3, 32), np.uint64)
    with pytest.raises(E.ExactoError) as ei:
        E.bfv_mul_and_relin_batch(big, z, z, E.RelinKey(np.zeros((big.gadget_digits, 2, 3, 32), np.uint64), big))
    assert "2^126" in str(ei.value)


def test_per_limb_tensor_path_on_gpu():
    """Batches large enough for tensor01_kernel (components 0/1 summed per output limb) vs the oracle, device and
    host entry points, and the same batch through the per-product kernel only (option "tensor_per_product")."""
    from exacto_b200 import batch
    S = H.u64_dbfv()
    P = S.bfv
    dp = E.u64_dbfv()
    rng = np.random.default_rng(2024)
    B = 26                                              # 26 pairs x 4 limb duos x 2 components = 208 CTAs: per-limb kernel
    ct1 = rng.integers(0, P.q, (B, S.d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (B, S.d, 2, P.n), dtype=np.uint64)
    half = O.ntt_fwd(np.full(P.n, P.q // 2, np.uint64), P.q)
    ct1[0] = half; ct2[0] = half                       # every |t_ij| at its bound, equal signs
    rlk_arr = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    want = np.stack([O.dbfv_mul(P, S.base, S.d, S.plain_modulus, a, b, rlk_arr, threads=O.max_threads()) for a, b in zip(ct1, ct2)])
    got = batch.to_host(batch.dbfv_mul(dp, batch.to_device(ct1), batch.to_device(ct2), rlk))
    assert np.array_equal(got, want)
    ctx = dp.bfv_params.context()
    ctx.set_option("tensor_per_product", 1)
    try:
        got = batch.to_host(batch.dbfv_mul(dp, batch.to_device(ct1), batch.to_device(ct2), rlk))
    finally:
        ctx.set_option("tensor_per_product", 0)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("plain", [(1 << 28) - 57, (1 << 29) + 11, (1 << 33) + 7])
def test_per_limb_rounding_sums_large_plain_modulus(plain):
    """tensor01_kernel's sums of rounding terms: p just below 2^28 with 8 products per limb puts the i32 sums at 2^30,
    p above 2^29 / above 2^32 takes the i64 variant (tensor01_kernel<true>); all word-exact at a batch size that selects
    the per-limb kernel."""
    from exacto_b200 import batch
    P = O.OracleParams(n=4096, q=1152921504606830593, aux=(18014398509998081, 36028797018972161), plain_modulus=plain,
                       gadget_base=256)
    d, b = 8, 256
    dp = to_dbfv_params(P, b, d, 0)
    q, n = P.q, P.n
    rng = np.random.default_rng(plain & 0xffff)
    rlk_arr = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    one = np.zeros(n, np.uint64); one[0] = 1
    half = O.ntt_fwd(np.full(n, q // 2, np.uint64), q)
    a = np.stack([np.stack([half, half])] * d)
    bb = np.stack([np.stack([O.ntt_fwd(one, q), O.ntt_fwd(one, q)])] * d)
    bb[1::3] = rng.integers(0, q, (len(bb[1::3]), 2, n), dtype=np.uint64)
    a2 = rng.integers(0, q, (d, 2, n), dtype=np.uint64)
    want = np.stack([O.dbfv_mul(P, b, d, 0, x, bb, rlk_arr, threads=O.max_threads()) for x in (a, a2)])
    ct1 = np.tile(np.stack([a, a2]), (13, 1, 1, 1))                      # 26 pairs: per-limb kernel when legal
    ct2 = np.tile(bb[None], (26, 1, 1, 1))
    got = batch.to_host(batch.dbfv_mul(dp, batch.to_device(ct1), batch.to_device(ct2), rlk))
    assert np.array_equal(got[:2], want) and np.array_equal(got[-2:], want)


@pytest.mark.parametrize("preset", ["compact", "u64", "cfg3", "toy16_noaux", "n64_base10"])
def test_mul_no_relin_relinearize_gadget_decompose(preset):
    """The halves of bfv_mul_and_relin as the reference exposes them (bfv/eval.rs:89-108, bfv/keyswitch.rs:11-101):
    word-for-word against the oracle, relinearize(mul_no_relin) == bfv_mul_and_relin, the degree-2 product decrypts
    (c0 + c1 s + c2 s^2), the reference's gadget KAT and guards."""
    from exacto_b200 import batch
    P = {"compact": H.compact_bfv(), "u64": H.u64_dbfv().bfv, "cfg3": H.cfg3_prime().bfv,
         "toy16_noaux": O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=29, gadget_base=8),
         "n64_base10": O.OracleParams(n=64, q=1152921504606830593, aux=(18014398509998081, 36028797018972161),
                                      plain_modulus=257, gadget_base=10)}[preset]
    params = to_params(P)
    q, n = P.q, P.n
    rng = np.random.default_rng(n + 5)
    s = H.gen_secret_key(P, rng)
    rlk_arr = H.gen_relin_key(P, s, rng)
    rlk = E.RelinKey(rlk_arr, params)
    B = 24 if n == 4096 else 3                         # n = 4096: enough limbs for the one-CTA-per-limb relin kernel too
    ct1 = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    ct1[0] = H.encrypt_sk(P, H.encode_scalar(P, 3), s, rng); ct2[0] = H.encrypt_sk(P, H.encode_scalar(P, 7), s, rng)
    want3 = np.stack([O.bfv_mul_no_relin(P, a, b) for a, b in zip(ct1, ct2)])
    got3 = E.bfv_mul_no_relin_batch(params, ct1, ct2)
    assert np.array_equal(got3, want3)
    want2 = O.relinearize(P, want3, rlk_arr)
    assert np.array_equal(E.relinearize_batch(params, want3, rlk), want2)
    assert np.array_equal(E.relinearize_batch(params, want3[:2], rlk), want2[:2])          # small batch: wide relin
    assert np.array_equal(want2, E.bfv_mul_and_relin_batch(params, ct1, ct2, rlk))
    d3 = batch.bfv_mul_no_relin(params, batch.to_device(ct1), batch.to_device(ct2))
    assert np.array_equal(batch.to_host(batch.relinearize(params, d3, rlk)), want2)
    sk = E.SecretKey.from_ntt(s, params)
    a, b = E.BfvCiphertext.from_array(ct1[0], params), E.BfvCiphertext.from_array(ct2[0], params)
    prod3 = E.bfv_mul_no_relin(a, b)
    t = P.plain_modulus
    assert len(prod3.c) == 3 and E.decode_scalar(E.decrypt(prod3, sk)) == 21 % t
    assert E.decode_scalar(E.decrypt(E.relinearize(prod3, rlk), sk)) == 21 % t
    assert E.relinearize(a, rlk) is a                                                    # keyswitch.rs:63-65
    with pytest.raises(E.ExactoError, match="relinearization only supports degree-2 ciphertexts"):
        E.relinearize(E.BfvCiphertext(prod3.c + prod3.c[:1], params), rlk)
    coeffs = rng.integers(0, q, n, dtype=np.uint64)
    coeffs[:6] = [0, 1, q - 1, q // 2, q // 2 + 1, 42]
    digs = E.gadget_decompose(E.CoeffPoly(coeffs, q), params)
    assert np.array_equal(np.stack([dg.coeffs for dg in digs]), O.gadget_decompose(coeffs, q, P.gadget_base, P.gadget_digits))
    if preset == "toy16_noaux":                                                          # keyswitch.rs:109-116 KAT
        kat = to_params(O.OracleParams(n=16, q=65537, aux=(), plain_modulus=5, gadget_base=16))   # G = 5 for this q
        c = np.zeros(16, np.uint64); c[0] = 42
        assert [int(dg.coeffs[0]) for dg in E.gadget_decompose(E.CoeffPoly(c, 65537), kat)] == [65531, 3, 0, 0, 0]


@pytest.mark.parametrize("name,base,d,p,gb", [("d4_b2^16", 1 << 16, 4, 34_359_738_367, 256), ("d16_b2^4", 1 << 4, 16, 12_289, 16)])
def test_paper_repro_other_profiles(name, base, d, p, gb):
    """The other two profiles of src/bin/paper_repro.rs:41-65 (the u64 profile is cfg 4): d = 4 with a 35-bit BFV
    plaintext modulus, d = 16 with gadget base 16 (G = 15) -- small batch (per-product kernels) and a batch large
    enough for the per-limb kernel where the plan allows it, vs the oracle; decrypt KAT 13579 * 24680 where the profile has the noise budget."""
    from exacto_b200 import batch
    P = O.OracleParams(n=4096, q=1152921504606830593, aux=(18014398509998081, 36028797018972161), plain_modulus=p, gadget_base=gb)
    S = H.DbfvSetup(P, base, d, 0)
    dp = to_dbfv_params(P, base, d, 0)
    rng = np.random.default_rng(1337 + d)
    s = H.gen_secret_key(P, rng)
    rlk_arr = H.gen_relin_key(P, s, rng)
    rlk = E.RelinKey(rlk_arr, dp.bfv_params)
    ca, cb = H.dbfv_encrypt_sk(S, 13_579, s, rng), H.dbfv_encrypt_sk(S, 24_680, s, rng)
    want = O.dbfv_mul(P, base, d, 0, ca, cb, rlk_arr, threads=O.max_threads())
    got = E.dbfv_mul(E.DbfvCiphertext.from_array(ca, dp), E.DbfvCiphertext.from_array(cb, dp), rlk)
    assert np.array_equal(got.to_array(), want)
    if d == 16:      # the d = 4 profile has no noise budget for one multiplication (reports/paper_reproduction.md: depth 0)
        assert H.dbfv_decrypt(S, got.to_array(), s) == 13_579 * 24_680
    B = 52 if d == 4 else 13        # 52 x 2 duos x 2 = 208, 13 x 8 duos x 2 = 208 CTAs: the per-limb kernel
    ct1 = rng.integers(0, P.q, (B, d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (B, d, 2, P.n), dtype=np.uint64)
    ct1[1] = ca; ct2[1] = cb
    outs = batch.to_host(batch.dbfv_mul(dp, batch.to_device(ct1), batch.to_device(ct2), rlk))
    assert np.array_equal(outs[1], want)
    for i in (0, B - 1):
        assert np.array_equal(outs[i], O.dbfv_mul(P, base, d, 0, ct1[i], ct2[i], rlk_arr, threads=O.max_threads()))


def test_bootstrap_fast_path_at_full_ring_degree():
    """The bootstrap pipeline beyond the reference's toy scale: n = 4096, original scheme without an auxiliary basis
    (50-bit q, t = 5), boot scheme = the u64 profile's primes with t_boot = 29, q' = 25.  Trivial ciphertexts take the
    fast path (modulus switch, re-encryption under bsk, degree-28 rounding polynomial): decode to m, and equal the
    oracle pipeline word for word."""
    from oracle import bootstrap_ref as B
    orig = O.OracleParams(n=4096, q=1125899906826241, aux=(), plain_modulus=5, gadget_base=256)
    boot = O.OracleParams(n=4096, q=1152921504606830593, aux=(18014398509998081, 36028797018972161), plain_modulus=29,
                          gadget_base=256)
    op, bp = to_params(orig), to_params(boot)
    rng = np.random.default_rng(4096)
    s = H.gen_secret_key(orig, rng)
    boot_s = B.create_boot_sk(orig, boot, s)
    s_pt = B._center_to(O.ntt_inv(s, orig.q), orig.q, boot.plain_modulus)
    bk = B.BootstrapKeyRef(H.encrypt_sk(boot, s_pt, boot_s, rng), boot, H.gen_relin_key(boot, boot_s, rng), {},
                           B.compute_rounding_poly(5, 25, 29), 5, 25)
    pk = _product_bootstrap_key(bk, bp)
    boot_sk = E.SecretKey.from_ntt(boot_s, bp)
    for m in (0, 3, 4):
        out = E.bfv_bootstrap(E.trivial_encrypt(m, op), pk)
        assert E.decode_scalar(E.decrypt(out, boot_sk)) % 5 == m
        if m == 3:
            assert np.array_equal(out.to_array(), B.bfv_bootstrap(orig, H.trivial_encrypt(orig, m), bk))
