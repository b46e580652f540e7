"""CPU tier: the product's CUDA kernels (exacto_b200/csrc/kernels.cu, compiled unchanged by g++
under tests/host_emul's CUDA shim: one OS thread per CUDA thread, pthread barriers for
__syncthreads) against the oracle and the golden fixtures.  Checks index math, barriers,
lazy-reduction bounds and the host-precomputed constants before any GPU time is spent."""
import numpy as np
import pytest

from common import CASES, Emulator, H, O, digest, golden, golden_inputs


@pytest.fixture(scope="module")
def emu():
    return Emulator()


@pytest.mark.parametrize("n,moduli", [(8, [65537]), (16, [65537]), (32, [65537, 1152921504606830593]), (128, [1099509805057]),
                                      (64, [1152921504606830593, 18014398509998081]),
                                      (1024, [1099509805057, 562949953443841]),
                                      (4096, [1152921504606830593, 18014398509998081, 36028797018972161])])
def test_ntt_kernels(emu, n, moduli):
    rc, h, err = emu.create(n, moduli[:1], moduli[1:], 2)
    assert rc == 0, err
    rng = np.random.default_rng(n)
    for base, m in enumerate(moduli):
        x = rng.integers(0, m, (2, n), dtype=np.uint64)
        x[0, :4] = [0, 1, m - 1, m // 2]
        y = emu.ntt(h, base, True, x)
        assert np.array_equal(y, O.ntt_fwd(x, m))
        assert np.array_equal(emu.ntt(h, base, False, y), x)
    assert emu.info(h)[3] == O.find_psi(n, moduli[0])


def test_poly_ops(emu):
    q = 1152921504606830593
    rc, h, err = emu.create(64, [q], [], 2)
    assert rc == 0, err
    rng = np.random.default_rng(8)
    a, b = rng.integers(0, q, 64, dtype=np.uint64), rng.integers(0, q, 64, dtype=np.uint64)
    a[:3] = [0, q - 1, 1]; b[:3] = [0, q - 1, q - 1]
    ao, bo = a.astype(object), b.astype(object)
    assert np.array_equal(emu.poly_op(h, 0, 0, a, b), np.array((ao + bo) % q, dtype=np.uint64))
    assert np.array_equal(emu.poly_op(h, 0, 1, a, b), np.array((ao - bo) % q, dtype=np.uint64))
    assert np.array_equal(emu.poly_op(h, 0, 2, a), np.array((-ao) % q, dtype=np.uint64))
    assert np.array_equal(emu.poly_op(h, 0, 3, a, b), np.array((ao * bo) % q, dtype=np.uint64))
    assert np.array_equal(emu.poly_op(h, 0, 4, a, scalar=2 ** 64 - 1), np.array((ao * ((2 ** 64 - 1) % q)) % q, dtype=np.uint64))


@pytest.mark.parametrize("name", ["toy16_a1", "n64_a2_rep", "n32_a2_base7", "compact_dbfv", "cfg3p_dbfv"])
def test_dbfv_mul_matches_golden(emu, name):
    P, base, d, pm, seed, full = CASES[name]
    g = golden()
    ct1, ct2, rlk = golden_inputs(P, d, seed)
    h = emu.from_oracle(P)
    rc, out, err = emu.dbfv_mul(h, base, d, pm, ct1, ct2, rlk)
    assert rc == 0, err
    assert digest(out) == str(g[f"{name}/dbfv_sha256"])
    # bfv_mul_and_relin is the d = 1 case of the same pipeline
    rc, one, err = emu.dbfv_mul(h, 2, 1, 0, ct1[:1], ct2[:1], rlk)
    assert rc == 0, err
    assert digest(one[0]) == str(g[f"{name}/bfv_sha256"])


def test_flags_masks_and_batches(emu):
    P, base, d, pm = CASES["n64_a2_rep"][:4]
    h = emu.from_oracle(P)
    rng = np.random.default_rng(21)
    ct1 = rng.integers(0, P.q, (3, d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (3, d, 2, P.n), dtype=np.uint64)
    rlk = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    want = np.stack([O.dbfv_mul(P, base, d, pm, a, b, rlk) for a, b in zip(ct1, ct2)])
    rc, got, err = emu.dbfv_mul(h, base, d, pm, ct1, ct2, rlk)
    assert rc == 0 and np.array_equal(got, want), err
    rc, got_all, _ = emu.dbfv_mul(h, base, d, pm, ct1, ct2, rlk, flags=1)       # EXB_DBFV_ALL_PRODUCTS
    assert np.array_equal(got_all, want)
    # zero small reps (p = b^d): dead products skipped vs computed give identical limbs
    want0 = np.stack([O.dbfv_mul(P, 16, 2, 256, a, b, rlk) for a, b in zip(ct1, ct2)])
    for flags in (0, 1):
        rc, g0, _ = emu.dbfv_mul(h, 16, 2, 256, ct1, ct2, rlk, flags=flags)
        assert np.array_equal(g0, want0)
    # limb masks: each rank's limbs are exact, other limbs untouched
    sentinel = np.full_like(ct1, 7)
    rc, part, _ = emu.dbfv_mul(h, base, d, pm, ct1, ct2, rlk, limb_mask=0b10, out=sentinel.copy())
    assert np.array_equal(part[:, 1], want[:, 1]) and np.array_equal(part[:, 0], sentinel[:, 0])
    # relinearize uses min(G, rlk.keys.len()) keys (bfv/keyswitch.rs:86-89)
    rc, short, _ = emu.dbfv_mul(h, 2, 1, 0, ct1[:1, :1], ct2[:1, :1], rlk[:2])
    full_key = O.bfv_mul_and_relin(P, ct1[0, 0], ct2[0, 0], np.concatenate([rlk[:2], np.zeros_like(rlk[2:])]))
    assert np.array_equal(short[0, 0], full_key)


def test_edge_inputs(emu):
    """All-zero, all-(q-1), half-boundary inputs: the centring / rounding branches."""
    P = H.cfg3_prime().bfv
    h = emu.from_oracle(P)
    q, n = P.q, P.n
    rlk = np.random.default_rng(4).integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    pats = [np.zeros(n, np.uint64), np.full(n, q - 1, np.uint64), np.full(n, q // 2, np.uint64),
            np.full(n, q // 2 + 1, np.uint64)]
    ct1 = np.stack([np.stack([O.ntt_fwd(pats[i], q), O.ntt_fwd(pats[(i + 1) % 4], q)]) for i in range(4)])
    ct2 = np.stack([np.stack([O.ntt_fwd(pats[(i + 2) % 4], q), O.ntt_fwd(pats[(i + 3) % 4], q)]) for i in range(4)])
    rc, got, err = emu.dbfv_mul(h, 2, 1, 0, ct1[:, None], ct2[:, None], rlk)
    assert rc == 0, err
    assert np.array_equal(got[:, 0], O.bfv_mul_and_relin(P, ct1, ct2, rlk, threads=4))


def test_dispatch_errors(emu):
    n = 4096
    rc, h, _ = emu.create(n, [18014398509506561], [36028797018972161], 1040407, 256)
    assert rc == 0
    st = emu.info(h)
    assert st[2] == 1 and "single aux prime too small for HPS centering" in st[4]
    rc, h, _ = emu.create(n, [18014398509506561], [], 1040407, 256)
    st = emu.info(h)
    assert st[2] == 9 and "schoolbook BFV multiplication can overflow i128" in st[4]
    rc, h, _ = emu.create(16, [65537, 1099509805057], [], 257, 8)
    st = emu.info(h)
    assert st[1] == 19 and st[2] == 0                    # the reference's multi-prime set runs on the device (rns_kernels.cu)
    rc, h, err = emu.create(4096, [0xFFFFFFFFFFE00001], [], 257)
    assert rc == 1 and "cannot create NTT plan" in err
    rc, h, err = emu.create(1000, [65537], [], 257)
    assert rc == 4


def test_internal_small_basis_selection(emu, monkeypatch):
    """The 27-bit internal auxiliary basis is used exactly when it is provably result-identical
    (n = 4096 fast path, two-aux configs where the reference's centred CRT cannot wrap)."""
    for name in ("cfg3p_dbfv", "u64_dbfv"):
        P = CASES[name][0]
        primes = emu.small_primes(emu.from_oracle(P))
        assert len(primes) == 3 and all(2 ** 26 < p < 2 ** 27 and p % 8192 == 1 and O.is_prime(p) for p in primes)
        prod = primes[0] * primes[1] * primes[2]
        assert prod >= (P.n * P.q // 2 + 2) << 8
    assert emu.small_primes(emu.from_oracle(CASES["compact_dbfv"][0])) == []          # n = 1024: generic path
    monkeypatch.setenv("EXB_AUX_BASIS", "reference")
    assert emu.small_primes(emu.from_oracle(CASES["cfg3p_dbfv"][0])) == []


def test_reference_aux_basis_path_still_exact(emu, monkeypatch):
    """The 64-bit reference-basis kernels (used whenever the internal basis is not allowed) stay covered."""
    monkeypatch.setenv("EXB_AUX_BASIS", "reference")
    P, base, d, pm, seed, _ = CASES["cfg3p_dbfv"]
    ct1, ct2, rlk = golden_inputs(P, d, seed)
    rc, out, err = emu.dbfv_mul(emu.from_oracle(P), base, d, pm, ct1, ct2, rlk)
    assert rc == 0, err
    assert digest(out) == str(golden()[f"cfg3p_dbfv/dbfv_sha256"])


def test_small_basis_adversarial_inputs(emu):
    """Extreme tensor magnitudes: all coefficients at +-q/2 so |m| reaches its n*q/2 bound."""
    P = H.u64_dbfv().bfv
    h = emu.from_oracle(P)
    assert emu.small_primes(h)
    q, n = P.q, P.n
    hi, lo = np.full(n, q // 2, np.uint64), np.full(n, q // 2 + 1, np.uint64)
    alt = np.where(np.arange(n) % 2 == 0, np.uint64(q // 2), np.uint64(q // 2 + 1))
    pats = [hi, lo, alt]
    rlk = np.random.default_rng(9).integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    ct1 = np.stack([np.stack([O.ntt_fwd(pats[i], q), O.ntt_fwd(pats[(i + 1) % 3], q)]) for i in range(3)])
    ct2 = np.stack([np.stack([O.ntt_fwd(pats[(i + 2) % 3], q), O.ntt_fwd(pats[i], q)]) for i in range(3)])
    rc, got, err = emu.dbfv_mul(h, 2, 1, 0, ct1[:, None], ct2[:, None], rlk)
    assert rc == 0, err
    assert np.array_equal(got[:, 0], O.bfv_mul_and_relin(P, ct1, ct2, rlk, threads=4))


# ---- Galois automorphism + key switch -------------------------------------------------
@pytest.mark.parametrize("preset,elements", [("toy16", [3, 5, 31]), ("compact", [3, 2047]), ("u64", [3, 8191]),
                                             ("cfg3", [4097])])
def test_galois_kernel_matches_oracle(emu, preset, elements):
    """galois_kernel (generic and n = 4096 paths; power-of-two gadget bases 2^8 / 2^16 and a general
    base) against the literal restatement of bfv/eval.rs:512-561 on uniform and edge inputs."""
    P = {"toy16": O.OracleParams(n=16, q=1152921504606830593, aux=(18014398509998081,), plain_modulus=17,
                                 gadget_base=10),
         "compact": H.compact_bfv(), "u64": H.u64_dbfv().bfv, "cfg3": H.cfg3_prime().bfv}[preset]
    h = emu.from_oracle(P)
    q, n = P.q, P.n
    rng = np.random.default_rng(n)
    gk = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    ct = rng.integers(0, q, (3, 2, n), dtype=np.uint64)
    edge = np.zeros(n, np.uint64)
    edge[:6] = [0, 1, q - 1, q // 2, q // 2 + 1, 2]
    ct[1, 0] = O.ntt_fwd(edge, q); ct[1, 1] = O.ntt_fwd(edge[::-1].copy(), q)
    ct[2] = 0
    for k in elements:
        got = emu.bfv_apply_automorphism(h, ct, k, gk)
        assert np.array_equal(got, O.bfv_apply_automorphism(P, ct, gk, k, threads=4)), (preset, k)


# ---- decrypt --------------------------------------------------------------------------
@pytest.mark.parametrize("preset", ["compact", "u64", "toy16"])
def test_decrypt_kernel_matches_definition(emu, preset):
    """decrypt_kernel against the big-int restatement of bfv/encrypt.rs:111-178 (oracle/harness.py:decrypt) on
    real encryptions (degree 1), a degree-2 product and uniform / edge phases."""
    P = {"compact": H.compact_bfv(), "u64": H.u64_dbfv().bfv,
         "toy16": O.OracleParams(n=16, q=1152921504606830593, aux=(18014398509998081,), plain_modulus=17,
                                 gadget_base=10)}[preset]
    h = emu.from_oracle(P)
    q, n, t = P.q, P.n, P.plain_modulus
    rng = np.random.default_rng(5)
    s = H.gen_secret_key(P, rng)
    pt = rng.integers(0, t, n, dtype=np.uint64)
    ct = np.stack([H.encrypt_sk(P, pt, s, rng), rng.integers(0, q, (2, n), dtype=np.uint64)])
    got = emu.bfv_decrypt(h, ct, s)
    assert np.array_equal(got[0], pt) and np.array_equal(got[0], H.decrypt(P, ct[0], s))
    assert np.array_equal(got[1], H.decrypt(P, ct[1], s))
    edge = np.zeros(n, np.uint64)
    edge[:8] = [0, 1, q - 1, q // 2, q // 2 + 1, q // t, q - q // (2 * t), q - q // (2 * t) - 1]
    c0 = np.stack([O.ntt_fwd(edge, q), np.zeros(n, np.uint64)])[None]
    assert np.array_equal(emu.bfv_decrypt(h, c0, s)[0], H.decrypt(P, c0[0], s))
    if preset != "toy16":
        c3 = O.bfv_mul_no_relin(P, ct[0], H.encrypt_sk(P, pt, s, rng))[None]       # degree 2: c0 + c1 s + c2 s^2
        assert np.array_equal(emu.bfv_decrypt(h, c3, s)[0], H.decrypt(P, c3[0], s))


# ---- parameter sets without an auxiliary basis (the reference's schoolbook branch) --------------------
NO_AUX = {
    "boot_orig": dict(P=O.OracleParams(n=16, q=65537, aux=(), plain_modulus=5), dbfv=None),
    "boot_scheme": dict(P=O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=29, gadget_base=8), dbfv=None),
    "boot_dbfv": dict(P=O.OracleParams(n=16, q=65537, aux=(), plain_modulus=97, gadget_base=8), dbfv=(4, 2, 16)),
    "n256_q40": dict(P=O.OracleParams(n=256, q=1099509805057, aux=(), plain_modulus=257), dbfv=(16, 2, 256)),
    "n4096_q50": dict(P=O.OracleParams(n=4096, q=1125899906826241, aux=(), plain_modulus=257, gadget_base=256), dbfv=None),
}


@pytest.mark.parametrize("name", list(NO_AUX))
def test_no_aux_params_match_schoolbook(emu, name):
    """bfv/eval.rs:415-464: without an auxiliary basis the reference convolves exactly in i128 and rounds.  The
    device runs its HPS pipeline on an internal auxiliary basis (two 61-bit primes, or the 27-bit basis at
    n = 4096); the result must equal the literal O(n^2) restatement word for word (bootstrap test parameter
    sets of bootstrap/bfv_host.rs:345-386 included)."""
    P, dbfv = NO_AUX[name]["P"], NO_AUX[name]["dbfv"]
    h = emu.from_oracle(P)
    assert emu.info(h)[2] == 0, emu.info(h)
    q, n = P.q, P.n
    rng = np.random.default_rng(len(name))
    rlk = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    B = 2 if n == 4096 else 4
    ct1 = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    edge = np.full(n, q // 2, np.uint64); edge[::2] = q // 2 + 1
    ct1[0] = O.ntt_fwd(np.stack([edge, edge[::-1].copy()]), q)           # worst-case |t| coefficients
    ct2[0] = O.ntt_fwd(np.stack([edge, edge]), q)
    rc, got, err = emu.dbfv_mul(h, 2, 1, 0, ct1[:, None], ct2[:, None], rlk)
    assert rc == 0, err
    assert np.array_equal(got[:, 0], O.bfv_mul_and_relin(P, ct1, ct2, rlk, threads=4))
    if dbfv:
        b, d, pm = dbfv
        a = rng.integers(0, q, (d, 2, n), dtype=np.uint64); c = rng.integers(0, q, (d, 2, n), dtype=np.uint64)
        rc, got, err = emu.dbfv_mul(h, b, d, pm, a[None], c[None], rlk)
        assert rc == 0, err
        assert np.array_equal(got[0], O.dbfv_mul(P, b, d, pm, a, c, rlk, threads=4))


def test_no_aux_overflow_risk_still_refused(emu):
    """README config 3 (n = 4096, 59-bit q, no aux): the reference errors (bfv/eval.rs:426-431); so do we."""
    P = O.OracleParams(n=4096, q=576460752308273153, aux=(), plain_modulus=65537)
    h = emu.from_oracle(P)
    info = emu.info(h)
    assert info[2] == 9
    rc, _, err = emu.dbfv_mul(h, 2, 1, 0, np.zeros((1, 1, 2, 4096), np.uint64), np.zeros((1, 1, 2, 4096), np.uint64),
                              np.zeros((P.gadget_digits, 2, 4096), np.uint64))
    assert rc == 9 and "schoolbook BFV multiplication can overflow i128" in err


def test_per_limb_tensor_path_is_taken_and_exact(emu):
    """tensor01_kernel (components 0/1 summed per output limb before ONE small-prime inverse transform) is what
    the n = 4096 presets run, including with all 64 products (limbs k >= d sum up to d-1 products), limb masks
    and worst-case inputs: every |t_ij| at its bound with equal signs maximises |sum m_ij|."""
    S = H.u64_dbfv()
    P = S.bfv
    h = emu.from_oracle(P)
    assert emu.tensor_per_limb(h, S.base, S.d, 0) == 1 and emu.tensor_per_limb(h, S.base, S.d, 0, flags=1) == 1
    assert emu.tensor_per_limb(h, 256, 2, 65536) == 1
    assert emu.tensor_per_limb(emu.from_oracle(H.compact_bfv()), 16, 2, 256) == 0        # n = 1024: generic kernels
    assert emu.tensor_per_limb(h, 2, 1, 0) == 0                                          # d = 1: nothing to hoist
    q, n = P.q, P.n
    rng = np.random.default_rng(77)
    rlk = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    half = np.full(n, q // 2, np.uint64)
    a = np.stack([np.stack([O.ntt_fwd(half, q), O.ntt_fwd(half, q)]) for _ in range(S.d)])
    b = a.copy()
    b[::2] = np.stack([O.ntt_fwd(half + np.uint64(1), q)] * 2)                       # -q/2 rows: mixed signs too
    for flags, mask in [(0, 0), (1, 0), (0, 0b10100000)]:
        want = O.dbfv_mul(P, S.base, S.d, S.plain_modulus, a, b, rlk, threads=8)
        init = np.zeros_like(a[None])
        rc, got, err = emu.dbfv_mul(h, S.base, S.d, S.plain_modulus, a[None], b[None], rlk, flags=flags, limb_mask=mask, out=init)
        assert rc == 0, err
        ks = [k for k in range(S.d) if not mask or (mask >> k) & 1]
        assert np.array_equal(got[0][ks], want[ks]), (flags, mask)
    # the per-product kernel (what small batches run on the GPU) on the same worst-case inputs, also with the
    # one-CTA-per-transform relinearisation of small batches (relin12_wide_kernel + relin_reduce_kernel)
    for fl in (0x80000000, 0xC0000000, 0x40000000, 0x40000001):
        rc, got, err = emu.dbfv_mul(h, S.base, S.d, S.plain_modulus, a[None], b[None], rlk, flags=fl)
        assert rc == 0 and np.array_equal(got[0], want), (hex(fl), err)


# ---- the halves of bfv_mul_and_relin as stand-alone entry points ---------------------------------------
@pytest.mark.parametrize("preset", ["compact", "u64", "cfg3", "toy16_noaux", "n64_base10"])
def test_mul_no_relin_relinearize_gadget_decompose(emu, preset):
    """bfv_mul_no_relin (bfv/eval.rs:89-108), relinearize (bfv/keyswitch.rs:59-101) and gadget_decompose (:11-52)
    against the literal restatement, and relinearize(mul_no_relin(a, b)) == bfv_mul_and_relin(a, b) word for word."""
    P = {"compact": H.compact_bfv(), "u64": H.u64_dbfv().bfv, "cfg3": H.cfg3_prime().bfv,
         "toy16_noaux": O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=29, gadget_base=8),
         "n64_base10": O.OracleParams(n=64, q=1152921504606830593, aux=(18014398509998081, 36028797018972161),
                                      plain_modulus=257, gadget_base=10)}[preset]
    h = emu.from_oracle(P)
    q, n = P.q, P.n
    rng = np.random.default_rng(n + 5)
    B = 2
    ct1 = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, q, (B, 2, n), dtype=np.uint64)
    edge = np.full(n, q // 2, np.uint64); edge[::2] = q // 2 + 1
    ct1[0] = O.ntt_fwd(np.stack([edge, edge[::-1].copy()]), q)
    rlk = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    rc, c3, err = emu.bfv_mul_no_relin(h, ct1, ct2)
    assert rc == 0, err
    want3 = np.stack([O.bfv_mul_no_relin(P, a, b) for a, b in zip(ct1, ct2)])
    assert np.array_equal(c3, want3)
    want2 = O.relinearize(P, want3, rlk)
    for wide in (False, True):
        rc, got2 = emu.bfv_relinearize(h, want3, rlk, wide=wide)
        assert rc == 0 and np.array_equal(got2, want2), wide
    assert np.array_equal(want2, O.bfv_mul_and_relin(P, ct1, ct2, rlk))
    rnd3 = rng.integers(0, q, (B, 3, n), dtype=np.uint64)                  # any degree-2 input, not only products
    rc, got = emu.bfv_relinearize(h, rnd3, rlk)
    assert rc == 0 and np.array_equal(got, O.relinearize(P, rnd3, rlk))
    coeffs = rng.integers(0, q, (2, n), dtype=np.uint64)
    coeffs[0, :6] = [0, 1, q - 1, q // 2, q // 2 + 1, 42]
    want_d = np.stack([O.gadget_decompose(c, q, P.gadget_base, P.gadget_digits) for c in coeffs])
    assert np.array_equal(emu.gadget_decompose(h, coeffs, P.gadget_digits), want_d)


def test_per_limb_tensor_path_sixteen_digits(emu):
    """d = 16 (paper_repro's third profile): limbs sum up to 16 products -- the largest count the exactness
    conditions allow for a 60-bit q -- on inputs where every centred residue and every |t_ij| sits at its bound."""
    P = O.OracleParams(n=4096, q=1152921504606830593, aux=(18014398509998081, 36028797018972161), plain_modulus=12289,
                       gadget_base=16)
    h = emu.from_oracle(P)
    d = 16
    assert emu.tensor_per_limb(h, 16, d, 0) == 1
    q, n = P.q, P.n
    rng = np.random.default_rng(16)
    rlk = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    half = O.ntt_fwd(np.full(n, q // 2, np.uint64), q)
    a = np.stack([np.stack([half, half])] * d)
    b = a.copy()
    b[1::2] = rng.integers(0, q, (d // 2, 2, n), dtype=np.uint64)
    want = O.dbfv_mul(P, 16, d, 0, a, b, rlk, threads=8)
    rc, got, err = emu.dbfv_mul(h, 16, d, 0, a[None], b[None], rlk)
    assert rc == 0 and np.array_equal(got[0], want), err


@pytest.mark.parametrize("plain,per_limb", [((1 << 28) - 57, 1), ((1 << 29) + 11, 1), ((1 << 33) + 7, 1)])
def test_per_limb_rounding_sums_i32_and_i64(emu, plain, per_limb):
    """tensor01_kernel keeps the per-limb sums of rounding terms in an i32 image while (products per limb) * (p/2 + 2)
    < 2^31 and in an i64 image otherwise (tensor01_kernel<true>).  p just below 2^28 with 8 products per limb and every
    residue at q/2 puts the i32 sum at 2^30; p above 2^29 and above 2^32 (64-bit rounding terms) take the wide variant.
    All word-exact."""
    P = O.OracleParams(n=4096, q=1152921504606830593, aux=(18014398509998081, 36028797018972161), plain_modulus=plain,
                       gadget_base=256)
    h = emu.from_oracle(P)
    d, b = 8, 256
    assert emu.tensor_per_limb(h, b, d, 0) == per_limb
    q, n = P.q, P.n
    rng = np.random.default_rng(plain & 0xffff)
    rlk = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    one = np.zeros(n, np.uint64); one[0] = 1
    half = O.ntt_fwd(np.full(n, q // 2, np.uint64), q)
    # (q/2) * 1 = q/2 per product and coefficient in component 0: every rounding term is round(p/2) with one sign
    a = np.stack([np.stack([half, half])] * d)
    bb = np.stack([np.stack([O.ntt_fwd(one, q), O.ntt_fwd(one, q)])] * d)
    bb[1::3] = rng.integers(0, q, (len(bb[1::3]), 2, n), dtype=np.uint64)
    want = O.dbfv_mul(P, b, d, 0, a, bb, rlk, threads=8)
    rc, got, err = emu.dbfv_mul(h, b, d, 0, a[None], bb[None], rlk)
    assert rc == 0 and np.array_equal(got[0], want), err


def test_schoolbook_middle_term_band_refused(emu):
    """The band the reference's overflow guard misses (bfv/eval.rs:457-464 bounds n (q/2)^2 p, the middle tensor
    term reaches twice that): refused, while the literal oracle wraps and the big-int definition does not."""
    P = O.OracleParams(n=64, q=1152921504606844417, aux=(), plain_modulus=8, gadget_base=10)
    h = emu.from_oracle(P)
    info = emu.info(h)
    assert info[2] == 9
    rc, _, err = emu.dbfv_mul(h, 2, 1, 0, np.zeros((1, 1, 2, 64), np.uint64), np.zeros((1, 1, 2, 64), np.uint64),
                              np.zeros((P.gadget_digits, 2, 64), np.uint64))
    assert rc == 9 and "overflows i128 in its middle tensor term" in err
    assert emu.info(emu.from_oracle(O.OracleParams(n=64, q=1152921504606844417, aux=(), plain_modulus=4, gadget_base=10)))[2] == 0


@pytest.mark.skipif(not __import__("os").environ.get("EXB_RUN_TSAN"), reason="set EXB_RUN_TSAN=1 (builds with -fsanitize=thread, ~2 min)")
def test_kernels_race_free_under_thread_sanitizer():
    """compute-sanitizer racecheck is closed on the GPU pool: the kernels run on the host emulator under
    ThreadSanitizer instead (tools/tsan_kernels.cpp); the self-test proves a missing barrier would be reported."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([os.path.join(root, "tools", "run_tsan.sh")], capture_output=True, text=True)
    assert res.returncode == 0 and "ThreadSanitizer" not in res.stdout + res.stderr, res.stdout + res.stderr
    st = subprocess.run(["/tmp/exb_tsan_kernels", "--selftest"], capture_output=True, text=True)
    assert "ThreadSanitizer: data race" in st.stdout + st.stderr


# ---- multi-prime ciphertext modulus: rns_kernels.cu / rns.cuh vs oracle/rns_ref.py --------------------------
@pytest.mark.parametrize("name", ["ref_n16", "n64_two60", "n32_three40_d2"])
def test_multi_prime_path_matches_bigint_oracle(emu, name):
    """bfv_mul_generic_rns (bfv/eval.rs:113-147) and the L > 1 relinearize (bfv/keyswitch.rs:59-101 on the
    truncating to_coeff_poly of ring/rns.rs:114-151): the product's extended-basis kernels, replayed on the CPU,
    word for word against the big-integer restatement."""
    from common import RNS_CASES, R, rns_inputs
    P, d, base, pm = RNS_CASES[name]
    rc, h, err = emu.create(P.n, list(P.moduli), [], P.plain_modulus, P.gadget_base, 0)
    assert rc == 0, err
    rc, L, K, ext, err = emu.rns_info(h)
    assert rc == 0 and L == len(P.moduli), err
    prod = 1
    for e in ext:
        assert O.is_prime(e) and e % (2 * P.n) == 1 and e not in P.moduli
        prod *= e
    assert prod > 4 * P.n * P.Q * P.Q                           # every tensor coefficient is represented exactly
    assert emu.info(h)[1] == P.G
    pairs = 2
    ct1, ct2, rlk = rns_inputs(P, d, pairs, 5)
    n = P.n
    if d == 1:
        rc, got3, err = emu.rns_mul(h, 2, 1, 0, ct1, ct2, rlk[:0], 1, (pairs, 3, L, n))
        assert rc == 0, err
        want3 = np.stack([R.bfv_mul_no_relin(P, a[0], b[0]) for a, b in zip(ct1, ct2)])
        assert np.array_equal(got3, want3)
        rc, got2, err = emu.rns_mul(h, 2, 1, 0, want3, want3, rlk, 2, (pairs, 2, L, n))
        assert rc == 0 and np.array_equal(got2, np.stack([R.relinearize(P, c, rlk) for c in want3]))
        rc, got2, err = emu.rns_mul(h, 2, 1, 0, want3, want3, rlk[:3], 2, (pairs, 2, L, n))     # fewer keys than digits
        assert rc == 0 and np.array_equal(got2, np.stack([R.relinearize(P, c, rlk[:3]) for c in want3]))
    rc, got, err = emu.rns_mul(h, base, d, pm, ct1, ct2, rlk, 0, (pairs, d, 2, L, n))
    assert rc == 0, err
    assert np.array_equal(got, np.stack([R.dbfv_mul(P, d, a, b, rlk) for a, b in zip(ct1, ct2)]))


def test_multi_prime_dispatch_limits(emu):
    """What the device path refuses for L > 1, with the reason (the single-prime messages are pinned above)."""
    q40 = [1099509805057, 1099510054913, 1099507695617]
    rc, h, err = emu.create(32, [1152921504606830593, 576460752308273153, 1099509805057], [], 257, 1 << 16)
    assert rc == 0
    assert emu.rns_info(h)[0] == 9 and "2^126" in emu.rns_info(h)[4]                  # Q ~ 2^159
    assert emu.info(h)[2] == 9
    rc, h, err = emu.create(32, q40 + [65537, 786433], [], 257, 1 << 16)            # five primes
    assert rc == 0 and emu.rns_info(h)[0] == 9 and "4 ciphertext primes" in emu.rns_info(h)[4]
    rc, h, err = emu.create(32, [q40[0], q40[0]], [], 257, 1 << 16)
    assert rc == 0 and emu.rns_info(h)[0] == 9 and "coprime" in emu.rns_info(h)[4]
