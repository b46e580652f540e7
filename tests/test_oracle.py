"""CPU tier: pins the oracle against every known-answer test the reference holds for the
ciphertext-multiplication path, against the independent definition oracle and against
the committed golden fixtures.  Citations are into /root/reference/src/."""
import numpy as np
import pytest

from common import CASES, H, O, digest, golden, golden_inputs
from oracle import definition as D


# ---- ring/modular.rs tests :127-202 (m = 65537) -----------------------------------------
def test_modular_ops_match_u128():
    m = 65537
    rng = np.random.default_rng(1)
    for a, b in rng.integers(0, m, (200, 2)):
        a, b = int(a), int(b)
        assert O.mod_mul(a, b, m) == a * b % m
        assert O.mod_add(a, b, m) == (a + b) % m
        assert O.mod_sub(a, b, m) == (a - b) % m
        assert O.mod_neg(a, m) == (-a) % m
    assert O.mod_pow(3, 65536, m) == 1
    assert O.mod_inv(3, m) * 3 % m == 1
    assert O.mod_inv(0, m) is None


# ---- ring/poly.rs:195-202: X^3 * X^3 = -X^2 in Z_17[X]/(X^4+1) ----------------------------
def test_negacyclic_wrap_sign():
    a = np.array([0, 0, 0, 1], np.uint64)
    assert O.poly_mul_naive(a, a, 17).tolist() == [0, 0, 16, 0]


# ---- ring/ntt.rs:170-212 (n = 16, q = 65537) -------------------------------------------------
def test_ntt_roundtrip_and_mul_vs_schoolbook():
    n, q = 16, 65537
    rng = np.random.default_rng(2)
    a, b = rng.integers(0, q, n, dtype=np.uint64), rng.integers(0, q, n, dtype=np.uint64)
    fa, fb = O.ntt_fwd(a, q), O.ntt_fwd(b, q)
    assert np.array_equal(O.ntt_inv(fa, q), a)
    prod = np.array((fa.astype(object) * fb.astype(object)) % q, dtype=np.uint64)
    assert np.array_equal(O.ntt_inv(prod, q), O.poly_mul_naive(a, b, q))
    s = np.array((fa.astype(object) + fb.astype(object)) % q, dtype=np.uint64)
    assert np.array_equal(O.ntt_inv(s, q), (a + b) % np.uint64(q))


@pytest.mark.parametrize("n,q", [(1024, 1099509805057), (4096, 1152921504606830593), (4096, 36028797018972161)])
def test_ntt_is_negacyclic_convolution(n, q):
    rng = np.random.default_rng(3)
    a = np.zeros(n, np.uint64); b = np.zeros(n, np.uint64)
    idx = rng.choice(n, 6, replace=False)
    a[idx] = rng.integers(1, q, 6, dtype=np.uint64); b[idx[::-1]] = rng.integers(1, q, 6, dtype=np.uint64)
    a[n - 1] = q - 1; b[n - 1] = 2                       # forces the X^n = -1 wrap
    fa, fb = O.ntt_fwd(a, q), O.ntt_fwd(b, q)
    prod = np.array((fa.astype(object) * fb.astype(object)) % q, dtype=np.uint64)
    assert np.array_equal(O.ntt_inv(prod, q), O.poly_mul_naive(a, b, q))


# ---- bfv/keyswitch.rs:109-152 gadget KATs ---------------------------------------------------------
def test_gadget_kats():
    assert O.gadget_decompose([42], 65537, 16, 2).ravel().tolist() == [65531, 3]
    q, base, G = 65537, 16, 4
    coeffs = np.array([12345, 54321, 100, 0], np.uint64)
    dg = O.gadget_decompose(coeffs, q, base, G)
    for pos in range(4):
        assert sum(int(dg[g, pos]) * base ** g for g in range(G)) % q == int(coeffs[pos])
    assert O.gadget_decompose([q - 1], q, 16, 4)[0, 0] == q - 1
    for c in [0, 1, 7, 8, 9, q // 2, q // 2 + 1, q - 9, q - 8]:
        assert O.gadget_decompose([c], q, 16, 4).ravel().tolist() == D.gadget_decompose(c, q, 16, 4)


# ---- oracle (literal HPS schedule) == definition oracle (round(p t / q)) ----------------------------
@pytest.mark.parametrize("name", ["toy16_a1", "n64_a2_rep", "n32_a2_base7"])
def test_literal_oracle_equals_definition(name):
    P, base, d, pm, seed, _ = CASES[name]
    ct1, ct2, rlk = golden_inputs(P, d, seed + 100)
    out = O.dbfv_mul(P, base, d, pm, ct1, ct2, rlk)
    want = D.dbfv_mul_coeff(O.ntt_inv(ct1, P.q), O.ntt_inv(ct2, P.q), O.ntt_inv(rlk, P.q), P.q, P.plain_modulus,
                            P.gadget_base, P.gadget_digits, base, d, pm)
    assert np.array_equal(O.ntt_inv(out, P.q), np.array(want, dtype=np.uint64))


def test_small_reps_zero_for_all_configs():
    """SURVEY finding 3: p = b^d in every preset, so every small representative is zero."""
    for base, d, pm in [(16, 2, 256), (256, 2, 65536), (256, 8, 0)]:
        assert not O.small_reps(base, d, pm).any()
        assert np.array_equal(O.small_reps(base, d, pm), np.array(D.small_reps(base, d, pm), dtype=np.int64))
    assert O.small_reps(16, 2, 250).tolist() == [[6, 0]]


# ---- golden fixtures ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name):
    P, base, d, pm, seed, full = CASES[name]
    g = golden()
    ct1, ct2, rlk = golden_inputs(P, d, seed)
    assert digest(ct1, ct2, rlk) == str(g[f"{name}/in_sha256"]), "seeded inputs drifted"
    out = O.dbfv_mul(P, base, d, pm, ct1, ct2, rlk, threads=O.max_threads())
    assert digest(out) == str(g[f"{name}/dbfv_sha256"])
    assert digest(O.bfv_mul_and_relin(P, ct1[0], ct2[0], rlk)) == str(g[f"{name}/bfv_sha256"])
    if full:
        assert np.array_equal(out, g[f"{name}/dbfv_out"])


# ---- decrypt KATs with valid keys (bfv/eval.rs:883-900, dbfv/eval.rs:224-290,345-382,521-564) --------------
def test_compact_bfv_mul_decrypts():
    P = H.compact_bfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(P, rng); rlk = H.gen_relin_key(P, s, rng)
    for a, b in [(3, 7), (0, 5), (10, 20)]:
        c1 = H.encrypt_sk(P, H.encode_scalar(P, a), s, rng); c2 = H.encrypt_sk(P, H.encode_scalar(P, b), s, rng)
        assert int(H.decrypt(P, O.bfv_mul_and_relin(P, c1, c2, rlk), s)[0]) == a * b % P.plain_modulus
        assert int(H.decrypt(P, O.bfv_mul_no_relin(P, c1, c2), s)[0]) == a * b % P.plain_modulus
        assert int(H.decrypt(P, O.bfv_add(P, c1, c2), s)[0]) == (a + b) % P.plain_modulus


def test_compact_dbfv_mul_decrypts():
    S = H.compact_dbfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(S.bfv, rng); rlk = H.gen_relin_key(S.bfv, s, rng)
    for a, b in [(3, 7), (15, 15), (10, 20), (12, 12)]:
        ca, cb = H.dbfv_encrypt_sk(S, a, s, rng), H.dbfv_encrypt_sk(S, b, s, rng)
        out = O.dbfv_mul(S.bfv, S.base, S.d, S.plain_modulus, ca, cb, rlk)
        assert H.dbfv_decrypt(S, out, s) == a * b % 256
    # (3 + X)(2 + X) = 6 + 5X + X^2   (dbfv/eval.rs:243-268)
    pa = np.zeros(S.bfv.n, np.uint64); pb = np.zeros(S.bfv.n, np.uint64)
    pa[:2] = [3, 1]; pb[:2] = [2, 1]
    out = O.dbfv_mul(S.bfv, S.base, S.d, S.plain_modulus, H.dbfv_encrypt_poly_sk(S, pa, s, rng),
                     H.dbfv_encrypt_poly_sk(S, pb, s, rng), rlk)
    assert H.dbfv_decrypt_poly(S, out, s)[:3].tolist() == [6, 5, 1]


def test_u64_profile_decrypts():
    S = H.u64_dbfv()
    rng = np.random.default_rng(101)
    s = H.gen_secret_key(S.bfv, rng); rlk = H.gen_relin_key(S.bfv, s, rng)
    for a, b in [(3, 7), (100, 100)]:                                  # dbfv/eval.rs:345-382
        c1 = H.encrypt_sk(S.bfv, H.encode_scalar(S.bfv, a), s, rng); c2 = H.encrypt_sk(S.bfv, H.encode_scalar(S.bfv, b), s, rng)
        assert int(H.decrypt(S.bfv, O.bfv_mul_and_relin(S.bfv, c1, c2, rlk), s)[0]) == a * b
    for v in [0, 255, 256, 2 ** 64 - 1]:                               # dbfv/eval.rs:315-327
        assert H.dbfv_decrypt(S, H.dbfv_encrypt_sk(S, v, s, rng), s) == v
    ca, cb = H.dbfv_encrypt_sk(S, 1000, s, rng), H.dbfv_encrypt_sk(S, 2000, s, rng)   # dbfv/eval.rs:548-564
    out = O.dbfv_mul(S.bfv, S.base, S.d, S.plain_modulus, ca, cb, rlk, threads=O.max_threads())
    assert H.dbfv_decrypt(S, out, s) == 2_000_000


# ---- error pins (dbfv/eval.rs:385-453) ------------------------------------------------------------------------
def test_error_pins():
    n = 4096
    z = np.zeros((2, n), np.uint64)
    P1 = O.OracleParams(n=n, q=18014398509506561, aux=(36028797018972161,), plain_modulus=1040407, gadget_base=256)
    with pytest.raises(O.OracleError, match="single aux prime too small"):
        O.bfv_mul_and_relin(P1, z, z, np.zeros((P1.gadget_digits, 2, n), np.uint64))
    P0 = O.OracleParams(n=n, q=18014398509506561, aux=(), plain_modulus=1040407, gadget_base=256)
    with pytest.raises(O.OracleError, match="schoolbook BFV multiplication can overflow i128") as ei:
        O.bfv_mul_and_relin(P0, z, z, np.zeros((P0.gadget_digits, 2, n), np.uint64))
    assert ei.value.kind == "NotImplemented"
    P3 = O.OracleParams(n=16, q=1099509805057, aux=(562949953443841, 18014398509998081 - 8192 * 0 + 0, 36028797018972161)[:1] * 3,
                        plain_modulus=257)
    with pytest.raises(O.OracleError, match="HPS scaling supports 1 or 2 aux primes"):
        O.bfv_mul_and_relin(P3, np.zeros((2, 16), np.uint64), np.zeros((2, 16), np.uint64),
                            np.zeros((P3.gadget_digits, 2, 16), np.uint64))


def test_schoolbook_branch_small_params():
    """bfv/eval.rs:416-454 on parameters where i128 does not overflow (n=16, q=65537)."""
    P = O.OracleParams(n=16, q=65537, aux=(), plain_modulus=17, gadget_base=16)
    rng = np.random.default_rng(5)
    ct1, ct2 = rng.integers(0, P.q, (2, 16), dtype=np.uint64), rng.integers(0, P.q, (2, 16), dtype=np.uint64)
    got = O.ntt_inv(O.bfv_mul_no_relin(P, O.ntt_fwd(ct1, P.q), O.ntt_fwd(ct2, P.q)), P.q)
    assert np.array_equal(got, np.array(D.bfv_mul_no_relin_coeff(ct1, ct2, P.q, 17), dtype=np.uint64))


# ---- Paterson-Stockmeyer evaluation (bootstrap/digit_extract.rs:100-157), SURVEY row f-2 ---------------
def test_eval_poly_homomorphic_oracle_decrypts():
    """f(x) = 1 + 2x + 3x^2 + x^3 at x = 5 (mod 257) on HPS parameters with a deep noise budget."""
    P = O.OracleParams(n=64, q=1152921504606830593, aux=(18014398509998081, 36028797018972161),
                       plain_modulus=257, gadget_base=256)
    rng = np.random.default_rng(7)
    s = H.gen_secret_key(P, rng); rlk = H.gen_relin_key(P, s, rng)
    ct = H.encrypt_sk(P, H.encode_scalar(P, 5), s, rng)
    out = H.eval_poly_homomorphic(P, ct, [1, 2, 3, 1], rlk)
    assert int(H.decrypt(P, out, s)[0]) == (1 + 10 + 75 + 125) % 257
    assert int(H.decrypt(P, H.eval_poly_homomorphic(P, ct, [42], rlk), s)[0]) == 42
    assert int(H.decrypt(P, H.trivial_encrypt(P, 100), s)[0]) == 100               # digit_extract.rs:271-288


# ---- Galois automorphism + key switch ------------------------------------------------
def test_apply_automorphism_signed_permutation():
    """bfv/keygen.rs:218-239: X^i -> X^(ik) with X^n = -1; composition sigma_a(sigma_b) = sigma_ab."""
    q, n = 65537, 16
    x = np.zeros(n, np.uint64); x[1] = 5
    y = O.apply_automorphism(x, q, 3)
    assert y[3] == 5 and y.sum() == 5
    x = np.zeros(n, np.uint64); x[7] = 2                          # 7*3 = 21 = n + 5 -> -2 X^5
    y = O.apply_automorphism(x, q, 3)
    assert y[5] == q - 2 and np.count_nonzero(y) == 1
    rng = np.random.default_rng(3)
    x = rng.integers(0, q, n, dtype=np.uint64)
    assert np.array_equal(O.apply_automorphism(O.apply_automorphism(x, q, 3), q, 5), O.apply_automorphism(x, q, 15))
    # ring homomorphism: sigma(a*b) = sigma(a)*sigma(b)
    a, b = rng.integers(0, q, n, dtype=np.uint64), rng.integers(0, q, n, dtype=np.uint64)
    assert np.array_equal(O.apply_automorphism(O.poly_mul_naive(a, b, q), q, 7),
                          O.poly_mul_naive(O.apply_automorphism(a, q, 7), O.apply_automorphism(b, q, 7), q))


def test_bfv_apply_automorphism_decrypts():
    """bfv/eval.rs:929-976: sigma_3 keeps a scalar, maps 1 + 2X to 1 + 2X^3."""
    P = H.compact_bfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(P, rng)
    gk = H.gen_galois_key(P, s, 3, rng)
    ct = H.encrypt_sk(P, H.encode_scalar(P, 10), s, rng)
    assert int(H.decrypt(P, O.bfv_apply_automorphism(P, ct, gk, 3), s)[0]) == 10
    pt = np.zeros(P.n, np.uint64); pt[:2] = [1, 2]
    dec = H.decrypt(P, O.bfv_apply_automorphism(P, H.encrypt_sk(P, pt, s, rng), gk, 3), s)
    assert dec[:4].tolist() == [1, 0, 0, 2] and np.count_nonzero(dec) == 2


def test_dbfv_apply_automorphism_scalar_decrypts():
    """dbfv/advanced.rs:182-193: every limb through sigma_3, value 42 preserved."""
    S = H.compact_dbfv()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(S.bfv, rng)
    gk = H.gen_galois_key(S.bfv, s, 3, rng)
    ct = H.dbfv_encrypt_sk(S, 42, s, rng)
    assert H.dbfv_decrypt(S, O.bfv_apply_automorphism(S.bfv, ct, gk, 3), s) == 42


# ---- bootstrap pipeline (bootstrap/bfv_host.rs tests :345-560), SURVEY row f-1 / BASELINE config 5 --------
def _boot_params():
    orig = O.OracleParams(n=16, q=65537, aux=(), plain_modulus=5)                          # :351-357
    boot = O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=29, gadget_base=8)   # :359-365
    return orig, boot, 25


def _dbfv_boot_params():
    orig = H.DbfvSetup(O.OracleParams(n=16, q=65537, aux=(), plain_modulus=97, gadget_base=8), 4, 2, 16)   # :372-380
    boot = O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=257, gadget_base=8)              # :382-388
    return orig, boot, 64


def expected_ring_bootstrap(orig, boot, qp, ct, s_ntt):
    """Clear-text model of bfv_bootstrap's ring path: mod-switch, phase mod t_boot, rounding function."""
    q, t, tb, n = orig.q, orig.plain_modulus, boot.plain_modulus, orig.n
    sw = lambda v: ((qp * int(v) + q // 2) // q) % qp % tb
    c0 = [sw(v) for v in O.ntt_inv(ct[0], q)]
    c1 = [sw(v) for v in O.ntt_inv(ct[1], q)]
    sc = [int(v) if int(v) <= q // 2 else int(v) - q for v in O.ntt_inv(s_ntt, q)]
    out = []
    for j in range(n):
        acc = c0[j]
        for i in range(n):
            k = j - i
            acc += c1[i] * sc[k] if k >= 0 else -c1[i] * sc[k + n]
        x = acc % tb
        out.append(((t * (x % qp) + qp // 2) // qp) % t)
    return np.array(out, dtype=np.uint64)


def test_rounding_poly_interpolates():
    """digit_extract.rs:19-29: f(x) = round(t x' / q') mod t on x' = x mod q' for every x in Z_{t_boot}."""
    from oracle import bootstrap_ref as B
    poly = B.compute_rounding_poly(5, 25, 29)
    assert len(poly) == 29
    for x in range(29):
        val = sum(c * pow(x, k, 29) for k, c in enumerate(poly)) % 29
        assert val == ((5 * (x % 25) + 12) // 25) % 5


def test_coeffs_to_slots_roundtrip_oracle():
    """coeffs_to_slots.rs tests: every coefficient extracted as a constant, then packed back."""
    from oracle import bootstrap_ref as B
    _, boot, _ = _boot_params()
    rng = np.random.default_rng(7)
    s = H.gen_secret_key(boot, rng)
    gks = B.gen_trace_galois_keys(boot, s, rng)
    assert sorted(gks) == list(range(3, 32, 2))
    pt = rng.integers(0, boot.plain_modulus, boot.n, dtype=np.uint64)
    ct = H.encrypt_sk(boot, pt, s, rng)
    slots = B.coeffs_to_slots(boot, ct, gks)
    for j, sl in enumerate(slots):
        dec = H.decrypt(boot, sl, s)
        assert int(dec[0]) == int(pt[j]) and not dec[1:].any()
    assert np.array_equal(H.decrypt(boot, B.slots_to_coeffs(boot, slots), s), pt)


def test_bootstrap_single_and_ring_oracle():
    """bfv_host.rs:398-453: trivial ciphertexts m = 0..4 (fast path) and a real encryption of 3 (ring path)."""
    from oracle import bootstrap_ref as B
    orig, boot, qp = _boot_params()
    rng = np.random.default_rng(42)
    s = H.gen_secret_key(orig, rng)
    bk = B.gen_bootstrap_key(orig, boot, s, qp, orig.plain_modulus, rng)
    boot_sk = B.create_boot_sk(orig, boot, s)
    for m in range(5):
        out = B.bfv_bootstrap(orig, H.trivial_encrypt(orig, m), bk)
        assert int(H.decrypt(boot, out, boot_sk)[0]) % 5 == m
    # Ring path (:428-453).  With the reference's toy sizes the degree-28 rounding polynomial needs ~8
    # multiplicative levels (~7.5 bits each, measured) on a 44-bit budget, and the integer phase
    # c0' + c1' s leaves [0, t_boot): the reference's `== 3` holds for its ChaCha20 stream only (sampling is
    # outside this path).  The pipeline itself is pinned on a parameter set that fits the budget, against a
    # clear-text model of every step (mod-switch, phase mod t_boot, rounding function), all n coefficients.
    orig2 = O.OracleParams(n=16, q=65537, aux=(), plain_modulus=2)
    boot2 = O.OracleParams(n=16, q=1125899906842817, aux=(), plain_modulus=5, gadget_base=8)
    s2 = H.gen_secret_key(orig2, rng)
    bk2 = B.gen_bootstrap_key(orig2, boot2, s2, 4, 2, rng)
    ct = H.encrypt_sk(orig2, rng.integers(0, 2, 16, dtype=np.uint64), s2, rng)
    out = B.bfv_bootstrap(orig2, ct, bk2)
    assert np.array_equal(H.decrypt(boot2, out, B.create_boot_sk(orig2, boot2, s2)),
                          expected_ring_bootstrap(orig2, boot2, 4, ct, s2))


def test_dbfv_mul_then_bootstrap_and_chain_oracle():
    """bfv_host.rs:455-560: metadata contract (params swapped to the boot set), a second multiplication
    under boot key material, decryptability under the boot key."""
    from oracle import bootstrap_ref as B
    S, boot, qp = _dbfv_boot_params()
    rng = np.random.default_rng(777)
    s = H.gen_secret_key(S.bfv, rng)
    rlk = H.gen_relin_key(S.bfv, s, rng)
    bk = B.gen_bootstrap_key(S.bfv, boot, s, qp, S.bfv.plain_modulus, rng)
    boot_sk = B.create_boot_sk(S.bfv, boot, s)
    pa = np.zeros(16, np.uint64); pa[:2] = [3, 1]
    pb = np.zeros(16, np.uint64); pb[0] = 2
    ca, cb = H.dbfv_encrypt_poly_sk(S, pa, s, rng), H.dbfv_encrypt_poly_sk(S, pb, s, rng)
    S2, refreshed = B.dbfv_mul_then_bootstrap(S, ca, cb, rlk, bk)
    assert S2.bfv.plain_modulus == 257 and refreshed.shape == (2, 2, 16)
    c3 = H.dbfv_encrypt_poly_sk(S2, pb, boot_sk, rng)
    nxt = O.dbfv_mul(S2.bfv, S2.base, S2.d, S2.plain_modulus, refreshed, c3, bk.boot_rlk)
    assert H.dbfv_decrypt_poly(S2, nxt, boot_sk).shape == (16,)
    mk = lambda m: H.dbfv_encrypt_poly_sk(S, np.array([m % 16] + [0] * 15, np.uint64), s, rng)
    S3, chained = B.dbfv_mul_chain_then_bootstrap([(S, mk(3)), (S, mk(2)), (S, mk(5))], rlk, bk)
    assert S3.bfv.plain_modulus == 257 and H.dbfv_decrypt_poly(S3, chained, boot_sk).shape == (16,)


def test_multi_prime_oracle_reference_kat():
    """oracle/rns_ref.py pinned on the reference's own multi-prime test (bfv/eval.rs:903-927): n = 16,
    ct_moduli [65537, 1099509805057], p = 257, gadget base 8 -- 3*7, 10*20 and 0*5 decrypt to 21, 200, 0 after
    bfv_mul_and_relin (and already after bfv_mul_no_relin with the three-component decrypt)."""
    from oracle import rns_ref as R
    P = R.RnsParams(16, (65537, 1099509805057), 257, 8)
    assert P.G == 19                                               # compute_gadget_digits: 8^19 >= Q
    rng = np.random.default_rng(1234)
    s = R.gen_secret_key(P, rng)
    rlk = R.gen_relin_key(P, s, rng)
    for a, b, want in [(3, 7, 21), (10, 20, 200), (0, 5, 0)]:
        c1 = R.encrypt_sk(P, [a] + [0] * 15, s, rng)
        c2 = R.encrypt_sk(P, [b] + [0] * 15, s, rng)
        assert R.decrypt(P, c1, s)[0] == a
        assert R.decrypt(P, R.bfv_mul_no_relin(P, c1, c2), s) == [want] + [0] * 15
        assert R.decrypt(P, R.bfv_mul_and_relin(P, c1, c2, rlk), s) == [want] + [0] * 15
    # exact negacyclic product: Kronecker substitution == the reference's O(n^2) loop (bfv/eval.rs:792-810)
    a = [int(x) for x in rng.integers(-10**6, 10**6, 16)]
    b = [int(x) for x in rng.integers(-10**6, 10**6, 16)]
    ref = [0] * 16
    for i in range(16):
        for j in range(16):
            if i + j < 16:
                ref[i + j] += a[i] * b[j]
            else:
                ref[i + j - 16] -= a[i] * b[j]
    assert ref == R.negacyclic_mul(a, b, 16)
    # scale_tensor_component_bigint rounds half away from zero on the magnitude (bfv/eval.rs:818-831)
    Q = P.Q
    assert R.scale_component(P, [Q, -Q, 0, (Q + 1) // 2]) == [257, -257, 0, 129]
    # gadget KAT of the reference (bfv/keyswitch.rs:112-116) through the i128 restatement
    assert [d[0] for d in R.gadget_decompose([42], 65537, 16, 2)] == [65531, 3]
