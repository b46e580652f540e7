"""GPU tier: differential fuzz over parameter sets -- ring degree, prime sizes (every lazy-reduction class of the
64-bit transforms), 0 / 1 / 2 auxiliary primes, plaintext modulus, power-of-two and general gadget bases, dBFV
digit counts -- against the literal oracle: dbfv_mul (host and device entry points), automorphism + key switch,
decrypt.  Deterministic (seeded); the parameter sets are derived, not hand-picked."""
import os

import numpy as np
import pytest

import exacto_b200 as E
from common import H, O, to_dbfv_params, to_params

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(native_lib):
    assert torch.cuda.is_available(), "the gpu tier needs a CUDA device"


def ntt_prime(bits: int, n: int, skip=()):
    """Largest prime < 2^bits with p = 1 mod 2n that is not in `skip`."""
    step = 2 * n
    c = ((1 << bits) - 1) // step * step + 1
    while c > (1 << (bits - 1)):
        if c not in skip and O.is_prime(c):
            return c
        c -= step
    raise AssertionError("no prime")


def overflow_risk(p, q, n):
    """bfv/eval.rs:457-464 with the factor 2 of the middle tensor term c0*d1 + c1*d0 that the reference's guard
    leaves out: in the band n*(q/2)^2*p <= i128::MAX < 2*n*(q/2)^2*p its schoolbook branch wraps on worst-case
    inputs (this fuzz found it: case 201 at EXB_FUZZ_CASES=300), so the device refuses those sets too."""
    mc = q // 2
    mt = 2 * n * mc * mc
    return mt > (1 << 127) - 1 or mt * p > (1 << 127) - 1


def make_cases():
    rng = np.random.default_rng(20241018)
    cases = []
    qbits = [20, 31, 36, 40, 47, 52, 55, 57, 59, 60, 61, 62]
    for idx in range(int(os.environ.get("EXB_FUZZ_CASES", "36"))):      # EXB_FUZZ_CASES=300 for a longer soak
        logn = int(rng.integers(4, 14))
        if idx % 6 == 0:
            logn = 12                                    # the tuned path gets its share
        n = 1 << logn
        qb = max(qbits[idx % len(qbits)], logn + 3)
        q = ntt_prime(qb, n)
        p = int(rng.integers(2, min(q, 1 << int(rng.integers(2, 21)))))
        A = idx % 3
        aux = ()
        if A == 1:
            need = (n * q) // 2 + 1
            ab = max(need.bit_length() + 1, 30)
            if ab > 62:
                A = 2
            else:
                aux = (ntt_prime(ab, n, skip=(q,)),)
        if A == 2:
            a0 = ntt_prime(int(rng.integers(50, 62)), n, skip=(q,))
            aux = (a0, ntt_prime(int(rng.integers(50, 62)), n, skip=(q, a0)))
        if A == 0 and overflow_risk(p, q, n):
            a0 = ntt_prime(61, n, skip=(q,))
            aux = (a0, ntt_prime(60, n, skip=(q, a0)))
        gb = [16, 256, 1 << 16, 10, 1000, 1 << 20][idx % 6]
        d = [1, 2, 3, 2][idx % 4]
        b = [4, 16, 256][idx % 3]
        pm = max(2, b ** d - 3) if idx % 5 == 0 else b ** d        # b^d - 3: non-zero small representatives (reduce)
        cases.append((idx, O.OracleParams(n=n, q=q, aux=aux, plain_modulus=p, gadget_base=gb), b, d, pm))
    return cases


CASES = make_cases()


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-n{c[1].n}-q{c[1].q.bit_length()}-A{len(c[1].aux)}-B{c[1].gadget_base}-d{c[3]}")
def test_fuzz_parameter_sets(case):
    from exacto_b200 import batch
    idx, P, b, d, pm = case
    dp = to_dbfv_params(P, b, d, pm)
    params = dp.bfv_params
    q, n = P.q, P.n
    rng = np.random.default_rng(idx)
    B = 3
    ct1 = rng.integers(0, q, (B, d, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, q, (B, d, 2, n), dtype=np.uint64)
    edge = np.full(n, q // 2, np.uint64); edge[::3] = q // 2 + 1; edge[1::5] = q - 1; edge[2::7] = 0
    ct1[0] = O.ntt_fwd(edge, q); ct2[0] = O.ntt_fwd(edge[::-1].copy(), q)
    karr = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    rlk = E.RelinKey(karr, params)
    want = np.stack([O.dbfv_mul(P, b, d, pm, x, y, karr, threads=O.max_threads()) for x, y in zip(ct1, ct2)])
    assert np.array_equal(E.dbfv_mul_batch(dp, ct1, ct2, rlk), want)
    assert np.array_equal(batch.to_host(batch.dbfv_mul(dp, batch.to_device(ct1), batch.to_device(ct2), rlk)), want)
    # the two halves on their own (limb 0 of each pair)
    want3 = np.stack([O.bfv_mul_no_relin(P, x, y) for x, y in zip(ct1[:, 0], ct2[:, 0])])
    assert np.array_equal(E.bfv_mul_no_relin_batch(params, ct1[:, 0], ct2[:, 0]), want3)
    assert np.array_equal(E.relinearize_batch(params, want3, rlk), O.relinearize(P, want3, karr))
    # automorphism + key switch on the limbs of the first operand, decrypt of products under a random "key"
    k = int(rng.integers(1, n)) * 2 + 1
    flat = ct1.reshape(B * d, 2, n)
    assert np.array_equal(E.bfv_apply_automorphism_batch(params, flat, E.GaloisKey(karr, k, params)),
                          O.bfv_apply_automorphism(P, flat, karr, k, threads=O.max_threads()))
    s = rng.integers(0, q, n, dtype=np.uint64)
    got = E.decrypt_batch(params, flat[:2], E.SecretKey.from_ntt(s, params))
    assert np.array_equal(got, np.stack([H.decrypt(P, c, s) for c in flat[:2]]))


def test_schoolbook_middle_term_band_is_refused():
    """n = 64, 60-bit q, p = 8, no aux: the reference's overflow guard passes (n*(q/2)^2*p < 2^127) but its middle
    tensor term needs 128 bits; the literal oracle wraps there while the exact result is what the device would
    compute -- no defined reference result, so the parameter set is refused like the overflow-risk ones."""
    P = O.OracleParams(n=64, q=1152921504606844417, aux=(), plain_modulus=8, gadget_base=10)
    params = to_params(P)
    z = np.zeros((1, 2, 64), np.uint64)
    with pytest.raises(E.ExactoError, match="overflows i128 in its middle tensor term"):
        E.bfv_mul_and_relin_batch(params, z, z, E.RelinKey(np.zeros((P.gadget_digits, 2, 64), np.uint64), params))
    aux = (ntt_prime(61, 64, skip=(P.q,)), ntt_prime(60, 64, skip=(P.q,)))
    P2 = O.OracleParams(n=64, q=P.q, aux=aux, plain_modulus=8, gadget_base=10)       # with an aux basis it multiplies
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, P.q, (1, 2, 64), dtype=np.uint64), rng.integers(0, P.q, (1, 2, 64), dtype=np.uint64)
    k = rng.integers(0, P.q, (P2.gadget_digits, 2, 64), dtype=np.uint64)
    p2 = to_params(P2)
    assert np.array_equal(E.bfv_mul_and_relin_batch(p2, a, b, E.RelinKey(k, p2)), O.bfv_mul_and_relin(P2, a, b, k))


@pytest.mark.parametrize("qb", [31, 36, 37, 40, 47, 52, 55, 56, 57, 59, 60, 61, 62])
@pytest.mark.parametrize("A", [0, 2])
def test_fuzz_n4096_prime_sizes(qb, A):
    """The tuned n = 4096 kernels across every size class of q: below 2^36 (generic 64-bit aux path), 2^36 .. 2^60
    (internal 27-bit basis; lazy classes 2 and 1 of the 64-bit transforms), above (exact Harvey class), with two
    auxiliary primes and with none (internal basis / refused when the schoolbook branch could overflow)."""
    from exacto_b200 import batch
    n = 4096
    q = ntt_prime(qb, n)
    rng = np.random.default_rng(qb * 3 + A)
    p = int(rng.integers(2, 1 << 16))
    if A == 0 and overflow_risk(p, q, n):
        params = to_params(O.OracleParams(n=n, q=q, aux=(), plain_modulus=p, gadget_base=256))
        z = np.zeros((1, 2, n), np.uint64)
        with pytest.raises(E.ExactoError, match="schoolbook BFV multiplication"):
            E.bfv_mul_and_relin_batch(params, z, z, E.RelinKey(np.zeros((params.gadget_digits, 2, n), np.uint64), params))
        return
    aux = ()
    if A == 2:
        a0 = ntt_prime(int(rng.integers(52, 62)), n, skip=(q,))
        aux = (a0, ntt_prime(int(rng.integers(52, 62)), n, skip=(q, a0)))
    P = O.OracleParams(n=n, q=q, aux=aux, plain_modulus=p, gadget_base=[256, 1 << 16, 1000][qb % 3])
    d, b = 3, 16
    dp = to_dbfv_params(P, b, d, b ** d)
    B = 2
    ct1 = rng.integers(0, q, (B, d, 2, n), dtype=np.uint64)
    ct2 = rng.integers(0, q, (B, d, 2, n), dtype=np.uint64)
    half = O.ntt_fwd(np.full(n, q // 2, np.uint64), q)
    ct1[0] = half; ct2[0] = half
    karr = rng.integers(0, q, (P.gadget_digits, 2, n), dtype=np.uint64)
    rlk = E.RelinKey(karr, dp.bfv_params)
    want = np.stack([O.dbfv_mul(P, b, d, b ** d, x, y, karr, threads=O.max_threads()) for x, y in zip(ct1, ct2)])
    assert np.array_equal(E.dbfv_mul_batch(dp, ct1, ct2, rlk), want)
    big1, big2 = np.tile(ct1, (40, 1, 1, 1)), np.tile(ct2, (40, 1, 1, 1))             # 80 pairs: the per-limb kernel
    got = batch.to_host(batch.dbfv_mul(dp, batch.to_device(big1), batch.to_device(big2), rlk))
    assert np.array_equal(got[:2], want) and np.array_equal(got[-2:], want)
