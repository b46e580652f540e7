"""GPU tier: the reference's property tests (tests/protocol_props.rs:55-155) restated on the product API --
key generation, encryption, evaluation and decryption all run through libexacto_b200.so; only the sampling
comes from the test harness (numpy) through the sampler protocol."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st

import exacto_b200 as E
from common import H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CASES = settings(max_examples=12, deadline=None, derandomize=True)


class NpSampler:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def ternary(self, n, q): return H.sample_ternary(n, q, self.rng)
    def uniform(self, n, q): return H.sample_uniform(n, q, self.rng)
    def gaussian(self, n, q, sigma): return H.sample_gaussian(n, q, sigma, self.rng)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(native_lib):
    assert torch.cuda.is_available(), "the gpu tier needs a CUDA device"


def _enc(pt, sk, params, smp):
    q, n = params.ct_basis.moduli[0], params.ring_degree
    return E.encrypt_sk_with_samples(pt, sk, params, E.CoeffPoly(smp.uniform(n, q), q),
                                     E.CoeffPoly(smp.gaussian(n, q, params.sigma), q))


def _denc(value_or_poly, sk, dparams, smp, poly=False):
    P = dparams.bfv_params
    q, n = P.ct_basis.moduli[0], P.ring_degree
    a = [E.CoeffPoly(smp.uniform(n, q), q) for _ in range(dparams.num_digits)]
    e = [E.CoeffPoly(smp.gaussian(n, q, P.sigma), q) for _ in range(dparams.num_digits)]
    fn = E.dbfv_encrypt_poly_sk_with_samples if poly else E.dbfv_encrypt_sk_with_samples
    return fn(value_or_poly, sk, dparams, a, e)


@CASES
@given(m=st.integers(0, 256), seed=st.integers(0, 2**32))
def test_prop_bfv_roundtrip_scalar(m, seed):                                    # protocol_props.rs:55-65
    params, smp = E.compact_bfv(), NpSampler(seed)
    sk = E.gen_secret_key_with_sampler(params, smp)
    ct = _enc(E.encode_scalar(m, params), sk, params, smp)
    assert E.decode_scalar(E.decrypt(ct, sk)) == m


@CASES
@given(a=st.integers(0, 63), b=st.integers(0, 63), seed=st.integers(0, 2**32))
def test_prop_bfv_mul_scalar(a, b, seed):                                       # :67-81
    params, smp = E.compact_bfv(), NpSampler(seed)
    sk = E.gen_secret_key_with_sampler(params, smp)
    rlk = E.gen_relin_key_with_sampler(sk, smp)
    prod = E.bfv_mul_and_relin(_enc(E.encode_scalar(a, params), sk, params, smp),
                               _enc(E.encode_scalar(b, params), sk, params, smp), rlk)
    assert E.decode_scalar(E.decrypt(prod, sk)) == (a * b) % params.plain_modulus


@CASES
@given(a=st.integers(0, 255), b=st.integers(0, 255), seed=st.integers(0, 2**32))
def test_prop_dbfv_add_mul(a, b, seed):                                         # :83-102
    dp, smp = E.compact_dbfv(), NpSampler(seed)
    p = dp.plain_modulus
    sk = E.gen_secret_key_with_sampler(dp.bfv_params, smp)
    rlk = E.gen_relin_key_with_sampler(sk, smp)
    ca, cb = _denc(a, sk, dp, smp), _denc(b, sk, dp, smp)
    assert E.dbfv_decrypt(E.dbfv_add(ca, cb), sk) == (a + b) % p
    assert E.dbfv_decrypt(E.dbfv_mul(ca, cb, rlk), sk) == (a * b) % p


def _terms_to_poly(terms, n, p):
    c = np.zeros(n, np.uint64)
    for idx, v in terms:
        c[idx] = (int(c[idx]) + v) % p
    return c


def _sparse_negacyclic_mul(a_terms, b_terms, n, p):
    out = [0] * n
    for i, x in a_terms:
        for j, y in b_terms:
            k = i + j
            if k < n:
                out[k] = (out[k] + x * y) % p
            else:
                out[k - n] = (out[k - n] - x * y) % p
    return out


terms = lambda hi, cnt: st.lists(st.tuples(st.integers(0, 1023), st.integers(0, hi)), min_size=0, max_size=cnt)


@CASES
@given(a_terms=terms(15, 5), b_terms=terms(15, 5), seed=st.integers(0, 2**32))
def test_prop_dbfv_poly_add_sparse(a_terms, b_terms, seed):                     # :108-130
    dp, smp = E.compact_dbfv(), NpSampler(seed)
    p, n = dp.plain_modulus, dp.bfv_params.ring_degree
    sk = E.gen_secret_key_with_sampler(dp.bfv_params, smp)
    pa, pb = _terms_to_poly(a_terms, n, p), _terms_to_poly(b_terms, n, p)
    ca = _denc(E.CoeffPoly(pa, p), sk, dp, smp, poly=True)
    cb = _denc(E.CoeffPoly(pb, p), sk, dp, smp, poly=True)
    got = E.dbfv_decrypt_poly(E.dbfv_add(ca, cb), sk).coeffs
    assert np.array_equal(got, (pa + pb) % np.uint64(p))


@CASES
@given(a_terms=terms(7, 4), b_terms=terms(7, 4), seed=st.integers(0, 2**32))
def test_prop_dbfv_poly_mul_sparse(a_terms, b_terms, seed):                     # :132-155
    dp, smp = E.compact_dbfv(), NpSampler(seed)
    p, n = dp.plain_modulus, dp.bfv_params.ring_degree
    sk = E.gen_secret_key_with_sampler(dp.bfv_params, smp)
    rlk = E.gen_relin_key_with_sampler(sk, smp)
    ca = _denc(E.CoeffPoly(_terms_to_poly(a_terms, n, p), p), sk, dp, smp, poly=True)
    cb = _denc(E.CoeffPoly(_terms_to_poly(b_terms, n, p), p), sk, dp, smp, poly=True)
    # the reference multiplies the term lists (duplicates accumulate before the product)
    agg = lambda ts: [(i, int(v)) for i, v in enumerate(_terms_to_poly(ts, n, p)) if v]
    got = E.dbfv_decrypt_poly(E.dbfv_mul(ca, cb, rlk), sk).coeffs
    assert [int(v) for v in got] == _sparse_negacyclic_mul(agg(a_terms), agg(b_terms), n, p)


def _binary(self, n, q):
    return self.rng.integers(0, 2, n).astype(np.uint64)


NpSampler.binary = _binary


@CASES
@given(value=st.integers(0, 255), seed=st.integers(0, 2**32))
def test_prop_public_key_roundtrip(value, seed):            # bfv/encrypt.rs tests + dbfv/decrypt.rs:110-124
    dp, smp = E.compact_dbfv(), NpSampler(seed)
    P = dp.bfv_params
    sk = E.gen_secret_key_with_sampler(P, smp)
    pk = E.gen_public_key_with_sampler(sk, smp)
    m = value % P.plain_modulus
    assert E.decode_scalar(E.decrypt(E.encrypt_pk_with_sampler(E.encode_scalar(m, P), pk, P, smp), sk)) == m
    assert E.dbfv_decrypt(E.dbfv_encrypt_with_sampler(value, pk, dp, smp), sk) == value
    assert E.dbfv_decrypt(E.dbfv_encrypt_sk_with_sampler(value, sk, dp, smp), sk) == value


def test_dbfv_div_by_base_and_change_base():
    """dbfv/advanced.rs:195-221: 48 / 16 = 3 with the plaintext modulus divided by the base; values survive a
    change of base 16 -> 4 (4 digits)."""
    dp = E.compact_dbfv()                                   # base 16, p = 256
    smp = NpSampler(43)
    sk = E.gen_secret_key_with_sampler(dp.bfv_params, smp)
    ct = E.dbfv_encrypt_sk_with_sampler(48, sk, dp, smp)
    div = E.dbfv_div_by_base(ct)
    assert E.dbfv_decrypt(div, sk) == 3 and div.params.plain_modulus == 16
    for value in [0, 1, 15, 42, 127, 255]:
        ct = E.dbfv_encrypt_sk_with_sampler(value, sk, dp, smp)
        b4 = E.dbfv_change_base(ct, 4, 4)
        assert E.dbfv_decrypt(b4, sk) == value and b4.params.base == 4 and b4.num_limbs() == 4
    with pytest.raises(E.ExactoError, match="new base must be >= 2"):
        E.dbfv_change_base(ct, 1, 4)
