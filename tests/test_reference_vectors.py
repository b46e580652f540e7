"""Reference-produced golden vectors (tests/golden/gen_reference_vectors.rs, run with cargo in the reference crate).

When `tests/golden/reference_vectors/<case>.*.u64` exist, the oracle (CPU tier) and the CUDA path (GPU tier) must
reproduce the reference's coefficient-domain outputs word for word; that turns the oracle's "restatement, pinned at
KAT level" into "pinned at ciphertext level".  Without the files the reference-vector tests skip; the format itself
is always exercised on a vector written by a Python twin of the Rust generator (same SplitMix64 stream)."""
import glob
import os

import numpy as np
import pytest

from common import H, O

HERE = os.path.dirname(os.path.abspath(__file__))
VEC_DIR = os.path.join(HERE, "golden", "reference_vectors")


def load_case(prefix):
    rd = lambda what: np.fromfile(f"{prefix}.{what}.u64", dtype="<u8")
    meta = [int(x) for x in rd("meta")]
    n, q, na = meta[0], meta[1], meta[2]
    aux = tuple(meta[3:3 + na])
    p, gb, G, base, d, pm = meta[3 + na:3 + na + 6]
    P = O.OracleParams(n=n, q=q, aux=aux, plain_modulus=p, gadget_base=gb, gadget_digits=G)
    return dict(P=P, base=base, d=d, pm=pm, in1=rd("in1").reshape(d, 2, n), in2=rd("in2").reshape(d, 2, n),
                rlk=rd("rlk").reshape(G, 2, n), out=rd("out").reshape(d, 2, n))


def splitmix_polys(seed, count, n, q):
    """The generator's stream: `count` polynomials of n words, each next() % q."""
    mask = (1 << 64) - 1
    out = np.zeros((count, n), np.uint64)
    s = seed
    for i in range(count):
        for j in range(n):
            s = (s + 0x9E3779B97F4A7C15) & mask
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
            out[i, j] = (z ^ (z >> 31)) % q
    return out


def write_case(prefix, P, base, d, pm, seed):
    """Python twin of run_case() in gen_reference_vectors.rs, with the oracle in the reference's place."""
    G = P.gadget_digits
    polys = splitmix_polys(seed, 4 * d + 2 * G, P.n, P.q)
    in1, in2, rlk = polys[:2 * d].reshape(d, 2, P.n), polys[2 * d:4 * d].reshape(d, 2, P.n), polys[4 * d:].reshape(G, 2, P.n)
    out = oracle_mul(dict(P=P, base=base, d=d, pm=pm, in1=in1, in2=in2, rlk=rlk))
    meta = [P.n, P.q, len(P.aux), *P.aux, P.plain_modulus, P.gadget_base, G, base, d, pm]
    for what, arr in (("meta", np.array(meta, "<u8")), ("in1", in1), ("in2", in2), ("rlk", rlk), ("out", out)):
        np.ascontiguousarray(arr, "<u8").tofile(f"{prefix}.{what}.u64")


def oracle_mul(c):
    P = c["P"]
    f = lambda a: O.ntt_fwd(a.reshape(-1, P.n), P.q).reshape(a.shape)
    prod = O.dbfv_mul(P, c["base"], c["d"], c["pm"], f(c["in1"]), f(c["in2"]), f(c["rlk"]))
    return O.ntt_inv(prod.reshape(-1, P.n), P.q).reshape(prod.shape)


def gpu_mul(c):
    import exacto_b200 as E
    from common import to_dbfv_params
    P = c["P"]
    dp = to_dbfv_params(P, c["base"], c["d"], c["pm"])
    plan = E.make_plan(P.n, P.q)
    f = lambda a: np.stack([E.NttPoly.from_coeff_poly(E.CoeffPoly(x, P.q), plan).evals for x in a.reshape(-1, P.n)]).reshape(a.shape)
    rlk = E.RelinKey(f(c["rlk"]), dp.bfv_params)
    prod = E.dbfv_mul_batch(dp, f(c["in1"])[None], f(c["in2"])[None], rlk)[0]
    return np.stack([E.NttPoly(x, P.q, plan).to_coeff_poly().coeffs for x in prod.reshape(-1, P.n)]).reshape(prod.shape)


def reference_cases():
    return sorted(p[:-len(".meta.u64")] for p in glob.glob(os.path.join(VEC_DIR, "*.meta.u64")))


def test_vector_format_round_trip(tmp_path):
    """The file format and the SplitMix64 stream of the Rust generator, on a small case the oracle finishes at once."""
    P = H.toy(16)
    P = O.OracleParams(n=P.n, q=P.q, aux=P.aux, plain_modulus=P.plain_modulus, gadget_base=P.gadget_base,
                       gadget_digits=P.gadget_digits)
    prefix = str(tmp_path / "toy")
    write_case(prefix, P, 16, 2, 256, 0xE8AC7002)
    c = load_case(prefix)
    assert c["P"] == P and (c["base"], c["d"], c["pm"]) == (16, 2, 256)
    assert c["in1"].shape == (2, 2, 16) and int(c["in1"].max()) < P.q
    assert np.array_equal(oracle_mul(c), c["out"])
    # first word of the stream: SplitMix64(seed).next() % q
    z = (0xE8AC7002 + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
    assert int(c["in1"][0, 0, 0]) == (z ^ (z >> 31)) % P.q


@pytest.mark.parametrize("prefix", reference_cases() or [None])
def test_oracle_matches_reference_vectors(prefix):
    if prefix is None:
        pytest.skip("no reference-produced vectors in tests/golden/reference_vectors (see gen_reference_vectors.rs)")
    c = load_case(prefix)
    assert np.array_equal(oracle_mul(c), c["out"]), os.path.basename(prefix)


@pytest.mark.gpu
@pytest.mark.parametrize("prefix", reference_cases() or [None])
def test_gpu_matches_reference_vectors(prefix, tmp_path):
    if prefix is None:                      # still run the whole loader path through the GPU on a generated case
        P = H.compact_dbfv().bfv
        prefix = str(tmp_path / "compact")
        write_case(prefix, P, 16, 2, 256, 0xE8AC7002)
    c = load_case(prefix)
    assert np.array_equal(gpu_mul(c), c["out"]), os.path.basename(prefix)
