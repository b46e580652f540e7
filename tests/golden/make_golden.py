"""Generates tests/golden/ctmul_golden.npz with the CPU oracle (oracle/exacto_oracle.c).

The reference is Rust and cannot run in this image, and it ships no ciphertext-level
golden vectors, so these fixtures are *oracle outputs* on seeded inputs,
cross-checked at generation time against the independent big-int definition oracle
(oracle/definition.py).  Small cases store inputs and outputs in full; the full-size
cases store SHA-256 digests of inputs and outputs (inputs are regenerated from the seed
and the digest guards against RNG drift).

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O                      # noqa: E402
from oracle import definition as D      # noqa: E402
from oracle import harness as H         # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctmul_golden.npz")

# name -> (OracleParams, dbfv base, d, dbfv plain modulus, seed, store_full)
CASES = {
    "toy16_a1": (H.toy(16), 16, 2, 256, 11, True),
    "n64_a2_rep": (O.OracleParams(n=64, q=1152921504606830593, aux=(18014398509998081, 36028797018972161),
                                  plain_modulus=1040407, gadget_base=256), 16, 2, 250, 12, True),
    "n32_a2_base7": (O.OracleParams(n=32, q=576460752308273153, aux=(18014398509998081, 36028797018972161),
                                    plain_modulus=65537, gadget_base=7), 3, 3, 20, 13, True),
    "compact_dbfv": (H.compact_dbfv().bfv, 16, 2, 256, 14, False),
    "cfg3p_dbfv": (H.cfg3_prime().bfv, 256, 2, 65536, 15, False),
    "u64_dbfv": (H.u64_dbfv().bfv, 256, 8, 0, 16, False),
}


def inputs(P, d, seed):
    rng = np.random.default_rng(seed)
    ct1 = rng.integers(0, P.q, (d, 2, P.n), dtype=np.uint64)
    ct2 = rng.integers(0, P.q, (d, 2, P.n), dtype=np.uint64)
    rlk = rng.integers(0, P.q, (P.gadget_digits, 2, P.n), dtype=np.uint64)
    # edge values the reference's centring / rounding branches care about
    edge = np.array([0, 1, P.q // 2, P.q // 2 + 1, P.q - 1], dtype=np.uint64)
    ct1[0, 0, :5] = edge
    ct2[0, 1, :5] = edge[::-1]
    return ct1, ct2, rlk


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return h.hexdigest()


def main():
    store = {}
    for name, (P, base, d, pm, seed, full) in CASES.items():
        ct1, ct2, rlk = inputs(P, d, seed)
        out = O.dbfv_mul(P, base, d, pm, ct1, ct2, rlk, threads=O.max_threads())
        bfv = O.bfv_mul_and_relin(P, ct1[0], ct2[0], rlk)
        if full:  # cross-check against the definition oracle in the coefficient domain
            want = D.dbfv_mul_coeff(O.ntt_inv(ct1, P.q), O.ntt_inv(ct2, P.q), O.ntt_inv(rlk, P.q), P.q,
                                    P.plain_modulus, P.gadget_base, P.gadget_digits, base, d, pm)
            assert np.array_equal(O.ntt_inv(out, P.q), np.array(want, dtype=np.uint64)), name
            store[f"{name}/ct1"], store[f"{name}/ct2"], store[f"{name}/rlk"] = ct1, ct2, rlk
            store[f"{name}/dbfv_out"], store[f"{name}/bfv_out"] = out, bfv
        store[f"{name}/in_sha256"] = np.array(digest(ct1, ct2, rlk))
        store[f"{name}/dbfv_sha256"] = np.array(digest(out))
        store[f"{name}/bfv_sha256"] = np.array(digest(bfv))
        print(name, "ok", store[f"{name}/dbfv_sha256"])
    # NTT vectors: forward transform of a fixed ramp for every config prime
    for n, q in [(16, 65537), (1024, 1099509805057), (1024, 562949953443841), (4096, 1152921504606830593),
                 (4096, 18014398509998081), (4096, 36028797018972161), (4096, 576460752308273153)]:
        x = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(12345)) % np.uint64(q)
        store[f"ntt/{n}_{q}"] = np.array(digest(O.ntt_fwd(x, q)))
        store[f"psi/{n}_{q}"] = np.array(O.find_psi(n, q), dtype=np.uint64)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
