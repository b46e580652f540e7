"""Shared helpers of the test-suite (oracle-side; never imported by the product)."""
import ctypes
import hashlib
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle as O                      # noqa: E402
from oracle import harness as H         # noqa: E402

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "ctmul_golden.npz")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES
golden_inputs = make_golden.inputs
digest = make_golden.digest


def golden():
    return np.load(GOLDEN_PATH)


def to_params(P: O.OracleParams):
    """OracleParams -> exacto_b200.BfvParams (same numbers through the reference-style builder)."""
    import exacto_b200 as E
    b = (E.BfvParamsBuilder().ring_degree(P.n).plain_modulus(P.plain_modulus).ct_moduli([P.q])
         .aux_moduli(list(P.aux)).gadget_base(P.gadget_base))
    return b.build()


def to_dbfv_params(P: O.OracleParams, base, d, pm):
    import exacto_b200 as E
    return E.DbfvParams.new(to_params(P), base, d, pm)


class Emulator:
    """ctypes front-end of tests/host_emul/libexb_emul.so (the product kernels on CPU threads)."""

    def __init__(self):
        import __graft_entry__ as g
        L = ctypes.CDLL(g.build_emulator())
        u64, u32, vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
        L.emu_last_error.restype = ctypes.c_char_p
        L.emu_create.argtypes = [u32, vp, u32, vp, u32, u64, u64, u32, ctypes.POINTER(vp)]
        L.emu_destroy.argtypes = [vp]
        L.emu_info.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(u64)]
        L.emu_ntt.argtypes = [vp, u32, ctypes.c_int, vp, vp, ctypes.c_size_t]
        L.emu_poly_op.argtypes = [vp, u32, ctypes.c_int, vp, vp, u64, vp, ctypes.c_size_t]
        L.emu_small_primes.argtypes = [vp, vp]
        L.emu_dbfv_mul.argtypes = [vp, u64, u32, u64, vp, vp, vp, u32, vp, ctypes.c_size_t, u32, u32]
        L.emu_bfv_apply_automorphism.argtypes = [vp, vp, u64, vp, vp, ctypes.c_size_t]
        L.emu_bfv_decrypt.argtypes = [vp, vp, u32, vp, vp, ctypes.c_size_t]
        L.emu_tensor_per_limb.argtypes = [vp, u64, u32, u64, u32, u32]
        L.emu_bfv_mul_no_relin.argtypes = [vp, vp, vp, vp, ctypes.c_size_t]
        L.emu_bfv_relinearize.argtypes = [vp, vp, vp, u32, vp, ctypes.c_size_t, ctypes.c_int]
        L.emu_gadget_decompose.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.emu_rns_info.argtypes = [vp, ctypes.POINTER(u32), ctypes.POINTER(u32), vp]
        L.emu_rns_mul.argtypes = [vp, u64, u32, u64, vp, vp, vp, u32, vp, ctypes.c_size_t, ctypes.c_int]
        self.L = L

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def create(self, n, ct_moduli, aux, plain, gadget_base=0, gadget_digits=0):
        ct = np.array(ct_moduli, np.uint64)
        ax = np.array(list(aux) or [0], np.uint64)
        h = ctypes.c_void_p()
        rc = self.L.emu_create(n, self._p(ct), len(ct_moduli), self._p(ax), len(aux), plain, gadget_base,
                               gadget_digits, ctypes.byref(h))
        return rc, h, self.L.emu_last_error().decode()

    def from_oracle(self, P: O.OracleParams):
        rc, h, err = self.create(P.n, [P.q], P.aux, P.plain_modulus, P.gadget_base, P.gadget_digits)
        assert rc == 0, err
        return h

    def info(self, h):
        gb, gd, ms, psi = ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_int(), ctypes.c_uint64()
        self.L.emu_info(h, ctypes.byref(gb), ctypes.byref(gd), ctypes.byref(ms), ctypes.byref(psi))
        return gb.value, gd.value, ms.value, psi.value, self.L.emu_last_error().decode()

    def small_primes(self, h):
        out = np.zeros(8, np.uint64)
        k = self.L.emu_small_primes(h, self._p(out))
        return [int(v) for v in out[:k]]

    def ntt(self, h, base, forward, x):
        x = np.ascontiguousarray(x, np.uint64)
        out = np.zeros_like(x)
        rc = self.L.emu_ntt(h, base, 1 if forward else 0, self._p(x), self._p(out), x.size // x.shape[-1])
        assert rc == 0
        return out

    def poly_op(self, h, base, op, a, b=None, scalar=0):
        a = np.ascontiguousarray(a, np.uint64)
        bb = np.ascontiguousarray(b if b is not None else a, np.uint64)
        out = np.zeros_like(a)
        self.L.emu_poly_op(h, base, op, self._p(a), self._p(bb), scalar, self._p(out), a.size)
        return out

    def bfv_apply_automorphism(self, h, ct, element, gk):
        ct, gk = np.ascontiguousarray(ct, np.uint64), np.ascontiguousarray(gk, np.uint64)
        out = np.zeros_like(ct)
        self.L.emu_bfv_apply_automorphism(h, self._p(ct), element, self._p(gk), self._p(out), ct.size // (2 * ct.shape[-1]))
        return out

    def bfv_mul_no_relin(self, h, ct1, ct2):
        ct1, ct2 = np.ascontiguousarray(ct1, np.uint64), np.ascontiguousarray(ct2, np.uint64)
        out = np.zeros(ct1.shape[:-2] + (3, ct1.shape[-1]), np.uint64)
        rc = self.L.emu_bfv_mul_no_relin(h, self._p(ct1), self._p(ct2), self._p(out), ct1.size // (2 * ct1.shape[-1]))
        return rc, out, self.L.emu_last_error().decode()

    def bfv_relinearize(self, h, ct3, rlk, wide=False):
        ct3, rlk = np.ascontiguousarray(ct3, np.uint64), np.ascontiguousarray(rlk, np.uint64)
        out = np.zeros(ct3.shape[:-2] + (2, ct3.shape[-1]), np.uint64)
        rc = self.L.emu_bfv_relinearize(h, self._p(ct3), self._p(rlk), rlk.shape[0], self._p(out),
                                        ct3.size // (3 * ct3.shape[-1]), 1 if wide else 0)
        return rc, out

    def gadget_decompose(self, h, coeffs, G):
        coeffs = np.ascontiguousarray(coeffs, np.uint64)
        out = np.zeros(coeffs.shape[:-1] + (G, coeffs.shape[-1]), np.uint64)
        self.L.emu_gadget_decompose(h, self._p(coeffs), self._p(out), coeffs.size // coeffs.shape[-1])
        return out

    def tensor_per_limb(self, h, base, d, pm, flags=0, limb_mask=0):
        return self.L.emu_tensor_per_limb(h, base, d, pm, flags, limb_mask)

    def bfv_decrypt(self, h, ct, sk_ntt):
        ct, sk_ntt = np.ascontiguousarray(ct, np.uint64), np.ascontiguousarray(sk_ntt, np.uint64)
        out = np.zeros(ct.shape[:-2] + ct.shape[-1:], np.uint64)
        self.L.emu_bfv_decrypt(h, self._p(ct), ct.shape[-2], self._p(sk_ntt), self._p(out), out.size // ct.shape[-1])
        return out

    def rns_info(self, h):
        L_, K_ = ctypes.c_uint32(), ctypes.c_uint32()
        primes = np.zeros(8, np.uint64)
        rc = self.L.emu_rns_info(h, ctypes.byref(L_), ctypes.byref(K_), self._p(primes))
        return rc, L_.value, K_.value, [int(v) for v in primes[:K_.value]], self.L.emu_last_error().decode()

    def rns_mul(self, h, base, d, pm, ct1, ct2, rlk, mode, out_shape):
        """Multi-prime path (rns_kernels.cu): mode 0 dbfv_mul, 1 bfv_mul_no_relin, 2 relinearize(ct1)."""
        ct1, ct2, rlk = (np.ascontiguousarray(v, np.uint64) for v in (ct1, ct2, rlk))
        out = np.zeros(out_shape, np.uint64)
        rc = self.L.emu_rns_mul(h, base, d, pm, self._p(ct1), self._p(ct2), self._p(rlk), rlk.shape[0], self._p(out),
                                out_shape[0], mode)
        return rc, out, self.L.emu_last_error().decode()

    def dbfv_mul(self, h, base, d, pm, ct1, ct2, rlk, flags=0, limb_mask=0, out=None):
        ct1, ct2, rlk = (np.ascontiguousarray(v, np.uint64) for v in (ct1, ct2, rlk))
        pairs = ct1.size // (d * 2 * ct1.shape[-1])
        out = np.zeros_like(ct1) if out is None else out
        rc = self.L.emu_dbfv_mul(h, base, d, pm, self._p(ct1), self._p(ct2), self._p(rlk), rlk.shape[0],
                                 self._p(out), pairs, flags, limb_mask)
        return rc, out, self.L.emu_last_error().decode()


# ---- multi-prime ciphertext modulus (oracle/rns_ref.py) ----------------------------------------------------
from oracle import rns_ref as R            # noqa: E402

RNS_CASES = {
    # the reference's own multi-prime test set (bfv/eval.rs:903-927): Q = 65537 * 1099509805057 < 2^64
    "ref_n16": (R.RnsParams(16, (65537, 1099509805057), 257, 8), 1, 2, 2),
    # two 60-bit primes: Q ~ 2^119 > 2^64, so relinearize runs on the reference's truncated (u64) reconstruction
    "n64_two60": (R.RnsParams(64, (1152921504606830593, 576460752308273153), 65537, 1 << 16), 1, 2, 2),
    # three 40-bit primes wrapped as dBFV d = 2, b = 16, p = 256
    "n32_three40_d2": (R.RnsParams(32, (1099509805057, 1099510054913, 1099507695617), 257, 1 << 20), 2, 16, 256),
}


def rns_inputs(P, d, pairs, seed):
    """Seeded multi-prime inputs: ct [pairs][d][2][L][n] x2 and rlk [G][2][L][n], with edge patterns in pair 0."""
    rng = np.random.default_rng(seed)

    def rnd(prefix):
        return np.stack([rng.integers(0, q, prefix + (P.n,), dtype=np.uint64) for q in P.moduli], axis=-2)
    ct1, ct2, rlk = rnd((pairs, d, 2)), rnd((pairs, d, 2)), rnd((P.G, 2))
    ct1[0, 0, 0, :, :4] = 0
    for l, q in enumerate(P.moduli):
        ct1[0, 0, 1, l, :4] = q - 1
        ct2[0, 0, 0, l, :4] = q // 2
        ct2[0, 0, 1, l, :4] = q // 2 + 1
    return ct1, ct2, rlk
