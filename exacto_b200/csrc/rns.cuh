// rns.cuh -- multi-prime ciphertext modulus (ct_basis with L > 1 primes): per-coefficient arithmetic of
//   bfv_mul_generic_rns           bfv/eval.rs:113-147   (exact CRT -> integer tensor -> round(p t / Q) -> RNS)
//   relinearize for L > 1         bfv/keyswitch.rs:59-101 + RnsPoly::to_coeff_poly ring/rns.rs:114-151
//
// The reference does the tensor with BigInt schoolbook products (O(n^2)).  Here the tensor is computed in an
// EXTENDED RNS basis E = e_1 .. e_K of 61-bit NTT primes with prod(E) > 4 n Q^2 >= 4 max|t|, so every tensor
// coefficient t is represented exactly; per coefficient it is then recovered as an integer (Garner mixed radix,
// centred), scaled with one multiword division r = sign(t) * floor((p |t| + floor(Q/2)) / Q) -- literally
// scale_tensor_component_bigint (bfv/eval.rs:818-831) -- and reduced mod every q_l.  All integer, bit-exact.
//
// Multiword integers are little-endian arrays of 32-bit limbs with compile-time capacities.
// Everything is __host__ __device__ so tests/host_emul replays it on the CPU.
#pragma once
#include "modarith.cuh"

namespace exb {

constexpr int kRnsMaxL = 4;                 // ciphertext primes
constexpr int kRnsMaxK = 6;                 // extended-basis primes
constexpr int kMwQ = 4;                     // limbs of Q  (Q < 2^127)
constexpr int kMwE = 2 * kRnsMaxK;          // limbs of prod(E) (K primes below 2^62)
constexpr int kMwNum = kMwE + 3;            // limbs of p * |t| + floor(Q/2)

struct RnsConsts {
    u32 L, K, n, logn;
    u32 gadget_digits, gadget_log2;
    u64 gadget_base, plain;
    Modulus q[kRnsMaxL];                    // ciphertext primes
    Modulus e[kRnsMaxK];                    // extended basis
    u32 Q[kMwQ], halfQ[kMwQ];               // Q = prod q_l, floor(Q / 2)
    u32 Qlen, Elen;
    u32 Qstar[kRnsMaxL][kMwQ];              // Q / q_l
    u64 crt_inv[kRnsMaxL], crt_inv_s[kRnsMaxL];     // (Q / q_l)^-1 mod q_l  (+ Shoup companion)
    u64 c32_q[kRnsMaxL], c32_q_s[kRnsMaxL];         // 2^32 mod q_l
    u64 c32_e[kRnsMaxK], c32_e_s[kRnsMaxK];         // 2^32 mod e_k
    u64 garner[kRnsMaxK][kRnsMaxK], garner_s[kRnsMaxK][kRnsMaxK];   // [k][j], j < k: e_j^-1 mod e_k
    u32 E[kMwE], halfE[kMwE];               // prod(E), floor(prod(E) / 2)
    // relinearize, the reference's u128 arithmetic (ring/rns.rs:133-150): big_q and Q / q_l as u128
    u64 bigq_lo, bigq_hi;
    u64 qstar_lo[kRnsMaxL], qstar_hi[kRnsMaxL];
};

// ---- multiword helpers ---------------------------------------------------------------------------------
EXB_HD int mw_cmp(const u32 *a, const u32 *b, int len) {
    for (int i = len - 1; i >= 0; i--)
        if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
    return 0;
}
// a -= b (a >= b)
EXB_HD void mw_sub(u32 *a, const u32 *b, int len) {
    u64 borrow = 0;
    for (int i = 0; i < len; i++) {
        const u64 d = (u64)a[i] - b[i] - borrow;
        a[i] = (u32)d;
        borrow = (d >> 32) & 1u;
    }
}
// a = b - a (b >= a)
EXB_HD void mw_rsub(u32 *a, const u32 *b, int len) {
    u64 borrow = 0;
    for (int i = 0; i < len; i++) {
        const u64 d = (u64)b[i] - a[i] - borrow;
        a[i] = (u32)d;
        borrow = (d >> 32) & 1u;
    }
}
// a = a * m + add over `len` limbs (the caller sizes len so that nothing is lost)
EXB_HD void mw_mul_add64(u32 *a, int len, u64 m, u64 add) {
    const u64 m0 = (u32)m, m1 = m >> 32;
    u64 c0 = (u32)add, c1 = add >> 32;        // carries into limb i and limb i + 1
    u32 prev = 0;                              // a[i-1] before the update (for the m1 cross term)
    for (int i = 0; i < len; i++) {
        const u32 ai = a[i];
        // new a[i] = low32( ai * m0 + prev * m1 + c0 ), with carries propagated in 64-bit pieces
        const u64 t0 = (u64)ai * m0;
        const u64 t1 = (u64)prev * m1;
        const u64 s = (t0 & 0xffffffffu) + (t1 & 0xffffffffu) + (c0 & 0xffffffffu);
        a[i] = (u32)s;
        const u64 carry = (s >> 32) + (t0 >> 32) + (t1 >> 32) + (c0 >> 32) + c1;
        c0 = carry;
        c1 = 0;
        prev = ai;
    }
}
// x mod m for a multiword x: Horner from the top limb with 2^32 mod m as a Shoup constant (m < 2^62)
EXB_HD u64 mw_mod64(const u32 *x, int len, u64 m, u64 c32, u64 c32_s) {
    u64 r = 0;
    for (int i = len - 1; i >= 0; i--) {
        r = shoup(r, c32, c32_s, m) + x[i];              // < m + 2^32
        if (r >= m) { r -= m; if (r >= m) r %= m; }      // one subtraction unless m is a small prime
    }
    return r;
}

EXB_HD int clz32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// Knuth algorithm D on 32-bit limbs: q = floor(u / v).  u has ulen limbs, v has vlen limbs with v[vlen-1] != 0,
// ulen >= vlen; q receives ulen - vlen + 1 limbs.  u is destroyed.  UCAP bounds ulen (+1 for the normalised copy).
template <int UCAP, int VCAP>
EXB_HD void mw_div(const u32 *u_in, int ulen, const u32 *v_in, int vlen, u32 *q) {
    if (vlen == 1) {
        u64 rem = 0;
        const u32 v0 = v_in[0];
        for (int j = ulen - 1; j >= 0; j--) {
            const u64 cur = (rem << 32) | u_in[j];
            q[j] = (u32)(cur / v0);
            rem = cur % v0;
        }
        return;
    }
    u32 un[UCAP + 1], vn[VCAP];
    const int s = clz32(v_in[vlen - 1]);
    for (int i = vlen - 1; i > 0; i--) vn[i] = s ? ((v_in[i] << s) | (v_in[i - 1] >> (32 - s))) : v_in[i];
    vn[0] = v_in[0] << s;
    un[ulen] = s ? (u_in[ulen - 1] >> (32 - s)) : 0u;
    for (int i = ulen - 1; i > 0; i--) un[i] = s ? ((u_in[i] << s) | (u_in[i - 1] >> (32 - s))) : u_in[i];
    un[0] = u_in[0] << s;
    for (int j = ulen - vlen; j >= 0; j--) {
        const u64 num = ((u64)un[j + vlen] << 32) | un[j + vlen - 1];
        u64 qhat = num / vn[vlen - 1];
        u64 rhat = num % vn[vlen - 1];
        while (qhat >= ((u64)1 << 32) || qhat * vn[vlen - 2] > ((rhat << 32) | un[j + vlen - 2])) {
            qhat--;
            rhat += vn[vlen - 1];
            if (rhat >= ((u64)1 << 32)) break;
        }
        // multiply and subtract
        i64 borrow = 0;
        u64 carry = 0;
        for (int i = 0; i < vlen; i++) {
            const u64 p = qhat * vn[i] + carry;
            carry = p >> 32;
            const i64 t = (i64)un[i + j] - borrow - (i64)(p & 0xffffffffu);
            un[i + j] = (u32)t;
            borrow = t < 0 ? 1 : 0;
        }
        const i64 t = (i64)un[j + vlen] - borrow - (i64)carry;
        un[j + vlen] = (u32)t;
        if (t < 0) {                          // qhat was one too large: add back
            qhat--;
            u64 c = 0;
            for (int i = 0; i < vlen; i++) {
                const u64 sum = (u64)un[i + j] + vn[i] + c;
                un[i + j] = (u32)sum;
                c = sum >> 32;
            }
            un[j + vlen] += (u32)c;
        }
        q[j] = (u32)qhat;
    }
}

// ---- reconstruct_centered_bigint (bfv/eval.rs:719-760) per coefficient ------------------------------------
// x[l] = coefficient residue mod q_l (canonical).  Writes |c| into mag[0..Qlen) and returns the sign (true: c < 0),
// c = the centred representative of the CRT value (c > floor(Q/2) => c - Q, strict like the reference).
EXB_HD bool rns_centered(const u64 *x, const RnsConsts &R, u32 *mag) {
    const int QL = (int)R.Qlen;
    u32 acc[kMwQ + 1];
    for (int i = 0; i <= QL; i++) acc[i] = 0;
    for (u32 l = 0; l < R.L; l++) {
        const u64 t = shoup(x[l], R.crt_inv[l], R.crt_inv_s[l], R.q[l].m);       // x_l * (Q/q_l)^-1 mod q_l
        // acc += t * Qstar_l   (each term < Q, so acc < L * Q fits QL + 1 limbs)
        const u64 t0 = (u32)t, t1 = t >> 32;
        u64 carry = 0;
        u32 prev = 0;
        for (int i = 0; i <= QL; i++) {
            const u32 si = i < QL ? R.Qstar[l][i] : 0u;
            const u64 a = (u64)si * t0, b = (u64)prev * t1;
            const u64 s = (a & 0xffffffffu) + (b & 0xffffffffu) + (carry & 0xffffffffu) + acc[i];
            acc[i] = (u32)s;
            carry = (s >> 32) + (a >> 32) + (b >> 32) + (carry >> 32);
            prev = si;
        }
    }
    u32 Qx[kMwQ + 1];
    for (int i = 0; i < QL; i++) Qx[i] = R.Q[i];
    Qx[QL] = 0;
    while (mw_cmp(acc, Qx, QL + 1) >= 0) mw_sub(acc, Qx, QL + 1);                // % Q
    bool neg = false;
    u32 hq[kMwQ + 1];
    for (int i = 0; i < QL; i++) hq[i] = R.halfQ[i];
    hq[QL] = 0;
    if (mw_cmp(acc, hq, QL + 1) > 0) { mw_rsub(acc, Qx, QL + 1); neg = true; }   // c - Q, magnitude Q - c
    for (int i = 0; i < QL; i++) mag[i] = acc[i];
    return neg;
}

// Residue of the centred value mod extended prime e_k (canonical).
EXB_HD u64 rns_mag_mod_e(const u32 *mag, bool neg, const RnsConsts &R, u32 k) {
    const u64 m = R.e[k].m;
    const u64 r = mw_mod64(mag, (int)R.Qlen, m, R.c32_e[k], R.c32_e_s[k]);
    return (neg && r) ? m - r : r;
}

// ---- exact tensor coefficient -> round(p t / Q) mod every q_l ---------------------------------------------
// tr[k] = t mod e_k (canonical).  out[l] = r mod q_l with r = scale_tensor_component_bigint(t)
// (bfv/eval.rs:818-831) and the reduction of centered_bigint_to_rns (:762-790).
EXB_HD void rns_scale_coeff(const u64 *tr, const RnsConsts &R, u64 *out) {
    const int K = (int)R.K, EL = (int)R.Elen, QL = (int)R.Qlen;
    // Garner: mixed-radix digits v_k with t = v_0 + v_1 e_0 + v_2 e_0 e_1 + ...  (t taken in [0, prod E))
    u64 v[kRnsMaxK];
    for (int k = 0; k < K; k++) {
        const u64 m = R.e[k].m;
        u64 u = tr[k];
        for (int j = 0; j < k; j++) {
            const u64 vj = v[j] >= m ? v[j] % m : v[j];
            u = shoup(mod_sub(u, vj, m), R.garner[k][j], R.garner_s[k][j], m);
        }
        v[k] = u;
    }
    u32 num[kMwNum];
    for (int i = 0; i < kMwNum; i++) num[i] = 0;
    // Horner: t = (((v_{K-1}) e_{K-2} + v_{K-2}) e_{K-3} + ...) e_0 + v_0
    num[0] = (u32)v[K - 1]; num[1] = (u32)(v[K - 1] >> 32);
    for (int k = K - 2; k >= 0; k--) mw_mul_add64(num, EL, R.e[k].m, v[k]);
    bool neg = false;
    if (mw_cmp(num, R.halfE, EL) > 0) { mw_rsub(num, R.E, EL); neg = true; }      // centre against prod(E)
    // num = p * |t| + floor(Q / 2)
    mw_mul_add64(num, kMwNum, R.plain, 0);
    {
        u64 carry = 0;
        for (int i = 0; i < kMwNum; i++) {
            const u64 s = (u64)num[i] + (i < QL ? R.halfQ[i] : 0u) + carry;
            num[i] = (u32)s;
            carry = s >> 32;
        }
    }
    int ulen = kMwNum;
    while (ulen > QL && num[ulen - 1] == 0) ulen--;
    u32 quo[kMwNum];
    for (int i = 0; i < kMwNum; i++) quo[i] = 0;
    if (ulen >= QL) mw_div<kMwNum, kMwQ>(num, ulen, R.Q, QL, quo);
    const int qlen = ulen >= QL ? ulen - QL + 1 : 0;
    for (u32 l = 0; l < R.L; l++) {
        const u64 m = R.q[l].m;
        const u64 r = mw_mod64(quo, qlen, m, R.c32_q[l], R.c32_q_s[l]);
        out[l] = (neg && r) ? m - r : r;
    }
}

// ---- decrypt for L > 1 (bfv/encrypt.rs:111-178): x = CRT value in [0, Q) of the phase coefficient,
// m = floor((x p + floor(Q/2)) / Q) mod p.  x[l] = phase residue mod q_l.
EXB_HD u64 rns_decrypt_coeff(const u64 *x, const RnsConsts &R) {
    u32 mag[kMwQ];
    const bool neg = rns_centered(x, R, mag);
    const int QL = (int)R.Qlen;
    u32 num[kMwQ + 3];
    for (int i = 0; i < kMwQ + 3; i++) num[i] = i < QL ? mag[i] : 0u;
    if (neg) mw_rsub(num, R.Q, QL);                       // centred value c = x - Q  =>  x = Q - |c|
    mw_mul_add64(num, kMwQ + 3, R.plain, 0);
    u64 carry = 0;
    for (int i = 0; i < kMwQ + 3; i++) {
        const u64 sum = (u64)num[i] + (i < QL ? R.halfQ[i] : 0u) + carry;
        num[i] = (u32)sum;
        carry = sum >> 32;
    }
    int ulen = kMwQ + 3;
    while (ulen > QL && num[ulen - 1] == 0) ulen--;
    u32 quo[kMwQ + 3];
    for (int i = 0; i < kMwQ + 3; i++) quo[i] = 0;
    mw_div<kMwQ + 3, kMwQ>(num, ulen, R.Q, QL, quo);
    unsigned __int128 r = 0;                              // quotient <= p < 2^64 (+1): two limbs and a bit
    for (int i = ulen - QL; i >= 0; i--) r = ((r << 32) | quo[i]) % R.plain;
    return (u64)r;
}

// ---- relinearize for L > 1: RnsPoly::to_coeff_poly (ring/rns.rs:133-150) then gadget_decompose
// (bfv/keyswitch.rs:11-52) on the truncated (coefficient, modulus) pair, i128 semantics ------------------
typedef unsigned __int128 exb_u128;
typedef __int128 exb_i128;

// x[l] = residue of c2 mod q_l.  Returns `val as u64`.
EXB_HD u64 rns_to_coeff_truncated(const u64 *x, const RnsConsts &R) {
    const exb_u128 bigq = ((exb_u128)R.bigq_hi << 64) | R.bigq_lo;
    exb_u128 val = 0;
    for (u32 l = 0; l < R.L; l++) {
        const u64 t = shoup(x[l], R.crt_inv[l], R.crt_inv_s[l], R.q[l].m);       // mod_mul(c, q_star_inv, q_l)
        const exb_u128 qs = ((exb_u128)R.qstar_hi[l] << 64) | R.qstar_lo[l];
        val = val + (exb_u128)t * qs;                                             // < 2 Q < 2^128
        if (val >= bigq) val -= bigq;                                             // % big_q
    }
    return (u64)val;
}

// Digit g of the running decomposition; `remaining` is updated like bfv/keyswitch.rs:33-42.  qm = big_q as u64.
EXB_HD u64 rns_gadget_digit(exb_i128 &remaining, u64 base, u64 qm) {
    const exb_i128 b = (exb_i128)base, hb = b / 2, qi = (exb_i128)qm;
    exb_i128 rem = remaining % b;
    if (rem < -hb) rem += b;
    else if (rem >= hb) rem -= b;
    const exb_i128 rq = ((rem % qi) + qi) % qi;
    remaining = (remaining - rem) / b;
    return (u64)rq;
}

}  // namespace exb
