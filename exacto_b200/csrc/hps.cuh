// hps.cuh -- per-coefficient steps of the HPS RNS multiplication and of the
// gadget decomposition, bit-exact restatements of
//   base_extend_centered  bfv/eval.rs:217-247   -> ext_centered()
//   hps_scale             bfv/eval.rs:257-413   -> hps_scale_coeff()
//   gadget_decompose      bfv/keyswitch.rs:11-52 -> gadget_digit_*()
// redesigned so no 128-bit division is needed: every `%`/`/` of the reference
// becomes a Shoup / Barrett step with host-precomputed constants (context.cpp),
// and every result is the same canonical residue the reference produces.
#pragma once
#include "modarith.cuh"

namespace exb {

struct ScaleConsts {
    u64 q, half_q;       // ciphertext prime, floor(q/2)
    u32 num_aux;         // A in {1, 2}
    u32 pad_;
    u64 pj[2];           // aux primes
    u64 pj_mu[2];        // floor(2^64 / p_j)
    // A = 1: K[0] = q^-1 mod P.          A = 2: K[j] = q^-1 * p_{1-j}^-1 mod p_j
    u64 K[2], K_s[2];
    u64 other[2];        // A = 2: other[0] = p1, other[1] = p0 (CRT cofactors)
    u128w big_p;         // P = prod p_j
    u128w half_big_p;    // floor(P / 2)
    u64 plain, plain_s;  // BFV plaintext modulus p (< q) and its Shoup companion mod q
    // A = 1: C[0] = p mod q.             A = 2: C[0] = p*p1 mod q, C[1] = p*p0 mod q
    u64 C[2], C_s[2];
    u64 CP;              // p * P mod q
    u64 CP2;             // 2 * p * P mod q
};

// The centred extension rule (bfv/eval.rs:230-240 and :307-314 / :357-375):
// value c in [0,q) is read as c - q when c > floor(q/2), then reduced to [0, pj).
EXB_HD u64 ext_centered(u64 c, u64 q, u64 half_q, u64 pj, u64 pj_mu) {
    const bool neg = c > half_q;
    const u64 v = neg ? q - c : c;
    const u64 rem = barrett_reduce(v, pj, pj_mu);
    return (neg && rem != 0) ? pj - rem : rem;
}

// round(p * a_c / q) mod q with the reference's sign-symmetric rounding
// (bfv/eval.rs:323-328).  |a_c| * p = k*q + r via Shoup (k = mulhi(|a_c|, p_s)),
// r in [0, 2q), so floor((p|a_c| + floor(q/2)) / q) = k + [r+h >= q] + [r+h >= 2q].
EXB_HD u64 round_term(u64 a, const ScaleConsts &c) {
    const bool neg = a > c.half_q;
    const u64 av = neg ? c.q - a : a;
    const u64 k = mulhi64(av, c.plain_s);
    const u64 r = av * c.plain - k * c.q;
    const u64 rh = r + c.half_q;
    const u64 rr = k + (rh >= c.q ? 1u : 0u) + (rh >= 2 * c.q ? 1u : 0u);   // <= p/2 + 1 < q
    return neg ? mod_neg(rr, c.q) : rr;
}

// The same rounding term as a signed integer (|value| <= p/2 + 1), for sums over products.
EXB_HD i64 round_term_signed(u64 a, const ScaleConsts &c) {
    const bool neg = a > c.half_q;
    const u64 av = neg ? c.q - a : a;
    const u64 k = mulhi64(av, c.plain_s);
    const u64 r = av * c.plain - k * c.q;
    const u64 rh = r + c.half_q;
    const i64 rr = (i64)(k + (rh >= c.q ? 1u : 0u) + (rh >= 2 * c.q ? 1u : 0u));
    return neg ? -rr : rr;
}

// hps_scale, one coefficient.  a = t mod q, b0/b1 = t mod p_j (all canonical).
EXB_HD u64 hps_scale_coeff(u64 a, u64 b0, u64 b1, const ScaleConsts &c) {
    const u64 q = c.q;
    const u64 rnd = round_term(a, c);
    if (c.num_aux == 1) {                                       // :301-332
        const u64 P = c.pj[0];
        const u64 ae = ext_centered(a, q, c.half_q, P, c.pj_mu[0]);
        const u64 diff = mod_sub(b0, ae, P);
        const u64 m_raw = shoup(diff, c.K[0], c.K_s[0], P);
        u64 x = shoup(m_raw, c.C[0], c.C_s[0], q);              // p * m_raw mod q
        if (m_raw > c.half_big_p.lo) x = mod_sub(x, c.CP, q);    // centred: m_raw - P
        return mod_add(rnd, x, q);
    }
    // :349-404
    const u64 ae0 = ext_centered(a, q, c.half_q, c.pj[0], c.pj_mu[0]);
    const u64 ae1 = ext_centered(a, q, c.half_q, c.pj[1], c.pj_mu[1]);
    const u64 t0 = shoup(mod_sub(b0, ae0, c.pj[0]), c.K[0], c.K_s[0], c.pj[0]);
    const u64 t1 = shoup(mod_sub(b1, ae1, c.pj[1]), c.K[1], c.K_s[1], c.pj[1]);
    u128w sum = add128(mul_wide(t0, c.other[0]), mul_wide(t1, c.other[1]));   // < 2P
    u32 wraps = 0;
    if (ge128(sum, c.big_p)) { sum = sub128(sum, c.big_p); wraps = 1; }       // crt_sum % P
    if (gt128(sum, c.half_big_p)) wraps += 1;                                 // m_crt - P
    u64 x = mod_add(shoup(t0, c.C[0], c.C_s[0], q), shoup(t1, c.C[1], c.C_s[1], q), q);
    if (wraps == 1) x = mod_sub(x, c.CP, q);
    else if (wraps == 2) x = mod_sub(x, c.CP2, q);
    return mod_add(rnd, x, q);
}

// Balanced base-B digits of the centred coefficient (bfv/keyswitch.rs:24-44).
// Returns the signed digit in [-B/2, B/2) and updates `remaining`.
// Power-of-two base B = 2^w: rem = ((v + B/2) mod B) - B/2, remaining = (v - rem) >> w.
EXB_HD i64 gadget_digit_pow2(i64 &remaining, u32 w) {
    const i64 half = (i64)1 << (w - 1);
    const i64 mask = ((i64)1 << w) - 1;
    const i64 rem = ((remaining + half) & mask) - half;
    remaining = (remaining - rem) >> w;
    return rem;
}
// Base 2^8, all digits at once: the balanced-digit carry chain (bfv/keyswitch.rs:33-42) IS the carry chain of one
// 64-bit addition.  With W = v + 0x8080..80 (mod 2^64), byte g of W is digit_g + 128, so W ^ 0x8080..80 holds the
// eight digits as int8 (valid while |v| < 2^62, i.e. the remainder after eight digits is zero).
EXB_HD u64 gadget_digits_base256(i64 v) {
    const u64 bias = 0x8080808080808080ull;
    return ((u64)v + bias) ^ bias;
}
// __byte_perm (PRMT): byte i of the result is byte ((s >> 4i) & 7) of the 8-byte pool {x, y}.
EXB_HD u32 byte_perm(u32 x, u32 y, u32 s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, s);
#else
    const u64 pool = ((u64)y << 32) | x;
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (u32)((pool >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}
// 4 x 4 byte transpose: out[g] byte j = in[j] byte g.
EXB_HD void transpose4x4_bytes(const u32 *in, u32 *out) {
    const u32 t0 = byte_perm(in[0], in[1], 0x5140), t1 = byte_perm(in[0], in[1], 0x7362);
    const u32 t2 = byte_perm(in[2], in[3], 0x5140), t3 = byte_perm(in[2], in[3], 0x7362);
    out[0] = byte_perm(t0, t2, 0x5410); out[1] = byte_perm(t0, t2, 0x7632);
    out[2] = byte_perm(t1, t3, 0x5410); out[3] = byte_perm(t1, t3, 0x7632);
}
// Eight coefficients' digit words (byte g of w[j] = digit g of coefficient j) -> eight plane words
// (byte j of plane[g] = digit g of coefficient j): an 8 x 8 byte transpose as four 4 x 4 ones.
EXB_HD void digits_to_planes(const u64 *w, u64 *plane) {
    u32 a[4], b[4], c[4], d[4], ta[4], tb[4], tc[4], td[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; j++) {
        a[j] = (u32)w[j]; b[j] = (u32)w[4 + j];                 // digits 0..3 of coefficients 0..3 / 4..7
        c[j] = (u32)(w[j] >> 32); d[j] = (u32)(w[4 + j] >> 32); // digits 4..7
    }
    transpose4x4_bytes(a, ta); transpose4x4_bytes(b, tb); transpose4x4_bytes(c, tc); transpose4x4_bytes(d, td);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < 4; g++) {
        plane[g] = (u64)ta[g] | ((u64)tb[g] << 32);
        plane[4 + g] = (u64)tc[g] | ((u64)td[g] << 32);
    }
}
// General base: the reference's truncating % and / on signed values.
EXB_HD i64 gadget_digit_general(i64 &remaining, i64 base) {
    const i64 half = base / 2;
    i64 rem = remaining % base;
    if (rem < -half) rem += base;
    else if (rem >= half) rem -= base;
    remaining = (remaining - rem) / base;
    return rem;
}

EXB_HD i64 center_i64(u64 c, u64 q, u64 half_q) { return c > half_q ? (i64)c - (i64)q : (i64)c; }
EXB_HD u64 signed_to_mod(i64 v, u64 q) { return v < 0 ? q - (u64)(-v) : (u64)v; }

}  // namespace exb
