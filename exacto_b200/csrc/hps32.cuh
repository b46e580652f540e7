// hps32.cuh -- hps_scale (bfv/eval.rs:257-413) on the internal 27-bit auxiliary basis.
//
// With a = t mod q (canonical) and b_i = t mod p_i for the small primes p_i (P' = prod p_i),
//   M = (t - a_c) / q           (exact integer, |M| <= n*q/2 + 1  <<  P'/2)
//   out = round(p * a_c / q) + p * M   (mod q)            -- bfv/eval.rs:323-331 / :390-403
// M is recovered without big integers (HPS "exact alpha"):
//   m'_i = (b_i - [a_c]_{p_i}) * Kp_i  mod p_i,   Kp_i = (q * P'/p_i)^-1 mod p_i
//        (b_i * Kp_i comes out of the inverse transform: Kp_i is folded into its n^-1 scaling),
//   M    = sum_i m'_i * (P'/p_i)  -  alpha * P',   alpha = round(sum_i m'_i / p_i)
// and alpha is exact in 2^-57 fixed point (error < 3 * 2^27 * 2^-57 < 2^-28) because |M| / P' < 2^-8
// (checked on the host).
// The result equals the reference's for every input because the reference's own centred CRT of m
// cannot wrap for the parameter sets this path is enabled for (host_setup.cpp: small_basis_ok).
#pragma once
#include "hps.cuh"
#include "ntt32_core.cuh"

namespace exb {

constexpr int kMaxSmall = 4;

struct Scale32Consts {
    u32 K;                       // number of small primes
    u32 sh;                      // bitlen(q) - 2  (q >> sh in [2, 4))
    u32 rq;                      // floor(2^(32 + sh) / q): Barrett companion for the final reduction
    u32 plain32;                 // 1 when p and floor(p * 2^64 / q) both fit 32 bits
    Mod32 m[kMaxSmall];          // ninv* hold n^-1 * Kp_i: the inverse transforms deliver b_i * Kp_i
    u32 Kp[kMaxSmall], Kp_s[kMaxSmall];   // Kp_i = (q * P'/p_i)^-1 mod p_i  + Shoup32 companion
    u32 RK[kMaxSmall], RK_s[kMaxSmall];   // 2^32 * Kp_i mod p_i             + Shoup32 companion
    u32 QK[kMaxSmall];                    // q * Kp_i mod p_i
    u32 g[kMaxSmall];                     // floor(2^57 / p_i)  (p_i > 2^26)
    u64 C[kMaxSmall];                     // p * (P'/p_i) mod q
    u64 CPn[kMaxSmall + 1];               // q - (alpha * p * P' mod q), alpha = 0..K
};

// Centred value of a (mod q) reduced to [0, p_i): the ext rule of bfv/eval.rs:230-240.
EXB_HD u32 ext32_centered(u64 a, u64 q, u64 half_q, const Mod32 &m) {
    const bool neg = a > half_q;
    const u32 r = reduce64_to_p(neg ? q - a : a, m);
    return (neg && r != 0) ? m.p - r : r;
}

// The same value up to a multiple of p, in [0, 2p], with ONE 32-bit Shoup multiply: |a_c| < 2^59 is split at bit 28,
// top * (2^28 mod p) + low < 2p + 2^28 < 6p (p > 2^26).  Feeds the lazy forward transform (which accepts it).
EXB_HD u32 ext32_centered_lazy(u64 a, u64 q, u64 half_q, const Mod32 &m) {
    const bool neg = a > half_q;
    const u64 v = neg ? q - a : a;
    const u32 u = shoup32_lazy((u32)(v >> 28), m.c28, m.c28_s, m.p) + ((u32)v & 0x0fffffffu);
    const u32 r = csub32(csub32(u, m.four_p), m.two_p);                  // [0, 2p)
    return neg ? m.two_p - r : r;
}

// round_term (hps.cuh) when p and its Shoup companion fit 32 bits: same k and r, half the multiplies.
EXB_HD u64 round_term32(u64 a, const ScaleConsts &c) {
    const bool neg = a > c.half_q;
    const u64 av = neg ? c.q - a : a;
    const u32 a1 = (u32)(av >> 32), a0 = (u32)av, s = (u32)c.plain_s, p = (u32)c.plain;
    const u32 k = (u32)(((u64)a1 * s + (((u64)a0 * s) >> 32)) >> 32);        // floor(av * s / 2^64)
    const u64 avp = (u64)a0 * p + ((u64)(a1 * p) << 32);                      // av * p   mod 2^64
    const u64 kq = (u64)k * (u32)c.q + ((u64)(k * (u32)(c.q >> 32)) << 32);   // k * q    mod 2^64
    const u64 r = avp - kq;                                                   // [0, 2q)
    const u64 rh = r + c.half_q;
    const u64 rr = (u64)k + (rh >= c.q ? 1u : 0u) + (rh >= 2 * c.q ? 1u : 0u);
    return neg ? mod_neg(rr, c.q) : rr;
}

EXB_HD i64 round_term32_signed(u64 a, const ScaleConsts &c) {
    const bool neg = a > c.half_q;
    const u64 av = neg ? c.q - a : a;
    const u32 a1 = (u32)(av >> 32), a0 = (u32)av, s = (u32)c.plain_s, p = (u32)c.plain;
    const u32 k = (u32)(((u64)a1 * s + (((u64)a0 * s) >> 32)) >> 32);
    const u64 avp = (u64)a0 * p + ((u64)(a1 * p) << 32);
    const u64 kq = (u64)k * (u32)c.q + ((u64)(k * (u32)(c.q >> 32)) << 32);
    const u64 rh = avp - kq + c.half_q;
    const i64 rr = (i64)((u64)k + (rh >= c.q ? 1u : 0u) + (rh >= 2 * c.q ? 1u : 0u));
    return neg ? -rr : rr;
}

// Sum over the products of one output limb (tensor01_kernel):
//   rsum = sum_ij round(p a_ij / q) as a signed integer (|rsum| < q),  ssum = sum_ij a_ij (centred, |ssum| < 2^63),
//   bk[i] = (T mod p_i) * Kp_i mod p_i with T = sum_ij t_ij.  M = (T - ssum) / q = sum_ij m_ij, so
//   out = rsum + p * M  (mod q)  =  sum_ij [ round(p a_ij / q) + p m_ij ]  (mod q).
EXB_HD u64 hps_scale32_sum(i64 rsum, i64 ssum, const u32 *bk, const ScaleConsts &c, const Scale32Consts &s) {
    const bool neg = ssum < 0;
    const u64 sa = neg ? (u64)(-ssum) : (u64)ssum;
    const u32 a1 = (u32)(sa >> 32), a0 = (u32)sa;
    u64 frac = 0;
    u64 lo = rsum < 0 ? c.q - (u64)(-rsum) : (u64)rsum;
    u32 hi = 0;
#pragma unroll
    for (u32 i = 0; i < (u32)kMaxSmall; i++) {
        if (i < s.K) {
            const Mod32 &m = s.m[i];
            const u32 u = shoup32_lazy(a1, s.RK[i], s.RK_s[i], m.p) + shoup32_lazy(a0, s.Kp[i], s.Kp_s[i], m.p);   // |S| Kp, [0,4p)
            const u32 t = neg ? bk[i] + u : bk[i] + m.four_p - u;                                                    // [0,5p)
            const u32 mp = csub32(csub32(csub32(t, m.four_p), m.two_p), m.p);
            frac += (u64)mp * s.g[i];
            const u64 p0 = (u64)mp * (u32)s.C[i], p1 = (u64)mp * (u32)(s.C[i] >> 32);
            u64 nl = lo + p0;
            hi += nl < lo ? 1u : 0u;
            lo = nl;
            nl = lo + (p1 << 32);
            hi += (u32)(p1 >> 32) + (nl < lo ? 1u : 0u);
            lo = nl;
        }
    }
    const u32 alpha = (u32)((frac + ((u64)1 << 56)) >> 57);
    {
        const u64 nl = lo + s.CPn[alpha];
        hi += nl < lo ? 1u : 0u;
        lo = nl;
    }
    const u64 mid = ((u64)hi << 32) | (lo >> 32);
    const u32 th = (u32)(mid >> (s.sh - 32));
    const u32 k = mulhi32(th, s.rq);
    const u64 kq = (u64)k * (u32)c.q + ((u64)(k * (u32)(c.q >> 32)) << 32);
    const u64 r = lo - kq;
    return csub(csub(r, 2 * c.q), c.q);
}

// a = t mod q (canonical), bk[i] = (t mod p_i) * Kp_i mod p_i (canonical).  Needs 2^36 <= q < 2^60.
//   m'_i = bk_i - a_c * Kp_i  mod p_i,  a_c = a - [a > q/2] q
//   out  = rnd + sum_i m'_i * C_i - alpha * p * P'   (mod q),  accumulated exactly in 96 bits and
//          reduced once (Barrett on the top 32 bits: quotient estimate within 2, remainder < 3q).
EXB_HD u64 hps_scale32_coeff(u64 a, const u32 *bk, const ScaleConsts &c, const Scale32Consts &s) {
    const u64 rnd = s.plain32 ? round_term32(a, c) : round_term(a, c);      // [0, q)
    const bool neg = a > c.half_q;
    const u32 a1 = (u32)(a >> 32), a0 = (u32)a;
    u64 frac = 0;
    u64 lo = rnd;                // 96-bit accumulator hi:lo
    u32 hi = 0;
#pragma unroll
    for (u32 i = 0; i < (u32)kMaxSmall; i++) {
        if (i < s.K) {
            const Mod32 &m = s.m[i];
            const u32 u = shoup32_lazy(a1, s.RK[i], s.RK_s[i], m.p) + shoup32_lazy(a0, s.Kp[i], s.Kp_s[i], m.p);   // [0,4p)
            const u32 t = bk[i] + m.four_p - u + (neg ? s.QK[i] : 0u);                                              // (0,6p)
            const u32 mp = csub32(csub32(csub32(t, m.four_p), m.two_p), m.p);
            frac += (u64)mp * s.g[i];
            const u64 p0 = (u64)mp * (u32)s.C[i], p1 = (u64)mp * (u32)(s.C[i] >> 32);
            u64 nl = lo + p0;
            hi += nl < lo ? 1u : 0u;
            lo = nl;
            nl = lo + (p1 << 32);
            hi += (u32)(p1 >> 32) + (nl < lo ? 1u : 0u);
            lo = nl;
        }
    }
    const u32 alpha = (u32)((frac + ((u64)1 << 56)) >> 57);
    {
        const u64 nl = lo + s.CPn[alpha];
        hi += nl < lo ? 1u : 0u;
        lo = nl;
    }
    // T = hi:lo < (K * 2^27 + 2) q;  Th = floor(T / 2^sh) < 2^32
    const u64 mid = ((u64)hi << 32) | (lo >> 32);
    const u32 th = (u32)(mid >> (s.sh - 32));
    const u32 k = mulhi32(th, s.rq);
    const u64 kq = (u64)k * (u32)c.q + ((u64)(k * (u32)(c.q >> 32)) << 32);
    const u64 r = lo - kq;                                                  // [0, 3q)
    return csub(csub(r, 2 * c.q), c.q);
}

}  // namespace exb
