// hps32.cuh -- hps_scale (bfv/eval.rs:257-413) on the internal 30-bit auxiliary basis.
//
// With a = t mod q (canonical) and b_i = t mod p_i for the small primes p_i (P' = prod p_i),
//   M = (t - a_c) / q           (exact integer, |M| <= n*q/2 + 1  <<  P'/2)
//   out = round(p * a_c / q) + p * M   (mod q)            -- bfv/eval.rs:323-331 / :390-403
// M is recovered without big integers (HPS "exact alpha"):
//   m'_i = (b_i - [a_c]_{p_i}) * (q * P'/p_i)^-1  mod p_i,
//   M    = sum_i m'_i * (P'/p_i)  -  alpha * P',   alpha = round(sum_i m'_i / p_i)
// and alpha is exact in 2^-60 fixed point (error < 2^-28) because |M| / P' < 2^-8 (checked on the host).
// The result equals the reference's for every input because the reference's own centred CRT of m
// cannot wrap for the parameter sets this path is enabled for (host_setup.cpp: small_basis_ok).
#pragma once
#include "hps.cuh"
#include "ntt32_core.cuh"

namespace exb {

constexpr int kMaxSmall = 4;

struct Scale32Consts {
    u32 K;                       // number of small primes
    u32 pad_;
    Mod32 m[kMaxSmall];
    u32 Kp[kMaxSmall], Kp_s[kMaxSmall];   // (q * P'/p_i)^-1 mod p_i  + Shoup32 companion
    u32 g[kMaxSmall];                     // floor(2^60 / p_i)
    u64 C[kMaxSmall], C_s[kMaxSmall];     // p * (P'/p_i) mod q       + Shoup64 companion
    u64 CP[kMaxSmall + 1];                // alpha * p * P' mod q, alpha = 0..K
};

// Centred value of a (mod q) reduced to [0, p_i): the ext rule of bfv/eval.rs:230-240.
EXB_HD u32 ext32_centered(u64 a, u64 q, u64 half_q, const Mod32 &m) {
    const bool neg = a > half_q;
    const u32 r = reduce64_to_p(neg ? q - a : a, m);
    return (neg && r != 0) ? m.p - r : r;
}

// addend + x * w mod q (+ up to 3q) for a 32-bit x: approximate Shoup quotient hi32(x * s1).
EXB_HD u64 shoup_small_mad(u32 x, u64 w, u64 s, u64 neg_q, u64 addend) {
    const u32 qh = mulhi32(x, (u32)(s >> 32));                // in [Q-2, Q]
    return addend + (u64)x * w + (u64)qh * neg_q;
}

// a = t mod q, b[i] = t mod p_i (all canonical).  Needs q < 2^60.
EXB_HD u64 hps_scale32_coeff(u64 a, const u32 *b, const ScaleConsts &c, const Scale32Consts &s,
                             const LazyC &lq) {
    u64 acc = round_term(a, c);                               // [0, q)
    u64 frac = 0;
#pragma unroll
    for (u32 i = 0; i < (u32)kMaxSmall; i++) {
        if (i < s.K) {
            const Mod32 &m = s.m[i];
            const u32 ae = ext32_centered(a, c.q, c.half_q, m);
            const u32 d = b[i] >= ae ? b[i] - ae : b[i] + m.p - ae;
            const u32 mp = shoup32(d, s.Kp[i], s.Kp_s[i], m.p);
            frac += (u64)mp * s.g[i];
            acc = shoup_small_mad(mp, s.C[i], s.C_s[i], lq.neg_q, acc);   // + [0, 4q)
        }
    }
    const u32 alpha = (u32)((frac + ((u64)1 << 59)) >> 60);
    acc += c.q - s.CP[alpha];                                 // total < (2 + 4K) q <= 14 q < 2^64
    return reduce_full(acc, lq);
}

}  // namespace exb
