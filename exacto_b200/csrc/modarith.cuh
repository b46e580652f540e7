// modarith.cuh -- 64-bit modular arithmetic for sm_100a integer pipes.
//
// Device restatement of ring/modular.rs (reference file:line cited per function)
// redesigned for 32-bit IMAD hardware: constant multipliers use Shoup's method
// (one mulhi64 + two mullo64), data x data products use Montgomery with one
// operand pre-converted, small-quotient reductions use one-word Barrett.  All
// results that leave a kernel are canonical residues in [0, m), i.e. bit-equal
// to the reference's `u128 %` results (ring/modular.rs:8-11).
//
// Every function is __host__ __device__ so tests/host_emul can check the exact
// same code on the CPU against the oracle.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EXB_HD __host__ __device__ __forceinline__
#else
#define EXB_HD inline
#endif

namespace exb {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

struct u128w {  // little-endian 128-bit value
    u64 lo, hi;
};

EXB_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

EXB_HD u128w mul_wide(u64 a, u64 b) {
    u128w r;
    r.lo = a * b;
    r.hi = mulhi64(a, b);
    return r;
}

EXB_HD u128w add128(u128w a, u128w b) {
    u128w r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1u : 0u);
    return r;
}

EXB_HD u128w sub128(u128w a, u128w b) {
    u128w r;
    r.lo = a.lo - b.lo;
    r.hi = a.hi - b.hi - (a.lo < b.lo ? 1u : 0u);
    return r;
}

EXB_HD bool ge128(u128w a, u128w b) { return a.hi > b.hi || (a.hi == b.hi && a.lo >= b.lo); }
EXB_HD bool gt128(u128w a, u128w b) { return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo); }

// x in [0, 2m) -> [0, m)
EXB_HD u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }

// mod_add ring/modular.rs:57, mod_sub :65, mod_neg :75 (canonical operands, m < 2^63)
EXB_HD u64 mod_add(u64 a, u64 b, u64 m) { return csub(a + b, m); }
EXB_HD u64 mod_sub(u64 a, u64 b, u64 m) { return a >= b ? a - b : a + m - b; }
EXB_HD u64 mod_neg(u64 a, u64 m) { return a == 0 ? 0 : m - a; }

// ---- Shoup multiplication by a constant w (< m) with wp = floor(w * 2^64 / m).
// Valid for ANY x < 2^64; returns x*w mod m up to one extra m: result in [0, 2m).
EXB_HD u64 shoup_lazy(u64 x, u64 w, u64 wp, u64 m) { return x * w - mulhi64(x, wp) * m; }
EXB_HD u64 shoup(u64 x, u64 w, u64 wp, u64 m) { return csub(shoup_lazy(x, w, wp, m), m); }

// ---- One-word Barrett: x mod m for any x < 2^64 with mu = floor(2^64 / m).
// (Replaces the `%` in base_extend_centered bfv/eval.rs:230-240.)
EXB_HD u64 barrett_reduce(u64 x, u64 m, u64 mu) { return csub(x - mulhi64(x, mu) * m, m); }

// ---- Montgomery (R = 2^64), m odd, m < 2^63, minv_neg = -m^{-1} mod 2^64
// (ring/modular.rs:34-53 has the same reduction; the reference does not use it
// on the hot path, it uses `u128 %` -- results are equal as residues.)
// REDC of z < m * 2^64: returns z * R^-1 mod m in [0, 2m).
EXB_HD u64 mont_redc_lazy(u128w z, u64 m, u64 minv_neg) {
    u64 k = z.lo * minv_neg;
    // low words of z and k*m cancel to 0 mod 2^64; the carry is 1 unless z.lo == 0
    return z.hi + mulhi64(k, m) + (z.lo != 0 ? 1u : 0u);
}
// a * b * R^-1 mod m, lazy [0, 2m); needs a * b < m * 2^64.
EXB_HD u64 mont_mul_lazy(u64 a, u64 b, u64 m, u64 minv_neg) {
    return mont_redc_lazy(mul_wide(a, b), m, minv_neg);
}
// (a*b + c*d) * R^-1 mod m, lazy [0, 2m); needs a*b + c*d < m * 2^64.
EXB_HD u64 mont_mul2_lazy(u64 a, u64 b, u64 c, u64 d, u64 m, u64 minv_neg) {
    return mont_redc_lazy(add128(mul_wide(a, b), mul_wide(c, d)), m, minv_neg);
}

// ---- Hand-scheduled Shoup multiply-accumulate for the NTT butterflies ----------------
// Measured on B200 (tools/ubench2.cu): IMAD.LO ~1 issue cycle per warp, IMAD.WIDE ~2,
// IMAD.HI ~2.5, a 64-bit compare + conditional subtract ~6.  So the butterfly uses an
// *approximate* quotient  Q' = hi64(y*s - y0*s0)  in [Q-1, Q]  (3 WIDE: only the y0*s0 partial
// product is dropped; IMAD.HI would cost more and needs addend register pairs) and folds every addition into the IMAD
// accumulate operands:  result = addend + y*w + Q' * (2^64 - m)  (mod 2^64)
//                              = addend + T,   T = y*w mod m + e*m,  T in [0, 4m).
// Valid for any y < 2^64 as long as addend + 4m < 2^64.
EXB_HD u64 shoup_mad4(u64 y, u64 w, u64 s, u64 neg_m, u64 addend) {
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u64 qq, acc, c1, c2;\n\t"
        ".reg .u32 y0, y1, w0, w1, s0, s1, n0, n1, q0, q1, t1, t2, l1, l2, lo, hi, x;\n\t"
        "mov.b64 {y0, y1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.wide.u32 qq, y1, s1;\n\t"           // y1*s1
        "mul.wide.u32 c1, y1, s0;\n\t"           // cross terms as full products: their low halves give
        "mul.wide.u32 c2, y0, s1;\n\t"           // the exact carry, and no IMAD.HI addend pairs are needed
        "mov.b64 {q0, q1}, qq;\n\t"
        "mov.b64 {l1, t1}, c1;\n\t"
        "mov.b64 {l2, t2}, c2;\n\t"
        "add.cc.u32 l1, l1, l2;\n\t"
        "addc.cc.u32 q0, q0, t1;\n\t"
        "addc.u32 q1, q1, 0;\n\t"
        "add.cc.u32 q0, q0, t2;\n\t"
        "addc.u32 q1, q1, 0;\n\t"
        "mad.wide.u32 acc, y0, w0, %5;\n\t"      // addend + y0*w0        (64-bit accumulate)
        "mad.wide.u32 acc, q0, n0, acc;\n\t"     // + q0*n0
        "mul.lo.u32 x, y0, w1;\n\t"              // cross terms, all land in the high word
        "mad.lo.u32 x, y1, w0, x;\n\t"
        "mad.lo.u32 x, q0, n1, x;\n\t"
        "mad.lo.u32 x, q1, n0, x;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "add.u32 hi, hi, x;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(w), "l"(s), "l"(neg_m), "l"(addend));
    return r;
#else
    const u64 y0 = (u32)y, y1 = y >> 32, s0 = (u32)s, s1 = s >> 32;
    const u64 c1 = y1 * s0, c2 = y0 * s1;
    const u64 q = y1 * s1 + (c1 >> 32) + (c2 >> 32) + (((c1 & 0xffffffffu) + (c2 & 0xffffffffu)) >> 32);
    return addend + y * w + q * neg_m;
#endif
}

// Cheap bound control: if the high word shows x > k*m, subtract k*m.  `hi_km` = (k*m) >> 32.
// After it, x < k*m + 2^32 whenever x < 2*k*m before.  3 instructions instead of ~6.
EXB_HD u64 csub_hi(u64 x, u64 km, u32 hi_km) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 lo, hi, k0, k1;\n\t"
        "mov.b64 {lo, hi}, %0;\n\t"
        "mov.b64 {k0, k1}, %1;\n\t"
        "setp.gt.u32 p, hi, %2;\n\t"
        "@p sub.cc.u32 lo, lo, k0;\n\t"
        "@p subc.u32 hi, hi, k1;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "+l"(x) : "l"(km), "r"(hi_km));
    return x;
#else
    return (u32)(x >> 32) > hi_km ? x - km : x;
#endif
}

// Reduce any x < 2^64 to [0, 2m) with one 32-bit high-word Barrett estimate:
// k = floor((x >> 32) * rhi / 2^rsh) is floor(x/m) or one less (needs m >= 2^36).
EXB_HD u64 reduce_to_2m(u64 x, u64 neg_m, u32 rhi, u32 rsh) {
#if defined(__CUDA_ARCH__)
    const u32 k = __umulhi((u32)(x >> 32), rhi) >> rsh;
#else
    const u32 k = (u32)(((u64)(u32)(x >> 32) * rhi) >> 32) >> rsh;
#endif
    return x + (u64)k * neg_m;
}

// Per-modulus constants, computed on the host (context.cpp).
struct Modulus {
    u64 m;          // the prime
    u64 two_m;      // 2m
    u64 mu;         // floor(2^64 / m)          (Barrett)
    u64 minv_neg;   // -m^-1 mod 2^64           (Montgomery)
    u64 r_mod;      // 2^64 mod m               (to-Montgomery multiplier)
    u64 r_mod_s;    // Shoup companion of r_mod
    u64 r2_mod;     // 2^128 mod m
    u64 ninv;       // n^-1 mod m
    u64 ninv_s;
    u64 ninv_w;     // n^-1 * psi_inv_rev[1] mod m  (last inverse stage, folded normalise)
    u64 ninv_w_s;
    u64 neg_m;      // 2^64 - m
    u64 four_m;     // 4m (lazy-bound bias of the approximate butterflies)
    u32 hi_four_m;  // (4m) >> 32
    u32 rhi;        // floor(2^(32+sh) / m), sh = bits(m) - 1   (high-word Barrett)
    u32 rsh;        // sh - 32
    u32 lazy;       // 0: exact Harvey path (m < 2^62); 1: m < 2^60; 2: m < 2^55 (see ntt_core.cuh)
};

}  // namespace exb
