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
};

}  // namespace exb
