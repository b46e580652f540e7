// ntt32_core.cuh -- 32-bit negacyclic NTT passes for the *internal* auxiliary basis.
//
// bfv_mul_hps (bfv/eval.rs:157-209) needs the integer tensor t of the centred inputs modulo q and
// modulo an auxiliary basis only to recover m = (t - [t]_q) / q (bfv/eval.rs:301-404).  m is an
// integer with |m| <= n*q/2 + 1, so ANY auxiliary basis whose product exceeds 2|m| gives the same m,
// and hence bit-identical output, as long as the reference's own centred CRT cannot wrap either
// (P_ref / 2 > n*q/2 + 1; true for every two-aux-prime BASELINE config, checked on the host).
// The device therefore replaces the reference's two 54/55-bit aux primes by three 27-bit NTT primes:
// a 32-bit lazy butterfly is 1 IMAD.HI + 2 IMAD.LO + 2 ALU, ~4x cheaper than the 64-bit one on the
// integer-multiply pipe that bounds these kernels (profiles/r01_ncu_tensor_kernel.json).
//
// Conventions are those of ntt_core.cuh (CT forward natural -> bit-reversed, GS inverse, psi from the
// same rule).  Primes are < 2^27 (kSmallPrimeBits) so 32p fits a u32: see the lazy ranges below.
#pragma once
#include "modarith.cuh"
#include "ntt_core.cuh"

namespace exb {

constexpr int kSmallPrimeBits = 27;

struct Tw32 {  // twiddle + Shoup companion floor(w * 2^32 / p)
    u32 w, s;
};
struct TwHead32 {
    Tw32 t[16];
    EXB_HD const Tw32 &operator[](u32 i) const { return t[i]; }
};

struct Mod32 {
    u32 p, two_p, neg_p;      // neg_p = 2^32 - p
    u32 four_p;
    u32 pinv_neg;             // -p^-1 mod 2^32 (Montgomery, R = 2^32)
    u32 r_mod, r_mod_s;       // 2^32 mod p and its Shoup companion (to-Montgomery / high-word fold)
    u32 one_s;                // floor(2^32 / p): Shoup companion of 1 (reduces any u32 to [0,2p))
    u32 c28, c28_s;           // 2^28 mod p and its Shoup companion (60-bit -> p reduction with one multiply)
    u32 ninv, ninv_s, ninv_w, ninv_w_s;
};

EXB_HD u32 mulhi32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
// x - m if x >= m else x (m > 0), as an unsigned min: x < m makes x - m wrap above x
EXB_HD u32 csub32(u32 x, u32 m) { const u32 d = x - m; return d < x ? d : x; }
// x * w mod p for any x < 2^32, result in [0, 2p)
EXB_HD u32 shoup32_lazy(u32 x, u32 w, u32 s, u32 p) { return x * w - mulhi32(x, s) * p; }
EXB_HD u32 shoup32(u32 x, u32 w, u32 s, u32 p) { return csub32(shoup32_lazy(x, w, s, p), p); }
// Montgomery REDC of z < p * 2^32: z * 2^-32 mod p in [0, 2p)
EXB_HD u32 mont32_redc_lazy(u64 z, u32 p, u32 pinv_neg) {
    const u32 lo = (u32)z;
    const u32 k = lo * pinv_neg;
    return (u32)(z >> 32) + mulhi32(k, p) + (lo != 0 ? 1u : 0u);
}
// any u64 v < 2^62 reduced to [0, p):  v = v1 * 2^32 + v0
EXB_HD u32 reduce64_to_p(u64 v, const Mod32 &m) {
    const u32 hi = shoup32_lazy((u32)(v >> 32), m.r_mod, m.r_mod_s, m.p);   // [0,2p)
    const u32 lo = shoup32_lazy((u32)v, 1u, m.one_s, m.p);                  // [0,2p)
    return csub32(csub32(hi + lo, m.two_p), m.p);
}

// Lazy ranges.  The primes are < 2^27, so 32p < 2^32 and a u32 holds every intermediate below
// without the per-butterfly conditional subtractions of a Harvey NTT:
//   forward (CT):  x' = x + T, y' = x - T + 2p with T = y*w mod p in [0, 2p) for ANY u32 y; a value
//                  grows by at most 2p per stage: canonical input -> < 25p after the 12 stages;
//   inverse (GS):  S = x + y, D = x - y + BIAS (BIAS = bound of y) and D*w mod p in [0, 2p) for any
//                  u32 D; only the sum path grows.  Over a 3-stage register pass with inputs < 4p:
//                  v0 < 32p, v1 < 8p, v2, v3 < 4p, v4..v7 < 2p; v0 and v1 are folded back to [0, 2p).
// Any u32 -> [0, 2p)
EXB_HD u32 fold32(u32 x, const Mod32 &m) { return x - mulhi32(x, m.one_s) * m.p; }

EXB_HD void ct32(u32 &x, u32 &y, const Tw32 t, const Mod32 &m) {
    const u32 Q = mulhi32(y, t.s);
    const u32 T = y * t.w + Q * m.neg_p;           // [0, 2p)
    y = x - T + m.two_p;
    x = x + T;
}
// bias = (bound of y) as a multiple of p, kept in a register by the pass
EXB_HD void gs32(u32 &x, u32 &y, const Tw32 t, const Mod32 &m, u32 bias) {
    const u32 S = x + y;
    const u32 D = x - y + bias;
    x = S;
    y = shoup32_lazy(D, t.w, t.s, m.p);
}

// CNT consecutive twiddles from entry `first`; like load_tws (ntt_core.cuh): a thread's own entries of a global
// table (S == 0) come as one vector load that bypasses L1.
template <int CNT, bool OWN, class TW>
EXB_HD void load_tws32(const TW &tw, u32 first, Tw32 (&w)[CNT]) {
#pragma unroll
    for (int g = 0; g < CNT; g++) w[g] = tw[first + g];
}
#if defined(__CUDA_ARCH__)
template <int CNT, bool OWN>
EXB_HD void load_tws32(const Tw32 *const &tw, u32 first, Tw32 (&w)[CNT]) {
    static_assert(CNT == 1 || CNT == 2 || CNT == 4, "one vector load");
    const Tw32 *p = tw + first;
    if constexpr (!OWN) {
#pragma unroll
        for (int g = 0; g < CNT; g++) w[g] = p[g];
    } else if constexpr (CNT == 1) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(w[0].w), "=r"(w[0].s) : "l"(p));
    } else if constexpr (CNT == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[0].w), "=r"(w[0].s), "=r"(w[1].w), "=r"(w[1].s) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w[0].w), "=r"(w[0].s), "=r"(w[1].w), "=r"(w[1].s), "=r"(w[2].w), "=r"(w[2].s), "=r"(w[3].w),
                       "=r"(w[3].s) : "l"(p));
    }
}
#endif

template <int LOGN, int S, int NB, int J, class TW>
EXB_HD void fwd_stage32(u32 (&v)[1 << NB], const TW &tw, u32 pre, const Mod32 &m) {
    constexpr int P = LOGN - NB - S;
    constexpr int half = (1 << NB) >> (J + 1);
    Tw32 w[1 << J];
    load_tws32<(1 << J), S == 0>(tw, (1u << (P + J)) + (pre << J), w);
#pragma unroll
    for (int g = 0; g < (1 << J); g++) {
#pragma unroll
        for (int u = 0; u < half; u++) ct32(v[g * 2 * half + u], v[g * 2 * half + u + half], w[g], m);
    }
}
// Values grow by < 6p over the pass (no reduction inside).
template <int LOGN, int S, int NB, class TW>
EXB_HD void fwd_pass32(u32 (&v)[1 << NB], const TW &tw, u32 t, const Mod32 &m) {
    static_assert(NB == 3, "lazy bounds are derived for 3-stage passes");
    const u32 pre = t >> S;
    fwd_stage32<LOGN, S, NB, 0>(v, tw, pre, m);
    fwd_stage32<LOGN, S, NB, 1>(v, tw, pre, m);
    fwd_stage32<LOGN, S, NB, 2>(v, tw, pre, m);
}

template <int LOGN, int S, int NB, int J, class TW>
EXB_HD void inv_stage32(u32 (&v)[1 << NB], const TW &tw, u32 pre, const Mod32 &m) {
    constexpr int P = LOGN - NB - S;
    constexpr int half = 1 << J;
    const u32 bias = m.four_p << J;
    constexpr int CNT = (1 << NB) >> (J + 1);
    Tw32 w[CNT];
    load_tws32<CNT, S == 0>(tw, (1u << (P + NB - 1 - J)) + (pre << (NB - 1 - J)), w);
#pragma unroll
    for (int g = 0; g < CNT; g++) {
#pragma unroll
        for (int u = 0; u < half; u++) gs32(v[g * 2 * half + u], v[g * 2 * half + u + half], w[g], m, bias);
    }
}
// Inputs < 4p.  !LAST: outputs < 4p.  LAST: the final stage folds n^-1, outputs canonical.
template <int LOGN, int S, int NB, bool LAST, class TW>
EXB_HD void inv_pass32(u32 (&v)[1 << NB], const TW &tw, u32 t, const Mod32 &m) {
    static_assert(NB == 3, "lazy bounds are derived for 3-stage passes");
    const u32 pre = t >> S;
    inv_stage32<LOGN, S, NB, 0>(v, tw, pre, m);
    inv_stage32<LOGN, S, NB, 1>(v, tw, pre, m);
    if constexpr (LAST) {
        constexpr int H = (1 << NB) / 2;
#pragma unroll
        for (int u = 0; u < H; u++) {
            const u32 S2 = v[u] + v[u + H];                       // < 32p
            const u32 D = v[u] - v[u + H] + (m.four_p << 2);
            v[u] = shoup32(S2, m.ninv, m.ninv_s, m.p);
            v[u + H] = shoup32(D, m.ninv_w, m.ninv_w_s, m.p);
        }
    } else {
        inv_stage32<LOGN, S, NB, NB - 1>(v, tw, pre, m);
        v[0] = fold32(v[0], m);
        v[1] = fold32(v[1], m);
    }
}

// Shared-memory image of a u32 polynomial: 16-byte chunk c = e >> 2 lives at chunk
// c ^ ((c >> 3) & 7) (the same chunk swizzle as the u64 image), conflict-free for all pass shapes.
EXB_HD u32 swz32(u32 e) {
    const u32 c = e >> 2;
    return ((c ^ ((c >> 3) & 7u)) << 2) | (e & 3u);
}
template <int NB, int S>
EXB_HD void load_vals32(u32 (&v)[1 << NB], const u32 *sm, u32 t) {
#pragma unroll
    for (int k = 0; k < (1 << NB); k++) v[k] = sm[swz32(elem_index<NB, S>(t, k))];
}
template <int NB, int S>
EXB_HD void store_vals32(const u32 (&v)[1 << NB], u32 *sm, u32 t) {
#pragma unroll
    for (int k = 0; k < (1 << NB); k++) sm[swz32(elem_index<NB, S>(t, k))] = v[k];
}

}  // namespace exb
