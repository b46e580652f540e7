// ntt_core.cuh -- register-resident radix-16 passes of the negacyclic NTT.
//
// Replaces concrete-ntt's Plan::fwd / Plan::inv + normalize as called from
// ring/ntt.rs:49,60-62.  Conventions (shared with oracle/exacto_oracle.c):
//   forward : Cooley-Tukey, natural order in  -> bit-reversed order out,
//             stage with m groups uses psi_rev[m + g]  (psi_rev[k] = psi^bitrev(k))
//   inverse : Gentleman-Sande, bit-reversed in -> natural out, stage with h
//             groups uses psi_inv_rev[h + g]; n^-1 is folded into the last stage.
// Lazy reduction: the exact Harvey butterflies (forward [0, 4q), inverse [0, 2q)) work for
// any prime q < 2^62; primes with head-room use cheaper approximate butterflies (LAZY 1/2).
//
// An n = 2^LOGN transform is done by n/16 threads; each thread keeps 16
// coefficients in registers and runs 4 butterfly stages per pass, exchanging
// through a swizzled shared-memory image between passes.  The per-thread code
// is __host__ __device__ so tests/host_emul can replay it thread by thread.
#pragma once
#include "modarith.cuh"

namespace exb {

struct alignas(16) Tw {  // twiddle + Shoup companion floor(w * 2^64 / q)
    u64 w, s;
};

// Shared-memory swizzle (16-byte granular, == TMA SWIZZLE_128B): conflict-free
// 64-bit accesses for all three pass shapes, and 128-bit accesses for S = 0.
EXB_HD u32 swz(u32 e) { return e ^ (((e >> 4) & 7u) << 1); }

// x, y in [0, 4q) -> [0, 4q)
EXB_HD void ct_bfly(u64 &x, u64 &y, const Tw t, const u64 q, const u64 q2) {
    const u64 X = csub(x, q2);
    const u64 T = shoup_lazy(y, t.w, t.s, q);
    x = X + T;
    y = X - T + q2;
}

// x, y in [0, 2q) -> [0, 2q)
EXB_HD void gs_bfly(u64 &x, u64 &y, const Tw t, const u64 q, const u64 q2) {
    const u64 S = csub(x + y, q2);
    const u64 D = x - y + q2;
    x = S;
    y = shoup_lazy(D, t.w, t.s, q);
}

// Element owned by thread `t` at local index k of the pass whose NB local bits sit at bit
// position S of the coefficient index (a pass = NB butterfly stages on 2^NB register values).
template <int NB, int S>
EXB_HD u32 elem_index(u32 t, u32 k) {
    const u32 lo = t & ((1u << S) - 1u);
    const u32 pre = t >> S;
    return (pre << (S + NB)) | (k << S) | lo;
}

// [0, 4q) -> [0, q)
EXB_HD u64 reduce4(u64 x, u64 q, u64 q2) { return csub(csub(x, q2), q); }

// ---- Lazy-bound butterfly variants ---------------------------------------------------
// LAZY = 0 : exact Harvey butterflies above, any prime < 2^62.
// LAZY = 1 : prime < 2^60.  Approximate Shoup quotient (T in [0,4q)), X' fused into the IMAD
//            chain, bounds kept by a cheap high-word conditional subtract of 4q:
//            forward values stay in [0, 8q + 2^32), inverse values in [0, 4q + 2^45).
// LAZY = 2 : prime < 2^55 (>= 9 bits of head-room).  No conditional subtracts at all in the
//            forward transform (values < 52q after 12 stages); the inverse lets sums double
//            and re-centres them once with a high-word Barrett step (see inv_stage).
struct LazyC {  // per-modulus constants kept in registers by the passes
    u64 q, q2, neg_q, four_q;
    u32 hi_four_q, rhi, rsh;
};
EXB_HD LazyC make_lazyc(const Modulus &m) {
    LazyC c;
    c.q = m.m; c.q2 = m.two_m; c.neg_q = m.neg_m; c.four_q = m.four_m;
    c.hi_four_q = m.hi_four_m; c.rhi = m.rhi; c.rsh = m.rsh;
    return c;
}

template <int LAZY>
EXB_HD void ct_bfly_l(u64 &x, u64 &y, const Tw t, const LazyC &c) {
    if (LAZY == 0) {
        ct_bfly(x, y, t, c.q, c.q2);
    } else {
        const u64 X = LAZY == 1 ? csub_hi(x, c.four_q, c.hi_four_q) : x;
        const u64 xn = shoup_mad4(y, t.w, t.s, c.neg_q, X);        // X + T, T in [0, 4q)
        y = (X + X + c.four_q) - xn;                                // X - T + 4q
        x = xn;
    }
}

// BIAS: a multiple of q that is >= the bound of y (LAZY >= 1).
template <int LAZY>
EXB_HD void gs_bfly_l(u64 &x, u64 &y, const Tw t, const LazyC &c, const u64 bias) {
    if (LAZY == 0) {
        gs_bfly(x, y, t, c.q, c.q2);
    } else {
        const u64 S = x + y;
        const u64 D = x + bias - y;
        x = LAZY == 1 ? csub_hi(S, c.four_q, c.hi_four_q) : S;
        y = shoup_mad4(D, t.w, t.s, c.neg_q, 0);                    // in [0, 4q)
    }
}

// Exact canonical value of any x < 2^64 (LAZY >= 1): high-word Barrett to [0, 2q), then one csub.
EXB_HD u64 reduce_full(u64 x, const LazyC &c) { return csub(reduce_to_2m(x, c.neg_q, c.rhi, c.rsh), c.q); }

// One butterfly stage J (0..3) of a forward pass over local bits [S, S+4): global
// stage P + J with P = LOGN-4-S.  Compile-time J keeps v[] in registers.
// First 16 twiddles by value: lives in the kernel-parameter constant bank, so the pass whose
// twiddles are the same for every thread (the first forward / last inverse pass) reads them
// as uniform operands instead of issuing loads.
struct TwHead {
    Tw t[16];
    EXB_HD const Tw &operator[](u32 i) const { return t[i]; }
};

// CNT consecutive twiddles starting at entry `first` (CNT-aligned).  Generic tables (the uniform head, shared-memory
// copies) and the passes whose entries are shared by groups of threads: plain indexing.  A global table in the pass
// where every thread has its own entries (OWN: S == 0) is read with one 128-bit or one / two 256-bit loads that do not
// allocate in L1: that part of a table (57 KB) is larger than the L1 the fused kernels leave beside their shared
// memory (228 KB - 2 x 80..96 KB), so caching it only evicts the small shared entries of the other passes; measured
// against keeping any share of it cacheable.
template <int CNT, bool OWN, class TW>
EXB_HD void load_tws(const TW &tw, u32 first, Tw (&w)[CNT]) {
#pragma unroll
    for (int g = 0; g < CNT; g++) w[g] = tw[first + g];
}
#if defined(__CUDA_ARCH__)
template <int CNT, bool OWN>
EXB_HD void load_tws(const Tw *const &tw, u32 first, Tw (&w)[CNT]) {
    if constexpr (!OWN) {
#pragma unroll
        for (int g = 0; g < CNT; g++) w[g] = tw[first + g];
    } else if constexpr (CNT == 1) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(w[0].w), "=l"(w[0].s) : "l"(tw + first));
    } else {
#pragma unroll
        for (int g = 0; g < CNT; g += 2)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];"
                         : "=l"(w[g].w), "=l"(w[g].s), "=l"(w[g + 1].w), "=l"(w[g + 1].s) : "l"(tw + first + g));
    }
}
#endif

// One butterfly stage J (0..NB-1) of a forward pass over local bits [S, S+NB): global stage
// P + J with P = LOGN-NB-S.  Compile-time J keeps v[] in registers.
template <int LOGN, int S, int NB, int J, int LAZY, class TW>
EXB_HD void fwd_stage(u64 (&v)[1 << NB], const TW &tw, u32 pre, const LazyC &c) {
    constexpr int P = LOGN - NB - S;
    constexpr int half = (1 << NB) >> (J + 1);
    Tw w[1 << J];
    load_tws<(1 << J), S == 0>(tw, (1u << (P + J)) + (pre << J), w);
#pragma unroll
    for (int g = 0; g < (1 << J); g++) {
#pragma unroll
        for (int u = 0; u < half; u++) ct_bfly_l<LAZY>(v[g * 2 * half + u], v[g * 2 * half + u + half], w[g], c);
    }
}

template <int LOGN, int S, int NB, int LAZY, class TW>
EXB_HD void fwd_pass(u64 (&v)[1 << NB], const TW &tw, u32 t, const LazyC &c) {
    const u32 pre = t >> S;
    fwd_stage<LOGN, S, NB, 0, LAZY>(v, tw, pre, c);
    fwd_stage<LOGN, S, NB, 1, LAZY>(v, tw, pre, c);
    fwd_stage<LOGN, S, NB, 2, LAZY>(v, tw, pre, c);
    if constexpr (NB >= 4) fwd_stage<LOGN, S, NB, 3, LAZY>(v, tw, pre, c);
}

// Final reduction of forward outputs to [0, q).
template <int LAZY>
EXB_HD u64 fwd_final(u64 x, const LazyC &c) { return LAZY == 0 ? reduce4(x, c.q, c.q2) : reduce_full(x, c); }

// Inverse stage J of a pass over local bits [S, S+NB): pairs local bit J; global stage index
// GSI = S + J (0 .. LOGN-1).  For LAZY == 2 the sum outputs double every stage; they are
// re-centred to [0, 2q) after global stage kRecentre, so every stage's bound is 4q << r with
// r = stages since the last re-centre (inputs of the transform must be < 4q).
constexpr int kRecentre = 5;
template <int LOGN, int S, int NB, int J, int LAZY, class TW>
EXB_HD void inv_stage(u64 (&v)[1 << NB], const TW &tw, u32 pre, const LazyC &c) {
    constexpr int P = LOGN - NB - S;
    constexpr int half = 1 << J;
    constexpr int GSI = S + J;
    constexpr int r = GSI <= kRecentre ? GSI : GSI - kRecentre - 1;
    const u64 bias = LAZY == 2 ? (c.four_q << r) : (c.four_q << 1);
    constexpr int CNT = (1 << NB) >> (J + 1);
    Tw w[CNT];
    load_tws<CNT, S == 0>(tw, (1u << (P + NB - 1 - J)) + (pre << (NB - 1 - J)), w);
#pragma unroll
    for (int g = 0; g < CNT; g++) {
#pragma unroll
        for (int u = 0; u < half; u++) {
            gs_bfly_l<LAZY>(v[g * 2 * half + u], v[g * 2 * half + u + half], w[g], c, bias);
            if (LAZY == 2 && GSI == kRecentre)
                v[g * 2 * half + u] = reduce_to_2m(v[g * 2 * half + u], c.neg_q, c.rhi, c.rsh);
        }
    }
}

// Last inverse stage (global h = 1): multiplies by n^-1 (x side) and n^-1 * psi_inv_rev[1]
// (y side) instead of the twiddle: plan.normalize (ring/ntt.rs:62) folded in.  Outputs canonical.
template <int LOGN, int NB, int LAZY>
EXB_HD void inv_last_stage(u64 (&v)[1 << NB], const Modulus &mod, const LazyC &c) {
    const u64 ni = mod.ninv, nis = mod.ninv_s, nw = mod.ninv_w, nws = mod.ninv_w_s;
    constexpr int GSI = LOGN - 1;
    constexpr int r = GSI <= kRecentre ? GSI : GSI - kRecentre - 1;
    constexpr int H = (1 << NB) / 2;
    const u64 bias = LAZY == 0 ? c.q2 : (LAZY == 2 ? (c.four_q << r) : (c.four_q << 1));
#pragma unroll
    for (int u = 0; u < H; u++) {
        const u64 S2 = v[u] + v[u + H];        // Shoup takes any u64
        const u64 D = v[u] + bias - v[u + H];
        if (LAZY == 0) {
            v[u] = csub(shoup_lazy(S2, ni, nis, c.q), c.q);
            v[u + H] = csub(shoup_lazy(D, nw, nws, c.q), c.q);
        } else {                                // [0, 4q) -> [0, q)
            v[u] = reduce4(shoup_mad4(S2, ni, nis, c.neg_q, 0), c.q, c.q2);
            v[u + H] = reduce4(shoup_mad4(D, nw, nws, c.neg_q, 0), c.q, c.q2);
        }
    }
}

// Inverse pass; LAST = the pass that contains global stage LOGN-1 (S + NB == LOGN).
template <int LOGN, int S, int NB, bool LAST, int LAZY, class TW>
EXB_HD void inv_pass(u64 (&v)[1 << NB], const TW &tw, u32 t, const Modulus &mod, const LazyC &c) {
    const u32 pre = t >> S;
    inv_stage<LOGN, S, NB, 0, LAZY>(v, tw, pre, c);
    inv_stage<LOGN, S, NB, 1, LAZY>(v, tw, pre, c);
    if constexpr (NB >= 4) inv_stage<LOGN, S, NB, 2, LAZY>(v, tw, pre, c);
    if constexpr (LAST) inv_last_stage<LOGN, NB, LAZY>(v, mod, c);
    else inv_stage<LOGN, S, NB, NB - 1, LAZY>(v, tw, pre, c);
}

// ---- register <-> swizzled shared-memory image -----------------------------------------
// `sm` holds n coefficients at swz(e).  The S = 0 pass owns 2^NB consecutive coefficients per
// thread and moves them as 128-bit words (swz keeps 16-byte pairs together and conflict-free).
template <int NB, int S>
EXB_HD void load_vals(u64 (&v)[1 << NB], const u64 *sm, u32 t) {
    if constexpr (S == 0) {
#pragma unroll
        for (int k = 0; k < (1 << NB); k += 2) {
            const u64 *p = sm + swz(elem_index<NB, 0>(t, k));
            v[k] = p[0];
            v[k + 1] = p[1];
        }
    } else {
#pragma unroll
        for (int k = 0; k < (1 << NB); k++) v[k] = sm[swz(elem_index<NB, S>(t, k))];
    }
}
template <int NB, int S>
EXB_HD void store_vals(const u64 (&v)[1 << NB], u64 *sm, u32 t) {
    if constexpr (S == 0) {
#pragma unroll
        for (int k = 0; k < (1 << NB); k += 2) {
            u64 *p = sm + swz(elem_index<NB, 0>(t, k));
            p[0] = v[k];
            p[1] = v[k + 1];
        }
    } else {
#pragma unroll
        for (int k = 0; k < (1 << NB); k++) sm[swz(elem_index<NB, S>(t, k))] = v[k];
    }
}

}  // namespace exb
