// ntt_core.cuh -- register-resident radix-16 passes of the negacyclic NTT.
//
// Replaces concrete-ntt's Plan::fwd / Plan::inv + normalize as called from
// ring/ntt.rs:49,60-62.  Conventions (shared with oracle/exacto_oracle.c):
//   forward : Cooley-Tukey, natural order in  -> bit-reversed order out,
//             stage with m groups uses psi_rev[m + g]  (psi_rev[k] = psi^bitrev(k))
//   inverse : Gentleman-Sande, bit-reversed in -> natural out, stage with h
//             groups uses psi_inv_rev[h + g]; n^-1 is folded into the last stage.
// Lazy reduction (Harvey): forward keeps values in [0, 4q), inverse in [0, 2q);
// any prime q < 2^62 works.
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

// Element owned by thread `t` at local index k of the pass whose 4 local bits sit
// at bit position S of the coefficient index.
template <int S>
EXB_HD u32 elem_index(u32 t, u32 k) {
    const u32 lo = t & ((1u << S) - 1u);
    const u32 pre = t >> S;
    return (pre << (S + 4)) | (k << S) | lo;
}

// One butterfly stage J (0..3) of a forward pass over local bits [S, S+4): global
// stage P + J with P = LOGN-4-S.  Compile-time J keeps v[] in registers.
template <int LOGN, int S, int J>
EXB_HD void fwd_stage(u64 (&v)[16], const Tw *__restrict__ tw, u32 pre, u64 q, u64 q2) {
    constexpr int P = LOGN - 4 - S;
    constexpr int half = 8 >> J;
#pragma unroll
    for (int g = 0; g < (1 << J); g++) {
        const Tw w = tw[(1u << (P + J)) + (pre << J) + g];
#pragma unroll
        for (int u = 0; u < half; u++) ct_bfly(v[g * 2 * half + u], v[g * 2 * half + u + half], w, q, q2);
    }
}

template <int LOGN, int S>
EXB_HD void fwd_pass16(u64 (&v)[16], const Tw *__restrict__ tw, u32 t, u64 q, u64 q2) {
    const u32 pre = t >> S;
    fwd_stage<LOGN, S, 0>(v, tw, pre, q, q2);
    fwd_stage<LOGN, S, 1>(v, tw, pre, q, q2);
    fwd_stage<LOGN, S, 2>(v, tw, pre, q, q2);
    fwd_stage<LOGN, S, 3>(v, tw, pre, q, q2);
}

// Inverse stage J of a pass over local bits [S, S+4): pairs local bit J.
template <int LOGN, int S, int J>
EXB_HD void inv_stage(u64 (&v)[16], const Tw *__restrict__ tw, u32 pre, u64 q, u64 q2) {
    constexpr int P = LOGN - 4 - S;
    constexpr int half = 1 << J;
#pragma unroll
    for (int g = 0; g < (8 >> J); g++) {
        const Tw w = tw[(1u << (P + 3 - J)) + (pre << (3 - J)) + g];
#pragma unroll
        for (int u = 0; u < half; u++) gs_bfly(v[g * 2 * half + u], v[g * 2 * half + u + half], w, q, q2);
    }
}

// Inverse pass.  If LAST, the final stage (global h = 1) multiplies by n^-1 (x side)
// and n^-1 * psi_inv_rev[1] (y side) instead: plan.normalize folded in.
template <int LOGN, int S, bool LAST>
EXB_HD void inv_pass16(u64 (&v)[16], const Tw *__restrict__ tw, u32 t, const Modulus &mod) {
    const u64 q = mod.m, q2 = mod.two_m;
    const u32 pre = t >> S;
    inv_stage<LOGN, S, 0>(v, tw, pre, q, q2);
    inv_stage<LOGN, S, 1>(v, tw, pre, q, q2);
    inv_stage<LOGN, S, 2>(v, tw, pre, q, q2);
    if constexpr (LAST) {
        const u64 ni = mod.ninv, nis = mod.ninv_s, nw = mod.ninv_w, nws = mod.ninv_w_s;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const u64 S2 = v[u] + v[u + 8];        // < 4q, Shoup takes any u64
            const u64 D = v[u] - v[u + 8] + q2;
            v[u] = shoup_lazy(S2, ni, nis, q);
            v[u + 8] = shoup_lazy(D, nw, nws, q);
        }
    } else {
        inv_stage<LOGN, S, 3>(v, tw, pre, q, q2);
    }
}

// ---- whole-transform helpers on a swizzled shared-memory image ------------------
// `sm` holds n = 2^LOGN coefficients at swz(e).  These are the per-thread bodies;
// the caller provides the barriers between passes.

template <int S>
EXB_HD void load16(u64 (&v)[16], const u64 *sm, u32 t) {
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = sm[swz(elem_index<S>(t, k))];
}
template <int S>
EXB_HD void store16(const u64 (&v)[16], u64 *sm, u32 t) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[swz(elem_index<S>(t, k))] = v[k];
}

// [0, 4q) -> [0, q)
EXB_HD u64 reduce4(u64 x, u64 q, u64 q2) { return csub(csub(x, q2), q); }

}  // namespace exb
