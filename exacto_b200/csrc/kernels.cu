// kernels.cu -- sm_100a kernels of the ciphertext-multiplication hot path and its callers.
//
//   K1/K2  ntt12_persist_kernel, ntt_*_kernel   ring/ntt.rs:42-67 (concrete-ntt fwd / inv+normalize)
//   K3     poly_op_kernel                       ring/ntt.rs:75-139, ring/rns.rs:159-217
//   K4     lift32_kernel / lift_kernel          bfv/eval.rs:217-247 base_extend_centered (once per limb)
//   K5     tensor01_kernel                      bfv/eval.rs:187-203 tensor + hps_scale, components 0/1 per output limb
//          tensor32_kernel / tensor_kernel      the same per product (component 2: keyswitch.rs:11-52 digits)
//   K6+K7  relin12_kernel / relin_kernel        bfv/keyswitch.rs:83-95 + dbfv/eval.rs:125-132 per-k sums
//          relin12_wide_kernel + relin_reduce_kernel   the same with one CTA per transform (small batches)
//          reduce_mac_kernel                    dbfv/reduction.rs:34-52 (general rep != 0 case)
//          galois_kernel                        bfv/eval.rs:512-561 automorphism + key switch
//          decrypt_kernel                       bfv/encrypt.rs:111-178
//
// No tensor cores: nothing here is a dense contraction; the work is 64-bit (and, on the internal auxiliary
// basis, 32-bit) modular arithmetic on the integer pipes, staged through shared memory.
#include "kernels.cuh"

#ifndef EXB_HOST_EMUL
#include <cuda.h>
#endif
#include <atomic>
#include <cstdlib>
#include <cstring>

// Dynamic shared memory.  tests/host_emul compiles this file with g++ and a shim
// (EXB_HOST_EMUL) that maps CUDA's execution model onto CPU threads + barriers.
#ifndef EXB_HOST_EMUL
#define EXB_DYN_SMEM(name) extern __shared__ __align__(16) u64 name[]
#endif

namespace exb {

std::atomic<unsigned long long> g_launch_count{0};

// Streaming 64-bit load that does not allocate in L1 (keeps L1 for the twiddle tables).
__device__ __forceinline__ u64 ld_stream(const u64 *p) {
#ifndef EXB_HOST_EMUL
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}

// Stores into peer GPUs' memory (k-sharded dbfv_mul, kernel-store transport) are made visible system-wide before
// the kernel ends.
__device__ __forceinline__ void peer_fence() {
#ifndef EXB_HOST_EMUL
    __threadfence_system();
#endif
}

// ---------------------------------------------------------------------------------
// Shared-memory transforms.  LOGN == 12: 256 threads, radix-16 register passes on a
// swizzled image.  LOGN == 0: any n = 2^logn, radix-2 stages on a linear image.
// Both start and end with a block barrier; outputs are canonical.
// ---------------------------------------------------------------------------------
template <int LOGN>
struct Lay {
    static __device__ __forceinline__ u32 at(u32 e) { return e; }
};
template <>
struct Lay<12> {
    static __device__ __forceinline__ u32 at(u32 e) { return swz(e); }
};

// n = 4096 transforms on the swizzled image by 4096 >> NB threads (NB = 3: 512 threads x 8
// values, four 3-stage passes; NB = 4: 256 threads x 16 values, three 4-stage passes).
// The first forward / last inverse pass takes its (thread-uniform) twiddles from `head`.
constexpr int kNB = 3;                       // values-per-thread exponent used by the fused kernels
constexpr int kThreads12 = 4096 >> kNB;

// CANON = false leaves the outputs lazily reduced (any u64 below the transform's growth bound): enough for a
// consumer that feeds them to a Montgomery REDC, which accepts any x < 2^64 against a canonical multiplier.
// `twm` serves the middle passes (table indices 8 .. 511: 8 KiB, staged in shared memory by the TMA kernel),
// `tw` the pass whose 7 twiddles per thread are all different (indices >= 512, coalesced 16-byte global loads).
#if defined(EXB_LAB) && defined(EXB_SHFL_STAGE) && !defined(EXB_HOST_EMUL)
// Lab (EXB_LAB_FLAGS=-DEXB_SHFL_STAGE sh tools/lab_build.sh; profiles/r02_shuffle_stage_ab.json): the S = 3 <-> S = 0
// regrouping is an 8 x 8 transpose inside groups of 8 consecutive lanes; done with warp shuffles (three xor stages,
// four 64-bit exchanges each) instead of a shared-memory round trip + block barrier.  Measured slower; not shipped.
__device__ __forceinline__ void transpose8_shfl(u64 (&v)[8]) {
    const u32 lane = threadIdx.x & 31u;
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            if (a & s) continue;
            const u64 send = up ? v[a] : v[a + s];
            const u64 recv = __shfl_xor_sync(0xffffffffu, send, s);
            if (up) v[a] = recv; else v[a + s] = recv;
        }
    }
}
#define EXB_REGROUP_F(v, sm, t) transpose8_shfl(v)
#define EXB_REGROUP_I(v, sm, t) transpose8_shfl(v)
#else
#define EXB_REGROUP_F(v, sm, t) store_vals<3, 3>(v, sm, t); __syncthreads(); load_vals<3, 0>(v, sm, t)
#define EXB_REGROUP_I(v, sm, t) store_vals<3, 0>(v, sm, t); __syncthreads(); load_vals<3, 3>(v, sm, t)
#endif
// `keep` (NB = 3): leave the outputs -- elements 8t .. 8t+7, the last pass's own layout -- in the caller's registers
// instead of storing them to the image: callers that go on element-wise skip a shared-memory round trip and a barrier.
template <int NB, int LAZY, bool CANON = true, class TWT = const Tw *, class TWM = const Tw *>
__device__ __forceinline__ void fwd_body12(u64 *sm, const TWT tw, const TWM twm, const TwHead &head,
                                           const LazyC &c, const u32 t, u64 *keep = nullptr, const u64 *feed = nullptr) {
    u64 v[1 << NB];
    if constexpr (NB == 4) {
        load_vals<4, 8>(v, sm, t); fwd_pass<12, 8, 4, LAZY>(v, head, t, c); store_vals<4, 8>(v, sm, t);
        __syncthreads();
        load_vals<4, 4>(v, sm, t); fwd_pass<12, 4, 4, LAZY>(v, tw, t, c); store_vals<4, 4>(v, sm, t);
        __syncthreads();
        load_vals<4, 0>(v, sm, t); fwd_pass<12, 0, 4, LAZY>(v, tw, t, c);
    } else {
        if (feed) {                        // inputs from the caller's registers: elements t + 512 k (first pass layout)
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = feed[k];
        } else {
            load_vals<3, 9>(v, sm, t);
        }
        fwd_pass<12, 9, 3, LAZY>(v, head, t, c); store_vals<3, 9>(v, sm, t);
        __syncthreads();
        load_vals<3, 6>(v, sm, t); fwd_pass<12, 6, 3, LAZY>(v, twm, t, c); store_vals<3, 6>(v, sm, t);
        __syncthreads();
        load_vals<3, 3>(v, sm, t); fwd_pass<12, 3, 3, LAZY>(v, twm, t, c);
        EXB_REGROUP_F(v, sm, t);
        fwd_pass<12, 0, 3, LAZY>(v, tw, t, c);
    }
    if constexpr (CANON) {
#pragma unroll
        for (int k = 0; k < (1 << NB); k++) v[k] = fwd_final<LAZY>(v[k], c);
    }
    if (keep) {
#pragma unroll
        for (int k = 0; k < (1 << NB); k++) keep[k] = v[k];
    } else {
        store_vals<NB, 0>(v, sm, t);
    }
}

// Inputs < 4q (LAZY >= 1) or < 2q (LAZY == 0); outputs canonical.
// `feed` (NB = 3): the inputs -- elements 8t .. 8t+7, the first pass's own layout -- come from the caller's registers
// instead of the image (callers that just computed them element-wise skip a shared-memory round trip and a barrier).
template <int NB, int LAZY, class TWT = const Tw *, class TWM = const Tw *>
__device__ __forceinline__ void inv_body12(u64 *sm, const TWT tw, const TWM twm, const TwHead &head,
                                           const Modulus &mod, const LazyC &c, const u32 t, const u64 *feed = nullptr,
                                           u64 *keep = nullptr) {
    u64 v[1 << NB];
    if constexpr (NB == 4) {
        load_vals<4, 0>(v, sm, t); inv_pass<12, 0, 4, false, LAZY>(v, tw, t, mod, c); store_vals<4, 0>(v, sm, t);
        __syncthreads();
        load_vals<4, 4>(v, sm, t); inv_pass<12, 4, 4, false, LAZY>(v, tw, t, mod, c); store_vals<4, 4>(v, sm, t);
        __syncthreads();
        load_vals<4, 8>(v, sm, t); inv_pass<12, 8, 4, true, LAZY>(v, head, t, mod, c); store_vals<4, 8>(v, sm, t);
    } else {
        if (feed) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = feed[k];
        } else {
            load_vals<3, 0>(v, sm, t);
        }
        inv_pass<12, 0, 3, false, LAZY>(v, tw, t, mod, c);
        EXB_REGROUP_I(v, sm, t);
        inv_pass<12, 3, 3, false, LAZY>(v, twm, t, mod, c); store_vals<3, 3>(v, sm, t);
        __syncthreads();
        load_vals<3, 6>(v, sm, t); inv_pass<12, 6, 3, false, LAZY>(v, twm, t, mod, c); store_vals<3, 6>(v, sm, t);
        __syncthreads();
        load_vals<3, 9>(v, sm, t); inv_pass<12, 9, 3, true, LAZY>(v, head, t, mod, c);
        if (keep) {                        // outputs stay in registers: elements t + 512 k (the last pass's layout)
#pragma unroll
            for (int k = 0; k < 8; k++) keep[k] = v[k];
        } else {
            store_vals<3, 9>(v, sm, t);
        }
    }
}

template <int LAZY, bool CANON = true>
__device__ __forceinline__ void fwd_sm12(u64 *sm, const Tw *__restrict__ tw, const TwHead &head, const Modulus &mod,
                                         u64 *keep = nullptr, const u64 *feed = nullptr) {
    const LazyC c = make_lazyc(mod);
    __syncthreads();                       // (fed from registers too: an earlier transform may have ended without a barrier)
    fwd_body12<kNB, LAZY, CANON>(sm, tw, tw, head, c, threadIdx.x, keep, feed);
    if (!keep) __syncthreads();            // kept outputs: the last pass read only this thread's own elements
}
// With `feed` there is no leading barrier: the first access to the image is this thread's store of its own elements,
// and every earlier use of the image must have ended with a barrier after its last strided access (the transforms do).
template <int LAZY>
__device__ __forceinline__ void inv_sm12(u64 *sm, const Tw *__restrict__ tw, const TwHead &head, const Modulus &mod,
                                         const u64 *feed = nullptr, u64 *keep = nullptr) {
    const LazyC c = make_lazyc(mod);
    if (!feed) __syncthreads();
    inv_body12<kNB, LAZY>(sm, tw, tw, head, mod, c, threadIdx.x, feed, keep);
    __syncthreads();                       // (kept outputs too: the last pass read other threads' elements)
}

// 32-bit transforms of the internal auxiliary basis (n = 4096, 512 threads x 8 values).
// Forward: canonical inputs, outputs in [0, 2p) (lazy: the point-wise Montgomery products accept them).
__device__ __forceinline__ void fwd32_sm(u32 *sm, const Tw32 *__restrict__ tw, const TwHead32 &head, const Mod32 &m) {
    const u32 t = threadIdx.x;
    u32 v[8];
    __syncthreads();
    load_vals32<3, 9>(v, sm, t); fwd_pass32<12, 9, 3>(v, head, t, m); store_vals32<3, 9>(v, sm, t);
    __syncthreads();
    load_vals32<3, 6>(v, sm, t); fwd_pass32<12, 6, 3>(v, tw, t, m); store_vals32<3, 6>(v, sm, t);
    __syncthreads();
    load_vals32<3, 3>(v, sm, t); fwd_pass32<12, 3, 3>(v, tw, t, m); store_vals32<3, 3>(v, sm, t);
    __syncthreads();
    load_vals32<3, 0>(v, sm, t); fwd_pass32<12, 0, 3>(v, tw, t, m);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = fold32(v[k], m);      // < 25p -> [0, 2p)
    store_vals32<3, 0>(v, sm, t);
    __syncthreads();
}
// Inputs in [0, 4p); outputs canonical.
__device__ __forceinline__ void inv32_sm(u32 *sm, const Tw32 *__restrict__ tw, const TwHead32 &head, const Mod32 &m) {
    const u32 t = threadIdx.x;
    u32 v[8];
    __syncthreads();
    load_vals32<3, 0>(v, sm, t); inv_pass32<12, 0, 3, false>(v, tw, t, m); store_vals32<3, 0>(v, sm, t);
    __syncthreads();
    load_vals32<3, 3>(v, sm, t); inv_pass32<12, 3, 3, false>(v, tw, t, m); store_vals32<3, 3>(v, sm, t);
    __syncthreads();
    load_vals32<3, 6>(v, sm, t); inv_pass32<12, 6, 3, false>(v, tw, t, m); store_vals32<3, 6>(v, sm, t);
    __syncthreads();
    load_vals32<3, 9>(v, sm, t); inv_pass32<12, 9, 3, true>(v, head, t, m); store_vals32<3, 9>(v, sm, t);
    __syncthreads();
}

// KK small primes at once: the KK independent 8-value butterfly networks of a thread are
// interleaved by the compiler (3x the ILP per barrier of one transform at a time).  Image i
// lives at sm + i * 4096.
// `io` (optional): inputs come from and outputs go to the caller's registers -- in: elements t + 512 k (first pass
// layout), out: elements 8t + k (last pass layout) -- instead of the images' first load / last store.
template <int KK>
__device__ __forceinline__ void fwd32_smK(u32 *sm, const SmallBasis &sb, u32 (*io)[8] = nullptr) {
    const u32 t = threadIdx.x;
    u32 v[KK][8];
    if (!io) __syncthreads();
#define EXB_PASS(S, TW)                                                                               \
    _Pragma("unroll") for (int i = 0; i < KK; i++) load_vals32<3, S>(v[i], sm + i * 4096, t);          \
    _Pragma("unroll") for (int i = 0; i < KK; i++) fwd_pass32<12, S, 3>(v[i], sb.TW[i], t, sb.sc.m[i]);
    if (io) {
#pragma unroll
        for (int i = 0; i < KK; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[i][k] = io[i][k];
        }
#pragma unroll
        for (int i = 0; i < KK; i++) fwd_pass32<12, 9, 3>(v[i], sb.headf[i], t, sb.sc.m[i]);
    } else {
        EXB_PASS(9, headf)
    }
#pragma unroll
    for (int i = 0; i < KK; i++) store_vals32<3, 9>(v[i], sm + i * 4096, t);
    __syncthreads();
    EXB_PASS(6, twf)
#pragma unroll
    for (int i = 0; i < KK; i++) store_vals32<3, 6>(v[i], sm + i * 4096, t);
    __syncthreads();
    EXB_PASS(3, twf)
#pragma unroll
    for (int i = 0; i < KK; i++) store_vals32<3, 3>(v[i], sm + i * 4096, t);
    __syncthreads();
    EXB_PASS(0, twf)
#undef EXB_PASS
#pragma unroll
    for (int i = 0; i < KK; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[i][k] = fold32(v[i][k], sb.sc.m[i]);
        if (io) {
#pragma unroll
            for (int k = 0; k < 8; k++) io[i][k] = v[i][k];
        } else {
            store_vals32<3, 0>(v[i], sm + i * 4096, t);
        }
    }
    if (!io) __syncthreads();
}
// Primes P0 .. P0 + KK - 1 of the basis; image i of the call is prime P0 + i.
// `feed` (optional): the inputs -- elements 8t .. 8t+7, the first pass's layout -- come from the caller's registers;
// the caller guarantees that nobody still reads the images (no leading barrier here).
template <int KK, int P0 = 0>
__device__ __forceinline__ void inv32_smK(u32 *sm, const SmallBasis &sb, const u32 (*feed)[8] = nullptr,
                                          u32 (*keep)[8] = nullptr) {
    const u32 t = threadIdx.x;
    u32 v[KK][8];
    if (!feed) __syncthreads();
#define EXB_PASS(S, LAST, TW)                                                                         \
    _Pragma("unroll") for (int i = 0; i < KK; i++) load_vals32<3, S>(v[i], sm + i * 4096, t);          \
    _Pragma("unroll") for (int i = 0; i < KK; i++) inv_pass32<12, S, 3, LAST>(v[i], sb.TW[P0 + i], t, sb.sc.m[P0 + i]); \
    _Pragma("unroll") for (int i = 0; i < KK; i++) store_vals32<3, S>(v[i], sm + i * 4096, t);         \
    __syncthreads();
    if (feed) {
#pragma unroll
        for (int i = 0; i < KK; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[i][k] = feed[i][k];
        }
#pragma unroll
        for (int i = 0; i < KK; i++) inv_pass32<12, 0, 3, false>(v[i], sb.twi[P0 + i], t, sb.sc.m[P0 + i]);
#pragma unroll
        for (int i = 0; i < KK; i++) store_vals32<3, 0>(v[i], sm + i * 4096, t);
        __syncthreads();
    } else {
        EXB_PASS(0, false, twi)
    }
    EXB_PASS(3, false, twi)
    EXB_PASS(6, false, twi)
    if (keep) {                            // outputs stay in registers: elements t + 512 k (the last pass's layout)
#pragma unroll
        for (int i = 0; i < KK; i++) load_vals32<3, 9>(v[i], sm + i * 4096, t);
#pragma unroll
        for (int i = 0; i < KK; i++) inv_pass32<12, 9, 3, true>(v[i], sb.headi[P0 + i], t, sb.sc.m[P0 + i]);
#pragma unroll
        for (int i = 0; i < KK; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) keep[i][k] = v[i][k];
        }
        __syncthreads();                   // the images may be rewritten once every thread has read its inputs
    } else {
        EXB_PASS(9, true, headi)
    }
#undef EXB_PASS
}

template <int LOGN, bool CANON = true>
__device__ __forceinline__ void fwd_sm(u64 *sm, const Tw *__restrict__ tw, const TwHead &head, const Modulus &mod,
                                       u32 logn, u64 *keep = nullptr, const u64 *feed = nullptr) {
    const u64 q = mod.m, q2 = mod.two_m;
    if constexpr (LOGN == 12) {
        static_assert(kNB == 3, "kept outputs assume 8 values per thread");
        if (mod.lazy == 2) fwd_sm12<2, CANON>(sm, tw, head, mod, keep, feed);
        else if (mod.lazy == 1) fwd_sm12<1, CANON>(sm, tw, head, mod, keep, feed);
        else fwd_sm12<0, CANON>(sm, tw, head, mod, keep, feed);
    } else {
        // generic path: two butterfly stages per shared-memory round trip (radix 4 on four register values),
        // one radix-2 stage first when logn is odd
        const u32 n = 1u << logn;
        u32 len = n, m = 1;
        if (logn & 1u) {
            len >>= 1;
            __syncthreads();
            for (u32 b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
                u64 *x = sm + b;                                     // m = 1: one block, j = b
                ct_bfly(x[0], x[len], tw[1], q, q2);
            }
            m = 2;
        }
        for (; m < n; m <<= 2) {
            const u32 L = len >> 1, H = L >> 1;                      // stage distances L and H
            __syncthreads();
            for (u32 b = threadIdx.x; b < (n >> 2); b += blockDim.x) {
                const u32 i = b / H, j = b - i * H;
                u64 *x = sm + 2 * i * L + j;
                u64 v0 = x[0], v1 = x[H], v2 = x[L], v3 = x[L + H];
                const Tw w = tw[m + i];
                ct_bfly(v0, v2, w, q, q2);
                ct_bfly(v1, v3, w, q, q2);
                ct_bfly(v0, v1, tw[2 * m + 2 * i], q, q2);
                ct_bfly(v2, v3, tw[2 * m + 2 * i + 1], q, q2);
                x[0] = v0; x[H] = v1; x[L] = v2; x[L + H] = v3;
            }
            len >>= 2;
        }
        __syncthreads();
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) sm[e] = reduce4(sm[e], q, q2);
        __syncthreads();
    }
}

// Inputs in [0, 2q).
template <int LOGN>
__device__ __forceinline__ void inv_sm(u64 *sm, const Tw *__restrict__ tw, const TwHead &head, const Modulus &mod,
                                       u32 logn, const u64 *feed = nullptr, u64 *keep = nullptr) {
    const u64 q = mod.m, q2 = mod.two_m;
    if constexpr (LOGN == 12) {
        if (mod.lazy == 2) inv_sm12<2>(sm, tw, head, mod, feed, keep);
        else if (mod.lazy == 1) inv_sm12<1>(sm, tw, head, mod, feed, keep);
        else inv_sm12<0>(sm, tw, head, mod, feed, keep);
    } else {
        // generic path: two stages per shared-memory round trip; one radix-2 stage first when logn is odd;
        // the very last stage folds n^-1 (plan.normalize, ring/ntt.rs:62)
        const u32 n = 1u << logn;
        u32 len = 1, m = n;
        if (logn & 1u) {
            const u32 h = m >> 1;
            __syncthreads();
            for (u32 b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
                u64 *x = sm + 2 * b;
                if (h == 1) {                                         // n = 2
                    const u64 s2 = x[0] + x[1], dd = x[0] - x[1] + q2;
                    x[0] = shoup_lazy(s2, mod.ninv, mod.ninv_s, q);
                    x[1] = shoup_lazy(dd, mod.ninv_w, mod.ninv_w_s, q);
                } else {
                    gs_bfly(x[0], x[1], tw[h + b], q, q2);
                }
            }
            len = 2; m >>= 1;
        }
        for (; m > 1; m >>= 2) {
            const u32 h = m >> 1, h2 = m >> 2;                        // twiddle bases of the two stages
            __syncthreads();
            for (u32 b = threadIdx.x; b < (n >> 2); b += blockDim.x) {
                const u32 i = b / len, j = b - i * len;
                u64 *x = sm + 4 * i * len + j;
                u64 v0 = x[0], v1 = x[len], v2 = x[2 * len], v3 = x[3 * len];
                gs_bfly(v0, v1, tw[h + 2 * i], q, q2);
                gs_bfly(v2, v3, tw[h + 2 * i + 1], q, q2);
                if (h2 == 1) {                                        // last stage
                    const u64 s02 = v0 + v2, d02 = v0 - v2 + q2, s13 = v1 + v3, d13 = v1 - v3 + q2;
                    v0 = shoup_lazy(s02, mod.ninv, mod.ninv_s, q);
                    v2 = shoup_lazy(d02, mod.ninv_w, mod.ninv_w_s, q);
                    v1 = shoup_lazy(s13, mod.ninv, mod.ninv_s, q);
                    v3 = shoup_lazy(d13, mod.ninv_w, mod.ninv_w_s, q);
                } else {
                    const Tw w = tw[h2 + i];
                    gs_bfly(v0, v2, w, q, q2);
                    gs_bfly(v1, v3, w, q, q2);
                }
                x[0] = v0; x[len] = v1; x[2 * len] = v2; x[3 * len] = v3;
            }
            len <<= 2;
        }
        __syncthreads();
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) sm[e] = csub(sm[e], q);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// K1 / K2: batched standalone transforms, one polynomial per CTA.
// ---------------------------------------------------------------------------------
template <int LOGN>
__global__ void __launch_bounds__(512)
ntt_fwd_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const Tw *__restrict__ tw,
               const __grid_constant__ TwHead head, const __grid_constant__ Modulus mod, u32 logn) {
    EXB_DYN_SMEM(smem);
    const u32 n = 1u << logn;
    const u64 *src = in + (size_t)blockIdx.x * n;
    u64 *dst = out + (size_t)blockIdx.x * n;
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) smem[Lay<LOGN>::at(e)] = ld_stream(src + e);
    fwd_sm<LOGN>(smem, tw, head, mod, logn);
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) dst[e] = smem[Lay<LOGN>::at(e)];
}

template <int LOGN>
__global__ void __launch_bounds__(512)
ntt_inv_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const Tw *__restrict__ tw,
               const __grid_constant__ TwHead head, const __grid_constant__ Modulus mod, u32 logn) {
    EXB_DYN_SMEM(smem);
    const u32 n = 1u << logn;
    const u64 *src = in + (size_t)blockIdx.x * n;
    u64 *dst = out + (size_t)blockIdx.x * n;
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) smem[Lay<LOGN>::at(e)] = ld_stream(src + e);
    inv_sm<LOGN>(smem, tw, head, mod, logn);
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) dst[e] = smem[Lay<LOGN>::at(e)];
}

// ---------------------------------------------------------------------------------
// K1 / K2 fast path (n = 4096): persistent CTAs, one polynomial in flight per CTA plus
// the next one being prefetched into the other shared-memory buffer with cp.async
// (L2 -> smem, no registers, no L1 allocation, so L1 keeps the twiddle tables).
// First-pass twiddles come from the kernel-parameter constant bank (TwHead).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
#ifndef EXB_HOST_EMUL
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#else
    memcpy(smem_dst, gsrc, 16);
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef EXB_HOST_EMUL
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef EXB_HOST_EMUL
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// 16-byte chunk c of a polynomial (elements 2c, 2c+1) lives at chunk swz16(c) of the image.
__device__ __forceinline__ u32 swz16(u32 c) { return c ^ ((c >> 3) & 7u); }

template <int THREADS>
__device__ __forceinline__ void prefetch_poly(u64 *buf, const u64 *src) {
#pragma unroll
    for (int i = 0; i < 2048 / THREADS; i++) {
        const u32 c = i * THREADS + threadIdx.x;
        cp_async16(buf + 2 * swz16(c), src + 2 * c);
    }
}

template <int THREADS>
__device__ __forceinline__ void store_poly(u64 *dst, const u64 *buf) {
#pragma unroll
    for (int i = 0; i < 2048 / THREADS; i++) {
        const u32 c = i * THREADS + threadIdx.x;
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(buf + 2 * swz16(c));
        *reinterpret_cast<ulonglong2 *>(dst + 2 * c) = v;
    }
}

// NB = 3: 512 threads x 8 values (<= 64 registers, 2 CTAs = 32 warps per SM);
// NB = 4: 256 threads x 16 values.  DBG (lab builds, env EXB_NTT_DBG): 1 = copy only, 2 = transforms without any
// global traffic, 3 = every twiddle load hits the same 16 table entries.
#ifdef EXB_LAB
struct TwMasked {
    const Tw *p;
    __device__ __forceinline__ Tw operator[](u32 i) const { return p[i & 15u]; }
};
#endif
template <bool FWD, int LAZY, int NB, int DBG = 0>
__global__ void __launch_bounds__(4096 >> NB, NB == 3 ? 2 : 2)
ntt12_persist_kernel(const u64 *in, u64 *out, const Tw *__restrict__ tw, const __grid_constant__ TwHead head,
                     const __grid_constant__ Modulus mod, u32 count) {
    EXB_DYN_SMEM(smem);
    constexpr u32 n = 4096;
    constexpr int THREADS = 4096 >> NB;
    const LazyC c = make_lazyc(mod);
    const u32 t = threadIdx.x;
    u32 poly = blockIdx.x;
    if (poly >= count) return;
    u32 cur = 0;
    if (DBG != 2) prefetch_poly<THREADS>(smem, in + (size_t)poly * n);
    cp_async_commit();
    for (; poly < count; poly += gridDim.x) {
        u64 *buf = smem + cur * n;
        const u32 next = poly + gridDim.x;
        if (DBG != 2 && next < count) prefetch_poly<THREADS>(smem + (cur ^ 1) * n, in + (size_t)next * n);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (DBG != 1) {
#ifdef EXB_LAB
            if (DBG == 3) {
                const TwMasked twm{tw};
                if (FWD) fwd_body12<NB, LAZY>(buf, twm, twm, head, c, t);
                else inv_body12<NB, LAZY>(buf, twm, twm, head, mod, c, t);
            } else
#endif
            if (FWD) fwd_body12<NB, LAZY>(buf, tw, tw, head, c, t);
            else inv_body12<NB, LAZY>(buf, tw, tw, head, mod, c, t);
        }
        __syncthreads();
        if (DBG != 2) store_poly<THREADS>(out + (size_t)poly * n, buf);
        __syncthreads();      // buf is the prefetch target of the next iteration
        cur ^= 1;
    }
    cp_async_wait<0>();
}


#ifndef EXB_HOST_EMUL
// ---------------------------------------------------------------------------------
// K1 / K2 with the Tensor Memory Accelerator (n = 4096).  One elected thread moves whole polynomials:
//   * load : cp.async.bulk.tensor.2d of a [256 rows][128 B] box with SWIZZLE_128B -- the hardware swizzle IS the
//            16-byte XOR swizzle of the shared-memory image (swz16), so a polynomial arrives transform-ready with
//            one instruction, completion signalled on an mbarrier (no per-thread copies, no address arithmetic);
//   * store: cp.async.bulk.tensor.2d shared -> global straight out of the image (bulk async-group);
//   * twiddles: table entries 0 .. 511 (the middle passes; 8 KiB) are staged into shared memory once per CTA with a
//            1-D cp.async.bulk; the first pass reads its 7 entries from the constant bank and the last pass keeps its
//            per-thread coalesced 16-byte loads (every entry is used by exactly one thread there).
// Three image buffers rotate: poly i is transformed while poly i+1 is landing and poly i-1 is draining, with ONE
// block barrier per polynomial (the cp.async version needs three plus per-thread LDGSTS / LDS / STG).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, u64 *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr u32 kTmaStages = 3;
constexpr u32 kTwStaged = 512;                                 // twiddle entries kept in shared memory
constexpr size_t kTmaSmemBytes = 1024 + kTmaStages * 4096 * 8 + kTwStaged * sizeof(Tw) + 64;

template <bool FWD, int LAZY, bool SMEM_TW = true>
__global__ void __launch_bounds__(512, 2)
ntt12_tma_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                 const Tw *__restrict__ tw, const __grid_constant__ TwHead head, const __grid_constant__ Modulus mod,
                 u32 count) {
    extern __shared__ unsigned char raw_smem[];
    // SWIZZLE_128B keys on address bits 7..9: the images must start on a 1024-byte boundary
    u64 *bufs = reinterpret_cast<u64 *>((reinterpret_cast<uintptr_t>(raw_smem) + 1023) & ~(uintptr_t)1023);
    Tw *stw = reinterpret_cast<Tw *>(bufs + kTmaStages * 4096);
    u64 *full = reinterpret_cast<u64 *>(stw + kTwStaged);      // full[stage], then the twiddle barrier
    u64 *twbar = full + kTmaStages;
    const LazyC c = make_lazyc(mod);
    const u32 t = threadIdx.x;
    if (blockIdx.x >= count) return;
    if (t == 0) {
        for (u32 s = 0; s < kTmaStages; s++) mbar_init(full + s, 1);
        mbar_init(twbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(twbar, kTwStaged * (u32)sizeof(Tw));
        bulk_load_1d(stw, tw, kTwStaged * (u32)sizeof(Tw), twbar);
        for (u32 s = 0; s < 2; s++) {
            const u32 p = blockIdx.x + s * gridDim.x;
            if (p < count) {
                mbar_expect_tx(full + s, 4096 * 8);
                tma_load_2d(bufs + s * 4096, &tin, 0, (int)(p * 256), full + s);
            }
        }
    }
    mbar_wait(twbar, 0);
    u32 it = 0;
    for (u32 poly = blockIdx.x; poly < count; poly += gridDim.x, it++) {
        const u32 b = it % kTmaStages;
        u64 *buf = bufs + b * 4096;
        mbar_wait(full + b, (it / kTmaStages) & 1u);
        if (SMEM_TW) {
            if (FWD) fwd_body12<3, LAZY>(buf, tw, stw, head, c, t);
            else inv_body12<3, LAZY>(buf, tw, stw, head, mod, c, t);
        } else {
            if (FWD) fwd_body12<3, LAZY>(buf, tw, tw, head, c, t);
            else inv_body12<3, LAZY>(buf, tw, tw, head, mod, c, t);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the TMA store
        __syncthreads();
        if (t == 0) {
            tma_store_2d(&tout, 0, (int)(poly * 256), buf);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            const u32 next = poly + 2 * gridDim.x;
            if (next < count) {
                // the buffer poly i+2 lands in was drained by the store of poly i-1: all but the newest group done reading
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                const u32 nb = (it + 2) % kTmaStages;
                mbar_expect_tx(full + nb, 4096 * 8);
                tma_load_2d(bufs + nb * 4096, &tin, 0, (int)(next * 256), full + nb);
            }
        }
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
#endif  // EXB_HOST_EMUL

// ---------------------------------------------------------------------------------
// K3: element-wise ring ops on canonical residues.
// ---------------------------------------------------------------------------------
__global__ void poly_op_kernel(Modulus mod, int op, const u64 *__restrict__ a, const u64 *__restrict__ b,
                               u64 scalar, u64 *__restrict__ out, size_t words) {
    const u64 m = mod.m;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words;
         i += (size_t)gridDim.x * blockDim.x) {
        const u64 x = a[i];
        u64 r;
        switch (op) {
            case OP_ADD: r = mod_add(x, b[i], m); break;
            case OP_SUB: r = mod_sub(x, b[i], m); break;
            case OP_NEG: r = mod_neg(x, m); break;
            case OP_MUL: {
                const u64 t = csub(mont_mul_lazy(x, b[i], m, mod.minv_neg), m);
                r = csub(mont_mul_lazy(t, mod.r2_mod, m, mod.minv_neg), m);
                break;
            }
            case OP_SCALAR_MUL: {
                const u64 t = csub(mont_mul_lazy(x, scalar, m, mod.minv_neg), m);
                r = csub(mont_mul_lazy(t, mod.r2_mod, m, mod.minv_neg), m);
                break;
            }
            default: r = csub(mont_mul_lazy(x, mod.r2_mod, m, mod.minv_neg), m); break;  // OP_TO_MONT
        }
        out[i] = r;
    }
}

// ---------------------------------------------------------------------------------
// K4: lift.  One CTA per input polynomial: INTT_q -> centred reduce mod p_j ->
// NTT_pj.  Right-hand-side polynomials are written in Montgomery form (x * 2^64)
// in every base so the tensor kernel needs one REDC per product.
// ext layout: [pair][side][limb][comp][1+A][n]; slot 0 of side 0 is unused.
// ---------------------------------------------------------------------------------
template <int LOGN>
__global__ void __launch_bounds__(LOGN == 12 ? kThreads12 : 256, LOGN == 12 ? 2 : 1)
lift_kernel(const __grid_constant__ DeviceParams P, u32 d, u32 need_lhs, u32 need_rhs, const u64 *__restrict__ ct1,
            const u64 *__restrict__ ct2, u64 *__restrict__ ext) {
    EXB_DYN_SMEM(smem);
    const u32 n = P.n, A = P.num_aux;
    u64 *coef = smem, *work = smem + n;
    const u32 idx = blockIdx.x;
    const u32 comp = idx & 1u;
    const u32 limb = (idx >> 1) % d;
    const u32 side = (idx / (2 * d)) & 1u;
    const size_t pair = idx / (4 * d);
    if (!(((side ? need_rhs : need_lhs) >> limb) & 1u)) return;      // no live product reads this limb
    const u64 *src = (side ? ct2 : ct1) + ((pair * d + limb) * 2 + comp) * (size_t)n;
    u64 *dst = ext + ((((pair * 2 + side) * d + limb) * 2 + comp) * (size_t)(1 + A)) * n;
    const Modulus &mq = P.mod[0];

    for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
        const u64 x = src[e];
        coef[Lay<LOGN>::at(e)] = x;
        if (side) dst[e] = shoup(x, mq.r_mod, mq.r_mod_s, mq.m);
    }
    inv_sm<LOGN>(coef, P.twi[0], P.headi[0], mq, P.logn);
    for (u32 j = 0; j < A; j++) {
        const Modulus &mp = P.mod[1 + j];
        for (u32 e = threadIdx.x; e < n; e += blockDim.x)
            work[Lay<LOGN>::at(e)] =
                ext_centered(coef[Lay<LOGN>::at(e)], mq.m, P.sc.half_q, mp.m, mp.mu);
        fwd_sm<LOGN>(work, P.twf[1 + j], P.headf[1 + j], mp, P.logn);
        u64 *o = dst + (size_t)(1 + j) * n;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            u64 x = work[Lay<LOGN>::at(e)];
            if (side) x = shoup(x, mp.r_mod, mp.r_mod_s, mp.m);
            o[e] = x;
        }
    }
}

// ---------------------------------------------------------------------------------
// K3+K5: tensor + scale.  One CTA per (pair, product, component): point-wise
// product in every base, INTT per base, then hps_scale per coefficient.
// Components 0/1 leave round(p t/q) mod q (coefficient domain) in r01;
// component 2 leaves its balanced gadget digits.
// ---------------------------------------------------------------------------------
template <int LOGN, typename DigT>
__global__ void __launch_bounds__(LOGN == 12 ? kThreads12 : 256, LOGN == 12 ? 2 : 1)
tensor_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
              const u64 *__restrict__ ct1, const u64 *__restrict__ ext, u64 *__restrict__ r01,
              DigT *__restrict__ digits, u32 raw3) {
    EXB_DYN_SMEM(smem);
    const u32 n = P.n, A = P.num_aux, d = M.d, NP = M.num_products;
    const u32 idx = blockIdx.x;
    const u32 comp = idx % 3u;
    const u32 prod = (idx / 3u) % NP;
    const size_t pair = idx / (3u * NP);
    const u32 li = M.prod_i[prod], lj = M.prod_j[prod];
    const size_t bases = 1 + A;

    for (u32 b = 0; b <= A; b++) {
        const Modulus &mb = P.mod[b];
        const u64 *l0, *l1;
        if (b == 0) {
            l0 = ct1 + ((pair * d + li) * 2) * (size_t)n;
            l1 = l0 + n;
        } else {
            l0 = ext + ((((pair * 2 + 0) * d + li) * 2 + 0) * bases + b) * n;
            l1 = l0 + bases * n;
        }
        const u64 *r0 = ext + ((((pair * 2 + 1) * d + lj) * 2 + 0) * bases + b) * n;
        const u64 *r1 = r0 + bases * n;
        u64 *buf = smem + (size_t)b * n;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            u64 v;
            if (comp == 0) v = mont_mul_lazy(l0[e], r0[e], mb.m, mb.minv_neg);
            else if (comp == 2) v = mont_mul_lazy(l1[e], r1[e], mb.m, mb.minv_neg);
            else v = mont_mul2_lazy(l0[e], r1[e], l1[e], r0[e], mb.m, mb.minv_neg);
            buf[Lay<LOGN>::at(e)] = v;
        }
        inv_sm<LOGN>(buf, P.twi[b], P.headi[b], mb, P.logn);
    }

    const u64 *ba = smem, *b0 = smem + n, *b1 = smem + 2 * (size_t)n;
    if (comp < 2 || raw3) {                // raw3 (bfv_mul_no_relin): all three scaled components, [pair][prod][3][n]
        u64 *o = r01 + ((pair * NP + prod) * (raw3 ? 3 : 2) + comp) * (size_t)n;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            const u32 p = Lay<LOGN>::at(e);
            o[e] = hps_scale_coeff(ba[p], b0[p], A == 2 ? b1[p] : 0, P.sc);
        }
    } else {
        const u32 G = P.gadget_digits;
        DigT *o = digits + ((pair * NP + prod) * (size_t)G) * n;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            const u32 p = Lay<LOGN>::at(e);
            const u64 c2 = hps_scale_coeff(ba[p], b0[p], A == 2 ? b1[p] : 0, P.sc);
            i64 rem = center_i64(c2, P.sc.q, P.sc.half_q);
            if (P.gadget_log2) {
                for (u32 g = 0; g < G; g++) o[(size_t)g * n + e] = (DigT)gadget_digit_pow2(rem, P.gadget_log2);
            } else {
                for (u32 g = 0; g < G; g++)
                    o[(size_t)g * n + e] = (DigT)gadget_digit_general(rem, (i64)P.gadget_base);
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// 8 consecutive coefficients per thread with 128-bit accesses (n = 4096 kernels): global rows
// are read/written as whole 64-byte (u64) / 32-byte (u32) / 16-byte (int16) pieces, and the
// swizzled shared-memory images keep every 16-byte pair together and conflict-free.
// ---------------------------------------------------------------------------------
struct alignas(16) Vec16 { u32 w[4]; };

// One 256-bit access per thread (sm_100: LDG.E.256 / STG.E.256): a thread's 32-byte piece is exactly one DRAM/L2
// sector, so operand rows -- read once per CTA -- need no L1 allocation and leave L1 to the twiddle tables.
#define EXB_LD256 "ld.global.nc.L1::no_allocate"
__device__ __forceinline__ void ldg_u64x4(const u64 *p, u64 *v) {   // 4 consecutive u64, 32-byte aligned
#ifndef EXB_HOST_EMUL
    asm volatile(EXB_LD256 ".v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
#else
    for (int i = 0; i < 4; i++) v[i] = p[i];
#endif
}
__device__ __forceinline__ void stg_u64x4(u64 *p, const u64 *v) {
#ifndef EXB_HOST_EMUL
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
#else
    for (int i = 0; i < 4; i++) p[i] = v[i];
#endif
}
__device__ __forceinline__ void ldg_u32x8(const u32 *p, u32 *v) {   // 8 consecutive u32, 32-byte aligned
#ifndef EXB_HOST_EMUL
    asm volatile(EXB_LD256 ".v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
#else
    for (int i = 0; i < 8; i++) v[i] = p[i];
#endif
}
__device__ __forceinline__ void stg_u32x8(u32 *p, const u32 *v) {
#ifndef EXB_HOST_EMUL
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
#else
    for (int i = 0; i < 8; i++) p[i] = v[i];
#endif
}
// swizzled u64 image: elements e0 .. e0+3 (e0 % 4 == 0) are two 16-byte chunks
__device__ __forceinline__ void lds_u64x4(const u64 *sm, u32 e0, u64 *v) {
    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(sm + swz(e0));
    const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(sm + swz(e0 + 2));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void sts_u64x4(u64 *sm, u32 e0, const u64 *v) {
    ulonglong2 a, b;
    a.x = v[0]; a.y = v[1]; b.x = v[2]; b.y = v[3];
    *reinterpret_cast<ulonglong2 *>(sm + swz(e0)) = a;
    *reinterpret_cast<ulonglong2 *>(sm + swz(e0 + 2)) = b;
}
// swizzled u32 image: elements e0 .. e0+7 (e0 % 8 == 0) are two 16-byte chunks
__device__ __forceinline__ void lds_u32x8(const u32 *sm, u32 e0, u32 *v) {
    const Vec16 a = *reinterpret_cast<const Vec16 *>(sm + swz32(e0)), b = *reinterpret_cast<const Vec16 *>(sm + swz32(e0 + 4));
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] = a.w[i]; v[4 + i] = b.w[i]; }
}
__device__ __forceinline__ void sts_u32x8(u32 *sm, u32 e0, const u32 *v) {
    Vec16 a, b;
#pragma unroll
    for (int i = 0; i < 4; i++) { a.w[i] = v[i]; b.w[i] = v[4 + i]; }
    *reinterpret_cast<Vec16 *>(sm + swz32(e0)) = a;
    *reinterpret_cast<Vec16 *>(sm + swz32(e0 + 4)) = b;
}
// 8 consecutive gadget digits (int16: one 16-byte piece, int32: two)
template <typename DigT>
__device__ __forceinline__ void ldg_dig8(const DigT *p, i64 *acc) {
    if constexpr (sizeof(DigT) == 1) {                     // 8 digits = one 8-byte piece
        const u64 a = *reinterpret_cast<const u64 *>(p);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += (i64)(int8_t)((a >> (8 * i)) & 0xffu);
    } else if constexpr (sizeof(DigT) == 2) {
        const Vec16 a = *reinterpret_cast<const Vec16 *>(p);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            acc[2 * i] += (i64)(int16_t)(a.w[i] & 0xffffu);
            acc[2 * i + 1] += (i64)(int16_t)(a.w[i] >> 16);
        }
    } else {
        const Vec16 a = reinterpret_cast<const Vec16 *>(p)[0], b = reinterpret_cast<const Vec16 *>(p)[1];
#pragma unroll
        for (int i = 0; i < 4; i++) { acc[i] += (i64)(int32_t)a.w[i]; acc[4 + i] += (i64)(int32_t)b.w[i]; }
    }
}
template <typename DigT>
__device__ __forceinline__ void stg_dig8(DigT *p, const i64 *d) {
    if constexpr (sizeof(DigT) == 1) {
        u64 a = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) a |= ((u64)d[i] & 0xffu) << (8 * i);
        *reinterpret_cast<u64 *>(p) = a;
    } else if constexpr (sizeof(DigT) == 2) {
        Vec16 a;
#pragma unroll
        for (int i = 0; i < 4; i++) a.w[i] = ((u32)d[2 * i] & 0xffffu) | ((u32)d[2 * i + 1] << 16);
        *reinterpret_cast<Vec16 *>(p) = a;
    } else {
        Vec16 a, b;
#pragma unroll
        for (int i = 0; i < 4; i++) { a.w[i] = (u32)d[i]; b.w[i] = (u32)d[4 + i]; }
        reinterpret_cast<Vec16 *>(p)[0] = a; reinterpret_cast<Vec16 *>(p)[1] = b;
    }
}

// A sum of at most 16 balanced digits (|s| <= 2^35) as a canonical residue.  For q >= 2^36 (`big_q`: the lazy prime
// classes) that is s or s + q; smaller primes take the general reduction.
__device__ __forceinline__ u64 signed_sum_to_mod(i64 s, u64 q, bool big_q) {
    if (big_q) return (u64)(s + ((s >> 63) & (i64)q));
    const u64 mag = s < 0 ? (u64)(-s) : (u64)s;
    const u64 r = mag < q ? mag : mag % q;
    return (s < 0 && r) ? q - r : r;
}

// Sum over the products of output limb k of digit plane g, for this thread's 8 coefficients.  int8 planes are summed
// as packed bytes: biased to unsigned (x ^ 0x80) and accumulated in 16-bit lanes, four coefficients per 64-bit word.
template <typename DigT>
__device__ __forceinline__ void digit_sums8(const DigT *__restrict__ digits, const MulPlan &M, size_t pair, u32 NP, u32 G,
                                            u32 g, u32 n, u32 e0, u32 k, u32 i_lo, u32 i_hi, i64 *ds) {
    if constexpr (sizeof(DigT) == 1) {
        const u64 bias = 0x8080808080808080ull, lanes = 0x00FF00FF00FF00FFull;
        u64 even = 0, odd = 0;
        for (u32 i = i_lo; i <= i_hi; i++) {
            const size_t pr = (size_t)M.prod_of[i][k - i];
                        const u64 x = ld_stream(reinterpret_cast<const u64 *>(digits + ((pair * NP + pr) * G + g) * n + e0)) ^ bias;
            even += x & lanes;
            odd += (x >> 8) & lanes;
        }
        const i64 off = 128 * (i64)(i_hi - i_lo + 1);          // at most 16 products: 16 * 255 fits a 16-bit lane
#pragma unroll
        for (int j = 0; j < 4; j++) {
            ds[2 * j] = (i64)((even >> (16 * j)) & 0xffffu) - off;
            ds[2 * j + 1] = (i64)((odd >> (16 * j)) & 0xffffu) - off;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) ds[j] = 0;
        for (u32 i = i_lo; i <= i_hi; i++) {
            const size_t pr = (size_t)M.prod_of[i][k - i];
            ldg_dig8<DigT>(digits + ((pair * NP + pr) * G + g) * n + e0, ds);
        }
    }
}

// ---------------------------------------------------------------------------------
// K4' / K5': lift and tensor+scale on the internal 27-bit auxiliary basis (n = 4096 only;
// see ntt32_core.cuh for why this is result-identical).  Layouts:
//   ext_s : [pair][side][limb][comp][K][n] u32   both operands mod the small primes, in [0, 2p)
// No operand is converted to Montgomery form: the point-wise REDC leaves a*b*R^-1 and R is folded into the
// n^-1 constants of the inverse transforms that follow (SmallBasis::mq_r for q, Mod32::ninv for the small primes).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads12, 2)
lift32_kernel(const __grid_constant__ DeviceParams P, u32 d, u32 need_lhs, u32 need_rhs, const u64 *__restrict__ ct1,
              const u64 *__restrict__ ct2, u32 *__restrict__ ext_s) {
    EXB_DYN_SMEM(smem);
    constexpr u32 n = 4096;
    const u32 K = P.sb.K;
    u64 *coef = smem;
    u32 *work = reinterpret_cast<u32 *>(smem + n);
    const u32 idx = blockIdx.x;
    const u32 comp = idx & 1u;
    const u32 limb = (idx >> 1) % d;
    const u32 side = (idx / (2 * d)) & 1u;
    const size_t pair = idx / (4 * d);
    if (!(((side ? need_rhs : need_lhs) >> limb) & 1u)) return;      // no live product reads this limb
    const u64 *src = (side ? ct2 : ct1) + ((pair * d + limb) * 2 + comp) * (size_t)n;
    u32 *ds = ext_s + ((((pair * 2 + side) * d + limb) * 2 + comp) * (size_t)K) * n;
    const Modulus &mq = P.mod[0];
    const u32 e0 = 8 * threadIdx.x;
    u64 x[8];
    ldg_u64x4(src + e0, x); ldg_u64x4(src + e0 + 4, x + 4);
    if (K == 3) {
        // registers all the way: the input row feeds the inverse transform, whose outputs (canonical coefficients
        // t + 512 k, its last pass's layout) are extended to the three small primes in place and feed the forward
        // transforms' first pass; their last pass leaves elements 8t .. 8t+7, which go straight to global memory
        inv_sm<12>(coef, P.twi[0], P.headi[0], mq, 12, x, x);
        u32 y[3][8];
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) y[i][k] = ext32_centered_lazy(x[k], mq.m, P.sc.half_q, P.sb.sc.m[i]);
        }
        fwd32_smK<3>(work, P.sb, y);
#pragma unroll
        for (int i = 0; i < 3; i++) stg_u32x8(ds + (size_t)i * n + e0, y[i]);
        return;
    }
    sts_u64x4(coef, e0, x); sts_u64x4(coef, e0 + 4, x + 4);
    inv_sm<12>(coef, P.twi[0], P.headi[0], mq, 12);
    lds_u64x4(coef, e0, x); lds_u64x4(coef, e0 + 4, x + 4);      // canonical coefficients stay in registers
    for (u32 i = 0; i < K; i++) {
        u32 y[8];
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = ext32_centered_lazy(x[k], mq.m, P.sc.half_q, P.sb.sc.m[i]);
        sts_u32x8(work + (size_t)i * n, e0, y);
    }
    for (u32 i = 0; i < K; i++) fwd32_sm(work + (size_t)i * n, P.sb.twf[i], P.sb.headf[i], P.sb.sc.m[i]);
    for (u32 i = 0; i < K; i++) {
        u32 y[8];
        lds_u32x8(work + (size_t)i * n, e0, y);
        stg_u32x8(ds + (size_t)i * n + e0, y);
    }
}

template <typename DigT>
__global__ void __launch_bounds__(kThreads12, 2)
tensor32_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
                const u64 *__restrict__ ct1, const u64 *__restrict__ ct2, const u32 *__restrict__ ext_s,
                u64 *__restrict__ r01, DigT *__restrict__ digits, u32 only_c2) {
    EXB_DYN_SMEM(smem);
    constexpr u32 n = 4096;
    const u32 K = P.sb.K, d = M.d, NP = M.num_products;
    const u32 idx = blockIdx.x;
    // mode bit 0 (only_c2): components 0 and 1 are produced per output limb by tensor01_kernel;
    // mode bit 1 (raw3, bfv_mul_no_relin): all three scaled components go to r01 as [pair][prod][3][n]
    const u32 raw3 = only_c2 & 2u;
    only_c2 &= 1u;
    const u32 comp = only_c2 ? 2u : idx % 3u;
    const u32 prod = only_c2 ? idx % NP : (idx / 3u) % NP;
    const size_t pair = only_c2 ? idx / NP : idx / (3u * NP);
    const u32 li = M.prod_i[prod], lj = M.prod_j[prod];
    u64 *bq = smem;
    u32 *bs = reinterpret_cast<u32 *>(smem + n);
    const u32 e0 = 8 * threadIdx.x;
    {   // base q: 64-bit Montgomery point-wise + INTT (two halves of 4 coefficients: register budget)
        const Modulus &mb = P.sb.mq_r;               // q with n^-1 * 2^64 as the inverse transform's scaling
        const u64 *l0 = ct1 + ((pair * d + li) * 2) * (size_t)n, *l1 = l0 + n;
        const u64 *r0 = ct2 + ((pair * d + lj) * 2) * (size_t)n, *r1 = r0 + n;
        u64 v[8];                          // the products feed the inverse transform from registers
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 eh = e0 + 4 * h;
            u64 a[4], b[4];
            if (comp == 1) {
                u64 c[4], dd[4];
                ldg_u64x4(l0 + eh, a); ldg_u64x4(r1 + eh, b); ldg_u64x4(l1 + eh, c); ldg_u64x4(r0 + eh, dd);
#pragma unroll
                for (int k = 0; k < 4; k++) v[4 * h + k] = mont_mul2_lazy(a[k], b[k], c[k], dd[k], mb.m, mb.minv_neg);
            } else {
                ldg_u64x4((comp == 0 ? l0 : l1) + eh, a); ldg_u64x4((comp == 0 ? r0 : r1) + eh, b);
#pragma unroll
                for (int k = 0; k < 4; k++) v[4 * h + k] = mont_mul_lazy(a[k], b[k], mb.m, mb.minv_neg);
            }
        }
        inv_sm<12>(bq, P.twi[0], P.headi[0], mb, 12, v);
    }
    u32 pw[3][8];                   // K == 3: the point-wise products feed the inverse transforms from registers
#pragma unroll
    for (u32 i = 0; i < (u32)kMaxSmall; i++) {   // small primes: 32-bit Montgomery point-wise, then the K INTTs together
        if (i >= K) break;
        const Mod32 &m = P.sb.sc.m[i];
        const u32 *l0 = ext_s + ((((pair * 2 + 0) * d + li) * 2) * (size_t)K + i) * n, *l1 = l0 + (size_t)K * n;
        const u32 *r0 = ext_s + ((((pair * 2 + 1) * d + lj) * 2) * (size_t)K + i) * n, *r1 = r0 + (size_t)K * n;
        u32 a[8], b[8], v[8];
        if (comp == 1) {
            u32 c[8], dd[8];
            ldg_u32x8(l0 + e0, a); ldg_u32x8(r1 + e0, b); ldg_u32x8(l1 + e0, c); ldg_u32x8(r0 + e0, dd);
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = mont32_redc_lazy((u64)a[k] * b[k] + (u64)c[k] * dd[k], m.p, m.pinv_neg);
        } else {
            ldg_u32x8((comp == 0 ? l0 : l1) + e0, a); ldg_u32x8((comp == 0 ? r0 : r1) + e0, b);
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = mont32_redc_lazy((u64)a[k] * b[k], m.p, m.pinv_neg);
        }
        if (K == 3) {
#pragma unroll
            for (int k = 0; k < 8; k++) pw[i < 3 ? i : 0][k] = v[k];
        } else {
            sts_u32x8(bs + (size_t)i * n, e0, v);
        }
    }
    if (K == 3) inv32_smK<3>(bs, P.sb, pw);          // bs is untouched so far: no barrier needed before its first store
    else for (u32 i = 0; i < K; i++) inv32_sm(bs + (size_t)i * n, P.sb.twi[i], P.sb.headi[i], P.sb.sc.m[i]);
    const u32 G = P.gadget_digits;
    u64 av[8];
    u32 bv[kMaxSmall][8];
    lds_u64x4(bq, e0, av); lds_u64x4(bq, e0 + 4, av + 4);
#pragma unroll
    for (u32 i = 0; i < (u32)kMaxSmall; i++)
        if (i < K) lds_u32x8(bs + (size_t)i * n, e0, bv[i]);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 b[kMaxSmall];
#pragma unroll
        for (u32 i = 0; i < (u32)kMaxSmall; i++) b[i] = i < K ? bv[i][k] : 0u;
        av[k] = hps_scale32_coeff(av[k], b, P.sc, P.sb.sc);
    }
    if (comp < 2 || raw3) {
        u64 *o01 = r01 + ((pair * NP + prod) * (raw3 ? 3 : 2) + comp) * (size_t)n;
        stg_u64x4(o01 + e0, av); stg_u64x4(o01 + e0 + 4, av + 4);
    } else {
        DigT *od = digits + ((pair * NP + prod) * (size_t)G) * n;
        i64 rem[8];
#pragma unroll
        for (int k = 0; k < 8; k++) rem[k] = center_i64(av[k], P.sc.q, P.sc.half_q);
        if (sizeof(DigT) == 1 && P.gadget_log2 == 8 && G <= 8) {
            // base 256 (the u64 profile): one addition per coefficient yields all eight digits, an 8 x 8 byte
            // transpose turns them into the eight 8-byte plane words of this thread
            u64 w[8], plane[8];
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] = gadget_digits_base256(rem[k]);
            digits_to_planes(w, plane);
#pragma unroll
            for (u32 g = 0; g < 8; g++)
                if (g < G) *reinterpret_cast<u64 *>(od + (size_t)g * n + e0) = plane[g];
            return;
        }
        for (u32 g = 0; g < G; g++) {
            i64 dg[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                dg[k] = P.gadget_log2 ? gadget_digit_pow2(rem[k], P.gadget_log2)
                                      : gadget_digit_general(rem[k], (i64)P.gadget_base);
            stg_dig8<DigT>(od + (size_t)g * n + e0, dg);
        }
    }
}

// tensor01_kernel, small primes P0 and P0 + 1 (those below K): sum the point-wise products of output limb k over its
// products, inverse-transform, and leave this thread's 8 coefficients of each -- elements t + 512 k, the transforms'
// last-pass layout -- in bv[prime].  Every earlier use of the images ended with a barrier after its last read.
template <int P0>
__device__ __forceinline__ void tensor01_small_round(const DeviceParams &P, const u32 *__restrict__ ext_s, u32 *bs,
                                                     size_t pair, u32 d, u32 k, u32 i_lo, u32 i_hi, u32 comp, u32 e0,
                                                     u32 (&bv)[kMaxSmall][8]) {
    constexpr u32 n = 4096;
    const u32 K = P.sb.K;
    const u32 kk = K - P0 >= 2 ? 2u : 1u;
    u32 accs[2][8];                        // the sums feed the inverse transforms from registers
#pragma unroll
    for (u32 r = 0; r < 2; r++) {
        u32 (&acc)[8] = accs[r];
#pragma unroll
        for (int t = 0; t < 8; t++) acc[t] = 0;
        if (r < kk) {
            const u32 pi = P0 + r;
            const Mod32 &m = P.sb.sc.m[pi];
            for (u32 i = i_lo; i <= i_hi; i++) {
                const u32 j = k - i;
                const u32 *l0 = ext_s + ((((pair * 2 + 0) * d + i) * 2) * (size_t)K + pi) * n, *l1 = l0 + (size_t)K * n;
                const u32 *r0 = ext_s + ((((pair * 2 + 1) * d + j) * 2) * (size_t)K + pi) * n, *r1 = r0 + (size_t)K * n;
                u32 a[8], b[8];
                if (comp == 1) {
                    u32 cc[8], dd[8];
                    ldg_u32x8(l0 + e0, a); ldg_u32x8(r1 + e0, b); ldg_u32x8(l1 + e0, cc); ldg_u32x8(r0 + e0, dd);
#pragma unroll
                    for (int t = 0; t < 8; t++) acc[t] += mont32_redc_lazy((u64)a[t] * b[t] + (u64)cc[t] * dd[t], m.p, m.pinv_neg);
                } else {
                    ldg_u32x8(l0 + e0, a); ldg_u32x8(r0 + e0, b);
#pragma unroll
                    for (int t = 0; t < 8; t++) acc[t] += mont32_redc_lazy((u64)a[t] * b[t], m.p, m.pinv_neg);
                }
            }
#pragma unroll
            for (int t = 0; t < 8; t++) acc[t] = fold32(acc[t], m);          // <= 16 * 2p < 2^32 -> [0, 2p)
        }
    }
    if (kk == 2) inv32_smK<2, P0>(bs, P.sb, accs, &bv[P0]);
    else inv32_smK<1, P0>(bs, P.sb, accs, &bv[P0]);
}

// ---------------------------------------------------------------------------------
// K5'' : components 0 and 1 per OUTPUT LIMB (n = 4096, internal basis).  relinearize only needs
//   sum_{i+j=k} [ round(p a_ij / q) + p m_ij ]  (mod q),   a_ij = t_ij mod q (centred),  m_ij = (t_ij - a_ij) / q,
// and sum_ij m_ij = (T_k - sum_ij a_ij) / q with T_k = sum_ij t_ij is LINEAR in the tensor, so the
// small-prime point-wise products are summed in the NTT domain and inverse-transformed once per limb
// instead of once per product; only the base-q transform and the rounding term stay per product
// (the rounding is the non-linear part).  Exact while cnt * q < 2^64 and P' >= 2^5 * cnt * |m|max
// (checked on the host: SmallBasis::max_terms).  Component 2 keeps the per-product kernel: its gadget
// digits are a non-linear function of each product.
// One CTA per (pair, two computed limbs, component in {0, 1}); output r01s [pair][limb][2][n], coefficient domain.
// ---------------------------------------------------------------------------------
// R64: the per-limb sums of rounding terms are kept as i64 (any plaintext modulus) instead of i32 (products per limb
// * (p/2 + 2) < 2^31: SmallBasis::max_terms_r32, every BASELINE config) -- 16 KB more shared memory per CTA.
template <bool R64>
__global__ void __launch_bounds__(kThreads12, 2)
tensor01_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
                const u64 *__restrict__ ct1, const u64 *__restrict__ ct2, const u32 *__restrict__ ext_s,
                u64 *__restrict__ r01s) {
    EXB_DYN_SMEM(smem);
    constexpr u32 n = 4096;
    const u32 K = P.sb.K, d = M.d, NL = M.num_limbs, ND = M.num_duos;
    const u32 comp = blockIdx.x & 1u;
    const u32 duo = (blockIdx.x >> 1) % ND;
    const size_t pair = (blockIdx.x >> 1) / ND;
    i64 *sacc = reinterpret_cast<i64 *>(smem);                 // sum of centred a_ij
    u32 *racc = reinterpret_cast<u32 *>(smem + n);             // sum of signed rounding terms: i32, or i64 rows (R64)
    u64 *bq = smem + n + (R64 ? n : n / 2);                    // base-q work image, later two u32 images per round
    u32 *bs = reinterpret_cast<u32 *>(bq);
    const Modulus &mb = P.sb.mq_r;                   // q with n^-1 * 2^64 as the inverse transform's scaling
    const ScaleConsts &c = P.sc;
    const Scale32Consts &sc = P.sb.sc;
    const u32 e0 = 8 * threadIdx.x;
  // two limbs per CTA (heaviest with lightest: every CTA runs about d + 1 products)
  for (u32 part = 0; part < 2; part++) {
    const u32 limb = part ? M.duo_b[duo] : M.duo_a[duo];
    if (limb == 0xFFu) break;
    const u32 k = M.limb_k[limb];
    const u32 i_lo = k >= d ? k - d + 1 : 0, i_hi = k < d ? k : d - 1;
    // Registers carry the data between the element-wise phases and the transforms: the point-wise products feed the
    // inverse transform (8 consecutive elements per thread), its outputs stay in registers in the last pass's layout
    // (elements t + 512 k), and the per-limb accumulators, the small-prime results and the final combine all use that
    // layout -- slot s of a thread's private accumulator rows belongs to element t + 512 s.

    for (u32 i = i_lo; i <= i_hi; i++) {   // base q: Montgomery point-wise + INTT per product
        const u32 j = k - i;
        const u64 *l0 = ct1 + ((pair * d + i) * 2) * (size_t)n, *l1 = l0 + n;
        const u64 *r0 = ct2 + ((pair * d + j) * 2) * (size_t)n, *r1 = r0 + n;
        u64 v[8];                          // the products feed the inverse transform from registers
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 eh = e0 + 4 * h;
            u64 a[4], b[4];
            if (comp == 1) {
                u64 cc[4], dd[4];
                ldg_u64x4(l0 + eh, a); ldg_u64x4(r1 + eh, b); ldg_u64x4(l1 + eh, cc); ldg_u64x4(r0 + eh, dd);
#pragma unroll
                for (int t = 0; t < 4; t++) v[4 * h + t] = mont_mul2_lazy(a[t], b[t], cc[t], dd[t], mb.m, mb.minv_neg);
            } else {
                ldg_u64x4(l0 + eh, a); ldg_u64x4(r0 + eh, b);
#pragma unroll
                for (int t = 0; t < 4; t++) v[4 * h + t] = mont_mul_lazy(a[t], b[t], mb.m, mb.minv_neg);
            }
        }
        inv_sm<12>(bq, P.twi[0], P.headi[0], mb, 12, v, v);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 eh = e0 + 4 * h;
            const u64 *a = v + 4 * h;
            i64 sv[4];
            if (i != i_lo) lds_u64x4(reinterpret_cast<const u64 *>(sacc), eh, reinterpret_cast<u64 *>(sv));
            else {
#pragma unroll
                for (int t = 0; t < 4; t++) sv[t] = 0;
            }
            // racc rows are private to the thread: plain 16-byte pieces, no swizzle
            if constexpr (R64) {
                u64 *rr = reinterpret_cast<u64 *>(racc) + eh;
                ulonglong2 r0, r1;
                r0.x = r0.y = r1.x = r1.y = 0;
                if (i != i_lo) { r0 = reinterpret_cast<const ulonglong2 *>(rr)[0]; r1 = reinterpret_cast<const ulonglong2 *>(rr)[1]; }
                r0.x += (u64)(sc.plain32 ? round_term32_signed(a[0], c) : round_term_signed(a[0], c));
                r0.y += (u64)(sc.plain32 ? round_term32_signed(a[1], c) : round_term_signed(a[1], c));
                r1.x += (u64)(sc.plain32 ? round_term32_signed(a[2], c) : round_term_signed(a[2], c));
                r1.y += (u64)(sc.plain32 ? round_term32_signed(a[3], c) : round_term_signed(a[3], c));
                reinterpret_cast<ulonglong2 *>(rr)[0] = r0; reinterpret_cast<ulonglong2 *>(rr)[1] = r1;
            } else {
                Vec16 r;
                if (i != i_lo) r = *reinterpret_cast<const Vec16 *>(racc + eh);
                else {
#pragma unroll
                    for (int t = 0; t < 4; t++) r.w[t] = 0;
                }
#pragma unroll
                for (int t = 0; t < 4; t++)
                    r.w[t] += (u32)(sc.plain32 ? round_term32_signed(a[t], c) : round_term_signed(a[t], c));
                *reinterpret_cast<Vec16 *>(racc + eh) = r;
            }
#pragma unroll
            for (int t = 0; t < 4; t++) sv[t] += center_i64(a[t], c.q, c.half_q);
            sts_u64x4(reinterpret_cast<u64 *>(sacc), eh, reinterpret_cast<const u64 *>(sv));
        }
    }
    // small primes: point-wise products summed over the limb's products, then one inverse transform per prime.  Two
    // primes per round share the 32 KB the base-q image occupied (a third image would cost 16 KB of the L1 that
    // holds the inverse twiddle table of the product loop above).
    u32 bv[kMaxSmall][8];
    tensor01_small_round<0>(P, ext_s, bs, pair, d, k, i_lo, i_hi, comp, e0, bv);
    if (K > 2) tensor01_small_round<2>(P, ext_s, bs, pair, d, k, i_lo, i_hi, comp, e0, bv);
    u64 *o = r01s + ((pair * NL + limb) * 2 + comp) * (size_t)n;
    u64 res[8];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        i64 sv[4], rs[4];
        if constexpr (R64) {
            const u64 *rr = reinterpret_cast<const u64 *>(racc) + e0 + 4 * h;
            const ulonglong2 r0 = reinterpret_cast<const ulonglong2 *>(rr)[0], r1 = reinterpret_cast<const ulonglong2 *>(rr)[1];
            rs[0] = (i64)r0.x; rs[1] = (i64)r0.y; rs[2] = (i64)r1.x; rs[3] = (i64)r1.y;
        } else {
            const Vec16 r = *reinterpret_cast<const Vec16 *>(racc + e0 + 4 * h);
#pragma unroll
            for (int t = 0; t < 4; t++) rs[t] = (i64)(int32_t)r.w[t];
        }
        lds_u64x4(reinterpret_cast<const u64 *>(sacc), e0 + 4 * h, reinterpret_cast<u64 *>(sv));
#pragma unroll
        for (int t = 0; t < 4; t++) {
            u32 b[kMaxSmall];
#pragma unroll
            for (u32 pi = 0; pi < (u32)kMaxSmall; pi++) b[pi] = pi < K ? bv[pi][4 * h + t] : 0u;
            res[4 * h + t] = hps_scale32_sum(rs[t], sv[t], b, c, sc);
        }
    }
    // r01s rows are stored in the transforms' strided layout: word 8t + s holds element t + 512 s, so the relinearisation
    // kernels feed their forward transform's first pass from one 64-byte piece per thread
    stg_u64x4(o + e0, res); stg_u64x4(o + e0 + 4, res + 4);
  }
}

// ---------------------------------------------------------------------------------
// gadget_decompose (bfv/keyswitch.rs:11-52) on coefficient-domain polynomials [count][n]:
//   digits_kernel     signed digits (DigT) in the layout the relin kernels read, [count][G][n]
//   gadget_kernel     the reference's return value: each digit reduced mod q, [count][G][n] u64
// ---------------------------------------------------------------------------------
template <typename OutT, bool MOD_Q>
__global__ void gadget_digits_kernel(const __grid_constant__ DeviceParams P, const u64 *__restrict__ coeffs,
                                     OutT *__restrict__ out, size_t count) {
    const u32 n = P.n, G = P.gadget_digits;
    const u64 q = P.mod[0].m, half_q = q >> 1;
    const size_t total = count * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = idx / n;
        const u32 e = (u32)(idx - poly * n);
        i64 rem = center_i64(coeffs[idx], q, half_q);
        for (u32 g = 0; g < G; g++) {
            const i64 dg = P.gadget_log2 ? gadget_digit_pow2(rem, P.gadget_log2) : gadget_digit_general(rem, (i64)P.gadget_base);
            if constexpr (MOD_Q) out[(poly * G + g) * n + e] = signed_to_mod(dg, q);
            else out[(poly * G + g) * n + e] = (OutT)dg;
        }
    }
}

// ---------------------------------------------------------------------------------
// K6+K7: relinearise + per-k accumulation.  One CTA per (pair, output limb k):
//   c0 = NTT(sum r0) + sum_g NTT(sum digits_g) * rlk0_g     (likewise c1)
// All sums are exact mod q, so any association is bit-equal to the reference's
// per-product relinearize followed by bfv_add (dbfv/eval.rs:125-132).
// ---------------------------------------------------------------------------------
template <int LOGN, typename DigT>
__global__ void __launch_bounds__(LOGN == 12 ? kThreads12 : 256, LOGN == 12 ? 2 : 1)
relin_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
             const u64 *__restrict__ r01, const DigT *__restrict__ digits,
             const u64 *__restrict__ rlk_mont, u64 *__restrict__ out, u64 *__restrict__ excess, u32 r01_ntt) {
    EXB_DYN_SMEM(smem);
    const u32 n = P.n, d = M.d, NP = M.num_products, NL = M.num_limbs, G = P.gadget_digits;
    const Modulus &mq = P.mod[0];
    const u64 q = mq.m;
    u64 *work = smem, *acc0 = smem + n, *acc1 = smem + 2 * (size_t)n;
    const u32 limb = blockIdx.x % NL;
    const size_t pair = blockIdx.x / NL;
    const u32 k = M.limb_k[limb];
    const u32 i_lo = k >= d ? k - d + 1 : 0, i_hi = k < d ? k : d - 1;

    for (u32 comp = 0; comp < 2; comp++) {
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            u64 s = 0;
            for (u32 i = i_lo; i <= i_hi; i++) {
                const size_t pr = (size_t)M.prod_of[i][k - i];
                s = mod_add(s, r01[((pair * NP + pr) * 2 + comp) * n + e], q);
            }
            work[Lay<LOGN>::at(e)] = s;
        }
        if (r01_ntt) __syncthreads();      // r0 / r1 already in the NTT domain (standalone relinearize)
        else fwd_sm<LOGN>(work, P.twf[0], P.headf[0], mq, P.logn);
        u64 *acc = comp ? acc1 : acc0;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) acc[e] = work[Lay<LOGN>::at(e)];
    }
    for (u32 g = 0; g < G; g++) {
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            i64 s = 0;
            for (u32 i = i_lo; i <= i_hi; i++) {
                const size_t pr = (size_t)M.prod_of[i][k - i];
                s += (i64)digits[((pair * NP + pr) * G + g) * n + e];
            }
            u64 v;
            const u64 mag = s < 0 ? (u64)(-s) : (u64)s;
            if (mag < q) v = (s < 0 && mag) ? q - mag : mag;
            else { const u64 r = mag % q; v = (s < 0 && r) ? q - r : r; }
            work[Lay<LOGN>::at(e)] = v;
        }
        fwd_sm<LOGN>(work, P.twf[0], P.headf[0], mq, P.logn);
        const u64 *k0 = rlk_mont + ((size_t)g * 2) * n, *k1 = k0 + n;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            const u64 x = work[Lay<LOGN>::at(e)];
            acc0[e] = mod_add(acc0[e], csub(mont_mul_lazy(x, k0[e], q, mq.minv_neg), q), q);
            acc1[e] = mod_add(acc1[e], csub(mont_mul_lazy(x, k1[e], q, mq.minv_neg), q), q);
        }
    }
    u64 *dst = k < d ? out + ((pair * d + k) * 2) * (size_t)n
                     : excess + ((pair * (NL - M.num_low) + (limb - M.num_low)) * 2) * (size_t)n;
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
        dst[e] = acc0[e];
        dst[n + e] = acc1[e];
    }
    if (k < d && M.num_peers) {            // k-sharded: the finished limb goes straight into every peer's output
        const size_t off = ((pair * d + k) * 2) * (size_t)n;
        for (u32 p = 0; p < M.num_peers; p++) {
            u64 *pd = M.peer_out[p] + off;
            for (u32 e = threadIdx.x; e < n; e += blockDim.x) { pd[e] = acc0[e]; pd[n + e] = acc1[e]; }
        }
        peer_fence();
    }
}

// n = 4096 specialisation of relin_kernel: 8 consecutive coefficients per thread, 128-bit accesses.
template <typename DigT>
__global__ void __launch_bounds__(kThreads12, 2)
relin12_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
               const u64 *__restrict__ r01, const DigT *__restrict__ digits,
               const u64 *__restrict__ rlk_mont, u64 *__restrict__ out, u64 *__restrict__ excess, u32 r01_summed) {
    EXB_DYN_SMEM(smem);
    constexpr u32 n = 4096;
    const u32 d = M.d, NP = M.num_products, NL = M.num_limbs, G = P.gadget_digits;
    const Modulus &mq = P.mod[0];
    const u64 q = mq.m;
    u64 *work = smem, *acc0 = smem + n, *acc1 = smem + 2 * (size_t)n;
    const u32 limb = blockIdx.x % NL;
    const size_t pair = blockIdx.x / NL;
    const u32 k = M.limb_k[limb];
    const u32 i_lo = k >= d ? k - d + 1 : 0, i_hi = k < d ? k : d - 1;
    const u32 e0 = 8 * threadIdx.x;

    for (u32 comp = 0; comp < 2; comp++) {
        u64 sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (r01_summed & 1u) {             // tensor01_kernel already summed the limb's products
            const u64 *src = r01 + ((pair * NL + limb) * 2 + comp) * n + e0;
            ldg_u64x4(src, sum); ldg_u64x4(src + 4, sum + 4);
        } else {
            for (u32 i = i_lo; i <= i_hi; i++) {
                const size_t pr = (size_t)M.prod_of[i][k - i];
                const u64 *src = r01 + ((pair * NP + pr) * 2 + comp) * n + e0;
                u64 v[8];
                ldg_u64x4(src, v); ldg_u64x4(src + 4, v + 4);
#pragma unroll
                for (int j = 0; j < 8; j++) sum[j] = mod_add(sum[j], v[j], q);
            }
        }
        u64 *acc = comp ? acc1 : acc0;
        if (!(r01_summed & 2u)) {          // bit 1: r0 / r1 are already in the NTT domain (standalone relinearize)
            if (r01_summed & 1u) {         // tensor01_kernel's rows: already in the first pass's register layout
                fwd_sm<12>(work, P.twf[0], P.headf[0], mq, 12, sum, sum);
            } else {
                sts_u64x4(work, e0, sum); sts_u64x4(work, e0 + 4, sum + 4);
                fwd_sm<12>(work, P.twf[0], P.headf[0], mq, 12, sum);      // outputs stay in registers
            }
        }
        sts_u64x4(acc, e0, sum); sts_u64x4(acc, e0 + 4, sum + 4);
    }
    for (u32 g = 0; g < G; g++) {
        i64 ds[8];
        digit_sums8<DigT>(digits, M, pair, NP, G, g, n, e0, k, i_lo, i_hi, ds);
        u64 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = signed_sum_to_mod(ds[j], q, mq.lazy != 0);
        sts_u64x4(work, e0, v); sts_u64x4(work, e0 + 4, v + 4);
        fwd_sm<12, false>(work, P.twf[0], P.headf[0], mq, 12, v);    // lazy outputs, kept in registers: they only feed the REDCs below
        const u64 *k0 = rlk_mont + ((size_t)g * 2) * n + e0, *k1 = k0 + n;
        // lazy accumulation for 2^36 <= q < 2^60 (mq.lazy >= 1): the sums stay below 4q + 2^32 (a REDC adds < 2q, then
        // a 3-instruction high-word conditional subtract of 4q) and are made canonical once, when the limb is written out
        const bool lazy_acc = mq.lazy != 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u64 kk[4], a[4];
            const u64 *x = v + 4 * h;
            ldg_u64x4(k0 + 4 * h, kk);
            lds_u64x4(acc0, e0 + 4 * h, a);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u64 t = mont_mul_lazy(x[j], kk[j], q, mq.minv_neg);
                a[j] = lazy_acc ? csub_hi(a[j] + t, mq.four_m, mq.hi_four_m) : mod_add(a[j], csub(t, q), q);
            }
            sts_u64x4(acc0, e0 + 4 * h, a);
            ldg_u64x4(k1 + 4 * h, kk);
            lds_u64x4(acc1, e0 + 4 * h, a);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u64 t = mont_mul_lazy(x[j], kk[j], q, mq.minv_neg);
                a[j] = lazy_acc ? csub_hi(a[j] + t, mq.four_m, mq.hi_four_m) : mod_add(a[j], csub(t, q), q);
            }
            sts_u64x4(acc1, e0 + 4 * h, a);
        }
    }
    u64 *dst = k < d ? out + ((pair * d + k) * 2) * (size_t)n
                     : excess + ((pair * (NL - M.num_low) + (limb - M.num_low)) * 2) * (size_t)n;
    const u32 np = k < d ? M.num_peers : 0u;   // k-sharded: the finished limb also goes into every peer's output
    const size_t off = ((pair * d + k) * 2) * (size_t)n;
    u64 v[8];
    lds_u64x4(acc0, e0, v); lds_u64x4(acc0, e0 + 4, v + 4);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = mq.lazy != 0 ? reduce4(csub(v[j], mq.four_m), q, mq.two_m) : v[j];   // < 4q + 2^32 -> [0, q)
    stg_u64x4(dst + e0, v); stg_u64x4(dst + e0 + 4, v + 4);
    for (u32 p = 0; p < np; p++) { u64 *pd = M.peer_out[p] + off + e0; stg_u64x4(pd, v); stg_u64x4(pd + 4, v + 4); }
    lds_u64x4(acc1, e0, v); lds_u64x4(acc1, e0 + 4, v + 4);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = mq.lazy != 0 ? reduce4(csub(v[j], mq.four_m), q, mq.two_m) : v[j];
    stg_u64x4(dst + n + e0, v); stg_u64x4(dst + n + e0 + 4, v + 4);
    for (u32 p = 0; p < np; p++) { u64 *pd = M.peer_out[p] + off + n + e0; stg_u64x4(pd, v); stg_u64x4(pd + 4, v + 4); }
    if (np) peer_fence();
}

// ---------------------------------------------------------------------------------
// K6+K7 for small batches: relin12_kernel runs (2 + G) transforms back to back in one CTA per output limb,
// which leaves most SMs idle when pairs * limbs < #SMs (a single dbfv_mul has 8 such CTAs).  The wide form
// gives every transform its own CTA -- role g < G: digit plane g (sum over the limb's products, NTT, times
// rlk0_g and rlk1_g); role G: the r0 / r1 sums -- writing 2 polynomials each, and relin_reduce_kernel adds
// the G + 1 partial pairs mod q.  Same sums in another order: exact.
//   partial : [pair][limb][G + 1][2][n]
// ---------------------------------------------------------------------------------
template <typename DigT>
__global__ void __launch_bounds__(kThreads12, 2)
relin12_wide_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
                    const u64 *__restrict__ r01, const DigT *__restrict__ digits,
                    const u64 *__restrict__ rlk_mont, u64 *__restrict__ partial, u32 r01_summed) {
    EXB_DYN_SMEM(smem);
    constexpr u32 n = 4096;
    const u32 d = M.d, NP = M.num_products, NL = M.num_limbs, G = P.gadget_digits;
    const Modulus &mq = P.mod[0];
    const u64 q = mq.m;
    u64 *work = smem;
    const u32 role = blockIdx.x % (G + 1);
    const u32 limb = (blockIdx.x / (G + 1)) % NL;
    const size_t pair = blockIdx.x / ((G + 1) * NL);
    const u32 k = M.limb_k[limb];
    const u32 i_lo = k >= d ? k - d + 1 : 0, i_hi = k < d ? k : d - 1;
    const u32 e0 = 8 * threadIdx.x;
    u64 *dst = partial + (((pair * NL + limb) * (G + 1) + role) * 2) * (size_t)n;
    if (role == G) {
        for (u32 comp = 0; comp < 2; comp++) {
            u64 sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (r01_summed & 1u) {
                const u64 *src = r01 + ((pair * NL + limb) * 2 + comp) * n + e0;
                ldg_u64x4(src, sum); ldg_u64x4(src + 4, sum + 4);
            } else {
                for (u32 i = i_lo; i <= i_hi; i++) {
                    const size_t pr = (size_t)M.prod_of[i][k - i];
                    const u64 *src = r01 + ((pair * NP + pr) * 2 + comp) * n + e0;
                    u64 v[8];
                    ldg_u64x4(src, v); ldg_u64x4(src + 4, v + 4);
#pragma unroll
                    for (int j = 0; j < 8; j++) sum[j] = mod_add(sum[j], v[j], q);
                }
            }
            if (!(r01_summed & 2u)) {
                if (r01_summed & 1u) {     // tensor01_kernel's rows: already in the first pass's register layout
                    fwd_sm<12>(work, P.twf[0], P.headf[0], mq, 12, sum, sum);
                } else {
                    sts_u64x4(work, e0, sum); sts_u64x4(work, e0 + 4, sum + 4);
                    fwd_sm<12>(work, P.twf[0], P.headf[0], mq, 12, sum);
                }
            }
            stg_u64x4(dst + comp * n + e0, sum); stg_u64x4(dst + comp * n + e0 + 4, sum + 4);
            __syncthreads();
        }
        return;
    }
    const u32 g = role;
    i64 ds[8];
    digit_sums8<DigT>(digits, M, pair, NP, G, g, n, e0, k, i_lo, i_hi, ds);
    u64 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = signed_sum_to_mod(ds[j], q, mq.lazy != 0);
    sts_u64x4(work, e0, v); sts_u64x4(work, e0 + 4, v + 4);
    fwd_sm<12, false>(work, P.twf[0], P.headf[0], mq, 12);           // lazy outputs: they only feed the REDCs below
    const u64 *k0 = rlk_mont + ((size_t)g * 2) * n + e0, *k1 = k0 + n;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        u64 x[4], kk[4], a[4];
        lds_u64x4(work, e0 + 4 * h, x);
        ldg_u64x4(k0 + 4 * h, kk);
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = csub(mont_mul_lazy(x[j], kk[j], q, mq.minv_neg), q);
        stg_u64x4(dst + e0 + 4 * h, a);
        ldg_u64x4(k1 + 4 * h, kk);
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = csub(mont_mul_lazy(x[j], kk[j], q, mq.minv_neg), q);
        stg_u64x4(dst + n + e0 + 4 * h, a);
    }
}

// out (or excess) limb = sum of the G + 1 partial pairs of relin12_wide_kernel, mod q.
__global__ void relin_reduce_kernel(const __grid_constant__ DeviceParams P, const __grid_constant__ MulPlan M,
                                    const u64 *__restrict__ partial, u64 *__restrict__ out, u64 *__restrict__ excess,
                                    size_t pairs) {
    const u32 n = P.n, d = M.d, NL = M.num_limbs, R = P.gadget_digits + 1;
    const u64 q = P.mod[0].m;
    const size_t total = pairs * NL * 2 * (size_t)n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 e = (u32)(idx % (2 * (size_t)n));
        const size_t pl = idx / (2 * (size_t)n);
        const u32 limb = (u32)(pl % NL);
        const size_t pair = pl / NL;
        const u64 *src = partial + (pl * R * 2) * (size_t)n + e;
        u64 acc = 0;
        for (u32 r = 0; r < R; r++) acc = mod_add(acc, src[(size_t)r * 2 * n], q);
        const u32 k = M.limb_k[limb];
        u64 *dst = k < d ? out + ((pair * d + k) * 2) * (size_t)n
                         : excess + ((pair * (NL - M.num_low) + (limb - M.num_low)) * 2) * (size_t)n;
        dst[e] = acc;
        if (k < d)
            for (u32 p = 0; p < M.num_peers; p++) M.peer_out[p][((pair * d + k) * 2) * (size_t)n + e] = acc;
    }
    if (M.num_peers) peer_fence();
}

// ---------------------------------------------------------------------------------
// Galois automorphism + key switch (bfv/eval.rs:512-561).  One CTA per ciphertext:
//   c0' = NTT(sigma_k(INTT c0)) + sum_g NTT(digit_g(sigma_k(INTT c1))) * gk0_g
//   c1' =                         sum_g NTT(digit_g(sigma_k(INTT c1))) * gk1_g
// sigma_k (bfv/keygen.rs:218-239) is a signed permutation for odd k, applied as the scatter
// out of the inverse transform.  The running `remaining` of gadget_decompose
// (bfv/keyswitch.rs:24-44) lives in shared memory as i64; c1' accumulates in `out` itself
// (each thread re-visits only its own words).
// ---------------------------------------------------------------------------------
template <int LOGN>
__global__ void __launch_bounds__(LOGN == 12 ? kThreads12 : 256, LOGN == 12 ? 2 : 1)
galois_kernel(const __grid_constant__ DeviceParams P, const u64 *__restrict__ ct,
              const u64 *__restrict__ gk_mont, u32 k, u64 *__restrict__ out) {
    EXB_DYN_SMEM(smem);
    const u32 n = P.n, G = P.gadget_digits, mask2n = 2 * n - 1;
    const Modulus &mq = P.mod[0];
    const u64 q = mq.m, half_q = q >> 1;
    u64 *work = smem, *acc0 = smem + n;
    i64 *remaining = reinterpret_cast<i64 *>(smem + 2 * (size_t)n);
    const u64 *src = ct + (size_t)blockIdx.x * 2 * n;
    u64 *dst = out + (size_t)blockIdx.x * 2 * n;

    for (u32 e = threadIdx.x; e < n; e += blockDim.x) work[Lay<LOGN>::at(e)] = ld_stream(src + e);
    inv_sm<LOGN>(work, P.twi[0], P.headi[0], mq, P.logn);
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
        const u64 c = work[Lay<LOGN>::at(e)];
        const u32 j = (e * k) & mask2n;
        acc0[Lay<LOGN>::at(j & (n - 1))] = j < n ? c : mod_neg(c, q);
    }
    fwd_sm<LOGN>(acc0, P.twf[0], P.headf[0], mq, P.logn);

    for (u32 e = threadIdx.x; e < n; e += blockDim.x) work[Lay<LOGN>::at(e)] = ld_stream(src + n + e);
    inv_sm<LOGN>(work, P.twi[0], P.headi[0], mq, P.logn);
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
        const u64 c = work[Lay<LOGN>::at(e)];
        const u32 j = (e * k) & mask2n;
        remaining[j & (n - 1)] = center_i64(j < n ? c : mod_neg(c, q), q, half_q);
    }
    __syncthreads();
    for (u32 g = 0; g < G; g++) {
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            i64 r = remaining[e];
            const i64 dg = P.gadget_log2 ? gadget_digit_pow2(r, P.gadget_log2)
                                         : gadget_digit_general(r, (i64)P.gadget_base);
            remaining[e] = r;
            work[Lay<LOGN>::at(e)] = signed_to_mod(dg, q);
        }
        fwd_sm<LOGN, false>(work, P.twf[0], P.headf[0], mq, P.logn);      // lazy outputs feed the REDCs below
        const u64 *k0 = gk_mont + ((size_t)g * 2) * n, *k1 = k0 + n;
        for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
            const u64 x = work[Lay<LOGN>::at(e)];
            const u32 a = Lay<LOGN>::at(e);
            acc0[a] = mod_add(acc0[a], csub(mont_mul_lazy(x, k0[e], q, mq.minv_neg), q), q);
            const u64 y = csub(mont_mul_lazy(x, k1[e], q, mq.minv_neg), q);
            dst[n + e] = g ? mod_add(dst[n + e], y, q) : y;
        }
    }
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) dst[e] = acc0[Lay<LOGN>::at(e)];
}

// ---------------------------------------------------------------------------------
// decrypt (bfv/encrypt.rs:111-178, single ciphertext prime).  One CTA per ciphertext:
//   phase = c0 + c1 s + c2 s^2 + ...  (NTT domain),  x = INTT(phase),
//   m = floor((p x + floor(q/2)) / q) mod p     with x in [0, q)  (the reference does not centre x).
// p x = k q + r by Shoup (k = mulhi(x, p_s), r in [0, 2q)), so the floor is k + [r+h >= q] + [r+h >= 2q].
// ---------------------------------------------------------------------------------
template <int LOGN>
__global__ void __launch_bounds__(LOGN == 12 ? kThreads12 : 256, LOGN == 12 ? 2 : 1)
decrypt_kernel(const __grid_constant__ DeviceParams P, const u64 *__restrict__ ct, u32 ncomp,
               const u64 *__restrict__ sk_ntt, u64 *__restrict__ out) {
    EXB_DYN_SMEM(smem);
    const u32 n = P.n;
    const Modulus &mq = P.mod[0];
    const u64 q = mq.m, half_q = q >> 1, p = P.sc.plain, p_s = P.sc.plain_s;
    const u64 *src = ct + (size_t)blockIdx.x * ncomp * n;
    u64 *dst = out + (size_t)blockIdx.x * n;
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
        const u64 s_m = csub(mont_mul_lazy(sk_ntt[e], mq.r2_mod, q, mq.minv_neg), q);      // s * 2^64
        u64 phase = ld_stream(src + e), spow = s_m;
        for (u32 i = 1; i < ncomp; i++) {
            phase = mod_add(phase, csub(mont_mul_lazy(ld_stream(src + (size_t)i * n + e), spow, q, mq.minv_neg), q), q);
            if (i + 1 < ncomp) spow = csub(mont_mul_lazy(spow, s_m, q, mq.minv_neg), q);
        }
        smem[Lay<LOGN>::at(e)] = phase;
    }
    inv_sm<LOGN>(smem, P.twi[0], P.headi[0], mq, P.logn);
    for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
        const u64 x = smem[Lay<LOGN>::at(e)];
        const u64 k = mulhi64(x, p_s);
        const u64 r = x * p - k * q;
        const u64 rh = r + half_q;
        const u64 m = k + (rh >= q ? 1u : 0u) + (rh >= 2 * q ? 1u : 0u);                    // <= p
        dst[e] = m >= p ? m - p : m;
    }
}

// out_limb += (+/-) s * excess_limb over 2n words per pair (dbfv/reduction.rs:34-52, :65-93).
__global__ void reduce_mac_kernel(Modulus mod, u64 *__restrict__ out_limb, const u64 *__restrict__ excess_limb,
                                  u64 s_mont, int negative, size_t out_stride, size_t excess_stride,
                                  u32 words, size_t pairs) {
    const u64 m = mod.m;
    const size_t total = pairs * words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t pair = i / words, e = i - pair * words;
        u64 v = csub(mont_mul_lazy(excess_limb[pair * excess_stride + e], s_mont, m, mod.minv_neg), m);
        if (negative) v = mod_neg(v, m);
        u64 *o = out_limb + pair * out_stride + e;
        *o = mod_add(*o, v, m);
    }
}

// SM count of the current device (grid sizing of the persistent / small-batch paths).
#ifndef EXB_HOST_EMUL
int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}
#else
int num_sms() { return 148; }
#endif

static u32 most_products_per_limb(const MulPlan &M) {
    u32 worst = 0;
    for (u32 l = 0; l < M.num_limbs; l++) {
        const u32 k = M.limb_k[l];
        const u32 cnt = (k < M.d ? k : M.d - 1) - (k >= M.d ? k - M.d + 1 : 0) + 1;
        worst = cnt > worst ? cnt : worst;
    }
    return worst;
}
bool tensor01_needs_r64(const DeviceParams &P, const MulPlan &M) { return most_products_per_limb(M) > P.sb.max_terms_r32; }

// Components 0/1 can be produced per output limb (tensor01_kernel) when the internal basis is on and
// every limb sums at most SmallBasis::max_terms products.
bool tensor_sums_per_limb(const DeviceParams &P, const MulPlan &M, size_t pairs) {
    if (!P.sb.enabled || P.logn != 12) return false;
    if (P.tensor_per_product) return false;                 // exb_context_set_option: the per-product kernel only
    const u32 worst = most_products_per_limb(M);
    if (worst > P.sb.max_terms) return false;               // launch_tensor picks the i32 / i64 rounding-sum variant
    // It pays when limbs sum several products (fewer small-prime inverse transforms) and the per-limb CTAs
    // (1..d products each) still fill the GPU; small batches keep the finer-grained per-product kernel.
    // measured crossover when the kernel has the GPU to itself: 24-28 pairs at d = 8 (~200 CTAs); chunks of the
    // host pipeline overlap with their neighbours' kernels, which fill the SMs a short grid leaves idle
    return M.num_products > M.num_limbs && pairs * M.num_duos * 2 >= (P.pipelined ? 100u : 200u);
}

// Small batches at n = 4096: one CTA per transform instead of one per output limb (relin12_wide_kernel).
bool relin_goes_wide(const DeviceParams &P, const MulPlan &M, size_t pairs) {
    return !P.relin_narrow && P.logn == 12 && P.gadget_digits > 0 && pairs * M.num_limbs < (size_t)num_sms();
}
size_t relin_wide_scratch_bytes(const DeviceParams &P, const MulPlan &M, size_t pairs) {
    return pairs * M.num_limbs * (size_t)(P.gadget_digits + 1) * 2 * P.n * sizeof(u64);
}

#ifndef EXB_HOST_EMUL
// ---------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------

static inline u32 block_threads(const DeviceParams &P) {
    if (P.logn == 12) return kThreads12;
    u32 t = P.n / 2;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}


template <bool FWD, int LAZY, int NB>
static void launch_ntt12_nb(const Modulus &mod, const Tw *tw, const TwHead &head, const u64 *in, u64 *out, size_t count,
                            cudaStream_t s) {
    const size_t sm = 2 * 4096 * 8;
    const size_t slots = (size_t)num_sms() * 2;               // persistent: 2 CTAs per SM
    const unsigned grid = (unsigned)(count < slots ? count : slots);
#ifdef EXB_LAB
    static const int dbg = getenv("EXB_NTT_DBG") ? atoi(getenv("EXB_NTT_DBG")) : 0;   // lab build: see the kernel
    if (dbg == 1) { ntt12_persist_kernel<FWD, LAZY, NB, 1><<<grid, 4096 >> NB, sm, s>>>(in, out, tw, head, mod, (u32)count); return; }
    if (dbg == 2) { ntt12_persist_kernel<FWD, LAZY, NB, 2><<<grid, 4096 >> NB, sm, s>>>(in, out, tw, head, mod, (u32)count); return; }
    if (dbg == 3) { ntt12_persist_kernel<FWD, LAZY, NB, 3><<<grid, 4096 >> NB, sm, s>>>(in, out, tw, head, mod, (u32)count); return; }
#endif
    ntt12_persist_kernel<FWD, LAZY, NB><<<grid, 4096 >> NB, sm, s>>>(in, out, tw, head, mod, (u32)count);
}
template <bool FWD, int LAZY>
static void launch_ntt12_l(const Modulus &mod, const Tw *tw, const TwHead &head, const u64 *in, u64 *out, size_t count,
                           cudaStream_t s) {
#ifdef EXB_LAB
    static const int nb = getenv("EXB_NTT_NB") ? atoi(getenv("EXB_NTT_NB")) : kNB;   // lab build: 16 values per thread
    if (nb == 4) { launch_ntt12_nb<FWD, LAZY, 4>(mod, tw, head, in, out, count, s); return; }
#endif
    launch_ntt12_nb<FWD, LAZY, 3>(mod, tw, head, in, out, count, s);
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// [count] polynomials of 4096 u64 as a 2-D tensor of 128-byte rows; one box = one polynomial (256 rows), SWIZZLE_128B.
static bool poly_tensor_map(CUtensorMap *map, const u64 *base, size_t count) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || count * 256 > 0xffffffffull) return false;
    const cuuint64_t dims[2] = {16, (cuuint64_t)count * 256};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {16, 256}, estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<u64 *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool FWD, int LAZY>
static bool launch_ntt12_tma(const Modulus &mod, const Tw *tw, const TwHead &head, const u64 *in, u64 *out, size_t count,
                             cudaStream_t s) {
    CUtensorMap tin, tout;
    if (!poly_tensor_map(&tin, in, count) || !poly_tensor_map(&tout, out, count)) return false;
    const size_t slots = (size_t)num_sms() * 2;               // persistent: 2 CTAs per SM
    const unsigned grid = (unsigned)(count < slots ? count : slots);
    // measured (tools/lab_ntt.py): the staged table wins everywhere except the csub-free inverse (LAZY 2), whose
    // middle passes are short enough for the L1-resident table to keep up
    constexpr bool kStaged = FWD || LAZY != 2;
#ifdef EXB_LAB
    static const int gtw = getenv("EXB_TMA_GLOBAL_TW") ? 1 : 0;
    if (gtw) { ntt12_tma_kernel<FWD, LAZY, false><<<grid, 512, kTmaSmemBytes, s>>>(tin, tout, tw, head, mod, (u32)count); return true; }
#endif
    ntt12_tma_kernel<FWD, LAZY, kStaged><<<grid, 512, kTmaSmemBytes, s>>>(tin, tout, tw, head, mod, (u32)count);
    return true;
}

std::atomic<int> g_ntt_path{0};                               // 0: TMA kernel, 1: cp.async kernel (exb_set_ntt_path)

template <bool FWD>
static void launch_ntt12(const Modulus &mod, const Tw *tw, const TwHead &head, const u64 *in, u64 *out, size_t count,
                         cudaStream_t s) {
    if (g_ntt_path.load(std::memory_order_relaxed) == 0) {
        bool ok;
        if (mod.lazy == 2) ok = launch_ntt12_tma<FWD, 2>(mod, tw, head, in, out, count, s);
        else if (mod.lazy == 1) ok = launch_ntt12_tma<FWD, 1>(mod, tw, head, in, out, count, s);
        else ok = launch_ntt12_tma<FWD, 0>(mod, tw, head, in, out, count, s);
        if (ok) return;
    }
    if (mod.lazy == 2) launch_ntt12_l<FWD, 2>(mod, tw, head, in, out, count, s);
    else if (mod.lazy == 1) launch_ntt12_l<FWD, 1>(mod, tw, head, in, out, count, s);
    else launch_ntt12_l<FWD, 0>(mod, tw, head, in, out, count, s);
}

void launch_ntt_plan(const Modulus &mod, const Tw *tw, const TwHead &head, u32 logn, bool forward, const u64 *in,
                     u64 *out, size_t count, cudaStream_t s) {
    if (count == 0) return;
    const u32 n = 1u << logn;
    const size_t sm = (size_t)n * 8;
    u32 threads = n / 2;
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    if (logn == 12) {
        if (forward) launch_ntt12<true>(mod, tw, head, in, out, count, s);
        else launch_ntt12<false>(mod, tw, head, in, out, count, s);
    } else if (forward) {
        ntt_fwd_kernel<0><<<(unsigned)count, threads, sm, s>>>(in, out, tw, head, mod, logn);
    } else {
        ntt_inv_kernel<0><<<(unsigned)count, threads, sm, s>>>(in, out, tw, head, mod, logn);
    }
    g_launch_count++;
}

void launch_ntt_fwd(const DeviceParams &P, int base, const u64 *in, u64 *out, size_t count, cudaStream_t s) {
    launch_ntt_plan(P.mod[base], P.twf[base], P.headf[base], P.logn, true, in, out, count, s);
}

void launch_ntt_inv(const DeviceParams &P, int base, const u64 *in, u64 *out, size_t count, cudaStream_t s) {
    launch_ntt_plan(P.mod[base], P.twi[base], P.headi[base], P.logn, false, in, out, count, s);
}

void launch_poly_op(const Modulus &m, PolyOp op, const u64 *a, const u64 *b, u64 scalar, u64 *out,
                    size_t words, cudaStream_t s) {
    if (words == 0) return;
    size_t blocks = (words + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    poly_op_kernel<<<(unsigned)blocks, 256, 0, s>>>(m, (int)op, a, b, scalar, out, words);
    g_launch_count++;
}

// With the internal small basis the lift workspace holds ext_s only.
static inline u32 *ext_small_part(u64 *ext) { return reinterpret_cast<u32 *>(ext); }

void launch_lift(const DeviceParams &P, const MulPlan &M, const u64 *ct1, const u64 *ct2, u64 *ext,
                 size_t pairs, cudaStream_t s) {
    if (pairs == 0) return;
    if (P.sb.enabled && P.logn == 12) {
        const size_t sm32 = 4096 * 8 + (size_t)P.sb.K * 4096 * 4;
        lift32_kernel<<<(unsigned)(pairs * 4 * M.d), kThreads12, sm32, s>>>(P, M.d, M.need_lhs, M.need_rhs, ct1, ct2, ext_small_part(ext));
        g_launch_count++;
        return;
    }
    const size_t sm = (size_t)P.n * 8 * 2;
    const unsigned grid = (unsigned)(pairs * 4 * M.d);
    if (P.logn == 12) {
        lift_kernel<12><<<grid, kThreads12, sm, s>>>(P, M.d, M.need_lhs, M.need_rhs, ct1, ct2, ext);
    } else {
        lift_kernel<0><<<grid, block_threads(P), sm, s>>>(P, M.d, M.need_lhs, M.need_rhs, ct1, ct2, ext);
    }
    g_launch_count++;
}

template <typename DigT>
static void launch_tensor_t(const DeviceParams &P, const MulPlan &M, const u64 *ct1, const u64 *ct2, const u64 *ext,
                            u64 *r01, DigT *digits, size_t pairs, cudaStream_t s, cudaEvent_t mid, bool raw3) {
    const size_t sm = (size_t)P.n * 8 * (1 + P.num_aux);
    const unsigned grid = (unsigned)(pairs * M.num_products * 3);
    if (P.sb.enabled && P.logn == 12) {
        const size_t sm32 = 4096 * 8 + (size_t)P.sb.K * 4096 * 4;
        const u32 *ext_s = ext_small_part(const_cast<u64 *>(ext));
        if (!raw3 && tensor_sums_per_limb(P, M, pairs)) {
            // components 0/1 per output limb (one small-prime inverse transform per limb), component 2 per product
            const bool r64 = tensor01_needs_r64(P, M);
            const size_t sm01 = smem_tensor01(r64);
            if (r64) tensor01_kernel<true><<<(unsigned)(pairs * M.num_duos * 2), kThreads12, sm01, s>>>(P, M, ct1, ct2, ext_s, r01);
            else tensor01_kernel<false><<<(unsigned)(pairs * M.num_duos * 2), kThreads12, sm01, s>>>(P, M, ct1, ct2, ext_s, r01);
            if (mid) cudaEventRecord(mid, s);
            tensor32_kernel<DigT><<<(unsigned)(pairs * M.num_products), kThreads12, sm32, s>>>(P, M, ct1, ct2, ext_s, r01,
                                                                                             digits, 1u);
            g_launch_count += 2;
            return;
        }
        tensor32_kernel<DigT><<<grid, kThreads12, sm32, s>>>(P, M, ct1, ct2, ext_s, r01, digits, raw3 ? 2u : 0u);
        if (mid) cudaEventRecord(mid, s);
        g_launch_count++;
        return;
    }
    if (P.logn == 12) {
        tensor_kernel<12, DigT><<<grid, kThreads12, sm, s>>>(P, M, ct1, ext, r01, digits, raw3 ? 1u : 0u);
    } else {
        tensor_kernel<0, DigT><<<grid, block_threads(P), sm, s>>>(P, M, ct1, ext, r01, digits, raw3 ? 1u : 0u);
    }
    if (mid) cudaEventRecord(mid, s);
    g_launch_count++;
}

void launch_tensor(const DeviceParams &P, const MulPlan &M, const u64 *ct1, const u64 *ct2, const u64 *ext, u64 *r01,
                   void *digits, int digit_kind, size_t pairs, cudaStream_t s, cudaEvent_t mid, bool raw3) {
    if (pairs == 0) { if (mid) cudaEventRecord(mid, s); return; }
    if (digit_kind == 2) launch_tensor_t<int8_t>(P, M, ct1, ct2, ext, r01, (int8_t *)digits, pairs, s, mid, raw3);
    else if (digit_kind == 1) launch_tensor_t<int32_t>(P, M, ct1, ct2, ext, r01, (int32_t *)digits, pairs, s, mid, raw3);
    else launch_tensor_t<int16_t>(P, M, ct1, ct2, ext, r01, (int16_t *)digits, pairs, s, mid, raw3);
}

void launch_gadget_digits(const DeviceParams &P, const u64 *coeffs, void *out, int out_kind, size_t count, cudaStream_t s) {
    if (count == 0) return;
    size_t blocks = (count * P.n + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    if (out_kind == 3) gadget_digits_kernel<int8_t, false><<<(unsigned)blocks, 256, 0, s>>>(P, coeffs, (int8_t *)out, count);
    else if (out_kind == 0) gadget_digits_kernel<int16_t, false><<<(unsigned)blocks, 256, 0, s>>>(P, coeffs, (int16_t *)out, count);
    else if (out_kind == 1) gadget_digits_kernel<int32_t, false><<<(unsigned)blocks, 256, 0, s>>>(P, coeffs, (int32_t *)out, count);
    else gadget_digits_kernel<u64, true><<<(unsigned)blocks, 256, 0, s>>>(P, coeffs, (u64 *)out, count);
    g_launch_count++;
}

template <typename DigT>
static void launch_relin_t(const DeviceParams &P, const MulPlan &M, const u64 *r01, const DigT *digits,
                           const u64 *rlk_mont, u64 *out, u64 *excess, size_t pairs, cudaStream_t s, u64 *wide_scratch,
                           bool r01_ntt) {
    const size_t sm = (size_t)P.n * 8 * 3;
    const unsigned grid = (unsigned)(pairs * M.num_limbs);
    // flags: bit 0 = r01 is [pair][limb][2][n] (per-limb sums), bit 1 = r0 / r1 already in the NTT domain
    const u32 flags12 = r01_ntt ? 3u : (tensor_sums_per_limb(P, M, pairs) ? 1u : 0u);
    if (wide_scratch && relin_goes_wide(P, M, pairs)) {
        const u32 summed = flags12;
        relin12_wide_kernel<DigT><<<grid * (P.gadget_digits + 1), kThreads12, (size_t)P.n * 8, s>>>(
            P, M, r01, digits, rlk_mont, wide_scratch, summed);
        size_t blocks = (pairs * M.num_limbs * 2 * (size_t)P.n + 255) / 256;
        if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
        relin_reduce_kernel<<<(unsigned)blocks, 256, 0, s>>>(P, M, wide_scratch, out, excess, pairs);
        g_launch_count += 2;
        return;
    }
    if (P.logn == 12) {
        relin12_kernel<DigT><<<grid, kThreads12, sm, s>>>(P, M, r01, digits, rlk_mont, out, excess, flags12);
    } else {
        relin_kernel<0, DigT><<<grid, block_threads(P), sm, s>>>(P, M, r01, digits, rlk_mont, out, excess, r01_ntt ? 1u : 0u);
    }
    g_launch_count++;
}

void launch_relin(const DeviceParams &P, const MulPlan &M, const u64 *r01, const void *digits, int digit_kind,
                  const u64 *rlk_mont, u64 *out, u64 *excess, size_t pairs, cudaStream_t s, u64 *wide_scratch,
                  bool r01_ntt) {
    if (pairs == 0) return;
    if (digit_kind == 2) launch_relin_t<int8_t>(P, M, r01, (const int8_t *)digits, rlk_mont, out, excess, pairs, s, wide_scratch, r01_ntt);
    else if (digit_kind == 1) launch_relin_t<int32_t>(P, M, r01, (const int32_t *)digits, rlk_mont, out, excess, pairs, s, wide_scratch, r01_ntt);
    else launch_relin_t<int16_t>(P, M, r01, (const int16_t *)digits, rlk_mont, out, excess, pairs, s, wide_scratch, r01_ntt);
}

void launch_galois(const DeviceParams &P, const u64 *ct, const u64 *gk_mont, u32 element, u64 *out,
                   size_t count, cudaStream_t s) {
    if (count == 0) return;
    const size_t sm = (size_t)P.n * 8 * 3;
    const u32 k = element & (2 * P.n - 1);
    if (P.logn == 12) {
        galois_kernel<12><<<(unsigned)count, kThreads12, sm, s>>>(P, ct, gk_mont, k, out);
    } else {
        galois_kernel<0><<<(unsigned)count, block_threads(P), sm, s>>>(P, ct, gk_mont, k, out);
    }
    g_launch_count++;
}

void launch_decrypt(const DeviceParams &P, const u64 *ct, u32 ncomp, const u64 *sk_ntt, u64 *out, size_t count,
                    cudaStream_t s) {
    if (count == 0) return;
    const size_t sm = (size_t)P.n * 8;
    if (P.logn == 12) {
        decrypt_kernel<12><<<(unsigned)count, kThreads12, sm, s>>>(P, ct, ncomp, sk_ntt, out);
    } else {
        decrypt_kernel<0><<<(unsigned)count, block_threads(P), sm, s>>>(P, ct, ncomp, sk_ntt, out);
    }
    g_launch_count++;
}

void launch_reduce_mac(const DeviceParams &P, u64 *out_limb, const u64 *excess_limb, u64 abs_scalar_mod_q,
                       bool negative, size_t out_stride, size_t excess_stride, size_t pairs, cudaStream_t s) {
    if (pairs == 0) return;
    const Modulus &m = P.mod[0];
    // scalar in Montgomery form so one REDC gives x * s mod q
    const u64 s_mont = (u64)(((unsigned __int128)abs_scalar_mod_q << 64) % m.m);
    const u32 words = 2 * P.n;
    size_t blocks = (pairs * words + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    reduce_mac_kernel<<<(unsigned)blocks, 256, 0, s>>>(m, out_limb, excess_limb, s_mont, negative ? 1 : 0,
                                                      out_stride, excess_stride, words, pairs);
    g_launch_count++;
}

// Opt every kernel that uses dynamic shared memory into the device's full budget, once per device (called from
// exb_context_create) instead of on every launch.
template <typename K>
static void opt_in(K kernel, int bytes) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }

void launch_prepare(int device) {
    static std::atomic<unsigned long long> done{0};
    if (device >= 0 && device < 64 && ((done.load() >> device) & 1ull)) return;
    int optin = 0;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess || optin <= 0) optin = 227 * 1024;
    opt_in(ntt_fwd_kernel<0>, optin); opt_in(ntt_inv_kernel<0>, optin);
#define EXB_OPT_NTT(F, L) opt_in(ntt12_persist_kernel<F, L, 3>, optin);
    EXB_OPT_NTT(true, 0) EXB_OPT_NTT(true, 1) EXB_OPT_NTT(true, 2) EXB_OPT_NTT(false, 0) EXB_OPT_NTT(false, 1) EXB_OPT_NTT(false, 2)
#undef EXB_OPT_NTT
#ifdef EXB_LAB
#define EXB_OPT_NTT(F, L) opt_in(ntt12_persist_kernel<F, L, 4>, optin); opt_in(ntt12_persist_kernel<F, L, 3, 1>, optin); opt_in(ntt12_persist_kernel<F, L, 4, 1>, optin); \
    opt_in(ntt12_persist_kernel<F, L, 3, 2>, optin); opt_in(ntt12_persist_kernel<F, L, 3, 3>, optin); \
    opt_in(ntt12_persist_kernel<F, L, 4, 2>, optin); opt_in(ntt12_persist_kernel<F, L, 4, 3>, optin);
    EXB_OPT_NTT(true, 0) EXB_OPT_NTT(true, 1) EXB_OPT_NTT(true, 2) EXB_OPT_NTT(false, 0) EXB_OPT_NTT(false, 1) EXB_OPT_NTT(false, 2)
#undef EXB_OPT_NTT
#endif
#define EXB_OPT_NTT(F, L) opt_in(ntt12_tma_kernel<F, L>, optin); opt_in(ntt12_tma_kernel<F, L, false>, optin);
    EXB_OPT_NTT(true, 0) EXB_OPT_NTT(true, 1) EXB_OPT_NTT(true, 2) EXB_OPT_NTT(false, 0) EXB_OPT_NTT(false, 1) EXB_OPT_NTT(false, 2)
#undef EXB_OPT_NTT
    opt_in(lift32_kernel, optin); opt_in(lift_kernel<12>, optin); opt_in(lift_kernel<0>, optin);
    opt_in(tensor01_kernel<false>, optin); opt_in(tensor01_kernel<true>, optin);
    opt_in(tensor32_kernel<int16_t>, optin); opt_in(tensor32_kernel<int32_t>, optin); opt_in(tensor32_kernel<int8_t>, optin);
    opt_in(tensor_kernel<12, int16_t>, optin); opt_in(tensor_kernel<12, int32_t>, optin); opt_in(tensor_kernel<12, int8_t>, optin);
    opt_in(tensor_kernel<0, int16_t>, optin); opt_in(tensor_kernel<0, int32_t>, optin); opt_in(tensor_kernel<0, int8_t>, optin);
    opt_in(relin12_kernel<int16_t>, optin); opt_in(relin12_kernel<int32_t>, optin); opt_in(relin12_kernel<int8_t>, optin);
    opt_in(relin12_wide_kernel<int16_t>, optin); opt_in(relin12_wide_kernel<int32_t>, optin); opt_in(relin12_wide_kernel<int8_t>, optin);
    opt_in(relin_kernel<0, int16_t>, optin); opt_in(relin_kernel<0, int32_t>, optin); opt_in(relin_kernel<0, int8_t>, optin);
    opt_in(galois_kernel<12>, optin); opt_in(galois_kernel<0>, optin);
    opt_in(decrypt_kernel<12>, optin); opt_in(decrypt_kernel<0>, optin);
    if (device >= 0 && device < 64) done.fetch_or(1ull << device);
}

#endif  // EXB_HOST_EMUL

}  // namespace exb
