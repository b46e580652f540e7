// emul.cpp -- TEST INFRASTRUCTURE: replays the product's CUDA kernels on the CPU.
//
// exacto_b200/csrc/kernels.cu is compiled *unchanged* by g++ under a small shim that
// maps CUDA's execution model onto OS threads: one std::thread per CUDA thread of a
// block, pthread barriers for __syncthreads, a heap buffer for dynamic shared
// memory, blocks executed one after another.  This lets the CPU-only test tier check
// the real kernel source (index math, barriers, lazy-reduction bounds, constants
// from host_setup.cpp) bit-for-bit against the oracle before any GPU time is spent.
// It is NOT a fallback: nothing in exacto_b200/ links or loads this library.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <type_traits>
#include <vector>

// ---- CUDA shim ---------------------------------------------------------------------
#define EXB_HOST_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(x) alignas(x)

struct emu_dim3 { unsigned x, y, z; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local void *emu_smem_ptr;
static pthread_barrier_t *emu_barrier;
static inline void __syncthreads() { pthread_barrier_wait(emu_barrier); }
#define EXB_DYN_SMEM(name) exb::u64 *name = (exb::u64 *)emu_smem_ptr

#include "../../exacto_b200/csrc/kernels.cu"
#include "../../exacto_b200/csrc/rns_kernels.cu"

using namespace exb;

// Run `grid` blocks of `block` threads; body() is the kernel call.
template <typename F>
static void emu_launch(unsigned grid, unsigned block, size_t smem_bytes, F body) {
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, block);
    emu_barrier = &bar;
    void *smem = nullptr;
    if (posix_memalign(&smem, 128, smem_bytes ? smem_bytes : 128)) abort();
    std::vector<std::thread> threads;
    threads.reserve(block);
    for (unsigned t = 0; t < block; t++)
        threads.emplace_back([=, &bar]() {
            threadIdx = {t, 0, 0};
            blockDim = {block, 1, 1};
            gridDim = {grid, 1, 1};
            emu_smem_ptr = smem;
            for (unsigned b = 0; b < grid; b++) {
                blockIdx = {b, 0, 0};
                body();
                pthread_barrier_wait(&bar);   // blocks run back to back on the same "SM"
            }
        });
    for (auto &th : threads) th.join();
    free(smem);
    pthread_barrier_destroy(&bar);
}

static unsigned emu_block_threads(const DeviceParams &P) {
    if (P.logn == 12) return kThreads12;
    unsigned t = P.n / 2;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}

struct emu_ctx {
    HostSetup hs;
    std::string err;
};

static thread_local std::string g_emu_err;

extern "C" {

const char *emu_last_error(void) { return g_emu_err.c_str(); }

int emu_create(uint32_t n, const uint64_t *ct_moduli, uint32_t num_ct, const uint64_t *aux, uint32_t num_aux,
               uint64_t plain, uint64_t gadget_base, uint32_t gadget_digits, emu_ctx **out) {
    exb_bfv_params p;
    p.ring_degree = n; p.num_ct_moduli = num_ct; p.ct_moduli = ct_moduli;
    p.num_aux_moduli = num_aux; p.aux_moduli = aux; p.plain_modulus = plain;
    p.gadget_base = gadget_base; p.gadget_digits = gadget_digits;
    emu_ctx *c = new emu_ctx();
    // test infrastructure: the aux-basis choice is a context flag in the product (EXB_CTX_REFERENCE_AUX_BASIS);
    // the emulator takes it from the environment so one fixture covers both
    const char *env = getenv("EXB_AUX_BASIS");
    const uint32_t cflags = (env && strcmp(env, "reference") == 0) ? (uint32_t)EXB_CTX_REFERENCE_AUX_BASIS : 0u;
    int rc = host_setup_build(&p, &c->hs, &g_emu_err, cflags);
    if (rc) { delete c; return rc; }
    for (int b = 0; b < kMaxBases; b++)
        if (c->hs.has_plan[b]) { c->hs.P.twf[b] = c->hs.twf[b].data(); c->hs.P.twi[b] = c->hs.twi[b].data(); }
    for (u32 i = 0; c->hs.P.sb.enabled && i < c->hs.P.sb.K; i++) {
        c->hs.P.sb.twf[i] = c->hs.twf32[i].data(); c->hs.P.sb.twi[i] = c->hs.twi32[i].data();
    }
    if (c->hs.rns_enabled) {
        for (u32 l = 0; l < c->hs.R.L; l++) { c->hs.T.twf_q[l] = c->hs.rns_twf_q[l].data(); c->hs.T.twi_q[l] = c->hs.rns_twi_q[l].data(); }
        for (u32 k = 0; k < c->hs.R.K; k++) { c->hs.T.twf_e[k] = c->hs.rns_twf_e[k].data(); c->hs.T.twi_e[k] = c->hs.rns_twi_e[k].data(); }
    }
    *out = c;
    return 0;
}

void emu_destroy(emu_ctx *c) { delete c; }

int emu_info(const emu_ctx *c, uint64_t *gadget_base, uint32_t *gadget_digits, int *mul_status, uint64_t *psi0) {
    *gadget_base = c->hs.gadget_base; *gadget_digits = c->hs.gadget_digits;
    *mul_status = c->hs.mul_status; *psi0 = c->hs.psi[0];
    g_emu_err = c->hs.mul_error;
    return 0;
}

// Internal small auxiliary basis chosen for this parameter set (0 = reference basis in use).
int emu_small_primes(const emu_ctx *c, uint64_t *primes) {
    if (!c->hs.P.sb.enabled) return 0;
    for (size_t i = 0; i < c->hs.small_primes.size(); i++) primes[i] = c->hs.small_primes[i];
    return (int)c->hs.small_primes.size();
}

int emu_ntt(emu_ctx *c, uint32_t base, int forward, const uint64_t *in, uint64_t *out, size_t count) {
    const DeviceParams &P = c->hs.P;
    if (base >= (uint32_t)kMaxBases || !c->hs.has_plan[base]) { g_emu_err = "no plan"; return EXB_MODULUS_MISMATCH; }
    const unsigned thr = emu_block_threads(P);
    const size_t sm = (size_t)P.n * 8;
    if (P.logn == 12) {
        // persistent fast path: few "CTAs" so every block loops over several polynomials;
        // both register tilings (8 and 16 values per thread) are exercised: odd/even count
        const unsigned grid = count < 3 ? (unsigned)count : 3u;
        const Modulus &m = P.mod[base];
        const u32 cnt = (u32)count;
        const bool nb4 = (count & 1) == 0;
#define EMU_NTT12(FWD, LZ) do { if (nb4) emu_launch(grid, 256, 2 * sm, [&]() { ntt12_persist_kernel<FWD, LZ, 4>(in, out, FWD ? P.twf[base] : P.twi[base], FWD ? P.headf[base] : P.headi[base], m, cnt); }); \
        else emu_launch(grid, 512, 2 * sm, [&]() { ntt12_persist_kernel<FWD, LZ, 3>(in, out, FWD ? P.twf[base] : P.twi[base], FWD ? P.headf[base] : P.headi[base], m, cnt); }); } while (0)
        if (forward) { if (m.lazy == 2) EMU_NTT12(true, 2); else if (m.lazy == 1) EMU_NTT12(true, 1); else EMU_NTT12(true, 0); }
        else { if (m.lazy == 2) EMU_NTT12(false, 2); else if (m.lazy == 1) EMU_NTT12(false, 1); else EMU_NTT12(false, 0); }
#undef EMU_NTT12
    } else {
        if (forward) emu_launch((unsigned)count, thr, sm, [&]() { ntt_fwd_kernel<0>(in, out, P.twf[base], P.headf[base], P.mod[base], P.logn); });
        else emu_launch((unsigned)count, thr, sm, [&]() { ntt_inv_kernel<0>(in, out, P.twi[base], P.headi[base], P.mod[base], P.logn); });
    }
    return 0;
}

int emu_poly_op(emu_ctx *c, uint32_t base, int op, const uint64_t *a, const uint64_t *b, uint64_t scalar,
                uint64_t *out, size_t words) {
    const Modulus m = c->hs.P.mod[base];
    if (op == OP_SCALAR_MUL) scalar %= m.m;
    emu_launch(4, 64, 0, [&]() { poly_op_kernel(m, op, a, b, scalar, out, words); });
    return 0;
}

// Same launch sequence as run_pairs() in exacto_b200/csrc/api.cu.
int emu_dbfv_mul(emu_ctx *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1, const uint64_t *ct2,
                 const uint64_t *rlk, uint32_t num_keys, uint64_t *out, size_t pairs, uint32_t flags,
                 uint32_t limb_mask) {
    const bool per_product = (flags & 0x80000000u) != 0;   // emulator-only flags
    const bool wide_relin = (flags & 0x40000000u) != 0;
    flags &= 0x3fffffffu;
    HostSetup &hs = c->hs;
    if (hs.mul_status != EXB_OK) { g_emu_err = hs.mul_error; return hs.mul_status; }
    HostPlan hp;
    int rc = host_build_plan(d, base, pm, flags, limb_mask, &hp, &g_emu_err);
    if (rc) return rc;
    DeviceParams P = hs.P;
    const uint32_t G = num_keys < hs.gadget_digits ? num_keys : hs.gadget_digits;
    P.gadget_digits = G;
    const size_t n = hs.n, A = hs.aux_moduli.size();
    const MulPlan &M = hp.M;
    const size_t nx = M.num_limbs - hp.num_low;
    const bool small = P.sb.enabled && P.logn == 12;
    std::vector<u64> ext(small ? (pairs * 2 * d * 2 * P.sb.K * n + 1) / 2 + 2
                               : pairs * 2 * d * 2 * (1 + A) * n),
        r01(pairs * M.num_products * 2 * n + 1);
    std::vector<u64> excess(pairs * nx * 2 * n + 1), rlk_mont((size_t)num_keys * 2 * n + 1);
    std::vector<u64> digbuf((pairs * M.num_products * (G ? G : 1) * n * hs.digit_bytes() + 7) / 8 + 1);
    const Modulus mq = P.mod[0];
    const size_t kw = (size_t)num_keys * 2 * n;
    if (kw) emu_launch(4, 64, 0, [&]() { poly_op_kernel(mq, OP_TO_MONT, rlk, nullptr, 0, rlk_mont.data(), kw); });

    const unsigned thr = emu_block_threads(P);
    u64 *extp = ext.data(), *r01p = r01.data(), *xp = excess.data();
    const u64 *rk = rlk_mont.data();
    if (small) {
        u32 *exts = reinterpret_cast<u32 *>(extp);
        emu_launch((unsigned)(pairs * 4 * d), thr, n * 8 + (size_t)P.sb.K * n * 4, [&]() { lift32_kernel(P, d, M.need_lhs, M.need_rhs, ct1, ct2, exts); });
        const size_t sm32 = n * 8 + (size_t)P.sb.K * n * 4;
        // the emulator always takes the per-limb kernel when it is legal (large-batch decision), so the CPU tier
        // covers it on small inputs; `per_product` forces the other kernel
        const bool per_limb = !per_product && tensor_sums_per_limb(P, M, (size_t)1 << 20);
        const u32 c2 = per_limb ? 1u : 0u;
        const unsigned tgrid = (unsigned)(pairs * M.num_products * (per_limb ? 1 : 3));
        if (per_limb) {
            const bool r64 = tensor01_needs_r64(P, M);
            const size_t sm01 = smem_tensor01(r64);
            if (r64) emu_launch((unsigned)(pairs * M.num_duos * 2), thr, sm01, [&]() { tensor01_kernel<true>(P, M, ct1, ct2, exts, r01p); });
            else emu_launch((unsigned)(pairs * M.num_duos * 2), thr, sm01, [&]() { tensor01_kernel<false>(P, M, ct1, ct2, exts, r01p); });
        }
        std::vector<u64> wide(wide_relin ? relin_wide_scratch_bytes(P, M, pairs) / 8 + 1 : 1);
        u64 *wp = wide.data();
        const unsigned rgrid = (unsigned)(pairs * M.num_limbs);
        auto run = [&](auto *dg) {
            typedef typename std::remove_pointer<decltype(dg)>::type DigT;
            emu_launch(tgrid, thr, sm32, [&]() { tensor32_kernel<DigT>(P, M, ct1, ct2, exts, r01p, dg, c2); });
            if (wide_relin) emu_launch(rgrid * (G + 1), thr, n * 8, [&]() { relin12_wide_kernel<DigT>(P, M, r01p, dg, rk, wp, c2); });
            else emu_launch(rgrid, thr, n * 24, [&]() { relin12_kernel<DigT>(P, M, r01p, dg, rk, out, xp, c2); });
        };
        if (hs.digit_kind() == 1) run(reinterpret_cast<int32_t *>(digbuf.data()));
        else if (hs.digit_kind() == 2) run(reinterpret_cast<int8_t *>(digbuf.data()));
        else run(reinterpret_cast<int16_t *>(digbuf.data()));
        if (wide_relin) emu_launch(4, 256, 0, [&]() { relin_reduce_kernel(P, M, wp, out, xp, pairs); });
    } else if (P.logn == 12) {
        emu_launch((unsigned)(pairs * 4 * d), thr, n * 16, [&]() { lift_kernel<12>(P, d, M.need_lhs, M.need_rhs, ct1, ct2, extp); });
        auto run = [&](auto *dg) {
            typedef typename std::remove_pointer<decltype(dg)>::type DigT;
            emu_launch((unsigned)(pairs * M.num_products * 3), thr, n * 8 * (1 + A), [&]() { tensor_kernel<12, DigT>(P, M, ct1, extp, r01p, dg, 0u); });
            emu_launch((unsigned)(pairs * M.num_limbs), thr, n * 24, [&]() { relin12_kernel<DigT>(P, M, r01p, dg, rk, out, xp, 0u); });
        };
        if (hs.digit_kind() == 1) run(reinterpret_cast<int32_t *>(digbuf.data()));
        else if (hs.digit_kind() == 2) run(reinterpret_cast<int8_t *>(digbuf.data()));
        else run(reinterpret_cast<int16_t *>(digbuf.data()));
    } else {
        emu_launch((unsigned)(pairs * 4 * d), thr, n * 16, [&]() { lift_kernel<0>(P, d, M.need_lhs, M.need_rhs, ct1, ct2, extp); });
        auto run = [&](auto *dg) {
            typedef typename std::remove_pointer<decltype(dg)>::type DigT;
            emu_launch((unsigned)(pairs * M.num_products * 3), thr, n * 8 * (1 + A), [&]() { tensor_kernel<0, DigT>(P, M, ct1, extp, r01p, dg, 0u); });
            emu_launch((unsigned)(pairs * M.num_limbs), thr, n * 24, [&]() { relin_kernel<0, DigT>(P, M, r01p, dg, rk, out, xp, 0u); });
        };
        if (hs.digit_kind() == 1) run(reinterpret_cast<int32_t *>(digbuf.data()));
        else if (hs.digit_kind() == 2) run(reinterpret_cast<int8_t *>(digbuf.data()));
        else run(reinterpret_cast<int16_t *>(digbuf.data()));
    }
    const u64 q = hs.ct_moduli[0];
    for (uint32_t j = d; j + 1 < 2 * d; j++) {
        if (hp.excess_index[j] < 0) continue;
        for (uint32_t i = 0; i < d; i++) {
            const int64_t rep = hp.reps[(size_t)(j - d) * d + i];
            if (rep == 0) continue;
            bool computed = false;
            for (uint32_t l = 0; l < M.num_limbs; l++) if (M.limb_k[l] == i) computed = true;
            if (!computed) continue;
            const u64 mag = rep < 0 ? (u64)(-(rep + 1)) + 1 : (u64)rep;
            const u64 s_mont = (u64)(((unsigned __int128)(mag % q) << 64) % q);
            u64 *ol = out + (size_t)i * 2 * n;
            const u64 *el = xp + (size_t)hp.excess_index[j] * 2 * n;
            emu_launch(4, 64, 0, [&]() {
                reduce_mac_kernel(mq, ol, el, s_mont, rep < 0 ? 1 : 0, (size_t)d * 2 * n, nx * 2 * n, (u32)(2 * n), pairs);
            });
        }
    }
    return 0;
}

// Same launch sequences as exb_bfv_mul_no_relin / exb_bfv_relinearize / exb_gadget_decompose in api.cu.
int emu_bfv_mul_no_relin(emu_ctx *c, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out3, size_t pairs) {
    HostSetup &hs = c->hs;
    if (hs.mul_status != EXB_OK) { g_emu_err = hs.mul_error; return hs.mul_status; }
    HostPlan hp;
    int rc = host_build_plan(1, 2, 0, 0, 0, &hp, &g_emu_err);
    if (rc) return rc;
    const DeviceParams P = hs.P;
    const MulPlan &M = hp.M;
    const size_t n = hs.n, A = hs.aux_moduli.size();
    const bool small = P.sb.enabled && P.logn == 12;
    std::vector<u64> ext(small ? (pairs * 2 * 2 * P.sb.K * n + 1) / 2 + 2 : pairs * 2 * 2 * (1 + A) * n), r01(pairs * 3 * n + 1);
    u64 *extp = ext.data(), *r01p = r01.data();
    const unsigned thr = emu_block_threads(P);
    int16_t *nodig = nullptr;
    if (small) {
        u32 *exts = reinterpret_cast<u32 *>(extp);
        emu_launch((unsigned)(pairs * 4), thr, n * 8 + (size_t)P.sb.K * n * 4, [&]() { lift32_kernel(P, 1, 1u, 1u, ct1, ct2, exts); });
        emu_launch((unsigned)(pairs * 3), thr, n * 8 + (size_t)P.sb.K * n * 4, [&]() { tensor32_kernel<int16_t>(P, M, ct1, ct2, exts, r01p, nodig, 2u); });
    } else if (P.logn == 12) {
        emu_launch((unsigned)(pairs * 4), thr, n * 16, [&]() { lift_kernel<12>(P, 1, 1u, 1u, ct1, ct2, extp); });
        emu_launch((unsigned)(pairs * 3), thr, n * 8 * (1 + A), [&]() { tensor_kernel<12, int16_t>(P, M, ct1, extp, r01p, nodig, 1u); });
    } else {
        emu_launch((unsigned)(pairs * 4), thr, n * 16, [&]() { lift_kernel<0>(P, 1, 1u, 1u, ct1, ct2, extp); });
        emu_launch((unsigned)(pairs * 3), thr, n * 8 * (1 + A), [&]() { tensor_kernel<0, int16_t>(P, M, ct1, extp, r01p, nodig, 1u); });
    }
    return emu_ntt(c, 0, 1, r01p, out3, pairs * 3);
}

int emu_bfv_relinearize(emu_ctx *c, const uint64_t *ct, const uint64_t *rlk, uint32_t num_keys, uint64_t *out,
                        size_t pairs, int wide) {
    HostSetup &hs = c->hs;
    HostPlan hp;
    int rc = host_build_plan(1, 2, 0, 0, 0, &hp, &g_emu_err);
    if (rc) return rc;
    DeviceParams P = hs.P;
    const uint32_t G = num_keys < hs.gadget_digits ? num_keys : hs.gadget_digits;
    P.gadget_digits = G;
    const MulPlan &M = hp.M;
    const size_t n = hs.n, kw = (size_t)num_keys * 2 * n;
    std::vector<u64> r01(pairs * 2 * n + 1), c2(pairs * n + 1), rlk_mont(kw + 1), widebuf(wide ? relin_wide_scratch_bytes(P, M, pairs) / 8 + 1 : 1);
    for (size_t b = 0; b < pairs; b++) {
        memcpy(r01.data() + b * 2 * n, ct + b * 3 * n, 2 * n * 8);
        memcpy(c2.data() + b * n, ct + b * 3 * n + 2 * n, n * 8);
    }
    const Modulus mq = P.mod[0];
    if (kw) emu_launch(4, 64, 0, [&]() { poly_op_kernel(mq, OP_TO_MONT, rlk, nullptr, 0, rlk_mont.data(), kw); });
    if ((rc = emu_ntt(c, 0, 0, c2.data(), c2.data(), pairs))) return rc;
    std::vector<u64> digbuf((pairs * (G ? G : 1) * n * hs.digit_bytes() + 7) / 8 + 1);
    const u64 *c2p = c2.data(), *r01p = r01.data(), *rk = rlk_mont.data();
    u64 *wp = widebuf.data(), *none = nullptr;
    const unsigned thr = emu_block_threads(P);
    auto run = [&](auto *dg) {
        typedef typename std::remove_pointer<decltype(dg)>::type DigT;
        emu_launch(4, 64, 0, [&]() { gadget_digits_kernel<DigT, false>(P, c2p, dg, pairs); });
        if (P.logn == 12 && wide) emu_launch((unsigned)(pairs * (G + 1)), thr, n * 8, [&]() { relin12_wide_kernel<DigT>(P, M, r01p, dg, rk, wp, 3u); });
        else if (P.logn == 12) emu_launch((unsigned)pairs, thr, n * 24, [&]() { relin12_kernel<DigT>(P, M, r01p, dg, rk, out, none, 3u); });
        else emu_launch((unsigned)pairs, thr, n * 24, [&]() { relin_kernel<0, DigT>(P, M, r01p, dg, rk, out, none, 1u); });
    };
    if (hs.digit_kind() == 1) run(reinterpret_cast<int32_t *>(digbuf.data()));
    else if (hs.digit_kind() == 2) run(reinterpret_cast<int8_t *>(digbuf.data()));
    else run(reinterpret_cast<int16_t *>(digbuf.data()));
    if (P.logn == 12 && wide) emu_launch(4, 256, 0, [&]() { relin_reduce_kernel(P, M, wp, out, none, pairs); });
    return 0;
}

int emu_gadget_decompose(emu_ctx *c, const uint64_t *coeffs, uint64_t *out, size_t count) {
    const DeviceParams P = c->hs.P;
    emu_launch(4, 64, 0, [&]() { gadget_digits_kernel<u64, true>(P, coeffs, out, count); });
    return 0;
}

// 1 when dbfv_mul with this plan takes the per-limb tensor01_kernel for components 0/1.
int emu_tensor_per_limb(emu_ctx *c, uint64_t base, uint32_t d, uint64_t pm, uint32_t flags, uint32_t limb_mask) {
    HostPlan hp;
    if (host_build_plan(d, base, pm, flags, limb_mask, &hp, &g_emu_err)) return -1;
    return tensor_sums_per_limb(c->hs.P, hp.M, (size_t)1 << 20) ? 1 : 0;
}

// Same launch as launch_galois() in exacto_b200/csrc/kernels.cu.
int emu_bfv_apply_automorphism(emu_ctx *c, const uint64_t *ct, uint64_t element, const uint64_t *gk,
                               uint64_t *out, size_t count) {
    HostSetup &hs = c->hs;
    const DeviceParams &P = hs.P;
    const size_t n = hs.n, kw = (size_t)hs.gadget_digits * 2 * n;
    std::vector<u64> gk_mont(kw + 1);
    const Modulus mq = P.mod[0];
    emu_launch(4, 64, 0, [&]() { poly_op_kernel(mq, OP_TO_MONT, gk, nullptr, 0, gk_mont.data(), kw); });
    const u64 *gm = gk_mont.data();
    const u32 k = (u32)(element % (2 * n));
    if (P.logn == 12) emu_launch((unsigned)count, kThreads12, n * 24, [&]() { galois_kernel<12>(P, ct, gm, k, out); });
    else emu_launch((unsigned)count, emu_block_threads(P), n * 24, [&]() { galois_kernel<0>(P, ct, gm, k, out); });
    return 0;
}

// Same launch as launch_decrypt() in exacto_b200/csrc/kernels.cu.
int emu_bfv_decrypt(emu_ctx *c, const uint64_t *ct, uint32_t ncomp, const uint64_t *sk_ntt, uint64_t *out, size_t count) {
    const DeviceParams &P = c->hs.P;
    const size_t n = c->hs.n;
    if (P.logn == 12) emu_launch((unsigned)count, kThreads12, n * 8, [&]() { decrypt_kernel<12>(P, ct, ncomp, sk_ntt, out); });
    else emu_launch((unsigned)count, emu_block_threads(P), n * 8, [&]() { decrypt_kernel<0>(P, ct, ncomp, sk_ntt, out); });
    return 0;
}

// ---- multi-prime ciphertext modulus: the launch sequence of launch_rns_mul / launch_rns_relinearize -----------
static void emu_ntt_plan(const Modulus &m, const Tw *tw, const TwHead &head, u32 logn, bool fwd, u64 *buf, size_t count) {
    const u32 n = 1u << logn;
    if (logn == 12) {
        const unsigned grid = count < 3 ? (unsigned)count : 3u;
        const u32 cnt = (u32)count;
        if (fwd) emu_launch(grid, 512, 2 * (size_t)n * 8, [&]() { ntt12_persist_kernel<true, 0, 3>(buf, buf, tw, head, m, cnt); });
        else emu_launch(grid, 512, 2 * (size_t)n * 8, [&]() { ntt12_persist_kernel<false, 0, 3>(buf, buf, tw, head, m, cnt); });
        return;
    }
    unsigned thr = n / 2;
    if (thr < 32) thr = 32;
    if (thr > 256) thr = 256;
    if (fwd) emu_launch((unsigned)count, thr, (size_t)n * 8, [&]() { ntt_fwd_kernel<0>(buf, buf, tw, head, m, logn); });
    else emu_launch((unsigned)count, thr, (size_t)n * 8, [&]() { ntt_inv_kernel<0>(buf, buf, tw, head, m, logn); });
}

int emu_rns_info(const emu_ctx *c, uint32_t *L, uint32_t *K, uint64_t *ext_primes) {
    if (!c->hs.rns_enabled) { g_emu_err = c->hs.mul_error; return c->hs.mul_status ? c->hs.mul_status : EXB_NOT_IMPLEMENTED; }
    *L = c->hs.R.L; *K = c->hs.R.K;
    for (size_t i = 0; i < c->hs.ext_primes.size(); i++) ext_primes[i] = c->hs.ext_primes[i];
    return 0;
}

// mode 0: dbfv_mul / bfv_mul_and_relin -> out [pairs][d][2][L][n]; mode 1: bfv_mul_no_relin -> [pairs][3][L][n];
// mode 2: relinearize of ct1 = [pairs][3][L][n] -> [pairs][2][L][n].  rlk [num_keys][2][L][n].
int emu_rns_mul(emu_ctx *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1, const uint64_t *ct2,
                const uint64_t *rlk, uint32_t num_keys, uint64_t *out, size_t pairs, int mode) {
    HostSetup &hs = c->hs;
    if (!hs.rns_enabled) { g_emu_err = hs.mul_error; return hs.mul_status ? hs.mul_status : EXB_NOT_IMPLEMENTED; }
    HostPlan hp;
    int rc = host_build_plan(d, base, pm, 0, 0, &hp, &g_emu_err);
    if (rc) return rc;
    const RnsConsts R = hs.R;
    const RnsPlans &T = hs.T;
    const MulPlan M = hp.M;
    const size_t n = hs.n;
    const u32 G = num_keys < hs.gadget_digits ? num_keys : hs.gadget_digits;
    std::vector<u64> key((size_t)num_keys * 2 * R.L * n + 1);
    for (size_t i = 0; i + 1 < key.size(); i++) key[i] = rlk[i];
    u64 *kp = key.data();
    const size_t kpolys = (size_t)num_keys * 2 * R.L;
    if (kpolys) emu_launch(2, 64, 0, [&]() { rns_to_mont_kernel(R, kp, kpolys); });
    if (mode == 2) {
        RnsShape S{1, 1, 1, pairs};
        const size_t nd = pairs * R.gadget_digits;
        std::vector<u64> c2((size_t)R.L * pairs * n + 1), dig((size_t)R.L * nd * n + 1);
        u64 *c2p = c2.data(), *dp = dig.data();
        emu_launch(2, 64, 0, [&]() { rns_gather_c2_kernel(R, pairs, ct1, c2p); });
        for (u32 l = 0; l < R.L; l++) emu_ntt_plan(R.q[l], T.twi_q[l], T.headi_q[l], R.logn, false, c2p + l * pairs * n, pairs);
        emu_launch(2, 64, 0, [&]() { rns_digits_kernel(R, M, S, c2p, 1u, dp); });
        for (u32 l = 0; l < R.L; l++) emu_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, dp + l * nd * n, nd);
        emu_launch(2, 64, 0, [&]() { rns_relin_mac_kernel(R, M, S, nullptr, ct1, dp, kp, G, out); });
        return 0;
    }
    const size_t npoly = pairs * 4 * M.d, nt = pairs * M.num_products * 3;
    RnsShape S{M.d, M.num_products, M.num_limbs, pairs};
    const size_t nc = pairs * M.num_limbs * 2, nd = pairs * M.num_limbs * R.gadget_digits;
    std::vector<u64> coef((size_t)R.L * npoly * n + 1), ext((size_t)R.K * npoly * n + 1), tens((size_t)R.K * nt * n + 1),
        res((size_t)R.L * nt * n + 1), c01((size_t)R.L * nc * n + 1), dig((size_t)R.L * nd * n + 1);
    u64 *cp = coef.data(), *ep = ext.data(), *tp = tens.data(), *rp = res.data(), *c01p = c01.data(), *dp = dig.data();
    emu_launch(2, 64, 0, [&]() { rns_gather_kernel(R, S, ct1, ct2, cp); });
    for (u32 l = 0; l < R.L; l++) emu_ntt_plan(R.q[l], T.twi_q[l], T.headi_q[l], R.logn, false, cp + l * npoly * n, npoly);
    emu_launch(2, 64, 0, [&]() { rns_extend_kernel(R, S, cp, ep); });
    for (u32 k = 0; k < R.K; k++) emu_ntt_plan(R.e[k], T.twf_e[k], T.headf_e[k], R.logn, true, ep + k * npoly * n, npoly);
    emu_launch(2, 64, 0, [&]() { rns_tensor_kernel(R, M, S, ep, tp); });
    for (u32 k = 0; k < R.K; k++) emu_ntt_plan(R.e[k], T.twi_e[k], T.headi_e[k], R.logn, false, tp + k * nt * n, nt);
    emu_launch(2, 64, 0, [&]() { rns_scale_kernel(R, nt, tp, rp); });
    if (mode == 1) {
        for (u32 l = 0; l < R.L; l++) emu_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, rp + l * nt * n, nt);
        emu_launch(2, 64, 0, [&]() { rns_scatter3_kernel(R, pairs, rp, out); });
        return 0;
    }
    emu_launch(2, 64, 0, [&]() { rns_sum01_kernel(R, M, S, rp, c01p); });
    emu_launch(2, 64, 0, [&]() { rns_digits_kernel(R, M, S, rp, 3u, dp); });
    for (u32 l = 0; l < R.L; l++) {
        emu_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, c01p + l * nc * n, nc);
        emu_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, dp + l * nd * n, nd);
    }
    emu_launch(2, 64, 0, [&]() { rns_relin_mac_kernel(R, M, S, c01p, nullptr, dp, kp, G, out); });
    return 0;
}

}  // extern "C"
