// rns_kernels.cu -- multi-prime ciphertext modulus (L > 1): element-wise kernels around the batched NTTs.
//
//   bfv_mul_generic_rns   bfv/eval.rs:113-147      gather -> INTT_q -> rns_extend -> NTT_e -> rns_tensor -> INTT_e
//                                                  -> rns_scale (round(p t / Q) mod q_l)
//   relinearize           bfv/keyswitch.rs:59-101  rns_digits (to_coeff_poly u128 + gadget_decompose) -> NTT_q
//                                                  -> rns_relin_mac
//   dbfv_mul              dbfv/eval.rs:82-149      the same over the live products; per-k sums by linearity
//
// The per-coefficient arithmetic is in rns.cuh.  Layout of the work buffers is prime-major, [prime][poly][n], so
// every transform batch is one contiguous launch of the single-prime NTT kernels.
#include "rns.cuh"
#include "kernels.cuh"

namespace exb {

// product -> (lhs limb, rhs limb) and computed limb -> k, as in MulPlan (host_setup.hpp)
struct RnsShape {
    u32 d, NP, NL;          // limbs per ciphertext, live products, computed output limbs
    size_t pairs;
};

// ct1 / ct2 [pairs][d][2][L][n]  ->  coef[l][p][n],  p = ((pair * 2 + side) * d + limb) * 2 + comp
__global__ void rns_gather_kernel(const __grid_constant__ RnsConsts R, RnsShape S, const u64 *__restrict__ ct1,
                                  const u64 *__restrict__ ct2, u64 *__restrict__ coef) {
    const size_t n = R.n, npoly = S.pairs * 4 * S.d, total = npoly * R.L * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, pl = idx / n, l = pl % R.L, p = pl / R.L;
        const size_t comp = p & 1, limb = (p >> 1) % S.d, side = (p / (2 * S.d)) & 1, pair = p / (4 * S.d);
        const u64 *src = side ? ct2 : ct1;
        coef[(l * npoly + p) * n + j] = src[((((pair * S.d + limb) * 2 + comp) * R.L) + l) * n + j];
    }
}

// coef[l][p][n] (coefficient domain) -> ext[k][p][n]: centred CRT value mod e_k; right-hand operands (side 1) in
// Montgomery form so the tensor needs one REDC per product.
__global__ void rns_extend_kernel(const __grid_constant__ RnsConsts R, RnsShape S, const u64 *__restrict__ coef,
                                  u64 *__restrict__ ext) {
    const size_t n = R.n, npoly = S.pairs * 4 * S.d, total = npoly * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, p = idx / n;
        const bool rhs = ((p / (2 * S.d)) & 1) != 0;
        u64 x[kRnsMaxL];
        for (u32 l = 0; l < R.L; l++) x[l] = coef[(l * npoly + p) * n + j];
        u32 mag[kMwQ];
        const bool neg = rns_centered(x, R, mag);
        for (u32 k = 0; k < R.K; k++) {
            u64 r = rns_mag_mod_e(mag, neg, R, k);
            if (rhs) r = shoup(r, R.e[k].r_mod, R.e[k].r_mod_s, R.e[k].m);
            ext[(k * npoly + p) * n + j] = r;
        }
    }
}

// ext (NTT domain mod e_k) -> tens[k][(pair * NP + prod) * 3 + comp][n]:  c0 d0,  c0 d1 + c1 d0,  c1 d1
__global__ void rns_tensor_kernel(const __grid_constant__ RnsConsts R, const __grid_constant__ MulPlan M, RnsShape S,
                                  const u64 *__restrict__ ext, u64 *__restrict__ tens) {
    const size_t n = R.n, npoly = S.pairs * 4 * S.d, nt = S.pairs * S.NP * 3, total = (size_t)R.K * S.pairs * S.NP * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, prod = r % S.NP, pair = (r / S.NP) % S.pairs, k = r / (S.NP * S.pairs);
        const Modulus &m = R.e[k];
        const u64 *base = ext + k * npoly * n;
        const size_t pl = ((pair * 2 + 0) * S.d + M.prod_i[prod]) * 2, pr = ((pair * 2 + 1) * S.d + M.prod_j[prod]) * 2;
        const u64 c0 = base[pl * n + j], c1 = base[(pl + 1) * n + j], d0 = base[pr * n + j], d1 = base[(pr + 1) * n + j];
        u64 *o = tens + ((k * nt) + (pair * S.NP + prod) * 3) * n + j;
        o[0] = csub(mont_mul_lazy(c0, d0, m.m, m.minv_neg), m.m);
        o[n] = mod_add(csub(mont_mul_lazy(c0, d1, m.m, m.minv_neg), m.m), csub(mont_mul_lazy(c1, d0, m.m, m.minv_neg), m.m), m.m);
        o[2 * n] = csub(mont_mul_lazy(c1, d1, m.m, m.minv_neg), m.m);
    }
}

// tens[k][t][n] (coefficient domain mod e_k) -> res[l][t][n] = round(p t / Q) mod q_l
__global__ void rns_scale_kernel(const __grid_constant__ RnsConsts R, size_t nt, const u64 *__restrict__ tens,
                                 u64 *__restrict__ res) {
    const size_t n = R.n, total = nt * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        u64 tr[kRnsMaxK], out[kRnsMaxL];
        for (u32 k = 0; k < R.K; k++) tr[k] = tens[k * nt * n + idx];
        rns_scale_coeff(tr, R, out);
        for (u32 l = 0; l < R.L; l++) res[l * nt * n + idx] = out[l];
    }
}

// res (coefficient domain) -> c01[l][(pair * NL + limb) * 2 + comp][n]: per-k sums of components 0 / 1
// (dbfv/eval.rs:125-132; sums mod q_l commute with the NTT that follows)
__global__ void rns_sum01_kernel(const __grid_constant__ RnsConsts R, const __grid_constant__ MulPlan M, RnsShape S,
                                 const u64 *__restrict__ res, u64 *__restrict__ c01) {
    const size_t n = R.n, nt = S.pairs * S.NP * 3, nc = S.pairs * S.NL * 2, total = (size_t)R.L * nc * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, comp = r & 1, limb = (r >> 1) % S.NL, pair = (r / (2 * S.NL)) % S.pairs;
        const size_t l = r / nc;
        const u64 q = R.q[l].m;
        const u32 k = M.limb_k[limb];
        const u32 i_lo = k >= S.d ? k - S.d + 1 : 0, i_hi = k < S.d ? k : S.d - 1;
        u64 s = 0;
        for (u32 i = i_lo; i <= i_hi; i++) {
            const size_t prod = (size_t)M.prod_of[i][k - i];
            s = mod_add(s, res[(l * nt + (pair * S.NP + prod) * 3 + comp) * n + j], q);
        }
        c01[idx] = s;
    }
}

// res component 2 (coefficient domain, every q_l) -> dig[l][(pair * NL + limb) * G + g][n]: gadget digits of the
// reference's truncated to_coeff_poly value, each reduced mod q_l (RnsPoly::from_coeff_poly), summed per limb.
// stride3 = 3 when res holds (c0, c1, c2) per product (the fused pipeline), 1 for a plain [count][n] c2 batch.
__global__ void rns_digits_kernel(const __grid_constant__ RnsConsts R, const __grid_constant__ MulPlan M, RnsShape S,
                                  const u64 *__restrict__ res, u32 stride3, u64 *__restrict__ dig) {
    const size_t n = R.n, nt = S.pairs * S.NP * stride3, nd = S.pairs * S.NL * R.gadget_digits, total = S.pairs * S.NL * n;
    const u32 G = R.gadget_digits;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, limb = r % S.NL, pair = r / S.NL;
        const u32 k = M.limb_k[limb];
        const u32 i_lo = k >= S.d ? k - S.d + 1 : 0, i_hi = k < S.d ? k : S.d - 1;
        for (u32 g = 0; g < G; g++)
            for (u32 l = 0; l < R.L; l++) dig[(l * nd + (pair * S.NL + limb) * G + g) * n + j] = 0;
        for (u32 i = i_lo; i <= i_hi; i++) {
            const size_t prod = (size_t)M.prod_of[i][k - i];
            u64 x[kRnsMaxL];
            for (u32 l = 0; l < R.L; l++) x[l] = res[(l * nt + (pair * S.NP + prod) * stride3 + (stride3 - 1)) * n + j];
            const u64 c = rns_to_coeff_truncated(x, R), qm = R.bigq_lo;
            exb_i128 remaining = c > qm / 2 ? (exb_i128)c - (exb_i128)qm : (exb_i128)c;
            for (u32 g = 0; g < G; g++) {
                const u64 dg = rns_gadget_digit(remaining, R.gadget_base, qm);
                for (u32 l = 0; l < R.L; l++) {
                    u64 *o = dig + (l * nd + (pair * S.NL + limb) * G + g) * n + j;
                    *o = mod_add(*o, dg % R.q[l].m, R.q[l].m);
                }
            }
        }
    }
}

// out[pair][k][comp][l][n] = c01_ntt + sum_g dig_ntt[g] * rlk_mont[g][comp][l]   (bfv/keyswitch.rs:86-95)
// c01 may instead be taken in place from a [pairs][3][L][n] NTT-domain ciphertext (standalone relinearize).
__global__ void rns_relin_mac_kernel(const __grid_constant__ RnsConsts R, const __grid_constant__ MulPlan M, RnsShape S,
                                     const u64 *__restrict__ c01, const u64 *__restrict__ ct3, const u64 *__restrict__ dig,
                                     const u64 *__restrict__ rlk_mont, u32 G, u64 *__restrict__ out) {
    const size_t n = R.n, nc = S.pairs * S.NL * 2, nd = S.pairs * S.NL * R.gadget_digits;
    const size_t total = S.pairs * S.NL * 2 * R.L * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, l = r % R.L, comp = (r / R.L) & 1, limb = (r / (2 * R.L)) % S.NL;
        const size_t pair = r / (2 * R.L * S.NL);
        const Modulus &m = R.q[l];
        u64 acc = ct3 ? ct3[((pair * 3 + comp) * R.L + l) * n + j]
                      : c01[(l * nc + (pair * S.NL + limb) * 2 + comp) * n + j];
        for (u32 g = 0; g < G; g++) {
            const u64 x = dig[(l * nd + (pair * S.NL + limb) * R.gadget_digits + g) * n + j];
            const u64 kk = rlk_mont[(((size_t)g * 2 + comp) * R.L + l) * n + j];
            acc = mod_add(acc, csub(mont_mul_lazy(x, kk, m.m, m.minv_neg), m.m), m.m);
        }
        const u32 k = M.limb_k[limb];
        out[(((pair * S.d + k) * 2 + comp) * R.L + l) * n + j] = acc;
    }
}

// res[l][t][n] (t = pair * 3 + comp, NTT domain) -> out3[pair][3][L][n]      (bfv_mul_no_relin)
__global__ void rns_scatter3_kernel(const __grid_constant__ RnsConsts R, size_t pairs, const u64 *__restrict__ res,
                                    u64 *__restrict__ out3) {
    const size_t n = R.n, nt = pairs * 3, total = nt * R.L * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, l = r % R.L, t = r / R.L;
        out3[idx] = res[(l * nt + t) * n + j];
    }
}

// ct3[pair][3][L][n] component 2 -> c2[l][pair][n]                                (standalone relinearize)
__global__ void rns_gather_c2_kernel(const __grid_constant__ RnsConsts R, size_t pairs, const u64 *__restrict__ ct3,
                                     u64 *__restrict__ c2) {
    const size_t n = R.n, total = pairs * R.L * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, pair = r % pairs, l = r / pairs;
        c2[idx] = ct3[((pair * 3 + 2) * R.L + l) * n + j];
    }
}

// decrypt (bfv/encrypt.rs:111-178), step 1: phase[l][b][n] = c0 + c1 s + c2 s^2 + ... mod q_l (NTT domain);
// ct [count][ncomp][L][n], sk [L][n].
__global__ void rns_phase_kernel(const __grid_constant__ RnsConsts R, const u64 *__restrict__ ct, u32 ncomp,
                                 const u64 *__restrict__ sk, u64 *__restrict__ phase, size_t count) {
    const size_t n = R.n, total = count * R.L * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx % n, r = idx / n, b = r % count, l = r / count;
        const Modulus &m = R.q[l];
        const u64 s_m = csub(mont_mul_lazy(sk[l * n + j], m.r2_mod, m.m, m.minv_neg), m.m);          // s * 2^64
        u64 acc = ct[((b * ncomp + 0) * R.L + l) * n + j], spow = s_m;
        for (u32 i = 1; i < ncomp; i++) {
            acc = mod_add(acc, csub(mont_mul_lazy(ct[((b * ncomp + i) * R.L + l) * n + j], spow, m.m, m.minv_neg), m.m), m.m);
            if (i + 1 < ncomp) spow = csub(mont_mul_lazy(spow, s_m, m.m, m.minv_neg), m.m);
        }
        phase[idx] = acc;
    }
}
// step 2 (after INTT per prime): out[b][n] = round(p x / Q) mod p with x the CRT value in [0, Q)
__global__ void rns_decrypt_kernel(const __grid_constant__ RnsConsts R, const u64 *__restrict__ phase,
                                   u64 *__restrict__ out, size_t count) {
    const size_t n = R.n, total = count * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        u64 x[kRnsMaxL];
        for (u32 l = 0; l < R.L; l++) x[l] = phase[l * total + idx];
        out[idx] = rns_decrypt_coeff(x, R);
    }
}

// key [G][2][L][n] -> Montgomery form per prime (in place)
__global__ void rns_to_mont_kernel(const __grid_constant__ RnsConsts R, u64 *__restrict__ key, size_t polys) {
    const size_t n = R.n, total = polys * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const Modulus &m = R.q[(idx / n) % R.L];
        key[idx] = csub(mont_mul_lazy(key[idx], m.r2_mod, m.m, m.minv_neg), m.m);
    }
}

#ifndef EXB_HOST_EMUL
static inline unsigned grid_for(size_t total) {
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(blocks > cap ? cap : (blocks ? blocks : 1));
}

size_t rns_workspace_words(const RnsConsts &R, const MulPlan &M, size_t pairs) {
    const size_t n = R.n, npoly = pairs * 4 * M.d, nt = pairs * M.num_products * 3;
    return (size_t)R.L * npoly * n + (size_t)R.K * npoly * n + (size_t)R.K * nt * n + (size_t)R.L * nt * n +
           (size_t)R.L * pairs * M.num_limbs * 2 * n + (size_t)R.L * pairs * M.num_limbs * R.gadget_digits * n;
}

void launch_rns_to_mont(const RnsConsts &R, u64 *key, size_t polys, cudaStream_t s) {
    if (!polys) return;
    rns_to_mont_kernel<<<grid_for(polys * R.n), 256, 0, s>>>(R, key, polys);
    g_launch_count++;
}

// mode 0: multiply + relinearise -> out [pairs][d][2][L][n];  mode 1: bfv_mul_no_relin -> out [pairs][3][L][n]
void launch_rns_mul(const RnsConsts &R, const RnsPlans &T, const MulPlan &M, const u64 *ct1, const u64 *ct2,
                    const u64 *rlk_mont, u32 G, u64 *ws, u64 *out, size_t pairs, int mode, cudaStream_t s) {
    if (!pairs) return;
    const size_t n = R.n, npoly = pairs * 4 * M.d, nt = pairs * M.num_products * 3;
    RnsShape S{M.d, M.num_products, M.num_limbs, pairs};
    u64 *coef = ws, *ext = coef + (size_t)R.L * npoly * n, *tens = ext + (size_t)R.K * npoly * n;
    u64 *res = tens + (size_t)R.K * nt * n, *c01 = res + (size_t)R.L * nt * n;
    u64 *dig = c01 + (size_t)R.L * pairs * M.num_limbs * 2 * n;
    rns_gather_kernel<<<grid_for(npoly * R.L * n), 256, 0, s>>>(R, S, ct1, ct2, coef);
    for (u32 l = 0; l < R.L; l++)
        launch_ntt_plan(R.q[l], T.twi_q[l], T.headi_q[l], R.logn, false, coef + l * npoly * n, coef + l * npoly * n, npoly, s);
    rns_extend_kernel<<<grid_for(npoly * n), 256, 0, s>>>(R, S, coef, ext);
    for (u32 k = 0; k < R.K; k++)
        launch_ntt_plan(R.e[k], T.twf_e[k], T.headf_e[k], R.logn, true, ext + k * npoly * n, ext + k * npoly * n, npoly, s);
    rns_tensor_kernel<<<grid_for((size_t)R.K * pairs * M.num_products * n), 256, 0, s>>>(R, M, S, ext, tens);
    for (u32 k = 0; k < R.K; k++)
        launch_ntt_plan(R.e[k], T.twi_e[k], T.headi_e[k], R.logn, false, tens + k * nt * n, tens + k * nt * n, nt, s);
    rns_scale_kernel<<<grid_for(nt * n), 256, 0, s>>>(R, nt, tens, res);
    g_launch_count += 4;
    if (mode == 1) {
        for (u32 l = 0; l < R.L; l++)
            launch_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, res + l * nt * n, res + l * nt * n, nt, s);
        rns_scatter3_kernel<<<grid_for(nt * R.L * n), 256, 0, s>>>(R, pairs, res, out);
        g_launch_count++;
        return;
    }
    const size_t nc = pairs * M.num_limbs * 2, nd = pairs * M.num_limbs * R.gadget_digits;
    rns_sum01_kernel<<<grid_for((size_t)R.L * nc * n), 256, 0, s>>>(R, M, S, res, c01);
    rns_digits_kernel<<<grid_for(pairs * M.num_limbs * n), 256, 0, s>>>(R, M, S, res, 3u, dig);
    for (u32 l = 0; l < R.L; l++) {
        launch_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, c01 + l * nc * n, c01 + l * nc * n, nc, s);
        launch_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, dig + l * nd * n, dig + l * nd * n, nd, s);
    }
    rns_relin_mac_kernel<<<grid_for(pairs * M.num_limbs * 2 * R.L * n), 256, 0, s>>>(R, M, S, c01, nullptr, dig, rlk_mont, G, out);
    g_launch_count += 3;
}

// relinearize (bfv/keyswitch.rs:59-101) of [pairs][3][L][n] -> [pairs][2][L][n]; M is the d = 1 plan
void launch_rns_relinearize(const RnsConsts &R, const RnsPlans &T, const MulPlan &M, const u64 *ct3,
                            const u64 *rlk_mont, u32 G, u64 *ws, u64 *out, size_t pairs, cudaStream_t s) {
    if (!pairs) return;
    const size_t n = R.n, nd = pairs * R.gadget_digits;
    RnsShape S{1, 1, 1, pairs};
    u64 *c2 = ws, *dig = c2 + (size_t)R.L * pairs * n;
    rns_gather_c2_kernel<<<grid_for(pairs * R.L * n), 256, 0, s>>>(R, pairs, ct3, c2);
    for (u32 l = 0; l < R.L; l++)
        launch_ntt_plan(R.q[l], T.twi_q[l], T.headi_q[l], R.logn, false, c2 + l * pairs * n, c2 + l * pairs * n, pairs, s);
    rns_digits_kernel<<<grid_for(pairs * n), 256, 0, s>>>(R, M, S, c2, 1u, dig);
    for (u32 l = 0; l < R.L; l++)
        launch_ntt_plan(R.q[l], T.twf_q[l], T.headf_q[l], R.logn, true, dig + l * nd * n, dig + l * nd * n, nd, s);
    rns_relin_mac_kernel<<<grid_for(pairs * 2 * R.L * n), 256, 0, s>>>(R, M, S, nullptr, ct3, dig, rlk_mont, G, out);
    g_launch_count += 3;
}
// decrypt of [count][ncomp][L][n] with sk [L][n] -> [count][n]; ws holds L * count * n words
void launch_rns_decrypt(const RnsConsts &R, const RnsPlans &T, const u64 *ct, u32 ncomp, const u64 *sk, u64 *ws,
                        u64 *out, size_t count, cudaStream_t s) {
    if (!count) return;
    const size_t n = R.n;
    rns_phase_kernel<<<grid_for(count * R.L * n), 256, 0, s>>>(R, ct, ncomp, sk, ws, count);
    for (u32 l = 0; l < R.L; l++)
        launch_ntt_plan(R.q[l], T.twi_q[l], T.headi_q[l], R.logn, false, ws + l * count * n, ws + l * count * n, count, s);
    rns_decrypt_kernel<<<grid_for(count * n), 256, 0, s>>>(R, ws, out, count);
    g_launch_count += 2;
}
size_t rns_relin_workspace_words(const RnsConsts &R, size_t pairs) {
    return (size_t)R.L * pairs * R.n * (1 + (size_t)R.gadget_digits);
}
#endif  // EXB_HOST_EMUL

}  // namespace exb
