// kernels.cuh -- launch-side declarations shared by kernels.cu and api.cu.
#pragma once
#ifndef EXB_HOST_EMUL
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <atomic>

#include "host_setup.hpp"

namespace exb {

// dynamic shared memory of tensor01_kernel (n = 4096): i64 and i32 (r64: i64) per-limb accumulators + the base-q image,
// which two u32 small-prime images at a time reuse
inline size_t smem_tensor01(bool r64) { return 4096 * 8 + (r64 ? 4096 * 8 : 4096 * 4) + 4096 * 8; }

enum PolyOp { OP_ADD = 0, OP_SUB = 1, OP_NEG = 2, OP_MUL = 3, OP_SCALAR_MUL = 4, OP_TO_MONT = 5 };

#ifndef EXB_HOST_EMUL
// Ring-level batched kernels (device pointers, `count` polynomials of n words).
void launch_ntt_fwd(const DeviceParams &P, int base, const u64 *in, u64 *out, size_t count, cudaStream_t s);
void launch_ntt_inv(const DeviceParams &P, int base, const u64 *in, u64 *out, size_t count, cudaStream_t s);

void launch_poly_op(const Modulus &m, PolyOp op, const u64 *a, const u64 *b, u64 scalar, u64 *out,
                    size_t words, cudaStream_t s);

// Fused ciphertext-multiplication pipeline.
//   ct1, ct2 : [pairs][d][2][n]  NTT domain, canonical mod q
//   ext      : [pairs][2 sides][d][2][1+A][n] u64, or [pairs][2 sides][d][2][K][n] u32 with the internal basis  (workspace)
//   r01      : [pairs][products][2][n] u64      (workspace, coefficient domain)
//   digits   : [pairs][products][G][n] int8|int16|int32 (workspace; digit_kind 2 / 0 / 1: base <= 2^8 / 2^16 / 2^32)
//   rlk_mont : [G][2][n] relin key in Montgomery form
//   out      : [pairs][d][2][n];  excess: [pairs][num_limbs - d'][2][n] for k >= d
void launch_lift(const DeviceParams &P, const MulPlan &M, const u64 *ct1, const u64 *ct2, u64 *ext,
                 size_t pairs, cudaStream_t s);
// `mid` (optional) is recorded after the first of the two tensor kernels (per-limb components 0/1), or after
// the only one.
void launch_tensor(const DeviceParams &P, const MulPlan &M, const u64 *ct1, const u64 *ct2, const u64 *ext, u64 *r01,
                   void *digits, int digit_kind, size_t pairs, cudaStream_t s, cudaEvent_t mid = nullptr,
                   bool raw3 = false);   // raw3 (bfv_mul_no_relin): r01 = [pairs][products][3][n], all scaled components
bool tensor01_needs_r64(const DeviceParams &P, const MulPlan &M);   // i64 instead of i32 rounding sums
bool tensor_sums_per_limb(const DeviceParams &P, const MulPlan &M, size_t pairs);   // r01 is [pairs][limbs][2][n] when true
// `wide_scratch` (relin_wide_scratch_bytes, optional): lets small batches use one CTA per transform.
bool relin_goes_wide(const DeviceParams &P, const MulPlan &M, size_t pairs);
size_t relin_wide_scratch_bytes(const DeviceParams &P, const MulPlan &M, size_t pairs);
void launch_relin(const DeviceParams &P, const MulPlan &M, const u64 *r01, const void *digits,
                  int digit_kind, const u64 *rlk_mont, u64 *out, u64 *excess, size_t pairs,
                  cudaStream_t s, u64 *wide_scratch = nullptr, bool r01_ntt = false);
// gadget_decompose of coefficient polynomials [count][n] -> [count][G][n]; out_kind 0: int16, 1: int32, 3: int8 signed
// digits (relin layout), 2: u64 digits mod q (the reference's return value)
void launch_gadget_digits(const DeviceParams &P, const u64 *coeffs, void *out, int out_kind, size_t count, cudaStream_t s);
// out[pairs][d][2][n] limb i += rep * excess limb (signed scalar, reduction.rs:34-52)
void launch_reduce_mac(const DeviceParams &P, u64 *out_limb, const u64 *excess_limb, u64 abs_scalar_mod_q,
                       bool negative, size_t out_stride, size_t excess_stride, size_t pairs,
                       cudaStream_t s);

// Galois automorphism + key switch on `count` degree-1 ciphertexts [count][2][n] (bfv/eval.rs:512-561);
// gk_mont [G][2][n] in Montgomery form, element odd.  `out` must not alias `ct`.
void launch_galois(const DeviceParams &P, const u64 *ct, const u64 *gk_mont, u32 element, u64 *out,
                   size_t count, cudaStream_t s);

// decrypt (bfv/encrypt.rs:111-178) of `count` ciphertexts [count][ncomp][n] with the secret key in the
// NTT domain -> plaintext coefficients mod p, [count][n].  Needs p < q.
void launch_decrypt(const DeviceParams &P, const u64 *ct, u32 ncomp, const u64 *sk_ntt, u64 *out, size_t count,
                    cudaStream_t s);

// Batched transform with an explicit plan (any prime below 2^62): the multi-prime path's ct and extended primes.
void launch_ntt_plan(const Modulus &mod, const Tw *tw, const TwHead &head, u32 logn, bool forward, const u64 *in,
                     u64 *out, size_t count, cudaStream_t s);

// ---- multi-prime ciphertext modulus (rns_kernels.cu) ----------------------------------------------------
// ct batches are [pairs][d][2][L][n] (BFV: d = 1), keys [G][2][L][n]; `ws` holds rns_workspace_words u64.
size_t rns_workspace_words(const RnsConsts &R, const MulPlan &M, size_t pairs);
size_t rns_relin_workspace_words(const RnsConsts &R, size_t pairs);
void launch_rns_to_mont(const RnsConsts &R, u64 *key, size_t polys, cudaStream_t s);
void launch_rns_mul(const RnsConsts &R, const RnsPlans &T, const MulPlan &M, const u64 *ct1, const u64 *ct2,
                    const u64 *rlk_mont, u32 G, u64 *ws, u64 *out, size_t pairs, int mode, cudaStream_t s);
void launch_rns_relinearize(const RnsConsts &R, const RnsPlans &T, const MulPlan &M, const u64 *ct3,
                            const u64 *rlk_mont, u32 G, u64 *ws, u64 *out, size_t pairs, cudaStream_t s);

void launch_rns_decrypt(const RnsConsts &R, const RnsPlans &T, const u64 *ct, u32 ncomp, const u64 *sk, u64 *ws,
                        u64 *out, size_t count, cudaStream_t s);

// Once per device: opt the kernels into their dynamic shared-memory sizes.
void launch_prepare(int device);
#endif

extern std::atomic<int> g_ntt_path;                      // 0: TMA transform kernel (default), 1: cp.async kernel

int num_sms();                                           // SM count of the current device (148 on B200)

// Count of kernels launched by this library (bench.py's gpu_launches).
extern std::atomic<unsigned long long> g_launch_count;

}  // namespace exb
