// host_setup.hpp -- pure host-side (no CUDA calls) construction of everything the
// kernels need: NTT plans, HPS constants, the bfv_mul_no_relin dispatch decision
// and the dBFV work plan.  Shared by api.cu (the product) and by the host
// emulator under tests/host_emul (which replays the kernels on CPU threads).
#pragma once
#include <string>
#include <vector>

#include "../../include/exacto_b200.h"
#include "hps.cuh"
#include "hps32.cuh"
#include "modarith.cuh"
#include "ntt_core.cuh"
#include "rns.cuh"

namespace exb {

constexpr int kMaxAux = 2;
constexpr int kMaxBases = 1 + kMaxAux;
constexpr int kMaxDigits = 16;                 // dBFV digits d
constexpr int kMaxProducts = kMaxDigits * kMaxDigits;
constexpr int kMaxLimbs = 2 * kMaxDigits - 1;
constexpr int kMaxPeers = 7;                   // other GPUs of one 8-GPU box (k-sharded dbfv_mul)

// Internal 27-bit auxiliary basis (see ntt32_core.cuh / hps32.cuh): replaces the reference's aux
// primes inside the lift / tensor kernels when that is provably result-identical.
struct SmallBasis {
    u32 enabled, K;
    u32 max_terms;                // most products a limb may sum in tensor01_kernel (0: per-product kernel only)
    u32 max_terms_r32;            // ... with the rounding-term sums still fitting an i32
    const Tw32 *twf[kMaxSmall];
    const Tw32 *twi[kMaxSmall];
    TwHead32 headf[kMaxSmall], headi[kMaxSmall];
    Scale32Consts sc;
    Modulus mq_r;                 // mod[0] with n^-1 * 2^64 in the ninv fields (inverse transforms after a plain REDC)
};

// Everything a fused kernel needs about the parameter set (passed by value).
struct DeviceParams {
    u32 n, logn;
    u32 num_aux;                 // A
    u32 gadget_digits;           // G (number of relin keys actually used)
    u64 gadget_base;             // B
    u32 gadget_log2;             // w if B == 2^w, else 0
    u32 pipelined;               // host hint: this call's kernels share the GPU with other chunks' (multi-stream pipeline)
    u32 tensor_per_product;      // option: never use tensor01_kernel (per-product tensor kernel only)
    u32 relin_narrow;            // option: never use relin12_wide_kernel
    Modulus mod[kMaxBases];      // [0] = q, [1..A] = aux primes
    const Tw *twf[kMaxBases];    // forward twiddles (psi_rev) per base
    const Tw *twi[kMaxBases];    // inverse twiddles (psi_inv_rev) per base
    TwHead headf[kMaxBases];     // twf[0..15] / twi[0..15] by value (constant-bank operands)
    TwHead headi[kMaxBases];
    ScaleConsts sc;
    SmallBasis sb;
};

// Work decomposition of one dbfv_mul (bfv_mul_and_relin is d = 1).
struct MulPlan {
    u32 d;                                   // limbs per ciphertext
    u32 num_products;                        // live products per pair
    u32 num_limbs;                           // output limbs k that are computed
    u32 num_low;                             // computed limbs with k < d (they come first)
    uint8_t prod_i[kMaxProducts];            // product -> lhs limb i
    uint8_t prod_j[kMaxProducts];            // product -> rhs limb j
    uint8_t limb_k[kMaxLimbs];               // computed limb index -> k
    uint8_t pad2_;
    // tensor01_kernel work items: two computed limbs per CTA (heaviest with lightest, 0xFF = none) so that
    // every CTA runs about the same number of products
    uint8_t duo_a[kMaxLimbs], duo_b[kMaxLimbs];
    uint8_t pad3_[2];
    u32 num_duos;
    int16_t prod_of[kMaxDigits][kMaxDigits]; // (i, j) -> product index or -1
    // Input limbs that feed a live product (bit i): the lift skips the others (k-sharded ranks need few of them).
    u32 need_lhs, need_rhs;
    // k-sharded dbfv_mul: every finished output limb k < d is also stored into the peers' output buffers
    // (same [pair][d][2][n] layout) over NVLink peer memory, so no separate gather pass is needed.
    u32 num_peers, pad4_;
    u64 *peer_out[kMaxPeers];
};

// Device tables of the multi-prime path: plans of the ciphertext primes q_l and the extended primes e_k.
struct RnsPlans {
    const Tw *twf_q[kRnsMaxL], *twi_q[kRnsMaxL];
    const Tw *twf_e[kRnsMaxK], *twi_e[kRnsMaxK];
    TwHead headf_q[kRnsMaxL], headi_q[kRnsMaxL], headf_e[kRnsMaxK], headi_e[kRnsMaxK];
};

struct HostSetup {
    u32 n = 0, logn = 0;
    std::vector<u64> ct_moduli, aux_moduli;
    u32 user_aux = 0;                        // aux primes supplied by the caller (0: aux_moduli may hold the internal pair)
    bool internal_aux = false;               // aux_moduli were synthesised (reference would run its schoolbook branch)
    std::vector<u64> psi;                    // per modulus index (0 = q_0, 1..A = aux, then q_1..)
    u64 plain = 0, gadget_base = 0;
    u32 gadget_digits = 0;
    bool digits32 = false;                   // gadget digits need 32-bit storage
    // storage of the balanced digits in [-B/2, B/2): 2 = int8 (B <= 2^8), 0 = int16 (B <= 2^16), 1 = int32
    int digit_kind() const { return digits32 ? 1 : (gadget_base <= 256 ? 2 : 0); }
    size_t digit_bytes() const { return digits32 ? 4 : (gadget_base <= 256 ? 1 : 2); }
    int mul_status = EXB_OK;                 // bfv_mul_no_relin dispatch result for these params
    std::string mul_error;
    DeviceParams P;                          // table pointers are filled by the owner
    bool has_plan[kMaxBases] = {false, false, false};
    std::vector<Tw> twf[kMaxBases], twi[kMaxBases];
    std::vector<Tw32> twf32[kMaxSmall], twi32[kMaxSmall];   // internal small basis (if P.sb.enabled)
    std::vector<u64> small_primes;
    // multi-prime ciphertext modulus (ct_moduli.size() > 1): constants and host copies of the plans
    bool rns_enabled = false;
    RnsConsts R;
    RnsPlans T;                              // table pointers are filled by the owner
    std::vector<Tw> rns_twf_q[kRnsMaxL], rns_twi_q[kRnsMaxL], rns_twf_e[kRnsMaxK], rns_twi_e[kRnsMaxK];
    std::vector<u64> ext_primes;
};

// BfvParamsBuilder::build (params/mod.rs:81-124) + RnsBasis::new (ring/rns.rs:35-63).
int host_setup_build(const exb_bfv_params *p, HostSetup *hs, std::string *err, uint32_t flags = 0);

struct HostPlan {
    MulPlan M;
    u32 num_low = 0;                         // computed limbs with k < d
    std::vector<int64_t> reps;               // [(d-1)][d] small representatives
    std::vector<int> excess_index;           // k -> index in the excess buffer or -1
};

// SmallReps::compute_simple (dbfv/lattice.rs:104-122).
int host_small_reps(u64 base, u32 d, u64 plain_modulus, int64_t *reps, std::string *err);
// Work items of dbfv_mul (dbfv/eval.rs:109-114) minus products whose limb reduce discards.
int host_build_plan(u32 d, u64 base, u64 plain_modulus, u32 flags, u32 limb_mask, HostPlan *hp,
                    std::string *err);

}  // namespace exb
