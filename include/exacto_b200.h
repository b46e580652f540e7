/*
 * exacto_b200.h -- C ABI of libexacto_b200.so, the B200 (sm_100a) implementation
 * of exacto's ciphertext-multiplication hot path.
 *
 * The reference (RajeshRk18/exacto, pure Rust, CPU only) has no FFI layer; its
 * boundary for this path is the `pub fn` surface cited beside each entry point
 * below (file:line under /root/reference/src/).  A Rust maintainer binds these
 * symbols with an `extern "C"` block (INTEGRATION.md shows it) and keeps the
 * reference's signatures on top.
 *
 * Conventions
 *   - Plain pointers and sizes only.  `*_dev` pointers are CUDA device pointers
 *     on the context's device, `*_host` pointers are host memory (pinned host
 *     memory makes the copies asynchronous).
 *     Device buffers passed to the multiplication / relinearisation entry points
 *     must be 32-byte aligned (the kernels use 256-bit accesses; cudaMalloc,
 *     exb_device_alloc and framework allocators are): EXB_INVALID_PARAM otherwise.
 *     Host buffers have no alignment requirement beyond that of uint64_t.
 *   - Polynomials are n 64-bit words of canonical residues in [0, modulus).
 *     Ciphertext polynomials are in the NTT (evaluation) domain like the
 *     reference's NttPoly (ring/ntt.rs:11-15).  The evaluation order is this
 *     library's own (natural -> bit-reversed, psi = x^((q-1)/2n) for the first
 *     x >= 2 whose psi has order 2n); concrete-ntt's order is unpinned by the
 *     reference.  Use exb_ntt_* to move between domains.
 *   - BFV ciphertext batch  : [batch][2][n]        (BfvCiphertext.c, bfv/mod.rs:19-24)
 *     dBFV ciphertext batch : [batch][d][2][n]     (DbfvCiphertext.limbs, dbfv/ciphertext.rs:10-22)
 *     relinearisation key   : [G][2][n]            (RelinKey.keys[g] = (rlk0, rlk1), bfv/keygen.rs:39-45)
 *   - Every function returns an exb_status (1:1 with ExactoError, error.rs:4-31);
 *     exb_last_error() gives the message of the calling thread's last failure.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls
 *     are stream-ordered and asynchronous unless they take host pointers.
 *   - A context owns a small pool of workspaces: device-resident calls on different
 *     streams (from one or several host threads) take different workspaces and overlap
 *     on the GPU; host-buffer calls are pipelined over the context's own streams.
 *     There is no CPU fallback: every entry point fails with EXB_NOT_IMPLEMENTED / a
 *     CUDA error rather than computing on the host.
 *   - NTT-domain words are a PRIVATE interchange format (see "Interchange format"
 *     below): data written by the upstream crate's concrete-ntt plan must cross the
 *     boundary in the coefficient domain.
 */
#ifndef EXACTO_B200_H
#define EXACTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum exb_status {
    EXB_OK = 0,
    EXB_INVALID_PARAM = 1,        /* ExactoError::InvalidParam        */
    EXB_DIMENSION_MISMATCH = 2,   /* ExactoError::DimensionMismatch   */
    EXB_MODULUS_MISMATCH = 3,     /* ExactoError::ModulusMismatch     */
    EXB_INVALID_RING_DEGREE = 4,  /* ExactoError::InvalidRingDegree   */
    EXB_DECRYPTION_ERROR = 5,     /* ExactoError::DecryptionError     */
    EXB_DECOMPOSITION_ERROR = 6,  /* ExactoError::DecompositionError  */
    EXB_LATTICE_ERROR = 7,        /* ExactoError::LatticeError        */
    EXB_MISSING_KEY = 8,          /* ExactoError::MissingKey          */
    EXB_NOT_IMPLEMENTED = 9,      /* ExactoError::NotImplemented      */
    EXB_CUDA_ERROR = 100          /* no reference counterpart         */
} exb_status;

/* BfvParams (params/mod.rs:11-27) restricted to what the hot path reads. */
typedef struct exb_bfv_params {
    uint32_t ring_degree;          /* n, power of two                               */
    uint32_t num_ct_moduli;        /* ct_basis.moduli.len()                         */
    const uint64_t *ct_moduli;
    uint32_t num_aux_moduli;       /* 0 = aux_basis None                            */
    const uint64_t *aux_moduli;
    uint64_t plain_modulus;        /* p                                             */
    uint64_t gadget_base;          /* 0 = auto 2^16 (params/mod.rs:102-108)         */
    uint32_t gadget_digits;        /* 0 = compute_gadget_digits (params/mod.rs:126) */
} exb_bfv_params;

/* Threading: the reference's functions are re-entrant (and dbfv_mul fans out on rayon, dbfv/eval.rs:117); here a
 * context may be shared by host threads.  Every device-resident entry point takes one of the context's workspace
 * slots while it enqueues; calls on different streams use different slots and run concurrently, and a slot that
 * is reused by another stream is ordered behind its previous work with a CUDA event (no host blocking, no stored
 * user stream is ever dereferenced).  Host-buffer entry points serialise their enqueue on a second lock and run
 * on the context's own streams.  exb_last_error() is thread-local. */
typedef struct exb_context exb_context;       /* BfvParams + plans + workspace on one GPU */
typedef struct exb_relin_key exb_relin_key;   /* device-resident RelinKey                 */

/* Flags of exb_dbfv_mul*. */
enum {
    /* Compute all d*d products like dbfv/eval.rs:109-122 even when the output
     * limbs they feed are discarded by reduce (dbfv/reduction.rs:28-52).  The
     * default skips those dead products; the result is bit-identical. */
    EXB_DBFV_ALL_PRODUCTS = 1u
};

/* Flags of exb_context_create_ex. */
enum {
    /* Keep the parameter set's own auxiliary primes inside the lift / tensor kernels instead of the internal
     * 27-bit basis (both are bit-identical to the reference; tests compare the two). */
    EXB_CTX_REFERENCE_AUX_BASIS = 1u
};

/* Handle of one asynchronous host-buffer call (exb_*_host_async); 0 = nothing to wait for. */
typedef uint64_t exb_ticket;

const char *exb_last_error(void);
const char *exb_version(void);

/* ---- context: BfvParamsBuilder::build (params/mod.rs:81-124) + RnsBasis::new
 * (ring/rns.rs:35-63) + make_plan (ring/ntt.rs:19-29) -------------------------- */
int exb_context_create(const exb_bfv_params *params, int device, exb_context **out);
int exb_context_create_ex(const exb_bfv_params *params, int device, uint32_t flags, exb_context **out);
void exb_context_destroy(exb_context *ctx);
/* Tuning knobs (the defaults are the measured optimum; tests use them to force every code path):
 *   "device_chunk_bytes"  workspace budget of one chunk of a device-resident call (default 4 GiB)
 *   "host_chunk_products" per-digit products per chunk of the host-buffer pipeline (default: 1024 for the
 *                         synchronous calls, 3996 for the asynchronous ones)
 *   "host_slots"          depth of that pipeline: 2..8 staging sets / streams in rotation (default 4)
 *   "tensor_per_product"  1 = never sum components 0/1 per output limb (tensor01_kernel off)
 *   "relin_narrow"        1 = never give each relinearisation transform its own CTA
 *   "kshard_kernel_stores" 1 = exb_dbfv_mul_scatter writes the peers from the relin kernel's epilogue instead of
 *                         using the copy engines (A/B; measured slower)
 *   "ntt_cp_async"        1 = batched n = 4096 transforms use the cp.async kernel instead of the TMA kernel
 *                         (process-wide; A/B measurements) */
int exb_context_set_option(exb_context *ctx, const char *name, int64_t value);
/* Effective values after defaults were applied. */
int exb_context_gadget(const exb_context *ctx, uint64_t *gadget_base, uint32_t *gadget_digits);
/* psi of modulus `modulus_index`: 0 = q_0, 1..A = the aux primes, A+1.. = the remaining ciphertext primes q_1..
 * (the same indices address exb_ntt_* and exb_poly_*). */
int exb_context_psi(const exb_context *ctx, uint32_t modulus_index, uint64_t *psi);
/* Number of kernels this library has launched in this process. */
unsigned long long exb_launch_count(void);

/* Per-kernel device timing of the ct-mul pipeline (CUDA events recorded on the launching stream
 * between the kernels; no extra synchronisation).  stage: 0 = lift, 1 = tensor+scale (components 0/1
 * per output limb when that kernel is used, otherwise all three components per product), 2 = tensor+scale
 * of component 2 per product (0 launches when stage 1 covered it), 3 = relin, 4 = reduce.
 * exb_profile_read synchronises the recorded events, adds their times (ms) and launch counts into the
 * arrays (EXB_PROFILE_STAGES = 5 entries each) and clears the record. */
#define EXB_PROFILE_STAGES 5
int exb_profile_enable(exb_context *ctx, int on);
int exb_profile_read(exb_context *ctx, double *stage_ms, unsigned long long *stage_launches);

/* ---- device memory plumbing (so a host without a CUDA runtime binding can drive it) */
int exb_device_alloc(exb_context *ctx, size_t bytes, void **dev_ptr);
int exb_device_free(exb_context *ctx, void *dev_ptr);
int exb_copy_to_device(exb_context *ctx, void *dst_dev, const void *src_host, size_t bytes, void *stream);
int exb_copy_to_host(exb_context *ctx, void *dst_host, const void *src_dev, size_t bytes, void *stream);
int exb_synchronize(exb_context *ctx, void *stream);

/* Page-locked host memory for the *_host entry points.  The reference owns plain Vec<u64> (bfv/mod.rs:19-24);
 * pageable memory works but the copies then run synchronously at a fraction of PCIe speed.  Either allocate
 * ciphertext storage here (exb_host_alloc / exb_host_free) or pin an existing allocation in place
 * (exb_host_register / exb_host_unregister, e.g. a long-lived Vec<u64>). */
int exb_host_alloc(exb_context *ctx, size_t bytes, void **host_ptr);
/* EXB_HOST_WRITE_COMBINED: for buffers the CPU only WRITES (inputs of the *_host calls): uploads skip the cache
 * snoop, CPU reads from such memory are very slow. */
enum { EXB_HOST_WRITE_COMBINED = 1u };
int exb_host_alloc_ex(exb_context *ctx, size_t bytes, uint32_t flags, void **host_ptr);
int exb_host_free(exb_context *ctx, void *host_ptr);
int exb_host_register(exb_context *ctx, void *host_ptr, size_t bytes);
int exb_host_unregister(exb_context *ctx, void *host_ptr);

/* ---- Interchange format.  NTT-domain polynomials are in THIS library's evaluation order: forward =
 * Cooley-Tukey, natural -> bit-reversed, psi = x^((q-1)/2n) for the first x >= 2 whose psi has order 2n
 * (exb_context_psi).  The upstream crate's NttPoly words come from concrete-ntt 0.2.0's own root and order,
 * which no reference test pins and which cannot be reproduced here: NTT-domain buffers written by the
 * upstream crate (ciphertexts, relin / Galois keys, secret keys) are NOT interchangeable with this library's
 * and nothing in the words themselves can reveal the difference.  This is a breaking format difference:
 * cross the boundary in the coefficient domain (unique, canonical residues) -- exb_ntt_forward[_host]
 * imports CoeffPoly data, exb_ntt_inverse[_host] exports it -- or keep every from_coeff_poly / to_coeff_poly
 * on this library (INTEGRATION.md replaces ring/ntt.rs wholesale, which does exactly that).
 * exb_ntt_format_id() identifies the format of the words a context produces: (version << 56) ^ a hash of
 * (n, q, psi, ordering); store it beside any serialised NTT-domain buffer and compare on load. */
#define EXB_NTT_FORMAT_VERSION 1u
int exb_ntt_format_id(const exb_context *ctx, uint32_t modulus_index, uint64_t *format_id);

/* ---- ring/: NttPoly::from_coeff_poly (ring/ntt.rs:42-55) and NttPoly::to_coeff_poly
 * (ring/ntt.rs:58-67), batched over `count` polynomials; in == out allowed. -------- */
int exb_ntt_forward(exb_context *ctx, uint32_t modulus_index, const uint64_t *in_dev, uint64_t *out_dev,
                    size_t count, void *stream);
int exb_ntt_inverse(exb_context *ctx, uint32_t modulus_index, const uint64_t *in_dev, uint64_t *out_dev,
                    size_t count, void *stream);
int exb_ntt_forward_host(exb_context *ctx, uint32_t modulus_index, const uint64_t *in_host,
                         uint64_t *out_host, size_t count);
int exb_ntt_inverse_host(exb_context *ctx, uint32_t modulus_index, const uint64_t *in_host,
                         uint64_t *out_host, size_t count);

/* Point-wise ops on `words` residues: NttPoly::add/sub/neg/mul/scalar_mul
 * (ring/ntt.rs:75-139), RnsPoly::* (ring/rns.rs:159-217), CoeffPoly::* (ring/poly.rs:40-136). */
int exb_poly_add(exb_context *ctx, uint32_t modulus_index, const uint64_t *a_dev, const uint64_t *b_dev,
                 uint64_t *out_dev, size_t words, void *stream);
int exb_poly_sub(exb_context *ctx, uint32_t modulus_index, const uint64_t *a_dev, const uint64_t *b_dev,
                 uint64_t *out_dev, size_t words, void *stream);
int exb_poly_neg(exb_context *ctx, uint32_t modulus_index, const uint64_t *a_dev, uint64_t *out_dev,
                 size_t words, void *stream);
int exb_poly_mul(exb_context *ctx, uint32_t modulus_index, const uint64_t *a_dev, const uint64_t *b_dev,
                 uint64_t *out_dev, size_t words, void *stream);
int exb_poly_scalar_mul(exb_context *ctx, uint32_t modulus_index, const uint64_t *a_dev, uint64_t scalar,
                        uint64_t *out_dev, size_t words, void *stream);

/* ---- RelinKey (bfv/keygen.rs:39-45).  `num_keys` entries of (rlk0, rlk1), NTT domain;
 * relinearize uses min(gadget_digits, num_keys) of them (bfv/keyswitch.rs:86-89). ---- */
int exb_relin_key_load(exb_context *ctx, const uint64_t *rlk_host, uint32_t num_keys, exb_relin_key **out);
int exb_relin_key_load_device(exb_context *ctx, const uint64_t *rlk_dev, uint32_t num_keys, void *stream,
                              exb_relin_key **out);
void exb_relin_key_destroy(exb_relin_key *key);

/* ---- bfv_mul_and_relin (bfv/eval.rs:73-82), batched over independent pairs.
 * Dispatch and errors follow bfv_mul_no_relin (bfv/eval.rs:89-108):
 *   single-aux P <= n*q/2      -> EXB_INVALID_PARAM  "single aux prime too small for HPS centering..."
 *   more than 2 aux primes     -> EXB_INVALID_PARAM  "HPS scaling supports 1 or 2 aux primes..."
 *   no aux basis, i128 overflow-> EXB_NOT_IMPLEMENTED "schoolbook BFV multiplication can overflow i128..."
 *   no aux basis, 2*n*(q/2)^2*p > i128::MAX (the reference's guard misses the middle tensor term and its
 *     convolution wraps)        -> EXB_NOT_IMPLEMENTED "...overflows i128 in its middle tensor term..."
 *   no aux basis otherwise: the exact schoolbook result, computed by the HPS pipeline on an internal
 *     auxiliary basis;
 *   more than one ciphertext prime (bfv_mul_generic_rns, bfv/eval.rs:113-147, the reference's BigInt branch):
 *     2..4 primes with Q = prod q_l < 2^126 run on the device -- the integer tensor is computed exactly in an
 *     internal extended RNS basis and scaled with one multiword division per coefficient, bit-identical to the
 *     BigInt result; relinearize reproduces RnsPoly::to_coeff_poly's u128 reconstruction INCLUDING its truncation to
 *     u64 (ring/rns.rs:133-150), so it is word-identical for Q >= 2^64 too.  Ciphertexts are then
 *     [batch][2][L][n] (dBFV: [batch][d][2][L][n]), keys [G][2][L][n], secret keys [L][n]: one row per prime, in
 *     ct_moduli order.  Larger Q / more primes -> EXB_NOT_IMPLEMENTED with the reason (beyond 2^127 the reference's
 *     own u128 products overflow).  Multi-prime dbfv_mul needs p = b^d (all-zero small representatives). */
int exb_bfv_mul_and_relin(exb_context *ctx, const uint64_t *ct1_dev, const uint64_t *ct2_dev,
                          const exb_relin_key *rlk, uint64_t *out_dev, size_t batch, void *stream);
int exb_bfv_mul_and_relin_host(exb_context *ctx, const uint64_t *ct1_host, const uint64_t *ct2_host,
                               const exb_relin_key *rlk, uint64_t *out_host, size_t batch);
/* The two halves of bfv_mul_and_relin as the reference exposes them, device-resident:
 *   bfv_mul_no_relin (bfv/eval.rs:89-108): ct [batch][2][n] x2 -> degree-2 ciphertexts [batch][3][n] (NTT domain);
 *   relinearize (bfv/keyswitch.rs:59-101): ct [batch][num_components][n] -> [batch][2][n]; fewer than 3
 *     components are returned unchanged ([batch][num_components][n], :63-65), more than 3 ->
 *     EXB_INVALID_PARAM "relinearization only supports degree-2 ciphertexts" (:66-70);
 *   gadget_decompose (bfv/keyswitch.rs:11-52): coefficient-domain polynomials [count][n] -> balanced base-B digits,
 *     each stored mod q, [count][gadget_digits][n] (gadget base and digit count of the context). */
int exb_bfv_mul_no_relin(exb_context *ctx, const uint64_t *ct1_dev, const uint64_t *ct2_dev, uint64_t *out3_dev,
                         size_t batch, void *stream);
int exb_bfv_relinearize(exb_context *ctx, const uint64_t *ct_dev, uint32_t num_components, const exb_relin_key *rlk,
                        uint64_t *out_dev, size_t batch, void *stream);
int exb_gadget_decompose(exb_context *ctx, const uint64_t *coeffs_dev, uint64_t *out_dev, size_t count, void *stream);
/* bfv_add (bfv/eval.rs:14-31) on degree-1 ciphertexts. */
int exb_bfv_add(exb_context *ctx, const uint64_t *a_dev, const uint64_t *b_dev, uint64_t *out_dev,
                size_t batch, void *stream);

/* ---- bfv_apply_automorphism (bfv/eval.rs:512-561): sigma_k: X -> X^k (bfv/keygen.rs:218-239) on
 * both components followed by the key switch from s(X^k) back to s, batched over independent
 * degree-1 ciphertexts [batch][2][n].  A GaloisKey (bfv/keygen.rs:47-54, keys[g] = (ks0_g, a_g),
 * [G][2][n] NTT domain) is loaded with exb_relin_key_load: it is the same key-switch key layout.
 * `element` must be odd (EXB_INVALID_PARAM otherwise; the "degree-1 ciphertext" guard :516-520
 * reads ciphertext metadata and lives in the host wrapper).  out must not alias ct. */
int exb_bfv_apply_automorphism(exb_context *ctx, const uint64_t *ct_dev, uint64_t element,
                               const exb_relin_key *gk, uint64_t *out_dev, size_t batch, void *stream);
int exb_bfv_apply_automorphism_host(exb_context *ctx, const uint64_t *ct_host, uint64_t element,
                                    const exb_relin_key *gk, uint64_t *out_host, size_t batch);

/* ---- decrypt (bfv/encrypt.rs:111-178), batched: phase = c0 + c1 s + c2 s^2 + ..., then
 * m = floor((p * INTT(phase) + floor(q/2)) / q) mod p per coefficient.  ct [batch][num_components][n]
 * (NTT domain), sk_ntt [n] = SecretKey.poly (NTT domain), out [batch][n] coefficients mod p.
 * A single ciphertext prime needs p < q.  Multi-prime sets (see exb_bfv_mul_and_relin) take ct
 * [batch][num_components][L][n] and sk_ntt [L][n] and run the reference's BigUint CRT scaling (:136-170) with
 * multiword arithmetic on the device. */
int exb_bfv_decrypt(exb_context *ctx, const uint64_t *ct_dev, uint32_t num_components, const uint64_t *sk_ntt_dev,
                    uint64_t *out_dev, size_t batch, void *stream);
int exb_bfv_decrypt_host(exb_context *ctx, const uint64_t *ct_host, uint32_t num_components,
                         const uint64_t *sk_ntt_host, uint64_t *out_host, size_t batch);

/* ---- dbfv_mul (dbfv/eval.rs:82-149) incl. reduction::reduce (dbfv/reduction.rs:15-60),
 * batched.  `base`, `num_digits`, `dbfv_plain_modulus` are DbfvParams (params/mod.rs:143-192;
 * 0 = 2^64).  The limb-count and mul_depth guards (dbfv/eval.rs:90-102) read ciphertext
 * metadata and therefore live in the host wrapper that owns it.
 * `limb_mask`: bit k set = compute output limb k (0 = all d limbs); lets ranks of a
 * multi-GPU job own disjoint limbs before an all-gather.  Limbs not selected are left
 * untouched in out. */
int exb_dbfv_mul(exb_context *ctx, uint64_t base, uint32_t num_digits, uint64_t dbfv_plain_modulus,
                 const uint64_t *ct1_dev, const uint64_t *ct2_dev, const exb_relin_key *rlk,
                 uint64_t *out_dev, size_t batch, uint32_t flags, uint32_t limb_mask, void *stream);
int exb_dbfv_mul_host(exb_context *ctx, uint64_t base, uint32_t num_digits, uint64_t dbfv_plain_modulus,
                      const uint64_t *ct1_host, const uint64_t *ct2_host, const exb_relin_key *rlk,
                      uint64_t *out_host, size_t batch, uint32_t flags);
/* ---- k-sharded dbfv_mul across the GPUs of one box (one process per GPU).  Every rank holds the same
 * ciphertext pairs and owns the output limbs of its `limb_mask` (products with equal i + j stay on one rank, so
 * the per-k accumulation of dbfv/eval.rs:125-136 is local).  The finished limbs go into this rank's `out_dev` and
 * into `peer_outs_dev[0..num_peers)` -- the other ranks' output buffers, same [batch][d][2][n] layout, mapped over
 * NVLink with exb_ipc_open: by default with one strided peer DMA per (peer, run of owned limbs) on per-peer
 * streams that the call's stream joins, or (option "kshard_kernel_stores") with stores from the relinearisation
 * kernel's epilogue.  Either way the path's only exchange step moves exactly 512 KiB * (N-1)/N per dbfv_mul into
 * each rank.  The caller synchronises the ranks (any barrier after the stream work) before reading.  Needs p = b^d (all-zero small
 * representatives, every BASELINE config): EXB_NOT_IMPLEMENTED otherwise. */
int exb_dbfv_mul_scatter(exb_context *ctx, uint64_t base, uint32_t num_digits, uint64_t dbfv_plain_modulus,
                         const uint64_t *ct1_dev, const uint64_t *ct2_dev, const exb_relin_key *rlk,
                         uint64_t *out_dev, uint64_t *const *peer_outs_dev, uint32_t num_peers, size_t batch,
                         uint32_t flags, uint32_t limb_mask, void *stream);
/* CUDA IPC for those buffers: `dev_ptr` must be the base of an exb_device_alloc allocation. */
int exb_ipc_export(exb_context *ctx, void *dev_ptr, uint8_t handle[64]);
int exb_ipc_open(exb_context *ctx, const uint8_t handle[64], void **peer_dev_ptr);
int exb_ipc_close(exb_context *ctx, void *peer_dev_ptr);

/* Asynchronous forms of the host-buffer calls: the work is enqueued on the context's pipeline and the call
 * returns; exb_wait(ticket) blocks until the output has landed in out_host.  Inputs must stay valid and
 * unmodified, and out_host unread, until then.  Calls are pipelined across each other: the next call's
 * uploads overlap this call's kernels and downloads, so a stream of calls runs at max(PCIe, kernel) speed
 * instead of their sum.  Host memory should be page-locked (exb_host_alloc / exb_host_register).
 * The synchronous forms above are async + wait.  The reference has no counterpart (its calls are CPU-synchronous,
 * bfv/eval.rs:73-82, dbfv/eval.rs:82-149); ownership is unchanged: inputs borrowed, output caller-owned. */
int exb_dbfv_mul_host_async(exb_context *ctx, uint64_t base, uint32_t num_digits, uint64_t dbfv_plain_modulus,
                            const uint64_t *ct1_host, const uint64_t *ct2_host, const exb_relin_key *rlk,
                            uint64_t *out_host, size_t batch, uint32_t flags, exb_ticket *ticket);
int exb_bfv_mul_and_relin_host_async(exb_context *ctx, const uint64_t *ct1_host, const uint64_t *ct2_host,
                                     const exb_relin_key *rlk, uint64_t *out_host, size_t batch, exb_ticket *ticket);
int exb_wait(exb_context *ctx, exb_ticket ticket);
/* SmallReps::compute_simple (dbfv/lattice.rs:104-122): reps[(d-1)][d]. */
int exb_dbfv_small_reps(uint64_t base, uint32_t num_digits, uint64_t dbfv_plain_modulus, int64_t *reps);

#ifdef __cplusplus
}
#endif
#endif /* EXACTO_B200_H */
