/*
 * exacto_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (RajeshRk18/exacto, Rust) CPU
 * algorithm for the ciphertext-multiplication hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (exacto_b200/) never links or calls it.
 *
 * Parity pinning: the Rust reference cannot be compiled in this image (no
 * cargo/rustc) and it ships no ciphertext-level golden vectors.  The oracle is
 * pinned against every known-answer test the reference holds for this path
 * (gadget KATs bfv/keyswitch.rs:109-152, negacyclic sign ring/poly.rs:195-202,
 * NTT-mul == schoolbook ring/ntt.rs:181-195, decrypt KATs bfv/eval.rs:883-900
 * and dbfv/eval.rs:224-290,345-382,521-564, error pins dbfv/eval.rs:292-313,
 * 385-453) and against an independent Python big-int definition oracle
 * (oracle/definition.py).  NTT-domain word order of the third-party crate
 * concrete-ntt 0.2.0 (absent from /root/reference) is "parity unpinned": the
 * reference's own tests never pin it; every public result is compared in the
 * coefficient domain where negacyclic arithmetic mod q is unique.
 *
 * All file:line citations are into /root/reference/src/.
 */
#ifndef EXACTO_ORACLE_H
#define EXACTO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: 1:1 with ExactoError (error.rs:4-31). */
enum {
    EXO_OK = 0,
    EXO_INVALID_PARAM = 1,
    EXO_DIMENSION_MISMATCH = 2,
    EXO_MODULUS_MISMATCH = 3,
    EXO_INVALID_RING_DEGREE = 4,
    EXO_DECRYPTION_ERROR = 5,
    EXO_DECOMPOSITION_ERROR = 6,
    EXO_LATTICE_ERROR = 7,
    EXO_MISSING_KEY = 8,
    EXO_NOT_IMPLEMENTED = 9,
};

#define EXO_MAX_AUX 4

/* BfvParams restricted to what the hot path reads (params/mod.rs:11-27). */
typedef struct {
    uint32_t n;                 /* ring_degree                                  */
    uint32_t num_ct;            /* ct_basis.moduli.len() (hot path needs 1)     */
    uint64_t q;                 /* ct_basis.moduli[0]                           */
    uint32_t num_aux;           /* aux_basis moduli count (0 = None)            */
    uint64_t aux[EXO_MAX_AUX];  /* aux_basis.moduli                             */
    uint64_t plain_modulus;     /* p                                            */
    uint64_t gadget_base;       /* B                                            */
    uint32_t gadget_digits;     /* G                                            */
} exo_params;

const char *exo_last_error(void);

/* ---- ring/modular.rs ---------------------------------------------------- */
uint64_t exo_mod_mul(uint64_t a, uint64_t b, uint64_t m);      /* :81  */
uint64_t exo_mod_add(uint64_t a, uint64_t b, uint64_t m);      /* :57  */
uint64_t exo_mod_sub(uint64_t a, uint64_t b, uint64_t m);      /* :65  */
uint64_t exo_mod_neg(uint64_t a, uint64_t m);                  /* :75  */
uint64_t exo_mod_pow(uint64_t b, uint64_t e, uint64_t m);      /* :87  */
int      exo_mod_inv(uint64_t a, uint64_t m, uint64_t *out);   /* :102 */
int      exo_is_prime(uint64_t m);

/* ---- ring/ntt.rs (semantics; eval order is ours: natural -> bit-reversed) */
/* psi = x^((q-1)/2n) for the first x = 2,3,.. with psi^n == q-1.            */
int  exo_find_psi(uint32_t n, uint64_t q, uint64_t *psi);
/* Tables: psi_rev[k] = psi^bitrev(k), psi_inv_rev[k] = psi^-bitrev(k).      */
int  exo_ntt_tables(uint32_t n, uint64_t q, uint64_t *psi_rev, uint64_t *psi_inv_rev,
                    uint64_t *n_inv);
int  exo_ntt_fwd(uint32_t n, uint64_t q, uint64_t *a);          /* :42-55 */
int  exo_ntt_inv(uint32_t n, uint64_t q, uint64_t *a);          /* :58-67 */
int  exo_ntt_fwd_batch(uint32_t n, uint64_t q, uint64_t *a, size_t count, int threads);
int  exo_ntt_inv_batch(uint32_t n, uint64_t q, uint64_t *a, size_t count, int threads);
/* ring/poly.rs:85 mul_naive (test oracle for the NTT).                      */
void exo_poly_mul_naive(uint32_t n, uint64_t q, const uint64_t *a, const uint64_t *b, uint64_t *out);

/* ---- bfv/keyswitch.rs ---------------------------------------------------- */
/* gadget_decompose :11-52.  out is [num_digits][n].                         */
void exo_gadget_decompose(uint32_t n, uint64_t q, const uint64_t *coeffs, uint64_t base,
                          uint32_t num_digits, uint64_t *out);

/* relinearize :59-101 on one degree-2 ciphertext c3 [3][n] (NTT domain), rlk [G][2][n] -> out [2][n]. */
int exo_relinearize(const exo_params *p, const uint64_t *c3, const uint64_t *rlk, uint64_t *out);

/* ---- bfv/eval.rs --------------------------------------------------------- */
/* bfv_add :14-31 on two degree-1 ciphertexts [2][n].                        */
void exo_bfv_add(const exo_params *p, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* bfv_mul_no_relin :89-108 -> out [3][n] (NTT domain).                      */
int exo_bfv_mul_no_relin(const exo_params *p, const uint64_t *ct1, const uint64_t *ct2,
                         uint64_t *out3);
/* bfv_mul_and_relin :73-82.  ct [2][n], rlk [G][2][n], out [2][n].          */
int exo_bfv_mul_and_relin(const exo_params *p, const uint64_t *ct1, const uint64_t *ct2,
                          const uint64_t *rlk, uint64_t *out);
int exo_bfv_mul_and_relin_batch(const exo_params *p, const uint64_t *ct1, const uint64_t *ct2,
                                const uint64_t *rlk, uint64_t *out, size_t batch, int threads);

/* ---- Galois automorphism + key switch ------------------------------------ */
/* apply_automorphism bfv/keygen.rs:218-239.                                 */
void exo_apply_automorphism(uint32_t n, uint64_t q, const uint64_t *in, uint64_t k, uint64_t *out);
/* bfv_apply_automorphism bfv/eval.rs:512-561. ct [2][n], gk [G][2][n], out [2][n]. */
int exo_bfv_apply_automorphism(const exo_params *p, const uint64_t *ct, const uint64_t *gk,
                               uint64_t k, uint64_t *out);
int exo_bfv_apply_automorphism_batch(const exo_params *p, const uint64_t *ct, const uint64_t *gk,
                                     uint64_t k, uint64_t *out, size_t batch, int threads);

/* ---- dbfv/ --------------------------------------------------------------- */
/* SmallReps::compute_simple lattice.rs:104-122. reps is [(d-1)][d] int64.   */
void exo_small_reps(uint64_t base, uint32_t d, uint64_t plain_modulus, int64_t *reps);
/* dbfv_mul eval.rs:82-149 (+ reduction.rs:15-60).  ct [d][2][n] -> out [d][2][n].
 * threads: OpenMP threads over the d*d products (rayon par_iter, eval.rs:117). */
int exo_dbfv_mul(const exo_params *p, uint64_t base, uint32_t d, uint64_t dbfv_plain_modulus,
                 const uint64_t *ct1, const uint64_t *ct2, const uint64_t *rlk, uint64_t *out,
                 int threads);

int exo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
