/*
 * exacto_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See exacto_oracle.h for scope, pinning status and who may load this.
 *
 * Literal restatement of the reference's CPU schedule: every product does
 * 4 centered base extensions, Q- and P-basis tensors, 3 x hps_scale, one INTT,
 * balanced gadget decomposition and G NTT + 2G multiply-accumulates, with the
 * same `u128 %` modular reduction (ring/modular.rs:8-11).  The NTT is a scalar
 * one in place of concrete-ntt's SIMD transform (crate absent from
 * /root/reference; its eval order is unpinned by the reference's own tests).
 *
 * Citations are into /root/reference/src/.
 */
#include "exacto_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef __int128 i128;
typedef uint64_t u64;

static __thread char g_err[256];

const char *exo_last_error(void) { return g_err; }

static int fail(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

int exo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* ring/modular.rs                                                           */
/* ------------------------------------------------------------------------ */

/* mod_mul :81 -> barrett_reduce :7-19.  For m > 2^32 the reference takes the
 * exact `a % m` branch; for m <= 2^32 its one-word Barrett is exact for
 * a < m^2, so `%` restates both branches. */
u64 exo_mod_mul(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }

/* mod_add :57-61 */
u64 exo_mod_add(u64 a, u64 b, u64 m) {
    u128 s = (u128)a + b;
    return s >= m ? (u64)(s - m) : (u64)s;
}

/* mod_sub :65-71 */
u64 exo_mod_sub(u64 a, u64 b, u64 m) { return a >= b ? a - b : m - b + a; }

/* mod_neg :75-77 */
u64 exo_mod_neg(u64 a, u64 m) { return a == 0 ? 0 : m - a; }

/* mod_pow :87-99 */
u64 exo_mod_pow(u64 base, u64 e, u64 m) {
    u64 r = 1;
    base %= m;
    while (e > 0) {
        if (e & 1) r = exo_mod_mul(r, base, m);
        e >>= 1;
        base = exo_mod_mul(base, base, m);
    }
    return r;
}

/* mod_inv :102-121 (extended Euclid on i128, Euclidean final reduction) */
int exo_mod_inv(u64 a, u64 m, u64 *out) {
    i128 old_r = a, r = m, old_s = 1, s = 0;
    while (r != 0) {
        i128 qq = old_r / r;
        i128 t = r;
        r = old_r - qq * r;
        old_r = t;
        t = s;
        s = old_s - qq * s;
        old_s = t;
    }
    if (old_r != 1) return 0;
    *out = (u64)(((old_s % (i128)m) + (i128)m) % (i128)m);
    return 1;
}

int exo_is_prime(u64 m) {
    if (m < 2) return 0;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (size_t i = 0; i < sizeof small / sizeof *small; i++) {
        if (m == small[i]) return 1;
        if (m % small[i] == 0) return 0;
    }
    u64 d = m - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (size_t i = 0; i < sizeof small / sizeof *small; i++) {  /* deterministic for 64-bit */
        u64 x = exo_mod_pow(small[i], d, m);
        if (x == 1 || x == m - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) {
            x = exo_mod_mul(x, x, m);
            if (x == m - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------ */
/* ring/ntt.rs -- plans                                                      */
/* ------------------------------------------------------------------------ */

static unsigned ilog2(uint32_t n) { unsigned l = 0; while ((1u << l) < n) l++; return l; }

static uint32_t bitrev(uint32_t x, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* make_plan ring/ntt.rs:19-29: n power of two, q prime, q == 1 mod 2n. */
int exo_find_psi(uint32_t n, u64 q, u64 *psi) {
    if (n < 2 || (n & (n - 1))) return fail(EXO_INVALID_RING_DEGREE, "ring degree must be a power of 2");
    if (!exo_is_prime(q) || (q - 1) % (2ull * n) != 0)
        return fail(EXO_INVALID_PARAM, "cannot create NTT plan (need prime q = 1 mod 2n)");
    u64 e = (q - 1) / (2ull * n);
    for (u64 x = 2; x < q; x++) {
        u64 c = exo_mod_pow(x, e, q);
        if (exo_mod_pow(c, n, q) == q - 1) { *psi = c; return EXO_OK; }
    }
    return fail(EXO_INVALID_PARAM, "no primitive 2n-th root found");
}

int exo_ntt_tables(uint32_t n, u64 q, u64 *psi_rev, u64 *psi_inv_rev, u64 *n_inv) {
    u64 psi, psi_inv, ninv;
    int rc = exo_find_psi(n, q, &psi);
    if (rc) return rc;
    if (!exo_mod_inv(psi, q, &psi_inv) || !exo_mod_inv(n % q, q, &ninv))
        return fail(EXO_INVALID_PARAM, "psi or n not invertible");
    unsigned bits = ilog2(n);
    u64 pw = 1, ipw = 1;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t r = bitrev(k, bits);
        psi_rev[r] = pw;
        psi_inv_rev[r] = ipw;
        pw = exo_mod_mul(pw, psi, q);
        ipw = exo_mod_mul(ipw, psi_inv, q);
    }
    *n_inv = ninv;
    return EXO_OK;
}

typedef struct plan {
    uint32_t n;
    u64 q, n_inv;
    u64 *psi_rev, *psi_inv_rev;
    struct plan *next;
} plan;

static plan *g_plans = NULL;

static const plan *get_plan(uint32_t n, u64 q) {
    const plan *found = NULL;
#pragma omp critical(exo_plan_cache)
    {
        for (plan *p = g_plans; p; p = p->next)
            if (p->n == n && p->q == q) { found = p; break; }
        if (!found) {
            plan *p = (plan *)calloc(1, sizeof *p);
            p->n = n; p->q = q;
            p->psi_rev = (u64 *)malloc(sizeof(u64) * n);
            p->psi_inv_rev = (u64 *)malloc(sizeof(u64) * n);
            if (exo_ntt_tables(n, q, p->psi_rev, p->psi_inv_rev, &p->n_inv) == EXO_OK) {
                p->next = g_plans; g_plans = p; found = p;
            } else {
                free(p->psi_rev); free(p->psi_inv_rev); free(p);
            }
        }
    }
    return found;
}

/* Forward negacyclic NTT, natural -> bit-reversed (NttPoly::from_coeff_poly
 * ring/ntt.rs:42-55 -> plan.fwd). */
static void ntt_fwd(const plan *pl, u64 *a) {
    const u64 q = pl->q;
    uint32_t t = pl->n;
    for (uint32_t m = 1; m < pl->n; m <<= 1) {
        t >>= 1;
        for (uint32_t i = 0; i < m; i++) {
            const u64 w = pl->psi_rev[m + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (uint32_t j = 0; j < t; j++) {
                u64 u = x[j], v = exo_mod_mul(y[j], w, q);
                x[j] = exo_mod_add(u, v, q);
                y[j] = exo_mod_sub(u, v, q);
            }
        }
    }
}

/* Inverse NTT, bit-reversed -> natural, then x n^-1 (NttPoly::to_coeff_poly
 * ring/ntt.rs:58-67 -> plan.inv + plan.normalize). */
static void ntt_inv(const plan *pl, u64 *a) {
    const u64 q = pl->q;
    uint32_t t = 1;
    for (uint32_t m = pl->n; m > 1; m >>= 1) {
        uint32_t h = m >> 1;
        for (uint32_t i = 0; i < h; i++) {
            const u64 w = pl->psi_inv_rev[h + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (uint32_t j = 0; j < t; j++) {
                u64 u = x[j], v = y[j];
                x[j] = exo_mod_add(u, v, q);
                y[j] = exo_mod_mul(exo_mod_sub(u, v, q), w, q);
            }
        }
        t <<= 1;
    }
    for (uint32_t j = 0; j < pl->n; j++) a[j] = exo_mod_mul(a[j], pl->n_inv, q);
}

int exo_ntt_fwd(uint32_t n, u64 q, u64 *a) {
    const plan *pl = get_plan(n, q);
    if (!pl) return EXO_INVALID_PARAM;
    ntt_fwd(pl, a);
    return EXO_OK;
}

int exo_ntt_inv(uint32_t n, u64 q, u64 *a) {
    const plan *pl = get_plan(n, q);
    if (!pl) return EXO_INVALID_PARAM;
    ntt_inv(pl, a);
    return EXO_OK;
}

int exo_ntt_fwd_batch(uint32_t n, u64 q, u64 *a, size_t count, int threads) {
    const plan *pl = get_plan(n, q);
    if (!pl) return EXO_INVALID_PARAM;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < count; i++) ntt_fwd(pl, a + i * n);
    return EXO_OK;
}

int exo_ntt_inv_batch(uint32_t n, u64 q, u64 *a, size_t count, int threads) {
    const plan *pl = get_plan(n, q);
    if (!pl) return EXO_INVALID_PARAM;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < count; i++) ntt_inv(pl, a + i * n);
    return EXO_OK;
}

/* CoeffPoly::mul_naive ring/poly.rs:85-121 */
void exo_poly_mul_naive(uint32_t n, u64 q, const u64 *a, const u64 *b, u64 *out) {
    memset(out, 0, sizeof(u64) * n);
    for (uint32_t i = 0; i < n; i++) {
        if (a[i] == 0) continue;
        for (uint32_t j = 0; j < n; j++) {
            if (b[j] == 0) continue;
            u64 prod = exo_mod_mul(a[i], b[j], q);
            uint32_t idx = i + j;
            if (idx < n) out[idx] = exo_mod_add(out[idx], prod, q);
            else out[idx - n] = exo_mod_sub(out[idx - n], prod, q);
        }
    }
}

/* NttPoly pointwise helpers ring/ntt.rs:75-129 (via RnsPoly ring/rns.rs:159-209) */
static void pw_mul(uint32_t n, u64 m, const u64 *a, const u64 *b, u64 *o) {
    for (uint32_t i = 0; i < n; i++) o[i] = exo_mod_mul(a[i], b[i], m);
}
static void pw_add(uint32_t n, u64 m, const u64 *a, const u64 *b, u64 *o) {
    for (uint32_t i = 0; i < n; i++) o[i] = exo_mod_add(a[i], b[i], m);
}

/* ------------------------------------------------------------------------ */
/* bfv/keyswitch.rs                                                          */
/* ------------------------------------------------------------------------ */

/* gadget_decompose :11-52 -- balanced digits of the centered coefficient,
 * i128 truncating % and /, each digit stored mod q. */
void exo_gadget_decompose(uint32_t n, u64 q, const u64 *coeffs, u64 base, uint32_t num_digits,
                          u64 *out) {
    const i128 base_i = (i128)base, half_base = base_i / 2, q_i = (i128)q;
    const u64 half_q = q / 2;
    for (uint32_t pos = 0; pos < n; pos++) {
        u64 c = coeffs[pos];
        i128 remaining = c > half_q ? (i128)c - q_i : (i128)c;
        for (uint32_t d = 0; d < num_digits; d++) {
            i128 rem = remaining % base_i;
            if (rem < -half_base) rem += base_i;
            else if (rem >= half_base) rem -= base_i;
            i128 rem_mod_q = ((rem % q_i) + q_i) % q_i;
            out[(size_t)d * n + pos] = (u64)rem_mod_q;
            remaining = (remaining - rem) / base_i;
        }
    }
}

/* relinearize :59-101 on a degree-2 ciphertext c3 = [3][n] (NTT domain). */
static int relinearize(const exo_params *p, const plan *pq, const u64 *c3, const u64 *rlk,
                       u64 *out) {
    const uint32_t n = p->n, G = p->gadget_digits;
    const u64 q = p->q;
    u64 *c2 = (u64 *)malloc(sizeof(u64) * n);
    u64 *digits = (u64 *)malloc(sizeof(u64) * n * G);
    u64 *prod = (u64 *)malloc(sizeof(u64) * n);
    memcpy(c2, c3 + 2 * (size_t)n, sizeof(u64) * n);
    ntt_inv(pq, c2);                                                 /* :76 */
    exo_gadget_decompose(n, q, c2, p->gadget_base, G, digits);       /* :79 */
    memcpy(out, c3, sizeof(u64) * 2 * n);                            /* :83-84 */
    for (uint32_t g = 0; g < G; g++) {                               /* :86-95 */
        u64 *dg = digits + (size_t)g * n;
        ntt_fwd(pq, dg);
        pw_mul(n, q, dg, rlk + ((size_t)g * 2 + 0) * n, prod);
        pw_add(n, q, out, prod, out);
        pw_mul(n, q, dg, rlk + ((size_t)g * 2 + 1) * n, prod);
        pw_add(n, q, out + n, prod, out + n);
    }
    free(c2); free(digits); free(prod);
    return EXO_OK;
}

/* relinearize :59-101 as a public entry: c3 [3][n] -> out [2][n]. */
int exo_relinearize(const exo_params *p, const u64 *c3, const u64 *rlk, u64 *out) {
    const plan *pq = get_plan(p->n, p->q);
    if (!pq) return fail(EXO_INVALID_PARAM, "cannot create NTT plan");
    return relinearize(p, pq, c3, rlk, out);
}

/* ------------------------------------------------------------------------ */
/* bfv/eval.rs                                                               */
/* ------------------------------------------------------------------------ */

/* bfv_add :14-31 (two degree-1 ciphertexts) */
void exo_bfv_add(const exo_params *p, const u64 *a, const u64 *b, u64 *out) {
    pw_add(2 * p->n, p->q, a, b, out);
}

/* The centered extension rule shared by base_extend_centered :230-240 and
 * hps_scale :307-314 / :357-375. */
static inline u64 ext_centered(u64 c, u64 q, u64 half_q, u64 pj) {
    if (c > half_q) {
        u64 rem = (q - c) % pj;
        return rem == 0 ? 0 : pj - rem;
    }
    return c % pj;
}

/* base_extend_centered :217-247: INTT_q, center/reduce mod each p_j, NTT_pj.
 * out is [A][n]. */
static void base_extend_centered(const exo_params *p, const plan *pq, const plan *const *pp,
                                 const u64 *poly_q, u64 *out) {
    const uint32_t n = p->n;
    const u64 q = p->q, half_q = q / 2;
    u64 *cq = (u64 *)malloc(sizeof(u64) * n);
    memcpy(cq, poly_q, sizeof(u64) * n);
    ntt_inv(pq, cq);
    for (uint32_t j = 0; j < p->num_aux; j++) {
        u64 *o = out + (size_t)j * n;
        for (uint32_t i = 0; i < n; i++) o[i] = ext_centered(cq[i], q, half_q, p->aux[j]);
        ntt_fwd(pp[j], o);
    }
    free(cq);
}

/* hps_scale :257-413.  t_q [n], t_p [A][n] in the NTT domain -> out [n] NTT_q. */
static int hps_scale(const exo_params *p, const plan *pq, const plan *const *pp, const u64 *t_q,
                     const u64 *t_p, u64 *out) {
    const uint32_t n = p->n, A = p->num_aux;
    const u64 q = p->q, half_q = q / 2, pl = p->plain_modulus;
    const i128 p_128 = (i128)pl, q_128 = (i128)q;
    u64 *a_poly = (u64 *)malloc(sizeof(u64) * n);
    u64 *b_polys = (u64 *)malloc(sizeof(u64) * n * (A ? A : 1));
    memcpy(a_poly, t_q, sizeof(u64) * n);
    ntt_inv(pq, a_poly);                                             /* :267 */
    for (uint32_t j = 0; j < A; j++) {                               /* :270-272 */
        memcpy(b_polys + (size_t)j * n, t_p + (size_t)j * n, sizeof(u64) * n);
        ntt_inv(pp[j], b_polys + (size_t)j * n);
    }
    u64 q_inv_pj[EXO_MAX_AUX] = {0};
    for (uint32_t j = 0; j < A && j < EXO_MAX_AUX; j++)              /* :275-286 */
        if (!exo_mod_inv(q % p->aux[j], p->aux[j], &q_inv_pj[j])) {
            free(a_poly); free(b_polys);
            return fail(EXO_INVALID_PARAM, "q not invertible mod p_j");
        }

    if (A == 1) {                                                    /* :294-332 */
        const u64 big_p = p->aux[0], half_p = big_p / 2;
        for (uint32_t i = 0; i < n; i++) {
            u64 a = a_poly[i], b = b_polys[i];
            i128 a_centered = a > half_q ? (i128)a - q_128 : (i128)a;
            u64 a_ext = ext_centered(a, q, half_q, big_p);
            u64 diff = b >= a_ext ? b - a_ext : big_p - a_ext + b;
            u64 m_raw = exo_mod_mul(diff, q_inv_pj[0], big_p);
            i128 m_centered = m_raw > half_p ? (i128)m_raw - (i128)big_p : (i128)m_raw;
            i128 pa = p_128 * a_centered;
            i128 round_pa_q = pa >= 0 ? (pa + q_128 / 2) / q_128 : -((-pa + q_128 / 2) / q_128);
            i128 scaled = round_pa_q + p_128 * m_centered;
            a_poly[i] = (u64)(((scaled % q_128) + q_128) % q_128);
        }
    } else if (A == 2) {                                             /* :333-404 */
        const u64 p0 = p->aux[0], p1 = p->aux[1];
        u64 p1_inv_p0, p0_inv_p1;
        if (!exo_mod_inv(p1 % p0, p0, &p1_inv_p0) || !exo_mod_inv(p0 % p1, p1, &p0_inv_p1)) {
            free(a_poly); free(b_polys);
            return fail(EXO_INVALID_PARAM, "aux primes not coprime");
        }
        const i128 big_p_128 = (i128)p0 * (i128)p1, half_big_p = big_p_128 / 2;
        const u64 *b0p = b_polys, *b1p = b_polys + n;
        for (uint32_t i = 0; i < n; i++) {
            u64 a = a_poly[i];
            i128 a_centered = a > half_q ? (i128)a - q_128 : (i128)a;
            u64 b0 = b0p[i], a_ext0 = ext_centered(a, q, half_q, p0);
            u64 diff0 = b0 >= a_ext0 ? b0 - a_ext0 : p0 - a_ext0 + b0;
            u64 m0 = exo_mod_mul(diff0, q_inv_pj[0], p0);
            u64 b1 = b1p[i], a_ext1 = ext_centered(a, q, half_q, p1);
            u64 diff1 = b1 >= a_ext1 ? b1 - a_ext1 : p1 - a_ext1 + b1;
            u64 m1 = exo_mod_mul(diff1, q_inv_pj[1], p1);
            i128 t0 = (i128)exo_mod_mul(m0, p1_inv_p0, p0);
            i128 t1 = (i128)exo_mod_mul(m1, p0_inv_p1, p1);
            i128 crt_sum = t0 * (i128)p1 + t1 * (i128)p0;
            i128 m_crt = crt_sum % big_p_128;
            i128 m_centered = m_crt > half_big_p ? m_crt - big_p_128 : m_crt;
            i128 m_mod_q = ((m_centered % q_128) + q_128) % q_128;
            i128 pa = p_128 * a_centered;
            i128 round_pa_q = pa >= 0 ? (pa + q_128 / 2) / q_128 : -((-pa + q_128 / 2) / q_128);
            u64 round_mod_q = (u64)(((round_pa_q % q_128) + q_128) % q_128);
            u64 pm_mod_q = exo_mod_mul(pl, (u64)m_mod_q, q);
            a_poly[i] = (u64)(((u128)round_mod_q + pm_mod_q) % q);
        }
    } else {                                                         /* :405-409 */
        free(a_poly); free(b_polys);
        char msg[96];
        snprintf(msg, sizeof msg, "HPS scaling supports 1 or 2 aux primes, got %u", A);
        return fail(EXO_INVALID_PARAM, msg);
    }
    ntt_fwd(pq, a_poly);                                             /* :411-412 (c % q is a no-op) */
    memcpy(out, a_poly, sizeof(u64) * n);
    free(a_poly); free(b_polys);
    return EXO_OK;
}

/* bfv_mul_hps :157-209 */
static int bfv_mul_hps(const exo_params *p, const plan *pq, const plan *const *pp, const u64 *ct1,
                       const u64 *ct2, u64 *out3) {
    const uint32_t n = p->n, A = p->num_aux;
    const u64 q = p->q;
    if (A == 1) {                                                    /* :170-178 */
        u128 min_required = ((u128)n * q) / 2;
        if ((u128)p->aux[0] <= min_required) {
            char msg[200], dec[48];
            int pos = (int)sizeof dec - 1;
            dec[pos] = 0;
            u128 v = min_required;
            do { dec[--pos] = (char)('0' + (int)(v % 10)); v /= 10; } while (v);
            snprintf(msg, sizeof msg,
                     "single aux prime too small for HPS centering: P=%llu <= n*Q/2=%s",
                     (unsigned long long)p->aux[0], dec + pos);
            return fail(EXO_INVALID_PARAM, msg);
        }
    }
    const size_t An = (size_t)A * n;
    u64 *ext = (u64 *)malloc(sizeof(u64) * An * 4);
    u64 *c0p = ext, *c1p = ext + An, *d0p = ext + 2 * An, *d1p = ext + 3 * An;
    const u64 *c0 = ct1, *c1 = ct1 + n, *d0 = ct2, *d1 = ct2 + n;
    base_extend_centered(p, pq, pp, c0, c0p);                        /* :181-184 */
    base_extend_centered(p, pq, pp, c1, c1p);
    base_extend_centered(p, pq, pp, d0, d0p);
    base_extend_centered(p, pq, pp, d1, d1p);

    u64 *tq = (u64 *)malloc(sizeof(u64) * n * 4);                    /* t0,t1,t2,tmp */
    pw_mul(n, q, c0, d0, tq);                                        /* :187-191 */
    pw_mul(n, q, c0, d1, tq + n);
    pw_mul(n, q, c1, d0, tq + 3 * (size_t)n);
    pw_add(n, q, tq + n, tq + 3 * (size_t)n, tq + n);
    pw_mul(n, q, c1, d1, tq + 2 * (size_t)n);

    u64 *tp = (u64 *)malloc(sizeof(u64) * An * 4);                   /* [comp][A][n] + tmp */
    for (uint32_t j = 0; j < A; j++) {                               /* :194-198 */
        const u64 pj = p->aux[j];
        const size_t o = (size_t)j * n;
        pw_mul(n, pj, c0p + o, d0p + o, tp + 0 * An + o);
        pw_mul(n, pj, c0p + o, d1p + o, tp + 1 * An + o);
        pw_mul(n, pj, c1p + o, d0p + o, tp + 3 * An + o);
        pw_add(n, pj, tp + 1 * An + o, tp + 3 * An + o, tp + 1 * An + o);
        pw_mul(n, pj, c1p + o, d1p + o, tp + 2 * An + o);
    }
    int rc = EXO_OK;
    for (int comp = 0; comp < 3 && rc == EXO_OK; comp++)             /* :201-203 */
        rc = hps_scale(p, pq, pp, tq + (size_t)comp * n, tp + (size_t)comp * An,
                       out3 + (size_t)comp * n);
    free(ext); free(tq); free(tp);
    return rc;
}

/* schoolbook_overflow_risk :457-464 (saturating u128 products) */
static u128 sat_mul(u128 a, u128 b) {
    if (a == 0 || b == 0) return 0;
    if (a > (~(u128)0) / b) return ~(u128)0;
    return a * b;
}
static int schoolbook_overflow_risk(u64 p, u64 q, uint32_t n) {
    const u128 i128_max = (~(u128)0) >> 1;
    u128 max_coeff = q / 2;
    u128 max_tensor = sat_mul(sat_mul((u128)n, max_coeff), max_coeff);
    u128 max_scaled = sat_mul(max_tensor, (u128)p);
    return max_tensor > i128_max || max_scaled > i128_max;
}

/* bfv_mul_schoolbook :416-454 with helpers :655-711 (small parameters only). */
static void poly_mul_i128(const i128 *a, const i128 *b, uint32_t n, i128 *r) {
    for (uint32_t i = 0; i < n; i++) r[i] = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (a[i] == 0) continue;
        for (uint32_t j = 0; j < n; j++) {
            if (b[j] == 0) continue;
            i128 prod = a[i] * b[j];
            uint32_t idx = i + j;
            if (idx < n) r[idx] += prod; else r[idx - n] -= prod;
        }
    }
}
static void scale_tensor_component(const i128 *x, u64 p, u64 q, uint32_t n, u64 *out) {
    const i128 pi = (i128)p, qi = (i128)q;
    for (uint32_t i = 0; i < n; i++) {
        i128 num = pi * x[i];
        i128 rounded = num >= 0 ? (num + qi / 2) / qi : (-(-num + qi / 2)) / qi;
        out[i] = (u64)(((rounded % qi) + qi) % qi);
    }
}
static int bfv_mul_schoolbook(const exo_params *p, const plan *pq, const u64 *ct1, const u64 *ct2,
                              u64 *out3) {
    const uint32_t n = p->n;
    const u64 q = p->q, half_q = q / 2;
    if (schoolbook_overflow_risk(p->plain_modulus, q, n))            /* :426-431 */
        return fail(EXO_NOT_IMPLEMENTED,
                    "schoolbook BFV multiplication can overflow i128 for these parameters; "
                    "use HPS auxiliary basis");
    i128 *cen = (i128 *)malloc(sizeof(i128) * n * 7);
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
    const u64 *src[4] = {ct1, ct1 + n, ct2, ct2 + n};
    for (int k = 0; k < 4; k++) {                                    /* :433-436 */
        memcpy(tmp, src[k], sizeof(u64) * n);
        ntt_inv(pq, tmp);
        for (uint32_t i = 0; i < n; i++)
            cen[(size_t)k * n + i] = tmp[i] > half_q ? (i128)tmp[i] - (i128)q : (i128)tmp[i];
    }
    i128 *c0 = cen, *c1 = cen + n, *d0 = cen + 2 * (size_t)n, *d1 = cen + 3 * (size_t)n;
    i128 *t0 = cen + 4 * (size_t)n, *t1 = cen + 5 * (size_t)n, *t2 = cen + 6 * (size_t)n;
    i128 *t1b = (i128 *)malloc(sizeof(i128) * n);
    poly_mul_i128(c0, d0, n, t0);                                    /* :438-440 */
    poly_mul_i128(c0, d1, n, t1);
    poly_mul_i128(c1, d0, n, t1b);
    for (uint32_t i = 0; i < n; i++) t1[i] += t1b[i];
    poly_mul_i128(c1, d1, n, t2);
    i128 *ts[3] = {t0, t1, t2};
    for (int comp = 0; comp < 3; comp++) {                           /* :442-448 */
        scale_tensor_component(ts[comp], p->plain_modulus, q, n, out3 + (size_t)comp * n);
        ntt_fwd(pq, out3 + (size_t)comp * n);
    }
    free(cen); free(tmp); free(t1b);
    return EXO_OK;
}

static int get_plans(const exo_params *p, const plan **pq, const plan **pp) {
    if (p->num_aux > EXO_MAX_AUX) return fail(EXO_INVALID_PARAM, "too many aux primes for the oracle");
    *pq = get_plan(p->n, p->q);
    if (!*pq) return EXO_INVALID_PARAM;
    for (uint32_t j = 0; j < p->num_aux; j++) {
        pp[j] = get_plan(p->n, p->aux[j]);
        if (!pp[j]) return EXO_INVALID_PARAM;
    }
    return EXO_OK;
}

/* bfv_mul_no_relin :89-108 (dispatch).  The `len != 2` check :93-97 lives in
 * the host mirror, which owns ciphertext shapes. */
int exo_bfv_mul_no_relin(const exo_params *p, const u64 *ct1, const u64 *ct2, u64 *out3) {
    const plan *pq, *pp[EXO_MAX_AUX];
    if (p->num_ct > 1)                                               /* :99-101 */
        return fail(EXO_NOT_IMPLEMENTED,
                    "multi-prime ciphertext modulus (BigInt path, bfv/eval.rs:113-147) is outside "
                    "the hot-path oracle");
    int rc = get_plans(p, &pq, pp);
    if (rc) return rc;
    if (p->num_aux > 0) return bfv_mul_hps(p, pq, pp, ct1, ct2, out3);   /* :102-104 */
    return bfv_mul_schoolbook(p, pq, ct1, ct2, out3);                    /* :105-107 */
}

/* bfv_mul_and_relin :73-82 */
int exo_bfv_mul_and_relin(const exo_params *p, const u64 *ct1, const u64 *ct2, const u64 *rlk,
                          u64 *out) {
    u64 *c3 = (u64 *)malloc(sizeof(u64) * 3 * p->n);
    int rc = exo_bfv_mul_no_relin(p, ct1, ct2, c3);
    if (rc == EXO_OK) rc = relinearize(p, get_plan(p->n, p->q), c3, rlk, out);
    free(c3);
    return rc;
}

int exo_bfv_mul_and_relin_batch(const exo_params *p, const u64 *ct1, const u64 *ct2,
                                const u64 *rlk, u64 *out, size_t batch, int threads) {
    int rc = EXO_OK;
    char msg[sizeof g_err] = {0};
    if (threads < 1) threads = 1;
    const size_t stride = 2 * (size_t)p->n;
    if (batch == 0) return EXO_OK;
    get_plan(p->n, p->q);
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t b = 0; b < batch; b++) {
        int r = exo_bfv_mul_and_relin(p, ct1 + b * stride, ct2 + b * stride, rlk, out + b * stride);
        if (r != EXO_OK) {
#pragma omp critical(exo_err)
            { rc = r; snprintf(msg, sizeof msg, "%s", g_err); }
        }
    }
    if (rc != EXO_OK) snprintf(g_err, sizeof g_err, "%s", msg);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Galois automorphism + key switch                            */
/* ------------------------------------------------------------------------ */

/* apply_automorphism bfv/keygen.rs:218-239: X^i -> X^(i*k) mod (X^n + 1), a
 * signed scatter with mod_add / mod_sub accumulation (zero coefficients skipped). */
void exo_apply_automorphism(uint32_t n, u64 q, const u64 *in, uint64_t k, u64 *out) {
    memset(out, 0, sizeof(u64) * n);
    for (uint32_t i = 0; i < n; i++) {
        u64 c = in[i];
        if (c == 0) continue;
        u64 new_exp = (u64)(((u128)i * k) % (2 * (u64)n));
        if (new_exp < n) out[new_exp] = exo_mod_add(out[new_exp], c, q);
        else out[new_exp - n] = exo_mod_sub(out[new_exp - n], c, q);
    }
}

/* bfv_apply_automorphism bfv/eval.rs:512-561.  ct [2][n], gk [G][2][n] (NTT domain),
 * out [2][n]. */
int exo_bfv_apply_automorphism(const exo_params *p, const u64 *ct, const u64 *gk, uint64_t k,
                               u64 *out) {
    const plan *pq = get_plan(p->n, p->q);
    if (!pq) return fail(EXO_INVALID_PARAM, "cannot create NTT plan");
    const uint32_t n = p->n, G = p->gadget_digits;
    const u64 q = p->q;
    u64 *c0 = (u64 *)malloc(sizeof(u64) * n), *c1 = (u64 *)malloc(sizeof(u64) * n);
    u64 *a0 = (u64 *)malloc(sizeof(u64) * n), *a1 = (u64 *)malloc(sizeof(u64) * n);
    u64 *digits = (u64 *)malloc(sizeof(u64) * n * G), *prod = (u64 *)malloc(sizeof(u64) * n);
    memcpy(c0, ct, sizeof(u64) * n);
    memcpy(c1, ct + n, sizeof(u64) * n);
    ntt_inv(pq, c0);                                                  /* :527 */
    ntt_inv(pq, c1);                                                  /* :528 */
    exo_apply_automorphism(n, q, c0, k, a0);                          /* :530 */
    exo_apply_automorphism(n, q, c1, k, a1);                          /* :531 */
    ntt_fwd(pq, a0);                                                  /* :533 */
    exo_gadget_decompose(n, q, a1, p->gadget_base, G, digits);        /* :538 */
    memcpy(out, a0, sizeof(u64) * n);
    memset(out + n, 0, sizeof(u64) * n);
    for (uint32_t g = 0; g < G; g++) {                                /* :543-555 */
        u64 *dg = digits + (size_t)g * n;
        ntt_fwd(pq, dg);
        pw_mul(n, q, dg, gk + ((size_t)g * 2 + 0) * n, prod);
        pw_add(n, q, out, prod, out);
        pw_mul(n, q, dg, gk + ((size_t)g * 2 + 1) * n, prod);
        pw_add(n, q, out + n, prod, out + n);
    }
    free(c0); free(c1); free(a0); free(a1); free(digits); free(prod);
    return EXO_OK;
}

int exo_bfv_apply_automorphism_batch(const exo_params *p, const u64 *ct, const u64 *gk, uint64_t k,
                                     u64 *out, size_t batch, int threads) {
    int rc = EXO_OK;
    if (!get_plan(p->n, p->q)) return fail(EXO_INVALID_PARAM, "cannot create NTT plan");
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic)
    for (size_t b = 0; b < batch; b++) {
        int r = exo_bfv_apply_automorphism(p, ct + b * 2 * (size_t)p->n, gk, k, out + b * 2 * (size_t)p->n);
        if (r != EXO_OK) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

/* ------------------------------------------------------------------------ */
/* dbfv/                                                                     */
/* ------------------------------------------------------------------------ */

/* mod_pow_u128 lattice.rs:234-247 */
static u128 mod_pow_u128(u128 base, u128 e, u128 m) {
    u128 r = 1;
    base %= m;
    while (e > 0) {
        if (e & 1) r = r * base % m;
        e >>= 1;
        if (e > 0) base = base * base % m;
    }
    return r;
}

/* SmallReps::compute_simple lattice.rs:104-122 + digit_decompose decomposition.rs:8-16 */
void exo_small_reps(u64 base, uint32_t d, u64 plain_modulus, int64_t *reps) {
    for (uint32_t j = d; j + 2 <= 2 * d; j++) {       /* j in d ..= 2d-2 */
        u64 val;
        if (plain_modulus == 0) {                     /* wrapping_pow :108-110 */
            val = 1;
            for (uint32_t k = 0; k < j; k++) val *= base;
        } else {
            val = (u64)mod_pow_u128(base, j, plain_modulus);
        }
        u64 remaining = val;
        for (uint32_t i = 0; i < d; i++) {
            reps[(size_t)(j - d) * d + i] = (int64_t)(remaining % base);
            remaining /= base;
        }
    }
}

/* scale_bfv_ciphertext reduction.rs:65-93: scalar_mul(|s|) then neg if s < 0. */
static void scale_bfv_ciphertext(const exo_params *p, const u64 *ct, int64_t scalar, u64 *out) {
    const u64 q = p->q;
    const size_t len = 2 * (size_t)p->n;
    if (scalar == 0) { memset(out, 0, sizeof(u64) * len); return; }
    u64 abs_s = scalar < 0 ? (u64)(-(scalar + 1)) + 1 : (u64)scalar;
    u64 s = abs_s % q;                                 /* NttPoly::scalar_mul ring/ntt.rs:132-139 */
    for (size_t i = 0; i < len; i++) {
        u64 v = exo_mod_mul(ct[i], s, q);
        out[i] = scalar < 0 ? exo_mod_neg(v, q) : v;
    }
}

/* dbfv_mul eval.rs:82-149.  The limb-count :90-94 and depth :96-102 guards
 * live in the host mirror (they read ciphertext metadata). */
int exo_dbfv_mul(const exo_params *p, u64 base, uint32_t d, u64 dbfv_plain_modulus, const u64 *ct1,
                 const u64 *ct2, const u64 *rlk, u64 *out, int threads) {
    const size_t ct_len = 2 * (size_t)p->n;
    const uint32_t result_len = 2 * d - 1, items = d * d;
    int rc = EXO_OK;
    char msg[sizeof g_err] = {0};
    if (threads < 1) threads = 1;
    if (p->num_ct == 1 && !get_plan(p->n, p->q)) return EXO_INVALID_PARAM;

    u64 *products = (u64 *)malloc(sizeof(u64) * ct_len * items);
    /* work_items (i, j, i+j) :109-114; par_iter :117-122 */
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (uint32_t w = 0; w < items; w++) {
        uint32_t i = w / d, j = w % d;
        int r = exo_bfv_mul_and_relin(p, ct1 + i * ct_len, ct2 + j * ct_len, rlk,
                                      products + (size_t)w * ct_len);
        if (r != EXO_OK) {
#pragma omp critical(exo_err)
            { rc = r; snprintf(msg, sizeof msg, "%s", g_err); }
        }
    }
    if (rc != EXO_OK) {
        free(products);
        snprintf(g_err, sizeof g_err, "%s", msg);
        return rc;
    }
    /* per-k accumulation with bfv_add in work-item order :125-132 */
    u64 *limbs = (u64 *)calloc(ct_len * result_len, sizeof(u64));
    char *have = (char *)calloc(result_len, 1);
    for (uint32_t w = 0; w < items; w++) {
        uint32_t k = w / d + w % d;
        u64 *dst = limbs + (size_t)k * ct_len;
        const u64 *src = products + (size_t)w * ct_len;
        if (have[k]) exo_bfv_add(p, dst, src, dst);
        else { memcpy(dst, src, sizeof(u64) * ct_len); have[k] = 1; }
    }
    /* reduction::reduce reduction.rs:15-60 */
    memcpy(out, limbs, sizeof(u64) * ct_len * d);                    /* :31 */
    if (result_len > d) {
        int64_t *reps = (int64_t *)malloc(sizeof(int64_t) * (d - 1) * d);
        u64 *scaled = (u64 *)malloc(sizeof(u64) * ct_len);
        exo_small_reps(base, d, dbfv_plain_modulus, reps);           /* :28 */
        for (uint32_t j = d; j < result_len; j++) {                  /* :34-52 */
            const int64_t *rep = reps + (size_t)(j - d) * d;
            for (uint32_t i = 0; i < d; i++) {
                if (rep[i] == 0) continue;
                scale_bfv_ciphertext(p, limbs + (size_t)j * ct_len, rep[i], scaled);
                exo_bfv_add(p, out + (size_t)i * ct_len, scaled, out + (size_t)i * ct_len);
            }
        }
        free(reps); free(scaled);
    }
    free(products); free(limbs); free(have);
    return EXO_OK;
}
