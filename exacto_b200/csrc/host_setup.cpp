// host_setup.cpp -- see host_setup.hpp.  Pure C++ (unsigned __int128), no CUDA.
//
// Host restatements: BfvParamsBuilder::build params/mod.rs:81-124,
// compute_gadget_digits params/mod.rs:126-140, RnsBasis::new ring/rns.rs:35-63,
// make_plan ring/ntt.rs:19-29, bfv_mul_no_relin dispatch bfv/eval.rs:89-108,
// single-aux guard bfv/eval.rs:170-178, schoolbook_overflow_risk bfv/eval.rs:457-464,
// hps_scale setup bfv/eval.rs:275-286,342-347, SmallReps::compute_simple
// dbfv/lattice.rs:104-122.
#include "host_setup.hpp"

#include <cstdlib>
#include <cstring>

namespace exb {

typedef unsigned __int128 u128;
typedef __int128 i128;

static int fail(std::string *err, int code, const std::string &msg) {
    if (err) *err = msg;
    return code;
}

static std::string u128_str(u128 v) {
    char buf[48];
    int pos = 47;
    buf[pos] = 0;
    do { buf[--pos] = (char)('0' + (int)(v % 10)); v /= 10; } while (v);
    return std::string(buf + pos);
}

// ---- number theory -----------------------------------------------------------------
static u64 h_mul(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
static u64 h_pow(u64 b, u64 e, u64 m) {
    u64 r = 1 % m;
    b %= m;
    while (e) { if (e & 1) r = h_mul(r, b, m); b = h_mul(b, b, m); e >>= 1; }
    return r;
}
static bool h_inv(u64 a, u64 m, u64 *out) {  // extended Euclid (ring/modular.rs:102-121)
    i128 old_r = a, r = m, old_s = 1, s = 0;
    while (r != 0) {
        i128 qq = old_r / r, t = r;
        r = old_r - qq * r; old_r = t;
        t = s; s = old_s - qq * s; old_s = t;
    }
    if (old_r != 1) return false;
    *out = (u64)(((old_s % (i128)m) + (i128)m) % (i128)m);
    return true;
}
static bool h_is_prime(u64 m) {
    if (m < 2) return false;
    static const u64 b[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : b) { if (m == p) return true; if (m % p == 0) return false; }
    u64 dd = m - 1; int s = 0;
    while (!(dd & 1)) { dd >>= 1; s++; }
    for (u64 a : b) {
        u64 x = h_pow(a, dd, m);
        if (x == 1 || x == m - 1) continue;
        bool comp = true;
        for (int r = 1; r < s; r++) { x = h_mul(x, x, m); if (x == m - 1) { comp = false; break; } }
        if (comp) return false;
    }
    return true;
}
static u64 shoup_of(u64 w, u64 m) { return (u64)(((u128)w << 64) / m); }
static u32 bitrev32(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

// minimal big integers (little-endian base 2^32) for compute_gadget_digits
static void bn_mul(std::vector<u32> &a, u64 m) {
    std::vector<u32> r(a.size() + 2, 0);
    const u32 mw[2] = {(u32)m, (u32)(m >> 32)};
    for (size_t i = 0; i < a.size(); i++)
        for (int j = 0; j < 2; j++) {
            u64 carry = (u64)a[i] * mw[j];
            for (size_t k = i + j; carry; k++) {
                const u64 t = (u64)r[k] + (carry & 0xffffffffu);
                r[k] = (u32)t;
                carry = (carry >> 32) + (t >> 32);
            }
        }
    while (r.size() > 1 && r.back() == 0) r.pop_back();
    a = r;
}
static int bn_cmp(const std::vector<u32> &a, const std::vector<u32> &b) {
    if (a.size() != b.size()) return a.size() < b.size() ? -1 : 1;
    for (size_t i = a.size(); i-- > 0;)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}

// make_plan (ring/ntt.rs:19-29) + the per-modulus constants of modarith.cuh.
// psi = x^((m-1)/2n) for the first x = 2, 3, .. with psi^n == m - 1.
static int build_modulus(u32 n, u32 logn, u64 m, Modulus *mod, std::vector<Tw> *twf, std::vector<Tw> *twi,
                         u64 *psi_out, std::string *err) {
    if (!h_is_prime(m) || (m - 1) % (2ull * n) != 0)
        return fail(err, EXB_INVALID_PARAM,
                    "cannot create NTT plan for n=" + std::to_string(n) + ", q=" + std::to_string(m) +
                        " (need prime q ≡ 1 mod " + std::to_string(2ull * n) + ")");
    if (m >> 62)
        return fail(err, EXB_NOT_IMPLEMENTED,
                    "device path supports NTT primes below 2^62, got " + std::to_string(m));
    const u64 e = (m - 1) / (2ull * n);
    u64 psi = 0;
    for (u64 x = 2; x < m; x++) {
        const u64 c = h_pow(x, e, m);
        if (h_pow(c, n, m) == m - 1) { psi = c; break; }
    }
    if (!psi) return fail(err, EXB_INVALID_PARAM, "no primitive 2n-th root of unity");
    u64 psi_inv = 0, ninv = 0;
    if (!h_inv(psi, m, &psi_inv) || !h_inv(n % m, m, &ninv))
        return fail(err, EXB_INVALID_PARAM, "psi or n not invertible");
    twf->resize(n); twi->resize(n);
    u64 pw = 1, ipw = 1;
    for (u32 k = 0; k < n; k++) {
        const u32 r = bitrev32(k, logn);
        (*twf)[r].w = pw;  (*twf)[r].s = shoup_of(pw, m);
        (*twi)[r].w = ipw; (*twi)[r].s = shoup_of(ipw, m);
        pw = h_mul(pw, psi, m); ipw = h_mul(ipw, psi_inv, m);
    }
    mod->m = m; mod->two_m = 2 * m;
    mod->mu = (u64)(((u128)1 << 64) / m);
    u64 inv = m;                                           // Newton: m^-1 mod 2^64
    for (int i = 0; i < 6; i++) inv *= 2 - m * inv;
    mod->minv_neg = (u64)0 - inv;
    mod->r_mod = (u64)(((u128)1 << 64) % m);
    mod->r_mod_s = shoup_of(mod->r_mod, m);
    mod->r2_mod = h_mul(mod->r_mod, mod->r_mod, m);
    mod->ninv = ninv; mod->ninv_s = shoup_of(ninv, m);
    mod->ninv_w = h_mul(ninv, (*twi)[1].w, m);
    mod->ninv_w_s = shoup_of(mod->ninv_w, m);
    mod->neg_m = (u64)0 - m;
    mod->four_m = 4 * m;
    mod->hi_four_m = (u32)(mod->four_m >> 32);
    u32 bits = 0;
    while (bits < 64 && (m >> bits)) bits++;
    mod->rhi = 0; mod->rsh = 0; mod->lazy = 0;
    if (bits >= 37 && bits <= 60) {                        // approximate-butterfly fast paths
        const u32 sh = bits - 1;                           // 2^(32+sh)/m in [2^31, 2^32)
        mod->rhi = (u32)(((u128)1 << (32 + sh)) / m);
        mod->rsh = sh - 32;
        mod->lazy = bits <= 55 ? 2 : 1;
    }
    *psi_out = psi;
    return EXB_OK;
}

static bool schoolbook_overflow_risk(u64 p, u64 q, u32 n) {   // bfv/eval.rs:457-464
    auto sat = [](u128 a, u128 b) -> u128 {
        if (a == 0 || b == 0) return 0;
        if (a > (~(u128)0) / b) return ~(u128)0;
        return a * b;
    };
    const u128 i128_max = (~(u128)0) >> 1;
    const u128 mc = q / 2, mt = sat(sat((u128)n, mc), mc), ms = sat(mt, (u128)p);
    return mt > i128_max || ms > i128_max;
}

// The reference's guard bounds n*(q/2)^2 (and times p), but the middle tensor component c0*d1 + c1*d0 reaches
// twice that: for 2*n*(q/2)^2*p > i128::MAX >= n*(q/2)^2*p its schoolbook branch wraps (release) or panics
// (debug) on worst-case inputs, so there is no well-defined result to reproduce.  Found by tests/test_gpu_fuzz.py.
static bool schoolbook_middle_term_overflow(u64 p, u64 q, u32 n) {
    auto sat = [](u128 a, u128 b) -> u128 {
        if (a == 0 || b == 0) return 0;
        if (a > (~(u128)0) / b) return ~(u128)0;
        return a * b;
    };
    const u128 i128_max = (~(u128)0) >> 1;
    const u128 mc = q / 2, mt = sat(sat(sat((u128)n, mc), mc), 2), ms = sat(mt, (u128)p);
    return mt > i128_max || ms > i128_max;
}


// ---- multi-prime ciphertext modulus: constants of rns.cuh ---------------------------------------------------
static std::vector<u32> bn_from(u64 v) { return std::vector<u32>{(u32)v, (u32)(v >> 32)}; }
static void bn_trim(std::vector<u32> &a) { while (a.size() > 1 && a.back() == 0) a.pop_back(); }
static std::vector<u32> bn_half(const std::vector<u32> &a) {
    std::vector<u32> r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = (a[i] >> 1) | (i + 1 < a.size() ? (a[i + 1] << 31) : 0u);
    return r;
}
static size_t bn_bits(std::vector<u32> a) {
    bn_trim(a);
    size_t b = (a.size() - 1) * 32;
    for (u32 t = a.back(); t; t >>= 1) b++;
    return b;
}
static void bn_store(const std::vector<u32> &a, u32 *dst, int cap) {
    for (int i = 0; i < cap; i++) dst[i] = (size_t)i < a.size() ? a[i] : 0u;
}

// Builds HostSetup::R / plans for 2 <= L <= 4 ciphertext primes with Q < 2^127 (the range in which the
// reference's own relinearize, whose to_coeff_poly multiplies in u128, is defined without overflow).
static int build_rns(HostSetup *c, std::string *err) {
    c->rns_enabled = false;
    c->mul_error.clear();
    const u32 L = (u32)c->ct_moduli.size(), n = c->n;
    if (L < 2) return EXB_OK;
    memset(&c->R, 0, sizeof c->R);
    memset(&c->T, 0, sizeof c->T);
    auto refuse = [&](const std::string &why) { c->mul_error = why; return EXB_OK; };
    if (L > (u32)kRnsMaxL) return refuse("multi-prime ciphertext modulus: the device path supports up to 4 ciphertext primes");
    for (u32 i = 0; i < L; i++)
        for (u32 j = i + 1; j < L; j++)
            if (c->ct_moduli[i] == c->ct_moduli[j]) return refuse("non-coprime ciphertext moduli");
    std::vector<u32> Q{1};
    for (u64 m : c->ct_moduli) bn_mul(Q, m);
    bn_trim(Q);
    const size_t qbits = bn_bits(Q);
    if (qbits > 126)
        return refuse("multi-prime ciphertext modulus: Q must stay below 2^126 (the reference's relinearize reconstructs c2 "
                      "with u128 products, ring/rns.rs:133-150, which overflow beyond that)");
    if (c->gadget_base < 2) return refuse("device path supports gadget bases >= 2");
    if (c->gadget_digits > 128) return refuse("device path supports up to 128 gadget digits");
    RnsConsts &R = c->R;
    R.L = L; R.n = n; R.logn = c->logn;
    R.gadget_digits = c->gadget_digits; R.gadget_base = c->gadget_base; R.plain = c->plain;
    R.gadget_log2 = c->P.gadget_log2;
    R.Qlen = (u32)Q.size();
    bn_store(Q, R.Q, kMwQ);
    bn_store(bn_half(Q), R.halfQ, kMwQ);
    u128 bigq = 1;
    for (u64 m : c->ct_moduli) bigq *= m;
    R.bigq_lo = (u64)bigq; R.bigq_hi = (u64)(bigq >> 64);
    if (R.bigq_lo == 0) return refuse("multi-prime ciphertext modulus: Q = 0 mod 2^64 (the reference's gadget_decompose divides by it)");
    for (u32 l = 0; l < L; l++) {
        const u64 ql = c->ct_moduli[l];
        u64 psi = 0;
        int rc = build_modulus(n, c->logn, ql, &R.q[l], &c->rns_twf_q[l], &c->rns_twi_q[l], &psi, err);
        if (rc != EXB_OK) return rc;
        for (u32 k = 0; k < 16 && k < n; k++) { c->T.headf_q[l].t[k] = c->rns_twf_q[l][k]; c->T.headi_q[l].t[k] = c->rns_twi_q[l][k]; }
        std::vector<u32> qs{1};
        u64 prod = 1;
        for (u32 j = 0; j < L; j++)
            if (j != l) { bn_mul(qs, c->ct_moduli[j]); prod = h_mul(prod, c->ct_moduli[j] % ql, ql); }
        bn_store(qs, R.Qstar[l], kMwQ);
        u64 inv = 0;
        if (!h_inv(prod, ql, &inv)) return refuse("non-coprime ciphertext moduli");
        R.crt_inv[l] = inv; R.crt_inv_s[l] = shoup_of(inv, ql);
        R.c32_q[l] = (u64)(((u128)1 << 32) % ql); R.c32_q_s[l] = shoup_of(R.c32_q[l], ql);
        const u128 qstar = bigq / ql;
        R.qstar_lo[l] = (u64)qstar; R.qstar_hi[l] = (u64)(qstar >> 64);
    }
    // extended basis: 61-bit primes = 1 mod 2n, prod(E) >= 4 n Q^2 > 4 max |t|
    const size_t need_bits = 2 * qbits + c->logn + 3;
    std::vector<u32> E{1};
    c->ext_primes.clear();
    const u64 step = 2ull * n, top = (u64)1 << 61;
    for (u64 cand = (top - 1) / step * step + 1; cand > top / 2 && bn_bits(E) <= need_bits; cand -= step) {
        bool clash = false;
        for (u64 m : c->ct_moduli) clash |= (m == cand);
        if (clash || !h_is_prime(cand)) continue;
        if (c->ext_primes.size() == (size_t)kRnsMaxK) return refuse("multi-prime ciphertext modulus: extended basis too large");
        c->ext_primes.push_back(cand);
        bn_mul(E, cand);
        bn_trim(E);
    }
    if (bn_bits(E) <= need_bits) return refuse("multi-prime ciphertext modulus: not enough NTT primes for the extended basis");
    const u32 K = (u32)c->ext_primes.size();
    R.K = K;
    R.Elen = (u32)E.size();
    if (R.Elen > (u32)kMwE) return refuse("multi-prime ciphertext modulus: extended basis too large");
    bn_store(E, R.E, kMwE);
    bn_store(bn_half(E), R.halfE, kMwE);
    for (u32 k = 0; k < K; k++) {
        const u64 ek = c->ext_primes[k];
        u64 psi = 0;
        int rc = build_modulus(n, c->logn, ek, &R.e[k], &c->rns_twf_e[k], &c->rns_twi_e[k], &psi, err);
        if (rc != EXB_OK) return rc;
        for (u32 t = 0; t < 16 && t < n; t++) { c->T.headf_e[k].t[t] = c->rns_twf_e[k][t]; c->T.headi_e[k].t[t] = c->rns_twi_e[k][t]; }
        R.c32_e[k] = (u64)(((u128)1 << 32) % ek); R.c32_e_s[k] = shoup_of(R.c32_e[k], ek);
        for (u32 j = 0; j < k; j++) {
            u64 inv = 0;
            if (!h_inv(c->ext_primes[j] % ek, ek, &inv)) return refuse("extended basis primes not coprime");
            R.garner[k][j] = inv; R.garner_s[k][j] = shoup_of(inv, ek);
        }
    }
    c->rns_enabled = true;
    return EXB_OK;
}

// The dispatch of bfv_mul_no_relin (bfv/eval.rs:89-108), evaluated once per parameter set.
static void decide_mul_support(HostSetup *c) {
    const u32 A = (u32)c->aux_moduli.size();
    const u64 q = c->ct_moduli[0];
    c->mul_status = EXB_OK;
    if (c->ct_moduli.size() > 1) {
        // bfv_mul_generic_rns (bfv/eval.rs:113-147): handled by the extended-basis path when build_rns succeeded
        if (!c->rns_enabled) {
            c->mul_status = EXB_NOT_IMPLEMENTED;
            if (c->mul_error.empty())
                c->mul_error = "multi-prime ciphertext modulus outside the device path's range";
        }
    } else if (A == 0) {
        c->mul_status = EXB_NOT_IMPLEMENTED;
        c->mul_error = schoolbook_overflow_risk(c->plain, q, c->n)
                           ? "schoolbook BFV multiplication can overflow i128 for these parameters; use HPS "
                             "auxiliary basis"
                           : schoolbook_middle_term_overflow(c->plain, q, c->n)
                                 ? "schoolbook BFV multiplication overflows i128 in its middle tensor term for these "
                                   "parameters (2*n*(q/2)^2*p > i128::MAX: the reference result is undefined); use "
                                   "HPS auxiliary basis"
                                 : "schoolbook BFV multiplication (no auxiliary basis) is not provided by the device "
                                   "library for these parameters; use HPS auxiliary basis";
    } else if (A == 1 && (u128)c->aux_moduli[0] <= ((u128)c->n * q) / 2) {   // bfv/eval.rs:170-178
        c->mul_status = EXB_INVALID_PARAM;
        c->mul_error = "single aux prime too small for HPS centering: P=" + std::to_string(c->aux_moduli[0]) +
                       " <= n*Q/2=" + u128_str(((u128)c->n * q) / 2);
    } else if (A > 2) {                                                       // bfv/eval.rs:405-409
        c->mul_status = EXB_INVALID_PARAM;
        c->mul_error = "HPS scaling supports 1 or 2 aux primes, got " + std::to_string(A);
    } else if (c->plain >= q) {
        c->mul_status = EXB_NOT_IMPLEMENTED;
        c->mul_error = "device path needs plain_modulus < ciphertext prime";
    } else if (c->gadget_base < 2 || c->gadget_base > (1ull << 32)) {
        c->mul_status = EXB_NOT_IMPLEMENTED;
        c->mul_error = "device path supports gadget bases in [2, 2^32]";
    }
}

static int fill_scale_consts(HostSetup *c, std::string *err) {
    ScaleConsts &s = c->P.sc;
    memset(&s, 0, sizeof s);
    const u32 A = (u32)c->aux_moduli.size();
    const u64 q = c->ct_moduli[0], p = c->plain;
    s.q = q; s.half_q = q / 2; s.num_aux = A;
    if (p < q) { s.plain = p; s.plain_s = shoup_of(p, q); }   // decrypt needs these without an aux basis
    if (A == 0 || A > 2 || p >= q) return EXB_OK;      // multiplication is refused by decide_mul_support
    u128 big_p = 1;
    for (u32 j = 0; j < A; j++) {
        s.pj[j] = c->aux_moduli[j];
        s.pj_mu[j] = (u64)(((u128)1 << 64) / s.pj[j]);
        big_p *= s.pj[j];
    }
    s.big_p.lo = (u64)big_p; s.big_p.hi = (u64)(big_p >> 64);
    const u128 hp = big_p / 2;
    s.half_big_p.lo = (u64)hp; s.half_big_p.hi = (u64)(hp >> 64);
    s.plain = p; s.plain_s = shoup_of(p, q);
    u64 qinv[2] = {0, 0};
    for (u32 j = 0; j < A; j++)
        if (!h_inv(q % s.pj[j], s.pj[j], &qinv[j]))
            return fail(err, EXB_INVALID_PARAM, "q not invertible mod p_j");
    if (A == 1) {
        s.K[0] = qinv[0];
        s.C[0] = p;
    } else {
        u64 p1_inv_p0, p0_inv_p1;
        if (!h_inv(s.pj[1] % s.pj[0], s.pj[0], &p1_inv_p0) || !h_inv(s.pj[0] % s.pj[1], s.pj[1], &p0_inv_p1))
            return fail(err, EXB_INVALID_PARAM, "aux primes not coprime");
        s.K[0] = h_mul(qinv[0], p1_inv_p0, s.pj[0]);
        s.K[1] = h_mul(qinv[1], p0_inv_p1, s.pj[1]);
        s.other[0] = s.pj[1]; s.other[1] = s.pj[0];
        s.C[0] = h_mul(p, s.pj[1] % q, q);
        s.C[1] = h_mul(p, s.pj[0] % q, q);
    }
    for (u32 j = 0; j < A; j++) {
        s.K_s[j] = shoup_of(s.K[j], s.pj[j]);
        s.C_s[j] = shoup_of(s.C[j], q);
    }
    s.CP = h_mul(p, (u64)(big_p % q), q);
    s.CP2 = (u64)(((u128)s.CP * 2) % q);
    return EXB_OK;
}

// ---- internal 27-bit auxiliary basis (ntt32_core.cuh, hps32.cuh) ---------------------------
static int build_mod32(u32 n, u32 logn, u32 p, Mod32 *m, std::vector<Tw32> *twf, std::vector<Tw32> *twi) {
    const u64 e = (p - 1) / (2ull * n);
    u64 psi = 0;
    for (u64 x = 2; x < p; x++) {
        const u64 c = h_pow(x, e, p);
        if (h_pow(c, n, p) == (u64)p - 1) { psi = c; break; }
    }
    u64 psi_inv = 0, ninv = 0;
    if (!psi || !h_inv(psi, p, &psi_inv) || !h_inv(n % p, p, &ninv)) return EXB_INVALID_PARAM;
    auto sh = [&](u64 w) { return (u32)((w << 32) / p); };
    twf->resize(n); twi->resize(n);
    u64 pw = 1, ipw = 1;
    for (u32 k = 0; k < n; k++) {
        const u32 r = bitrev32(k, logn);
        (*twf)[r].w = (u32)pw;  (*twf)[r].s = sh(pw);
        (*twi)[r].w = (u32)ipw; (*twi)[r].s = sh(ipw);
        pw = pw * psi % p; ipw = ipw * psi_inv % p;
    }
    memset(m, 0, sizeof *m);
    m->p = p; m->two_p = 2 * p; m->neg_p = (u32)0 - p; m->four_p = 4 * p;
    u32 inv = p;
    for (int i = 0; i < 5; i++) inv *= 2 - p * inv;
    m->pinv_neg = (u32)0 - inv;
    m->r_mod = (u32)(((u64)1 << 32) % p); m->r_mod_s = sh(m->r_mod);
    m->one_s = (u32)(((u64)1 << 32) / p);
    m->c28 = (u32)(((u64)1 << 28) % p); m->c28_s = sh(m->c28);
    m->ninv = (u32)ninv; m->ninv_s = sh(ninv);
    m->ninv_w = (u32)(ninv * (*twi)[1].w % p); m->ninv_w_s = sh(m->ninv_w);
    return EXB_OK;
}

// Enable the internal basis only when the result is provably identical to the reference's:
//   |m| <= Mmax = n*q/2 + 2 for every input (|t| <= 2 n (q/2)^2 for the middle tensor term);
//   the reference's centred CRT of m (bfv/eval.rs:316-321, :385-388) cannot wrap: P_ref/2 > Mmax;
//   ours cannot either and alpha is exact: P' >= 2^8 * Mmax (fixed-point error < 2^-28).
static void build_small_basis(HostSetup *c, uint32_t flags) {
    SmallBasis &sb = c->P.sb;
    memset(&sb, 0, sizeof sb);
    if (flags & EXB_CTX_REFERENCE_AUX_BASIS) return;
    const u32 A = (u32)c->aux_moduli.size(), n = c->n;
    if (c->logn != 12 || c->mul_status != EXB_OK || A < 1 || A > (u32)kMaxAux) return;
    const u64 q = c->ct_moduli[0];
    if (c->P.mod[0].lazy == 0) return;                       // needs 2^36 <= q < 2^60
    const u128 mmax = ((u128)n * q) / 2 + 2;
    u128 pref = 1;
    for (u64 a : c->aux_moduli) pref *= a;
    if (pref / 2 <= mmax) return;
    std::vector<u32> primes;
    u128 prod = 1;
    const u128 need = mmax << 8;
    // primes in (2^26, 2^27): 32p fits a u32 (lazy NTT ranges, ntt32_core.cuh) and 2^57 / p fits a u32
    constexpr u64 top = (u64)1 << kSmallPrimeBits;
    for (u64 cand = (top - 1) / (2ull * n) * (2ull * n) + 1; cand > top / 2 && prod < need; cand -= 2ull * n) {
        if (cand >= top || !h_is_prime(cand) || cand == q) continue;
        bool clash = false;
        for (u64 a : c->aux_moduli) clash |= (a == cand);
        if (clash) continue;
        primes.push_back((u32)cand);
        prod *= cand;
        if (primes.size() > (size_t)kMaxSmall) return;
    }
    if (prod < need || primes.size() > (size_t)kMaxSmall) return;
    const u32 K = (u32)primes.size();
    Scale32Consts &s = sb.sc;
    s.K = K;
    const u64 p = c->plain;
    for (u32 i = 0; i < K; i++) {
        if (build_mod32(n, c->logn, primes[i], &s.m[i], &c->twf32[i], &c->twi32[i]) != EXB_OK) return;
        for (u32 k = 0; k < 16; k++) { sb.headf[i].t[k] = c->twf32[i][k]; sb.headi[i].t[k] = c->twi32[i][k]; }
        const u64 pi = primes[i];
        const u128 cof = prod / pi;                            // P'/p_i
        u64 inv = 0;
        if (!h_inv((u64)((u128)(q % pi) * (u64)(cof % pi) % pi), pi, &inv)) return;
        auto sh32 = [&](u64 w) { return (u32)((w << 32) / pi); };
        s.Kp[i] = (u32)inv; s.Kp_s[i] = sh32(inv);
        const u64 rk = (((u64)1 << 32) % pi) * inv % pi;
        s.RK[i] = (u32)rk; s.RK_s[i] = sh32(rk);
        s.QK[i] = (u32)((q % pi) * inv % pi);
        // the inverse transforms of this basis only feed hps_scale32_coeff: fold Kp_i into n^-1
        Mod32 &m = s.m[i];
        // ... and 2^32: the point-wise products are plain REDCs of operands that are not in Montgomery form
        const u64 r32 = ((u64)1 << 32) % pi;
        const u64 nk = (u64)m.ninv * inv % pi * r32 % pi, nwk = (u64)m.ninv_w * inv % pi * r32 % pi;
        m.ninv = (u32)nk; m.ninv_s = sh32(nk);
        m.ninv_w = (u32)nwk; m.ninv_w_s = sh32(nwk);
        s.g[i] = (u32)(((u64)1 << 57) / pi);
        s.C[i] = h_mul(p % q, (u64)(cof % q), q);
    }
    const u64 pp = h_mul(p % q, (u64)(prod % q), q);
    for (u32 a = 0; a <= K; a++) s.CPn[a] = q - (u64)((u128)pp * a % q);
    u32 bits = 0;
    while (bits < 64 && (q >> bits)) bits++;
    s.sh = bits - 2;
    s.rq = (u32)((((u128)1) << (32 + s.sh)) / q);
    s.plain32 = (p < ((u64)1 << 32) && c->P.sc.plain_s < ((u64)1 << 32)) ? 1u : 0u;
    // tensor01_kernel sums the products of one output limb before the small-prime inverse transforms:
    // cnt terms are exact while cnt*|m|max still leaves the alpha margin (2^5), the sum of cnt centred residues
    // (|.| <= cnt*q/2) fits an i64, and the summed rounding terms (each <= p/2 + 1) stay below q in magnitude;
    // max_terms_r32: ... and below 2^31 (tensor01_kernel then keeps them in an i32 image).
    u32 mt = 0, mt32 = 0;
    for (u32 cnt = 1; cnt <= 16; cnt++) {
        if (prod < (mmax * cnt) << 5) break;
        if ((u128)cnt * q >= ((u128)1 << 64)) break;
        if ((u128)cnt * (p / 2 + 2) >= q) break;
        mt = cnt;
        if ((u128)cnt * (p / 2 + 2) < ((u128)1 << 31)) mt32 = cnt;
    }
    sb.max_terms = mt;
    sb.max_terms_r32 = mt32;
    sb.mq_r = c->P.mod[0];
    {
        const u64 r64 = (u64)((((u128)1) << 64) % q);
        sb.mq_r.ninv = h_mul(c->P.mod[0].ninv, r64, q);     sb.mq_r.ninv_s = shoup_of(sb.mq_r.ninv, q);
        sb.mq_r.ninv_w = h_mul(c->P.mod[0].ninv_w, r64, q); sb.mq_r.ninv_w_s = shoup_of(sb.mq_r.ninv_w, q);
    }
    c->small_primes.assign(primes.begin(), primes.end());
    sb.K = K;
    sb.enabled = 1;
}

int host_setup_build(const exb_bfv_params *p, HostSetup *c, std::string *err, uint32_t flags) {
    if (!p || !c) return fail(err, EXB_INVALID_PARAM, "null argument");
    const u32 n = p->ring_degree;
    if (n < 2 || (n & (n - 1)))                                           // params/mod.rs:82-84
        return fail(err, EXB_INVALID_RING_DEGREE, "ring degree must be a power of 2, got " + std::to_string(n));
    if (p->num_ct_moduli == 0 || !p->ct_moduli)                           // :85-87
        return fail(err, EXB_INVALID_PARAM, "must specify at least one ciphertext modulus");
    if (p->plain_modulus < 2)                                             // :88-90
        return fail(err, EXB_INVALID_PARAM, "plaintext modulus must be >= 2");
    if (n > 8192) return fail(err, EXB_NOT_IMPLEMENTED, "device path supports ring degrees up to 8192");
    if (p->num_aux_moduli && !p->aux_moduli) return fail(err, EXB_INVALID_PARAM, "null aux_moduli");

    c->n = n; c->logn = 0;
    while ((1u << c->logn) < n) c->logn++;
    c->ct_moduli.assign(p->ct_moduli, p->ct_moduli + p->num_ct_moduli);
    c->aux_moduli.clear();
    if (p->num_aux_moduli) c->aux_moduli.assign(p->aux_moduli, p->aux_moduli + p->num_aux_moduli);
    c->plain = p->plain_modulus;
    c->user_aux = (u32)c->aux_moduli.size();
    c->internal_aux = false;
    // No auxiliary basis: the reference multiplies by exact i128 schoolbook convolution and rounds
    // round(p * t / q) per coefficient (bfv/eval.rs:415-464), refusing when i128 could overflow
    // (:457-464).  That result is the HPS closed form with ANY auxiliary basis large enough to hold
    // m = (t - [t]_q) / q without wrapping (q is an odd prime that divides neither 2p nor a non-multiple t,
    // so the rounding never ties), so the device synthesises two 61-bit NTT primes (P ~ 2^122 >> n*q) and runs
    // its HPS pipeline; tests compare it with the literal O(n^2) restatement.
    if (c->user_aux == 0 && c->ct_moduli.size() == 1 && n >= 2 && c->plain < c->ct_moduli[0] &&
        !schoolbook_overflow_risk(c->plain, c->ct_moduli[0], n) &&
        !schoolbook_middle_term_overflow(c->plain, c->ct_moduli[0], n)) {
        const u64 step = 2ull * n, top = (u64)1 << 61;
        for (u64 cand = (top - 1) / step * step + 1; cand > top / 2 && c->aux_moduli.size() < 2; cand -= step)
            if (cand != c->ct_moduli[0] && h_is_prime(cand)) c->aux_moduli.push_back(cand);
        if (c->aux_moduli.size() == 2) c->internal_aux = true;
        else c->aux_moduli.clear();
    }
    // gadget parameters: params/mod.rs:98-112, compute_gadget_digits :126-140
    c->gadget_base = p->gadget_base ? p->gadget_base : (1ull << 16);
    if (p->gadget_digits) {
        c->gadget_digits = p->gadget_digits;
    } else if (c->gadget_base < 2) {
        c->gadget_digits = 1;
    } else {
        std::vector<u32> Q{1}, pw{1};                  // smallest k with base^k >= Q = prod q_i
        for (u64 m : c->ct_moduli) bn_mul(Q, m);
        u32 digits = 0;
        while (bn_cmp(pw, Q) < 0) { bn_mul(pw, c->gadget_base); digits++; }
        c->gadget_digits = digits ? digits : 1;
    }
    c->digits32 = c->gadget_base > 65536;

    memset(&c->P, 0, sizeof c->P);
    c->P.n = n; c->P.logn = c->logn;
    const u32 A = (u32)c->aux_moduli.size();
    c->P.num_aux = A <= (u32)kMaxAux ? A : 0;
    c->P.gadget_digits = c->gadget_digits;
    c->P.gadget_base = c->gadget_base;
    c->P.gadget_log2 = 0;
    if (c->gadget_base >= 2 && (c->gadget_base & (c->gadget_base - 1)) == 0)
        for (u32 w = 1; w < 64; w++) if ((1ull << w) == c->gadget_base) c->P.gadget_log2 = w;

    // modulus indices: 0 = q_0, 1..A = aux primes, then the remaining ciphertext primes.
    // Device plans: q_0 always, aux primes when the HPS path can use them (A <= 2); any
    // other modulus is validated (make_plan) but gets no device plan.
    std::vector<u64> all;
    all.push_back(c->ct_moduli[0]);
    for (u64 a : c->aux_moduli) all.push_back(a);
    for (size_t i = 1; i < c->ct_moduli.size(); i++) all.push_back(c->ct_moduli[i]);
    c->psi.assign(all.size(), 0);
    for (int b = 0; b < kMaxBases; b++) { c->has_plan[b] = false; c->twf[b].clear(); c->twi[b].clear(); }
    for (size_t b = 0; b < all.size(); b++) {
        Modulus mod; std::vector<Tw> twf, twi;
        int rc = build_modulus(n, c->logn, all[b], &mod, &twf, &twi, &c->psi[b], err);
        if (rc != EXB_OK) return rc;
        if (b == 0 || (b <= A && A <= (u32)kMaxAux)) {
            c->P.mod[b] = mod; c->has_plan[b] = true;
            for (u32 k = 0; k < 16 && k < n; k++) { c->P.headf[b].t[k] = twf[k]; c->P.headi[b].t[k] = twi[k]; }
            c->twf[b].swap(twf); c->twi[b].swap(twi);
        }
    }
    int rc = fill_scale_consts(c, err);
    if (rc != EXB_OK) return rc;
    if ((rc = build_rns(c, err)) != EXB_OK) return rc;
    decide_mul_support(c);
    build_small_basis(c, flags);
    return EXB_OK;
}

// ---- dBFV plan ------------------------------------------------------------------------
static u128 pow_mod_u128(u128 b, u128 e, u128 m) {       // dbfv/lattice.rs:234-247
    u128 r = 1;
    b %= m;
    while (e > 0) { if (e & 1) r = r * b % m; e >>= 1; if (e > 0) b = b * b % m; }
    return r;
}

int host_small_reps(u64 base, u32 d, u64 pm, int64_t *reps, std::string *err) {
    if (base < 2 || d < 1 || !reps) return fail(err, EXB_INVALID_PARAM, "bad small-reps arguments");
    for (u32 j = d; j + 2 <= 2 * d; j++) {
        u64 val;
        if (pm == 0) { val = 1; for (u32 k = 0; k < j; k++) val *= base; }   // wrapping_pow :108-110
        else val = (u64)pow_mod_u128(base, j, pm);
        for (u32 i = 0; i < d; i++) { reps[(size_t)(j - d) * d + i] = (int64_t)(val % base); val /= base; }
    }
    return EXB_OK;
}

int host_build_plan(u32 d, u64 base, u64 pm, u32 flags, u32 limb_mask, HostPlan *hp, std::string *err) {
    if (d < 1 || d > (u32)kMaxDigits)
        return fail(err, EXB_NOT_IMPLEMENTED, "device path supports 1..16 dBFV digits");
    MulPlan &M = hp->M;
    memset(&M, 0, sizeof M);
    M.d = d;
    hp->num_low = 0;
    hp->reps.assign((size_t)(d > 1 ? d - 1 : 0) * d, 0);
    if (d > 1) {
        int rc = host_small_reps(base, d, pm, hp->reps.data(), err);
        if (rc) return rc;
    }
    const u32 all = (1u << d) - 1u;
    const u32 mask = (limb_mask & all) ? (limb_mask & all) : all;
    std::vector<bool> need(2 * d - 1, false);
    for (u32 k = 0; k < d; k++) need[k] = (mask >> k) & 1u;
    for (u32 j = d; j + 1 < 2 * d; j++) {
        bool used = (flags & EXB_DBFV_ALL_PRODUCTS) != 0;
        for (u32 i = 0; i < d && !used; i++)
            if (((mask >> i) & 1u) && hp->reps[(size_t)(j - d) * d + i] != 0) used = true;
        need[j] = used;
    }
    hp->excess_index.assign(2 * d - 1, -1);
    u32 nl = 0, nx = 0;
    for (u32 k = 0; k < 2 * d - 1; k++) {
        if (!need[k]) continue;
        M.limb_k[nl++] = (uint8_t)k;
        if (k < d) hp->num_low++;
        else hp->excess_index[k] = (int)nx++;
    }
    M.num_limbs = nl;
    M.num_low = hp->num_low;
    u32 np = 0;
    for (u32 i = 0; i < d; i++)
        for (u32 j = 0; j < d; j++) {
            if (need[i + j]) {
                M.prod_i[np] = (uint8_t)i; M.prod_j[np] = (uint8_t)j; M.prod_of[i][j] = (int16_t)np; np++;
                M.need_lhs |= 1u << i; M.need_rhs |= 1u << j;
            } else {
                M.prod_of[i][j] = -1;
            }
        }
    M.num_products = np;
    {   // pair the computed limbs heaviest-with-lightest (a limb k sums min(k, d-1) - max(0, k-d+1) + 1 products)
        auto cnt = [&](u32 l) { const u32 k = M.limb_k[l]; return (k < d ? k : d - 1) - (k >= d ? k - d + 1 : 0) + 1; };
        std::vector<u32> order(nl);
        for (u32 l = 0; l < nl; l++) order[l] = l;
        for (u32 a = 1; a < nl; a++)
            for (u32 b = a; b > 0 && cnt(order[b]) > cnt(order[b - 1]); b--) { const u32 t = order[b]; order[b] = order[b - 1]; order[b - 1] = t; }
        M.num_duos = (nl + 1) / 2;
        for (u32 i = 0; i < M.num_duos; i++) {
            M.duo_a[i] = (uint8_t)order[i];
            M.duo_b[i] = (nl - 1 - i > i) ? (uint8_t)order[nl - 1 - i] : (uint8_t)0xFF;
        }
    }
    return EXB_OK;
}

}  // namespace exb
