// exacto_b200.hpp -- C++ host mirror of the reference's public surface for the
// ciphertext-multiplication path, header-only, over the C ABI of exacto_b200.h.
//
// The reference is Rust; there is no Rust toolchain in this image, so the compiled-language host
// side is C++: same type and function names, argument meaning, ownership (inputs borrowed, fresh
// outputs sharing the parameter object) and error behaviour (ExactoError variants and the message
// substrings the reference's tests pin).  Citations are into /root/reference/src/.
//
//   ring/    CoeffPoly (ring/poly.rs:6-9), NttPoly (ring/ntt.rs:11-15), RnsPoly (ring/rns.rs:14-17)
//   params/  BfvParams + BfvParamsBuilder (params/mod.rs:11-124), DbfvParams (params/mod.rs:143-192),
//            compact_bfv / compact_dbfv / u64_dbfv (params/presets.rs:24-98)
//   bfv/     BfvCiphertext (bfv/mod.rs:19-24), RelinKey (bfv/keygen.rs:39-45),
//            bfv_add (bfv/eval.rs:14-31), bfv_mul_and_relin (bfv/eval.rs:73-82)
//   dbfv/    DbfvCiphertext (dbfv/ciphertext.rs:10-22), dbfv_add (dbfv/eval.rs:11-33), dbfv_mul (dbfv/eval.rs:82-149)
#pragma once
#include <cstdint>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "exacto_b200.h"

namespace exacto {

// ---- error.rs:4-31 ---------------------------------------------------------------------------
class ExactoError : public std::runtime_error {
public:
    enum Kind { InvalidParam = 1, DimensionMismatch, ModulusMismatch, InvalidRingDegree, DecryptionError,
                DecompositionError, LatticeError, MissingKey, NotImplemented, Cuda = 100 };
    ExactoError(Kind k, const std::string &msg) : std::runtime_error(render(k, msg)), kind(k), detail(msg) {}
    Kind kind;
    std::string detail;
private:
    static std::string render(Kind k, const std::string &m) {
        switch (k) {
            case InvalidParam: return "invalid parameter: " + m;
            case DimensionMismatch: return "dimension mismatch: " + m;
            case ModulusMismatch: return "modulus mismatch";
            case InvalidRingDegree: return m;
            case NotImplemented: return "not yet implemented: " + m;
            case Cuda: return "CUDA error: " + m;
            default: return m;
        }
    }
};

inline void check(int rc) {
    if (rc != EXB_OK) throw ExactoError(static_cast<ExactoError::Kind>(rc), exb_last_error());
}

// ---- params/ -----------------------------------------------------------------------------------
struct BfvParams {                                   // params/mod.rs:11-27
    size_t ring_degree = 4096;
    uint64_t plain_modulus = 65537;
    std::vector<uint64_t> ct_moduli, aux_moduli;
    double sigma = 3.2;
    uint64_t gadget_base = 0;
    uint32_t gadget_digits = 0;

    ~BfvParams() { if (ctx_) exb_context_destroy(ctx_); }
    BfvParams() = default;
    BfvParams(const BfvParams &) = delete;

    // The GPU context (plans, HPS constants, workspace) is created on first use.
    exb_context *context(int device = 0) const {
        if (!ctx_) {
            exb_bfv_params p{};
            p.ring_degree = (uint32_t)ring_degree;
            p.num_ct_moduli = (uint32_t)ct_moduli.size(); p.ct_moduli = ct_moduli.data();
            p.num_aux_moduli = (uint32_t)aux_moduli.size(); p.aux_moduli = aux_moduli.data();
            p.plain_modulus = plain_modulus; p.gadget_base = gadget_base; p.gadget_digits = gadget_digits;
            check(exb_context_create(&p, device, &ctx_));
        }
        return ctx_;
    }
    uint64_t q() const { return ct_moduli.at(0); }
private:
    mutable exb_context *ctx_ = nullptr;
};

class BfvParamsBuilder {                             // params/mod.rs:29-124
public:
    BfvParamsBuilder &ring_degree(size_t n) { p_->ring_degree = n; return *this; }
    BfvParamsBuilder &plain_modulus(uint64_t p) { p_->plain_modulus = p; return *this; }
    BfvParamsBuilder &ct_moduli(std::vector<uint64_t> m) { p_->ct_moduli = std::move(m); return *this; }
    BfvParamsBuilder &aux_moduli(std::vector<uint64_t> m) { p_->aux_moduli = std::move(m); return *this; }
    BfvParamsBuilder &sigma(double s) { p_->sigma = s; return *this; }
    BfvParamsBuilder &gadget_base(uint64_t b) { p_->gadget_base = b; return *this; }
    // Validation (ring degree, moduli, plans) and the gadget-digit default run in exb_context_create,
    // with the reference's error variants; build() forces it so errors surface here like in Rust.
    std::shared_ptr<BfvParams> build() {
        std::shared_ptr<BfvParams> out = p_;
        exb_context *c = out->context();
        uint64_t gb = 0; uint32_t gd = 0;
        check(exb_context_gadget(c, &gb, &gd));
        out->gadget_base = gb; out->gadget_digits = gd;
        p_ = std::make_shared<BfvParams>();
        return out;
    }
private:
    std::shared_ptr<BfvParams> p_ = std::make_shared<BfvParams>();
};

struct DbfvParams {                                  // params/mod.rs:143-192
    std::shared_ptr<BfvParams> bfv_params;
    uint64_t base;
    size_t num_digits;
    uint64_t plain_modulus;                          // 0 == 2^64
    static std::shared_ptr<DbfvParams> create(std::shared_ptr<BfvParams> bfv, uint64_t base, size_t d, uint64_t p) {
        if (base < 2) throw ExactoError(ExactoError::InvalidParam, "base must be >= 2");
        if (d < 1) throw ExactoError(ExactoError::InvalidParam, "num_digits must be >= 1");
        unsigned __int128 bd = 1;
        for (size_t i = 0; i < d; i++) { if (bd > (~(unsigned __int128)0) / base) { bd = ~(unsigned __int128)0; break; } bd *= base; }
        const unsigned __int128 p128 = p == 0 ? ((unsigned __int128)1 << 64) : p;
        if (bd < p128) throw ExactoError(ExactoError::InvalidParam, "base^digits < plain_modulus");
        return std::make_shared<DbfvParams>(DbfvParams{std::move(bfv), base, d, p});
    }
};

inline std::shared_ptr<BfvParams> compact_bfv() {    // params/presets.rs:24-35
    return BfvParamsBuilder().ring_degree(1024).plain_modulus(257).ct_moduli({1099509805057ull})
        .aux_moduli({562949953443841ull}).build();
}
inline std::shared_ptr<DbfvParams> compact_dbfv() {  // params/presets.rs:86-98
    auto bfv = BfvParamsBuilder().ring_degree(1024).plain_modulus(929).ct_moduli({1099509805057ull})
                   .aux_moduli({562949953443841ull}).build();
    return DbfvParams::create(bfv, 16, 2, 256);
}
inline std::shared_ptr<DbfvParams> u64_dbfv() {      // params/presets.rs:61-75
    auto bfv = BfvParamsBuilder().ring_degree(4096).plain_modulus(1040407).ct_moduli({1152921504606830593ull})
                   .aux_moduli({18014398509998081ull, 36028797018972161ull}).gadget_base(256).build();
    return DbfvParams::create(bfv, 256, 8, 0);
}

// ---- ring/ (single ciphertext prime: what the hot path uses) ---------------------------------------
struct CoeffPoly {                                   // ring/poly.rs:6-9
    std::vector<uint64_t> coeffs;
    uint64_t modulus;
};
struct NttPoly {                                     // ring/ntt.rs:11-15 (the plan is the params' context)
    std::vector<uint64_t> evals;
    uint64_t modulus;
    static NttPoly from_coeff_poly(const CoeffPoly &p, const BfvParams &params) {   // ring/ntt.rs:42-55
        if (p.modulus != params.q()) throw ExactoError(ExactoError::ModulusMismatch, "");
        NttPoly out{std::vector<uint64_t>(p.coeffs.size()), p.modulus};
        check(exb_ntt_forward_host(params.context(), 0, p.coeffs.data(), out.evals.data(), 1));
        return out;
    }
    CoeffPoly to_coeff_poly(const BfvParams &params) const {                        // ring/ntt.rs:58-67
        CoeffPoly out{std::vector<uint64_t>(evals.size()), modulus};
        check(exb_ntt_inverse_host(params.context(), 0, evals.data(), out.coeffs.data(), 1));
        return out;
    }
};
struct RnsPoly {                                     // ring/rns.rs:14-17
    std::vector<NttPoly> components;
    size_t ring_degree;
};

// ---- bfv/ --------------------------------------------------------------------------------------------
struct BfvCiphertext {                               // bfv/mod.rs:19-24
    std::vector<RnsPoly> c;
    std::shared_ptr<BfvParams> params;
    size_t degree() const { return c.size() - 1; }
};

class RelinKey {                                     // bfv/keygen.rs:39-45
public:
    std::vector<std::pair<RnsPoly, RnsPoly>> keys;
    std::shared_ptr<BfvParams> params;
    RelinKey(std::vector<std::pair<RnsPoly, RnsPoly>> k, std::shared_ptr<BfvParams> p) : keys(std::move(k)), params(std::move(p)) {}
    RelinKey(const RelinKey &) = delete;
    ~RelinKey() { if (dev_) exb_relin_key_destroy(dev_); }
    const exb_relin_key *device() const {            // uploaded once, [G][2][n]
        if (!dev_) {
            const size_t n = params->ring_degree;
            std::vector<uint64_t> flat;
            flat.reserve(keys.size() * 2 * n);
            for (const auto &k : keys) {
                const auto &a = k.first.components.at(0).evals, &b = k.second.components.at(0).evals;
                flat.insert(flat.end(), a.begin(), a.end());
                flat.insert(flat.end(), b.begin(), b.end());
            }
            check(exb_relin_key_load(params->context(), flat.data(), (uint32_t)keys.size(), &dev_));
        }
        return dev_;
    }
private:
    mutable exb_relin_key *dev_ = nullptr;
};

class GaloisKey : public RelinKey {                  // bfv/keygen.rs:47-54: key switch s(X^element) -> s(X)
public:
    size_t element;
    GaloisKey(std::vector<std::pair<RnsPoly, RnsPoly>> k, size_t elem, std::shared_ptr<BfvParams> p)
        : RelinKey(std::move(k), std::move(p)), element(elem) {}
};

namespace detail {
inline void flatten(const BfvCiphertext &ct, std::vector<uint64_t> &out) {
    for (const RnsPoly &p : ct.c) {
        const auto &e = p.components.at(0).evals;
        out.insert(out.end(), e.begin(), e.end());
    }
}
inline BfvCiphertext unflatten(const uint64_t *src, size_t polys, const std::shared_ptr<BfvParams> &params) {
    const size_t n = params->ring_degree;
    BfvCiphertext ct{{}, params};
    for (size_t i = 0; i < polys; i++)
        ct.c.push_back(RnsPoly{{NttPoly{std::vector<uint64_t>(src + i * n, src + (i + 1) * n), params->q()}}, n});
    return ct;
}
}  // namespace detail

inline BfvCiphertext bfv_add(const BfvCiphertext &ct1, const BfvCiphertext &ct2) {   // bfv/eval.rs:14-31
    if (ct1.c.size() != 2 || ct2.c.size() != 2)
        throw ExactoError(ExactoError::NotImplemented, "C++ mirror adds degree-1 ciphertexts only");
    std::vector<uint64_t> a, b;
    detail::flatten(ct1, a); detail::flatten(ct2, b);
    exb_context *ctx = ct1.params->context();
    void *da = nullptr, *db = nullptr;
    const size_t bytes = a.size() * 8;
    check(exb_device_alloc(ctx, bytes, &da)); check(exb_device_alloc(ctx, bytes, &db));
    check(exb_copy_to_device(ctx, da, a.data(), bytes, nullptr)); check(exb_copy_to_device(ctx, db, b.data(), bytes, nullptr));
    check(exb_bfv_add(ctx, (const uint64_t *)da, (const uint64_t *)db, (uint64_t *)da, 1, nullptr));
    check(exb_copy_to_host(ctx, a.data(), da, bytes, nullptr)); check(exb_synchronize(ctx, nullptr));
    exb_device_free(ctx, da); exb_device_free(ctx, db);
    return detail::unflatten(a.data(), 2, ct1.params);
}

inline BfvCiphertext bfv_mul_and_relin(const BfvCiphertext &ct1, const BfvCiphertext &ct2, const RelinKey &rlk) {   // bfv/eval.rs:73-82
    if (ct1.c.size() != 2 || ct2.c.size() != 2)                                                                     // :93-97
        throw ExactoError(ExactoError::InvalidParam, "multiplication requires degree-1 ciphertexts");
    std::vector<uint64_t> a, b;
    detail::flatten(ct1, a); detail::flatten(ct2, b);
    std::vector<uint64_t> out(a.size());
    check(exb_bfv_mul_and_relin_host(ct1.params->context(), a.data(), b.data(), rlk.device(), out.data(), 1));
    return detail::unflatten(out.data(), 2, ct1.params);
}

namespace detail {
// host vector -> HBM -> `fn(device in..., device out)` -> host vector (device memory through the C ABI only)
template <typename F>
inline std::vector<uint64_t> device_round_trip(exb_context *ctx, const std::vector<const std::vector<uint64_t> *> &ins,
                                               size_t out_words, F fn) {
    std::vector<void *> dev(ins.size() + 1, nullptr);
    auto release = [&] { for (void *p : dev) if (p) exb_device_free(ctx, p); };
    try {
        for (size_t i = 0; i < ins.size(); i++) {
            check(exb_device_alloc(ctx, ins[i]->size() * 8, &dev[i]));
            check(exb_copy_to_device(ctx, dev[i], ins[i]->data(), ins[i]->size() * 8, nullptr));
        }
        check(exb_device_alloc(ctx, out_words * 8, &dev.back()));
        check(fn(dev));
        std::vector<uint64_t> out(out_words);
        check(exb_copy_to_host(ctx, out.data(), dev.back(), out_words * 8, nullptr));
        check(exb_synchronize(ctx, nullptr));
        release();
        return out;
    } catch (...) { release(); throw; }
}
}  // namespace detail

inline BfvCiphertext bfv_mul_no_relin(const BfvCiphertext &ct1, const BfvCiphertext &ct2) {   // bfv/eval.rs:89-108
    if (ct1.c.size() != 2 || ct2.c.size() != 2)                                               // :93-97
        throw ExactoError(ExactoError::InvalidParam, "multiplication requires degree-1 ciphertexts");
    std::vector<uint64_t> a, b;
    detail::flatten(ct1, a); detail::flatten(ct2, b);
    exb_context *ctx = ct1.params->context();
    const size_t n = ct1.params->ring_degree;
    auto out = detail::device_round_trip(ctx, {&a, &b}, 3 * n, [&](std::vector<void *> &d) {
        return exb_bfv_mul_no_relin(ctx, (const uint64_t *)d[0], (const uint64_t *)d[1], (uint64_t *)d[2], 1, nullptr);
    });
    return detail::unflatten(out.data(), 3, ct1.params);
}

inline BfvCiphertext relinearize(const BfvCiphertext &ct, const RelinKey &rlk) {              // bfv/keyswitch.rs:59-101
    if (ct.c.size() < 3) return ct;                                                           // :63-65
    if (ct.c.size() > 3)                                                                      // :66-70
        throw ExactoError(ExactoError::InvalidParam, "relinearization only supports degree-2 ciphertexts");
    std::vector<uint64_t> a;
    detail::flatten(ct, a);
    exb_context *ctx = ct.params->context();
    const size_t n = ct.params->ring_degree;
    auto out = detail::device_round_trip(ctx, {&a}, 2 * n, [&](std::vector<void *> &d) {
        return exb_bfv_relinearize(ctx, (const uint64_t *)d[0], 3, rlk.device(), (uint64_t *)d[1], 1, nullptr);
    });
    return detail::unflatten(out.data(), 2, ct.params);
}

struct SecretKey {                                   // bfv/keygen.rs:13-17: s in RNS-NTT form
    RnsPoly poly;
    std::shared_ptr<BfvParams> params;
};

// bfv/encrypt.rs:111-178: m = round(p (c0 + c1 s + c2 s^2 + ...) / q) mod p, any ciphertext degree.
inline CoeffPoly decrypt(const BfvCiphertext &ct, const SecretKey &sk) {
    std::vector<uint64_t> a;
    detail::flatten(ct, a);
    const size_t n = ct.params->ring_degree;
    std::vector<uint64_t> out(n);
    check(exb_bfv_decrypt_host(ct.params->context(), a.data(), (uint32_t)ct.c.size(), sk.poly.components.at(0).evals.data(),
                               out.data(), 1));
    return CoeffPoly{std::move(out), ct.params->plain_modulus};
}

inline BfvCiphertext bfv_apply_automorphism(const BfvCiphertext &ct, const GaloisKey &gk) {   // bfv/eval.rs:512-561
    if (ct.c.size() != 2)                                                                     // :516-520
        throw ExactoError(ExactoError::InvalidParam, "automorphism requires degree-1 ciphertext");
    std::vector<uint64_t> a;
    detail::flatten(ct, a);
    std::vector<uint64_t> out(a.size());
    check(exb_bfv_apply_automorphism_host(ct.params->context(), a.data(), gk.element, gk.device(), out.data(), 1));
    return detail::unflatten(out.data(), 2, ct.params);
}

// bfv/eval.rs:573-588: result <- result + sigma_k(result) for each k in order.
inline BfvCiphertext bfv_trace(const BfvCiphertext &ct, const std::vector<size_t> &galois_elements,
                               const std::map<size_t, std::shared_ptr<GaloisKey>> &galois_keys) {
    BfvCiphertext result = ct;
    for (size_t k : galois_elements) {
        auto it = galois_keys.find(k);
        if (it == galois_keys.end())
            throw ExactoError(ExactoError::InvalidParam, "missing Galois key for element " + std::to_string(k));
        result = bfv_add(result, bfv_apply_automorphism(result, *it->second));
    }
    return result;
}


// ---- plaintext-side operations used by the bootstrap (bfv/encrypt.rs:181-229, bfv/eval.rs:468-486, :613-634,
// bootstrap/digit_extract.rs:161-197) ---------------------------------------------------------------------
namespace detail {
inline std::vector<uint64_t> pointwise(const BfvParams &params, int op, const std::vector<uint64_t> &a,
                                       const std::vector<uint64_t> &b) {                  // op 0: add, 1: mul
    exb_context *ctx = params.context();
    return device_round_trip(ctx, {&a, &b}, a.size(), [&](std::vector<void *> &d) {
        return op == 0 ? exb_poly_add(ctx, 0, (const uint64_t *)d[0], (const uint64_t *)d[1], (uint64_t *)d[2], a.size(), nullptr)
                       : exb_poly_mul(ctx, 0, (const uint64_t *)d[0], (const uint64_t *)d[1], (uint64_t *)d[2], a.size(), nullptr);
    });
}
inline NttPoly ntt_of(const std::vector<uint64_t> &coeffs_mod_q, const BfvParams &params) {
    return NttPoly::from_coeff_poly(CoeffPoly{coeffs_mod_q, params.q()}, params);
}
inline BfvCiphertext times_ntt(const BfvCiphertext &ct, const NttPoly &m) {               // every component * m
    BfvCiphertext out{{}, ct.params};
    for (const RnsPoly &c : ct.c)
        out.c.push_back(RnsPoly{{NttPoly{pointwise(*ct.params, 1, c.components.at(0).evals, m.evals), ct.params->q()}}, c.ring_degree});
    return out;
}
}  // namespace detail

inline RnsPoly scale_plaintext(const CoeffPoly &pt, const std::shared_ptr<BfvParams> &params) {   // bfv/encrypt.rs:181-229
    const uint64_t q = params->q(), delta = q / params->plain_modulus;
    std::vector<uint64_t> c(pt.coeffs.size());
    for (size_t i = 0; i < c.size(); i++) c[i] = (uint64_t)((unsigned __int128)(pt.coeffs[i] % q) * delta % q);
    return RnsPoly{{detail::ntt_of(c, *params)}, params->ring_degree};
}
inline BfvCiphertext trivial_encrypt_poly(const CoeffPoly &pt, const std::shared_ptr<BfvParams> &params) {   // digit_extract.rs:180-189
    const size_t n = params->ring_degree;
    return BfvCiphertext{{scale_plaintext(pt, params), RnsPoly{{NttPoly{std::vector<uint64_t>(n, 0), params->q()}}, n}}, params};
}
inline BfvCiphertext trivial_encrypt(uint64_t m, const std::shared_ptr<BfvParams> &params) {       // digit_extract.rs:161-177
    CoeffPoly pt{std::vector<uint64_t>(params->ring_degree, 0), params->plain_modulus};
    pt.coeffs[0] = m % params->plain_modulus;
    return trivial_encrypt_poly(pt, params);
}
inline BfvCiphertext bfv_plain_mul(const BfvCiphertext &ct, const CoeffPoly &pt) {                 // bfv/eval.rs:468-486
    std::vector<uint64_t> c(pt.coeffs.size());
    for (size_t i = 0; i < c.size(); i++) c[i] = pt.coeffs[i] % ct.params->q();
    return detail::times_ntt(ct, detail::ntt_of(c, *ct.params));
}
inline BfvCiphertext bfv_scalar_mul(const BfvCiphertext &ct, uint64_t scalar) {                    // digit_extract.rs:192-197
    CoeffPoly pt{std::vector<uint64_t>(ct.params->ring_degree, 0), ct.params->plain_modulus};
    pt.coeffs[0] = scalar % ct.params->plain_modulus;
    return bfv_plain_mul(ct, pt);
}
inline BfvCiphertext bfv_monomial_mul(const BfvCiphertext &ct, size_t j) {                         // bfv/eval.rs:613-634
    const size_t n = ct.params->ring_degree;
    j %= 2 * n;
    if (j == 0) return ct;
    std::vector<uint64_t> c(n, 0);
    c[j % n] = j < n ? 1 : ct.params->q() - 1;                                                     // X^n = -1
    return detail::times_ntt(ct, detail::ntt_of(c, *ct.params));
}

// bootstrap/digit_extract.rs:100-157: Paterson-Stockmeyer, every product a bfv_mul_and_relin on the GPU.
inline BfvCiphertext eval_poly_homomorphic(const BfvCiphertext &ct_x, const std::vector<uint64_t> &coeffs, const RelinKey &rlk) {
    const auto &params = ct_x.params;
    const size_t d = coeffs.empty() ? 0 : coeffs.size() - 1;
    if (d == 0) return trivial_encrypt(coeffs.empty() ? 0 : coeffs[0], params);
    size_t k = (size_t)std::ceil(std::sqrt((double)d + 1.0));
    if (k < 2) k = 2;
    std::vector<BfvCiphertext> baby{trivial_encrypt(1, params), ct_x};
    for (size_t i = 2; i <= k; i++) baby.push_back(bfv_mul_and_relin(baby[i / 2], baby[i - i / 2], rlk));
    const size_t groups_n = (d + k) / k;
    std::vector<BfvCiphertext> groups;
    for (size_t i = 0; i < groups_n; i++) {
        BfvCiphertext g = trivial_encrypt(0, params);
        for (size_t j = 0; j < k; j++) {
            const size_t idx = i * k + j;
            if (idx >= coeffs.size()) break;
            if (coeffs[idx] == 0) continue;
            g = bfv_add(g, bfv_scalar_mul(baby[j], coeffs[idx]));
        }
        groups.push_back(std::move(g));
    }
    BfvCiphertext result = std::move(groups.back());
    groups.pop_back();
    while (!groups.empty()) {
        result = bfv_add(bfv_mul_and_relin(result, baby[k], rlk), groups.back());
        groups.pop_back();
    }
    return result;
}

// ---- bootstrap/coeffs_to_slots.rs ---------------------------------------------------------------------------------
using GaloisKeys = std::map<size_t, std::shared_ptr<GaloisKey>>;
inline std::vector<size_t> required_trace_elements(size_t n) {                                     // :167-181
    std::vector<size_t> e;
    if (n <= 32 || (n & (n - 1))) { for (size_t k = 3; k < 2 * n; k += 2) e.push_back(k); return e; }
    for (size_t step = n; step >= 2; step >>= 1) e.push_back(step + 1);
    return e;
}
inline BfvCiphertext extract_coefficient(const BfvCiphertext &ct, size_t j, const GaloisKeys &gks) {   // :21-95
    const size_t n = ct.params->ring_degree;
    const uint64_t t = ct.params->plain_modulus;
    const BfvCiphertext shifted = j == 0 ? ct : bfv_monomial_mul(ct, 2 * n - j);
    auto key = [&](size_t k) -> const GaloisKey & {
        auto it = gks.find(k);
        if (it == gks.end()) throw ExactoError(ExactoError::InvalidParam, "missing Galois key for element " + std::to_string(k));
        return *it->second;
    };
    BfvCiphertext result = shifted;
    if (n <= 32 || (n & (n - 1))) {                                                                // naive trace
        for (size_t k = 3; k < 2 * n; k += 2) result = bfv_add(result, bfv_apply_automorphism(shifted, key(k)));
    } else {
        for (size_t k : required_trace_elements(n)) result = bfv_add(result, bfv_apply_automorphism(result, key(k)));
    }
    uint64_t n_inv = 0;                                                                            // n^-1 mod t
    for (uint64_t x = 1; x < t; x++) if ((unsigned __int128)(n % t) * x % t == 1) { n_inv = x; break; }
    if (!n_inv) throw ExactoError(ExactoError::InvalidParam, "n not invertible mod t");
    return bfv_scalar_mul(result, n_inv);
}
inline std::vector<BfvCiphertext> coeffs_to_slots(const BfvCiphertext &ct, const GaloisKeys &gks) {    // :103-116
    std::vector<BfvCiphertext> out;
    for (size_t j = 0; j < ct.params->ring_degree; j++) out.push_back(extract_coefficient(ct, j, gks));
    return out;
}
inline BfvCiphertext slots_to_coeffs(const std::vector<BfvCiphertext> &slots) {                    // :122-142
    if (slots.empty()) throw ExactoError(ExactoError::InvalidParam, "empty slots");
    BfvCiphertext result = slots[0];
    for (size_t j = 1; j < slots.size(); j++) result = bfv_add(result, bfv_monomial_mul(slots[j], j));
    return result;
}

// ---- bootstrap/bfv_host.rs ------------------------------------------------------------------------------------------
struct BootstrapKey {                                // :19-41
    std::shared_ptr<BfvParams> boot_params;
    uint64_t q_prime = 0;
    BfvCiphertext bsk;                               // encryption of s under the boot scheme
    std::shared_ptr<RelinKey> boot_rlk;
    std::vector<uint64_t> rounding_poly;             // coefficients mod t_boot
    GaloisKeys galois_keys;
};

inline BfvCiphertext bfv_bootstrap(const BfvCiphertext &ct, const BootstrapKey &bsk) {             // :134-209
    if (ct.c.size() != 2) throw ExactoError(ExactoError::InvalidParam, "bootstrap requires degree-1 ciphertext");
    const uint64_t q = ct.params->q(), qp = bsk.q_prime, tb = bsk.boot_params->plain_modulus;
    const size_t n = ct.params->ring_degree;
    const CoeffPoly c0 = ct.c[0].components.at(0).to_coeff_poly(*ct.params), c1 = ct.c[1].components.at(0).to_coeff_poly(*ct.params);
    CoeffPoly p0{std::vector<uint64_t>(n), tb}, p1{std::vector<uint64_t>(n), tb};
    bool trivial = true;
    for (size_t i = 0; i < n; i++) {                                                               // :150-171
        p0.coeffs[i] = (uint64_t)(((unsigned __int128)qp * c0.coeffs[i] + q / 2) / q) % qp % tb;
        p1.coeffs[i] = (uint64_t)(((unsigned __int128)qp * c1.coeffs[i] + q / 2) / q) % qp % tb;
        trivial = trivial && c1.coeffs[i] == 0;
    }
    const BfvCiphertext phase = bfv_add(trivial_encrypt_poly(p0, bsk.boot_params), bfv_plain_mul(bsk.bsk, p1));   // :173-176
    if (trivial) return eval_poly_homomorphic(phase, bsk.rounding_poly, *bsk.boot_rlk);            // :179-186
    std::vector<BfvCiphertext> rounded;
    for (const BfvCiphertext &slot : coeffs_to_slots(phase, bsk.galois_keys))                      // :189-194
        rounded.push_back(eval_poly_homomorphic(slot, bsk.rounding_poly, *bsk.boot_rlk));
    return slots_to_coeffs(rounded);                                                               // :197-201
}

// ---- dbfv/ -----------------------------------------------------------------------------------------------
struct DbfvCiphertext {                              // dbfv/ciphertext.rs:10-22
    std::vector<BfvCiphertext> limbs;
    size_t degree;
    size_t mul_depth;
    std::shared_ptr<DbfvParams> params;
    size_t num_limbs() const { return limbs.size(); }
    bool needs_reduction() const { return degree > params->num_digits; }
};

inline DbfvCiphertext dbfv_add(const DbfvCiphertext &ct1, const DbfvCiphertext &ct2) {   // dbfv/eval.rs:11-33
    if (ct1.num_limbs() != ct2.num_limbs())
        throw ExactoError(ExactoError::DimensionMismatch, "expected " + std::to_string(ct1.num_limbs()) + ", got " + std::to_string(ct2.num_limbs()));
    DbfvCiphertext out{{}, std::max(ct1.degree, ct2.degree), std::max(ct1.mul_depth, ct2.mul_depth), ct1.params};
    for (size_t i = 0; i < ct1.num_limbs(); i++) out.limbs.push_back(bfv_add(ct1.limbs[i], ct2.limbs[i]));
    return out;
}

inline DbfvCiphertext dbfv_mul(const DbfvCiphertext &ct1, const DbfvCiphertext &ct2, const RelinKey &rlk) {   // dbfv/eval.rs:82-149
    const auto &params = ct1.params;
    const size_t d = params->num_digits;
    if (ct1.num_limbs() != d || ct2.num_limbs() != d)                                                         // :90-94
        throw ExactoError(ExactoError::InvalidParam, "multiplication requires d-limb ciphertexts");
    const size_t next_depth = std::max(ct1.mul_depth, ct2.mul_depth) + 1;
    if (next_depth > 1)                                                                                       // :96-102
        throw ExactoError(ExactoError::NotImplemented,
                          "chained dBFV multiplication requires ciphertext-level lattice reduction (paper §4.6.2)");
    std::vector<uint64_t> a, b;
    for (const auto &l : ct1.limbs) {
        if (l.c.size() != 2) throw ExactoError(ExactoError::InvalidParam, "multiplication requires degree-1 ciphertexts");
        detail::flatten(l, a);
    }
    for (const auto &l : ct2.limbs) {
        if (l.c.size() != 2) throw ExactoError(ExactoError::InvalidParam, "multiplication requires degree-1 ciphertexts");
        detail::flatten(l, b);
    }
    std::vector<uint64_t> out(a.size());
    check(exb_dbfv_mul_host(params->bfv_params->context(), params->base, (uint32_t)d, params->plain_modulus, a.data(), b.data(),
                            rlk.device(), out.data(), 1, 0));       // the d^2 products, per-k sums and reduce() in one batched sequence
    DbfvCiphertext res{{}, d, next_depth, params};                   // :138-146, reduction.rs:54-59
    const size_t n = params->bfv_params->ring_degree;
    for (size_t i = 0; i < d; i++) res.limbs.push_back(detail::unflatten(out.data() + i * 2 * n, 2, params->bfv_params));
    return res;
}

// dbfv/advanced.rs:15-30: the BFV automorphism + key switch on every limb, one batched call.
inline DbfvCiphertext dbfv_apply_automorphism(const DbfvCiphertext &ct, const GaloisKey &gk) {
    std::vector<uint64_t> a;
    for (const auto &l : ct.limbs) {
        if (l.c.size() != 2) throw ExactoError(ExactoError::InvalidParam, "automorphism requires degree-1 ciphertext");
        detail::flatten(l, a);
    }
    std::vector<uint64_t> out(a.size());
    const auto &bfv = ct.params->bfv_params;
    check(exb_bfv_apply_automorphism_host(bfv->context(), a.data(), gk.element, gk.device(), out.data(), ct.num_limbs()));
    DbfvCiphertext res{{}, ct.degree, ct.mul_depth, ct.params};
    const size_t n = bfv->ring_degree;
    for (size_t i = 0; i < ct.num_limbs(); i++) res.limbs.push_back(detail::unflatten(out.data() + i * 2 * n, 2, bfv));
    return res;
}

// bootstrap/bfv_host.rs:212-236: every limb refreshed; dBFV metadata kept, BFV params swapped, mul_depth restarted.
inline DbfvCiphertext dbfv_bootstrap(const DbfvCiphertext &ct, const BootstrapKey &bsk) {
    auto refreshed = DbfvParams::create(bsk.boot_params, ct.params->base, ct.params->num_digits, ct.params->plain_modulus);
    DbfvCiphertext out{{}, ct.degree, 0, refreshed};
    for (const BfvCiphertext &limb : ct.limbs) out.limbs.push_back(bfv_bootstrap(limb, bsk));
    return out;
}

// bootstrap/bfv_host.rs:242-250
inline DbfvCiphertext dbfv_mul_then_bootstrap(const DbfvCiphertext &ct1, const DbfvCiphertext &ct2, const RelinKey &rlk,
                                              const BootstrapKey &bsk) {
    return dbfv_bootstrap(dbfv_mul(ct1, ct2, rlk), bsk);
}

// bootstrap/bfv_host.rs:258-288: fold with the relinearisation key chosen by parameter equality (:271-284).
inline DbfvCiphertext dbfv_mul_chain_then_bootstrap(const std::vector<DbfvCiphertext> &cts, const RelinKey &rlk,
                                                    const BootstrapKey &bsk) {
    if (cts.empty())
        throw ExactoError(ExactoError::InvalidParam, "dbfv_mul_chain_then_bootstrap requires at least one ciphertext");
    auto same = [](const BfvParams &a, const BfvParams &b) {
        return a.plain_modulus == b.plain_modulus && a.ring_degree == b.ring_degree && a.ct_moduli == b.ct_moduli;
    };
    DbfvCiphertext acc = cts[0];
    for (size_t i = 1; i < cts.size(); i++) {
        const bool use_boot_rlk = same(*acc.params->bfv_params, *bsk.boot_params);
        const DbfvCiphertext rhs = same(*acc.params->bfv_params, *cts[i].params->bfv_params) ? cts[i] : dbfv_bootstrap(cts[i], bsk);
        acc = dbfv_mul_then_bootstrap(acc, rhs, use_boot_rlk ? *bsk.boot_rlk : rlk, bsk);
    }
    return acc;
}

}  // namespace exacto
