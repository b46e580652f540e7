// host_mirror_driver.cpp -- exercises the C++ host mirror (include/exacto_b200.hpp) on the GPU.
// Usage: host_mirror_driver <preset: compact_dbfv|u64_dbfv> <in.bin> <out.bin>
//   in.bin : ct1 [d][2][n] | ct2 [d][2][n] | rlk [G][2][n]   (u64, NTT domain)
//   out.bin: dbfv_mul limbs [d][2][n] | bfv_mul_and_relin(limb0, limb0) [2][n] | dbfv_add [d][2][n]
//            | NTT round trip of ct1 limb 0 comp 0 [n] | dbfv_apply_automorphism(ct1, sigma_3 with rlk as key) [d][2][n]
//            | bfv_trace(ct1 limb 0, {3}) [2][n] | decrypt(ct1 limb 0, sk = rlk[0][1]) [n]
// Prints "guards ok" after checking the reference's error behaviour (dbfv/eval.rs:90-102, bfv/eval.rs:93-97).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../../include/exacto_b200.hpp"

using namespace exacto;

static std::vector<uint64_t> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    std::vector<uint64_t> v((size_t)f.tellg() / 8);
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * 8));
    return v;
}

static BfvCiphertext make_ct(const uint64_t *src, size_t polys, const std::shared_ptr<BfvParams> &p) {
    return detail::unflatten(src, polys, p);
}

template <typename F>
static bool throws(ExactoError::Kind kind, const char *needle, F f) {
    try { f(); } catch (const ExactoError &e) { return e.kind == kind && std::string(e.what()).find(needle) != std::string::npos; }
    return false;
}

int main(int argc, char **argv) {
    if (argc != 4) { std::fprintf(stderr, "usage: %s preset in.bin out.bin\n", argv[0]); return 2; }
    try {
        std::shared_ptr<DbfvParams> params = std::strcmp(argv[1], "u64_dbfv") == 0 ? u64_dbfv() : compact_dbfv();
        const auto &bfv = params->bfv_params;
        const size_t n = bfv->ring_degree, d = params->num_digits, G = bfv->gadget_digits;
        const std::vector<uint64_t> in = slurp(argv[2]);
        if (in.size() != (2 * d * 2 + G * 2) * n) { std::fprintf(stderr, "bad input size\n"); return 3; }
        const uint64_t *p1 = in.data(), *p2 = p1 + d * 2 * n, *pk = p2 + d * 2 * n;
        DbfvCiphertext a{{}, d, 0, params}, b{{}, d, 0, params};
        for (size_t i = 0; i < d; i++) { a.limbs.push_back(make_ct(p1 + i * 2 * n, 2, bfv)); b.limbs.push_back(make_ct(p2 + i * 2 * n, 2, bfv)); }
        std::vector<std::pair<RnsPoly, RnsPoly>> keys;
        for (size_t g = 0; g < G; g++) {
            BfvCiphertext k = make_ct(pk + g * 2 * n, 2, bfv);
            keys.emplace_back(k.c[0], k.c[1]);
        }
        auto gk = std::make_shared<GaloisKey>(keys, 3, bfv);              // same [G][2][n] words as a Galois key
        RelinKey rlk(std::move(keys), bfv);

        const DbfvCiphertext prod = dbfv_mul(a, b, rlk);
        const BfvCiphertext one = bfv_mul_and_relin(a.limbs[0], b.limbs[0], rlk);
        const DbfvCiphertext sum = dbfv_add(a, b);
        const CoeffPoly back = NttPoly::from_coeff_poly(a.limbs[0].c[0].components[0].to_coeff_poly(*bfv), *bfv).to_coeff_poly(*bfv);
        const CoeffPoly ref = a.limbs[0].c[0].components[0].to_coeff_poly(*bfv);

        bool ok = prod.degree == d && prod.mul_depth == 1 && prod.num_limbs() == d && one.c.size() == 2;
        ok = ok && back.coeffs == ref.coeffs;
        ok = ok && throws(ExactoError::NotImplemented, "chained dBFV multiplication requires ciphertext-level lattice reduction",
                          [&] { dbfv_mul(prod, b, rlk); });
        DbfvCiphertext shorty = a; shorty.limbs.pop_back();
        ok = ok && throws(ExactoError::InvalidParam, "multiplication requires d-limb ciphertexts", [&] { dbfv_mul(shorty, b, rlk); });
        BfvCiphertext deg2 = a.limbs[0]; deg2.c.push_back(deg2.c[0]);
        ok = ok && throws(ExactoError::InvalidParam, "multiplication requires degree-1 ciphertexts",
                          [&] { bfv_mul_and_relin(deg2, b.limbs[0], rlk); });
        ok = ok && throws(ExactoError::InvalidParam, "cannot create NTT plan",
                          [&] { BfvParamsBuilder().ring_degree(4096).ct_moduli({0xFFFFFFFFFFE00001ull}).build(); });
        ok = ok && throws(ExactoError::NotImplemented, "schoolbook BFV multiplication can overflow i128", [&] {
                 auto p = BfvParamsBuilder().ring_degree(4096).plain_modulus(1040407).ct_moduli({18014398509506561ull}).gadget_base(256).build();
                 std::vector<std::pair<RnsPoly, RnsPoly>> none;
                 RelinKey empty(std::move(none), p);
                 BfvCiphertext z = detail::unflatten(std::vector<uint64_t>(2 * 4096, 0).data(), 2, p);
                 bfv_mul_and_relin(z, z, empty);
             });
        const BfvCiphertext two_step = relinearize(bfv_mul_no_relin(a.limbs[0], b.limbs[0]), rlk);   // == bfv_mul_and_relin
        ok = ok && two_step.c.size() == 2 && two_step.c[0].components[0].evals == one.c[0].components[0].evals &&
             two_step.c[1].components[0].evals == one.c[1].components[0].evals;
        ok = ok && relinearize(a.limbs[0], rlk).c.size() == 2;
        const DbfvCiphertext rot = dbfv_apply_automorphism(a, *gk);
        const BfvCiphertext tr = bfv_trace(a.limbs[0], {3}, {{3, gk}});
        ok = ok && rot.degree == a.degree && rot.mul_depth == a.mul_depth && rot.num_limbs() == d;
        ok = ok && throws(ExactoError::InvalidParam, "automorphism requires degree-1 ciphertext", [&] { bfv_apply_automorphism(deg2, *gk); });
        ok = ok && throws(ExactoError::InvalidParam, "missing Galois key for element 5", [&] { bfv_trace(a.limbs[0], {5}, {{3, gk}}); });
        std::puts(ok ? "guards ok" : "guards FAILED");

        std::ofstream out(argv[3], std::ios::binary);
        auto dump = [&](const BfvCiphertext &ct) { for (const auto &c : ct.c) out.write(reinterpret_cast<const char *>(c.components[0].evals.data()), (std::streamsize)(n * 8)); };
        for (const auto &l : prod.limbs) dump(l);
        dump(one);
        for (const auto &l : sum.limbs) dump(l);
        out.write(reinterpret_cast<const char *>(back.coeffs.data()), (std::streamsize)(n * 8));
        for (const auto &l : rot.limbs) dump(l);
        dump(tr);
        const SecretKey sk{make_ct(pk + n, 1, bfv).c[0], bfv};            // any NTT-domain polynomial serves as s here
        const CoeffPoly dec = decrypt(a.limbs[0], sk);
        out.write(reinterpret_cast<const char *>(dec.coeffs.data()), (std::streamsize)(n * 8));
        return ok ? 0 : 4;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
