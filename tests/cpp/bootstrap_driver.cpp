// bootstrap_driver.cpp -- exercises dbfv_mul_then_bootstrap / dbfv_mul_chain_then_bootstrap of the C++ host mirror
// (include/exacto_b200.hpp, bootstrap/bfv_host.rs:242-288) on the GPU at the reference's toy scale.
// Usage: bootstrap_driver <in.bin> <out.bin>      (u64 words, NTT domain)
//   in : n q p gb | qb tb gbb | q_prime base d dbfv_p | R NG NC
//        rounding_poly[R] | rlk[G][2][n] | bsk[2][n] | boot_rlk[Gb][2][n] | NG x (element, key[Gb][2][n]) | NC x ct[d][2][n]
//   out: dbfv_mul_then_bootstrap(ct0, ct1) [d][2][n] | dbfv_mul_chain_then_bootstrap(ct0..) [d][2][n]
//        | dbfv_mul(refreshed, dbfv_bootstrap(ct2), boot_rlk) [d][2][n]
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../include/exacto_b200.hpp"

using namespace exacto;

int main(int argc, char **argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
        std::vector<uint64_t> in((size_t)f.tellg() / 8);
        f.seekg(0);
        f.read(reinterpret_cast<char *>(in.data()), (std::streamsize)(in.size() * 8));
        const uint64_t *w = in.data();
        const size_t n = (size_t)w[0];
        auto orig = BfvParamsBuilder().ring_degree(n).plain_modulus(w[2]).ct_moduli({w[1]}).gadget_base(w[3]).build();
        auto boot = BfvParamsBuilder().ring_degree(n).plain_modulus(w[5]).ct_moduli({w[4]}).gadget_base(w[6]).build();
        const uint64_t q_prime = w[7], base = w[8];
        const size_t d = (size_t)w[9];
        auto params = DbfvParams::create(orig, base, d, w[10]);
        const size_t R = (size_t)w[11], NG = (size_t)w[12], NC = (size_t)w[13];
        w += 14;
        BootstrapKey bsk;
        bsk.boot_params = boot;
        bsk.q_prime = q_prime;
        bsk.rounding_poly.assign(w, w + R); w += R;
        auto key_pairs = [&](size_t G, const std::shared_ptr<BfvParams> &p) {
            std::vector<std::pair<RnsPoly, RnsPoly>> keys;
            for (size_t g = 0; g < G; g++) {
                BfvCiphertext k = detail::unflatten(w, 2, p);
                keys.emplace_back(k.c[0], k.c[1]);
                w += 2 * n;
            }
            return keys;
        };
        RelinKey rlk(key_pairs(orig->gadget_digits, orig), orig);
        bsk.bsk = detail::unflatten(w, 2, boot); w += 2 * n;
        bsk.boot_rlk = std::make_shared<RelinKey>(key_pairs(boot->gadget_digits, boot), boot);
        for (size_t i = 0; i < NG; i++) {
            const size_t element = (size_t)*w++;
            bsk.galois_keys[element] = std::make_shared<GaloisKey>(key_pairs(boot->gadget_digits, boot), element, boot);
        }
        std::vector<DbfvCiphertext> cts;
        for (size_t c = 0; c < NC; c++) {
            DbfvCiphertext ct{{}, d, 0, params};
            for (size_t i = 0; i < d; i++) { ct.limbs.push_back(detail::unflatten(w, 2, orig)); w += 2 * n; }
            cts.push_back(std::move(ct));
        }
        if ((size_t)(w - in.data()) != in.size()) { std::fprintf(stderr, "bad input size\n"); return 3; }

        const DbfvCiphertext refreshed = dbfv_mul_then_bootstrap(cts[0], cts[1], rlk, bsk);
        const DbfvCiphertext chained = dbfv_mul_chain_then_bootstrap(cts, rlk, bsk);
        const DbfvCiphertext again = dbfv_mul(refreshed, dbfv_bootstrap(cts[2], bsk), *bsk.boot_rlk);
        bool ok = refreshed.mul_depth == 0 && refreshed.degree == d && refreshed.params->bfv_params == boot &&
                  chained.mul_depth == 0 && chained.params->bfv_params->plain_modulus == boot->plain_modulus && again.mul_depth == 1;
        try { dbfv_mul_chain_then_bootstrap({}, rlk, bsk); ok = false; }
        catch (const ExactoError &e) { ok = ok && e.kind == ExactoError::InvalidParam; }
        BfvCiphertext deg2 = cts[0].limbs[0]; deg2.c.push_back(deg2.c[0]);
        try { bfv_bootstrap(deg2, bsk); ok = false; }
        catch (const ExactoError &e) { ok = ok && std::string(e.what()).find("bootstrap requires degree-1 ciphertext") != std::string::npos; }
        std::puts(ok ? "metadata ok" : "metadata FAILED");

        std::ofstream out(argv[2], std::ios::binary);
        for (const DbfvCiphertext *r : {&refreshed, &chained, &again})
            for (const auto &l : r->limbs)
                for (const auto &c : l.c) out.write(reinterpret_cast<const char *>(c.components[0].evals.data()), (std::streamsize)(n * 8));
        return ok ? 0 : 4;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
