// tsan_kernels.cpp -- TEST TOOLING: the product's kernels (tests/host_emul: kernels.cu under the CUDA shim, one OS
// thread per CUDA thread, pthread barriers for __syncthreads) run under ThreadSanitizer.  compute-sanitizer's
// racecheck is closed on this GPU pool; a shared-memory access pair that is not ordered by a barrier shows up here
// as a data race between the two OS threads.  Build + run: tools/run_tsan.sh
#include "../tests/host_emul/emul.cpp"

#include <cstdio>

static uint64_t lcg = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(uint64_t m) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (lcg >> 1) % m; }
static std::vector<uint64_t> rand_vec(size_t words, uint64_t m) { std::vector<uint64_t> v(words); for (auto &x : v) x = rnd(m); return v; }

static int run_set(const char *name, uint32_t n, uint64_t q, std::vector<uint64_t> aux, uint64_t p, uint64_t gb, uint64_t base, uint32_t d, uint64_t pm) {
    emu_ctx *c = nullptr;
    int rc = emu_create(n, &q, 1, aux.data(), (uint32_t)aux.size(), p, gb, 0, &c);
    if (rc) { std::printf("%s: create failed %d %s\n", name, rc, emu_last_error()); return 1; }
    const uint32_t G = c->hs.gadget_digits;
    auto x = rand_vec(3 * (size_t)n, q), y = x;
    emu_ntt(c, 0, 1, x.data(), y.data(), 3);
    emu_ntt(c, 0, 0, y.data(), x.data(), 3);
    emu_ntt(c, 0, 1, x.data(), y.data(), 2);                 // even count: the other register tiling at n = 4096
    auto ct1 = rand_vec((size_t)d * 2 * n, q), ct2 = rand_vec((size_t)d * 2 * n, q), key = rand_vec((size_t)G * 2 * n, q);
    std::vector<uint64_t> out((size_t)d * 2 * n);
    for (uint32_t flags : {0u, 0x80000000u, 0x40000000u, 0xC0000000u, 1u}) {
        rc = emu_dbfv_mul(c, base, d, pm, ct1.data(), ct2.data(), key.data(), G, out.data(), 1, flags, 0);
        if (rc) { std::printf("%s: dbfv_mul flags %x failed %d %s\n", name, flags, rc, emu_last_error()); return 1; }
    }
    std::vector<uint64_t> out3(3 * (size_t)n), out2(2 * (size_t)n), dec(n);
    emu_bfv_mul_no_relin(c, ct1.data(), ct2.data(), out3.data(), 1);
    emu_bfv_relinearize(c, out3.data(), key.data(), G, out2.data(), 1, 0);
    emu_bfv_relinearize(c, out3.data(), key.data(), G, out2.data(), 1, 1);
    emu_bfv_apply_automorphism(c, ct1.data(), 3, key.data(), out2.data(), 1);
    emu_bfv_decrypt(c, out3.data(), 3, key.data(), dec.data(), 1);
    std::vector<uint64_t> digs((size_t)G * n);
    emu_gadget_decompose(c, x.data(), digs.data(), 1);
    emu_destroy(c);
    std::printf("%s: ran\n", name);
    return 0;
}

// `--selftest`: a deliberately unsynchronised neighbour read; ThreadSanitizer must report it (proves the set-up
// sees shared-memory races between emulated CUDA threads).
static void selftest() {
    emu_launch(1, 64, 64 * 8, [&]() {
        exb::u64 *sm = (exb::u64 *)emu_smem_ptr;
        sm[threadIdx.x] = threadIdx.x;
        volatile exb::u64 v = sm[threadIdx.x ^ 1u];          // missing __syncthreads()
        (void)v;
    });
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--selftest") { selftest(); return 0; }
    int bad = 0;
    bad += run_set("u64 profile (internal 27-bit basis)", 4096, 1152921504606830593ull, {18014398509998081ull, 36028797018972161ull}, 1040407, 256, 256, 2, 65536);
    bad += run_set("n=4096 reference-basis path (q < 2^36)", 4096, 2147377153ull, {18014398509998081ull, 36028797018972161ull}, 257, 65536, 16, 2, 256);
    bad += run_set("compact (n=1024, one aux prime)", 1024, 1099509805057ull, {562949953443841ull}, 929, 65536, 16, 2, 256);
    bad += run_set("n=32 generic, odd log2 n, no aux", 32, 1125899906842817ull, {}, 29, 8, 4, 2, 16);
    return bad;
}
