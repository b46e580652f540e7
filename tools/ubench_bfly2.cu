// ubench_bfly2.cu -- register-resident butterfly throughput of the product's own butterflies
// (ntt_core.cuh) vs values-per-thread and resident warps (tools, not product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../exacto_b200/csrc/ntt_core.cuh"
using namespace exb;
#define ITERS 1024

template <int NV, int LAZY, bool INV, int MINB>
__global__ void __launch_bounds__(256, MINB) k(u64 *out, const Tw *tw, Modulus mod, long long *cycles) {
    u64 v[NV];
    const LazyC c = make_lazyc(mod);
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = (u64)(threadIdx.x * 977 + i * 13 + 1) * 0x9E3779B97F4A7C15ull % mod.m;
    Tw t[4];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = tw[(threadIdx.x + i) & 63];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        // log2(NV) stages like a real register pass (stage j pairs elements half = NV >> (j+1) apart)
#pragma unroll
        for (int half = NV / 2; half >= 1; half >>= 1) {
#pragma unroll
            for (int i = 0; i < NV; i++) {
                if ((i & half) == 0) {
                    if (INV) gs_bfly_l<LAZY>(v[i], v[i + half], t[(i + half) & 3], c, c.four_q << 1);
                    else ct_bfly_l<LAZY>(v[i], v[i + half], t[(i + half) & 3], c);
                }
            }
        }
        if (LAZY == 2 || INV) {   // keep values bounded like the real transforms do
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = reduce_to_2m(v[i], c.neg_q, c.rhi, c.rsh);
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < NV; i++) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NV, int LAZY, bool INV, int MINB>
void run(const char *name, int bps, u64 *out, Tw *tw, Modulus mod, long long *cyc) {
    int blocks = 148 * bps;
    k<NV, LAZY, INV, MINB><<<blocks, 256>>>(out, tw, mod, cyc); cudaDeviceSynchronize();
    k<NV, LAZY, INV, MINB><<<blocks, 256>>>(out, tw, mod, cyc); cudaDeviceSynchronize();
    static long long h[148 * 8];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    int stages = 0; for (int x = NV; x > 1; x >>= 1) stages++;
    double bf = (double)ITERS * (NV / 2) * stages * 256 * bps;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<NV, LAZY, INV, MINB>);
    printf("%-28s NV=%2d regs=%3d blocks/SM=%d : %.3f bfly/clk/SM (%.1f%% of roofline rate 8.47) %s\n", name, NV, fa.numRegs, bps,
           bf / avg, 100.0 * bf / avg / 8.47, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    u64 *out; Tw *tw; long long *cyc;
    cudaMalloc(&out, 8 * 256 * 148 * 8); cudaMalloc(&tw, sizeof(Tw) * 64); cudaMalloc(&cyc, 8 * 148 * 8);
    const u64 q = 18014398509998081ull;   // 54-bit aux prime (LAZY 2 capable)
    Tw h[64];
    for (int i = 0; i < 64; i++) { h[i].w = 0x123456789abcdefull * (i + 1) % q; h[i].s = (u64)(((unsigned __int128)h[i].w << 64) / q); }
    cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
    Modulus m{}; m.m = q; m.two_m = 2 * q; m.neg_m = 0 - q; m.four_m = 4 * q; m.hi_four_m = (u32)((4 * q) >> 32);
    m.rhi = (u32)(((unsigned __int128)1 << (32 + 53)) / q); m.rsh = 53 - 32; m.lazy = 2;
    run<16, 2, false, 1>("fwd lazy2", 1, out, tw, m, cyc);
    run<16, 2, false, 2>("fwd lazy2", 2, out, tw, m, cyc);
    run<16, 2, false, 3>("fwd lazy2", 3, out, tw, m, cyc);
    run<16, 2, false, 4>("fwd lazy2", 4, out, tw, m, cyc);
    run<8, 2, false, 2>("fwd lazy2", 2, out, tw, m, cyc);
    run<8, 2, false, 4>("fwd lazy2", 4, out, tw, m, cyc);
    run<8, 2, false, 6>("fwd lazy2", 6, out, tw, m, cyc);
    run<8, 2, false, 8>("fwd lazy2", 8, out, tw, m, cyc);
    run<16, 1, false, 2>("fwd lazy1", 2, out, tw, m, cyc);
    run<16, 1, false, 4>("fwd lazy1", 4, out, tw, m, cyc);
    run<16, 0, false, 2>("fwd lazy0 (Harvey exact)", 2, out, tw, m, cyc);
    run<16, 0, false, 4>("fwd lazy0 (Harvey exact)", 4, out, tw, m, cyc);
    run<16, 2, true, 2>("inv lazy2", 2, out, tw, m, cyc);
    run<16, 2, true, 4>("inv lazy2", 4, out, tw, m, cyc);
    run<16, 1, true, 2>("inv lazy1", 2, out, tw, m, cyc);
    run<8, 2, true, 4>("inv lazy2", 4, out, tw, m, cyc);
    run<8, 2, true, 8>("inv lazy2", 8, out, tw, m, cyc);
    return 0;
}
