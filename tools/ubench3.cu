// ubench3.cu -- FP64 / FP32 FMA issue cost on sm_100a and co-issue with IMAD (tools, not product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(256) k(u32 *out, long long *cycles, u32 seed) {
    u32 a[16]; double d[16]; float f[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = seed * (threadIdx.x + 3) + i * 77; d[i] = (double)a[i] * 1e-9; f[i] = (float)a[i] * 1e-9f; }
    u32 m = seed * threadIdx.x | 1, m2 = (seed + threadIdx.x) * 3;
    double dm = 1.0 + 1e-9 * threadIdx.x, dm2 = 1e-7 * threadIdx.x;
    float fm = 1.0f + 1e-6f * threadIdx.x, fm2 = 1e-5f * threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (OP == 0 || OP == 2 || OP == 5) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(dm2));
            if (OP == 1 || OP == 2 || OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 3 || OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fm), "f"(fm2));
            if (OP == 5) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 6) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dm2));
            if (OP == 7) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dm));
            if (OP == 8) asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(d[i]));
            if (OP == 9) asm volatile("{.reg .f64 t; cvt.rn.f64.u32 t, %1; add.rn.f64 %0, %0, t;}" : "+d"(d[i]) : "r"(a[i]));
            if (OP == 10) asm volatile("{.reg .s32 t; cvt.rni.s32.f64 t, %0; add.u32 %1, %1, t;}" : "+d"(d[i]), "+r"(a[i]));
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= a[i] ^ (u32)__double_as_longlong(d[i]) ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, u32 *out, long long *cyc) {
    for (int bps = 2; bps <= 4; bps += 2) {
        int blocks = 148 * bps;
        k<OP><<<blocks, 256>>>(out, cyc, 12345); cudaDeviceSynchronize();
        k<OP><<<blocks, 256>>>(out, cyc, 12345); cudaDeviceSynchronize();
        static long long h[148 * 8];
        cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
        double groups_per_smsp = (double)ITERS * 16 * (256 / 32) * bps / 4;
        printf("%-44s blocks/SM=%d: %.2f SMSP-cycles per warp-group\n", name, bps, avg / groups_per_smsp);
    }
}

int main() {
    u32 *out; long long *cyc;
    cudaMalloc(&out, 4 * 256 * 148 * 8); cudaMalloc(&cyc, 8 * 148 * 8);
    run<0>("fma.f64", out, cyc);
    run<6>("add.f64", out, cyc);
    run<7>("mul.f64", out, cyc);
    run<8>("cvt.rni.f64.f64", out, cyc);
    run<9>("cvt.f64.u32 + add.f64", out, cyc);
    run<10>("cvt.rni.s32.f64 + add.u32", out, cyc);
    run<1>("mad.lo.u32", out, cyc);
    run<2>("fma.f64 + mad.lo.u32 (group of 2)", out, cyc);
    run<5>("fma.f64 + mad.hi.u32 (group of 2)", out, cyc);
    run<3>("fma.f32", out, cyc);
    run<4>("fma.f32 + mad.lo.u32 (group of 2)", out, cyc);
    return 0;
}
