// ubench.cu -- integer-pipe micro-benchmark for sm_100a (tools, not product).
// Measures per-SM per-clock throughput (thread-instructions / clock / SM) of the
// instructions the 64-bit modular butterflies are made of, alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8   // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint64_t *out64, uint32_t seed, long long *cycles) {
    uint32_t a[CH], b[CH];
    uint64_t w[CH];
    double dd[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 11 + 1 + threadIdx.x * 13; w[i] = ((uint64_t)a[i] << 32) | b[i]; dd[i] = 1.0 + a[i] * 1e-9; }
    const uint32_t m = seed | 1, m2 = seed * 5 + 3;
    const double dm = 1.0000001, da = 1e-7;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 6) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a[i]) : "r"(m), "r"(b[i]));
            if (OP == 7) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
            if (OP == 8) {  // wide + add interleaved (do they overlap?)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            }
            if (OP == 9) {  // mad.lo + dfma interleaved
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
            }
            if (OP == 10) {  // mad.wide + 2 adds
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m2));
            }
            if (OP == 11) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)m << 20 | m2));
            if (OP == 12) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)m << 20 | m2));
            if (OP == 13) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(m2));
            if (OP == 14) {  // mad.lo + mad.hi pair (alternative to wide)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(a[i]), "r"(m));
            }
            if (OP == 15) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"(a[i]), "r"(m), "l"(w[i]));  // same, keeps dependency
            if (OP == 16) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"(m));        // wide without addend
            if (OP == 17) asm volatile("min.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[i] - m));
            if (OP == 18) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&a[i]) : "f"(1.0001f), "f"(0.5f));
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0; uint64_t acc64 = 0; double accd = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) { acc += a[i] + b[i]; acc64 += w[i]; accd += dd[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)accd;
    out64[blockIdx.x * blockDim.x + threadIdx.x] = acc64;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter, uint32_t *out, uint64_t *out64, long long *cyc, int blocks) {
    k<OP><<<blocks, 1024>>>(out, out64, 12345, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 1024>>>(out, out64, 12345, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double inst = (double)ITERS * CH * per_iter * 1024;           // thread-instructions per block (1 block/SM at 1024 thr... 2 may co-reside)
    cudaError_t e = cudaGetLastError();
    printf("%-34s %8.2f thread-inst/clk/SM (per resident block: %.2f)  %.3f ms %s\n", name, inst * blocks / 148.0 / avg / ((blocks + 147) / 148), inst / avg, ms, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    int blocks = 148;
    uint32_t *out; uint64_t *out64; long long *cyc;
    cudaMalloc(&out, 4 * 1024 * 2048); cudaMalloc(&out64, 8 * 1024 * 2048); cudaMalloc(&cyc, 8 * 2048);
    run<0>("mad.lo.u32 (IMAD)", 1, out, out64, cyc, blocks);
    run<1>("mad.hi.u32 (IMAD.HI)", 1, out, out64, cyc, blocks);
    run<2>("mad.wide.u32 acc (IMAD.WIDE)", 1, out, out64, cyc, blocks);
    run<15>("mad.wide.u32 (dep addend)", 1, out, out64, cyc, blocks);
    run<16>("mul.wide.u32", 1, out, out64, cyc, blocks);
    run<3>("add.u32 (IADD3)", 1, out, out64, cyc, blocks);
    run<4>("lop3", 1, out, out64, cyc, blocks);
    run<5>("shf", 1, out, out64, cyc, blocks);
    run<6>("setp+selp", 2, out, out64, cyc, blocks);
    run<13>("add.cc+addc (64-bit add)", 2, out, out64, cyc, blocks);
    run<17>("min.u64 (+sub)", 1, out, out64, cyc, blocks);
    run<7>("fma.f64 (DFMA)", 1, out, out64, cyc, blocks);
    run<18>("fma.f32 (FFMA)", 1, out, out64, cyc, blocks);
    run<8>("wide + add", 2, out, out64, cyc, blocks);
    run<10>("wide + 2 add", 3, out, out64, cyc, blocks);
    run<9>("mad.lo + dfma", 2, out, out64, cyc, blocks);
    run<14>("mad.lo + mad.hi", 2, out, out64, cyc, blocks);
    run<11>("mul.hi.u64", 1, out, out64, cyc, blocks);
    run<12>("mul.lo.u64", 1, out, out64, cyc, blocks);
    return 0;
}
