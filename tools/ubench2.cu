// ubench2.cu -- clean per-instruction issue-cost measurement on sm_100a (tools, not product).
// Each kernel's loop body is one asm block of 16 independent instances of one instruction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 2048

#define R16(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)

template <int OP>
__global__ void __launch_bounds__(256) k(u32 *out, long long *cycles, u32 seed) {
    u32 a[16]; u64 w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = seed * (threadIdx.x + 3) + i * 77; w[i] = (u64)a[i] * 0x9E3779B97F4A7C15ull; }
    u32 m = seed * threadIdx.x | 1, m2 = (seed + threadIdx.x) * 3;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(m2));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(m), "r"(m2));
            if (OP == 3) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, hi, %1, %0;}" : "+l"(w[i]) : "r"(m));
            if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            if (OP == 5) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; add.cc.u32 lo, lo, %1; addc.u32 hi, hi, %2; mov.b64 %0, {lo, hi};}" : "+l"(w[i]) : "r"(m), "r"(m2));
            if (OP == 6) asm volatile("{.reg .pred p; .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; setp.ge.u32 p, hi, %1; @p sub.u32 hi, hi, %1; mov.b64 %0, {lo, hi};}" : "+l"(w[i]) : "r"(m));
            if (OP == 7) asm volatile("{.reg .pred p; setp.ge.u64 p, %0, %1; @p sub.u64 %0, %0, %1;}" : "+l"(w[i]) : "l"((u64)m << 29));
            if (OP == 8) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a[i]) : "r"(m));
            if (OP == 9) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            if (OP == 10) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= a[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, double ops_per_iter, u32 *out, long long *cyc) {
    for (int bps = 2; bps <= 4; bps += 2) {
        int blocks = 148 * bps;
        k<OP><<<blocks, 256>>>(out, cyc, 12345); cudaDeviceSynchronize();
        k<OP><<<blocks, 256>>>(out, cyc, 12345); cudaDeviceSynchronize();
        static long long h[148 * 8];
        cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
        double warp_ops_per_smsp = (double)ITERS * 16 * ops_per_iter * (256 / 32) * bps / 4;
        printf("%-30s blocks/SM=%d: %.2f SMSP-cycles per warp-op  (%.1f thread-op/clk/SM)\n", name, bps, avg / warp_ops_per_smsp,
               32.0 * 4 * warp_ops_per_smsp / avg);
    }
}

int main() {
    u32 *out; long long *cyc;
    cudaMalloc(&out, 4 * 256 * 148 * 8); cudaMalloc(&cyc, 8 * 148 * 8);
    run<0>("mad.lo.u32", 1, out, cyc);
    run<10>("mul.lo.u32", 1, out, cyc);
    run<1>("mad.hi.u32", 1, out, cyc);
    run<9>("mul.hi.u32", 1, out, cyc);
    run<2>("mad.wide.u32 (acc chain)", 1, out, cyc);
    run<3>("mad.wide.u32 (dep on hi)", 1, out, cyc);
    run<4>("add.u32", 1, out, cyc);
    run<5>("add.cc+addc (u64 add)", 1, out, cyc);
    run<6>("setp+@p sub (32-bit csub)", 1, out, cyc);
    run<7>("setp.u64+@p sub.u64 (csub64)", 1, out, cyc);
    run<8>("shf.l.wrap", 1, out, cyc);
    return 0;
}
