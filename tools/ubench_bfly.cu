// ubench_bfly.cu -- butterfly-level micro-benchmark (tools, not product): how many 64-bit
// Shoup butterflies per clock per SM can sm_100a sustain for different instruction selections?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 2048
#define NB 8   // butterflies in flight per thread (16 values)

struct TwC { u64 w, s; };

// V0: straightforward C++ (what kernels.cu does today)
__device__ __forceinline__ void bfly_v0(u64 &x, u64 &y, u64 w, u64 s, u64 q, u64 q2) {
    u64 X = x >= q2 ? x - q2 : x;
    u64 T = y * w - __umul64hi(y, s) * q;
    x = X + T; y = X - T + q2;
}
// V1: hand-scheduled PTX. Q = approx hi64(y*s) (1 wide + 2 hi), X' = X + y*w + Q*nq in one IMAD chain,
//     Y' = X + X + c*q - X'.  nq = 2^64 - q.  T < 4q  (quotient error <= 2).
template <bool CSUB>
__device__ __forceinline__ void bfly_v1(u64 &x, u64 &y, u64 w, u64 s, u64 nq, u64 cq, u64 q2) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    u32 w0 = (u32)w, w1 = (u32)(w >> 32), s0 = (u32)s, s1 = (u32)(s >> 32);
    u32 n0 = (u32)nq, n1 = (u32)(nq >> 32), c0 = (u32)cq, c1 = (u32)(cq >> 32);
    if (CSUB) {
        u64 X = x >= q2 ? x - q2 : x;
        x0 = (u32)X; x1 = (u32)(X >> 32);
    }
    u32 q0, q1, t1, t2, r0, r1, z0, z1;
    asm("{\n\t"
        ".reg .u64 qq, acc;\n\t"
        "mul.wide.u32 qq, %8, %12;\n\t"          // y1*s1
        "mov.b64 {%0, %1}, qq;\n\t"
        "mul.hi.u32 %2, %8, %11;\n\t"            // hi(y1*s0)
        "mul.hi.u32 %3, %7, %12;\n\t"            // hi(y0*s1)
        "add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "add.cc.u32 %0, %0, %3;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "mov.b64 acc, {%5, %6};\n\t"             // acc = X
        "mad.wide.u32 acc, %7, %9, acc;\n\t"     // + y0*w0
        "mov.b64 {%4, %2}, acc;\n\t"             // r0 = lo, t1 = hi
        "mad.lo.u32 %2, %7, %10, %2;\n\t"        // + y0*w1
        "mad.lo.u32 %2, %8, %9, %2;\n\t"         // + y1*w0
        "mov.b64 acc, {%4, %2};\n\t"
        "mad.wide.u32 acc, %0, %13, acc;\n\t"    // + q0*n0
        "mov.b64 {%4, %2}, acc;\n\t"
        "mad.lo.u32 %2, %0, %14, %2;\n\t"        // + q0*n1
        "mad.lo.u32 %2, %1, %13, %2;\n\t"        // + q1*n0
        "}\n"
        : "=&r"(q0), "=&r"(q1), "=&r"(t1), "=&r"(t2), "=&r"(r0)
        : "r"(x0), "r"(x1), "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(s0), "r"(s1), "r"(n0), "r"(n1));
    r1 = t1;
    // Y' = (X + cq) + X - X'
    asm("{\n\t"
        ".reg .u32 a0, a1;\n\t"
        "add.cc.u32 a0, %2, %4;\n\t"
        "addc.u32 a1, %3, %5;\n\t"
        "add.cc.u32 a0, a0, %2;\n\t"
        "addc.u32 a1, a1, %3;\n\t"
        "sub.cc.u32 %0, a0, %6;\n\t"
        "subc.u32 %1, a1, %7;\n\t"
        "}\n"
        : "=r"(z0), "=r"(z1) : "r"(x0), "r"(x1), "r"(c0), "r"(c1), "r"(r0), "r"(r1));
    x = ((u64)r1 << 32) | r0;
    y = ((u64)z1 << 32) | z0;
}

template <int V>
__global__ void __launch_bounds__(256) k(u64 *out, const TwC *tw, u64 q, long long *cycles) {
    u64 v[2 * NB];
#pragma unroll
    for (int i = 0; i < 2 * NB; i++) v[i] = (u64)(threadIdx.x * 977 + i * 13 + 1) * 0x9E3779B97F4A7C15ull % q;
    const u64 q2 = 2 * q, nq = 0 - q, cq = 4 * q;
    TwC t[4];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = tw[(threadIdx.x + i) & 63];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NB; i++) {
            TwC c = t[i & 3];
            if (V == 0) bfly_v0(v[i], v[i + NB], c.w, c.s, q, q2);
            if (V == 1) bfly_v1<true>(v[i], v[i + NB], c.w, c.s, nq, cq, q2);
            if (V == 2) bfly_v1<false>(v[i], v[i + NB], c.w, c.s, nq, cq, q2);
        }
        // rotate roles so values stay "live" like consecutive stages (no extra instructions: register renaming)
        u64 tmp = v[0];
#pragma unroll
        for (int i = 0; i < 2 * NB - 1; i++) v[i] = v[i + 1];
        v[2 * NB - 1] = tmp;
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 2 * NB; i++) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char *name, int blocks_per_sm, u64 *out, TwC *tw, long long *cyc) {
    int blocks = 148 * blocks_per_sm;
    k<V><<<blocks, 256>>>(out, tw, 1152921504606830593ull, cyc);
    cudaDeviceSynchronize();
    k<V><<<blocks, 256>>>(out, tw, 1152921504606830593ull, cyc);
    cudaDeviceSynchronize();
    static long long h[148 * 8];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double bf = (double)ITERS * NB * 256 * blocks_per_sm;   // butterflies per SM
    printf("%-40s blocks/SM=%d : %.3f butterflies/clk/SM  (%.1f%% of the 8.47 needed for 100%% HBM roofline) %s\n", name,
           blocks_per_sm, bf / avg, 100.0 * (bf / avg) / 8.47, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    u64 *out; TwC *tw; long long *cyc;
    cudaMalloc(&out, 8 * 256 * 148 * 8); cudaMalloc(&tw, sizeof(TwC) * 64); cudaMalloc(&cyc, 8 * 148 * 8);
    TwC h[64];
    for (int i = 0; i < 64; i++) { h[i].w = 0x123456789abcdefull * (i + 1) % 1152921504606830593ull; h[i].s = (u64)(((unsigned __int128)h[i].w << 64) / 1152921504606830593ull); }
    cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
    for (int b = 1; b <= 4; b++) {
        run<0>("V0 C++ (umul64hi, csub)", b, out, tw, cyc);
        run<1>("V1 PTX approx-hi, fused X', csub", b, out, tw, cyc);
        run<2>("V2 PTX approx-hi, fused X', lazy", b, out, tw, cyc);
    }
    return 0;
}
