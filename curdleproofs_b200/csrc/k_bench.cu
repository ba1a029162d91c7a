// Micro-benchmarks that give the integer-pipe roofline its denominators (MEASURED_PEAKS.json only holds HBM and bf16):
//   k_bench_imad   independent IMAD.WIDE.U32 chains, no memory traffic      -> peak 32x32->64 multiply-adds per second
//   k_bench_fpmul  dependent chain of Montgomery products per thread        -> achieved Fp multiplications per second
#include "launch.h"
#include "fp381.cuh"

namespace cdp {

__global__ void __launch_bounds__(256) k_bench_imad(uint32_t *out, uint32_t seed, int iters) {
    uint32_t x = seed + threadIdx.x, y = seed ^ (blockIdx.x * 2654435761u);
    uint32_t a0 = x, a1 = y, b0 = y + 1, b1 = x + 1, c0 = x ^ y, c1 = 7, d0 = 11, d1 = x * 3;
    uint32_t e0 = x + 5, e1 = y + 9, f0 = y ^ 3, f1 = x ^ 9, g0 = 13, g1 = y * 5, h0 = 17, h1 = x * 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {  // 8 independent 64-bit accumulators, 128 wide multiply-adds per iteration
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a0), "+r"(a1) : "r"(x), "r"(y));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(b0), "+r"(b1) : "r"(y), "r"(x));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(c0), "+r"(c1) : "r"(x), "r"(x));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(d0), "+r"(d1) : "r"(y), "r"(y));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(e0), "+r"(e1) : "r"(x), "r"(y));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(f0), "+r"(f1) : "r"(y), "r"(x));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(g0), "+r"(g1) : "r"(x), "r"(x));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(h0), "+r"(h1) : "r"(y), "r"(y));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ b0 ^ b1 ^ c0 ^ c1 ^ d0 ^ d1 ^ e0 ^ e1 ^ f0 ^ f1 ^ g0 ^ g1 ^ h0 ^ h1;
}

__global__ void __launch_bounds__(256) k_bench_fpmul(uint32_t *out, uint32_t seed, int iters, int sqr) {
    fp x, y;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        x.v[i] = seed * (i + 1) + threadIdx.x;
        y.v[i] = (seed ^ 0x9e3779b9u) * (i + 3) + blockIdx.x;
    }
    x.v[11] &= 0x0fffffffu;
    y.v[11] &= 0x0fffffffu;
    if (sqr) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) fp_sqr(x, x);
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            fp z;
            fp_mul(z, x, y);
            x = y;
            y = z;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= x.v[i] ^ y.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// FP64-pipe probes: k_bench_dfma = 128 independent DFMAs per iteration; mix = 1 adds the 128 IMAD.WIDE of k_bench_imad to the same
// loop body (do the two pipes overlap?), mix = 2 adds 128 64-bit integer additions instead (ALU pipe).
template <int mix>
__global__ void __launch_bounds__(256) k_bench_dfma(uint32_t *out, uint32_t seed, int iters) {
    double x = 1.0 + 1e-9 * (double)(seed + threadIdx.x), y = 1.0 - 1e-9 * (double)(blockIdx.x + 1);
    double a = x, b = y, c = x + y, d = x - y, e = x * 3, f = y * 5, g = x * 7, h = y * 9;
    uint32_t ux = seed + threadIdx.x, uy = seed ^ (blockIdx.x * 2654435761u);
    uint32_t a0 = ux, a1 = uy, b0 = uy + 1, b1 = ux + 1, c0 = ux ^ uy, c1 = 7, d0 = 11, d1 = ux * 3;
    uint32_t e0 = ux + 5, e1 = uy + 9, f0 = uy ^ 3, f1 = ux ^ 9, g0 = 13, g1 = uy * 5, h0 = 17, h1 = ux * 7;
    unsigned long long s0 = ux, s1 = uy, s2 = ux * 3ull, s3 = uy * 5ull, s4 = 1, s5 = 2, s6 = 3, s7 = 4;
    const unsigned long long inc = ((unsigned long long)uy << 32) | ux;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a = fma(a, x, y); b = fma(b, y, x); c = fma(c, x, x); d = fma(d, y, y);
            e = fma(e, x, y); f = fma(f, y, x); g = fma(g, x, x); h = fma(h, y, y);
            if (mix == 1) {
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a0), "+r"(a1) : "r"(ux), "r"(uy));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(b0), "+r"(b1) : "r"(uy), "r"(ux));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(c0), "+r"(c1) : "r"(ux), "r"(ux));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(d0), "+r"(d1) : "r"(uy), "r"(uy));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(e0), "+r"(e1) : "r"(ux), "r"(uy));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(f0), "+r"(f1) : "r"(uy), "r"(ux));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(g0), "+r"(g1) : "r"(ux), "r"(ux));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(h0), "+r"(h1) : "r"(uy), "r"(uy));
            } else if (mix == 2) {
                asm volatile("add.u64 %0, %0, %1;" : "+l"(s0) : "l"(inc)); asm volatile("add.u64 %0, %0, %1;" : "+l"(s1) : "l"(inc));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(s2) : "l"(inc)); asm volatile("add.u64 %0, %0, %1;" : "+l"(s3) : "l"(inc));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(s4) : "l"(inc)); asm volatile("add.u64 %0, %0, %1;" : "+l"(s5) : "l"(inc));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(s6) : "l"(inc)); asm volatile("add.u64 %0, %0, %1;" : "+l"(s7) : "l"(inc));
            }
        }
    }
    double t = a + b + c + d + e + f + g + h;
    unsigned long long u = s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2hiint(t) ^ (uint32_t)__double2loint(t) ^ a0 ^ a1 ^ b0 ^ b1 ^ c0 ^ c1 ^ d0 ^ d1 ^
                                                 e0 ^ e1 ^ f0 ^ f1 ^ g0 ^ g1 ^ h0 ^ h1 ^ (uint32_t)u ^ (uint32_t)(u >> 32);
}

cudaError_t launch_bench(cudaStream_t st, int which, uint32_t *out, int blocks, int threads, int iters) {
    if (which == 0) k_bench_imad<<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 3) k_bench_dfma<0><<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 4) k_bench_dfma<1><<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 5) k_bench_dfma<2><<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else k_bench_fpmul<<<blocks, threads, 0, st>>>(out, 12345u, iters, which == 2);
    return cudaGetLastError();
}

}  // namespace cdp
