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

cudaError_t launch_bench(cudaStream_t st, int which, uint32_t *out, int blocks, int threads, int iters) {
    if (which == 0) k_bench_imad<<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else k_bench_fpmul<<<blocks, threads, 0, st>>>(out, 12345u, iters, which == 2);
    return cudaGetLastError();
}

}  // namespace cdp
