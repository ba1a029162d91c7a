// Micro-benchmarks that give the integer-pipe roofline its denominators (MEASURED_PEAKS.json only holds HBM and bf16):
//   k_bench_imad   independent IMAD.WIDE.U32 chains, no memory traffic      -> peak 32x32->64 multiply-adds per second
//   k_bench_fpmul  dependent chain of Montgomery products per thread        -> achieved Fp multiplications per second
#include "launch.h"
#include "fp381.cuh"
#include "g1.cuh"

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

// Mixed-addition chains (acc += P, `iters` times per thread) at the occupancy of the production kernels (128 threads, 3 CTAs / SM):
//   which = 6  g1j_add_mixed as the kernels use it: field products called with by-value register arguments (ptxas marshals them with
//              IMAD.MOV.U32, i.e. on the fmaheavy pipe that the products themselves saturate)
//   which = 7  the same formula with the products called through POINTERS to local-memory operands: the marshalling becomes LDL / STL
//              (load/store pipe) instead
static __device__ __noinline__ void fp_mul_pfn(fp *r, const fp *a, const fp *b) {
    fp z;
    fp_mul_eo(z, *a, *b);
    *r = z;
}
static __device__ __noinline__ void fp_sqr_pfn(fp *r, const fp *a) {
    fp z;
    fp_sqr_rw(z, *a);
    *r = z;
}
__device__ __forceinline__ void g1j_add_mixed_ptr(g1j &r, const g1j &p, const g1a &q) {
    fp Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;
    fp_sqr_pfn(&Z1Z1, &p.Z);
    fp_mul_pfn(&U2, &q.x, &Z1Z1);
    fp_mul_pfn(&S2, &q.y, &p.Z);
    fp_mul_pfn(&S2, &S2, &Z1Z1);
    fp_sub(H, U2, p.X);
    fp_sub(rr, S2, p.Y);
    fp_dbl(rr, rr);
    fp_sqr_pfn(&HH, &H);
    fp_dbl(I, HH);
    fp_dbl(I, I);
    fp_mul_pfn(&J, &H, &I);
    fp_mul_pfn(&V, &p.X, &I);
    fp_sqr_pfn(&X3, &rr);
    fp_sub(X3, X3, J);
    fp_sub(X3, X3, V);
    fp_sub(X3, X3, V);
    fp_sub(t, V, X3);
    fp_mul_pfn(&Y3, &rr, &t);
    fp_mul_pfn(&t, &p.Y, &J);
    fp_dbl(t, t);
    fp_sub(Y3, Y3, t);
    fp_add(Z3, p.Z, H);
    fp_sqr_pfn(&Z3, &Z3);
    fp_sub(Z3, Z3, Z1Z1);
    fp_sub(Z3, Z3, HH);
    r.X = X3;
    r.Y = Y3;
    r.Z = Z3;
}
//   which = 8  the accumulator in XYZZ coordinates (X, Y, ZZ, ZZZ; madd-2008-s: 8M + 2S instead of 7M + 4S, 12 more live registers)
template <int MODE>
__global__ void __launch_bounds__(128, 3) k_bench_madd(uint32_t *out, uint32_t seed, int iters) {
    g1j acc;
    g1a P;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        acc.X.v[i] = seed * (i + 1) + threadIdx.x;
        acc.Y.v[i] = (seed ^ 0x9e3779b9u) * (i + 3) + blockIdx.x;
        acc.Z.v[i] = seed + 77u * i + threadIdx.x * 3u;
        P.x.v[i] = (seed ^ 0x85ebca6bu) * (i + 5) + threadIdx.x;
        P.y.v[i] = (seed ^ 0xc2b2ae35u) * (i + 7) + blockIdx.x;
    }
    acc.X.v[11] &= 0x0fffffffu; acc.Y.v[11] &= 0x0fffffffu; acc.Z.v[11] &= 0x0fffffffu; P.x.v[11] &= 0x0fffffffu; P.y.v[11] &= 0x0fffffffu;
    if (MODE == 2) {
        g1x ax;
        ax.X = acc.X; ax.Y = acc.Y; ax.ZZ = acc.Z; ax.ZZZ = P.x;
#pragma unroll 1
        for (int i = 0; i < iters; i++) g1x_add_mixed(ax, ax, P);
        g1x_to_jac(acc, ax);
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            if (MODE == 0) g1j_add_mixed(acc, acc, P);
            else g1j_add_mixed_ptr(acc, acc, P);
        }
    }
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) a ^= acc.X.v[i] ^ acc.Y.v[i] ^ acc.Z.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

cudaError_t launch_bench(cudaStream_t st, int which, uint32_t *out, int blocks, int threads, int iters) {
    if (which == 0) k_bench_imad<<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 3) k_bench_dfma<0><<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 4) k_bench_dfma<1><<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 5) k_bench_dfma<2><<<blocks, threads, 0, st>>>(out, 12345u, iters);
    else if (which == 6) k_bench_madd<0><<<blocks, 128, 0, st>>>(out, 12345u, iters);
    else if (which == 7) k_bench_madd<1><<<blocks, 128, 0, st>>>(out, 12345u, iters);
    else if (which == 8) k_bench_madd<2><<<blocks, 128, 0, st>>>(out, 12345u, iters);
    else k_bench_fpmul<<<blocks, threads, 0, st>>>(out, 12345u, iters, which == 2);
    return cudaGetLastError();
}

}  // namespace cdp
