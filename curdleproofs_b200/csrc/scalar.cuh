// Scalar recoding helpers shared by the kernels (GLV split, endomorphism constant).
#pragma once
#include "g1.cuh"

namespace cdp {

// ------------------------------------------------------------------------------------------------ scalars
// GLV split of a canonical scalar k < r < 2^255:  k = k2 * lambda + k1 (plain Euclidean division by
// lambda = z^2 - 1 ~ 2^127.4), so k*P = k1*P + k2*phi(P) with phi(x, y) = (beta*x, y) and k1, k2 < 2^128, both
// non-negative -- no lattice rounding, no signs.  Bit-serial restoring division: ~4k simple integer instructions,
// <1% of the scalar multiplication it feeds.
struct glv_t {
    uint32_t k1[4], k2[4];
};
__device__ __forceinline__ void glv_split(glv_t &g, const uint32_t kin[8]) {
    const uint32_t L0 = FR_LAMBDA[0], L1 = FR_LAMBDA[1], L2 = FR_LAMBDA[2], L3 = FR_LAMBDA[3];
    uint32_t k0 = kin[0], k1 = kin[1], k2 = kin[2], k3 = kin[3], k4 = kin[4], k5 = kin[5], k6 = kin[6], k7 = kin[7];
    // bit 255 of k is always clear (k < r < 2^255): pre-shift so that the loop consumes bits 254..0 from the top
    k7 = (k7 << 1) | (k6 >> 31); k6 = (k6 << 1) | (k5 >> 31); k5 = (k5 << 1) | (k4 >> 31); k4 = (k4 << 1) | (k3 >> 31);
    k3 = (k3 << 1) | (k2 >> 31); k2 = (k2 << 1) | (k1 >> 31); k1 = (k1 << 1) | (k0 >> 31); k0 <<= 1;
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;  // remainder (< 2*lambda < 2^129)
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;          // quotient < 2^128: bits shifted out of the top are zero
#pragma unroll 1
    for (int i = 254; i >= 0; i--) {
        uint32_t bit = k7 >> 31;
        k7 = (k7 << 1) | (k6 >> 31); k6 = (k6 << 1) | (k5 >> 31); k5 = (k5 << 1) | (k4 >> 31); k4 = (k4 << 1) | (k3 >> 31);
        k3 = (k3 << 1) | (k2 >> 31); k2 = (k2 << 1) | (k1 >> 31); k1 = (k1 << 1) | (k0 >> 31); k0 <<= 1;
        r4 = (r4 << 1) | (r3 >> 31);
        r3 = (r3 << 1) | (r2 >> 31);
        r2 = (r2 << 1) | (r1 >> 31);
        r1 = (r1 << 1) | (r0 >> 31);
        r0 = (r0 << 1) | bit;
        uint32_t t0, t1, t2, t3, t4, borrow;
        asm("sub.cc.u32 %0, %6, %11;\n\t"
            "subc.cc.u32 %1, %7, %12;\n\t"
            "subc.cc.u32 %2, %8, %13;\n\t"
            "subc.cc.u32 %3, %9, %14;\n\t"
            "subc.cc.u32 %4, %10, 0;\n\t"
            "subc.u32 %5, 0, 0;"
            : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(borrow)
            : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(L0), "r"(L1), "r"(L2), "r"(L3));
        uint32_t qbit = borrow == 0;
        if (qbit) { r0 = t0; r1 = t1; r2 = t2; r3 = t3; r4 = t4; }
        q3 = (q3 << 1) | (q2 >> 31); q2 = (q2 << 1) | (q1 >> 31); q1 = (q1 << 1) | (q0 >> 31); q0 = (q0 << 1) | qbit;
    }
    g.k1[0] = r0; g.k1[1] = r1; g.k1[2] = r2; g.k1[3] = r3;
    g.k2[0] = q0; g.k2[1] = q1; g.k2[2] = q2; g.k2[3] = q3;
}

__device__ __forceinline__ void fp_mul_beta(fp &r, const fp &a) {
    fp beta;
#pragma unroll
    for (int i = 0; i < 12; i++) beta.v[i] = FP_BETA_MONT[i];
    fp_mul(r, a, beta);
}

}  // namespace cdp
