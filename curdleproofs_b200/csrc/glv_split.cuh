// GLV split of a canonical scalar k < r < 2^255:  k = k2 * lambda + k1 (plain Euclidean division by lambda = z^2 - 1 ~ 2^127.4), so
// k*P = k1*P + k2*phi(P) with phi(x, y) = (beta*x, y) and k1, k2 < 2^128, both non-negative -- no lattice rounding, no signs.
// Barrett: q' = floor(k mu / 2^256) with mu = floor(2^256 / lambda) is floor(k / lambda) or one less; one conditional correction.
// ~50 wide multiply-adds (the bit-serial division this replaces was ~9 000 instructions per scalar: 1 ms of the 2^22-pair MSM's digit kernel).
// Plain C++ so that the same source is checked on the CPU (tests/host/glv_check.cpp).
#pragma once
#include <stdint.h>

namespace cdp {

#ifndef __CUDACC__
#define CDP_GLV_FN inline
#else
#define CDP_GLV_FN __device__ __forceinline__
#endif

struct glv_t {
    uint32_t k1[4], k2[4];
};

CDP_GLV_FN void glv_split(glv_t &g, const uint32_t k[8]) {
    const uint32_t MU[4] = {0xf6cfee30u, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u};  // mu = 2^128 + MU
    const uint32_t LAM[4] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u};
    uint32_t acc[13];
#pragma unroll
    for (int i = 0; i < 13; i++) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t t = (uint64_t)k[i] * MU[j] + acc[i + j] + carry;
            acc[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        const uint64_t t = (uint64_t)k[i] + acc[i + 4] + carry;  // the 2^128 term of mu
        acc[i + 4] = (uint32_t)t;
        acc[i + 5] += (uint32_t)(t >> 32);  // acc[i + 5] is still zero here: no carry out
    }
    uint32_t q[4] = {acc[8], acc[9], acc[10], acc[11]};
    // r = k - q lambda, 5 limbs (r < 2 lambda < 2^129)
    uint32_t pr[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i + j < 5) {
                const uint64_t t = (uint64_t)q[i] * LAM[j] + pr[i + j] + carry;
                pr[i + j] = (uint32_t)t;
                carry = t >> 32;
            }
        }
        if (i + 4 < 5) pr[i + 4] += (uint32_t)carry;
    }
    uint32_t r[5];
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint64_t d = (uint64_t)k[i] - pr[i] - borrow;
        r[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 32) & 1u;
    }
    // r >= lambda: one more lambda fits
    uint32_t t[5];
    borrow = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint64_t d = (uint64_t)r[i] - (i < 4 ? LAM[i] : 0u) - borrow;
        t[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 32) & 1u;
    }
    const bool ge = borrow == 0;
    uint32_t inc = ge ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        g.k1[i] = ge ? t[i] : r[i];
        const uint32_t s = q[i] + inc;
        inc = (s < q[i]) ? 1u : 0u;
        g.k2[i] = s;
    }
}

}  // namespace cdp
