// Fp inversion by the Bernstein-Yang "safegcd" division steps (Fast constant-time gcd computation and modular inversion, TCHES 2019),
// variable-time form, in plain C++ so that the same source is checked on the CPU (tests/host/fp_inv_check.cpp).
//
// A division step on (delta, f, g), f odd:   delta > 0 and g odd: (1 - delta, g, (g - f) / 2)     otherwise: (1 + delta, f, (g + (g mod 2) f) / 2).
// Started from (1, p, a) it reaches g = 0 with f = +-gcd = +-1 after at most 1 101 steps for 381-bit inputs.  The steps only look at the low
// bits, so 30 of them are run on the low words of f and g alone, collecting a 2x2 integer transition matrix t (entries up to 2^30 in
// magnitude); t is then applied once to the full-length (f, g) -- an exact division by 2^30 -- and to (d, e) modulo p, which keep
// d a = f, e a = g (mod p).  Numbers are 13 signed limbs of 30 bits, so a row of the matrix product is 32 x 32 -> 64-bit multiply-adds.
// ~25 batches of ~500 instructions against ~600 steps of ~250 for the binary Euclid of fp_inv_euclid.cuh: what makes one inversion cheap
// enough to be shared by as few as 16 batched affine additions (batch_affine.cuh).  The same code serves the scalar field (mod_fr: 9 limbs,
// the per-round challenge inversions of the device-side prover).
#pragma once
#include <stdint.h>

namespace cdp {
namespace safegcd {

#ifndef __CUDACC__
#define CDP_SAFEGCD_FN inline
#define CDP_SAFEGCD_CTZ(x) __builtin_ctz(x)
#else
#define CDP_SAFEGCD_FN __device__ __forceinline__
#define CDP_SAFEGCD_CTZ(x) (__ffs((int)(x)) - 1)
#endif

constexpr int32_t M30 = 0x3fffffff;
// the modulus: NL limbs of 30 bits (room for values in (-2m, m)), NW 32-bit words, m^-1 mod 2^30
struct mod_fp {  // BLS12-381 base field
    static constexpr int NL = 13, NW = 12;
    static constexpr uint32_t INV30 = 0x30003u;
    static CDP_SAFEGCD_FN int32_t limb(int i) {
        const int32_t t[13] = {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673,
                               0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111};
        return t[i];
    }
};
struct mod_fr {  // BLS12-381 scalar field
    static constexpr int NL = 9, NW = 8;
    static constexpr uint32_t INV30 = 0x1u;
    static CDP_SAFEGCD_FN int32_t limb(int i) {
        const int32_t t[9] = {0x00000001, 0x3ffffffc, 0x3fe5bfef, 0x2f6900bf, 0x21d80553, 0x27602026, 0x17d48333, 0x29d4ca67, 0x000073ed};
        return t[i];
    }
};

struct mat {
    int32_t u, v, q, r;
};

// up to 30 division steps on the low words; returns the new eta = -delta
CDP_SAFEGCD_FN int32_t divsteps30(int32_t eta, uint32_t f, uint32_t g, mat &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;
    int i = 30;
    for (;;) {
        // g even: halve it (f's row doubles instead); all pending halvings at once
        const int zeros = CDP_SAFEGCD_CTZ(g | (0xffffffffu << i));
        g >>= zeros;
        u <<= zeros;
        v <<= zeros;
        eta -= zeros;
        i -= zeros;
        if (i == 0) break;
        if (eta < 0) {  // delta > 0, g odd: swap, (f, g) <- (g, -f)
            eta = -eta;
            uint32_t x = f; f = g; g = 0u - x;
            x = u; u = q; q = 0u - x;
            x = v; v = r; r = 0u - x;
        }
        g += f;  // both odd: g becomes even
        q += u;
        r += v;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return eta;
}

// (f, g) <- t (f, g) / 2^30, exact
template <class MOD>
CDP_SAFEGCD_FN void update_fg(int32_t *f, int32_t *g, const mat &t) {
    constexpr int NL = MOD::NL;
    int64_t cf = (int64_t)t.u * f[0] + (int64_t)t.v * g[0], cg = (int64_t)t.q * f[0] + (int64_t)t.r * g[0];
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < NL; i++) {
        const int32_t fi = f[i], gi = g[i];
        cf += (int64_t)t.u * fi + (int64_t)t.v * gi;
        cg += (int64_t)t.q * fi + (int64_t)t.r * gi;
        f[i - 1] = (int32_t)cf & M30;
        g[i - 1] = (int32_t)cg & M30;
        cf >>= 30;
        cg >>= 30;
    }
    f[NL - 1] = (int32_t)cf;
    g[NL - 1] = (int32_t)cg;
}

// (d, e) <- t (d, e) / 2^30 mod p; both stay in (-2p, p)
template <class MOD>
CDP_SAFEGCD_FN void update_de(int32_t *d, int32_t *e, const mat &t) {
    constexpr int NL = MOD::NL;
    const int32_t sd = d[NL - 1] >> 31, se = e[NL - 1] >> 31;
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);  // + p for a negative input
    int64_t cd = (int64_t)t.u * d[0] + (int64_t)t.v * e[0], ce = (int64_t)t.q * d[0] + (int64_t)t.r * e[0];
    // the multiple of p that clears the low 30 bits
    md -= (int32_t)((MOD::INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((MOD::INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)MOD::limb(0) * md;
    ce += (int64_t)MOD::limb(0) * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < NL; i++) {
        const int32_t di = d[i], ei = e[i];
        cd += (int64_t)t.u * di + (int64_t)t.v * ei + (int64_t)MOD::limb(i) * md;
        ce += (int64_t)t.q * di + (int64_t)t.r * ei + (int64_t)MOD::limb(i) * me;
        d[i - 1] = (int32_t)cd & M30;
        e[i - 1] = (int32_t)ce & M30;
        cd >>= 30;
        ce >>= 30;
    }
    d[NL - 1] = (int32_t)cd;
    e[NL - 1] = (int32_t)ce;
}

// out = a^-1 mod m as an integer (a in [0, m), MOD::NW 32-bit limbs); a = 0 gives 0.  `all_done` as in euclid::inverse_int.
template <class MOD, class AllDone>
CDP_SAFEGCD_FN void inverse_mod(uint32_t *out, const uint32_t *a, AllDone all_done) {
    constexpr int NL = MOD::NL, NW = MOD::NW;
    int32_t f[NL], g[NL], d[NL], e[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        // bits [30 i, 30 i + 30) of a
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint32_t x = w < NW ? a[w] >> sh : 0u;
        if (sh > 2 && w + 1 < NW) x |= a[w + 1] << (32 - sh);
        g[i] = (int32_t)(x & (uint32_t)M30);
        f[i] = MOD::limb(i);
        d[i] = 0;
        e[i] = 0;
    }
    e[0] = 1;
    int32_t eta = -1;
#pragma unroll 1
    for (int it = 0; it < 40; it++) {
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < NL; i++) nz |= (uint32_t)g[i];
        if (all_done(nz == 0)) break;
        mat t;
        eta = divsteps30(eta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), t);
        update_de<MOD>(d, e, t);
        update_fg<MOD>(f, g, t);
    }
    // f = +-1: the inverse is f d, brought into [0, m)
    const int32_t neg = f[NL - 1] >> 31;
    int32_t s = d[NL - 1] >> 31;
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
        int32_t x = d[i] + (MOD::limb(i) & s);
        x = (x ^ neg) - neg + carry;
        carry = x >> 30;
        d[i] = i < NL - 1 ? (x & M30) : x;
    }
    s = d[NL - 1] >> 31;
    carry = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
        int32_t x = d[i] + (MOD::limb(i) & s) + carry;
        carry = x >> 30;
        d[i] = i < NL - 1 ? (x & M30) : x;
    }
#pragma unroll
    for (int w = 0; w < NW; w++) {
        // bits [32 w, 32 w + 32) of d
        const int bit = 32 * w, i = bit / 30, sh = bit % 30;
        uint32_t x = (uint32_t)d[i] >> sh;
        if (i + 1 < NL) x |= (uint32_t)d[i + 1] << (30 - sh);
        if (sh > 28 && i + 2 < NL) x |= (uint32_t)d[i + 2] << (60 - sh);
        out[w] = x;
    }
}
template <class AllDone>
CDP_SAFEGCD_FN void inverse_int(uint32_t *out, const uint32_t *a, AllDone all_done) {  // mod p
    inverse_mod<mod_fp>(out, a, all_done);
}
template <class AllDone>
CDP_SAFEGCD_FN void inverse_int_fr(uint32_t *out, const uint32_t *a, AllDone all_done) {  // mod r
    inverse_mod<mod_fr>(out, a, all_done);
}

}  // namespace safegcd
}  // namespace cdp
