// BLS12-381 scalar field Fr on the device for the verifier's scalar algebra (k_vcoeffs.cu, the verifier transcript kernels in
// k_transcript.cu): Montgomery form, R = 2^256, 8 x 32-bit limbs with 64-bit products -- the same representation as the host's Fr
// (host/fr.hpp, 4 x u64) and ark-ff's `Fr`.  A few thousand products per proof: nothing here is throughput-critical, so this is plain
// C++ (no PTX) and also compiles with g++ for the CPU harnesses under tests/host/.
#pragma once
#include <stdint.h>
#include "fp_inv_safegcd.cuh"

namespace cdp {

namespace vcoef {

struct fr_t {
    uint32_t v[8];
};
// r, R mod r, -r^-1 mod 2^32 (the same constants as host/fr.hpp and constants.cuh, as literals so that the CPU harness sees them too)
__device__ __forceinline__ uint32_t fr_mod(int i) {
    const uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return t[i];
}
__device__ __forceinline__ uint32_t fr_r2(int i) {
    const uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    return t[i];
}
__device__ __forceinline__ uint32_t fr_r1(int i) {
    const uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
    return t[i];
}
constexpr uint32_t FR_NINV = 0xffffffffu;

__device__ __forceinline__ fr_t fr_zero() {
    fr_t r;
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
__device__ __forceinline__ fr_t fr_one() {
    fr_t r;
    for (int i = 0; i < 8; i++) r.v[i] = fr_r1(i);
    return r;
}
__device__ __forceinline__ bool fr_geq_mod(const uint32_t *a) {
    for (int i = 7; i >= 0; i--) {
        if (a[i] > fr_mod(i)) return true;
        if (a[i] < fr_mod(i)) return false;
    }
    return true;
}
__device__ __forceinline__ void fr_sub_mod(uint32_t *a) {
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - fr_mod(i) - borrow;
        a[i] = (uint32_t)d;
        borrow = (d >> 32) & 1;
    }
}
__device__ __forceinline__ fr_t fr_add(const fr_t &a, const fr_t &b) {
    fr_t r;
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    if (c || fr_geq_mod(r.v)) fr_sub_mod(r.v);
    return r;
}
__device__ __forceinline__ fr_t fr_sub(const fr_t &a, const fr_t &b) {
    fr_t r;
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - borrow;
        r.v[i] = (uint32_t)d;
        borrow = (d >> 32) & 1;
    }
    if (borrow) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)r.v[i] + fr_mod(i);
            r.v[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    return r;
}
__device__ __forceinline__ fr_t fr_neg(const fr_t &a) { return fr_sub(fr_zero(), a); }
// Montgomery product (R = 2^256), coarsely integrated operand scanning
static __device__ __noinline__ fr_t fr_mul(const fr_t &a, const fr_t &b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[j] * b.v[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        const uint32_t m = t[0] * FR_NINV;
        c = (uint64_t)m * fr_mod(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * fr_mod(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    fr_t r;
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    if (t[8] || fr_geq_mod(r.v)) fr_sub_mod(r.v);
    return r;
}
__device__ __forceinline__ fr_t fr_load(const uint32_t *p) {
    fr_t r;
    for (int i = 0; i < 8; i++) r.v[i] = p[i];
    return r;
}
// canonical little-endian value -> Montgomery form
__device__ __forceinline__ fr_t fr_to_mont(const fr_t &a) {
    fr_t r2;
    for (int i = 0; i < 8; i++) r2.v[i] = fr_r2(i);
    return fr_mul(a, r2);
}
// Montgomery form -> canonical value, stored as 8 words
__device__ __forceinline__ void fr_store_canonical(uint32_t *dst, const fr_t &a) {
    fr_t one = fr_zero();
    one.v[0] = 1;
    const fr_t c = fr_mul(a, one);
    for (int i = 0; i < 8; i++) dst[i] = c.v[i];
}
// x^e for a small exponent
__device__ __forceinline__ fr_t fr_pow_u32(const fr_t &x, uint32_t e) {
    fr_t acc = fr_one();
    bool started = false;
    for (int i = 31; i >= 0; i--) {
        if (started) acc = fr_mul(acc, acc);
        if ((e >> i) & 1) {
            acc = started ? fr_mul(acc, x) : x;
            started = true;
        }
    }
    return acc;
}
// s_i = prod_{j : bit (m-1-j) of i set} g_j
__device__ __forceinline__ fr_t s_value(const uint32_t *g, uint32_t m, uint32_t i) {
    fr_t acc = fr_one();
    bool started = false;
    for (uint32_t j = 0; j < m; j++)
        if ((i >> (m - 1 - j)) & 1) {
            const fr_t gj = fr_load(g + 8 * j);
            acc = started ? fr_mul(acc, gj) : gj;
            started = true;
        }
    return acc;
}


// x^-1 = x^(r-2) (Fermat; 0 -> 0)
static __device__ __noinline__ fr_t fr_inverse(const fr_t &x) {
    fr_t acc = fr_one();
    bool started = false;
    for (int i = 254; i >= 0; i--) {
        // r - 2: r ends in ...ffffffff 00000001, so r - 2 ends in ...fffffffe ffffffff
        const uint32_t w = (i >> 5) == 0 ? 0xffffffffu : (i >> 5) == 1 ? 0xfffffffeu : fr_mod(i >> 5);
        if (started) acc = fr_mul(acc, acc);
        if ((w >> (i & 31)) & 1) {
            acc = started ? fr_mul(acc, x) : x;
            started = true;
        }
    }
    return acc;
}
// x^-1 by the binary extended Euclidean algorithm on plain integers (0 -> 0).  ~765 shift / subtract steps of 8-limb words instead of
// the ~380 dependent Montgomery products of the Fermat ladder: about 4x less latency for a single thread, which is what matters where one
// thread inverts a fresh Fiat-Shamir challenge on the critical path of every folding round (k_prove.cu).  Input and output in Montgomery form:
// the integer inverse of xR is x^-1 R^-1, one product with R^3 gives x^-1 R.
namespace detail {
__device__ __forceinline__ bool u256_is_one(const uint32_t *a) { return a[0] == 1 && (a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0; }
__device__ __forceinline__ void u256_shr1(uint32_t *a, uint32_t top) {
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] = (a[7] >> 1) | (top << 31);
}
// a <- a / 2 mod r
__device__ __forceinline__ void fr_halve(uint32_t *a) {
    uint32_t top = 0;
    if (a[0] & 1) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)a[i] + fr_mod(i);
            a[i] = (uint32_t)c;
            c >>= 32;
        }
        top = (uint32_t)c;
    }
    u256_shr1(a, top);
}
// a >= b
__device__ __forceinline__ bool u256_geq(const uint32_t *a, const uint32_t *b) {
    for (int i = 7; i >= 0; i--) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
__device__ __forceinline__ void u256_sub(uint32_t *a, const uint32_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) {
        const uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        a[i] = (uint32_t)d;
        borrow = (d >> 32) & 1;
    }
}
}  // namespace detail
static __device__ __noinline__ fr_t fr_inverse_euclid(const fr_t &x) {
    using namespace detail;
    fr_t u = x, v, x1 = fr_zero(), x2 = fr_zero();
    bool zero = true;
    for (int i = 0; i < 8; i++) {
        v.v[i] = fr_mod(i);
        zero = zero && x.v[i] == 0;
    }
    if (zero) return x;
    x1.v[0] = 1;
    // invariants: x1 * x = u, x2 * x = v (mod r); u, v odd after the halving loops, gcd(u, v) = 1
    while (!u256_is_one(u.v) && !u256_is_one(v.v)) {
        while (!(u.v[0] & 1)) {
            u256_shr1(u.v, 0);
            fr_halve(x1.v);
        }
        while (!(v.v[0] & 1)) {
            u256_shr1(v.v, 0);
            fr_halve(x2.v);
        }
        if (u256_geq(u.v, v.v)) {
            u256_sub(u.v, v.v);
            x1 = fr_sub(x1, x2);
        } else {
            u256_sub(v.v, u.v);
            x2 = fr_sub(x2, x1);
        }
    }
    fr_t r3;
    const uint32_t R3[8] = {0x439b73afu, 0xc62c1807u, 0x8cf06990u, 0x1b3e0d18u, 0xc7b5f418u, 0x73d13c71u, 0xc8db33e9u, 0x6e2a5bb9u};  // 2^768 mod r
    for (int i = 0; i < 8; i++) r3.v[i] = R3[i];
    return fr_mul(u256_is_one(u.v) ? x1 : x2, r3);
}
// x^-1 by division steps (fp_inv_safegcd.cuh, the scalar field's modulus: ~17 batches of ~350 instructions; 0 -> 0): what the device code uses.
// The Fermat ladder and the binary Euclid above stay as cross-checks (tests/host/cpu_engine_mock.cpp compares all three).
static __device__ __noinline__ fr_t fr_inverse_safegcd(const fr_t &x) {
    fr_t t, r3;
    safegcd::inverse_int_fr(t.v, x.v, [](bool done) { return done; });
    const uint32_t R3[8] = {0x439b73afu, 0xc62c1807u, 0x8cf06990u, 0x1b3e0d18u, 0xc7b5f418u, 0x73d13c71u, 0xc8db33e9u, 0x6e2a5bb9u};  // 2^768 mod r
    for (int i = 0; i < 8; i++) r3.v[i] = R3[i];
    return fr_mul(t, r3);
}
// xs[k] <- xs[k]^-1 for k < n <= 16 with one inversion (Montgomery's trick; no element may be zero: challenges never are)
__device__ __forceinline__ void fr_batch_inverse(fr_t *xs, uint32_t n) {
    fr_t pre[16];
    fr_t acc = fr_one();
    for (uint32_t k = 0; k < n; k++) {
        pre[k] = acc;
        acc = fr_mul(acc, xs[k]);
    }
    fr_t inv = fr_inverse_safegcd(acc);
    for (uint32_t k = n; k-- > 0;) {
        const fr_t t = fr_mul(inv, pre[k]);
        inv = fr_mul(inv, xs[k]);
        xs[k] = t;
    }
}

}  // namespace vcoef

// positions inside a proof's challenge block (Montgomery scalars, 8 words each; cdp_verify_coeffs_dev in include/cdp_msm.h); the four
// challenge vectors follow at CH_VEC
enum {
    CH_RHO = 0, CH_ALPHA_SP = 12, CH_BETA_SP, CH_ALPHA_G, CH_BETA_INV, CH_ALPHA_I, CH_BETA_I, CH_Z, CH_C, CH_D, CH_X, CH_ALPHA_SM,
    CH_ALPHA_SS, CH_ZK, CH_ZT, CH_ZU, CH_VEC = 27
};

}  // namespace cdp
