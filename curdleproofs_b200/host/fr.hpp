// Host-side scalar field Fr of BLS12-381 (r = 0x73eda753...00000001), Montgomery form on 4 x u64.
// The engine keeps ALL group arithmetic on the GPU; the host driver only needs the cheap O(n) Fr bookkeeping of the
// protocol (challenge algebra, vector folding, inner products) -- what the reference does with ark-ff's `Fr`
// (e.g. /root/reference/src/inner_product_argument.rs:136-139,175-176, src/grand_product_argument.rs:66-131).
#pragma once
#include <cstdint>
#include <cstring>

namespace cdp_host {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t v[4];

    static constexpr uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    static constexpr uint64_t R1[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};  // 2^256 mod r
    static constexpr uint64_t R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};  // 2^512 mod r
    static constexpr uint64_t NINV = 0xfffffffeffffffffULL;  // -r^-1 mod 2^64

    static Fr zero() { Fr r; memset(r.v, 0, 32); return r; }
    static Fr one() { Fr r; memcpy(r.v, R1, 32); return r; }
    static Fr raw(const uint64_t w[4]) { Fr r; memcpy(r.v, w, 32); return r; }
    static Fr from_u64(uint64_t x) { uint64_t c[4] = {x, 0, 0, 0}; return from_canonical_unchecked(c); }
    static Fr from_canonical_unchecked(const uint64_t c[4]) { return raw(c) * raw(R2); }
    static bool geq_mod(const uint64_t a[4]) {
        for (int i = 3; i >= 0; i--) {
            if (a[i] > MOD[i]) return true;
            if (a[i] < MOD[i]) return false;
        }
        return true;
    }
    // canonical little-endian bytes -> Fr; false when the value is >= r
    static bool from_bytes(const uint8_t b[32], Fr &out) {
        uint64_t c[4];
        memcpy(c, b, 32);
        if (geq_mod(c)) return false;
        out = from_canonical_unchecked(c);
        return true;
    }
    void to_canonical(uint64_t c[4]) const {
        uint64_t one_[4] = {1, 0, 0, 0};
        Fr t = *this * raw(one_);
        memcpy(c, t.v, 32);
    }
    void to_bytes(uint8_t out[32]) const {
        uint64_t c[4];
        to_canonical(c);
        memcpy(out, c, 32);
    }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const Fr &o) const { return memcmp(v, o.v, 32) == 0; }

    Fr operator+(const Fr &o) const {
        Fr r;
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)v[i] + o.v[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
        if (c || geq_mod(r.v)) r.sub_mod();
        return r;
    }
    Fr operator-(const Fr &o) const {
        Fr r;
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)v[i] - o.v[i] - borrow;
            r.v[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            u128 c = 0;
            for (int i = 0; i < 4; i++) {
                c += (u128)r.v[i] + MOD[i];
                r.v[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    Fr neg() const { return zero() - *this; }
    // Montgomery product, coarsely integrated operand scanning without the extra carry word (r < 2^255: the top limb of the
    // modulus has its high bit clear, so t never exceeds 2r - 1 between sweeps)
    Fr operator*(const Fr &o) const {
        uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define CDP_FR_SWEEP(bi)                                                                 \
    {                                                                                    \
        u128 A = (u128)v[0] * (bi) + t0;                                                 \
        uint64_t m = (uint64_t)A * NINV;                                                 \
        u128 C = (u128)m * MOD[0] + (uint64_t)A;                                         \
        A = (u128)v[1] * (bi) + t1 + (uint64_t)(A >> 64);                                \
        C = (u128)m * MOD[1] + (uint64_t)A + (uint64_t)(C >> 64);                        \
        t0 = (uint64_t)C;                                                                \
        A = (u128)v[2] * (bi) + t2 + (uint64_t)(A >> 64);                                \
        C = (u128)m * MOD[2] + (uint64_t)A + (uint64_t)(C >> 64);                        \
        t1 = (uint64_t)C;                                                                \
        A = (u128)v[3] * (bi) + t3 + (uint64_t)(A >> 64);                                \
        C = (u128)m * MOD[3] + (uint64_t)A + (uint64_t)(C >> 64);                        \
        t2 = (uint64_t)C;                                                                \
        t3 = (uint64_t)(C >> 64) + (uint64_t)(A >> 64);                                  \
    }
        CDP_FR_SWEEP(o.v[0]) CDP_FR_SWEEP(o.v[1]) CDP_FR_SWEEP(o.v[2]) CDP_FR_SWEEP(o.v[3])
#undef CDP_FR_SWEEP
        Fr r;
        r.v[0] = t0; r.v[1] = t1; r.v[2] = t2; r.v[3] = t3;
        if (geq_mod(r.v)) r.sub_mod();
        return r;
    }
    Fr &operator+=(const Fr &o) { return *this = *this + o; }
    Fr &operator-=(const Fr &o) { return *this = *this - o; }
    Fr &operator*=(const Fr &o) { return *this = *this * o; }
    Fr pow(const uint64_t *e, int limbs) const {
        Fr acc = one();
        bool started = false;
        for (int i = limbs * 64 - 1; i >= 0; i--) {
            if (started) acc = acc * acc;
            if ((e[i / 64] >> (i % 64)) & 1) {
                acc = acc * *this;
                started = true;
            }
        }
        return acc;
    }
    Fr pow_u64(uint64_t e) const { return pow(&e, 1); }
    Fr inverse() const {  // Fermat; inverse of zero is zero (callers never invert zero: challenges are non-zero)
        uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
        return pow(e, 4);
    }

   private:
    void sub_mod() {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)v[i] - MOD[i] - borrow;
            v[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
};

inline Fr inner_product(const Fr *a, const Fr *b, size_t n) {
    Fr acc = Fr::zero();
    for (size_t i = 0; i < n; i++) acc += a[i] * b[i];
    return acc;
}

}  // namespace cdp_host
