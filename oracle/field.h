/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Fp (381-bit base field) and Fr (255-bit scalar field) of BLS12-381 in Montgomery form on
 * 64-bit limbs.  Restates the arithmetic the reference gets from ark-ff ^0.4.0
 * (`/root/reference/Cargo.toml:24`; the crate is NOT vendored under /root/reference):
 *   Fp384 = [u64;6] little-endian limbs, Montgomery R = 2^384,
 *   Fr    = [u64;4] little-endian limbs, Montgomery R = 2^256.
 * Used at every reference call site listed in SURVEY.md section 8c (e.g. `inverse`
 * src/inner_product_argument.rs:69-71,171; `pow` src/grand_product_argument.rs:101,125;
 * `batch_inversion` src/inner_product_argument.rs:234,247).
 * The algorithm is textbook CIOS Montgomery multiplication; nothing here is copied.
 */
#ifndef CDP_ORACLE_FIELD_H
#define CDP_ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>
#include "constants.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp_t;
typedef struct { uint64_t l[4]; } fr_t;

#include <immintrin.h>
#define FORCE_INLINE static inline __attribute__((always_inline))

/* ---- generic N-limb helpers (N is a compile-time constant at every call site) ---- */
FORCE_INLINE int limbs_geq(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
FORCE_INLINE int limbs_is_zero(const uint64_t *a, int n) {
    uint64_t acc = 0;
    for (int i = 0; i < n; i++) acc |= a[i];
    return acc == 0;
}
FORCE_INLINE uint64_t limbs_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    unsigned char c = 0;
    for (int i = 0; i < n; i++) c = _addcarry_u64(c, a[i], b[i], (unsigned long long *)&r[i]);
    return c;
}
FORCE_INLINE uint64_t limbs_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    unsigned char c = 0;
    for (int i = 0; i < n; i++) c = _subborrow_u64(c, a[i], b[i], (unsigned long long *)&r[i]);
    return c;
}
/* (a + b) mod m and (a - b) mod m for a, b < m: the correction is selected with a mask instead of a branch on unpredictable data */
FORCE_INLINE void mod_add(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int n) {
    uint64_t t[6], u[6];
    const uint64_t carry = limbs_add(t, a, b, n);
    const uint64_t borrow = limbs_sub(u, t, m, n);
    const uint64_t keep = (uint64_t)0 - (uint64_t)(borrow & (carry ^ 1));  /* all ones: t < m, keep t */
    for (int i = 0; i < n; i++) r[i] = (t[i] & keep) | (u[i] & ~keep);
}
FORCE_INLINE void mod_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int n) {
    uint64_t t[6];
    const uint64_t mask = (uint64_t)0 - limbs_sub(t, a, b, n);
    unsigned char c = 0;
    for (int i = 0; i < n; i++) c = _addcarry_u64(c, t[i], m[i] & mask, (unsigned long long *)&r[i]);
}
/* CIOS Montgomery product: r = a*b*R^-1 mod m */
FORCE_INLINE void mont_mul(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m,
                           uint64_t inv, int n) {
    uint64_t t[8] = {0};
    for (int i = 0; i < n; i++) {
        u128 c = 0;
        for (int j = 0; j < n; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[n]; t[n] = (uint64_t)c; t[n + 1] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * inv;
        c = (u128)q * m[0] + t[0]; c >>= 64;
        for (int j = 1; j < n; j++) {
            c += (u128)q * m[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[n]; t[n - 1] = (uint64_t)c;
        t[n] = t[n + 1] + (uint64_t)(c >> 64);
    }
    if (t[n] || limbs_geq(t, m, n)) limbs_sub(t, t, m, n);
    memcpy(r, t, 8 * n);
}

/* ------------------------------------------------------------------ Fp */
/* Fp product: CIOS specialised to six limbs, without the extra carry word -- the top limb of p has its three high bits clear, so the
 * running value stays below 2p between sweeps.  (The generic mont_mul above serves Fr.)  This and the dedicated squaring / Euclidean
 * inversion below exist so that the CPU baseline timed from this port is not flattered by a slow field (bench.py cpu_baseline). */
#define CDP_FP_SWEEP(bi)                                                                                    \
    {                                                                                                       \
        u128 A = (u128)x[0] * (bi) + t0;                                                                    \
        const uint64_t q = (uint64_t)A * FP_INV64;                                                          \
        u128 C = (u128)q * FP_P[0] + (uint64_t)A;                                                           \
        A = (u128)x[1] * (bi) + t1 + (uint64_t)(A >> 64); C = (u128)q * FP_P[1] + (uint64_t)A + (uint64_t)(C >> 64); t0 = (uint64_t)C; \
        A = (u128)x[2] * (bi) + t2 + (uint64_t)(A >> 64); C = (u128)q * FP_P[2] + (uint64_t)A + (uint64_t)(C >> 64); t1 = (uint64_t)C; \
        A = (u128)x[3] * (bi) + t3 + (uint64_t)(A >> 64); C = (u128)q * FP_P[3] + (uint64_t)A + (uint64_t)(C >> 64); t2 = (uint64_t)C; \
        A = (u128)x[4] * (bi) + t4 + (uint64_t)(A >> 64); C = (u128)q * FP_P[4] + (uint64_t)A + (uint64_t)(C >> 64); t3 = (uint64_t)C; \
        A = (u128)x[5] * (bi) + t5 + (uint64_t)(A >> 64); C = (u128)q * FP_P[5] + (uint64_t)A + (uint64_t)(C >> 64); t4 = (uint64_t)C; \
        t5 = (uint64_t)(C >> 64) + (uint64_t)(A >> 64);                                                     \
    }
FORCE_INLINE void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) {
    const uint64_t *x = a->l, *y = b->l;
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    CDP_FP_SWEEP(y[0]) CDP_FP_SWEEP(y[1]) CDP_FP_SWEEP(y[2]) CDP_FP_SWEEP(y[3]) CDP_FP_SWEEP(y[4]) CDP_FP_SWEEP(y[5])
    uint64_t t[6] = {t0, t1, t2, t3, t4, t5};
    if (limbs_geq(t, FP_P, 6)) limbs_sub(t, t, FP_P, 6);
    memcpy(r->l, t, 48);
}
#undef CDP_FP_SWEEP
/* Fp square: the 15 cross products once, doubled, plus the 6 diagonal squares (21 limb products instead of 36), then six reduction sweeps */
FORCE_INLINE void fp_sqr(fp_t *r, const fp_t *a) {
    const uint64_t *x = a->l;
    uint64_t o[12] = {0}, t[13];
    for (int i = 0; i < 5; i++) {
        uint64_t carry = 0;
        for (int j = i + 1; j < 6; j++) {
            u128 c = (u128)x[i] * x[j] + o[i + j] + carry;
            o[i + j] = (uint64_t)c; carry = (uint64_t)(c >> 64);
        }
        o[i + 6] = carry;
    }
    uint64_t top = 0, carry = 0;
    for (int k = 0; k < 12; k++) { const uint64_t v = o[k]; o[k] = (v << 1) | top; top = v >> 63; }
    for (int i = 0; i < 6; i++) {
        u128 c = (u128)x[i] * x[i] + o[2 * i] + carry;
        t[2 * i] = (uint64_t)c;
        c = (u128)o[2 * i + 1] + (uint64_t)(c >> 64);
        t[2 * i + 1] = (uint64_t)c; carry = (uint64_t)(c >> 64);
    }
    t[12] = 0;
    for (int i = 0; i < 6; i++) {
        const uint64_t q = t[i] * FP_INV64;
        uint64_t cy = 0;
        for (int j = 0; j < 6; j++) {
            u128 c = (u128)q * FP_P[j] + t[i + j] + cy;
            t[i + j] = (uint64_t)c; cy = (uint64_t)(c >> 64);
        }
        for (int k = i + 6; cy && k < 13; k++) { u128 c = (u128)t[k] + cy; t[k] = (uint64_t)c; cy = (uint64_t)(c >> 64); }
    }
    if (t[12] || limbs_geq(t + 6, FP_P, 6)) limbs_sub(t + 6, t + 6, FP_P, 6);
    memcpy(r->l, t + 6, 48);
}
FORCE_INLINE void fp_add(fp_t *r, const fp_t *a, const fp_t *b) { mod_add(r->l, a->l, b->l, FP_P, 6); }
FORCE_INLINE void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) { mod_sub(r->l, a->l, b->l, FP_P, 6); }
FORCE_INLINE void fp_dbl(fp_t *r, const fp_t *a) { mod_add(r->l, a->l, a->l, FP_P, 6); }
FORCE_INLINE int fp_is_zero(const fp_t *a) { return limbs_is_zero(a->l, 6); }
FORCE_INLINE int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a, b, sizeof(fp_t)) == 0; }
FORCE_INLINE void fp_neg(fp_t *r, const fp_t *a) {
    if (fp_is_zero(a)) { *r = *a; return; }
    limbs_sub(r->l, FP_P, a->l, 6);
}
FORCE_INLINE void fp_one(fp_t *r) { memcpy(r->l, FP_R_MOD_P, 48); }
FORCE_INLINE void fp_zero(fp_t *r) { memset(r, 0, sizeof *r); }
/* canonical integer (6 limbs) <-> Montgomery */
FORCE_INLINE void fp_from_canon(fp_t *r, const uint64_t c[6]) {
    fp_t t, r2; memcpy(t.l, c, 48); memcpy(r2.l, FP_R2_MOD_P, 48); fp_mul(r, &t, &r2);
}
FORCE_INLINE void fp_to_canon(uint64_t c[6], const fp_t *a) {
    fp_t one = {{1, 0, 0, 0, 0, 0}}, t; fp_mul(&t, a, &one); memcpy(c, t.l, 48);
}
static void fp_pow(fp_t *r, const fp_t *a, const uint64_t *e, int nlimbs) {
    fp_t acc; fp_one(&acc);
    int started = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) fp_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) { fp_mul(&acc, &acc, a); started = 1; }
    }
    *r = acc;
}
/* a^-1 (0 -> 0) by the binary extended Euclidean algorithm on the integer representative: for the Montgomery form aR the integer inverse
 * is a^-1 R^-1, and one product with R^3 mod p (= 2^1152 mod p) gives a^-1 R.  ~5x fewer limb operations than the Fermat ladder
 * fp_pow(a, p - 2); the same element either way (tests/test_oracle_golden.py checks both). */
static const uint64_t FP_R3_MOD_P[6] = {0xed48ac6bd94ca1e0ULL, 0x315f831e03a7adf8ULL, 0x9a53352a615e29ddULL, 0x34c04e5e921e1761ULL, 0x2512d43565724728ULL, 0x0aa6346091755d4dULL};
static void fp_inv_fermat(fp_t *r, const fp_t *a) { fp_pow(r, a, FP_P_MINUS_2, 6); }
FORCE_INLINE void limbs_shr1(uint64_t *a, uint64_t top, int n) {
    for (int i = 0; i + 1 < n; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[n - 1] = (a[n - 1] >> 1) | (top << 63);
}
FORCE_INLINE void fp_halve_int(uint64_t *a) { /* a <- a / 2 mod p on integers < p */
    uint64_t top = 0;
    if (a[0] & 1) top = limbs_add(a, a, FP_P, 6);
    limbs_shr1(a, top, 6);
}
FORCE_INLINE int limbs_is_one(const uint64_t *a, int n) {
    if (a[0] != 1) return 0;
    for (int i = 1; i < n; i++) if (a[i]) return 0;
    return 1;
}
static void fp_inv(fp_t *r, const fp_t *a) {
    if (fp_is_zero(a)) { *r = *a; return; }
    uint64_t u[6], v[6], x1[6] = {1, 0, 0, 0, 0, 0}, x2[6] = {0};
    memcpy(u, a->l, 48); memcpy(v, FP_P, 48);
    while (!limbs_is_one(u, 6) && !limbs_is_one(v, 6)) {
        while (!(u[0] & 1)) { limbs_shr1(u, 0, 6); fp_halve_int(x1); }
        while (!(v[0] & 1)) { limbs_shr1(v, 0, 6); fp_halve_int(x2); }
        if (limbs_geq(u, v, 6)) { limbs_sub(u, u, v, 6); mod_sub(x1, x1, x2, FP_P, 6); }
        else { limbs_sub(v, v, u, 6); mod_sub(x2, x2, x1, FP_P, 6); }
    }
    fp_t t, r3;
    memcpy(t.l, limbs_is_one(u, 6) ? x1 : x2, 48); memcpy(r3.l, FP_R3_MOD_P, 48);
    fp_mul(r, &t, &r3);
}
/* sqrt for p = 3 mod 4; returns 1 and writes r when a is a square */
static int fp_sqrt(fp_t *r, const fp_t *a) {
    fp_t s, s2; fp_pow(&s, a, FP_P_PLUS_1_DIV_4, 6); fp_sqr(&s2, &s);
    if (!fp_eq(&s2, a)) return 0;
    *r = s; return 1;
}
/* compare canonical integer values: returns 1 when a > b */
static int fp_canon_gt(const fp_t *a, const fp_t *b) {
    uint64_t ca[6], cb[6]; fp_to_canon(ca, a); fp_to_canon(cb, b);
    for (int i = 5; i >= 0; i--) { if (ca[i] > cb[i]) return 1; if (ca[i] < cb[i]) return 0; }
    return 0;
}

/* ------------------------------------------------------------------ Fr */
FORCE_INLINE void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { mont_mul(r->l, a->l, b->l, FR_R, FR_INV64, 4); }
FORCE_INLINE void fr_add(fr_t *r, const fr_t *a, const fr_t *b) { mod_add(r->l, a->l, b->l, FR_R, 4); }
FORCE_INLINE void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) { mod_sub(r->l, a->l, b->l, FR_R, 4); }
FORCE_INLINE int fr_is_zero(const fr_t *a) { return limbs_is_zero(a->l, 4); }
FORCE_INLINE int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, sizeof(fr_t)) == 0; }
FORCE_INLINE void fr_neg(fr_t *r, const fr_t *a) {
    if (fr_is_zero(a)) { *r = *a; return; }
    limbs_sub(r->l, FR_R, a->l, 4);
}
FORCE_INLINE void fr_one(fr_t *r) { memcpy(r->l, FR_R_MOD_R, 32); }
FORCE_INLINE void fr_zero(fr_t *r) { memset(r, 0, sizeof *r); }
FORCE_INLINE void fr_from_canon(fr_t *r, const uint64_t c[4]) {
    fr_t t, r2; memcpy(t.l, c, 32); memcpy(r2.l, FR_R2_MOD_R, 32); fr_mul(r, &t, &r2);
}
FORCE_INLINE void fr_to_canon(uint64_t c[4], const fr_t *a) {
    fr_t one = {{1, 0, 0, 0}}, t; fr_mul(&t, a, &one); memcpy(c, t.l, 32);
}
FORCE_INLINE void fr_from_u64(fr_t *r, uint64_t v) { uint64_t c[4] = {v, 0, 0, 0}; fr_from_canon(r, c); }
static void fr_pow(fr_t *r, const fr_t *a, const uint64_t *e, int nlimbs) {
    fr_t acc; fr_one(&acc);
    int started = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) fr_mul(&acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) { fr_mul(&acc, &acc, a); started = 1; }
    }
    *r = acc;
}
static void fr_pow_u64(fr_t *r, const fr_t *a, uint64_t e) { fr_pow(r, a, &e, 1); }
static void fr_inv(fr_t *r, const fr_t *a) { fr_pow(r, a, FR_R_MINUS_2, 4); }
/* Montgomery-trick inversion of every non-zero entry (ark-ff `batch_inversion` semantics:
 * zero entries are left untouched). */
static void fr_batch_inv(fr_t *v, size_t n, fr_t *scratch /* n entries */) {
    fr_t acc; fr_one(&acc);
    for (size_t i = 0; i < n; i++) {
        scratch[i] = acc;
        if (!fr_is_zero(&v[i])) fr_mul(&acc, &acc, &v[i]);
    }
    fr_inv(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        if (fr_is_zero(&v[i])) continue;
        fr_t t; fr_mul(&t, &acc, &scratch[i]);
        fr_mul(&acc, &acc, &v[i]);
        v[i] = t;
    }
}
static void fr_inner_product(fr_t *r, const fr_t *a, const fr_t *b, size_t n) {
    fr_t acc; fr_zero(&acc);
    for (size_t i = 0; i < n; i++) { fr_t t; fr_mul(&t, &a[i], &b[i]); fr_add(&acc, &acc, &t); }
    *r = acc;
}
#endif
