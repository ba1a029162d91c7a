/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Deterministic RNG chain the reference's tests use: `StdRng::seed_from_u64(0)`
 * (e.g. `/root/reference/src/whisk.rs:383,418`, `src/crs.rs:63`) followed by arkworks sampling
 * (`Fr::rand`, `G1Projective::rand`) and `SliceRandom::shuffle` (`src/whisk.rs:153`).
 * None of rand 0.8 / rand_chacha 0.3 / rand_core 0.6 / ark-ff / ark-ec is vendored under
 * /root/reference; this restates their published behaviour:
 *   seed_from_u64 : PCG32 (mul 6364136223846793005, inc 11634580027462260723) expands the u64
 *                   into the 32-byte seed, 4 bytes at a time
 *   StdRng        : ChaCha12, 64-bit block counter in words 12-13, stream 0, 4 blocks (64 u32)
 *                   buffered per refill; next_u64 = two consecutive u32 (lo, hi), straddling
 *                   a refill when one word is left
 *   Fp::rand      : N x next_u64 little-endian limbs, top bits masked to the modulus width,
 *                   rejected when >= modulus, and the limbs are taken *as the Montgomery form*
 *   bool          : top bit of next_u32
 *   G1::rand      : loop { x <- Fq::rand; greatest <- bool; point from x (larger y iff greatest) }
 *                   then multiplication by the cofactor h
 *   shuffle       : Fisher-Yates from the top, index <- gen_range(0..i+1) on u32 using the
 *                   widening-multiply rejection zone
 * The only judge of these statements is the pair of seed-0 golden proofs
 * (src/whisk.rs:401 and :455), checked in tests/test_oracle_golden.py.
 */
#ifndef CDP_ORACLE_RNG_H
#define CDP_ORACLE_RNG_H
#include "g1.h"

typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[64];
    int index;
} stdrng_t;

#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define CHACHA_QR(a, b, c, d) \
    a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); \
    a += b; d ^= a; d = ROTL32(d, 8);  c += d; b ^= c; b = ROTL32(b, 7);

static void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
    uint32_t x[16]; memcpy(x, s, sizeof x);
    for (int i = 0; i < 6; i++) {
        CHACHA_QR(x[0], x[4], x[8], x[12]) CHACHA_QR(x[1], x[5], x[9], x[13])
        CHACHA_QR(x[2], x[6], x[10], x[14]) CHACHA_QR(x[3], x[7], x[11], x[15])
        CHACHA_QR(x[0], x[5], x[10], x[15]) CHACHA_QR(x[1], x[6], x[11], x[12])
        CHACHA_QR(x[2], x[7], x[8], x[13]) CHACHA_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}
static void stdrng_refill(stdrng_t *r) {
    for (int b = 0; b < 4; b++) chacha12_block(r->key, r->counter + b, r->buf + 16 * b);
    r->counter += 4;
}
static void stdrng_seed_from_u64(stdrng_t *r, uint64_t state) {
    for (int i = 0; i < 8; i++) {
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        r->key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
    r->counter = 0; r->index = 64;
}
static uint32_t stdrng_next_u32(stdrng_t *r) {
    if (r->index >= 64) { stdrng_refill(r); r->index = 0; }
    return r->buf[r->index++];
}
static uint64_t stdrng_next_u64(stdrng_t *r) {
    if (r->index < 63) {
        uint64_t v = ((uint64_t)r->buf[r->index + 1] << 32) | r->buf[r->index];
        r->index += 2; return v;
    } else if (r->index >= 64) {
        stdrng_refill(r); r->index = 2;
        return ((uint64_t)r->buf[1] << 32) | r->buf[0];
    } else {
        uint64_t lo = r->buf[63];
        stdrng_refill(r); r->index = 1;
        return ((uint64_t)r->buf[0] << 32) | lo;
    }
}
static void fr_rand(fr_t *out, stdrng_t *r) {
    for (;;) {
        for (int i = 0; i < 4; i++) out->l[i] = stdrng_next_u64(r);
        out->l[3] &= 0xFFFFFFFFFFFFFFFFULL >> 1;
        if (!limbs_geq(out->l, FR_R, 4)) return;
    }
}
static void fp_rand(fp_t *out, stdrng_t *r) {
    for (;;) {
        for (int i = 0; i < 6; i++) out->l[i] = stdrng_next_u64(r);
        out->l[5] &= 0xFFFFFFFFFFFFFFFFULL >> 3;
        if (!limbs_geq(out->l, FP_P, 6)) return;
    }
}
static int stdrng_bool(stdrng_t *r) { return (int)(stdrng_next_u32(r) >> 31); }
static void g1j_rand(g1j_t *out, stdrng_t *r) {
    for (;;) {
        fp_t x, rhs, b, y; fp_rand(&x, r);
        int greatest = stdrng_bool(r);
        fp_sqr(&rhs, &x); fp_mul(&rhs, &rhs, &x); memcpy(b.l, FP_B_MONT, 48); fp_add(&rhs, &rhs, &b);
        if (!fp_sqrt(&y, &rhs)) continue;
        fp_t ny; fp_neg(&ny, &y);
        int y_is_larger = fp_canon_gt(&y, &ny);
        if (y_is_larger != greatest) y = ny;
        g1j_t p; p.X = x; p.Y = y; fp_one(&p.Z);
        g1j_mul_limbs(out, &p, G1_COFACTOR, 2);
        return;
    }
}
static uint32_t stdrng_gen_range_u32(stdrng_t *r, uint32_t range /* samples [0, range) */) {
    int lz = __builtin_clz(range);
    uint32_t zone = (range << lz) - 1;
    for (;;) {
        uint64_t m = (uint64_t)stdrng_next_u32(r) * range;
        if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
    }
}
static void stdrng_shuffle_u32(uint32_t *v, size_t n, stdrng_t *r) {
    for (size_t i = n; i-- > 1;) {
        uint32_t j = stdrng_gen_range_u32(r, (uint32_t)(i + 1));
        uint32_t t = v[i]; v[i] = v[j]; v[j] = t;
    }
}
#endif
