/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * BLS12-381 G1 group law, scalar multiplication, batch normalisation, the ZCash-format
 * 48-byte encoding, and a Pippenger MSM.  Restates what the reference obtains from
 * ark-ec ^0.4.0 / ark-bls12-381 ^0.4.0 (`/root/reference/Cargo.toml:23,25`; not vendored):
 *   - `G1Projective::msm`            <- /root/reference/src/util.rs:19-22
 *   - `CurveGroup::normalize_batch`  <- /root/reference/src/util.rs:25-29
 *   - `Mul<Fr>` + `into_affine()`    <- src/inner_product_argument.rs:174-179,
 *                                       src/same_multiscalar_argument.rs:126-131,
 *                                       src/grand_product_argument.rs:92-102
 *   - `serialize_compressed`         <- src/transcript.rs:29-33, src/whisk.rs:308-316
 * Every observable output of those calls is a canonical group element, so the
 * formulas below (EFD dbl-2009-l / add-2007-bl / madd-2007-bl for a = 0) need only be
 * mathematically right; they are pinned by tests/test_oracle_*.py against oracle/py_ref.py,
 * the generator KAT (src/whisk.rs:365) and the two seed-0 golden proofs (src/whisk.rs:401,455).
 */
#ifndef CDP_ORACLE_G1_H
#define CDP_ORACLE_G1_H
#include <stdlib.h>
#include "field.h"

typedef struct { fp_t x, y; } g1a_t;          /* affine; infinity <=> x == y == 0 */
typedef struct { fp_t X, Y, Z; } g1j_t;       /* Jacobian; infinity <=> Z == 0     */

FORCE_INLINE int g1a_is_inf(const g1a_t *p) { return fp_is_zero(&p->x) && fp_is_zero(&p->y); }
FORCE_INLINE int g1j_is_inf(const g1j_t *p) { return fp_is_zero(&p->Z); }
FORCE_INLINE void g1j_set_inf(g1j_t *p) { fp_one(&p->X); fp_one(&p->Y); fp_zero(&p->Z); }
FORCE_INLINE void g1a_set_inf(g1a_t *p) { memset(p, 0, sizeof *p); }
FORCE_INLINE void g1j_from_affine(g1j_t *r, const g1a_t *p) {
    if (g1a_is_inf(p)) { g1j_set_inf(r); return; }
    r->X = p->x; r->Y = p->y; fp_one(&r->Z);
}
FORCE_INLINE void g1a_generator(g1a_t *g) { memcpy(g->x.l, FP_GX_MONT, 48); memcpy(g->y.l, FP_GY_MONT, 48); }

static void g1j_dbl(g1j_t *r, const g1j_t *p) {
    if (g1j_is_inf(p)) { *r = *p; return; }
    fp_t A, B, C, D, E, F, t;
    fp_sqr(&A, &p->X); fp_sqr(&B, &p->Y); fp_sqr(&C, &B);
    fp_add(&t, &p->X, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_dbl(&D, &t);
    fp_dbl(&E, &A); fp_add(&E, &E, &A);
    fp_sqr(&F, &E);
    fp_t Z3; fp_mul(&Z3, &p->Y, &p->Z); fp_dbl(&Z3, &Z3);
    fp_t X3; fp_dbl(&t, &D); fp_sub(&X3, &F, &t);
    fp_t Y3; fp_sub(&t, &D, &X3); fp_mul(&Y3, &E, &t);
    fp_dbl(&C, &C); fp_dbl(&C, &C); fp_dbl(&C, &C); fp_sub(&Y3, &Y3, &C);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1j_add(g1j_t *r, const g1j_t *p, const g1j_t *q) {
    if (g1j_is_inf(p)) { *r = *q; return; }
    if (g1j_is_inf(q)) { *r = *p; return; }
    fp_t Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
    fp_sqr(&Z1Z1, &p->Z); fp_sqr(&Z2Z2, &q->Z);
    fp_mul(&U1, &p->X, &Z2Z2); fp_mul(&U2, &q->X, &Z1Z1);
    fp_mul(&S1, &p->Y, &q->Z); fp_mul(&S1, &S1, &Z2Z2);
    fp_mul(&S2, &q->Y, &p->Z); fp_mul(&S2, &S2, &Z1Z1);
    fp_sub(&H, &U2, &U1); fp_sub(&rr, &S2, &S1);
    if (fp_is_zero(&H)) {
        if (fp_is_zero(&rr)) { g1j_dbl(r, p); return; }
        g1j_set_inf(r); return;
    }
    fp_dbl(&rr, &rr);
    fp_dbl(&I, &H); fp_sqr(&I, &I); fp_mul(&J, &H, &I); fp_mul(&V, &U1, &I);
    fp_t X3, Y3, Z3;
    fp_sqr(&X3, &rr); fp_sub(&X3, &X3, &J); fp_sub(&X3, &X3, &V); fp_sub(&X3, &X3, &V);
    fp_sub(&t, &V, &X3); fp_mul(&Y3, &rr, &t); fp_mul(&t, &S1, &J); fp_dbl(&t, &t); fp_sub(&Y3, &Y3, &t);
    fp_add(&Z3, &p->Z, &q->Z); fp_sqr(&Z3, &Z3); fp_sub(&Z3, &Z3, &Z1Z1); fp_sub(&Z3, &Z3, &Z2Z2); fp_mul(&Z3, &Z3, &H);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1j_add_affine(g1j_t *r, const g1j_t *p, const g1a_t *q) {
    if (g1a_is_inf(q)) { *r = *p; return; }
    if (g1j_is_inf(p)) { g1j_from_affine(r, q); return; }
    fp_t Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    fp_sqr(&Z1Z1, &p->Z); fp_mul(&U2, &q->x, &Z1Z1);
    fp_mul(&S2, &q->y, &p->Z); fp_mul(&S2, &S2, &Z1Z1);
    fp_sub(&H, &U2, &p->X); fp_sub(&rr, &S2, &p->Y);
    if (fp_is_zero(&H)) {
        if (fp_is_zero(&rr)) { g1j_dbl(r, p); return; }
        g1j_set_inf(r); return;
    }
    fp_dbl(&rr, &rr);
    fp_sqr(&HH, &H); fp_dbl(&I, &HH); fp_dbl(&I, &I); fp_mul(&J, &H, &I); fp_mul(&V, &p->X, &I);
    fp_t X3, Y3, Z3;
    fp_sqr(&X3, &rr); fp_sub(&X3, &X3, &J); fp_sub(&X3, &X3, &V); fp_sub(&X3, &X3, &V);
    fp_sub(&t, &V, &X3); fp_mul(&Y3, &rr, &t); fp_mul(&t, &p->Y, &J); fp_dbl(&t, &t); fp_sub(&Y3, &Y3, &t);
    fp_add(&Z3, &p->Z, &H); fp_sqr(&Z3, &Z3); fp_sub(&Z3, &Z3, &Z1Z1); fp_sub(&Z3, &Z3, &HH);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
FORCE_INLINE void g1j_neg(g1j_t *r, const g1j_t *p) { *r = *p; fp_neg(&r->Y, &p->Y); }
FORCE_INLINE void g1a_neg(g1a_t *r, const g1a_t *p) { *r = *p; fp_neg(&r->y, &p->y); }
static void g1j_sub(g1j_t *r, const g1j_t *p, const g1j_t *q) { g1j_t n; g1j_neg(&n, q); g1j_add(r, p, &n); }

/* `into_affine()` */
static void g1j_to_affine(g1a_t *r, const g1j_t *p) {
    if (g1j_is_inf(p)) { g1a_set_inf(r); return; }
    fp_t zi, zi2, zi3; fp_inv(&zi, &p->Z); fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
    fp_mul(&r->x, &p->X, &zi2); fp_mul(&r->y, &p->Y, &zi3);
}
/* `normalize_batch` (one inversion, Montgomery trick) -- src/util.rs:27 */
static void g1j_batch_to_affine(g1a_t *out, const g1j_t *in, size_t n) {
    if (n == 0) return;
    fp_t *pre = (fp_t *)malloc(n * sizeof(fp_t));
    fp_t acc; fp_one(&acc);
    for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!g1j_is_inf(&in[i])) fp_mul(&acc, &acc, &in[i].Z); }
    fp_inv(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        if (g1j_is_inf(&in[i])) { g1a_set_inf(&out[i]); continue; }
        fp_t zi, zi2, zi3; fp_mul(&zi, &acc, &pre[i]); fp_mul(&acc, &acc, &in[i].Z);
        fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
        fp_mul(&out[i].x, &in[i].X, &zi2); fp_mul(&out[i].y, &in[i].Y, &zi3);
    }
    free(pre);
}
static int g1j_eq(const g1j_t *p, const g1j_t *q) {
    int pi = g1j_is_inf(p), qi = g1j_is_inf(q);
    if (pi || qi) return pi && qi;
    fp_t z1, z2, a, b; fp_sqr(&z1, &p->Z); fp_sqr(&z2, &q->Z);
    fp_mul(&a, &p->X, &z2); fp_mul(&b, &q->X, &z1); if (!fp_eq(&a, &b)) return 0;
    fp_mul(&z1, &z1, &p->Z); fp_mul(&z2, &z2, &q->Z);
    fp_mul(&a, &p->Y, &z2); fp_mul(&b, &q->Y, &z1); return fp_eq(&a, &b);
}
static int g1a_on_curve(const g1a_t *p) {
    if (g1a_is_inf(p)) return 1;
    fp_t l, r, b; fp_sqr(&l, &p->y); fp_sqr(&r, &p->x); fp_mul(&r, &r, &p->x);
    memcpy(b.l, FP_B_MONT, 48); fp_add(&r, &r, &b); return fp_eq(&l, &r);
}

/* scalar multiplication by a little-endian multi-limb integer (`mul_bigint`): MSB-first double-and-add */
static void g1j_mul_limbs(g1j_t *r, const g1j_t *p, const uint64_t *k, int nlimbs) {
    g1j_t acc; g1j_set_inf(&acc);
    int started = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) g1j_dbl(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) { g1j_add(&acc, &acc, p); started = 1; }
    }
    *r = acc;
}
/* `P.mul(s)` for s in Fr (Montgomery form in, as arkworks holds it) */
static void g1j_mul_fr(g1j_t *r, const g1j_t *p, const fr_t *s) {
    uint64_t c[4]; fr_to_canon(c, s); g1j_mul_limbs(r, p, c, 4);
}
static void g1a_mul_fr(g1j_t *r, const g1a_t *p, const fr_t *s) {
    g1j_t j; g1j_from_affine(&j, p); g1j_mul_fr(r, &j, s);
}
static int g1a_in_subgroup(const g1a_t *p) {
    g1j_t j, r; g1j_from_affine(&j, p); g1j_mul_limbs(&r, &j, FR_R, 4); return g1j_is_inf(&r);
}

/* ---- 48-byte compressed encoding (big-endian x; bit7 compressed, bit6 infinity, bit5 y>-y) ---- */
static void g1a_compress(uint8_t out[48], const g1a_t *p) {
    if (g1a_is_inf(p)) { memset(out, 0, 48); out[0] = 0xC0; return; }
    uint64_t c[6]; fp_to_canon(c, &p->x);
    for (int i = 0; i < 48; i++) out[i] = (uint8_t)(c[5 - i / 8] >> (8 * (7 - i % 8)));
    fp_t ny; fp_neg(&ny, &p->y);
    out[0] |= 0x80;
    if (fp_canon_gt(&p->y, &ny)) out[0] |= 0x20;
}
static void g1j_compress(uint8_t out[48], const g1j_t *p) { g1a_t a; g1j_to_affine(&a, p); g1a_compress(out, &a); }
/* returns 0 on success; performs the on-curve and (optionally) subgroup checks that
 * `deserialize_compressed` performs */
static int g1a_decompress(g1a_t *p, const uint8_t in[48], int check_subgroup) {
    if (!(in[0] & 0x80)) return -1;
    if (in[0] & 0x40) {
        if (in[0] & 0x3F) return -1;
        for (int i = 1; i < 48; i++) if (in[i]) return -1;
        g1a_set_inf(p); return 0;
    }
    uint64_t c[6] = {0};
    for (int i = 0; i < 48; i++) {
        uint8_t b = in[i]; if (i == 0) b &= 0x1F;
        c[5 - i / 8] |= (uint64_t)b << (8 * (7 - i % 8));
    }
    if (limbs_geq(c, FP_P, 6)) return -1;
    fp_t x, rhs, b, y; fp_from_canon(&x, c);
    fp_sqr(&rhs, &x); fp_mul(&rhs, &rhs, &x); memcpy(b.l, FP_B_MONT, 48); fp_add(&rhs, &rhs, &b);
    if (!fp_sqrt(&y, &rhs)) return -1;
    fp_t ny; fp_neg(&ny, &y);
    int y_big = fp_canon_gt(&y, &ny);
    if (y_big != ((in[0] >> 5) & 1)) y = ny;
    p->x = x; p->y = y;
    if (check_subgroup && !g1a_in_subgroup(p)) return -1;
    return 0;
}

/* ---- Pippenger MSM following the ark-ec 0.4 `VariableBaseMSM` shape: signed radix-2^c digits,
 * window c = 3 for n < 32 else ceil(log2 n)*69/100 + 2, 2^(c-1) buckets per window, one task per
 * window (ark's rayon per-window parallelism -> OpenMP here), running-sum bucket reduction,
 * windows combined high->low.  Scalars are canonical 4x64 integers (`into_bigint`). ---- */
static int msm_window_bits(size_t n) {
    if (n < 32) return 3;
    int lg = 0; while (((size_t)1 << lg) < n) lg++;
    return lg * 69 / 100 + 2;
}
static void g1_msm_canon(g1j_t *out, const g1a_t *bases, const uint64_t (*scalars)[4], size_t n, int threads) {
    if (n == 0) { g1j_set_inf(out); return; }
    const int c = msm_window_bits(n);
    const int nwin = (255 + c - 1) / c;
    const size_t nb = (size_t)1 << (c - 1);
    int32_t *digits = (int32_t *)malloc(sizeof(int32_t) * n * nwin);
    for (size_t i = 0; i < n; i++) {
        uint64_t carry = 0;
        for (int w = 0; w < nwin; w++) {
            int bit = w * c, li = bit / 64, sh = bit % 64;
            uint64_t v = scalars[i][li] >> sh;
            if (sh + c > 64 && li + 1 < 4) v |= scalars[i][li + 1] << (64 - sh);
            v = (v & (((uint64_t)1 << c) - 1)) + carry;
            carry = (v + ((uint64_t)1 << (c - 1))) >> c;
            int64_t d = (int64_t)v - (int64_t)(carry << c);
            if (w == nwin - 1) d += (int64_t)(carry << c);
            digits[i * nwin + w] = (int32_t)d;
        }
    }
    g1j_t *wsum = (g1j_t *)malloc(sizeof(g1j_t) * nwin);
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (int w = 0; w < nwin; w++) {
        /* the top window may hold digits up to 2^c (unrecentred carry) -> size buckets generously */
        size_t nbw = (w == nwin - 1) ? ((size_t)1 << c) + 1 : nb;
        g1j_t *bk = (g1j_t *)malloc(sizeof(g1j_t) * nbw);
        for (size_t b = 0; b < nbw; b++) g1j_set_inf(&bk[b]);
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * nwin + w];
            if (d > 0) g1j_add_affine(&bk[d - 1], &bk[d - 1], &bases[i]);
            else if (d < 0) { g1a_t nq; g1a_neg(&nq, &bases[i]); g1j_add_affine(&bk[-d - 1], &bk[-d - 1], &nq); }
        }
        g1j_t run, acc; g1j_set_inf(&run); g1j_set_inf(&acc);
        for (size_t b = nbw; b-- > 0;) { g1j_add(&run, &run, &bk[b]); g1j_add(&acc, &acc, &run); }
        wsum[w] = acc;
        free(bk);
    }
    g1j_t total = wsum[nwin - 1];
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) g1j_dbl(&total, &total);
        g1j_add(&total, &total, &wsum[w]);
    }
    *out = total;
    free(wsum); free(digits);
}
/* `util::msm(points, scalars)` with Fr scalars in Montgomery form -- src/util.rs:19-22 */
static void g1_msm(g1j_t *out, const g1a_t *bases, const fr_t *scalars, size_t n, int threads) {
    uint64_t (*c)[4] = (uint64_t (*)[4])malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) fr_to_canon(c[i], &scalars[i]);
    g1_msm_canon(out, bases, (const uint64_t (*)[4])c, n, threads);
    free(c);
}
/* `util::msm_from_projective` -- src/util.rs:25-29 */
static void g1_msm_from_projective(g1j_t *out, const g1j_t *pts, const fr_t *scalars, size_t n, int threads) {
    g1a_t *aff = (g1a_t *)malloc(sizeof(g1a_t) * (n ? n : 1));
    g1j_batch_to_affine(aff, pts, n);
    g1_msm(out, aff, scalars, n, threads);
    free(aff);
}
#endif
