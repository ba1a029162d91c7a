/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Sequential, single-proof restatement of the Curdleproofs prover and verifier, following the
 * reference file by file (cited at each function).  It exists to (1) reproduce the reference's
 * seed-0 golden proofs so that the arithmetic/transcript/RNG restatements above are PINNED, and
 * (2) act as the checker for the B200 engine: same inputs, same RNG stream => byte-identical proof.
 * It is also the "port" CPU baseline timed by bench.py.
 */
#ifndef CDP_ORACLE_PROTOCOL_H
#define CDP_ORACLE_PROTOCOL_H
#include "g1.h"
#include "merlin.h"
#include "rng.h"

#define N_BLINDERS 4 /* src/lib.rs:35 */
#define MAX_LOG_N 24

/* ----------------------------------------------------------------- transcript glue: src/transcript.rs:28-61 */
static void fr_to_bytes(uint8_t out[32], const fr_t *a) { uint64_t c[4]; fr_to_canon(c, a); memcpy(out, c, 32); }
static void t_append_g1j(transcript_t *t, const char *label, const g1j_t *p) { uint8_t b[48]; g1j_compress(b, p); transcript_append_message(t, label, b, 48); }
static void t_append_g1a(transcript_t *t, const char *label, const g1a_t *p) { uint8_t b[48]; g1a_compress(b, p); transcript_append_message(t, label, b, 48); }
static void t_append_fr(transcript_t *t, const char *label, const fr_t *a) { uint8_t b[32]; fr_to_bytes(b, a); transcript_append_message(t, label, b, 32); }
/* a `Vec<_>` item serialises as u64-LE length followed by the elements (ark-serialize) */
static void t_append_g1a_vec(transcript_t *t, const char *label, const g1a_t *v, size_t n) {
    uint8_t *buf = (uint8_t *)malloc(8 + 48 * n);
    uint64_t len = n; memcpy(buf, &len, 8);
    for (size_t i = 0; i < n; i++) g1a_compress(buf + 8 + 48 * i, &v[i]);
    transcript_append_message(t, label, buf, 8 + 48 * n); free(buf);
}
static void t_append_fr_vec(transcript_t *t, const char *label, const fr_t *v, size_t n) {
    uint8_t *buf = (uint8_t *)malloc(8 + 32 * n);
    uint64_t len = n; memcpy(buf, &len, 8);
    for (size_t i = 0; i < n; i++) fr_to_bytes(buf + 8 + 32 * i, &v[i]);
    transcript_append_message(t, label, buf, 8 + 32 * n); free(buf);
}
/* get_and_append_challenge, src/transcript.rs:41-54.  `Fr::from_random_bytes(&[u8;64])` keeps the
 * first 32 bytes, clears bit 255 and fails when the value is >= r. */
static void t_challenge(transcript_t *t, const char *label, fr_t *out) {
    for (;;) {
        uint8_t buf[64]; transcript_challenge_bytes(t, label, buf, 64);
        uint64_t c[4]; memcpy(c, buf, 32); c[3] &= 0x7FFFFFFFFFFFFFFFULL;
        if (limbs_geq(c, FR_R, 4) || limbs_is_zero(c, 4)) continue;
        fr_from_canon(out, c);
        t_append_fr(t, label, out);
        return;
    }
}

/* ----------------------------------------------------------------- CRS: src/crs.rs:19-74 */
typedef struct {
    size_t ell;
    g1a_t *vec_G, *vec_H; /* ell, N_BLINDERS */
    g1j_t H, G_t, G_u;
    g1a_t G_sum, H_sum;
} crs_t;
static void sum_affine(g1a_t *out, const g1a_t *v, size_t n) { /* src/util.rs:108-114 */
    g1j_t acc; g1j_set_inf(&acc);
    for (size_t i = 0; i < n; i++) g1j_add_affine(&acc, &acc, &v[i]);
    g1j_to_affine(out, &acc);
}
static void crs_from_points(crs_t *crs, size_t ell, const g1a_t *pts /* ell+7 */) { /* src/crs.rs:37-58 */
    crs->ell = ell;
    crs->vec_G = (g1a_t *)malloc(sizeof(g1a_t) * ell);
    crs->vec_H = (g1a_t *)malloc(sizeof(g1a_t) * N_BLINDERS);
    memcpy(crs->vec_G, pts, sizeof(g1a_t) * ell);
    memcpy(crs->vec_H, pts + ell, sizeof(g1a_t) * N_BLINDERS);
    sum_affine(&crs->G_sum, crs->vec_G, ell);
    sum_affine(&crs->H_sum, crs->vec_H, N_BLINDERS);
    size_t n = ell + N_BLINDERS;
    g1j_from_affine(&crs->H, &pts[n]); g1j_from_affine(&crs->G_t, &pts[n + 1]); g1j_from_affine(&crs->G_u, &pts[n + 2]);
}
static void crs_generate(crs_t *crs, size_t ell) { /* src/crs.rs:61-69: its own seed-0 StdRng */
    size_t np = ell + N_BLINDERS + 3;
    stdrng_t rng; stdrng_seed_from_u64(&rng, 0);
    g1j_t *pj = (g1j_t *)malloc(sizeof(g1j_t) * np);
    g1a_t *pa = (g1a_t *)malloc(sizeof(g1a_t) * np);
    for (size_t i = 0; i < np; i++) g1j_rand(&pj[i], &rng);
    g1j_batch_to_affine(pa, pj, np);
    crs_from_points(crs, ell, pa);
    free(pj); free(pa);
}
static void crs_free(crs_t *crs) { free(crs->vec_G); free(crs->vec_H); }

/* ----------------------------------------------------------------- GroupCommitment: src/commitments.rs:44-79 */
typedef struct { g1j_t T_1, T_2; } gcomm_t;
static void gcomm_new(gcomm_t *c, const g1j_t *G, const g1j_t *H, const g1j_t *T, const fr_t *r) {
    g1j_t t; g1j_mul_fr(&c->T_1, G, r); g1j_mul_fr(&t, H, r); g1j_add(&c->T_2, T, &t);
}
static void gcomm_add(gcomm_t *r, const gcomm_t *a, const gcomm_t *b) { g1j_add(&r->T_1, &a->T_1, &b->T_1); g1j_add(&r->T_2, &a->T_2, &b->T_2); }
static void gcomm_mul(gcomm_t *r, const gcomm_t *a, const fr_t *s) { g1j_mul_fr(&r->T_1, &a->T_1, s); g1j_mul_fr(&r->T_2, &a->T_2, s); }
static int gcomm_eq(const gcomm_t *a, const gcomm_t *b) { return g1j_eq(&a->T_1, &b->T_1) && g1j_eq(&a->T_2, &b->T_2); }

/* ----------------------------------------------------------------- MsmAccumulator: src/msm_accumulator.rs:22-68 */
typedef struct { g1a_t base; fr_t scalar; int used; } acc_slot_t;
typedef struct { g1j_t A_c; acc_slot_t *slots; size_t cap, count; } msm_acc_t;
static void acc_init(msm_acc_t *a, size_t expected) {
    g1j_set_inf(&a->A_c); a->cap = 64; while (a->cap < 4 * expected) a->cap <<= 1;
    a->slots = (acc_slot_t *)calloc(a->cap, sizeof(acc_slot_t)); a->count = 0;
}
static void acc_free(msm_acc_t *a) { free(a->slots); }
static void acc_accumulate_check(msm_acc_t *a, const g1j_t *C, const fr_t *vec_x, const g1a_t *vec_V, size_t n, stdrng_t *rng) {
    fr_t rf; fr_rand(&rf, rng);                               /* :44 */
    g1j_t t; g1j_mul_fr(&t, C, &rf); g1j_add(&a->A_c, &a->A_c, &t); /* :46 */
    for (size_t i = 0; i < n; i++) {                          /* :48-51 */
        size_t h = (size_t)(vec_V[i].x.l[0] * 0x9E3779B97F4A7C15ULL ^ vec_V[i].y.l[1]) & (a->cap - 1);
        while (a->slots[h].used && memcmp(&a->slots[h].base, &vec_V[i], sizeof(g1a_t)) != 0) h = (h + 1) & (a->cap - 1);
        if (!a->slots[h].used) { a->slots[h].used = 1; a->slots[h].base = vec_V[i]; fr_zero(&a->slots[h].scalar); a->count++; }
        fr_t p; fr_mul(&p, &rf, &vec_x[i]); fr_add(&a->slots[h].scalar, &a->slots[h].scalar, &p);
    }
}
/* returns 1 when all accumulated checks hold (:55-68); optionally exports the final MSM instance */
static int acc_verify(msm_acc_t *a, int threads, g1a_t **out_bases, fr_t **out_scalars, size_t *out_n) {
    g1a_t *bases = (g1a_t *)malloc(sizeof(g1a_t) * (a->count ? a->count : 1));
    fr_t *scalars = (fr_t *)malloc(sizeof(fr_t) * (a->count ? a->count : 1));
    size_t k = 0;
    for (size_t i = 0; i < a->cap; i++) if (a->slots[i].used) { bases[k] = a->slots[i].base; scalars[k] = a->slots[i].scalar; k++; }
    g1j_t m, d; g1_msm(&m, bases, scalars, k, threads); g1j_sub(&d, &m, &a->A_c);
    if (out_bases) { *out_bases = bases; *out_scalars = scalars; *out_n = k; } else { free(bases); free(scalars); }
    return g1j_is_inf(&d);
}

/* ----------------------------------------------------------------- InnerProductProof: src/inner_product_argument.rs */
typedef struct {
    g1j_t B_c, B_d;
    g1j_t L_C[MAX_LOG_N], R_C[MAX_LOG_N], L_D[MAX_LOG_N], R_D[MAX_LOG_N];
    fr_t c_final, d_final; int m;
} ipa_proof_t;

/* generate_ipa_blinders :42-82 */
static void ipa_blinders(fr_t *r, fr_t *z, const fr_t *c, const fr_t *d, size_t n, stdrng_t *rng) {
    for (size_t i = 0; i < n; i++) fr_rand(&r[i], rng);
    for (size_t i = 0; i + 2 < n; i++) fr_rand(&z[i], rng);
    fr_t omega, delta, t, inv_c, num, den, last_z, pen_z;
    fr_inner_product(&omega, r, d, n); fr_inner_product(&t, z, c, n - 2); fr_add(&omega, &omega, &t);
    fr_inner_product(&delta, r, z, n - 2);
    fr_inv(&inv_c, &c[n - 2]);
    fr_mul(&num, &r[n - 2], &inv_c); fr_mul(&num, &num, &omega); fr_sub(&num, &num, &delta);
    fr_neg(&den, &r[n - 2]); fr_mul(&den, &den, &inv_c); fr_mul(&den, &den, &c[n - 1]); fr_add(&den, &den, &r[n - 1]);
    fr_inv(&den, &den); fr_mul(&last_z, &num, &den);
    fr_mul(&t, &last_z, &c[n - 1]); fr_add(&t, &t, &omega); fr_neg(&pen_z, &inv_c); fr_mul(&pen_z, &pen_z, &t);
    z[n - 2] = pen_z; z[n - 1] = last_z;
}
/* InnerProductProof::new :98-198.  G, Gp, c, d are consumed (folded in place). */
static void ipa_prove(ipa_proof_t *pf, g1a_t *G, g1a_t *Gp, const g1j_t *crs_H, const g1j_t *C, const g1j_t *D,
                      const fr_t *z, fr_t *c, fr_t *d, size_t n, transcript_t *tr, stdrng_t *rng, int threads) {
    fr_t *r_c = (fr_t *)malloc(sizeof(fr_t) * n), *r_d = (fr_t *)malloc(sizeof(fr_t) * n);
    ipa_blinders(r_c, r_d, c, d, n, rng);                                     /* :124 */
    g1_msm(&pf->B_c, G, r_c, n, threads); g1_msm(&pf->B_d, Gp, r_d, n, threads); /* :126-127 */
    t_append_g1j(tr, "ipa_step1", C); t_append_g1j(tr, "ipa_step1", D);
    t_append_fr(tr, "ipa_step1", z);
    t_append_g1j(tr, "ipa_step1", &pf->B_c); t_append_g1j(tr, "ipa_step1", &pf->B_d);
    fr_t alpha, beta; t_challenge(tr, "ipa_alpha", &alpha); t_challenge(tr, "ipa_beta", &beta);
    for (size_t i = 0; i < n; i++) {                                          /* :136-139 */
        fr_t t; fr_mul(&t, &alpha, &c[i]); fr_add(&c[i], &r_c[i], &t);
        fr_mul(&t, &alpha, &d[i]); fr_add(&d[i], &r_d[i], &t);
    }
    g1j_t H; g1j_mul_fr(&H, crs_H, &beta);                                    /* :140 */
    int round = 0;
    while (n > 1) {                                                           /* :150-186 */
        n /= 2;
        fr_t *c_L = c, *c_R = c + n, *d_L = d, *d_R = d + n;
        g1a_t *G_L = G, *G_R = G + n, *Gp_L = Gp, *Gp_R = Gp + n;
        g1j_t L_C, L_D, R_C, R_D, t; fr_t ip;
        g1_msm(&L_C, G_R, c_L, n, threads); fr_inner_product(&ip, c_L, d_R, n); g1j_mul_fr(&t, &H, &ip); g1j_add(&L_C, &L_C, &t);
        g1_msm(&L_D, Gp_L, d_R, n, threads);
        g1_msm(&R_C, G_L, c_R, n, threads); fr_inner_product(&ip, c_R, d_L, n); g1j_mul_fr(&t, &H, &ip); g1j_add(&R_C, &R_C, &t);
        g1_msm(&R_D, Gp_R, d_L, n, threads);
        pf->L_C[round] = L_C; pf->L_D[round] = L_D; pf->R_C[round] = R_C; pf->R_D[round] = R_D;
        t_append_g1j(tr, "ipa_loop", &L_C); t_append_g1j(tr, "ipa_loop", &L_D);
        t_append_g1j(tr, "ipa_loop", &R_C); t_append_g1j(tr, "ipa_loop", &R_D);
        fr_t gamma, gamma_inv; t_challenge(tr, "ipa_gamma", &gamma); fr_inv(&gamma_inv, &gamma);
        for (size_t i = 0; i < n; i++) {                                      /* :174-179 (the fold) */
            fr_t u; fr_mul(&u, &gamma_inv, &c_R[i]); fr_add(&c_L[i], &c_L[i], &u);
            fr_mul(&u, &gamma, &d_R[i]); fr_add(&d_L[i], &d_L[i], &u);
            g1j_t s; g1a_mul_fr(&s, &G_R[i], &gamma); g1j_add_affine(&s, &s, &G_L[i]); g1j_to_affine(&G_L[i], &s);
            g1a_mul_fr(&s, &Gp_R[i], &gamma_inv); g1j_add_affine(&s, &s, &Gp_L[i]); g1j_to_affine(&Gp_L[i], &s);
        }
        round++;
    }
    pf->m = round; pf->c_final = c[0]; pf->d_final = d[0];
    free(r_c); free(r_d);
}
/* get_verification_scalars_bitstring (src/util.rs:40-64) + the product loop
 * (src/inner_product_argument.rs:237-243): s_i = prod_{j : bit (m-1-j) of i set} gamma_j */
static void verification_scalars_s(fr_t *s, const fr_t *gamma, int m, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fr_one(&s[i]);
        for (int j = 0; j < m; j++) if ((i >> (m - 1 - j)) & 1) fr_mul(&s[i], &s[i], &gamma[j]);
    }
}
/* InnerProductProof::verify :264-326; returns 0 on Ok */
static int ipa_verify(const ipa_proof_t *pf, const g1a_t *G, size_t n, const g1j_t *crs_H, const g1j_t *C, const g1j_t *D,
                      const fr_t *z, const fr_t *vec_u, transcript_t *tr, msm_acc_t *acc, stdrng_t *rng, int threads) {
    t_append_g1j(tr, "ipa_step1", C); t_append_g1j(tr, "ipa_step1", D);
    t_append_fr(tr, "ipa_step1", z);
    t_append_g1j(tr, "ipa_step1", &pf->B_c); t_append_g1j(tr, "ipa_step1", &pf->B_d);
    fr_t alpha, beta; t_challenge(tr, "ipa_alpha", &alpha); t_challenge(tr, "ipa_beta", &beta);
    int m = pf->m;
    if (m >= 32 || n != ((size_t)1 << m)) return -1;                           /* :208-214 */
    fr_t gam[MAX_LOG_N], gam_inv[MAX_LOG_N], scratch[MAX_LOG_N];
    for (int i = 0; i < m; i++) {
        t_append_g1j(tr, "ipa_loop", &pf->L_C[i]); t_append_g1j(tr, "ipa_loop", &pf->L_D[i]);
        t_append_g1j(tr, "ipa_loop", &pf->R_C[i]); t_append_g1j(tr, "ipa_loop", &pf->R_D[i]);
        t_challenge(tr, "ipa_gamma", &gam[i]); gam_inv[i] = gam[i];
    }
    fr_batch_inv(gam_inv, m, scratch);
    fr_t *s = (fr_t *)malloc(sizeof(fr_t) * n), *s_inv = (fr_t *)malloc(sizeof(fr_t) * n), *sc = (fr_t *)malloc(sizeof(fr_t) * n);
    verification_scalars_s(s, gam, m, n);
    memcpy(s_inv, s, sizeof(fr_t) * n); fr_batch_inv(s_inv, n, sc);
    /* first check :289-309 */
    fr_t *rhs = (fr_t *)malloc(sizeof(fr_t) * (n + 1));
    g1a_t *GH = (g1a_t *)malloc(sizeof(g1a_t) * (n + 1));
    for (size_t i = 0; i < n; i++) fr_mul(&rhs[i], &pf->c_final, &s[i]);
    fr_mul(&rhs[n], &pf->c_final, &pf->d_final); fr_mul(&rhs[n], &rhs[n], &beta);
    memcpy(GH, G, sizeof(g1a_t) * n); g1j_to_affine(&GH[n], crs_H);
    g1j_t H, C_a, t, lhs; fr_t aaz;
    g1j_mul_fr(&H, crs_H, &beta);
    g1j_mul_fr(&t, C, &alpha); g1j_add(&C_a, &pf->B_c, &t);
    fr_mul(&aaz, &alpha, &alpha); fr_mul(&aaz, &aaz, z); g1j_mul_fr(&t, &H, &aaz); g1j_add(&C_a, &C_a, &t);
    g1_msm_from_projective(&lhs, pf->L_C, gam, m, threads); g1j_add(&lhs, &lhs, &C_a);
    g1_msm_from_projective(&t, pf->R_C, gam_inv, m, threads); g1j_add(&lhs, &lhs, &t);
    acc_accumulate_check(acc, &lhs, rhs, GH, n + 1, rng);
    /* second check :311-323 */
    for (size_t i = 0; i < n; i++) { fr_mul(&rhs[i], &s_inv[i], &vec_u[i]); fr_mul(&rhs[i], &pf->d_final, &rhs[i]); }
    g1j_t D_a; g1j_mul_fr(&t, D, &alpha); g1j_add(&D_a, &pf->B_d, &t);
    g1_msm_from_projective(&lhs, pf->L_D, gam, m, threads); g1j_add(&lhs, &lhs, &D_a);
    g1_msm_from_projective(&t, pf->R_D, gam_inv, m, threads); g1j_add(&lhs, &lhs, &t);
    acc_accumulate_check(acc, &lhs, rhs, G, n, rng);
    free(s); free(s_inv); free(sc); free(rhs); free(GH);
    return 0;
}

/* ----------------------------------------------------------------- GrandProductProof: src/grand_product_argument.rs */
typedef struct { g1j_t C; fr_t r_p; ipa_proof_t ipa; } gprod_proof_t;
/* ::new :43-166 */
static void gprod_prove(gprod_proof_t *pf, const g1a_t *crs_G, size_t ell, const g1a_t *crs_Hv, const g1j_t *crs_U,
                        const g1j_t *B, const fr_t *gprod_result, const fr_t *vec_b, const fr_t *vec_b_blinders,
                        transcript_t *tr, stdrng_t *rng, int threads) {
    const size_t nb = N_BLINDERS, n = ell + nb;
    t_append_g1j(tr, "gprod_step1", B); t_append_fr(tr, "gprod_step1", gprod_result);
    fr_t alpha; t_challenge(tr, "gprod_alpha", &alpha);
    fr_t *vec_c = (fr_t *)malloc(sizeof(fr_t) * n), *vec_d = (fr_t *)malloc(sizeof(fr_t) * n);
    fr_one(&vec_c[0]);
    for (size_t i = 0; i + 1 < ell; i++) fr_mul(&vec_c[i + 1], &vec_c[i], &vec_b[i]);      /* :69-73 */
    fr_t c_bl[N_BLINDERS]; for (size_t i = 0; i < nb; i++) fr_rand(&c_bl[i], rng);          /* :75 */
    g1j_t t; g1_msm(&pf->C, crs_G, vec_c, ell, threads); g1_msm(&t, crs_Hv, c_bl, nb, threads); g1j_add(&pf->C, &pf->C, &t);
    fr_t rb_alpha[N_BLINDERS];
    for (size_t i = 0; i < nb; i++) fr_add(&rb_alpha[i], &vec_b_blinders[i], &alpha);
    fr_inner_product(&pf->r_p, rb_alpha, c_bl, nb);
    t_append_g1j(tr, "gprod_step2", &pf->C); t_append_fr(tr, "gprod_step2", &pf->r_p);
    fr_t beta, beta_inv; t_challenge(tr, "gprod_beta", &beta); fr_inv(&beta_inv, &beta);
    /* G' and H' :90-102 */
    g1a_t *vec_G = (g1a_t *)malloc(sizeof(g1a_t) * n), *vec_Gp = (g1a_t *)malloc(sizeof(g1a_t) * n);
    fr_t pw = beta_inv;
    for (size_t i = 0; i < ell; i++) { g1j_t s; g1a_mul_fr(&s, &crs_G[i], &pw); g1j_to_affine(&vec_Gp[i], &s); fr_mul(&pw, &pw, &beta_inv); }
    fr_t beta_inv_l1, beta_l1, beta_l; fr_pow_u64(&beta_inv_l1, &beta_inv, ell + 1); fr_pow_u64(&beta_l1, &beta, ell + 1); fr_pow_u64(&beta_l, &beta, ell);
    for (size_t i = 0; i < nb; i++) { g1j_t s; g1a_mul_fr(&s, &crs_Hv[i], &beta_inv_l1); g1j_to_affine(&vec_Gp[ell + i], &s); }
    /* b', d, beta powers :104-121 */
    fr_t *beta_pows = (fr_t *)malloc(sizeof(fr_t) * ell);
    fr_t pb = beta, p1; fr_one(&p1);
    for (size_t i = 0; i < ell; i++) {
        fr_t bp; fr_mul(&bp, &vec_b[i], &pb); fr_mul(&pb, &pb, &beta);
        fr_sub(&vec_d[i], &bp, &p1); beta_pows[i] = p1; fr_mul(&p1, &p1, &beta);
    }
    for (size_t i = 0; i < nb; i++) fr_mul(&vec_d[ell + i], &beta_l1, &rb_alpha[i]);          /* :124-127 */
    fr_t ab[N_BLINDERS]; for (size_t i = 0; i < nb; i++) fr_mul(&ab[i], &alpha, &beta_l1);    /* :130-131 */
    g1j_t D, m1, m2; g1_msm(&m1, vec_Gp, beta_pows, ell, threads); g1_msm(&m2, vec_Gp + ell, ab, nb, threads);
    g1j_sub(&D, B, &m1); g1j_add(&D, &D, &m2);                                               /* :132 */
    memcpy(vec_G, crs_G, sizeof(g1a_t) * ell); memcpy(vec_G + ell, crs_Hv, sizeof(g1a_t) * nb);
    fr_t inner_prod, u, one; fr_one(&one);
    fr_mul(&inner_prod, &pf->r_p, &beta_l1); fr_mul(&u, gprod_result, &beta_l); fr_add(&inner_prod, &inner_prod, &u); fr_sub(&inner_prod, &inner_prod, &one);
    memcpy(vec_c + ell, c_bl, sizeof(fr_t) * nb);
    ipa_prove(&pf->ipa, vec_G, vec_Gp, crs_U, &pf->C, &D, &inner_prod, vec_c, vec_d, n, tr, rng, threads);
    free(vec_c); free(vec_d); free(vec_G); free(vec_Gp); free(beta_pows);
}
/* ::verify :180-246 */
static int gprod_verify(const gprod_proof_t *pf, const g1a_t *crs_G, size_t ell, const g1a_t *crs_Hv, const g1j_t *crs_U,
                        const g1a_t *G_sum, const g1a_t *H_sum, const g1j_t *B, const fr_t *gprod_result,
                        transcript_t *tr, msm_acc_t *acc, stdrng_t *rng, int threads) {
    const size_t nb = N_BLINDERS, n = ell + nb;
    t_append_g1j(tr, "gprod_step1", B); t_append_fr(tr, "gprod_step1", gprod_result);
    fr_t alpha; t_challenge(tr, "gprod_alpha", &alpha);
    t_append_g1j(tr, "gprod_step2", &pf->C); t_append_fr(tr, "gprod_step2", &pf->r_p);
    fr_t beta, beta_inv; t_challenge(tr, "gprod_beta", &beta); fr_inv(&beta_inv, &beta);
    fr_t *vec_u = (fr_t *)malloc(sizeof(fr_t) * n);
    fr_t pw = beta_inv;
    for (size_t i = 0; i < ell; i++) { vec_u[i] = pw; fr_mul(&pw, &pw, &beta_inv); }
    fr_t beta_inv_l1, beta_l1, beta_l; fr_pow_u64(&beta_inv_l1, &beta_inv, ell + 1); fr_pow_u64(&beta_l1, &beta, ell + 1); fr_pow_u64(&beta_l, &beta, ell);
    for (size_t i = 0; i < nb; i++) vec_u[ell + i] = beta_inv_l1;
    g1j_t D, t; g1a_mul_fr(&t, G_sum, &beta_inv); g1j_sub(&D, B, &t); g1a_mul_fr(&t, H_sum, &alpha); g1j_add(&D, &D, &t); /* :223 */
    g1a_t *vec_G = (g1a_t *)malloc(sizeof(g1a_t) * n);
    memcpy(vec_G, crs_G, sizeof(g1a_t) * ell); memcpy(vec_G + ell, crs_Hv, sizeof(g1a_t) * nb);
    fr_t inner_prod, u, one; fr_one(&one);
    fr_mul(&inner_prod, &pf->r_p, &beta_l1); fr_mul(&u, gprod_result, &beta_l); fr_add(&inner_prod, &inner_prod, &u); fr_sub(&inner_prod, &inner_prod, &one);
    int rc = ipa_verify(&pf->ipa, vec_G, n, crs_U, &pf->C, &D, &inner_prod, vec_u, tr, acc, rng, threads);
    free(vec_u); free(vec_G);
    return rc;
}

/* ----------------------------------------------------------------- SamePermutationProof: src/same_permutation_argument.rs */
typedef struct { g1j_t B; gprod_proof_t gprod; } sameperm_proof_t;
/* ::new :40-99 */
static void sameperm_prove(sameperm_proof_t *pf, const g1a_t *crs_G, size_t ell, const g1a_t *crs_Hv, const g1j_t *crs_U,
                           const g1j_t *A, const g1j_t *M, const fr_t *vec_a, const uint32_t *perm,
                           const fr_t *vec_a_blinders, const fr_t *vec_m_blinders, transcript_t *tr, stdrng_t *rng, int threads) {
    t_append_g1j(tr, "same_perm_step1", A); t_append_g1j(tr, "same_perm_step1", M);
    t_append_fr_vec(tr, "same_perm_step1", vec_a, ell);
    fr_t alpha, beta; t_challenge(tr, "same_perm_alpha", &alpha); t_challenge(tr, "same_perm_beta", &beta);
    fr_t *factors = (fr_t *)malloc(sizeof(fr_t) * ell), *beta_rep = (fr_t *)malloc(sizeof(fr_t) * ell);
    fr_t prod; fr_one(&prod);
    for (size_t i = 0; i < ell; i++) {                                           /* :66-73 */
        fr_t m, t; fr_from_u64(&m, perm[i]); fr_mul(&t, &m, &alpha);
        fr_add(&factors[i], &vec_a[perm[i]], &t); fr_add(&factors[i], &factors[i], &beta);
        fr_mul(&prod, &prod, &factors[i]); beta_rep[i] = beta;
    }
    g1j_t t; g1j_mul_fr(&t, M, &alpha); g1j_add(&pf->B, A, &t);
    g1_msm(&t, crs_G, beta_rep, ell, threads); g1j_add(&pf->B, &pf->B, &t);       /* :75-76 */
    fr_t b_bl[N_BLINDERS];
    for (int i = 0; i < N_BLINDERS; i++) { fr_t u; fr_mul(&u, &alpha, &vec_m_blinders[i]); fr_add(&b_bl[i], &vec_a_blinders[i], &u); }
    gprod_prove(&pf->gprod, crs_G, ell, crs_Hv, crs_U, &pf->B, &prod, factors, b_bl, tr, rng, threads);
    free(factors); free(beta_rep);
}
/* ::verify :112-171 */
static int sameperm_verify(const sameperm_proof_t *pf, const g1a_t *crs_G, size_t ell, const g1a_t *crs_Hv, const g1j_t *crs_U,
                           const g1a_t *G_sum, const g1a_t *H_sum, const g1j_t *A, const g1j_t *M, const fr_t *vec_a,
                           transcript_t *tr, msm_acc_t *acc, stdrng_t *rng, int threads) {
    t_append_g1j(tr, "same_perm_step1", A); t_append_g1j(tr, "same_perm_step1", M);
    t_append_fr_vec(tr, "same_perm_step1", vec_a, ell);
    fr_t alpha, beta; t_challenge(tr, "same_perm_alpha", &alpha); t_challenge(tr, "same_perm_beta", &beta);
    fr_t prod; fr_one(&prod);
    fr_t *beta_rep = (fr_t *)malloc(sizeof(fr_t) * ell);
    for (size_t i = 0; i < ell; i++) {
        fr_t m, t, f; fr_from_u64(&m, i); fr_mul(&t, &m, &alpha); fr_add(&f, &vec_a[i], &t); fr_add(&f, &f, &beta);
        fr_mul(&prod, &prod, &f); beta_rep[i] = beta;
    }
    g1j_t lhs, t; g1j_sub(&lhs, &pf->B, A); g1j_mul_fr(&t, M, &alpha); g1j_sub(&lhs, &lhs, &t);
    acc_accumulate_check(acc, &lhs, beta_rep, crs_G, ell, rng);                   /* :149-154 */
    free(beta_rep);
    return gprod_verify(&pf->gprod, crs_G, ell, crs_Hv, crs_U, G_sum, H_sum, &pf->B, &prod, tr, acc, rng, threads);
}

/* ----------------------------------------------------------------- SameScalarProof: src/same_scalar_argument.rs */
typedef struct { gcomm_t cm_A, cm_B; fr_t z_k, z_t, z_u; } samescalar_proof_t;
static void samescalar_transcript(transcript_t *tr, const g1j_t *R, const g1j_t *S, const gcomm_t *cm_T, const gcomm_t *cm_U,
                                  const gcomm_t *cm_A, const gcomm_t *cm_B, fr_t *alpha) {
    const g1j_t *pts[10] = {R, S, &cm_T->T_1, &cm_T->T_2, &cm_U->T_1, &cm_U->T_2, &cm_A->T_1, &cm_A->T_2, &cm_B->T_1, &cm_B->T_2};
    for (int i = 0; i < 10; i++) t_append_g1j(tr, "sameexp_points", pts[i]);
    t_challenge(tr, "same_scalar_alpha", alpha);
}
/* ::new :39-84 */
static void samescalar_prove(samescalar_proof_t *pf, const g1j_t *G_t, const g1j_t *G_u, const g1j_t *H, const g1j_t *R, const g1j_t *S,
                             const gcomm_t *cm_T, const gcomm_t *cm_U, const fr_t *k, const fr_t *r_t, const fr_t *r_u,
                             transcript_t *tr, stdrng_t *rng) {
    fr_t r_a, r_b, r_k; fr_rand(&r_a, rng); fr_rand(&r_b, rng); fr_rand(&r_k, rng);
    g1j_t t; g1j_mul_fr(&t, R, &r_k); gcomm_new(&pf->cm_A, G_t, H, &t, &r_a);
    g1j_mul_fr(&t, S, &r_k); gcomm_new(&pf->cm_B, G_u, H, &t, &r_b);
    fr_t alpha, u; samescalar_transcript(tr, R, S, cm_T, cm_U, &pf->cm_A, &pf->cm_B, &alpha);
    fr_mul(&u, k, &alpha); fr_add(&pf->z_k, &r_k, &u);
    fr_mul(&u, r_t, &alpha); fr_add(&pf->z_t, &r_a, &u);
    fr_mul(&u, r_u, &alpha); fr_add(&pf->z_u, &r_b, &u);
}
/* ::verify :96-137 */
static int samescalar_verify(const samescalar_proof_t *pf, const g1j_t *G_t, const g1j_t *G_u, const g1j_t *H, const g1j_t *R, const g1j_t *S,
                             const gcomm_t *cm_T, const gcomm_t *cm_U, transcript_t *tr) {
    fr_t alpha; samescalar_transcript(tr, R, S, cm_T, cm_U, &pf->cm_A, &pf->cm_B, &alpha);
    gcomm_t e1, e2, l1, l2; g1j_t t;
    g1j_mul_fr(&t, R, &pf->z_k); gcomm_new(&e1, G_t, H, &t, &pf->z_t);
    g1j_mul_fr(&t, S, &pf->z_k); gcomm_new(&e2, G_u, H, &t, &pf->z_u);
    gcomm_mul(&l1, cm_T, &alpha); gcomm_add(&l1, &pf->cm_A, &l1);
    gcomm_mul(&l2, cm_U, &alpha); gcomm_add(&l2, &pf->cm_B, &l2);
    return (gcomm_eq(&l1, &e1) && gcomm_eq(&l2, &e2)) ? 0 : -1;
}

/* ----------------------------------------------------------------- SameMultiscalarProof: src/same_multiscalar_argument.rs */
typedef struct {
    g1j_t B_a, B_t, B_u;
    g1j_t L_A[MAX_LOG_N], L_T[MAX_LOG_N], L_U[MAX_LOG_N], R_A[MAX_LOG_N], R_T[MAX_LOG_N], R_U[MAX_LOG_N];
    fr_t x_final; int m;
} samemsm_proof_t;
/* ::new :54-150.  G, T, U, x are consumed. */
static void samemsm_prove(samemsm_proof_t *pf, g1a_t *G, const g1j_t *A, const g1j_t *Z_t, const g1j_t *Z_u, g1a_t *T, g1a_t *U,
                          fr_t *x, size_t n, transcript_t *tr, stdrng_t *rng, int threads) {
    fr_t *r = (fr_t *)malloc(sizeof(fr_t) * n);
    for (size_t i = 0; i < n; i++) fr_rand(&r[i], rng);                                        /* :78 */
    g1_msm(&pf->B_a, G, r, n, threads); g1_msm(&pf->B_t, T, r, n, threads); g1_msm(&pf->B_u, U, r, n, threads);
    t_append_g1j(tr, "same_msm_step1", A); t_append_g1j(tr, "same_msm_step1", Z_t); t_append_g1j(tr, "same_msm_step1", Z_u);
    t_append_g1a_vec(tr, "same_msm_step1", T, n); t_append_g1a_vec(tr, "same_msm_step1", U, n);
    t_append_g1j(tr, "same_msm_step1", &pf->B_a); t_append_g1j(tr, "same_msm_step1", &pf->B_t); t_append_g1j(tr, "same_msm_step1", &pf->B_u);
    fr_t alpha; t_challenge(tr, "same_msm_alpha", &alpha);
    for (size_t i = 0; i < n; i++) { fr_t t; fr_mul(&t, &alpha, &x[i]); fr_add(&x[i], &r[i], &t); }
    int round = 0;
    while (n > 1) {                                                                            /* :99-136 */
        n /= 2;
        g1_msm(&pf->L_A[round], G + n, x, n, threads); g1_msm(&pf->L_T[round], T + n, x, n, threads); g1_msm(&pf->L_U[round], U + n, x, n, threads);
        g1_msm(&pf->R_A[round], G, x + n, n, threads); g1_msm(&pf->R_T[round], T, x + n, n, threads); g1_msm(&pf->R_U[round], U, x + n, n, threads);
        t_append_g1j(tr, "same_msm_loop", &pf->L_A[round]); t_append_g1j(tr, "same_msm_loop", &pf->L_T[round]); t_append_g1j(tr, "same_msm_loop", &pf->L_U[round]);
        t_append_g1j(tr, "same_msm_loop", &pf->R_A[round]); t_append_g1j(tr, "same_msm_loop", &pf->R_T[round]); t_append_g1j(tr, "same_msm_loop", &pf->R_U[round]);
        fr_t gamma, gamma_inv; t_challenge(tr, "same_msm_gamma", &gamma); fr_inv(&gamma_inv, &gamma);
        for (size_t i = 0; i < n; i++) {                                                       /* :126-131 (the fold) */
            fr_t u; fr_mul(&u, &gamma_inv, &x[n + i]); fr_add(&x[i], &x[i], &u);
            g1j_t s;
            g1a_mul_fr(&s, &T[n + i], &gamma); g1j_add_affine(&s, &s, &T[i]); g1j_to_affine(&T[i], &s);
            g1a_mul_fr(&s, &U[n + i], &gamma); g1j_add_affine(&s, &s, &U[i]); g1j_to_affine(&U[i], &s);
            g1a_mul_fr(&s, &G[n + i], &gamma); g1j_add_affine(&s, &s, &G[i]); g1j_to_affine(&G[i], &s);
        }
        round++;
    }
    pf->m = round; pf->x_final = x[0];
    free(r);
}
/* ::verify :213-261 */
static int samemsm_verify(const samemsm_proof_t *pf, const g1a_t *G, const g1j_t *A, const g1j_t *Z_t, const g1j_t *Z_u,
                          const g1a_t *T, const g1a_t *U, size_t n, transcript_t *tr, msm_acc_t *acc, stdrng_t *rng, int threads) {
    t_append_g1j(tr, "same_msm_step1", A); t_append_g1j(tr, "same_msm_step1", Z_t); t_append_g1j(tr, "same_msm_step1", Z_u);
    t_append_g1a_vec(tr, "same_msm_step1", T, n); t_append_g1a_vec(tr, "same_msm_step1", U, n);
    t_append_g1j(tr, "same_msm_step1", &pf->B_a); t_append_g1j(tr, "same_msm_step1", &pf->B_t); t_append_g1j(tr, "same_msm_step1", &pf->B_u);
    fr_t alpha; t_challenge(tr, "same_msm_alpha", &alpha);
    int m = pf->m;
    if (m >= 32 || n != ((size_t)1 << m)) return -1;
    fr_t gam[MAX_LOG_N], gam_inv[MAX_LOG_N], scratch[MAX_LOG_N];
    for (int i = 0; i < m; i++) {
        t_append_g1j(tr, "same_msm_loop", &pf->L_A[i]); t_append_g1j(tr, "same_msm_loop", &pf->L_T[i]); t_append_g1j(tr, "same_msm_loop", &pf->L_U[i]);
        t_append_g1j(tr, "same_msm_loop", &pf->R_A[i]); t_append_g1j(tr, "same_msm_loop", &pf->R_T[i]); t_append_g1j(tr, "same_msm_loop", &pf->R_U[i]);
        t_challenge(tr, "same_msm_gamma", &gam[i]); gam_inv[i] = gam[i];
    }
    fr_batch_inv(gam_inv, m, scratch);
    fr_t *xs = (fr_t *)malloc(sizeof(fr_t) * n);
    verification_scalars_s(xs, gam, m, n);
    for (size_t i = 0; i < n; i++) fr_mul(&xs[i], &pf->x_final, &xs[i]);
    const g1j_t *Bs[3] = {&pf->B_a, &pf->B_t, &pf->B_u}, *Cs[3] = {A, Z_t, Z_u};
    const g1j_t *Ls[3] = {pf->L_A, pf->L_T, pf->L_U}, *Rs[3] = {pf->R_A, pf->R_T, pf->R_U};
    const g1a_t *Vs[3] = {G, T, U};
    for (int k = 0; k < 3; k++) {
        g1j_t t, base, lhs; g1j_mul_fr(&t, Cs[k], &alpha); g1j_add(&base, Bs[k], &t);
        g1_msm_from_projective(&lhs, Ls[k], gam, m, threads); g1j_add(&lhs, &lhs, &base);
        g1_msm_from_projective(&t, Rs[k], gam_inv, m, threads); g1j_add(&lhs, &lhs, &t);
        acc_accumulate_check(acc, &lhs, xs, Vs[k], n, rng);
    }
    free(xs);
    return 0;
}

/* ----------------------------------------------------------------- CurdleproofsProof: src/curdleproofs.rs */
typedef struct {
    g1j_t A; gcomm_t cm_T, cm_U; g1j_t R, S;
    sameperm_proof_t same_perm; samescalar_proof_t same_scalar; samemsm_proof_t same_msm;
} curdle_proof_t;

static void build_blinded_vectors(const crs_t *crs, const g1a_t *vec_T, const g1a_t *vec_U, g1a_t *Gb, g1a_t *Tb, g1a_t *Ub) {
    size_t ell = crs->ell; g1a_t Ha; g1j_to_affine(&Ha, &crs->H);             /* :136-155 / :259-278 */
    memcpy(Gb, crs->vec_G, sizeof(g1a_t) * ell); Gb[ell] = crs->vec_H[0]; Gb[ell + 1] = crs->vec_H[1];
    g1j_to_affine(&Gb[ell + 2], &crs->G_t); g1j_to_affine(&Gb[ell + 3], &crs->G_u);
    memcpy(Tb, vec_T, sizeof(g1a_t) * ell); g1a_set_inf(&Tb[ell]); g1a_set_inf(&Tb[ell + 1]); Tb[ell + 2] = Ha; g1a_set_inf(&Tb[ell + 3]);
    memcpy(Ub, vec_U, sizeof(g1a_t) * ell); g1a_set_inf(&Ub[ell]); g1a_set_inf(&Ub[ell + 1]); g1a_set_inf(&Ub[ell + 2]); Ub[ell + 3] = Ha;
}
/* ::new :59-184 */
static void curdle_prove(curdle_proof_t *pf, const crs_t *crs, const g1a_t *vec_R, const g1a_t *vec_S, const g1a_t *vec_T, const g1a_t *vec_U,
                         const g1j_t *M, const uint32_t *perm, const fr_t *k, const fr_t *vec_m_blinders, stdrng_t *rng, int threads) {
    const size_t ell = crs->ell, n = ell + N_BLINDERS;
    transcript_t tr; transcript_init(&tr, "curdleproofs");
    t_append_g1a_vec(&tr, "curdleproofs_step1", vec_R, ell); t_append_g1a_vec(&tr, "curdleproofs_step1", vec_S, ell);
    t_append_g1a_vec(&tr, "curdleproofs_step1", vec_T, ell); t_append_g1a_vec(&tr, "curdleproofs_step1", vec_U, ell);
    t_append_g1j(&tr, "curdleproofs_step1", M);
    fr_t *vec_a = (fr_t *)malloc(sizeof(fr_t) * ell), *a_perm = (fr_t *)malloc(sizeof(fr_t) * n);
    for (size_t i = 0; i < ell; i++) t_challenge(&tr, "curdleproofs_vec_a", &vec_a[i]);
    fr_t r_a_prime[N_BLINDERS]; fr_rand(&r_a_prime[0], rng); fr_rand(&r_a_prime[1], rng); fr_zero(&r_a_prime[2]); fr_zero(&r_a_prime[3]); /* :86-89 */
    for (size_t i = 0; i < ell; i++) a_perm[i] = vec_a[perm[i]];
    g1j_t t; g1_msm(&pf->A, crs->vec_G, a_perm, ell, threads); g1_msm(&t, crs->vec_H, r_a_prime, N_BLINDERS, threads); g1j_add(&pf->A, &pf->A, &t); /* :93 */
    sameperm_prove(&pf->same_perm, crs->vec_G, ell, crs->vec_H, &crs->H, &pf->A, M, vec_a, perm, r_a_prime, vec_m_blinders, &tr, rng, threads);
    fr_t r_t, r_u; fr_rand(&r_t, rng); fr_rand(&r_u, rng);                                   /* :110-111 */
    g1_msm(&pf->R, vec_R, vec_a, ell, threads); g1_msm(&pf->S, vec_S, vec_a, ell, threads);
    g1j_mul_fr(&t, &pf->R, k); gcomm_new(&pf->cm_T, &crs->G_t, &crs->H, &t, &r_t);
    g1j_mul_fr(&t, &pf->S, k); gcomm_new(&pf->cm_U, &crs->G_u, &crs->H, &t, &r_u);
    samescalar_prove(&pf->same_scalar, &crs->G_t, &crs->G_u, &crs->H, &pf->R, &pf->S, &pf->cm_T, &pf->cm_U, k, &r_t, &r_u, &tr, rng);
    g1j_t A_prime; g1j_add(&A_prime, &pf->A, &pf->cm_T.T_1); g1j_add(&A_prime, &A_prime, &pf->cm_U.T_1);
    g1a_t *Gb = (g1a_t *)malloc(sizeof(g1a_t) * n), *Tb = (g1a_t *)malloc(sizeof(g1a_t) * n), *Ub = (g1a_t *)malloc(sizeof(g1a_t) * n);
    build_blinded_vectors(crs, vec_T, vec_U, Gb, Tb, Ub);
    a_perm[ell] = r_a_prime[0]; a_perm[ell + 1] = r_a_prime[1]; a_perm[ell + 2] = r_t; a_perm[ell + 3] = r_u; /* :157-160 */
    samemsm_prove(&pf->same_msm, Gb, &A_prime, &pf->cm_T.T_2, &pf->cm_U.T_2, Tb, Ub, a_perm, n, &tr, rng, threads);
    free(vec_a); free(a_perm); free(Gb); free(Tb); free(Ub);
}
/* ::verify :197-298; returns 0 on Ok.  When out_* are non-NULL the final accumulated MSM instance is exported. */
static int curdle_verify_ex(const curdle_proof_t *pf, const crs_t *crs, const g1a_t *vec_R, const g1a_t *vec_S, const g1a_t *vec_T, const g1a_t *vec_U,
                            const g1j_t *M, stdrng_t *rng, int threads, g1a_t **out_bases, fr_t **out_scalars, size_t *out_n) {
    const size_t ell = crs->ell, n = ell + N_BLINDERS;
    transcript_t tr; transcript_init(&tr, "curdleproofs");
    if (g1a_is_inf(&vec_T[0])) return -1;                                                    /* :218-220 */
    msm_acc_t acc; acc_init(&acc, 5 * ell + 16);
    t_append_g1a_vec(&tr, "curdleproofs_step1", vec_R, ell); t_append_g1a_vec(&tr, "curdleproofs_step1", vec_S, ell);
    t_append_g1a_vec(&tr, "curdleproofs_step1", vec_T, ell); t_append_g1a_vec(&tr, "curdleproofs_step1", vec_U, ell);
    t_append_g1j(&tr, "curdleproofs_step1", M);
    fr_t *vec_a = (fr_t *)malloc(sizeof(fr_t) * ell);
    for (size_t i = 0; i < ell; i++) t_challenge(&tr, "curdleproofs_vec_a", &vec_a[i]);
    int rc = sameperm_verify(&pf->same_perm, crs->vec_G, ell, crs->vec_H, &crs->H, &crs->G_sum, &crs->H_sum, &pf->A, M, vec_a, &tr, &acc, rng, threads);
    if (rc == 0) rc = samescalar_verify(&pf->same_scalar, &crs->G_t, &crs->G_u, &crs->H, &pf->R, &pf->S, &pf->cm_T, &pf->cm_U, &tr);
    if (rc == 0) {
        g1j_t A_prime; g1j_add(&A_prime, &pf->A, &pf->cm_T.T_1); g1j_add(&A_prime, &A_prime, &pf->cm_U.T_1);
        g1a_t *Gb = (g1a_t *)malloc(sizeof(g1a_t) * n), *Tb = (g1a_t *)malloc(sizeof(g1a_t) * n), *Ub = (g1a_t *)malloc(sizeof(g1a_t) * n);
        build_blinded_vectors(crs, vec_T, vec_U, Gb, Tb, Ub);
        rc = samemsm_verify(&pf->same_msm, Gb, &A_prime, &pf->cm_T.T_2, &pf->cm_U.T_2, Tb, Ub, n, &tr, &acc, rng, threads);
        free(Gb); free(Tb); free(Ub);
    }
    if (rc == 0) {
        acc_accumulate_check(&acc, &pf->R, vec_a, vec_R, ell, rng);                          /* :293-294 */
        acc_accumulate_check(&acc, &pf->S, vec_a, vec_S, ell, rng);
        rc = acc_verify(&acc, threads, out_bases, out_scalars, out_n) ? 0 : -1;
    }
    acc_free(&acc); free(vec_a);
    return rc;
}

/* ----------------------------------------------------------------- serialisation: src/curdleproofs.rs:300-323 and the per-argument (de)serialise fns */
static size_t curdle_proof_size(int m) { return 48 * (1 + 4 + 2 + 1 + 1 + 2 + 4 * m + 4 + 3 + 6 * m) + 32 * (1 + 2 + 3 + 1); }
static uint8_t *put_g1(uint8_t *w, const g1j_t *p) { g1j_compress(w, p); return w + 48; }
static uint8_t *put_fr(uint8_t *w, const fr_t *a) { fr_to_bytes(w, a); return w + 32; }
static size_t curdle_serialize(uint8_t *out, const curdle_proof_t *pf) {
    uint8_t *w = out;
    w = put_g1(w, &pf->A); w = put_g1(w, &pf->cm_T.T_1); w = put_g1(w, &pf->cm_T.T_2); w = put_g1(w, &pf->cm_U.T_1); w = put_g1(w, &pf->cm_U.T_2);
    w = put_g1(w, &pf->R); w = put_g1(w, &pf->S);
    const gprod_proof_t *gp = &pf->same_perm.gprod; const ipa_proof_t *ip = &gp->ipa;
    w = put_g1(w, &pf->same_perm.B); w = put_g1(w, &gp->C); w = put_fr(w, &gp->r_p);
    w = put_g1(w, &ip->B_c); w = put_g1(w, &ip->B_d);
    for (int i = 0; i < ip->m; i++) w = put_g1(w, &ip->L_C[i]);
    for (int i = 0; i < ip->m; i++) w = put_g1(w, &ip->R_C[i]);
    for (int i = 0; i < ip->m; i++) w = put_g1(w, &ip->L_D[i]);
    for (int i = 0; i < ip->m; i++) w = put_g1(w, &ip->R_D[i]);
    w = put_fr(w, &ip->c_final); w = put_fr(w, &ip->d_final);
    const samescalar_proof_t *ss = &pf->same_scalar;
    w = put_g1(w, &ss->cm_A.T_1); w = put_g1(w, &ss->cm_A.T_2); w = put_g1(w, &ss->cm_B.T_1); w = put_g1(w, &ss->cm_B.T_2);
    w = put_fr(w, &ss->z_k); w = put_fr(w, &ss->z_t); w = put_fr(w, &ss->z_u);
    const samemsm_proof_t *sm = &pf->same_msm;
    w = put_g1(w, &sm->B_a); w = put_g1(w, &sm->B_t); w = put_g1(w, &sm->B_u);
    const g1j_t *vs[6] = {sm->L_A, sm->L_T, sm->L_U, sm->R_A, sm->R_T, sm->R_U};
    for (int v = 0; v < 6; v++) for (int i = 0; i < sm->m; i++) w = put_g1(w, &vs[v][i]);
    w = put_fr(w, &sm->x_final);
    return (size_t)(w - out);
}
static int get_g1(const uint8_t **r, g1j_t *p) { g1a_t a; if (g1a_decompress(&a, *r, 1)) return -1; g1j_from_affine(p, &a); *r += 48; return 0; }
static int get_fr(const uint8_t **r, fr_t *a) { uint64_t c[4]; memcpy(c, *r, 32); if (limbs_geq(c, FR_R, 4)) return -1; fr_from_canon(a, c); *r += 32; return 0; }
static int curdle_deserialize(curdle_proof_t *pf, const uint8_t *in, int m) {
    const uint8_t *r = in; int e = 0;
    e |= get_g1(&r, &pf->A); e |= get_g1(&r, &pf->cm_T.T_1); e |= get_g1(&r, &pf->cm_T.T_2); e |= get_g1(&r, &pf->cm_U.T_1); e |= get_g1(&r, &pf->cm_U.T_2);
    e |= get_g1(&r, &pf->R); e |= get_g1(&r, &pf->S);
    gprod_proof_t *gp = &pf->same_perm.gprod; ipa_proof_t *ip = &gp->ipa;
    e |= get_g1(&r, &pf->same_perm.B); e |= get_g1(&r, &gp->C); e |= get_fr(&r, &gp->r_p);
    e |= get_g1(&r, &ip->B_c); e |= get_g1(&r, &ip->B_d); ip->m = m;
    for (int i = 0; i < m && !e; i++) e |= get_g1(&r, &ip->L_C[i]);
    for (int i = 0; i < m && !e; i++) e |= get_g1(&r, &ip->R_C[i]);
    for (int i = 0; i < m && !e; i++) e |= get_g1(&r, &ip->L_D[i]);
    for (int i = 0; i < m && !e; i++) e |= get_g1(&r, &ip->R_D[i]);
    if (e) return -1;
    e |= get_fr(&r, &ip->c_final); e |= get_fr(&r, &ip->d_final);
    samescalar_proof_t *ss = &pf->same_scalar;
    e |= get_g1(&r, &ss->cm_A.T_1); e |= get_g1(&r, &ss->cm_A.T_2); e |= get_g1(&r, &ss->cm_B.T_1); e |= get_g1(&r, &ss->cm_B.T_2);
    e |= get_fr(&r, &ss->z_k); e |= get_fr(&r, &ss->z_t); e |= get_fr(&r, &ss->z_u);
    samemsm_proof_t *sm = &pf->same_msm; sm->m = m;
    e |= get_g1(&r, &sm->B_a); e |= get_g1(&r, &sm->B_t); e |= get_g1(&r, &sm->B_u);
    g1j_t *vs[6] = {sm->L_A, sm->L_T, sm->L_U, sm->R_A, sm->R_T, sm->R_U};
    for (int v = 0; v < 6 && !e; v++) for (int i = 0; i < m && !e; i++) e |= get_g1(&r, &vs[v][i]);
    if (e) return -1;
    e |= get_fr(&r, &sm->x_final);
    return e ? -1 : 0;
}

/* shuffle_permute_and_commit_input: src/util.rs:83-106 */
static void shuffle_permute_and_commit(const crs_t *crs, const g1a_t *vec_R, const g1a_t *vec_S, const uint32_t *perm, const fr_t *k,
                                       stdrng_t *rng, g1a_t *vec_T, g1a_t *vec_U, g1j_t *M, fr_t *vec_m_blinders, int threads) {
    size_t ell = crs->ell;
    g1a_t *tT = (g1a_t *)malloc(sizeof(g1a_t) * ell), *tU = (g1a_t *)malloc(sizeof(g1a_t) * ell);
    for (size_t i = 0; i < ell; i++) { g1j_t s; g1a_mul_fr(&s, &vec_R[i], k); g1j_to_affine(&tT[i], &s); g1a_mul_fr(&s, &vec_S[i], k); g1j_to_affine(&tU[i], &s); }
    for (size_t i = 0; i < ell; i++) { vec_T[i] = tT[perm[i]]; vec_U[i] = tU[perm[i]]; }
    fr_t *sigma = (fr_t *)malloc(sizeof(fr_t) * ell);
    for (size_t i = 0; i < ell; i++) fr_from_u64(&sigma[i], perm[i]);
    for (int i = 0; i < N_BLINDERS; i++) fr_rand(&vec_m_blinders[i], rng);
    g1j_t t; g1_msm(M, crs->vec_G, sigma, ell, threads); g1_msm(&t, crs->vec_H, vec_m_blinders, N_BLINDERS, threads); g1j_add(M, M, &t);
    free(tT); free(tU); free(sigma);
}
#endif
