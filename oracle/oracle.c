/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Exported (ctypes) surface of the oracle: the MSM / fold / normalise boundary of the reference
 * (`/root/reference/src/util.rs:19-29` and the inline fold loops) in the SAME byte layouts as
 * include/cdp_msm.h, the whisk golden-vector drivers (`/root/reference/src/whisk.rs:381-456`), and a
 * generic prove/verify pair used by the parity tests and by bench.py's cpu_baseline / reference leg.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py may load this library.
 */
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "protocol.h"

#define EXPORT __attribute__((visibility("default")))

/* ---- layout converters (include/cdp_msm.h layouts: Fp Montgomery 6xu64 LE; scalars canonical 4xu64 LE) ---- */
static void load_affine(g1a_t *p, const uint8_t *b) { memcpy(p, b, 96); }
static void store_affine(uint8_t *b, const g1a_t *p) { memcpy(b, p, 96); }
static void load_jac(g1j_t *p, const uint8_t *b) { memcpy(p, b, 144); }
static void store_jac(uint8_t *b, const g1j_t *p) { memcpy(b, p, 144); }

EXPORT int oracle_msm(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t out_jac[144], int threads) {
    g1j_t r; g1_msm_canon(&r, (const g1a_t *)pts, (const uint64_t (*)[4])scalars, n, threads);
    store_jac(out_jac, &r); return 0;
}
/* naive sum of double-and-adds: an MSM that shares nothing with Pippenger */
EXPORT int oracle_msm_naive(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t out_jac[144]) {
    g1j_t acc; g1j_set_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        g1a_t p; load_affine(&p, pts + 96 * i); g1j_t j, t; g1j_from_affine(&j, &p);
        g1j_mul_limbs(&t, &j, (const uint64_t *)(scalars + 32 * i), 4); g1j_add(&acc, &acc, &t);
    }
    store_jac(out_jac, &acc); return 0;
}
EXPORT int oracle_msm_from_projective(const uint8_t *jac, const uint8_t *scalars, size_t n, uint8_t out_jac[144]) {
    g1a_t *aff = (g1a_t *)malloc(sizeof(g1a_t) * (n ? n : 1));
    g1j_batch_to_affine(aff, (const g1j_t *)jac, n);
    g1j_t r; g1_msm_canon(&r, aff, (const uint64_t (*)[4])scalars, n, 1);
    store_jac(out_jac, &r); free(aff); return 0;
}
/* out[i] = (s_i * P_i).into_affine()   -- src/grand_product_argument.rs:92-102, src/util.rs:94-95 */
EXPORT int oracle_scalar_mul_batch(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out_affine) {
    for (size_t i = 0; i < n; i++) {
        g1a_t p, o; load_affine(&p, pts + 96 * i); g1j_t j, t; g1j_from_affine(&j, &p);
        g1j_mul_limbs(&t, &j, (const uint64_t *)(scalars + 32 * i), 4); g1j_to_affine(&o, &t); store_affine(out_affine + 96 * i, &o);
    }
    return 0;
}
/* out[i] = (L_i + gamma * R_i).into_affine()  -- src/inner_product_argument.rs:177-178, src/same_multiscalar_argument.rs:128-130 */
EXPORT int oracle_fold(const uint8_t *L, const uint8_t *R, const uint8_t gamma[32], size_t n, uint8_t *out_affine) {
    for (size_t i = 0; i < n; i++) {
        g1a_t l, r, o; load_affine(&l, L + 96 * i); load_affine(&r, R + 96 * i);
        g1j_t j, t; g1j_from_affine(&j, &r); g1j_mul_limbs(&t, &j, (const uint64_t *)gamma, 4);
        g1j_add_affine(&t, &t, &l); g1j_to_affine(&o, &t); store_affine(out_affine + 96 * i, &o);
    }
    return 0;
}
EXPORT int oracle_normalize_batch(const uint8_t *jac, size_t n, uint8_t *out_affine) {
    g1j_batch_to_affine((g1a_t *)out_affine, (const g1j_t *)jac, n); return 0;
}
EXPORT int oracle_compress(const uint8_t *affine, size_t n, uint8_t *out48) {
    for (size_t i = 0; i < n; i++) g1a_compress(out48 + 48 * i, (const g1a_t *)(affine + 96 * i));
    return 0;
}
EXPORT int oracle_decompress(const uint8_t *in48, size_t n, uint8_t *out_affine, int check_subgroup) {
    for (size_t i = 0; i < n; i++) if (g1a_decompress((g1a_t *)(out_affine + 96 * i), in48 + 48 * i, check_subgroup)) return -1;
    return 0;
}
EXPORT int oracle_on_curve(const uint8_t *affine, size_t n) {
    for (size_t i = 0; i < n; i++) if (!g1a_on_curve((const g1a_t *)(affine + 96 * i))) return 0;
    return 1;
}
EXPORT void oracle_generator(uint8_t out_affine[96]) { g1a_t g; g1a_generator(&g); store_affine(out_affine, &g); }
EXPORT void oracle_fp_from_canon(const uint8_t in[48], uint8_t out[48]) { fp_t t; fp_from_canon(&t, (const uint64_t *)in); memcpy(out, &t, 48); }
EXPORT void oracle_fp_to_canon(const uint8_t in[48], uint8_t out[48]) { fp_t t; memcpy(&t, in, 48); fp_to_canon((uint64_t *)out, &t); }
EXPORT void oracle_fp_mul(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) { fp_t x, y, z; memcpy(&x, a, 48); memcpy(&y, b, 48); fp_mul(&z, &x, &y); memcpy(out, &z, 48); }

/* ---- transcript KAT hook: new(label); append_message(l1, m1); challenge_bytes(l2, out) ---- */
EXPORT void oracle_merlin_kat(const char *proto, const char *l1, const uint8_t *m1, size_t n1, const char *l2, uint8_t *out, size_t nout) {
    transcript_t t; transcript_init(&t, proto); transcript_append_message(&t, l1, m1, n1); transcript_challenge_bytes(&t, l2, out, nout);
}
EXPORT void oracle_keccak_f1600(uint64_t st[25]) { keccak_f1600(st); }
/* ---- RNG KAT hooks ---- */
EXPORT void oracle_stdrng_u32s(uint64_t seed, uint32_t *out, size_t n) { stdrng_t r; stdrng_seed_from_u64(&r, seed); for (size_t i = 0; i < n; i++) out[i] = stdrng_next_u32(&r); }
EXPORT void oracle_stdrng_seed_bytes(uint64_t seed, uint8_t out[32]) { stdrng_t r; stdrng_seed_from_u64(&r, seed); memcpy(out, r.key, 32); }

/* ---- whisk golden vectors ---- */
/* whisk_tracker_proof test, src/whisk.rs:381-402 (uses :50-67, :231-263) */
/* inputs of the same test for a caller that runs the product API: k (32) | tracker r_G, k_r_G (96) | k_commitment (48) | u64 rng words consumed before generate_whisk_tracker_proof */
EXPORT int oracle_whisk_tracker_inputs_seed0(uint8_t out[184]) {
    stdrng_t rng; stdrng_seed_from_u64(&rng, 0);
    fr_t k, r; fr_rand(&k, &rng); fr_rand(&r, &rng);
    g1a_t G; g1a_generator(&G);
    g1j_t r_G, k_r_G, k_G; g1a_mul_fr(&r_G, &G, &r);
    g1a_t r_Ga; g1j_to_affine(&r_Ga, &r_G); g1a_mul_fr(&k_r_G, &r_Ga, &k); g1a_mul_fr(&k_G, &G, &k);
    fr_to_bytes(out, &k); g1j_compress(out + 32, &r_G); g1j_compress(out + 80, &k_r_G); g1j_compress(out + 128, &k_G);
    uint64_t words = (rng.counter / 4 - 1) * 64 + (uint64_t)rng.index;
    memcpy(out + 176, &words, 8);
    return 0;
}
EXPORT int oracle_whisk_tracker_proof_seed0(uint8_t out[128]) {
    stdrng_t rng; stdrng_seed_from_u64(&rng, 0);
    fr_t k, r, blinder; fr_rand(&k, &rng); fr_rand(&r, &rng);
    g1a_t G; g1a_generator(&G);
    g1j_t r_G, k_r_G, k_G, A, B; g1a_mul_fr(&r_G, &G, &r);
    g1a_t r_Ga; g1j_to_affine(&r_Ga, &r_G); g1a_mul_fr(&k_r_G, &r_Ga, &k);
    g1a_mul_fr(&k_G, &G, &k);
    fr_rand(&blinder, &rng);
    g1a_mul_fr(&A, &G, &blinder); g1a_mul_fr(&B, &r_Ga, &blinder);
    transcript_t tr; transcript_init(&tr, "whisk_opening_proof");
    t_append_g1j(&tr, "tracker_opening_proof", &k_G); t_append_g1a(&tr, "tracker_opening_proof", &G);
    t_append_g1j(&tr, "tracker_opening_proof", &k_r_G); t_append_g1a(&tr, "tracker_opening_proof", &r_Ga);
    t_append_g1j(&tr, "tracker_opening_proof", &A); t_append_g1j(&tr, "tracker_opening_proof", &B);
    fr_t ch, s; t_challenge(&tr, "tracker_opening_proof_challenge", &ch);
    fr_mul(&s, &ch, &k); fr_sub(&s, &blinder, &s);
    g1j_compress(out, &A); g1j_compress(out + 48, &B); fr_to_bytes(out + 96, &s);
    return 0;
}
/* whisk_shuffle_proof test, src/whisk.rs:416-456 with N = 128 (generic in ell here so that the same
 * driver also produces the seed-0 instances for other sizes).  Outputs:
 *   proof_out : 48 (M) + curdle_proof_size(m) bytes              (the 4496-byte golden string at ell=124)
 *   inst_out  : optional, 4*ell affine points R,S,T,U (96 B each), M (144 B jacobian), permutation (ell x u32), k (32 B),
 *               vec_m_blinders (4 x 32 B) and, as a u64, the number of u32 words the rng had produced when
 *               CurdleproofsProof::new was entered (so that a prover can be handed the same stream), then the same count
 *               at the entry of generate_whisk_shuffle_proof
 *   verified  : result of is_valid_whisk_shuffle_proof on the serialised proof with the same rng */
EXPORT int oracle_whisk_shuffle_proof_seed0(size_t ell, uint8_t *proof_out, uint8_t *inst_out, int *verified, int threads) {
    stdrng_t rng; stdrng_seed_from_u64(&rng, 0);
    crs_t crs; crs_generate(&crs, ell);
    size_t n = ell + N_BLINDERS; int m = 0; while (((size_t)1 << m) < n) m++;
    g1a_t *vR = (g1a_t *)malloc(sizeof(g1a_t) * ell), *vS = (g1a_t *)malloc(sizeof(g1a_t) * ell);
    g1a_t *vT = (g1a_t *)malloc(sizeof(g1a_t) * ell), *vU = (g1a_t *)malloc(sizeof(g1a_t) * ell);
    g1a_t G; g1a_generator(&G);
    for (size_t i = 0; i < ell; i++) { /* WhiskTracker::from_rand :65-68 */
        fr_t k, r; fr_rand(&k, &rng); fr_rand(&r, &rng);
        g1j_t t; g1a_mul_fr(&t, &G, &r); g1j_to_affine(&vR[i], &t); g1a_mul_fr(&t, &vR[i], &k); g1j_to_affine(&vS[i], &t);
    }
    uint64_t words_at_entry = (rng.counter / 4 - 1) * 64 + (uint64_t)rng.index;   /* rng position when generate_whisk_shuffle_proof is entered */
    uint32_t *perm = (uint32_t *)malloc(4 * ell);
    for (size_t i = 0; i < ell; i++) perm[i] = (uint32_t)i;
    stdrng_shuffle_u32(perm, ell, &rng);                                   /* :153 */
    fr_t k; fr_rand(&k, &rng);                                             /* :154 */
    g1j_t M; fr_t m_bl[N_BLINDERS];
    shuffle_permute_and_commit(&crs, vR, vS, perm, &k, &rng, vT, vU, &M, m_bl, threads);
    uint64_t words_before = (rng.counter / 4 - 1) * 64 + (uint64_t)rng.index;
    curdle_proof_t pf; curdle_prove(&pf, &crs, vR, vS, vT, vU, &M, perm, &k, m_bl, &rng, threads);
    g1j_compress(proof_out, &M); size_t sz = curdle_serialize(proof_out + 48, &pf);
    if (inst_out) {
        memcpy(inst_out, vR, 96 * ell); memcpy(inst_out + 96 * ell, vS, 96 * ell);
        memcpy(inst_out + 192 * ell, vT, 96 * ell); memcpy(inst_out + 288 * ell, vU, 96 * ell);
        memcpy(inst_out + 384 * ell, &M, 144);
        uint8_t *w = inst_out + 384 * ell + 144;
        memcpy(w, perm, 4 * ell); w += 4 * ell;
        fr_to_bytes(w, &k); w += 32;
        for (int i = 0; i < N_BLINDERS; i++) { fr_to_bytes(w, &m_bl[i]); w += 32; }
        memcpy(w, &words_before, 8);
        memcpy(w + 8, &words_at_entry, 8);
    }
    if (verified) { /* is_valid_whisk_shuffle_proof :106-130 */
        curdle_proof_t pf2; g1j_t M2; const uint8_t *rd = proof_out;
        int rc = get_g1(&rd, &M2); if (!rc) rc = curdle_deserialize(&pf2, proof_out + 48, m);
        if (!rc) rc = curdle_verify_ex(&pf2, &crs, vR, vS, vT, vU, &M2, &rng, threads, NULL, NULL, NULL);
        *verified = (rc == 0);
    }
    free(vR); free(vS); free(vT); free(vU); free(perm); crs_free(&crs);
    return (int)sz + 48;
}

/* ---- generic prove / verify on explicit inputs ---- */
/* crs_pts: ell+7 affine points (vec_G | vec_H | H | G_t | G_u), as CurdleproofsCrs::from_points (src/crs.rs:37-58) */
EXPORT int oracle_generate_crs_points(size_t ell, uint8_t *out_pts /* (ell+7)*96 */) {
    crs_t crs; crs_generate(&crs, ell);
    memcpy(out_pts, crs.vec_G, 96 * ell); memcpy(out_pts + 96 * ell, crs.vec_H, 96 * N_BLINDERS);
    g1a_t t; g1j_to_affine(&t, &crs.H); memcpy(out_pts + 96 * (ell + 4), &t, 96);
    g1j_to_affine(&t, &crs.G_t); memcpy(out_pts + 96 * (ell + 5), &t, 96);
    g1j_to_affine(&t, &crs.G_u); memcpy(out_pts + 96 * (ell + 6), &t, 96);
    crs_free(&crs); return 0;
}
/* Random instance from a seeded StdRng, in the draw order of the reference's own tests
 * (src/curdleproofs.rs:336-361): permutation.shuffle, k, vec_R, vec_S, then shuffle_permute_and_commit.
 * `fast_points` != 0 replaces G1Projective::rand by (Fr::rand)*G -- same distribution on the subgroup, ~8x cheaper;
 * it changes the rng stream and is therefore never used for the golden vectors. */
EXPORT int oracle_random_instance(size_t ell, const uint8_t *crs_pts, uint64_t seed, int fast_points, uint8_t *vec_R, uint8_t *vec_S, uint8_t *vec_T,
                                  uint8_t *vec_U, uint8_t M_jac[144], uint32_t *perm, uint8_t k_out[32], uint8_t *m_blinders /* 4*32 */, int threads) {
    stdrng_t rng; stdrng_seed_from_u64(&rng, seed);
    crs_t crs; crs_from_points(&crs, ell, (const g1a_t *)crs_pts);
    for (size_t i = 0; i < ell; i++) perm[i] = (uint32_t)i;
    stdrng_shuffle_u32(perm, ell, &rng);
    fr_t k; fr_rand(&k, &rng);
    g1a_t G; g1a_generator(&G);
    g1j_t *tmp = (g1j_t *)malloc(sizeof(g1j_t) * 2 * ell);
    for (size_t i = 0; i < 2 * ell; i++) {
        if (fast_points) { fr_t s; fr_rand(&s, &rng); g1a_mul_fr(&tmp[i], &G, &s); } else g1j_rand(&tmp[i], &rng);
    }
    g1j_batch_to_affine((g1a_t *)vec_R, tmp, ell); g1j_batch_to_affine((g1a_t *)vec_S, tmp + ell, ell);
    g1j_t M; fr_t mb[N_BLINDERS];
    shuffle_permute_and_commit(&crs, (const g1a_t *)vec_R, (const g1a_t *)vec_S, perm, &k, &rng, (g1a_t *)vec_T, (g1a_t *)vec_U, &M, mb, threads);
    store_jac(M_jac, &M); fr_to_bytes(k_out, &k);
    for (int i = 0; i < N_BLINDERS; i++) fr_to_bytes(m_blinders + 32 * i, &mb[i]);
    free(tmp); crs_free(&crs); return 0;
}
static void fr_load_canon(fr_t *r, const uint8_t *b) { uint64_t c[4]; memcpy(c, b, 32); fr_from_canon(r, c); }
/* CurdleproofsProof::new (src/curdleproofs.rs:59-184) + serialize (:300-310); prover randomness = StdRng::seed_from_u64(rng_seed) */
EXPORT int oracle_prove(size_t ell, const uint8_t *crs_pts, const uint8_t *vec_R, const uint8_t *vec_S, const uint8_t *vec_T, const uint8_t *vec_U,
                        const uint8_t M_jac[144], const uint32_t *perm, const uint8_t k_bytes[32], const uint8_t *m_blinders, uint64_t rng_seed,
                        uint8_t *proof_out, int threads) {
    crs_t crs; crs_from_points(&crs, ell, (const g1a_t *)crs_pts);
    stdrng_t rng; stdrng_seed_from_u64(&rng, rng_seed);
    g1j_t M; load_jac(&M, M_jac); fr_t k, mb[N_BLINDERS]; fr_load_canon(&k, k_bytes);
    for (int i = 0; i < N_BLINDERS; i++) fr_load_canon(&mb[i], m_blinders + 32 * i);
    curdle_proof_t pf;
    curdle_prove(&pf, &crs, (const g1a_t *)vec_R, (const g1a_t *)vec_S, (const g1a_t *)vec_T, (const g1a_t *)vec_U, &M, perm, &k, mb, &rng, threads);
    size_t sz = curdle_serialize(proof_out, &pf);
    crs_free(&crs); return (int)sz;
}
/* CurdleproofsProof::deserialize + verify (src/curdleproofs.rs:197-298, :312-323).  Returns 1 = Ok, 0 = VerificationError,
 * -1 = deserialisation error.  If acc_bases/acc_scalars are non-NULL the final accumulated MSM (msm_accumulator.rs:55-68)
 * is exported: bases 96 B affine, scalars 32 B canonical, count in *acc_n (capacity must be >= 5*ell+16). */
EXPORT int oracle_verify(size_t ell, const uint8_t *crs_pts, const uint8_t *vec_R, const uint8_t *vec_S, const uint8_t *vec_T, const uint8_t *vec_U,
                         const uint8_t M_jac[144], const uint8_t *proof, uint64_t rng_seed, int threads,
                         uint8_t *acc_bases, uint8_t *acc_scalars, size_t *acc_n) {
    size_t n = ell + N_BLINDERS; int m = 0; while (((size_t)1 << m) < n) m++;
    crs_t crs; crs_from_points(&crs, ell, (const g1a_t *)crs_pts);
    stdrng_t rng; stdrng_seed_from_u64(&rng, rng_seed);
    g1j_t M; load_jac(&M, M_jac);
    curdle_proof_t pf;
    if (curdle_deserialize(&pf, proof, m)) { crs_free(&crs); return -1; }
    g1a_t *ab = NULL; fr_t *as = NULL; size_t an = 0;
    int rc = curdle_verify_ex(&pf, &crs, (const g1a_t *)vec_R, (const g1a_t *)vec_S, (const g1a_t *)vec_T, (const g1a_t *)vec_U, &M, &rng, threads,
                              acc_bases ? &ab : NULL, acc_bases ? &as : NULL, acc_bases ? &an : NULL);
    if (acc_bases && ab) {
        memcpy(acc_bases, ab, 96 * an);
        for (size_t i = 0; i < an; i++) fr_to_bytes(acc_scalars + 32 * i, &as[i]);
        *acc_n = an; free(ab); free(as);
    }
    crs_free(&crs);
    return rc == 0 ? 1 : 0;
}
EXPORT size_t oracle_proof_size(size_t ell) { size_t n = ell + N_BLINDERS; int m = 0; while (((size_t)1 << m) < n) m++; return curdle_proof_size(m); }

/* ---- CPU baseline timers (bench.py): `count` independent proofs, one per OpenMP thread, each proof single-threaded
 * except that arkworks' per-window MSM parallelism is modelled by `msm_threads` when count == 1 ---- */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
EXPORT double oracle_time_prove(size_t ell, const uint8_t *crs_pts, const uint8_t *vec_R, const uint8_t *vec_S, const uint8_t *vec_T, const uint8_t *vec_U,
                                const uint8_t M_jac[144], const uint32_t *perm, const uint8_t k_bytes[32], const uint8_t *m_blinders,
                                int count, int threads, uint8_t *last_proof) {
    size_t psz = oracle_proof_size(ell);
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int i = 0; i < count; i++) {
        uint8_t *buf = (uint8_t *)malloc(psz);
        oracle_prove(ell, crs_pts, vec_R, vec_S, vec_T, vec_U, M_jac, perm, k_bytes, m_blinders, (uint64_t)i, buf, 1);
        if (i == count - 1 && last_proof) memcpy(last_proof, buf, psz);
        free(buf);
    }
    return now_s() - t0;
}
EXPORT double oracle_time_verify(size_t ell, const uint8_t *crs_pts, const uint8_t *vec_R, const uint8_t *vec_S, const uint8_t *vec_T, const uint8_t *vec_U,
                                 const uint8_t M_jac[144], const uint8_t *proof, int count, int threads, int *all_ok) {
    int ok = 1;
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(&& : ok)
    for (int i = 0; i < count; i++) ok = ok && (oracle_verify(ell, crs_pts, vec_R, vec_S, vec_T, vec_U, M_jac, proof, (uint64_t)i, 1, NULL, NULL, NULL) == 1);
    if (all_ok) *all_ok = ok;
    return now_s() - t0;
}
EXPORT double oracle_time_msm(const uint8_t *pts, const uint8_t *scalars, size_t n, int reps, int threads, uint8_t out_jac[144]) {
    double t0 = now_s();
    for (int i = 0; i < reps; i++) oracle_msm(pts, scalars, n, out_jac, threads);
    return now_s() - t0;
}
/* single-thread unit costs of this port (bench.py prints them next to the baseline so that its speed can be judged) */
EXPORT double oracle_time_fp_mul_ns(int reps) {
    fp_t a, b; memcpy(a.l, FP_GX_MONT, 48); memcpy(b.l, FP_GY_MONT, 48);
    double t0 = now_s();
    for (int i = 0; i < reps; i++) fp_mul(&a, &a, &b);   /* dependent chain */
    double dt = now_s() - t0;
    volatile uint64_t sink = a.l[0]; (void)sink;
    return dt / reps * 1e9;
}
EXPORT double oracle_time_mixed_add_ns(int reps) {
    g1a_t G; g1a_generator(&G); g1j_t J; g1j_from_affine(&J, &G); g1j_dbl(&J, &J);
    double t0 = now_s();
    for (int i = 0; i < reps; i++) g1j_add_affine(&J, &J, &G);
    double dt = now_s() - t0;
    volatile uint64_t sink = J.X.l[0]; (void)sink;
    return dt / reps * 1e9;
}
EXPORT int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
