// The PROVER's Fiat-Shamir transcript and scalar-field algebra on the device -- SURVEY.md section 8(f) rank 1, prover half.
//
// `CurdleproofsProof::new` (/root/reference/src/curdleproofs.rs:59-184) alternates between group work (MSMs, folds: the kernels of
// k_fixed.cu / k_msm.cu / k_smul.cu) and, between two such launches, (a) appending the freshly computed points to the merlin transcript
// and drawing the next challenges (/root/reference/src/transcript.rs:28-61) and (b) O(n) arithmetic in Fr that turns the challenges
// into the next launch's scalars.  This file is (a) + (b) for every step of the protocol, so that a batch of proofs advances from the
// transcript opening to the serialised proof bytes as ONE stream of kernels with no host round trip:
//
//   stage            follows                                                     consumes            emits (scalars of)
//   PS_S1            curdleproofs.rs:86-116 (witness vectors, blinders)          vec_a, randomness   stage 1: A, R, S, B_a, B_t, B_u, T_1, U_1, A_1, B_1
//   PS_SAMEPERM      same_permutation_argument.rs:60-82                          stage 1 points      stage 2: B, T_2, A_2, U_2, B_2, A'
//   PS_GPROD1        grand_product_argument.rs:63-83                             stage 2 points      stage 3: C
//   PS_GPROD2        grand_product_argument.rs:85-147, inner_product_argument    stage 3 point       stage 4: D, B_c, B_d
//                    .rs:53-77 (generate_ipa_blinders), :124-127
//   PS_IPA0          inner_product_argument.rs:129-140                           stage 4 points      IPA round 0: L_C, L_D, R_C, R_D
//   PS_IPA_ROUND k   inner_product_argument.rs:150-186; after the last round     round k points      IPA round k + 1, or SameMSM round 0:
//                    same_scalar_argument.rs:64-75 and                                               L_A, L_T, L_U, R_A, R_T, R_U
//                    same_multiscalar_argument.rs:84-91
//   PS_SM_ROUND k    same_multiscalar_argument.rs:99-136                         round k points      fold scalar gamma, SameMSM round k + 1
//
// One CTA per proof.  Thread 0 runs the Keccak-f[1600] permutations (STROBE-128 state in shared memory); message bytes are XORed into the
// state by all threads in parallel; the vector algebra is one thread per vector index with block reductions / scans through a per-proof
// scratch area.  The round MSMs are written over the ORIGINAL CRS bases (host/prover.cpp header), so a round's scalars are
// prefix weight x folded vector entry; the prefix weights live in the per-proof work area and double each round.
// Every proof byte is written here too (`CurdleproofsProof::serialize`, curdleproofs.rs:300-310 and the per-argument serialisers).
//
// Integer / byte work only.  The same source compiles as plain C++ (CDP_PROVE_HOST_HARNESS: a CTA is emulated by running every
// "for each thread" loop sequentially, which is exact because no phase between two barriers has cross-thread dependences) -- that is how
// tests/host/prove_dev_check.cpp checks it on the CPU against the oracle, proof bytes included.
#ifndef CDP_PROVE_HOST_HARNESS
#include "launch.h"
#endif
#include "../../include/cdp_msm.h"
#include "fr256.cuh"

#ifdef CDP_PROVE_HOST_HARNESS
#define CTA_FOR(i, cnt) for (uint32_t i = 0; i < (uint32_t)(cnt); i++)
#define CTA_SYNC() ((void)0)
#define CTA_LEADER true
#define CTA_SHARED static thread_local
#define CTA_PROOF_INDEX (harness_proof_index)
#else
#define CTA_FOR(i, cnt) for (uint32_t i = threadIdx.x; i < (uint32_t)(cnt); i += blockDim.x)
#define CTA_SYNC() __syncthreads()
#define CTA_LEADER (threadIdx.x == 0)
#define CTA_SHARED __shared__
#define CTA_PROOF_INDEX (blockIdx.x)
#endif

// Keccak / STROBE / merlin with the state shared by the CTA: cstrobe.cuh (the transcript opening of k_transcript.cu uses it as well)
#include "cstrobe.cuh"
namespace cdp {
namespace prove {

using namespace vcoef;
using namespace cstr;
#ifdef CDP_PROVE_HOST_HARNESS
static thread_local uint32_t harness_proof_index = 0;  // lanes are host threads
#endif

__device__ const uint8_t L_SP1[] = "same_perm_step1";
__device__ const uint8_t L_SPA[] = "same_perm_alpha";
__device__ const uint8_t L_SPB[] = "same_perm_beta";
__device__ const uint8_t L_GP1[] = "gprod_step1";
__device__ const uint8_t L_GPA[] = "gprod_alpha";
__device__ const uint8_t L_GP2[] = "gprod_step2";
__device__ const uint8_t L_GPB[] = "gprod_beta";
__device__ const uint8_t L_IP1[] = "ipa_step1";
__device__ const uint8_t L_IPA[] = "ipa_alpha";
__device__ const uint8_t L_IPB[] = "ipa_beta";
__device__ const uint8_t L_IPL[] = "ipa_loop";
__device__ const uint8_t L_IPG[] = "ipa_gamma";
__device__ const uint8_t L_SEP[] = "sameexp_points";
__device__ const uint8_t L_SSA[] = "same_scalar_alpha";
__device__ const uint8_t L_SM1[] = "same_msm_step1";
__device__ const uint8_t L_SMA[] = "same_msm_alpha";
__device__ const uint8_t L_SML[] = "same_msm_loop";
__device__ const uint8_t L_SMG[] = "same_msm_gamma";

// ---------------------------------------------------------------------------------------------------------------- per-proof context
// work area of one proof, in 32-byte scalars (cdp_prove_work_scalars): Montgomery form unless stated
struct work_t {
    fr_t *a_perm, *fac, *c, *d, *ucan /* canonical */, *x, *wc /* 2 x n/2, canonical */, *wd /* 2 x n/2 */, *ws /* 2 x n/2, canonical */, *sa, *sb, *sm;
};
// small per-proof scalars (Montgomery) in work_t::sm
enum {
    SM_K = 0, SM_MBL = 1 /* 4 */, SM_ALPHA_SP = 5, SM_BETA_SP, SM_GPROD, SM_BBL /* 4 */, SM_ALPHA_G = 12, SM_RBA /* 4 */, SM_RP = 17, SM_BETA_G, SM_Z,
    SM_ALPHA_I, SM_BETA_I, SM_T0, SM_T1, SM_T2, SM_T3, SM_COUNT = 64
};
// the randomness block of one proof (cdp_prove_random_scalars): raw Fr::rand outputs = Montgomery representations, in the reference's draw order
// except that r_d has its two solved-for entries (inner_product_argument.rs:53-77) in place
struct rnd_t {
    uint32_t n;
    fr_t *base;
    __device__ fr_t *a_bl() const { return base; }               // 2   curdleproofs.rs:86
    __device__ fr_t *c_bl() const { return base + 2; }           // 4   grand_product_argument.rs:75
    __device__ fr_t *r_c() const { return base + 6; }            // n   inner_product_argument.rs:46
    __device__ fr_t *r_d() const { return base + 6 + n; }        // n   :47 (n - 2 drawn, 2 solved)
    __device__ fr_t *tail() const { return base + 6 + 2 * n; }   // r_t r_u (curdleproofs.rs:110-111) r_a r_b r_k (same_scalar_argument.rs:56-58)
    __device__ fr_t *r_sm() const { return base + 11 + 2 * n; }  // n   same_multiscalar_argument.rs:78
};
enum { T_RT = 0, T_RU, T_RA, T_RB, T_RK };

struct pctx {
    uint32_t pr, ell, n, m;
    work_t w;
    rnd_t r;
    cstrobe s;
    uint8_t *buf;  // 64 shared bytes
    uint8_t *proof;
    uint8_t *side;        // A', D encodings
    uint32_t *scal;       // this proof's block of the emitted stage scalars
    const uint8_t *comp;  // the consumed stage's encodings (whole batch)
    const cdp_prove_dev *P;
};

// proof layout: byte offset of point `idx` in serialisation order (curdleproofs.rs:300-310), the scalars interleaved as the argument serialisers
// place them: r_p after C; c_final, d_final after the IPA vectors; z_k, z_t, z_u after the SameScalar commitments; x_final last
__device__ __forceinline__ uint32_t pt_off(uint32_t m, uint32_t idx) {
    return 48 * idx + (idx >= 9 ? 32u : 0u) + (idx >= 11 + 4 * m ? 64u : 0u) + (idx >= 15 + 4 * m ? 96u : 0u);
}
enum { P_A = 0, P_T1, P_T2, P_U1, P_U2, P_R, P_S, P_B, P_C, P_BC, P_BD, P_LC };
__device__ __forceinline__ uint32_t P_RC(uint32_t m) { return 11 + m; }
__device__ __forceinline__ uint32_t P_LD(uint32_t m) { return 11 + 2 * m; }
__device__ __forceinline__ uint32_t P_RD(uint32_t m) { return 11 + 3 * m; }
__device__ __forceinline__ uint32_t P_A1(uint32_t m) { return 11 + 4 * m; }
__device__ __forceinline__ uint32_t P_BA(uint32_t m) { return 15 + 4 * m; }
__device__ __forceinline__ uint32_t P_LA(uint32_t m) { return 18 + 4 * m; }
__device__ __forceinline__ uint32_t off_rp() { return 48 * 9; }
__device__ __forceinline__ uint32_t off_cfinal(uint32_t m) { return 48 * (11 + 4 * m) + 32; }
__device__ __forceinline__ uint32_t off_zk(uint32_t m) { return 48 * (15 + 4 * m) + 96; }
__device__ __forceinline__ uint32_t off_xfinal(uint32_t m) { return 48 * (18 + 10 * m) + 192; }

// encoding q of this proof among the consumed stage's outputs (launch order: sub-launch major, then proof, then segment)
__device__ __forceinline__ const uint8_t *stage_out(const pctx &c, uint32_t q) {
    const uint32_t e = c.P->out_map[q];
    return c.comp + 48 * ((size_t)c.P->batch * (e >> 16) + (size_t)c.pr * ((e >> 8) & 0xFF) + (e & 0xFF));
}
__device__ void copy48(uint8_t *dst, const uint8_t *src) { CTA_FOR(t, 48) dst[t] = src[t]; }
__device__ __forceinline__ void put_raw(uint32_t *dst, const fr_t &x) {
    for (int k = 0; k < 8; k++) dst[k] = x.v[k];
}
__device__ __forceinline__ fr_t canonical_one() {
    fr_t r = fr_zero();
    r.v[0] = 1;
    return r;
}
__device__ __forceinline__ fr_t fr_from_mont(const fr_t &a) { return fr_mul(a, canonical_one()); }
__device__ __forceinline__ fr_t fr_from_u32(uint32_t v) {
    fr_t r = fr_zero();
    r.v[0] = v;
    return fr_to_mont(r);
}
// a canonical scalar stored as bytes at a possibly unaligned address (proof bytes)
__device__ void store_scalar_bytes(uint8_t *dst, const fr_t &canon) {
    for (int k = 0; k < 8; k++)
        for (int b = 0; b < 4; b++) dst[4 * k + b] = (uint8_t)(canon.v[k] >> (8 * b));
}

// `transcript.append(label, &Fr)` (transcript.rs:29-33): the canonical 32-byte value
__device__ void t_append_fr(pctx &c, const uint8_t *label, uint32_t llen, const fr_t &x_mont) {
    if (CTA_LEADER) store_scalar_bytes(c.buf, fr_from_mont(x_mont));
    CTA_SYNC();
    cs_append(c.s, label, llen, c.buf, 32);
    CTA_SYNC();
}
__device__ void t_append_point(pctx &c, const uint8_t *label, uint32_t llen, const uint8_t *comp48) { cs_append(c.s, label, llen, comp48, 48); }
// get_and_append_challenge (transcript.rs:41-54): 64 challenge bytes, the first 32 with bit 255 cleared must be a non-zero value < r, else draw
// again; the accepted 32 bytes are appended under the same label.  Returns the challenge in Montgomery form (to every thread).
__device__ fr_t t_challenge(pctx &c, const uint8_t *label, uint32_t llen) {
    fr_t v;
    for (;;) {
        cs_challenge_bytes(c.s, label, llen, c.buf, 64);
        uint32_t nz = 0;
        for (int k = 0; k < 8; k++) {
            v.v[k] = (uint32_t)c.buf[4 * k] | ((uint32_t)c.buf[4 * k + 1] << 8) | ((uint32_t)c.buf[4 * k + 2] << 16) | ((uint32_t)c.buf[4 * k + 3] << 24);
            if (k == 7) v.v[k] &= 0x7FFFFFFFu;
            nz |= v.v[k];
        }
        const bool ok = nz != 0 && !fr_geq_mod(v.v);
        CTA_SYNC();  // everyone has read the bytes before the next draw may overwrite them
        if (ok) break;
    }
    if (CTA_LEADER) c.buf[31] &= 0x7F;
    CTA_SYNC();
    cs_append(c.s, label, llen, c.buf, 32);
    CTA_SYNC();
    return fr_to_mont(v);
}

// arr[0] <- sum of arr[0 .. cnt); the other entries are destroyed.  Callers synchronise before (after filling arr) -- and this returns synchronised.
__device__ void cta_reduce_add(fr_t *arr, uint32_t cnt) {
    uint32_t s = 1;
    while (s < cnt) s <<= 1;
    for (s >>= 1; s > 0; s >>= 1) {
        CTA_FOR(i, s) if (i + s < cnt) arr[i] = fr_add(arr[i], arr[i + s]);
        CTA_SYNC();
        cnt = cnt < s ? cnt : s;
    }
}
// x^-1 computed by the leader and handed to every thread through `slot`
__device__ fr_t cta_inverse(const fr_t &x, fr_t *slot) {
    if (CTA_LEADER) *slot = fr_inverse_safegcd(x);
    CTA_SYNC();
    const fr_t r = *slot;
    CTA_SYNC();
    return r;
}

// ---------------------------------------------------------------------------------------------------------------- round scalar emission
// IPA round k (split h = n >> (k + 1)), inner_product_argument.rs:152-161 over the original bases; scalars: cw (n) | ipL | ipR | dw (n)
//   cw[j] = Wc(j / 2h) c[sel j],  dw[j] = Wd(j / 2h) u_j d[sel j],  sel j = j mod h when bit h of j is set (R-half base, L-half scalar), h + j mod h otherwise
//   ipL = beta_i <c_L, d_R>, ipR = beta_i <c_R, d_L>   (the H term: H = beta crs.H, :140)
__device__ void emit_ipa_round(pctx &c, uint32_t k) {
    const uint32_t n = c.n, h = n >> (k + 1), Q = n / (2 * h);
    const fr_t *Wc = c.w.wc + (k & 1) * (n / 2), *Wd = c.w.wd + (k & 1) * (n / 2);
    CTA_FOR(j, n) {
        const uint32_t q = j / (2 * h), i = j & (h - 1), sel = (j & h) ? i : h + i;
        put_raw(c.scal + 8 * j, fr_mul(Wc[q], c.w.c[sel]));
        put_raw(c.scal + 8 * (n + 2 + j), fr_mul(fr_mul(Wd[q], c.w.ucan[j]), c.w.d[sel]));
    }
    (void)Q;
    CTA_FOR(i, h) {
        c.w.sa[i] = fr_mul(c.w.c[i], c.w.d[h + i]);
        c.w.sb[i] = fr_mul(c.w.c[h + i], c.w.d[i]);
    }
    CTA_SYNC();
    cta_reduce_add(c.w.sa, h);
    cta_reduce_add(c.w.sb, h);
    if (CTA_LEADER) {
        const fr_t beta_i = c.w.sm[SM_BETA_I];
        put_raw(c.scal + 8 * n, fr_from_mont(fr_mul(beta_i, c.w.sa[0])));
        put_raw(c.scal + 8 * (n + 1), fr_from_mont(fr_mul(beta_i, c.w.sb[0])));
    }
    CTA_SYNC();
}
// SameMSM round k, same_multiscalar_argument.rs:101-112; scalars: xw (n, over the original G_with_blinders) | x[0 .. 2h) (for the folded T, U)
__device__ void emit_sm_round(pctx &c, uint32_t k) {
    const uint32_t n = c.n, h = n >> (k + 1);
    const fr_t *Ws = c.w.ws + (k & 1) * (n / 2);
    CTA_FOR(j, n) {
        const uint32_t q = j / (2 * h), i = j & (h - 1), sel = (j & h) ? i : h + i;
        put_raw(c.scal + 8 * j, fr_mul(Ws[q], c.w.x[sel]));
    }
    CTA_FOR(i, 2 * h) put_raw(c.scal + 8 * (n + i), fr_from_mont(c.w.x[i]));
}
// From the switch round k0 on (cdp_prove_dev::switch_round) the folded bases are MATERIALISED: G^(k0)_i = sum_q Wc[q] G_{i + q n'} (n' = n >> k0
// entries, Q = 2^k0 original bases each: one strided fixed-base segment per entry), likewise G'^(k0) and G_with_blinders^(k0); the remaining rounds
// then run like the reference's (inner_product_argument.rs:158-179): MSMs of h pairs over the folded vectors, which are folded with gamma.
// Unfolded, every round costs n / 2 table pairs per cross term however short the vectors have become.
//   materialisation scalars, entry (i, q) at i * Q + q:   IPA: m1 = Wc[q] | m2 = Wd[q] u_{i + q n'}  (n each);  SameMSM: ms = Ws[q]  (n)
__device__ void emit_ipa_materialise(pctx &c, uint32_t k0) {
    const uint32_t n = c.n, ns = n >> k0, Q = 1u << k0;
    const fr_t *Wc = c.w.wc + (k0 & 1) * (n / 2), *Wd = c.w.wd + (k0 & 1) * (n / 2);
    CTA_FOR(e, n) {
        const uint32_t i = e / Q, q = e - i * Q;
        put_raw(c.scal + 8 * e, Wc[q]);
        put_raw(c.scal + 8 * (n + e), fr_mul(Wd[q], c.w.ucan[i + q * ns]));
    }
}
__device__ void emit_sm_materialise(pctx &c, uint32_t k0) {
    const uint32_t n = c.n, Q = 1u << k0;
    const fr_t *Ws = c.w.ws + (k0 & 1) * (n / 2);
    CTA_FOR(e, n) put_raw(c.scal + 8 * e, Ws[e & (Q - 1)]);
}
// IPA round k over the materialised vectors; scalars at offset 2n: c_L (h) | ipL | c_R (h) | ipR | d (2h), all canonical
__device__ void emit_ipa_round_folded(pctx &c, uint32_t k) {
    const uint32_t n = c.n, h = n >> (k + 1);
    uint32_t *sc = c.scal + 8 * (2 * n);
    CTA_FOR(i, h) {
        put_raw(sc + 8 * i, fr_from_mont(c.w.c[i]));
        put_raw(sc + 8 * (h + 1 + i), fr_from_mont(c.w.c[h + i]));
        c.w.sa[i] = fr_mul(c.w.c[i], c.w.d[h + i]);
        c.w.sb[i] = fr_mul(c.w.c[h + i], c.w.d[i]);
    }
    CTA_FOR(i, 2 * h) put_raw(sc + 8 * (2 * h + 2 + i), fr_from_mont(c.w.d[i]));
    CTA_SYNC();
    cta_reduce_add(c.w.sa, h);
    cta_reduce_add(c.w.sb, h);
    if (CTA_LEADER) {
        const fr_t beta_i = c.w.sm[SM_BETA_I];
        put_raw(sc + 8 * h, fr_from_mont(fr_mul(beta_i, c.w.sa[0])));
        put_raw(sc + 8 * (2 * h + 1), fr_from_mont(fr_mul(beta_i, c.w.sb[0])));
    }
    CTA_SYNC();
}
// SameMSM round k over the materialised G_with_blinders: only x[0 .. 2h) is needed (at offset n, where the unfolded form keeps it too)
__device__ void emit_sm_round_folded(pctx &c, uint32_t k) {
    const uint32_t n = c.n, h = n >> (k + 1);
    CTA_FOR(i, 2 * h) put_raw(c.scal + 8 * (n + i), fr_from_mont(c.w.x[i]));
}
// prefix weights of the next round: the new low bit of the prefix is bit h of the base index; set = R half = weight x g
__device__ void grow_weights(fr_t *W, uint32_t half, uint32_t k, const fr_t &g) {
    const uint32_t Q = 1u << k;
    const fr_t *src = W + (k & 1) * half;
    fr_t *dst = W + ((k + 1) & 1) * half;
    CTA_FOR(q, Q) {
        const fr_t w = src[q];
        dst[2 * q] = w;
        dst[2 * q + 1] = fr_mul(w, g);
    }
}

// ---------------------------------------------------------------------------------------------------------------- stages
// curdleproofs.rs:86-116: a_perm, the blinders, and the scalars of everything that depends only on vec_a and the prover's randomness
__device__ void stage_s1(pctx &c) {
    const uint32_t ell = c.ell, n = c.n;
    const cdp_prove_dev &P = *c.P;
    const uint32_t *va = reinterpret_cast<const uint32_t *>(P.d_vec_a) + 8 * (size_t)c.pr * ell;
    const uint32_t *perm = P.d_perm + (size_t)c.pr * ell;
    const uint32_t *wit = reinterpret_cast<const uint32_t *>(P.d_witness) + 8 * (size_t)c.pr * 5;
    // scalars: a_perm | r_a' (n) | vec_a (ell) | r_sm (n) | r_t | r_u | r_a | r_b        (host/prover.cpp stage 1)
    const uint32_t sa = n, sR = n + ell, sx = 2 * n + ell;
    CTA_FOR(i, ell) {
        const fr_t a = fr_load(va + 8 * perm[i]);   // vec_a_permuted, curdleproofs.rs:87
        put_raw(c.scal + 8 * i, a);
        c.w.a_perm[i] = fr_to_mont(a);
        put_raw(c.scal + 8 * (sa + i), fr_load(va + 8 * i));
    }
    CTA_FOR(i, 4) {
        put_raw(c.scal + 8 * (ell + i), i < 2 ? fr_from_mont(c.r.a_bl()[i]) : fr_zero());  // r_a' = (r_a0, r_a1, 0, 0), :88-92
        put_raw(c.scal + 8 * (sx + i), fr_from_mont(c.r.tail()[i]));                        // r_t, r_u, r_a, r_b
        c.w.sm[SM_MBL + i] = fr_to_mont(fr_load(wit + 8 * (1 + i)));
    }
    CTA_FOR(i, n) put_raw(c.scal + 8 * (sR + i), fr_from_mont(c.r.r_sm()[i]));
    if (CTA_LEADER) c.w.sm[SM_K] = fr_to_mont(fr_load(wit));
}

// same_permutation_argument.rs:60-82
__device__ void stage_sameperm(pctx &c) {
    const uint32_t ell = c.ell, m = c.m;
    const cdp_prove_dev &P = *c.P;
    // stage-1 outputs (host/prover.cpp Stage1Out order: A, R, S, B_a, B_t, B_u, T_1, U_1, A_1, B_1) into the proof
    {
        const uint32_t dst[10] = {P_A, P_R, P_S, P_BA(m), P_BA(m) + 1, P_BA(m) + 2, P_T1, P_U1, P_A1(m), P_A1(m) + 2};
        for (int q = 0; q < 10; q++) copy48(c.proof + pt_off(m, dst[q]), stage_out(c, q));
    }
    const uint8_t *va = P.d_vec_a + 32 * (size_t)c.pr * ell;
    t_append_point(c, L_SP1, 15, stage_out(c, 0));
    t_append_point(c, L_SP1, 15, P.d_comp0_M + 48 * (size_t)c.pr);
    cs_append_header(c.s, L_SP1, 15, 8 + 32 * ell);  // a Vec<Fr>: u64 length, then the elements
    cs_value(c.s, (uint64_t)ell, 8);
    cs_absorb(c.s, va, 32 * ell);
    const fr_t alpha = t_challenge(c, L_SPA, 15), beta = t_challenge(c, L_SPB, 14);
    const uint32_t *perm = P.d_perm + (size_t)c.pr * ell;
    // b_i = a_sigma(i) + sigma(i) alpha + beta (:70-74); c = exclusive prefix product (grand_product_argument.rs:65-71), gprod = prod b_i
    CTA_FOR(i, ell) {
        const fr_t f = fr_add(fr_add(c.w.a_perm[i], fr_mul(fr_from_u32(perm[i]), alpha)), beta);
        c.w.fac[i] = f;
        c.w.sa[i] = f;
    }
    CTA_SYNC();
    fr_t *src = c.w.sa, *dst = c.w.sb;
    for (uint32_t d = 1; d < ell; d <<= 1) {  // inclusive scan of the products, Hillis-Steele
        CTA_FOR(i, ell) dst[i] = i >= d ? fr_mul(src[i - d], src[i]) : src[i];
        CTA_SYNC();
        fr_t *t = src; src = dst; dst = t;
    }
    CTA_FOR(i, ell) c.w.c[i] = i == 0 ? fr_one() : src[i - 1];
    CTA_FOR(i, 4) {
        const fr_t rap = i < 2 ? c.r.a_bl()[i] : fr_zero();
        c.w.sm[SM_BBL + i] = fr_add(rap, fr_mul(alpha, c.w.sm[SM_MBL + i]));  // vec_b_blinders, :78-81
    }
    if (CTA_LEADER) {
        c.w.sm[SM_ALPHA_SP] = alpha;
        c.w.sm[SM_BETA_SP] = beta;
        c.w.sm[SM_GPROD] = src[ell - 1];
        // stage 2: B = 1 A + alpha M + beta sum(G) | T_2 = k R + r_t H | A_2 = r_k R + r_a H | U_2 = k S + r_u H | B_2 = r_k S + r_b H | A' = 1 A + r_t G_t + r_u G_u
        const fr_t one = canonical_one(), k = fr_from_mont(c.w.sm[SM_K]), r_t = fr_from_mont(c.r.tail()[T_RT]), r_u = fr_from_mont(c.r.tail()[T_RU]),
                   r_a = fr_from_mont(c.r.tail()[T_RA]), r_b = fr_from_mont(c.r.tail()[T_RB]), r_k = fr_from_mont(c.r.tail()[T_RK]);
        const fr_t v2[14] = {one, fr_from_mont(alpha), fr_from_mont(beta), k, r_t, r_k, r_a, k, r_u, r_k, r_b, one, r_t, r_u};
        for (int i = 0; i < 14; i++) put_raw(c.scal + 8 * i, v2[i]);
    }
}

// grand_product_argument.rs:63-83
__device__ void stage_gprod1(pctx &c) {
    const uint32_t ell = c.ell, n = c.n, m = c.m;
    // stage-2 outputs: B, T_2, A_2, U_2, B_2, A'
    {
        const uint32_t dst[5] = {P_B, P_T2, P_A1(m) + 1, P_U2, P_A1(m) + 3};
        for (int q = 0; q < 5; q++) copy48(c.proof + pt_off(m, dst[q]), stage_out(c, q));
        copy48(c.side, stage_out(c, 5));
    }
    t_append_point(c, L_GP1, 11, stage_out(c, 0));
    t_append_fr(c, L_GP1, 11, c.w.sm[SM_GPROD]);
    const fr_t alpha = t_challenge(c, L_GPA, 11);
    CTA_FOR(i, 4) {
        c.w.c[ell + i] = c.r.c_bl()[i];                                // vec_c | r_c blinders, :75-76
        c.w.sm[SM_RBA + i] = fr_add(c.w.sm[SM_BBL + i], alpha);        // r_b + alpha, :78-81
    }
    CTA_SYNC();
    if (CTA_LEADER) {
        c.w.sm[SM_ALPHA_G] = alpha;
        fr_t rp = fr_zero();
        for (int i = 0; i < 4; i++) rp = fr_add(rp, fr_mul(c.w.sm[SM_RBA + i], c.r.c_bl()[i]));  // r_p = <r_b + alpha, r_c>, :82
        c.w.sm[SM_RP] = rp;
        store_scalar_bytes(c.proof + off_rp(), fr_from_mont(rp));
    }
    CTA_FOR(i, n) put_raw(c.scal + 8 * i, fr_from_mont(c.w.c[i]));  // stage 3: C = msm(G | Hvec, c | r_c)
}

// grand_product_argument.rs:85-147 and generate_ipa_blinders (inner_product_argument.rs:53-77), :124-127
__device__ void stage_gprod2(pctx &c) {
    const uint32_t ell = c.ell, n = c.n, m = c.m;
    copy48(c.proof + pt_off(m, P_C), stage_out(c, 0));
    t_append_point(c, L_GP2, 11, stage_out(c, 0));
    t_append_fr(c, L_GP2, 11, c.w.sm[SM_RP]);
    const fr_t beta = t_challenge(c, L_GPB, 10);
    fr_t *rc = c.r.r_c(), *rd = c.r.r_d();
    // one field inversion for beta^-1, c[n-2]^-1 and the blinder denominator: with E = r_c[n-1] c[n-2] - r_c[n-2] c[n-1],
    // 1 / (r_c[n-1] - r_c[n-2] c[n-1] / c[n-2]) = c[n-2] / E
    if (CTA_LEADER) {
        const fr_t c2 = c.w.c[n - 2], E = fr_sub(fr_mul(rc[n - 1], c2), fr_mul(rc[n - 2], c.w.c[n - 1]));
        const fr_t p01 = fr_mul(beta, c2), inv = fr_inverse_safegcd(fr_mul(p01, E));
        const fr_t invE = fr_mul(inv, p01), inv01 = fr_mul(inv, E);
        c.w.sm[SM_T0] = fr_mul(inv01, c2);      // beta^-1
        c.w.sm[SM_T1] = fr_mul(inv01, beta);    // c[n-2]^-1
        c.w.sm[SM_T2] = fr_mul(invE, c2);       // the blinder denominator's inverse
        c.w.sm[SM_BETA_G] = beta;
    }
    CTA_SYNC();
    const fr_t beta_inv = c.w.sm[SM_T0], inv_c = c.w.sm[SM_T1], inv2 = c.w.sm[SM_T2];
    const fr_t beta_l = fr_pow_u32(beta, ell), beta_l1 = fr_mul(beta_l, beta);
    // u_i = beta^-(i+1), beta^-(ell+1) on the blinder positions (:92-102, :125): kept canonical, it only ever multiplies other scalars
    // d_i = b_i beta^(i+1) - beta^i (:113-123), beta^(ell+1) (r_b + alpha) on the blinder positions (:126-129)
    CTA_FOR(i, n) {
        const uint32_t e = i < ell ? i + 1 : ell + 1;
        c.w.ucan[i] = fr_from_mont(fr_pow_u32(beta_inv, e));
        if (i < ell) {
            const fr_t bi = i == 0 ? fr_one() : fr_pow_u32(beta, i);
            c.w.d[i] = fr_sub(fr_mul(c.w.fac[i], fr_mul(bi, beta)), bi);
        } else {
            c.w.d[i] = fr_mul(beta_l1, c.w.sm[SM_RBA + (i - ell)]);
        }
    }
    CTA_SYNC();
    // omega = <r_c, d> + <r_d, c>, delta = <r_c, r_d> over the n - 2 drawn entries of r_d
    CTA_FOR(i, n) {
        fr_t t = fr_mul(rc[i], c.w.d[i]);
        if (i + 2 < n) {
            t = fr_add(t, fr_mul(rd[i], c.w.c[i]));
            c.w.sb[i] = fr_mul(rc[i], rd[i]);
        } else {
            c.w.sb[i] = fr_zero();
        }
        c.w.sa[i] = t;
    }
    CTA_SYNC();
    cta_reduce_add(c.w.sa, n);
    cta_reduce_add(c.w.sb, n);
    if (CTA_LEADER) {
        const fr_t omega = c.w.sa[0], delta = c.w.sb[0];
        const fr_t last_z = fr_mul(fr_sub(fr_mul(fr_mul(rc[n - 2], inv_c), omega), delta), inv2);
        const fr_t pen_z = fr_mul(fr_neg(inv_c), fr_add(fr_mul(last_z, c.w.c[n - 1]), omega));
        rd[n - 2] = pen_z;
        rd[n - 1] = last_z;
        // inner_prod = r_p beta^(ell+1) + gprod beta^ell - 1   (:131-141)
        c.w.sm[SM_Z] = fr_sub(fr_add(fr_mul(c.w.sm[SM_RP], beta_l1), fr_mul(c.w.sm[SM_GPROD], beta_l)), fr_one());
        // stage 4: D = 1 B - beta^-1 sum(G) + alpha_g sum(Hvec) | B_c = msm(G | Hvec, r_c) | B_d = msm(G', r_d) = msm(G | Hvec, r_d o u)
        put_raw(c.scal, canonical_one());
        put_raw(c.scal + 8, fr_from_mont(fr_neg(beta_inv)));
        put_raw(c.scal + 16, fr_from_mont(c.w.sm[SM_ALPHA_G]));
    }
    CTA_SYNC();
    CTA_FOR(i, n) {
        put_raw(c.scal + 8 * (3 + i), fr_from_mont(rc[i]));
        put_raw(c.scal + 8 * (3 + n + i), fr_mul(rd[i], c.w.ucan[i]));
    }
}

// inner_product_argument.rs:129-140, then the scalars of round 0
__device__ void stage_ipa0(pctx &c) {
    const uint32_t n = c.n, m = c.m;
    copy48(c.side + 48, stage_out(c, 0));  // D
    copy48(c.proof + pt_off(m, P_BC), stage_out(c, 1));
    copy48(c.proof + pt_off(m, P_BD), stage_out(c, 2));
    CTA_SYNC();  // C is read back from the proof bytes written by an earlier kernel; D, B_c, B_d from the stage outputs
    t_append_point(c, L_IP1, 9, c.proof + pt_off(m, P_C));
    t_append_point(c, L_IP1, 9, stage_out(c, 0));
    t_append_fr(c, L_IP1, 9, c.w.sm[SM_Z]);
    t_append_point(c, L_IP1, 9, stage_out(c, 1));
    t_append_point(c, L_IP1, 9, stage_out(c, 2));
    const fr_t alpha = t_challenge(c, L_IPA, 9), beta = t_challenge(c, L_IPB, 8);
    const fr_t *rc = c.r.r_c(), *rd = c.r.r_d();
    CTA_FOR(i, n) {  // c <- r_c + alpha c, d <- r_d + alpha d   (:133-138)
        c.w.c[i] = fr_add(rc[i], fr_mul(alpha, c.w.c[i]));
        c.w.d[i] = fr_add(rd[i], fr_mul(alpha, c.w.d[i]));
    }
    if (CTA_LEADER) {
        c.w.sm[SM_ALPHA_I] = alpha;
        c.w.sm[SM_BETA_I] = beta;
        c.w.wc[0] = canonical_one();
        c.w.wd[0] = fr_one();
    }
    CTA_SYNC();
    emit_ipa_round(c, 0);
}

// same_scalar_argument.rs:64-75 and same_multiscalar_argument.rs:84-91, then the scalars of SameMSM round 0
__device__ void stage_same_scalar_and_msm_step1(pctx &c) {
    const uint32_t ell = c.ell, n = c.n, m = c.m;
    const cdp_prove_dev &P = *c.P;
    {
        const uint32_t order[10] = {P_R, P_S, P_T1, P_T2, P_U1, P_U2, P_A1(m), P_A1(m) + 1, P_A1(m) + 2, P_A1(m) + 3};
        for (int q = 0; q < 10; q++) t_append_point(c, L_SEP, 14, c.proof + pt_off(m, order[q]));
    }
    const fr_t alpha = t_challenge(c, L_SSA, 17);
    if (CTA_LEADER) {
        const fr_t *t = c.r.tail();
        const fr_t z_k = fr_add(t[T_RK], fr_mul(c.w.sm[SM_K], alpha)), z_t = fr_add(t[T_RA], fr_mul(t[T_RT], alpha)),
                   z_u = fr_add(t[T_RB], fr_mul(t[T_RU], alpha));
        uint8_t *w = c.proof + off_zk(m);
        store_scalar_bytes(w, fr_from_mont(z_k));
        store_scalar_bytes(w + 32, fr_from_mont(z_t));
        store_scalar_bytes(w + 64, fr_from_mont(z_u));
    }
    t_append_point(c, L_SM1, 14, c.side);  // A'
    t_append_point(c, L_SM1, 14, c.proof + pt_off(m, P_T2));
    t_append_point(c, L_SM1, 14, c.proof + pt_off(m, P_U2));
    for (uint32_t v = 0; v < 2; v++) {  // vec_T_with_blinders = T | O O H O, vec_U_with_blinders = U | O O O H   (curdleproofs.rs:142-155)
        cs_append_header(c.s, L_SM1, 14, 8 + 48 * n);
        cs_value(c.s, (uint64_t)n, 8);
        cs_absorb(c.s, P.d_comp0_vecs + 48 * (((size_t)c.pr * 4 + 2 + v) * ell), 48 * ell);
        for (uint32_t q = 0; q < 4; q++) {
            if (q == 2 + v) {
                cs_absorb(c.s, P.d_comp_H, 48);
            } else {
                cs_byte(c.s, 0xC0);
                for (int z = 1; z < 48; z++) cs_byte(c.s, 0);
            }
        }
    }
    for (uint32_t q = 0; q < 3; q++) t_append_point(c, L_SM1, 14, c.proof + pt_off(m, P_BA(m) + q));
    const fr_t a_sm = t_challenge(c, L_SMA, 14);
    const fr_t *rs = c.r.r_sm();
    CTA_FOR(i, n) {  // x = r + alpha (a_perm | r_a0 r_a1 r_t r_u)   (:93-97, curdleproofs.rs:157-160)
        fr_t a;
        if (i < ell) a = c.w.a_perm[i];
        else if (i < ell + 2) a = c.r.a_bl()[i - ell];
        else a = c.r.tail()[i - ell - 2];
        c.w.x[i] = fr_add(rs[i], fr_mul(a_sm, a));
    }
    if (CTA_LEADER) c.w.ws[0] = canonical_one();
    CTA_SYNC();
    emit_sm_round(c, 0);
}

// inner_product_argument.rs:162-186 for round k
__device__ void stage_ipa_round(pctx &c, uint32_t k) {
    const uint32_t n = c.n, m = c.m, h = n >> (k + 1);
    {   // round outputs L_C, L_D, R_C, R_D; serialised as vec_L_C, vec_R_C, vec_L_D, vec_R_D
        const uint32_t dst[4] = {P_LC + k, P_LD(m) + k, P_RC(m) + k, P_RD(m) + k};
        for (int q = 0; q < 4; q++) {
            copy48(c.proof + pt_off(m, dst[q]), stage_out(c, q));
            t_append_point(c, L_IPL, 8, stage_out(c, q));
        }
    }
    const fr_t gamma = t_challenge(c, L_IPG, 9);
    const fr_t gamma_inv = cta_inverse(gamma, c.w.sm + SM_T0);
    CTA_FOR(i, h) {  // c = c_L + gamma^-1 c_R, d = d_L + gamma d_R   (:174-176)
        c.w.c[i] = fr_add(c.w.c[i], fr_mul(gamma_inv, c.w.c[h + i]));
        c.w.d[i] = fr_add(c.w.d[i], fr_mul(gamma, c.w.d[h + i]));
    }
    if (h > 1) {  // G = G_L + gamma G_R, G' = G'_L + gamma^-1 G'_R (:177-178): as weights on the original bases, or -- from the switch round on -- a fold launch
        const uint32_t k0 = c.P->switch_round;
        if (k + 1 <= k0) {
            grow_weights(c.w.wc, n / 2, k, gamma);
            grow_weights(c.w.wd, n / 2, k, gamma_inv);
        }
        if (k >= k0 && CTA_LEADER) {
            uint32_t *fs = reinterpret_cast<uint32_t *>(c.P->d_fold_scalars) + 16 * (size_t)c.pr;
            put_raw(fs, fr_from_mont(gamma));
            put_raw(fs + 8, fr_from_mont(gamma_inv));
        }
        CTA_SYNC();
        if (k + 1 < k0) {
            emit_ipa_round(c, k + 1);
        } else {
            if (k + 1 == k0) emit_ipa_materialise(c, k0);
            emit_ipa_round_folded(c, k + 1);
        }
        return;
    }
    CTA_SYNC();
    if (CTA_LEADER) {
        store_scalar_bytes(c.proof + off_cfinal(m), fr_from_mont(c.w.c[0]));
        store_scalar_bytes(c.proof + off_cfinal(m) + 32, fr_from_mont(c.w.d[0]));
    }
    stage_same_scalar_and_msm_step1(c);
}

// same_multiscalar_argument.rs:113-136 for round k
__device__ void stage_sm_round(pctx &c, uint32_t k) {
    const uint32_t n = c.n, m = c.m, h = n >> (k + 1);
    for (uint32_t q = 0; q < 6; q++) {  // L_A, L_T, L_U, R_A, R_T, R_U -- also the serialisation order of the six vectors
        copy48(c.proof + pt_off(m, P_LA(m) + q * m + k), stage_out(c, q));
        t_append_point(c, L_SML, 13, stage_out(c, q));
    }
    const fr_t gamma = t_challenge(c, L_SMG, 14);
    const fr_t gamma_inv = cta_inverse(gamma, c.w.sm + SM_T0);
    CTA_FOR(i, h) c.w.x[i] = fr_add(c.w.x[i], fr_mul(gamma_inv, c.w.x[h + i]));  // x = x_L + gamma^-1 x_R   (:126)
    if (h > 1) {
        // T = T_L + gamma T_R, U likewise: the fold launch reads gamma; G_with_blinders = G_L + gamma G_R as weights, or folded too   (:128-130)
        const uint32_t k0 = c.P->switch_round;
        if (CTA_LEADER) put_raw(reinterpret_cast<uint32_t *>(c.P->d_fold_scalars) + 8 * (size_t)c.pr, fr_from_mont(gamma));
        if (k + 1 <= k0) grow_weights(c.w.ws, n / 2, k, gamma);
        CTA_SYNC();
        if (k + 1 < k0) {
            emit_sm_round(c, k + 1);
        } else {
            if (k + 1 == k0) emit_sm_materialise(c, k0);
            emit_sm_round_folded(c, k + 1);
        }
        return;
    }
    CTA_SYNC();
    if (CTA_LEADER) store_scalar_bytes(c.proof + off_xfinal(m), fr_from_mont(c.w.x[0]));
}

}  // namespace prove

// One CTA per proof; `stage` / `round` select the step (cdp_prove_stage_dev in include/cdp_msm.h).
__global__ void __launch_bounds__(256) k_prove_stage(const __grid_constant__ cdp_prove_dev P, const int stage, const uint32_t round) {
    using namespace prove;
    CTA_SHARED uint64_t sh_state[25];
    CTA_SHARED uint8_t sh_buf[64];
    pctx c;
    c.pr = CTA_PROOF_INDEX;
    if (c.pr >= P.batch) return;
    c.ell = P.ell; c.n = P.ell + 4; c.m = P.m;
    c.P = &P;
    const uint32_t n = c.n;
    fr_t *w = reinterpret_cast<fr_t *>(P.d_work) + (size_t)c.pr * (11 * (size_t)n + SM_COUNT);
    c.w.a_perm = w; c.w.fac = w + n; c.w.c = w + 2 * n; c.w.d = w + 3 * n; c.w.ucan = w + 4 * n; c.w.x = w + 5 * n;
    c.w.wc = w + 6 * n; c.w.wd = w + 7 * n; c.w.ws = w + 8 * n; c.w.sa = w + 9 * n; c.w.sb = w + 10 * n; c.w.sm = w + 11 * n;
    c.r.n = n;
    c.r.base = reinterpret_cast<fr_t *>(P.d_random) + (size_t)c.pr * (3 * (size_t)n + 11);
    c.buf = sh_buf;
    c.proof = P.d_proofs + (size_t)c.pr * P.proof_bytes;
    c.side = P.d_side + (size_t)c.pr * 96;
    c.scal = reinterpret_cast<uint32_t *>(P.d_scalars) + 8 * (size_t)c.pr * P.scalars_per_proof;
    c.comp = P.d_comp;
    uint64_t *gstate = reinterpret_cast<uint64_t *>(P.d_state) + (size_t)c.pr * 26;
    CTA_FOR(i, 25) sh_state[i] = gstate[i];
    c.s.st = reinterpret_cast<uint8_t *>(sh_state);
    c.s.pos = (uint32_t)(gstate[25] & 0xFF);
    c.s.pos_begin = (uint32_t)((gstate[25] >> 8) & 0xFF);
    CTA_SYNC();
    switch (stage) {
        case CDP_PS_S1: stage_s1(c); break;
        case CDP_PS_SAMEPERM: stage_sameperm(c); break;
        case CDP_PS_GPROD1: stage_gprod1(c); break;
        case CDP_PS_GPROD2: stage_gprod2(c); break;
        case CDP_PS_IPA0: stage_ipa0(c); break;
        case CDP_PS_IPA_ROUND: stage_ipa_round(c, round); break;
        case CDP_PS_SM_ROUND: stage_sm_round(c, round); break;
        default: break;
    }
    CTA_SYNC();
    CTA_FOR(i, 25) gstate[i] = sh_state[i];
    if (CTA_LEADER) gstate[25] = (uint64_t)c.s.pos | ((uint64_t)c.s.pos_begin << 8);
}

#ifndef CDP_PROVE_HOST_HARNESS
// ---- the prover's randomness on the device.  The reference takes `rng: &mut impl RngCore` (/root/reference/src/curdleproofs.rs:74; its tests:
// `StdRng` = ChaCha12, rand 0.8) and every draw of `CurdleproofsProof::new` is `Fr::rand` (ark-ff: four u64 = eight consecutive words of the
// stream, top bit cleared, rejected when >= r, taken as the Montgomery representation).  So attempt a of a proof sits at words
// [skip + 8a, skip + 8a + 8) of its keystream -- independent of every other attempt -- and the i-th ACCEPTED one is the i-th draw.  One warp
// per proof: lane l computes the ChaCha12 blocks of attempt a0 + l, a ballot ranks the accepted ones, and they are stored in the draw order
// of host/prover.cpp (the two r_d entries the device solves for are left zero).  The host hands over 32-byte keys instead of 25 KB of
// scalars per proof and no longer runs the cipher (it was most of the host's time per step).
__device__ __forceinline__ uint32_t cc_rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
__device__ __forceinline__ void cc_qr(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    a += b; d = cc_rotl(d ^ a, 16);
    c += d; b = cc_rotl(b ^ c, 12);
    a += b; d = cc_rotl(d ^ a, 8);
    c += d; b = cc_rotl(b ^ c, 7);
}
// block `ctr` of the ChaCha12 keystream (64-bit counter, stream 0: rand_chacha's ChaCha12Rng) to out[0..16)
__device__ __forceinline__ void chacha12_block(const uint32_t *key, uint64_t ctr, uint32_t *out) {
    const uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                             key[4], key[5], key[6], key[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3], x4 = in[4], x5 = in[5], x6 = in[6], x7 = in[7], x8 = in[8], x9 = in[9], x10 = in[10],
             x11 = in[11], x12 = in[12], x13 = in[13], x14 = in[14], x15 = in[15];
#pragma unroll 1
    for (int r = 0; r < 6; r++) {
        cc_qr(x0, x4, x8, x12); cc_qr(x1, x5, x9, x13); cc_qr(x2, x6, x10, x14); cc_qr(x3, x7, x11, x15);
        cc_qr(x0, x5, x10, x15); cc_qr(x1, x6, x11, x12); cc_qr(x2, x7, x8, x13); cc_qr(x3, x4, x9, x14);
    }
    out[0] = x0 + in[0]; out[1] = x1 + in[1]; out[2] = x2 + in[2]; out[3] = x3 + in[3]; out[4] = x4 + in[4]; out[5] = x5 + in[5];
    out[6] = x6 + in[6]; out[7] = x7 + in[7]; out[8] = x8 + in[8]; out[9] = x9 + in[9]; out[10] = x10 + in[10]; out[11] = x11 + in[11];
    out[12] = x12 + in[12]; out[13] = x13 + in[13]; out[14] = x14 + in[14]; out[15] = x15 + in[15];
}
__global__ void __launch_bounds__(32) k_prove_random(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ skip_words, uint32_t n,
                                                     uint32_t nrnd, uint32_t *__restrict__ out) {
    __shared__ uint32_t sm[32][33];  // lane l's two blocks (32 words) in row l
    const uint32_t pr = blockIdx.x, lane = threadIdx.x;
    uint32_t key[8];
    for (int k = 0; k < 8; k++) key[k] = keys[8 * (size_t)pr + k];
    const uint64_t sk = skip_words ? skip_words[pr] : 0;
    uint32_t *o = out + 8 * (size_t)pr * nrnd;
    const uint32_t need = 3 * n + 9, solved = 6 + 2 * n - 2;  // draws; first of the two slots the device solves for (left zero)
    if (lane < 16) o[8 * (size_t)solved + lane] = 0;
    uint32_t accepted = 0;
#pragma unroll 1
    for (uint64_t a0 = 0; accepted < need; a0 += 32) {
        const uint64_t wo = sk + 8 * (a0 + lane);
        const uint32_t off = (uint32_t)(wo & 15);
        chacha12_block(key, wo >> 4, &sm[lane][0]);
        if (off > 8) chacha12_block(key, (wo >> 4) + 1, &sm[lane][16]);  // the eight words run into the next block
        vcoef::fr_t v;
        for (int k = 0; k < 8; k++) v.v[k] = sm[lane][off + k];
        v.v[7] &= 0x7FFFFFFFu;
        const bool ok = !vcoef::fr_geq_mod(v.v);
        const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
        const uint32_t pos = accepted + __popc(ballot & ((1u << lane) - 1u));
        if (ok && pos < need) {
            uint32_t *dst = o + 8 * (size_t)(pos < solved ? pos : pos + 2);
            for (int k = 0; k < 8; k++) dst[k] = v.v[k];
        }
        accepted += __popc(ballot);
    }
}
cudaError_t launch_prove_random(cudaStream_t st, const uint32_t *keys, const uint64_t *skip_words, uint32_t ell, uint32_t batch, uint32_t *out) {
    if (batch == 0) return cudaSuccess;
    k_prove_random<<<batch, 32, 0, st>>>(keys, skip_words, ell + 4, 3 * (ell + 4) + 11, out);
    return cudaGetLastError();
}
cudaError_t launch_prove_stage(cudaStream_t st, const cdp_prove_dev &P, int stage, uint32_t round) {
    if (P.batch == 0) return cudaSuccess;
    // one warp per proof by default: the steps are latency-bound (one thread hashes), so a CTA should hold as little of an SM as possible
    // while the throughput kernels of the other lanes run beside it; CDP_PROVE_THREADS (32 .. 256) for tuning runs
    static const int threads = [] { const char *e = getenv("CDP_PROVE_THREADS"); int v = e ? atoi(e) : 32; return v >= 32 && v <= 256 && v % 32 == 0 ? v : 32; }();
    k_prove_stage<<<P.batch, threads, 0, st>>>(P, stage, round);
    return cudaGetLastError();
}
#endif

}  // namespace cdp
