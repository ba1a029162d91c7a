// Device-side opening of the Curdleproofs Fiat-Shamir transcript -- SURVEY.md section 8(f) rank 1, the part that carries the bulk
// of the hashing: `CurdleproofsProof::new` / `verify` start with (/root/reference/src/curdleproofs.rs:78-83 and :213-225)
//     transcript = Transcript::new(b"curdleproofs");
//     transcript.append_list(b"curdleproofs_step1", &[vec_R, vec_S, vec_T, vec_U]);   // 4 * ell compressed points
//     transcript.append(b"curdleproofs_step1", M);
//     vec_a = transcript.get_and_append_challenges(b"curdleproofs_vec_a", ell);        // ell challenges
// which is ~640 of the ~900 Keccak-f[1600] permutations a whole ell = 252 proof needs.  The compressed encodings are produced on the
// GPU anyway, so one thread per proof runs merlin 3.0.0 / STROBE-128 over them here and hands the host vec_a plus the 200-byte STROBE
// state to continue from (/root/reference/src/transcript.rs:28-61).  Integer / byte work only.
#ifndef CDP_TRANSCRIPT_HOST_HARNESS  // tests/host/transcript_dev_check.cpp compiles this file with g++ to check it on the CPU
#include "launch.h"
#include "constants.cuh"
#endif
#include "fr256.cuh"

namespace cdp {

namespace {

__device__ __forceinline__ uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }

__device__ __noinline__ void keccak_f1600(uint64_t *A) {
    const uint64_t RC[24] = {0x1ULL, 0x8082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x808bULL, 0x80000001ULL,
                             0x8000000080008081ULL, 0x8000000000008009ULL, 0x8aULL, 0x88ULL, 0x80008009ULL, 0x8000000aULL,
                             0x8000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                             0x8000000000008002ULL, 0x8000000000000080ULL, 0x800aULL, 0x800000008000000aULL,
                             0x8000000080008081ULL, 0x8000000000008080ULL, 0x80000001ULL, 0x8000000080008008ULL};
    uint64_t a00 = A[0], a10 = A[1], a20 = A[2], a30 = A[3], a40 = A[4], a01 = A[5], a11 = A[6], a21 = A[7], a31 = A[8], a41 = A[9],
             a02 = A[10], a12 = A[11], a22 = A[12], a32 = A[13], a42 = A[14], a03 = A[15], a13 = A[16], a23 = A[17], a33 = A[18],
             a43 = A[19], a04 = A[20], a14 = A[21], a24 = A[22], a34 = A[23], a44 = A[24];
    uint64_t c0, c1, c2, c3, c4, d0, d1, d2, d3, d4;
    uint64_t b00, b10, b20, b30, b40, b01, b11, b21, b31, b41, b02, b12, b22, b32, b42, b03, b13, b23, b33, b43, b04, b14, b24, b34, b44;
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
#include "../host/keccak_round.inc"
        a00 ^= RC[round];
    }
    A[0] = a00; A[1] = a10; A[2] = a20; A[3] = a30; A[4] = a40; A[5] = a01; A[6] = a11; A[7] = a21; A[8] = a31; A[9] = a41;
    A[10] = a02; A[11] = a12; A[12] = a22; A[13] = a32; A[14] = a42; A[15] = a03; A[16] = a13; A[17] = a23; A[18] = a33; A[19] = a43;
    A[20] = a04; A[21] = a14; A[22] = a24; A[23] = a34; A[24] = a44;
}

// STROBE-128 (rate 166) restricted to the three operations merlin uses: meta-AD, AD, PRF
struct strobe_t {
    uint64_t st[25];
    uint32_t pos, pos_begin;
};
constexpr uint32_t STROBE_R = 166;
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_M = 16 };

__device__ __forceinline__ void st_xor_byte(strobe_t &s, uint32_t at, uint32_t v) { s.st[at >> 3] ^= (uint64_t)v << (8 * (at & 7)); }
__device__ __forceinline__ void strobe_run_f(strobe_t &s) {
    st_xor_byte(s, s.pos, s.pos_begin);
    st_xor_byte(s, s.pos + 1, 0x04);
    st_xor_byte(s, STROBE_R + 1, 0x80);
    keccak_f1600(s.st);
    s.pos = 0;
    s.pos_begin = 0;
}
__device__ __forceinline__ void strobe_absorb_byte(strobe_t &s, uint32_t v) {
    st_xor_byte(s, s.pos, v);
    if (++s.pos == STROBE_R) strobe_run_f(s);
}
__device__ void strobe_absorb(strobe_t &s, const uint8_t *d, uint32_t n) {
    uint32_t i = 0;
    // whole 64-bit words: the source is read with aligned 8-byte loads and re-aligned with a funnel shift, the state is updated at its
    // (generally unaligned) byte position through the two lanes the word touches; bytes only around a permutation and for the tail
    if (n >= 24) {
        const uint32_t m = (uint32_t)((uintptr_t)d & 7);
        const uint64_t *base = reinterpret_cast<const uint64_t *>(d - m);
        uint64_t w0 = *base++;
#pragma unroll 1
        while (n - i >= 16) {  // keeps the look-ahead load inside the message
            uint64_t v = w0;
            if (m) {
                const uint64_t w1 = *base;
                v = (w0 >> (8 * m)) | (w1 << (64 - 8 * m));
                w0 = w1;
            } else {
                w0 = *base;
            }
            base++;
            if (STROBE_R - s.pos >= 8) {
                const uint32_t sh = 8 * (s.pos & 7);
                s.st[s.pos >> 3] ^= v << sh;
                if (sh) s.st[(s.pos >> 3) + 1] ^= v >> (64 - sh);
                s.pos += 8;
                if (s.pos == STROBE_R) strobe_run_f(s);
            } else {  // this word straddles a permutation
#pragma unroll 1
                for (int k = 0; k < 8; k++) strobe_absorb_byte(s, (uint32_t)(v >> (8 * k)) & 0xFF);
            }
            i += 8;
        }
    }
    while (i < n) strobe_absorb_byte(s, d[i++]);
}
__device__ __forceinline__ void strobe_begin(strobe_t &s, uint32_t flags, bool more) {
    if (more) return;
    const uint32_t old_begin = s.pos_begin;
    s.pos_begin = s.pos + 1;
    strobe_absorb_byte(s, old_begin);
    strobe_absorb_byte(s, flags);
    if ((flags & FLAG_C) && s.pos != 0) strobe_run_f(s);
}
__device__ void strobe_meta_ad(strobe_t &s, const uint8_t *d, uint32_t n, bool more) {
    strobe_begin(s, FLAG_M | FLAG_A, more);
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) strobe_absorb_byte(s, d[i]);
}
__device__ void strobe_prf(strobe_t &s, uint8_t *out, uint32_t n) {
    strobe_begin(s, FLAG_I | FLAG_A | FLAG_C, false);
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        out[i] = (uint8_t)(s.st[s.pos >> 3] >> (8 * (s.pos & 7)));
        s.st[s.pos >> 3] &= ~((uint64_t)0xFF << (8 * (s.pos & 7)));
        if (++s.pos == STROBE_R) strobe_run_f(s);
    }
}
// little-endian bytes of a value, absorbed without going through a byte array in local memory
__device__ __forceinline__ void strobe_absorb_value(strobe_t &s, uint64_t v, uint32_t nbytes) {
#pragma unroll 1
    for (uint32_t i = 0; i < nbytes; i++) strobe_absorb_byte(s, (uint32_t)(v >> (8 * i)) & 0xFF);
}
// merlin: append_message(label, msg) = meta-AD(label) meta-AD(u32 len, more) AD(msg); here the message is an optional little-endian
// prefix value (the u64 element count of a serialised Vec) followed by a body
__device__ void merlin_append(strobe_t &s, const uint8_t *label, uint32_t llen, uint64_t prefix, uint32_t plen, const uint8_t *body,
                              uint32_t blen) {
    strobe_meta_ad(s, label, llen, false);
    strobe_absorb_value(s, plen + blen, 4);  // meta-AD continued (more = true): no new operation header
    strobe_begin(s, FLAG_A, false);
    strobe_absorb_value(s, prefix, plen);
    strobe_absorb(s, body, blen);
}
__device__ void merlin_challenge(strobe_t &s, const uint8_t *label, uint32_t llen, uint8_t *out, uint32_t n) {
    strobe_meta_ad(s, label, llen, false);
    strobe_absorb_value(s, n, 4);
    strobe_prf(s, out, n);
}

__device__ const uint8_t L_MERLIN[] = "Merlin v1.0";
__device__ const uint8_t L_DOMSEP[] = "dom-sep";
__device__ const uint8_t L_PROTO[] = "curdleproofs";
__device__ const uint8_t L_STEP1[] = "curdleproofs_step1";
__device__ const uint8_t L_VEC_A[] = "curdleproofs_vec_a";
__device__ const uint8_t L_STROBE[] = "STROBEv1.0.2";

}  // namespace

// comp_vecs: B x 4 x ell encodings (proof-major, R | S | T | U), comp_M: B encodings.  vec_a_out: B x ell canonical 32-byte scalars.
// state_out: B x 26 u64 = the 25 STROBE lanes, then pos | pos_begin << 8.
__global__ void __launch_bounds__(32) k_transcript_open(const uint8_t *__restrict__ comp_vecs, const uint8_t *__restrict__ comp_M, uint32_t ell,
                                                        uint32_t B, uint8_t *__restrict__ vec_a_out, uint64_t *__restrict__ state_out) {
    const uint32_t pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (pr >= B) return;
    strobe_t s;
    // Strobe128::new("Merlin v1.0"), then Transcript::new(b"curdleproofs") = append_message(b"dom-sep", label)
#pragma unroll 1
    for (int i = 0; i < 25; i++) s.st[i] = 0;
    {
        const uint8_t init[6] = {1, (uint8_t)(STROBE_R + 2), 1, 0, 1, 96};
        for (uint32_t i = 0; i < 6; i++) st_xor_byte(s, i, init[i]);
        for (uint32_t i = 0; i < 12; i++) st_xor_byte(s, 6 + i, L_STROBE[i]);
    }
    keccak_f1600(s.st);
    s.pos = 0;
    s.pos_begin = 0;
    strobe_meta_ad(s, L_MERLIN, 11, false);
    merlin_append(s, L_DOMSEP, 7, 0, 0, L_PROTO, 12);
    // append_list: every Vec<G1Affine> is one message, u64-LE length then the elements (ark-serialize)
#pragma unroll 1
    for (int v = 0; v < 4; v++) merlin_append(s, L_STEP1, 18, (uint64_t)ell, 8, comp_vecs + ((size_t)pr * 4 + v) * ell * 48, ell * 48);
    merlin_append(s, L_STEP1, 18, 0, 0, comp_M + (size_t)pr * 48, 48);
    // get_and_append_challenge (src/transcript.rs:41-54): 64 challenge bytes, the first 32 with bit 255 cleared must be a non-zero
    // value < r, otherwise draw again; the accepted challenge is appended
#pragma unroll 1
    for (uint32_t i = 0; i < ell; i++) {
        uint8_t buf[64];
        for (;;) {
            merlin_challenge(s, L_VEC_A, 18, buf, 64);
            buf[31] &= 0x7F;
            uint32_t w[8];
            uint32_t nz = 0;
            for (int k = 0; k < 8; k++) {
                w[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) | ((uint32_t)buf[4 * k + 3] << 24);
                nz |= w[k];
            }
            bool lt = false, decided = false;
            for (int k = 7; k >= 0; k--) {
                if (!decided && w[k] != FR_R[k]) {
                    lt = w[k] < FR_R[k];
                    decided = true;
                }
            }
            if (lt && nz) break;
        }
        merlin_append(s, L_VEC_A, 18, 0, 0, buf, 32);
        uint8_t *o = vec_a_out + ((size_t)pr * ell + i) * 32;
        for (int k = 0; k < 32; k++) o[k] = buf[k];
    }
    uint64_t *so = state_out + (size_t)pr * 26;
    for (int i = 0; i < 25; i++) so[i] = s.st[i];
    so[25] = (uint64_t)s.pos | ((uint64_t)s.pos_begin << 8);
}


// ---------------------------------------------------------------------------------------------------------------------------------
// The REST of the verifier's transcript on the device (verifier side): `CurdleproofsProof::verify`
// (/root/reference/src/curdleproofs.rs:226-296) reaches SamePerm / GrandProduct / IPA / SameScalar / SameMSM `verify`, each of which only
// appends proof points and scalars and draws challenges.  One thread per proof continues the STROBE state k_transcript_open left, in two
// kernels because the transcript needs two points computed on the GPU in between (D and A', a fixed-base launch):
//   k_verify_transcript_a   same_perm (src/same_permutation_argument.rs:134-145) and gprod (src/grand_product_argument.rs:202-223) up to
//                           the gprod beta: alpha, beta of same_perm, the grand product of (a_i + i alpha + beta), alpha_g, beta_g and
//                           beta_g^-1; writes the six scalars of D = B - beta^-1 sum(G) + alpha_g sum(Hvec) and A' = A + T_1 + U_1
//   k_verify_transcript_b   z, then IPA (src/inner_product_argument.rs:282-304), SameScalar (src/same_scalar_argument.rs:110-125) and
//                           SameMSM (src/same_multiscalar_argument.rs:231-241); inverts the 2m round challenges with one inversion each
//                           (the reference's batch_inversion, inner_product_argument.rs:234); completes the proof's challenge block for
//                           k_verify_coeffs (entries 12..26 and the four challenge vectors; 0..11, the random factors, come from the host)
// Proof points are read as the 48-byte encodings of the serialised proof, proof scalars as canonical 32-byte values.
namespace {

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

__device__ __forceinline__ void state_load(strobe_t &s, const uint64_t *src) {
#pragma unroll 1
    for (int i = 0; i < 25; i++) s.st[i] = src[i];
    s.pos = (uint32_t)(src[25] & 0xFF);
    s.pos_begin = (uint32_t)((src[25] >> 8) & 0xFF);
}
__device__ __forceinline__ void state_store(uint64_t *dst, const strobe_t &s) {
#pragma unroll 1
    for (int i = 0; i < 25; i++) dst[i] = s.st[i];
    dst[25] = (uint64_t)s.pos | ((uint64_t)s.pos_begin << 8);
}
__device__ __forceinline__ void append_point(strobe_t &s, const uint8_t *label, uint32_t llen, const uint8_t *comp) {
    merlin_append(s, label, llen, 0, 0, comp, 48);
}
// `append(label, &Fr)`: the canonical 32-byte little-endian value (src/transcript.rs:29-33)
__device__ void append_fr(strobe_t &s, const uint8_t *label, uint32_t llen, const vcoef::fr_t &x_mont) {
    uint32_t w[8];
    vcoef::fr_store_canonical(w, x_mont);
    strobe_meta_ad(s, label, llen, false);
    strobe_absorb_value(s, 32, 4);
    strobe_begin(s, FLAG_A, false);
#pragma unroll 1
    for (int k = 0; k < 8; k++) strobe_absorb_value(s, w[k], 4);
}
// get_and_append_challenge (src/transcript.rs:41-54); returns the challenge in Montgomery form
__device__ vcoef::fr_t draw_challenge(strobe_t &s, const uint8_t *label, uint32_t llen) {
    uint8_t buf[64];
    vcoef::fr_t c;
    for (;;) {
        merlin_challenge(s, label, llen, buf, 64);
        buf[31] &= 0x7F;
        uint32_t nz = 0;
        for (int k = 0; k < 8; k++) {
            c.v[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) | ((uint32_t)buf[4 * k + 3] << 24);
            nz |= c.v[k];
        }
        if (nz && !vcoef::fr_geq_mod(c.v)) break;
    }
    merlin_append(s, label, llen, 0, 0, buf, 32);
    return vcoef::fr_to_mont(c);
}
__device__ __forceinline__ void chal_store(uint32_t *dst, const vcoef::fr_t &x) {
    for (int k = 0; k < 8; k++) dst[k] = x.v[k];
}
// the thread that writes a proof's results: with one thread per proof, that thread
__device__ __forceinline__ bool is_leader(const strobe_t &) { return true; }

}  // namespace

// pcomp: B x np encodings (the proof's points in serialisation order), pscal: B x 7 canonical scalars (r_p, c_final, d_final, z_k, z_t,
// z_u, x_final), comp_M: B encodings, vec_a: B x ell canonical, state: B x 26 u64 (in/out).  Outputs: chal[pr][12..15] (Montgomery),
// tmp[pr][0..1] = grand product, beta_g (Montgomery), stage_scal[pr][6] canonical = {1, -beta_g^-1, alpha_g, 1, 1, 1}, flags[pr] = 1 when
// vec_T[0] (first encoding of the T block of comp_vecs, B x 4 x ell encodings) is the identity.
template <class S>
__device__ __forceinline__ void verify_transcript_a_body(S &s, uint32_t pr, const uint8_t *__restrict__ pcomp, const uint8_t *__restrict__ pscal,
                                                         const uint8_t *__restrict__ comp_M, const uint8_t *__restrict__ vec_a, uint32_t ell, uint32_t np,
                                                         uint32_t vch, uint64_t *__restrict__ state, uint32_t *__restrict__ chal, uint32_t *__restrict__ tmp,
                                                         uint32_t *__restrict__ stage_scal, const uint8_t *__restrict__ comp_vecs, uint8_t *__restrict__ flags) {
    using namespace vcoef;
    state_load(s, state + (size_t)pr * 26);
    const uint8_t *pc = pcomp + (size_t)pr * np * 48;
    const uint8_t *va = vec_a + (size_t)pr * ell * 32;
    // same_perm: A, M, vec_a -> alpha, beta
    append_point(s, L_SP1, 15, pc + 48 * 0);
    append_point(s, L_SP1, 15, comp_M + (size_t)pr * 48);
    merlin_append(s, L_SP1, 15, (uint64_t)ell, 8, va, ell * 32);
    const fr_t alpha_sp = draw_challenge(s, L_SPA, 15), beta_sp = draw_challenge(s, L_SPB, 14);
    fr_t gprod = fr_one(), i_alpha = beta_sp;  // i * alpha + beta, advanced by addition
#pragma unroll 1
    for (uint32_t i = 0; i < ell; i++) {
        const fr_t a = fr_to_mont(fr_load(reinterpret_cast<const uint32_t *>(va + 32 * i)));
        gprod = fr_mul(gprod, fr_add(a, i_alpha));
        i_alpha = fr_add(i_alpha, alpha_sp);
    }
    // gprod: B, grand product -> alpha; C, r_p -> beta
    append_point(s, L_GP1, 11, pc + 48 * 7);
    append_fr(s, L_GP1, 11, gprod);
    const fr_t alpha_g = draw_challenge(s, L_GPA, 11);
    append_point(s, L_GP2, 11, pc + 48 * 8);
    const fr_t r_p = fr_to_mont(fr_load(reinterpret_cast<const uint32_t *>(pscal + ((size_t)pr * 7 + 0) * 32)));
    append_fr(s, L_GP2, 11, r_p);
    const fr_t beta_g = draw_challenge(s, L_GPB, 10);
    const fr_t beta_inv = fr_inverse_safegcd(beta_g);
    state_store(state + (size_t)pr * 26, s);
    if (!is_leader(s)) return;
    uint32_t *ch = chal + 8 * (size_t)pr * vch;
    chal_store(ch + 8 * CH_ALPHA_SP, alpha_sp); chal_store(ch + 8 * CH_BETA_SP, beta_sp);
    chal_store(ch + 8 * CH_ALPHA_G, alpha_g); chal_store(ch + 8 * CH_BETA_INV, beta_inv);
    chal_store(tmp + 8 * ((size_t)pr * 2 + 0), gprod); chal_store(tmp + 8 * ((size_t)pr * 2 + 1), beta_g);
    uint32_t *sc = stage_scal + 8 * (size_t)pr * 6;
    const fr_t one = fr_one();
    fr_store_canonical(sc + 0, one); fr_store_canonical(sc + 8, fr_neg(beta_inv)); fr_store_canonical(sc + 16, alpha_g);
    fr_store_canonical(sc + 24, one); fr_store_canonical(sc + 32, one); fr_store_canonical(sc + 40, one);
    // vec_T[0] is the identity -> Err(VerificationError) (curdleproofs.rs:218-220): bit 6 of the first byte of its encoding
    if (flags) flags[pr] = (comp_vecs[(((size_t)pr * 4 + 2) * ell) * 48] & 0x40) ? 1 : 0;
}

__global__ void __launch_bounds__(32) k_verify_transcript_a(const uint8_t *__restrict__ pcomp, const uint8_t *__restrict__ pscal,
                                                            const uint8_t *__restrict__ comp_M, const uint8_t *__restrict__ vec_a, uint32_t ell,
                                                            uint32_t np, uint32_t vch, uint32_t B, uint64_t *__restrict__ state,
                                                            uint32_t *__restrict__ chal, uint32_t *__restrict__ tmp, uint32_t *__restrict__ stage_scal,
                                                            const uint8_t *__restrict__ comp_vecs, uint8_t *__restrict__ flags) {
    const uint32_t pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (pr >= B) return;
    strobe_t s;
    verify_transcript_a_body(s, pr, pcomp, pscal, comp_M, vec_a, ell, np, vch, state, chal, tmp, stage_scal, comp_vecs, flags);
}

// da_comp: B x 2 encodings (D, A'), comp_T / comp_U: the instance's vec_T / vec_U encodings of proof pr at comp_vecs + ((pr * 4 + 2 | 3) * ell) * 48,
// H_comp: the encoding of crs.H.  Completes chal[pr][16..26] and the four challenge vectors.
template <class S>
__device__ __forceinline__ void verify_transcript_b_body(S &s, uint32_t pr, const uint8_t *__restrict__ pcomp, const uint8_t *__restrict__ pscal,
                                                         const uint8_t *__restrict__ da_comp, const uint8_t *__restrict__ comp_vecs,
                                                         const uint8_t *__restrict__ H_comp, uint32_t ell, uint32_t m, uint32_t np, uint32_t vch,
                                                         uint64_t *__restrict__ state, uint32_t *__restrict__ chal, const uint32_t *__restrict__ tmp) {
    using namespace vcoef;
    state_load(s, state + (size_t)pr * 26);
    const uint8_t *pc = pcomp + (size_t)pr * np * 48;
    const uint32_t n = ell + 4;
    // proof point indices in serialisation order (host/verifier.cpp ProofLayout)
    const uint32_t L_T1 = 1, L_T2 = 2, L_U1 = 3, L_U2 = 4, L_R = 5, L_S = 6, L_C = 8, L_Bc = 9, L_Bd = 10;
    const uint32_t L_LC = 11, L_RC = L_LC + m, L_LD = L_RC + m, L_RD = L_LD + m, L_A1 = L_RD + m, L_A2 = L_A1 + 1, L_B1 = L_A2 + 1,
                   L_B2 = L_B1 + 1, L_Ba = L_B2 + 1, L_Bt = L_Ba + 1, L_Bu = L_Bt + 1, L_LA = L_Bu + 1, L_LT = L_LA + m, L_LU = L_LT + m,
                   L_RA = L_LU + m, L_RT = L_RA + m, L_RU = L_RT + m;
    auto PS = [&](uint32_t k) { return fr_to_mont(fr_load(reinterpret_cast<const uint32_t *>(pscal + ((size_t)pr * 7 + k) * 32))); };
    const fr_t r_p = PS(0);
    const fr_t gprod = fr_load(tmp + 8 * ((size_t)pr * 2 + 0)), beta = fr_load(tmp + 8 * ((size_t)pr * 2 + 1));
    // z = r_p beta^(ell+1) + gprod beta^ell - 1   (grand_product_argument.rs:225-229)
    const fr_t beta_l = fr_pow_u32(beta, ell);
    const fr_t z = fr_sub(fr_add(fr_mul(r_p, fr_mul(beta_l, beta)), fr_mul(gprod, beta_l)), fr_one());
    // IPA
    append_point(s, L_IP1, 9, pc + 48 * L_C);
    append_point(s, L_IP1, 9, da_comp + ((size_t)pr * 2 + 0) * 48);
    append_fr(s, L_IP1, 9, z);
    append_point(s, L_IP1, 9, pc + 48 * L_Bc);
    append_point(s, L_IP1, 9, pc + 48 * L_Bd);
    const fr_t alpha_i = draw_challenge(s, L_IPA, 9), beta_i = draw_challenge(s, L_IPB, 8);
    uint32_t *ch = chal + 8 * (size_t)pr * vch;
    fr_t g[16];
#pragma unroll 1
    for (uint32_t k = 0; k < m; k++) {
        append_point(s, L_IPL, 8, pc + 48 * (L_LC + k)); append_point(s, L_IPL, 8, pc + 48 * (L_LD + k));
        append_point(s, L_IPL, 8, pc + 48 * (L_RC + k)); append_point(s, L_IPL, 8, pc + 48 * (L_RD + k));
        g[k] = draw_challenge(s, L_IPG, 9);
        if (is_leader(s)) chal_store(ch + 8 * (CH_VEC + k), g[k]);
    }
    fr_batch_inverse(g, m);
    if (is_leader(s)) for (uint32_t k = 0; k < m; k++) chal_store(ch + 8 * (CH_VEC + m + k), g[k]);
    // SameScalar
    {
        const uint32_t ss[10] = {L_R, L_S, L_T1, L_T2, L_U1, L_U2, L_A1, L_A2, L_B1, L_B2};
#pragma unroll 1
        for (int q = 0; q < 10; q++) append_point(s, L_SEP, 14, pc + 48 * ss[q]);
    }
    const fr_t alpha_ss = draw_challenge(s, L_SSA, 17);
    // SameMSM: A', cm_T.T_2, cm_U.T_2, vec_T | inf inf H inf, vec_U | inf inf inf H, B_a, B_t, B_u
    append_point(s, L_SM1, 14, da_comp + ((size_t)pr * 2 + 1) * 48);
    append_point(s, L_SM1, 14, pc + 48 * L_T2);
    append_point(s, L_SM1, 14, pc + 48 * L_U2);
#pragma unroll 1
    for (int v = 0; v < 2; v++) {  // a Vec<G1Affine> of n elements: u64-LE count, the ell instance points, the four blinder slots
        strobe_meta_ad(s, L_SM1, 14, false);
        strobe_absorb_value(s, 8 + n * 48, 4);
        strobe_begin(s, FLAG_A, false);
        strobe_absorb_value(s, (uint64_t)n, 8);
        strobe_absorb(s, comp_vecs + (((size_t)pr * 4 + 2 + v) * ell) * 48, ell * 48);
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            if (q == 2 + v) {
                strobe_absorb(s, H_comp, 48);
            } else {
                strobe_absorb_byte(s, 0xC0);
#pragma unroll 1
                for (int k = 1; k < 48; k++) strobe_absorb_byte(s, 0);
            }
        }
    }
    append_point(s, L_SM1, 14, pc + 48 * L_Ba);
    append_point(s, L_SM1, 14, pc + 48 * L_Bt);
    append_point(s, L_SM1, 14, pc + 48 * L_Bu);
    const fr_t alpha_sm = draw_challenge(s, L_SMA, 14);
#pragma unroll 1
    for (uint32_t k = 0; k < m; k++) {
        const uint32_t o[6] = {L_LA, L_LT, L_LU, L_RA, L_RT, L_RU};
#pragma unroll 1
        for (int q = 0; q < 6; q++) append_point(s, L_SML, 13, pc + 48 * (o[q] + k));
        g[k] = draw_challenge(s, L_SMG, 14);
        if (is_leader(s)) chal_store(ch + 8 * (CH_VEC + 2 * m + k), g[k]);
    }
    fr_batch_inverse(g, m);
    if (is_leader(s)) for (uint32_t k = 0; k < m; k++) chal_store(ch + 8 * (CH_VEC + 3 * m + k), g[k]);
    state_store(state + (size_t)pr * 26, s);
    if (!is_leader(s)) return;
    chal_store(ch + 8 * CH_ALPHA_I, alpha_i); chal_store(ch + 8 * CH_BETA_I, beta_i); chal_store(ch + 8 * CH_Z, z);
    chal_store(ch + 8 * CH_C, PS(1)); chal_store(ch + 8 * CH_D, PS(2)); chal_store(ch + 8 * CH_X, PS(6));
    chal_store(ch + 8 * CH_ALPHA_SM, alpha_sm); chal_store(ch + 8 * CH_ALPHA_SS, alpha_ss);
    chal_store(ch + 8 * CH_ZK, PS(3)); chal_store(ch + 8 * CH_ZT, PS(4)); chal_store(ch + 8 * CH_ZU, PS(5));
}

__global__ void __launch_bounds__(32) k_verify_transcript_b(const uint8_t *__restrict__ pcomp, const uint8_t *__restrict__ pscal,
                                                            const uint8_t *__restrict__ da_comp, const uint8_t *__restrict__ comp_vecs,
                                                            const uint8_t *__restrict__ H_comp, uint32_t ell, uint32_t m, uint32_t np, uint32_t vch,
                                                            uint32_t B, uint64_t *__restrict__ state, uint32_t *__restrict__ chal,
                                                            const uint32_t *__restrict__ tmp) {
    const uint32_t pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (pr >= B) return;
    strobe_t s;
    verify_transcript_b_body(s, pr, pcomp, pscal, da_comp, comp_vecs, H_comp, ell, m, np, vch, state, chal, tmp);
}

#ifndef CDP_TRANSCRIPT_HOST_HARNESS
// ---- the same opening with ONE WARP per proof (cstrobe.cuh): the STROBE state in shared memory, message bytes XORed in by all lanes,
// Keccak-f[1600] by 25 lanes.  One thread per proof (k_transcript_open above, kept: it is the form the CPU harness checks, and the GPU test
// compares this kernel's output with it byte for byte) spends ~8 ms per launch on ~640 dependent permutations of ~9 us each; here a
// permutation is ~2 us.
#define CTA_FOR(i, cnt) for (uint32_t i = threadIdx.x; i < (uint32_t)(cnt); i += blockDim.x)
#define CTA_SYNC() __syncwarp()
#define CTA_LEADER (threadIdx.x == 0)
}  // namespace cdp
#include "cstrobe.cuh"
namespace cdp {
__global__ void __launch_bounds__(32) k_transcript_open_warp(const uint8_t *__restrict__ comp_vecs, const uint8_t *__restrict__ comp_M, uint32_t ell,
                                                             uint32_t B, uint8_t *__restrict__ vec_a_out, uint64_t *__restrict__ state_out) {
    using namespace cstr;
    __shared__ __align__(8) uint8_t st[200];
    __shared__ __align__(8) uint8_t buf[64];
    const uint32_t pr = blockIdx.x, lane = threadIdx.x;
    cstrobe s{st, 0, 0};
    for (uint32_t i = lane; i < 200; i += 32) st[i] = 0;
    __syncwarp();
    if (lane == 0) {  // Strobe128::new("Merlin v1.0")
        const uint8_t init[6] = {1, (uint8_t)(SR + 2), 1, 0, 1, 96};
        for (uint32_t i = 0; i < 6; i++) st[i] = init[i];
        for (uint32_t i = 0; i < 12; i++) st[6 + i] = L_STROBE[i];
    }
    keccak_f1600_warp(reinterpret_cast<uint64_t *>(st));
    cs_begin(s, cstr::FLAG_M | cstr::FLAG_A);
    cs_absorb(s, L_MERLIN, 11);
    cs_append(s, L_DOMSEP, 7, L_PROTO, 12);  // Transcript::new(b"curdleproofs")
    // append_list: every Vec<G1Affine> is one message, u64-LE length then the elements (ark-serialize)
#pragma unroll 1
    for (int v = 0; v < 4; v++) {
        cs_append_header(s, L_STEP1, 18, 8 + ell * 48);
        cs_value(s, (uint64_t)ell, 8);
        cs_absorb(s, comp_vecs + ((size_t)pr * 4 + v) * ell * 48, ell * 48);
    }
    cs_append(s, L_STEP1, 18, comp_M + (size_t)pr * 48, 48);
    // get_and_append_challenge (src/transcript.rs:41-54), as in k_transcript_open
#pragma unroll 1
    for (uint32_t i = 0; i < ell; i++) {
        for (;;) {
            cs_challenge_bytes(s, L_VEC_A, 18, buf, 64);
            uint32_t w[8];
            uint32_t nz = 0;
            for (int k = 0; k < 8; k++) {
                w[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) | ((uint32_t)buf[4 * k + 3] << 24);
                if (k == 7) w[k] &= 0x7FFFFFFFu;
                nz |= w[k];
            }
            bool lt = false, decided = false;
            for (int k = 7; k >= 0; k--) {
                if (!decided && w[k] != FR_R[k]) {
                    lt = w[k] < FR_R[k];
                    decided = true;
                }
            }
            __syncwarp();  // every lane has read the bytes before another draw may overwrite them
            if (lt && nz) break;
        }
        if (lane == 0) buf[31] &= 0x7F;
        __syncwarp();
        cs_append(s, L_VEC_A, 18, buf, 32);
        vec_a_out[((size_t)pr * ell + i) * 32 + lane] = buf[lane];
        __syncwarp();
    }
    uint64_t *so = state_out + (size_t)pr * 26;
    if (lane < 25) so[lane] = reinterpret_cast<const uint64_t *>(st)[lane];
    if (lane == 0) so[25] = (uint64_t)s.pos | ((uint64_t)s.pos_begin << 8);
}
// ---- the verifier's transcript kernels with one warp per proof: the same bodies (verify_transcript_a_body / _b_body) over a STROBE state in
// shared memory.  Every lane runs the scalar algebra redundantly (same latency as one thread), lane 0 writes the results.
namespace {
struct wstrobe {
    cstr::cstrobe c;
    uint8_t *buf;  // 64 shared bytes
};
__device__ __forceinline__ bool is_leader(const wstrobe &) { return threadIdx.x == 0; }
__device__ __forceinline__ void state_load(wstrobe &s, const uint64_t *src) {
    if (threadIdx.x < 25) reinterpret_cast<uint64_t *>(s.c.st)[threadIdx.x] = src[threadIdx.x];
    s.c.pos = (uint32_t)(src[25] & 0xFF);
    s.c.pos_begin = (uint32_t)((src[25] >> 8) & 0xFF);
    __syncwarp();
}
__device__ __forceinline__ void state_store(uint64_t *dst, const wstrobe &s) {
    __syncwarp();
    if (threadIdx.x < 25) dst[threadIdx.x] = reinterpret_cast<const uint64_t *>(s.c.st)[threadIdx.x];
    if (threadIdx.x == 0) dst[25] = (uint64_t)s.c.pos | ((uint64_t)s.c.pos_begin << 8);
}
__device__ __forceinline__ void strobe_begin(wstrobe &s, uint32_t flags, bool more) {
    if (!more) cstr::cs_begin(s.c, flags);
}
__device__ __forceinline__ void strobe_meta_ad(wstrobe &s, const uint8_t *d, uint32_t n, bool more) {
    if (!more) cstr::cs_begin(s.c, cstr::FLAG_M | cstr::FLAG_A);
    cstr::cs_absorb(s.c, d, n);
}
__device__ __forceinline__ void strobe_absorb_value(wstrobe &s, uint64_t v, uint32_t nbytes) { cstr::cs_value(s.c, v, nbytes); }
__device__ __forceinline__ void strobe_absorb(wstrobe &s, const uint8_t *d, uint32_t n) { cstr::cs_absorb(s.c, d, n); }
__device__ __forceinline__ void strobe_absorb_byte(wstrobe &s, uint32_t v) { cstr::cs_byte(s.c, v); }
__device__ __forceinline__ void merlin_append(wstrobe &s, const uint8_t *label, uint32_t llen, uint64_t prefix, uint32_t plen, const uint8_t *body,
                                              uint32_t blen) {
    cstr::cs_append_header(s.c, label, llen, plen + blen);
    cstr::cs_value(s.c, prefix, plen);
    cstr::cs_absorb(s.c, body, blen);
}
__device__ __forceinline__ void append_point(wstrobe &s, const uint8_t *label, uint32_t llen, const uint8_t *comp) {
    cstr::cs_append(s.c, label, llen, comp, 48);
}
__device__ void append_fr(wstrobe &s, const uint8_t *label, uint32_t llen, const vcoef::fr_t &x_mont) {
    if (threadIdx.x == 0) vcoef::fr_store_canonical(reinterpret_cast<uint32_t *>(s.buf), x_mont);
    __syncwarp();
    cstr::cs_append(s.c, label, llen, s.buf, 32);
    __syncwarp();
}
__device__ vcoef::fr_t draw_challenge(wstrobe &s, const uint8_t *label, uint32_t llen) {
    vcoef::fr_t c;
    for (;;) {
        cstr::cs_challenge_bytes(s.c, label, llen, s.buf, 64);
        uint32_t nz = 0;
        for (int k = 0; k < 8; k++) {
            c.v[k] = reinterpret_cast<const uint32_t *>(s.buf)[k];
            if (k == 7) c.v[k] &= 0x7FFFFFFFu;
            nz |= c.v[k];
        }
        const bool ok = nz && !vcoef::fr_geq_mod(c.v);
        __syncwarp();  // every lane has read the bytes before another draw may overwrite them
        if (ok) break;
    }
    if (threadIdx.x == 0) s.buf[31] &= 0x7F;
    __syncwarp();
    cstr::cs_append(s.c, label, llen, s.buf, 32);
    __syncwarp();
    return vcoef::fr_to_mont(c);
}
}  // namespace
__global__ void __launch_bounds__(32) k_verify_transcript_a_warp(const uint8_t *__restrict__ pcomp, const uint8_t *__restrict__ pscal,
                                                                 const uint8_t *__restrict__ comp_M, const uint8_t *__restrict__ vec_a, uint32_t ell,
                                                                 uint32_t np, uint32_t vch, uint32_t B, uint64_t *__restrict__ state,
                                                                 uint32_t *__restrict__ chal, uint32_t *__restrict__ tmp,
                                                                 uint32_t *__restrict__ stage_scal, const uint8_t *__restrict__ comp_vecs,
                                                                 uint8_t *__restrict__ flags) {
    __shared__ __align__(8) uint8_t st[200];
    __shared__ __align__(8) uint8_t buf[64];
    wstrobe s{{st, 0, 0}, buf};
    verify_transcript_a_body(s, blockIdx.x, pcomp, pscal, comp_M, vec_a, ell, np, vch, state, chal, tmp, stage_scal, comp_vecs, flags);
}
__global__ void __launch_bounds__(32) k_verify_transcript_b_warp(const uint8_t *__restrict__ pcomp, const uint8_t *__restrict__ pscal,
                                                                 const uint8_t *__restrict__ da_comp, const uint8_t *__restrict__ comp_vecs,
                                                                 const uint8_t *__restrict__ H_comp, uint32_t ell, uint32_t m, uint32_t np, uint32_t vch,
                                                                 uint32_t B, uint64_t *__restrict__ state, uint32_t *__restrict__ chal,
                                                                 const uint32_t *__restrict__ tmp) {
    __shared__ __align__(8) uint8_t st[200];
    __shared__ __align__(8) uint8_t buf[64];
    wstrobe s{{st, 0, 0}, buf};
    verify_transcript_b_body(s, blockIdx.x, pcomp, pscal, da_comp, comp_vecs, H_comp, ell, m, np, vch, state, chal, tmp);
}
#undef CTA_FOR
#undef CTA_SYNC
#undef CTA_LEADER
#endif

#ifndef CDP_TRANSCRIPT_HOST_HARNESS
// one warp per proof (default) or the one-thread-per-proof kernels (CDP_TRANSCRIPT_WARP=0)
static bool transcript_warp() {
    static const bool warp = [] { const char *e = getenv("CDP_TRANSCRIPT_WARP"); return !e || atoi(e) != 0; }();
    return warp;
}
cudaError_t launch_transcript_open(cudaStream_t st, const uint8_t *comp_vecs, const uint8_t *comp_M, uint32_t ell, uint32_t B, uint8_t *vec_a_out,
                                   uint64_t *state_out) {
    if (B == 0) return cudaSuccess;
    if (transcript_warp()) k_transcript_open_warp<<<B, 32, 0, st>>>(comp_vecs, comp_M, ell, B, vec_a_out, state_out);
    else k_transcript_open<<<(B + 31) / 32, 32, 0, st>>>(comp_vecs, comp_M, ell, B, vec_a_out, state_out);
    return cudaGetLastError();
}
cudaError_t launch_verify_transcript_a(cudaStream_t st, const uint8_t *pcomp, const uint8_t *pscal, const uint8_t *comp_M, const uint8_t *vec_a,
                                       uint32_t ell, uint32_t np, uint32_t vch, uint32_t B, uint64_t *state, uint32_t *chal, uint32_t *tmp,
                                       uint32_t *stage_scal, const uint8_t *comp_vecs, uint8_t *flags) {
    if (B == 0) return cudaSuccess;
    if (transcript_warp()) k_verify_transcript_a_warp<<<B, 32, 0, st>>>(pcomp, pscal, comp_M, vec_a, ell, np, vch, B, state, chal, tmp, stage_scal, comp_vecs, flags);
    else k_verify_transcript_a<<<(B + 31) / 32, 32, 0, st>>>(pcomp, pscal, comp_M, vec_a, ell, np, vch, B, state, chal, tmp, stage_scal, comp_vecs, flags);
    return cudaGetLastError();
}
cudaError_t launch_verify_transcript_b(cudaStream_t st, const uint8_t *pcomp, const uint8_t *pscal, const uint8_t *da_comp, const uint8_t *comp_vecs,
                                       const uint8_t *H_comp, uint32_t ell, uint32_t m, uint32_t np, uint32_t vch, uint32_t B, uint64_t *state,
                                       uint32_t *chal, const uint32_t *tmp) {
    if (B == 0) return cudaSuccess;
    if (transcript_warp()) k_verify_transcript_b_warp<<<B, 32, 0, st>>>(pcomp, pscal, da_comp, comp_vecs, H_comp, ell, m, np, vch, B, state, chal, tmp);
    else k_verify_transcript_b<<<(B + 31) / 32, 32, 0, st>>>(pcomp, pscal, da_comp, comp_vecs, H_comp, ell, m, np, vch, B, state, chal, tmp);
    return cudaGetLastError();
}
#endif

}  // namespace cdp
