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

#ifndef CDP_TRANSCRIPT_HOST_HARNESS
cudaError_t launch_transcript_open(cudaStream_t st, const uint8_t *comp_vecs, const uint8_t *comp_M, uint32_t ell, uint32_t B, uint8_t *vec_a_out,
                                   uint64_t *state_out) {
    if (B == 0) return cudaSuccess;
    k_transcript_open<<<(B + 31) / 32, 32, 0, st>>>(comp_vecs, comp_M, ell, B, vec_a_out, state_out);
    return cudaGetLastError();
}
#endif

}  // namespace cdp
