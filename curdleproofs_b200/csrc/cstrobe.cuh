// STROBE-128 / merlin over a state SHARED by a CTA, and Keccak-f[1600] spread over the lanes of a warp.
//
// The transcript steps of the device-side prover (k_prove.cu) and the transcript opening (k_transcript.cu) are latency chains of Keccak
// permutations (~900 per ell = 252 proof): one warp per proof, message bytes XORed into the shared state by all lanes, and the permutation
// itself run by 25 lanes, one 64-bit lane of the state each -- theta, rho-pi and chi become 4 + 2 + 1 + 2 shuffles per round, ~45
// instructions deep instead of ~180 for one thread holding all 25 lanes (~2 us instead of ~9 us per permutation for a lone warp).
// The includer defines CTA_FOR / CTA_SYNC / CTA_LEADER (k_prove.cu: device, or the CPU harness where a CTA is emulated sequentially and the
// permutation is the plain one-thread version).
#pragma once
#include <stdint.h>

namespace cdp {
namespace cstr {

__device__ __forceinline__ uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }

static __device__ __noinline__ void keccak_f1600(uint64_t *A) {
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

#if defined(__CUDACC__) && !defined(CDP_PROVE_HOST_HARNESS)
// Keccak-f[1600] by one warp: lane t < 25 owns A[t] = a[x][y], t = x + 5 y.  All 32 lanes must call (full-mask shuffles).
__device__ __forceinline__ void keccak_f1600_warp(uint64_t *st) {
    const uint64_t RC[24] = {0x1ULL, 0x8082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x808bULL, 0x80000001ULL,
                             0x8000000080008081ULL, 0x8000000000008009ULL, 0x8aULL, 0x88ULL, 0x80008009ULL, 0x8000000aULL,
                             0x8000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                             0x8000000000008002ULL, 0x8000000000000080ULL, 0x800aULL, 0x800000008000000aULL,
                             0x8000000080008081ULL, 0x8000000000008080ULL, 0x80000001ULL, 0x8000000080008008ULL};
    const int lane = (int)(threadIdx.x & 31), t = lane < 25 ? lane : 0, x = t % 5, y = t / 5;
    // rho offsets r[x][y] by lane index
    const uint8_t RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    const int rho = (int)RHO[t];
    const int th1 = x + 5 * ((y + 1) % 5), th2 = x + 5 * ((y + 2) % 5), th3 = x + 5 * ((y + 3) % 5), th4 = x + 5 * ((y + 4) % 5);
    const int dm = (x + 4) % 5 + 5 * y, dp = (x + 1) % 5 + 5 * y, c2 = (x + 2) % 5 + 5 * y;
    const int pi_src = (x + 3 * y) % 5 + 5 * x;  // B[X][Y] = rot(a[(X + 3Y) % 5][X])
    __syncwarp();
    uint64_t a = lane < 25 ? st[lane] : 0;  // the idle lanes only take part in the shuffles
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        const uint64_t c = a ^ __shfl_sync(0xffffffffu, a, th1) ^ __shfl_sync(0xffffffffu, a, th2) ^ __shfl_sync(0xffffffffu, a, th3) ^
                           __shfl_sync(0xffffffffu, a, th4);
        const uint64_t cp = __shfl_sync(0xffffffffu, c, dp);
        a ^= __shfl_sync(0xffffffffu, c, dm) ^ ((cp << 1) | (cp >> 63));
        uint64_t b = rho ? ((a << rho) | (a >> (64 - rho))) : a;
        b = __shfl_sync(0xffffffffu, b, pi_src);
        a = b ^ (~__shfl_sync(0xffffffffu, b, dp) & __shfl_sync(0xffffffffu, b, c2));
        if (lane == 0) a ^= RC[round];
    }
    if (lane < 25) st[lane] = a;
    __syncwarp();
}
#endif

// STROBE-128 (rate 166) restricted to what merlin uses (meta-AD, AD, PRF), with the state shared by the CTA.  `pos` / `pos_begin` depend only
// on the message lengths, so every thread tracks them in its own registers and all control flow below is uniform across the CTA.
struct cstrobe {
    uint8_t *st;  // 200 bytes, shared memory
    uint32_t pos, pos_begin;
};
constexpr uint32_t SR = 166;
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_M = 16 };

static __device__ void cs_run_f(cstrobe &s) {
    CTA_SYNC();  // all pending XORs into the state are visible
#if defined(__CUDACC__) && !defined(CDP_PROVE_HOST_HARNESS)
    if (threadIdx.x < 32) {  // the CTA's first warp: lane 0 closes the block, 25 lanes permute
        if (threadIdx.x == 0) {
            s.st[s.pos] ^= (uint8_t)s.pos_begin;
            s.st[s.pos + 1] ^= 0x04;
            s.st[SR + 1] ^= 0x80;
        }
        keccak_f1600_warp(reinterpret_cast<uint64_t *>(s.st));
    }
#else
    if (CTA_LEADER) {
        s.st[s.pos] ^= (uint8_t)s.pos_begin;
        s.st[s.pos + 1] ^= 0x04;
        s.st[SR + 1] ^= 0x80;
        keccak_f1600(reinterpret_cast<uint64_t *>(s.st));
    }
#endif
    CTA_SYNC();
    s.pos = 0;
    s.pos_begin = 0;
}
static __device__ void cs_byte(cstrobe &s, uint32_t v) {
    if (CTA_LEADER) s.st[s.pos] ^= (uint8_t)v;
    if (++s.pos == SR) cs_run_f(s);
}
// message bytes: thread t takes byte t of the part that fits before the next permutation
static __device__ void cs_absorb(cstrobe &s, const uint8_t *src, uint32_t n) {
    uint32_t done = 0;
    while (done < n) {
        const uint32_t room = SR - s.pos, chunk = n - done < room ? n - done : room;
        CTA_FOR(t, chunk) s.st[s.pos + t] ^= src[done + t];
        s.pos += chunk;
        done += chunk;
        if (s.pos == SR) cs_run_f(s);
    }
}
static __device__ void cs_value(cstrobe &s, uint64_t v, uint32_t nbytes) {
    for (uint32_t i = 0; i < nbytes; i++) cs_byte(s, (uint32_t)(v >> (8 * i)) & 0xFF);
}
static __device__ void cs_begin(cstrobe &s, uint32_t flags) {
    const uint32_t old_begin = s.pos_begin;
    s.pos_begin = s.pos + 1;
    cs_byte(s, old_begin);
    cs_byte(s, flags);
    if ((flags & FLAG_C) && s.pos != 0) cs_run_f(s);
}
// merlin append_message(label, msg): meta-AD(label), meta-AD(u32 length, continued), AD(msg); the message body follows through cs_value / cs_absorb
static __device__ void cs_append_header(cstrobe &s, const uint8_t *label, uint32_t llen, uint32_t msg_len) {
    cs_begin(s, FLAG_M | FLAG_A);
    cs_absorb(s, label, llen);
    cs_value(s, msg_len, 4);
    cs_begin(s, FLAG_A);
}
static __device__ void cs_append(cstrobe &s, const uint8_t *label, uint32_t llen, const uint8_t *body, uint32_t blen) {
    cs_append_header(s, label, llen, blen);
    cs_absorb(s, body, blen);
}
// merlin challenge_bytes(label, out): meta-AD(label), meta-AD(u32 length, continued), PRF(out).  `out` is CTA-shared; valid for all threads on return.
static __device__ void cs_challenge_bytes(cstrobe &s, const uint8_t *label, uint32_t llen, uint8_t *out, uint32_t n) {
    cs_begin(s, FLAG_M | FLAG_A);
    cs_absorb(s, label, llen);
    cs_value(s, n, 4);
    cs_begin(s, FLAG_I | FLAG_A | FLAG_C);
    // PRF: the output is the state's bytes, which are then cleared; thread t takes byte t of the part before the next permutation
    uint32_t done = 0;
    while (done < n) {
        const uint32_t room = SR - s.pos, chunk = n - done < room ? n - done : room;
        CTA_FOR(t, chunk) {
            out[done + t] = s.st[s.pos + t];
            s.st[s.pos + t] = 0;
        }
        s.pos += chunk;
        done += chunk;
        if (s.pos == SR) cs_run_f(s);
    }
    CTA_SYNC();
}

}  // namespace cstr
}  // namespace cdp
