// Fp inversion by the binary extended Euclidean algorithm on the integer representative, in plain C++ (no PTX) so that the same source is
// checked on the CPU (tests/host/fp_inv_check.cpp).  Replaces the Fermat ladder a^(p-2) (381 squarings + ~95 products, all dependent) in
// `into_affine` / `normalize_batch` (k_normalize), where one inversion per thread is the whole latency of the launch: ~760 shift /
// subtract steps on 12 limbs -- integer-ALU work, off the multiply pipe the MSM kernels saturate -- instead of ~285 k multiply-pipe
// instructions.  The steps are branch-free (selects) and the loop runs until every lane of the warp is done, so lanes do not diverge.
//
//   u = a, v = p, x1 = 1, x2 = 0;  invariants  x1 a = u, x2 a = v (mod p)
//   step: u even: u /= 2, x1 /= 2 | v even: v /= 2, x2 /= 2 | u >= v: u = (u - v) / 2, x1 = (x1 - x2) / 2 | else: v = (v - u) / 2, x2 = (x2 - x1) / 2
//   until u = 1 (result x1) or v = 1 (result x2).  Halving mod p: (x + (x odd ? p : 0)) >> 1.
#pragma once
#include <stdint.h>

namespace cdp {
namespace euclid {

#ifndef __CUDACC__
#define CDP_EUCLID_FN inline
#else
#define CDP_EUCLID_FN __device__ __forceinline__
#endif

CDP_EUCLID_FN uint32_t p_limb(int i) {
    const uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                            0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return t[i];
}
// r = a - b, returns the borrow
CDP_EUCLID_FN uint32_t sub12(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        r[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 32) & 1u;
    }
    return borrow;
}
// x <- x / 2 mod p  (x < p)
CDP_EUCLID_FN void halve_mod_p(uint32_t *x) {
    const uint32_t odd = 0u - (x[0] & 1u);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        c += (uint64_t)x[i] + (p_limb(i) & odd);
        x[i] = (uint32_t)c;
        c >>= 32;
    }
#pragma unroll
    for (int i = 0; i < 11; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[11] = (x[11] >> 1) | ((uint32_t)c << 31);
}
CDP_EUCLID_FN bool is_one(const uint32_t *a) {
    uint32_t acc = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 12; i++) acc |= a[i];
    return acc == 0;
}

// out = a^-1 mod p as an integer (a in [1, p)); a = 0 gives 0.  `all_done` folds the per-lane "finished" flags of the executing warp
// (device: __all_sync; host: identity).
template <class AllDone>
CDP_EUCLID_FN void inverse_int(uint32_t *out, const uint32_t *a, AllDone all_done) {
    uint32_t u[12], v[12], x1[12], x2[12];
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u[i] = a[i];
        v[i] = p_limb(i);
        x1[i] = 0;
        x2[i] = 0;
        nz |= a[i];
    }
    x1[0] = 1;
    bool done = nz == 0 || is_one(u);
#pragma unroll 1
    for (int it = 0; it < 768; it++) {
        if (all_done(done)) break;
        // both differences, then one of four updates selected per lane; a finished lane keeps its state
        uint32_t d_uv[12], d_vu[12], dx12[12], dx21[12];
        const uint32_t b_uv = sub12(d_uv, u, v);  // borrow <=> u < v
        sub12(d_vu, v, u);
        const uint32_t bx12 = sub12(dx12, x1, x2), bx21 = sub12(dx21, x2, x1);
        {   // reduce the x differences into [0, p)
            const uint32_t m12 = 0u - bx12, m21 = 0u - bx21;
            uint64_t c1 = 0, c2 = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                c1 += (uint64_t)dx12[i] + (p_limb(i) & m12);
                dx12[i] = (uint32_t)c1;
                c1 >>= 32;
                c2 += (uint64_t)dx21[i] + (p_limb(i) & m21);
                dx21[i] = (uint32_t)c2;
                c2 >>= 32;
            }
        }
        const bool u_even = (u[0] & 1u) == 0, v_even = (v[0] & 1u) == 0, u_ge_v = b_uv == 0;
        // which pair changes: (u, x1) when u is even, or when both are odd and u >= v; otherwise (v, x2)
        const bool upd_u = !done && (u_even || (!v_even && u_ge_v));
        const bool upd_v = !done && !upd_u;
        const bool sub_u = !u_even, sub_v = !v_even;  // (only meaningful for the pair being updated) take the difference first
        uint32_t nu[12], nx[12];
#pragma unroll
        for (int i = 0; i < 12; i++) {
            nu[i] = upd_u ? (sub_u ? d_uv[i] : u[i]) : (sub_v ? d_vu[i] : v[i]);
            nx[i] = upd_u ? (sub_u ? dx12[i] : x1[i]) : (sub_v ? dx21[i] : x2[i]);
        }
#pragma unroll
        for (int i = 0; i < 11; i++) nu[i] = (nu[i] >> 1) | (nu[i + 1] << 31);
        nu[11] >>= 1;
        halve_mod_p(nx);
#pragma unroll
        for (int i = 0; i < 12; i++) {
            if (upd_u) { u[i] = nu[i]; x1[i] = nx[i]; }
            if (upd_v) { v[i] = nu[i]; x2[i] = nx[i]; }
        }
        done = done || is_one(u) || is_one(v);
    }
    const bool from_u = is_one(u);
#pragma unroll
    for (int i = 0; i < 12; i++) out[i] = nz == 0 ? 0u : (from_u ? x1[i] : x2[i]);
}

}  // namespace euclid
}  // namespace cdp
