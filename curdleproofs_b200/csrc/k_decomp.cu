// k_decompress: 48-byte ZCash-format G1 encodings -> affine points, with the checks ark-serialize's
// `deserialize_compressed` (Validate::Yes) performs: canonical x < p, on the curve, in the prime-order subgroup.
// This is the step immediately before the verifier's hot path (`CurdleproofsProof::deserialize`,
// /root/reference/src/curdleproofs.rs:312-323; `from_bytes_g1affine`, src/whisk.rs:313-315).
//
// Subgroup test without a 255-bit multiplication (Scott, ePrint 2021/1130 section 6): with phi(x, y) = (beta x, y) acting as
// lambda = z^2 - 1 on G1 and 1 + phi + phi^2 = 0,   P in G1  <=>  [z^2] P == -phi^2(P) = (beta^2 x, -y).
// [z^2]P is two multiplications by |z| = 0xd201000000010000 (Hamming weight 6): 126 doublings + 10 additions.
#include "launch.h"
#include "scalar.cuh"

namespace cdp {

__device__ __forceinline__ void mul_by_z(g1j &r, const g1j &p) {
    // |z| = 0xd201000000010000, MSB first: bits 63,62,60,57,48,16 set
    g1j acc = p;
    const unsigned long long Z = 0xd201000000010000ULL;
#pragma unroll 1
    for (int bit = 62; bit >= 0; bit--) {
        g1j_dbl(acc, acc);
        if ((Z >> bit) & 1ULL) g1j_add(acc, acc, p);
    }
    r = acc;
}

// status: 0 ok, 1 malformed encoding / x >= p, 2 not on curve, 3 not in subgroup
__global__ void __launch_bounds__(128, 3) k_decompress(const uint8_t *__restrict__ comp, const uint32_t *__restrict__ dst_idx,
                                                    uint32_t *__restrict__ out_affine, uint8_t *__restrict__ status, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t *c = reinterpret_cast<const uint32_t *>(comp + 48 * (size_t)i);
    uint32_t w[12];
#pragma unroll
    for (int k = 0; k < 12; k++) w[11 - k] = __byte_perm(c[k], 0, 0x0123);  // big-endian bytes -> little-endian limbs
    const uint32_t flags = w[11] >> 29;
    w[11] &= 0x1FFFFFFFu;
    uint32_t *dst = out_affine + 24 * (size_t)(dst_idx ? dst_idx[i] : i);
    uint8_t st = 0;
    g1a P;
    g1a_set_inf(P);
    if (!(flags & 4u)) {
        st = 1;  // uncompressed form is not accepted here
    } else if (flags & 2u) {
        uint32_t nz = flags & 1u;
#pragma unroll
        for (int k = 0; k < 12; k++) nz |= w[k];
        if (nz) st = 1;  // infinity must be 0xC0 00 .. 00
    } else {
        fp xc, t;
#pragma unroll
        for (int k = 0; k < 12; k++) xc.v[k] = w[k];
        if (fp_sub_p(t, xc) == 0) {
            st = 1;  // x >= p
        } else {
            fp x, rhs, y, b4;
            fp_to_mont(x, xc);
#pragma unroll
            for (int k = 0; k < 12; k++) b4.v[k] = FP_B_MONT[k];
            fp_sqr(rhs, x);
            fp_mul(rhs, rhs, x);
            fp_add(rhs, rhs, b4);
            if (!fp_sqrt(y, rhs)) {
                st = 2;
            } else {
                fp yc, ny;
                fp_from_mont(yc, y);
                fp_neg(ny, y);
                bool largest = fp_canon_is_lexicographically_largest(yc);
                if (largest != ((flags & 1u) != 0)) y = ny;
                P.x = x;
                P.y = y;
                // subgroup: [z^2]P == (beta^2 x, -y)
                g1j J, Q;
                g1j_from_affine(J, P);
                mul_by_z(Q, J);
                mul_by_z(Q, Q);
                fp bx, z2, z3, lhs, rx;
                fp_mul_beta(bx, P.x);
                fp_mul_beta(bx, bx);
                fp_sqr(z2, Q.Z);
                fp_mul(z3, z2, Q.Z);
                fp_mul(rx, bx, z2);
                fp_neg(ny, P.y);
                fp_mul(lhs, ny, z3);
                if (fp_is_zero(Q.Z) || !fp_eq(rx, Q.X) || !fp_eq(lhs, Q.Y)) {
                    st = 3;
                    g1a_set_inf(P);
                }
            }
        }
    }
    g1a_store(dst, P);
    status[i] = st;
}

cudaError_t launch_decompress(cudaStream_t st, const uint8_t *comp, const uint32_t *dst_idx, uint32_t *out_affine, uint8_t *status, uint32_t n) {
    if (n == 0) return cudaSuccess;
    k_decompress<<<(n + 127) / 128, 128, 0, st>>>(comp, dst_idx, out_affine, status, n);
    return cudaGetLastError();
}

}  // namespace cdp
