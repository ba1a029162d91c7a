#include "launch.h"
#include "g1.cuh"

namespace cdp {

// ------------------------------------------------------------------------------------------------ k_normalize
// Jacobian -> affine with Montgomery's simultaneous-inversion trick over CHUNK points per thread.
// out_affine (nullable): 96 B Montgomery x||y, infinity -> all zero.   out_comp (nullable): 48 B ZCash encoding.
__device__ __forceinline__ void g1a_compress_store(uint8_t *dst, const g1a &p, bool inf) {
    uint32_t w[12];
    if (inf) {
#pragma unroll
        for (int i = 0; i < 12; i++) w[i] = 0;
        w[11] = 0xC0000000u;
    } else {
        fp xc, yc;
        fp_from_mont(xc, p.x);
        fp_from_mont(yc, p.y);
#pragma unroll
        for (int i = 0; i < 12; i++) w[i] = xc.v[i];
        w[11] |= 0x80000000u;
        if (fp_canon_is_lexicographically_largest(yc)) w[11] |= 0x20000000u;
    }
    // big-endian byte order: most significant limb first, bytes swapped
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = __byte_perm(w[11 - i], 0, 0x0123);
}

template <int CHUNK>
__global__ void __launch_bounds__(128) k_normalize(const uint32_t *__restrict__ jac, uint32_t *__restrict__ out_affine,
                                                   uint8_t *__restrict__ out_comp, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t base = t * CHUNK;
    if (base >= n) return;
    fp pre[CHUNK];
    fp acc;
    fp_set_one(acc);
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        pre[j] = acc;
        if (base + j < n) {
            fp Z;
            fp_load(Z, jac + 36 * (size_t)(base + j) + 24);
            if (!fp_is_zero(Z)) fp_mul(acc, acc, Z);
        }
    }
    fp inv;
    fp_inv(inv, acc);
#pragma unroll
    for (int j = CHUNK - 1; j >= 0; j--) {
        if (base + j < n) {
            g1j P;
            g1j_load(P, jac + 36 * (size_t)(base + j));
            bool inf = fp_is_zero(P.Z);
            g1a A;
            if (inf) {
                g1a_set_inf(A);
            } else {
                fp zi, zi2;
                fp_mul(zi, inv, pre[j]);
                fp_mul(inv, inv, P.Z);
                fp_sqr(zi2, zi);
                fp_mul(A.x, P.X, zi2);
                fp_mul(zi2, zi2, zi);
                fp_mul(A.y, P.Y, zi2);
            }
            if (out_affine) g1a_store(out_affine + 24 * (size_t)(base + j), A);
            if (out_comp) g1a_compress_store(out_comp + 48 * (size_t)(base + j), A, inf);
        }
    }
}

cudaError_t launch_normalize(cudaStream_t st, int chunk, const uint32_t *jac, uint32_t *out_affine, uint8_t *out_comp, uint32_t n) {
    uint32_t threads = (n + chunk - 1) / chunk, blocks = (threads + 127) / 128;
    if (chunk == 8) k_normalize<8><<<blocks, 128, 0, st>>>(jac, out_affine, out_comp, n);
    else if (chunk == 2) k_normalize<2><<<blocks, 128, 0, st>>>(jac, out_affine, out_comp, n);
    else k_normalize<1><<<blocks, 128, 0, st>>>(jac, out_affine, out_comp, n);
    return cudaGetLastError();
}

}  // namespace cdp
