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

// With `jobs` != nullptr the affine result of element i = j * elems_per_job + e is scattered to out_affine[jobs[j].out_off + e]
// (the fold writes its result back into the L half of the proof's working vector); otherwise it goes to out_affine[i].
template <int CHUNK>
__global__ void __launch_bounds__(128) k_normalize(const uint32_t *__restrict__ jac, uint32_t *__restrict__ out_affine,
                                                   uint8_t *__restrict__ out_comp, uint32_t n, const smul_job_t *__restrict__ jobs,
                                                   uint32_t elems_per_job) {
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
            if (out_affine) {
                size_t dst = base + j;
                if (jobs) {
                    uint32_t jj = (base + j) / elems_per_job, e = (base + j) - jj * elems_per_job;
                    dst = (size_t)jobs[jj].out_off + e;
                }
                g1a_store(out_affine + 24 * dst, A);
            }
            if (out_comp) g1a_compress_store(out_comp + 48 * (size_t)(base + j), A, inf);
        }
    }
}

// affine points (gathered through idx when given) -> 48-byte encodings: `serialize_compressed` of the instance vectors
// vec_R, vec_S, vec_T, vec_U that open the transcript (/root/reference/src/curdleproofs.rs:81).
__global__ void __launch_bounds__(128) k_compress_affine(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ idx,
                                                         uint8_t *__restrict__ out_comp, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1a A;
    g1a_load(A, pts + 24 * (size_t)(idx ? idx[i] : i));
    g1a_compress_store(out_comp + 48 * (size_t)i, A, g1a_is_inf(A));
}

// pts[dst_idx[i]] = src[src_idx[i]]  (96-byte points; builds the per-proof working vectors out of the CRS and the inputs)
__global__ void __launch_bounds__(256) k_gather_points(uint32_t *__restrict__ pts, const uint32_t *__restrict__ src,
                                                       const uint32_t *__restrict__ src_idx, const uint32_t *__restrict__ dst_idx, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = t / 6, q = t % 6;  // six 16-byte pieces per point: consecutive threads move consecutive 16 B
    if (i >= n) return;
    const uint4 *s = reinterpret_cast<const uint4 *>(src + 24 * (size_t)src_idx[i]);
    uint4 *d = reinterpret_cast<uint4 *>(pts + 24 * (size_t)dst_idx[i]);
    d[q] = s[q];
}

cudaError_t launch_normalize(cudaStream_t st, int chunk, const uint32_t *jac, uint32_t *out_affine, uint8_t *out_comp, uint32_t n,
                             const smul_job_t *jobs, uint32_t elems_per_job) {
    uint32_t threads = (n + chunk - 1) / chunk, blocks = (threads + 127) / 128;
    if (chunk == 8) k_normalize<8><<<blocks, 128, 0, st>>>(jac, out_affine, out_comp, n, jobs, elems_per_job);
    else if (chunk == 2) k_normalize<2><<<blocks, 128, 0, st>>>(jac, out_affine, out_comp, n, jobs, elems_per_job);
    else k_normalize<1><<<blocks, 128, 0, st>>>(jac, out_affine, out_comp, n, jobs, elems_per_job);
    return cudaGetLastError();
}
cudaError_t launch_compress_affine(cudaStream_t st, const uint32_t *pts, const uint32_t *idx, uint8_t *out_comp, uint32_t n) {
    k_compress_affine<<<(n + 127) / 128, 128, 0, st>>>(pts, idx, out_comp, n);
    return cudaGetLastError();
}
cudaError_t launch_gather_points(cudaStream_t st, uint32_t *pts, const uint32_t *src, const uint32_t *src_idx, const uint32_t *dst_idx, uint32_t n) {
    k_gather_points<<<(n * 6 + 255) / 256, 256, 0, st>>>(pts, src, src_idx, dst_idx, n);
    return cudaGetLastError();
}

}  // namespace cdp
