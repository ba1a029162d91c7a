#include "launch.h"
#include "msm_common.cuh"

namespace cdp {

// Window combine of a batch of MSMs from their bucket sums [msm][window][bucket].  One thread per (MSM, bucket index b), the
// nb = 2^(c-1) lanes of an MSM side by side in a warp:
//   S_b = sum_w 2^(c w) B_{w,b}           Horner from the top window down, all lanes doubling in parallel
//   out = sum_b (b + 1) S_b               suffix scan + tree sum with warp shuffles (2 (c-1) full additions deep), once per MSM
// -- the same group element as reducing every window separately and combining the window sums, with nwin times fewer reductions.
__global__ void __launch_bounds__(128) k_msm_combine(const uint32_t *__restrict__ bucket_sums, uint32_t *__restrict__ out_jac, uint32_t n_msm,
                                                     int c, int nwin) {
    const int nb = 1 << (c - 1);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = gid / nb;
    const int b = (int)(gid % nb);
    const bool valid = i < n_msm;
    g1j acc;
    g1j_set_inf(acc);
    if (valid) {
        const uint32_t *B = bucket_sums + 36 * ((size_t)i * nwin * nb + b);
        // top window: bucket b < nbt is spread over sp slots b*sp .. b*sp + sp - 1 (k_msm_buckets); the other lanes start from infinity
        const int tb = 128 - c * (nwin - 1), nbt = tb > 0 ? (1 << tb) : 1, sp = nb / nbt >= 1 ? nb / nbt : 1;
        if (b < nbt) {
            const uint32_t *T = bucket_sums + 36 * (((size_t)i * nwin + (nwin - 1)) * nb + (size_t)b * sp);
            g1j_load(acc, T);
#pragma unroll 1
            for (int s = 1; s < sp; s++) {
                g1j W;
                g1j_load(W, T + 36 * (size_t)s);
                g1j_add(acc, acc, W);
            }
        }
#pragma unroll 1
        for (int w = nwin - 2; w >= 0; w--) {
#pragma unroll 1
            for (int k = 0; k < c; k++) g1j_dbl(acc, acc);
            g1j W;
            g1j_load(W, B + 36 * (size_t)w * nb);
            g1j_add(acc, acc, W);
        }
    }
#pragma unroll 1
    for (int step = 0; step < 2 * (c - 1); step++) {
        const bool scan = step < c - 1;
        const int d = scan ? (1 << step) : (nb >> (step - (c - 1) + 1));
        const bool take = scan ? (b + d < nb) : (b < d);
        g1j o;
        shfl_down_g1j(o, acc, d, nb);
        if (take) g1j_add(acc, acc, o);
    }
    if (valid && b == 0) g1j_store(out_jac + 36 * (size_t)i, acc);
}


// out[g] = sum_{s < per_a} A[s * sa + g * ga]  (+ sum_{s < per_b} B[s * sb + g * gb])   -- one warp per output point
__global__ void __launch_bounds__(128) k_sum_groups(const uint32_t *__restrict__ A, uint32_t per_a, uint32_t sa, uint32_t ga,
                                                    const uint32_t *__restrict__ Bsrc, uint32_t per_b, uint32_t sb, uint32_t gb,
                                                    uint32_t *__restrict__ out, uint32_t n_out) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_out) return;
    g1j acc;
    g1j_set_inf(acc);
    const uint32_t total = per_a + per_b;
#pragma unroll 1
    for (uint32_t s = lane; s < ((total + 31) & ~31u); s += 32) {
        g1j q;
        g1j_set_inf(q);
        if (s < per_a) g1j_load(q, A + 36 * ((size_t)s * sa + (size_t)warp * ga));
        else if (s < total) g1j_load(q, Bsrc + 36 * ((size_t)(s - per_a) * sb + (size_t)warp * gb));
        g1j_add(acc, acc, q);
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        g1j o;
        shfl_down_g1j(o, acc, d, 32);
        g1j_add(acc, acc, o);
    }
    if (lane == 0) g1j_store(out + 36 * (size_t)warp, acc);
}

cudaError_t launch_msm_buckets(cudaStream_t st, int c, const uint32_t *pts, const uint32_t *scalars, const msm_seg_t *segs, uint32_t count,
                               uint32_t nmax, int8_t *dig, uint32_t *win_sums) {
    switch (c) {
        case 6: return launch_msm_buckets_c6(st, pts, scalars, segs, count, nmax, dig, win_sums);
        case 5: return launch_msm_buckets_c5(st, pts, scalars, segs, count, nmax, dig, win_sums);
        case 4: return launch_msm_buckets_c4(st, pts, scalars, segs, count, nmax, dig, win_sums);
        case 3: return launch_msm_buckets_c3(st, pts, scalars, segs, count, nmax, dig, win_sums);
        case 2: return launch_msm_buckets_c2(st, pts, scalars, segs, count, nmax, dig, win_sums);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t launch_msm_combine(cudaStream_t st, const uint32_t *bucket_sums, uint32_t *out_jac, uint32_t n_msm, int c, int nwin) {
    const uint64_t threads = (uint64_t)n_msm << (c - 1);
    k_msm_combine<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(bucket_sums, out_jac, n_msm, c, nwin);
    return cudaGetLastError();
}
cudaError_t launch_sum_groups(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n_out, uint32_t per_out, uint32_t group_stride) {
    k_sum_groups<<<(n_out * 32 + 127) / 128, 128, 0, st>>>(in, per_out, group_stride, 1, nullptr, 0, 0, 0, out, n_out);
    return cudaGetLastError();
}
cudaError_t launch_sum_groups2(cudaStream_t st, const uint32_t *A, uint32_t per_a, uint32_t sa, uint32_t ga, const uint32_t *B, uint32_t per_b,
                               uint32_t sb, uint32_t gb, uint32_t *out, uint32_t n_out) {
    k_sum_groups<<<(n_out * 32 + 127) / 128, 128, 0, st>>>(A, per_a, sa, ga, B, per_b, sb, gb, out, n_out);
    return cudaGetLastError();
}

}  // namespace cdp
