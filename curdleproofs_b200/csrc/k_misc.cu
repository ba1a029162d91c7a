#include <cstdlib>

#include "launch.h"
#include "msm_common.cuh"
#include "g1_quad.cuh"

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
    if (valid && nwin == 1) {  // the windows are already combined (k_msm_horner_quad): bucket_sums = S_b, only the reduction over b is left
        g1j_load(acc, bucket_sums + 36 * ((size_t)i * nb + b));
    } else if (valid) {
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


// The Horner part of k_msm_combine with a QUAD of lanes per (MSM, bucket index) -- g1_quad.cuh: the ~130 dependent doublings and nwin
// additions run at 3 resp. 5 product latencies each instead of 7 resp. 16.  Writes S_b = sum_w 2^(c w) B_{w,b} per (MSM, b); the weighted
// reduction over b is then k_msm_combine with nwin = 1.  For launches of FEW MSMs only (a lone `util::msm`, small proof batches): with many
// MSMs in flight the plain kernel keeps the machine just as busy with a quarter of the threads.
__global__ void __launch_bounds__(128) k_msm_horner_quad(const uint32_t *__restrict__ bucket_sums, uint32_t *__restrict__ S_out, uint32_t n_msm, int c,
                                                         int nwin) {
    const int nb = 1 << (c - 1);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, quad = gid >> 2;
    const int s = (int)(threadIdx.x & 3);
    const uint32_t i = quad / nb;
    const int b = (int)(quad % nb);
    const bool valid = i < n_msm;
    g1j acc;
    g1j_set_inf(acc);
    const uint32_t *B = bucket_sums + 36 * ((size_t)(valid ? i : 0) * nwin * nb + b);
    // top window: bucket b < nbt is spread over sp slots b*sp .. b*sp + sp - 1 (k_msm_buckets); the other quads start from infinity
    const int tb = 128 - c * (nwin - 1), nbt = tb > 0 ? (1 << tb) : 1, sp = nb / nbt >= 1 ? nb / nbt : 1;
    const bool has_top = valid && b < nbt;
    const uint32_t *T = bucket_sums + 36 * (((size_t)(valid ? i : 0) * nwin + (nwin - 1)) * nb + (size_t)(has_top ? b : 0) * sp);
#pragma unroll 1
    for (int k = 0; k < sp; k++) {  // sp is the same for every quad: the shuffles inside stay warp-uniform
        g1j W;
        g1j_set_inf(W);
        if (has_top) g1j_load(W, T + 36 * (size_t)k);
        g1j_add_quad(acc, acc, W, s);
    }
#pragma unroll 1
    for (int w = nwin - 2; w >= 0; w--) {
#pragma unroll 1
        for (int k = 0; k < c; k++) g1j_dbl_quad(acc, s);
        g1j W;
        g1j_set_inf(W);
        if (valid) g1j_load(W, B + 36 * (size_t)w * nb);
        g1j_add_quad(acc, acc, W, s);
    }
    if (valid && s == 0) g1j_store(S_out + 36 * ((size_t)i * nb + b), acc);
}

// k_msm_horner_quad with the ADDITIONS of the chain taken out of it: the windows are first combined in groups of G (k_msm_horner_part1: a quad
// per (MSM, group, bucket index), all in parallel, (G - 1) c doublings and G additions deep), then one quad per (MSM, bucket index) runs Horner
// over the ceil(nwin / G) group sums (k_msm_horner_part2).  The ~128 doublings of the chain are inherent, but its additions drop from nwin (26 at
// c = 5, ~10 us each for a lone quad) to G + ceil(nwin / G): ~0.1 ms of the ~0.6 ms a lone `util::msm` spends here.
__global__ void __launch_bounds__(128) k_msm_horner_part1(const uint32_t *__restrict__ bucket_sums, uint32_t *__restrict__ V, uint32_t n_msm, int c, int nwin,
                                                          int G) {
    const int nb = 1 << (c - 1), ng = (nwin + G - 1) / G;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, quad = gid >> 2;
    const int s = (int)(threadIdx.x & 3);
    const uint32_t i = quad / (uint32_t)(ng * nb);
    const int j = (int)(quad / nb) % ng, b = (int)(quad % nb);
    const bool valid = i < n_msm;
    const int tb = 128 - c * (nwin - 1), nbt = tb > 0 ? (1 << tb) : 1, sp = nb / nbt >= 1 ? nb / nbt : 1;
    // the top window: bucket b < nbt is spread over sp slots (k_msm_buckets); only the last group's quads have something to add here
    const bool has_top = valid && j == ng - 1 && b < nbt;
    const uint32_t *T = bucket_sums + 36 * (((size_t)(valid ? i : 0) * nwin + (nwin - 1)) * nb + (size_t)(has_top ? b : 0) * sp);
    g1j top;
    g1j_set_inf(top);
#pragma unroll 1
    for (int k = 0; k < sp; k++) {
        g1j W;
        g1j_set_inf(W);
        if (has_top) g1j_load(W, T + 36 * (size_t)k);
        g1j_add_quad(top, top, W, s);
    }
    const uint32_t *B = bucket_sums + 36 * ((size_t)(valid ? i : 0) * nwin * nb + b);
    g1j acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (int t = G - 1; t >= 0; t--) {
        if (t != G - 1) {
#pragma unroll 1
            for (int k = 0; k < c; k++) g1j_dbl_quad(acc, s);
        }
        const int w = j * G + t;
        g1j W;
        g1j_set_inf(W);
        if (valid && w < nwin - 1) g1j_load(W, B + 36 * (size_t)w * nb);
        else if (valid && w == nwin - 1) W = top;
        g1j_add_quad(acc, acc, W, s);
    }
    if (valid && s == 0) g1j_store(V + 36 * (((size_t)i * ng + j) * nb + b), acc);
}
__global__ void __launch_bounds__(128) k_msm_horner_part2(const uint32_t *__restrict__ V, uint32_t *__restrict__ S_out, uint32_t n_msm, int nb, int ng,
                                                          int dbl_per_step) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, quad = gid >> 2;
    const int s = (int)(threadIdx.x & 3);
    const uint32_t i = quad / (uint32_t)nb;
    const int b = (int)(quad % nb);
    const bool valid = i < n_msm;
    const uint32_t *Vi = V + 36 * ((size_t)(valid ? i : 0) * ng * nb + b);
    g1j acc;
    g1j_set_inf(acc);
    if (valid) g1j_load(acc, Vi + 36 * (size_t)(ng - 1) * nb);
#pragma unroll 1
    for (int jj = ng - 2; jj >= 0; jj--) {
#pragma unroll 1
        for (int k = 0; k < dbl_per_step; k++) g1j_dbl_quad(acc, s);
        g1j W;
        g1j_set_inf(W);
        if (valid) g1j_load(W, Vi + 36 * (size_t)jj * nb);
        g1j_add_quad(acc, acc, W, s);
    }
    if (valid && s == 0) g1j_store(S_out + 36 * ((size_t)i * nb + b), acc);
}

// The reduction over the bucket indices (out = sum_b (b + 1) S_b: suffix scan + tree sum, as the tail of k_msm_combine) with a quad per
// bucket index: one CTA per MSM, partners exchanged through shared memory (nb quads = up to 128 lanes do not fit a warp's shuffles).
__global__ void __launch_bounds__(128) k_msm_reduce_quad(const uint32_t *__restrict__ S, uint32_t *__restrict__ out_jac, int c) {
    __shared__ uint32_t sm[32 * 36];
    const int nb = 1 << (c - 1);
    const int b = (int)(threadIdx.x >> 2), s = (int)(threadIdx.x & 3);
    const uint32_t i = blockIdx.x;
    g1j acc;
    g1j_set_inf(acc);
    if (b < nb) g1j_load(acc, S + 36 * ((size_t)i * nb + b));  // the CTA is padded to a whole warp: quads beyond nb carry infinity
#pragma unroll 1
    for (int step = 0; step < 2 * (c - 1); step++) {
        const bool scan = step < c - 1;
        const int d = scan ? (1 << step) : (nb >> (step - (c - 1) + 1));
        const bool take = b < nb && (scan ? (b + d < nb) : (b < d));
        if (s == 0 && b < nb) g1j_store(sm + 36 * b, acc);
        __syncthreads();
        g1j o;
        g1j_set_inf(o);
        if (take) g1j_load(o, sm + 36 * (b + d));
        __syncthreads();
        g1j_add_quad(acc, acc, o, s);
    }
    if (b == 0 && s == 0) g1j_store(out_jac + 36 * (size_t)i, acc);
}

// k_sum_groups with a quad per partial sum (8 quads per output point): for launches of few outputs, where the plain kernel is a chain of
// ~per_out / 32 + 5 dependent full additions
__global__ void __launch_bounds__(128) k_sum_groups_quad(const uint32_t *__restrict__ A, uint32_t per_a, uint32_t sa, uint32_t ga, uint32_t *__restrict__ out,
                                                         uint32_t n_out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_out) return;  // whole warps leave together
    const uint32_t qd = lane >> 2;
    const int s = (int)(lane & 3);
    g1j acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (uint32_t base = 0; base < per_a; base += 8) {
        g1j q;
        g1j_set_inf(q);
        if (base + qd < per_a) g1j_load(q, A + 36 * ((size_t)(base + qd) * sa + (size_t)warp * ga));
        g1j_add_quad(acc, acc, q, s);
    }
#pragma unroll 1
    for (int d = 4; d >= 1; d >>= 1) {
        g1j o;
        shfl_down_g1j(o, acc, 4 * d, 32);
        g1j_add_quad(acc, acc, o, s);
    }
    if (lane == 0) g1j_store(out + 36 * (size_t)warp, acc);
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
uint32_t horner_groups_max() {
    static const uint32_t v = [] { const char *e = getenv("CDP_HORNER_GROUP_MAX"); return e ? (uint32_t)atoi(e) : 64u; }();
    return v;
}
size_t horner_groups_scratch_bytes(uint32_t n_msm, int c, int nwin) {  // for any G >= 2
    return (size_t)n_msm * ((nwin + 1) / 2) * ((size_t)1 << (c - 1)) * 144;
}
cudaError_t launch_msm_horner_quad(cudaStream_t st, const uint32_t *bucket_sums, uint32_t *S_out, uint32_t n_msm, int c, int nwin, uint32_t *scratch) {
    // few MSMs (a lone `util::msm`, small batches): the windows combined in groups first; CDP_HORNER_GROUP = 0 keeps the plain chain.
    // `scratch` holds n_msm * ceil(nwin / G) * nb Jacobian group sums
    static const int G = [] { const char *e = getenv("CDP_HORNER_GROUP"); return e ? atoi(e) : 3; }();  // measured: 2, 3 < 4 < 6
    if (G >= 2 && scratch && n_msm <= horner_groups_max()) {
        const int nb = 1 << (c - 1), ng = (nwin + G - 1) / G;
        const uint64_t q1 = (uint64_t)n_msm * ng * nb * 4, q2 = (uint64_t)n_msm * nb * 4;
        k_msm_horner_part1<<<(unsigned)((q1 + 127) / 128), 128, 0, st>>>(bucket_sums, scratch, n_msm, c, nwin, G);
        k_msm_horner_part2<<<(unsigned)((q2 + 127) / 128), 128, 0, st>>>(scratch, S_out, n_msm, nb, ng, G * c);
        return cudaGetLastError();
    }
    const uint64_t threads = ((uint64_t)n_msm << (c - 1)) * 4;
    k_msm_horner_quad<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(bucket_sums, S_out, n_msm, c, nwin);
    return cudaGetLastError();
}
cudaError_t launch_msm_reduce_quad(cudaStream_t st, const uint32_t *S, uint32_t *out_jac, uint32_t n_msm, int c) {
    const unsigned threads = 4u << (c - 1);
    k_msm_reduce_quad<<<n_msm, threads < 32 ? 32 : threads, 0, st>>>(S, out_jac, c);
    return cudaGetLastError();
}
cudaError_t launch_sum_groups(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n_out, uint32_t per_out, uint32_t group_stride) {
    if (n_out <= 2048 && per_out >= 2) {  // few outputs: latency matters, not lane efficiency
        k_sum_groups_quad<<<(n_out * 32 + 127) / 128, 128, 0, st>>>(in, per_out, group_stride, 1, out, n_out);
        return cudaGetLastError();
    }
    k_sum_groups<<<(n_out * 32 + 127) / 128, 128, 0, st>>>(in, per_out, group_stride, 1, nullptr, 0, 0, 0, out, n_out);
    return cudaGetLastError();
}
cudaError_t launch_sum_groups2(cudaStream_t st, const uint32_t *A, uint32_t per_a, uint32_t sa, uint32_t ga, const uint32_t *B, uint32_t per_b,
                               uint32_t sb, uint32_t gb, uint32_t *out, uint32_t n_out) {
    k_sum_groups<<<(n_out * 32 + 127) / 128, 128, 0, st>>>(A, per_a, sa, ga, B, per_b, sb, gb, out, n_out);
    return cudaGetLastError();
}

}  // namespace cdp
