// Large Pippenger MSM (n >= 2^13): the standalone sweep of BASELINE.json config 5 and any accumulated-verify MSM that is
// merged across a batch.  `util::msm` semantics (/root/reference/src/util.rs:19-22): out = sum_i s_i * P_i.
//
//   k_big_digits     one thread per pair: GLV split (2n half-width points), signed radix-2^c recoding; writes one
//                    (bucket key, point id | sign) pair per (window, half-point) -- coalesced 128-bit scalar loads, coalesced stores
//   one cub::DeviceRadixSort over (window | bucket) keys groups the point ids of each bucket (every window keeps exactly 2n items,
//                    so window w is the slice [w * 2n, (w+1) * 2n) of the sorted array)
//   k_big_offsets    bucket boundaries from the sorted keys
//   k_big_loads      + a 12-bit radix sort: the accumulate slots in descending order of their load (balanced warps)
//   k_big_accumulate one thread per (window, bucket): mixed additions over the bucket's points (gathered 96-byte loads, the
//                    base array is L2-resident up to ~2^20 points); the top window's few long buckets are spread over sp_top threads
//                    (k_big_fold_top adds their partial sums); a bucket whose per-thread share exceeds BIG_HEAVY points (skewed
//                    scalars) goes to a device-side work list instead: k_big_heavy (one CTA per 4096-point chunk), k_big_heavy_fold
//   k_big_reduce_level   sum_b (b+1) B_b per window as a hierarchy of 16-ary running sums (A = plain sum, Bv = index-weighted sum per node)
//   k_big_horner     Horner over the windows.
#include <cub/device/device_radix_sort.cuh>

#include "launch.h"
#include "msm_common.cuh"
#include "g1_quad.cuh"

namespace cdp {

__global__ void __launch_bounds__(256) k_big_digits(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ scalars, uint32_t n, int c,
                                                    int nwin, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ bx) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t k[8];
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * (size_t)j);
    uint4 a = sp[0], b = sp[1];
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    const uint4 *pp = reinterpret_cast<const uint4 *>(pts + 24 * (size_t)j);
    uint32_t nz = 0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        uint4 v = pp[q];
        nz |= v.x | v.y | v.z | v.w;
    }
    if (bx) {  // beta x, the x coordinate of the endomorphism image: once per base here instead of once per (window, point) in the gather
        fp x;
        fp_load(x, pts + 24 * (size_t)j);
        fp_mul_beta(x, x);
        fp_store(bx + 12 * (size_t)j, x);
    }
    glv_t g;
    glv_split(g, k);
    if (nz == 0) {  // infinity base contributes nothing
#pragma unroll
        for (int q = 0; q < 4; q++) g.k1[q] = g.k2[q] = 0;
    }
    const uint32_t nb = 1u << (c - 1), n2 = 2 * n;
    // both halves of a pair go out as one 8-byte store per window (ids 2j, 2j + 1 are neighbours): fully coalesced
    uint32_t kk[2][5] = {{g.k1[0], g.k1[1], g.k1[2], g.k1[3], 0}, {g.k2[0], g.k2[1], g.k2[2], g.k2[3], 0}};
    uint32_t carry[2] = {0, 0};
    for (int w = 0; w < nwin; w++) {
        const int bit = w * c, li = bit >> 5, sh = bit & 31;
        uint32_t key[2], val[2];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t v = 0;
            if (li < 4) {
                v = kk[half][li] >> sh;
                if (sh + c > 32) v |= kk[half][li + 1] << (32 - sh);
            }
            v = (v & ((1u << c) - 1)) + carry[half];
            carry[half] = (v + nb) >> c;
            const int d = (int)v - (int)(carry[half] << c);
            const uint32_t ad = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            // key = window | bucket; bucket nb = "digit zero", sorted behind every real bucket of its window
            key[half] = ((uint32_t)w << c) | (ad ? ad - 1 : nb);
            val[half] = (2 * j + half) | (d < 0 ? 0x80000000u : 0u);
        }
        *reinterpret_cast<uint2 *>(keys + (size_t)w * n2 + 2 * j) = make_uint2(key[0], key[1]);
        *reinterpret_cast<uint2 *>(vals + (size_t)w * n2 + 2 * j) = make_uint2(val[0], val[1]);
    }
}

// start[w][b] = first position (inside window w's sorted segment) whose key is >= b, for b = 0..nb (inclusive): one thread per (window, b),
// a binary search in the sorted keys (neighbouring threads walk almost the same path, so the probes coalesce).  Scanning the keys for
// boundaries instead leaves long runs of empty buckets -- the top window uses 2^14 of its 2^18 -- to single threads: 1 ms of serial stores.
__global__ void __launch_bounds__(256) k_big_offsets(const uint32_t *__restrict__ keys_sorted, uint32_t n2, int nwin, uint32_t nb, int c,
                                                     uint32_t *__restrict__ start) {
    const uint32_t w = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    const uint32_t mask = (1u << c) - 1;
    const uint32_t *k = keys_sorted + (size_t)w * n2;
    uint32_t lo = 0, hi = n2;  // first i with key(i) >= b
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((k[mid] & mask) >= b) hi = mid;
        else lo = mid + 1;
    }
    start[(size_t)w * (nb + 1) + b] = lo;
}

// Every window owns nb slots (threads).  In an ordinary window slot = bucket.  The top window only sees the few leftover
// bits (plus the recoding carry), i.e. a handful of very long buckets: there slot = (bucket, s) and the sp_top threads of a
// bucket take every sp_top-th point, producing sp_top partial sums that all carry the bucket's weight.
// A bucket with more than BIG_HEAVY points (skewed scalars: all-equal scalars put ALL points of a window into one bucket, 32-bit scalars put
// half of them into the carry bucket of the third window) is not run by its one thread -- that would be seconds of sequential additions --
// but cut into chunks of BIG_CHUNK points, appended to a device-side work list and summed by k_big_heavy (one CTA per chunk) and
// k_big_heavy_fold (one warp per heavy bucket).  Uniform scalars never take this path: buckets hold ~2n / 2^(c-1) <= 256 points.
constexpr uint32_t BIG_HEAVY = 2048, BIG_CHUNK = 4096;
struct big_heavy_item_t {
    uint32_t slot, lo, hi, first;  // bucket slot (index into buckets_jac), range in the window's sorted array, first item of this bucket in the list
};
__global__ void __launch_bounds__(128, 3) k_big_accumulate(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ vals_sorted,
                                                        const uint32_t *__restrict__ start, uint32_t n2, int nwin, uint32_t nb, uint32_t sp_top,
                                                        const uint32_t *__restrict__ order, uint32_t *__restrict__ buckets_jac, uint32_t *__restrict__ heavy_count,
                                                        big_heavy_item_t *__restrict__ heavy_items, uint32_t heavy_cap) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nwin * nb) return;
    // slots are handed out in descending order of their load (k_big_loads + a 12-bit radix sort): the 32 lanes of a warp then run the same
    // number of additions (Poisson bucket sizes left 24 of 32 lanes active on average), and the longest buckets start first
    t = order[t];
    uint32_t w = (uint32_t)(t / nb), slot = (uint32_t)(t % nb);
    const bool top = (int)w == nwin - 1;
    const uint32_t nbt = nb / sp_top;  // buckets of the top window; slot = r * nbt + b there (b fastest, so that the fold is a strided sum)
    const uint32_t b = top ? slot % nbt : slot, stride = top ? sp_top : 1u, first = top ? slot / nbt : 0u;
    const uint32_t *st = start + (size_t)w * (nb + 1);
    uint32_t lo = st[b] + first, hi = st[b + 1];
    const uint32_t *v = vals_sorted + (size_t)w * n2;
    g1j acc;
    g1j_set_inf(acc);
    // heavy bucket (per-thread share above BIG_HEAVY; in the top window a bucket is already spread over sp_top threads): the thread of
    // its first slot appends the bucket's chunks to the work list, every thread of the bucket leaves infinity in its slot
    const uint32_t blo = st[b], cnt = hi > blo ? hi - blo : 0u;
    if (cnt / stride > BIG_HEAVY) {
        const uint32_t nch = (cnt + BIG_CHUNK - 1) / BIG_CHUNK;
        bool listed = true;
        if (first == 0) {
            const uint32_t base = atomicAdd(heavy_count, nch);
            listed = base + nch <= heavy_cap;  // always true: the list is sized for the worst case
            if (listed)
                for (uint32_t j = 0; j < nch; j++) {
                    const uint32_t clo = blo + j * BIG_CHUNK, chi = clo + BIG_CHUNK < hi ? clo + BIG_CHUNK : hi;
                    heavy_items[base + j] = {(uint32_t)t, clo, chi, j == 0 ? 0x80000000u : 0u};
                }
        }
        if (listed) {
            g1j_store(buckets_jac + 36 * t, acc);  // infinity; k_big_heavy_fold writes the bucket's sum into its first slot
            return;
        }
    }
    g1x ax;  // running sum in XYZZ coordinates (8M + 2S per point), stored as the Jacobian (X ZZ, Y ZZZ, ZZ)
    g1x_set_inf(ax);
#pragma unroll 1
    for (uint32_t e = lo; e < hi; e += stride) {
        uint32_t id = v[e];
        uint32_t p = id & 0x7FFFFFFFu;
        g1a q;
        g1a_load(q, pts + 24 * (size_t)(p >> 1));
        if (p & 1) fp_mul_beta(q.x, q.x);
        if (id & 0x80000000u) fp_neg(q.y, q.y);
        g1x_add_mixed(ax, ax, q);
    }
    g1x_to_jac(acc, ax);
    g1j_store(buckets_jac + 36 * t, acc);
}

// load of every accumulate slot (number of points its thread adds; 0 for a heavy bucket, which only lists its chunks), clamped to 12 bits
__global__ void __launch_bounds__(256) k_big_loads(const uint32_t *__restrict__ start, int nwin, uint32_t nb, uint32_t sp_top, uint32_t *__restrict__ keys,
                                                   uint32_t *__restrict__ vals) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nwin * nb) return;
    const uint32_t w = (uint32_t)(t / nb), slot = (uint32_t)(t % nb);
    const bool top = (int)w == nwin - 1;
    const uint32_t nbt = nb / sp_top;
    const uint32_t b = top ? slot % nbt : slot, stride = top ? sp_top : 1u, first = top ? slot / nbt : 0u;
    const uint32_t *st = start + (size_t)w * (nb + 1);
    const uint32_t cnt = st[b + 1] - st[b];
    uint32_t load = cnt > first ? (cnt - first + stride - 1) / stride : 0u;
    if (cnt / stride > BIG_HEAVY) load = 0;
    keys[t] = load > 4095u ? 4095u : load;
    vals[t] = (uint32_t)t;
}

// One CTA per work item (a chunk of <= BIG_CHUNK points of one heavy bucket): the 128 threads stride over the chunk with mixed additions,
// then a shuffle / shared-memory reduction; partial sum to heavy_partial[item].  The grid is sized for the worst case; CTAs beyond the
// list's length leave at once.
__global__ void __launch_bounds__(128, 3) k_big_heavy(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ vals_sorted, uint32_t n2,
                                                   uint32_t nb, const uint32_t *__restrict__ heavy_count,
                                                   const big_heavy_item_t *__restrict__ heavy_items, uint32_t heavy_cap,
                                                   uint32_t *__restrict__ heavy_partial) {
    const uint32_t count = *heavy_count < heavy_cap ? *heavy_count : heavy_cap;
    __shared__ uint32_t sm[4 * 36];
    for (uint32_t item = blockIdx.x; item < count; item += gridDim.x) {
        const big_heavy_item_t it = heavy_items[item];
        const uint32_t w = it.slot / nb;
        const uint32_t *v = vals_sorted + (size_t)w * n2;
        g1j acc;
        g1x ax;
        g1x_set_inf(ax);
#pragma unroll 1
        for (uint32_t e = it.lo + threadIdx.x; e < it.hi; e += blockDim.x) {
            const uint32_t id = v[e], p = id & 0x7FFFFFFFu;
            g1a q;
            g1a_load(q, pts + 24 * (size_t)(p >> 1));
            if (p & 1) fp_mul_beta(q.x, q.x);
            if (id & 0x80000000u) fp_neg(q.y, q.y);
            g1x_add_mixed(ax, ax, q);
        }
        g1x_to_jac(acc, ax);
#pragma unroll 1
        for (int d = 16; d >= 1; d >>= 1) {
            g1j o;
            shfl_down_g1j(o, acc, d, 32);
            g1j_add(acc, acc, o);
        }
        __syncthreads();  // the previous item's readers are done with sm
        if ((threadIdx.x & 31) == 0) g1j_store(sm + 36 * (threadIdx.x >> 5), acc);
        __syncthreads();
        if (threadIdx.x == 0) {
            g1j_load(acc, sm);
#pragma unroll 1
            for (int k = 1; k < 4; k++) {
                g1j o;
                g1j_load(o, sm + 36 * k);
                g1j_add(acc, acc, o);
            }
            g1j_store(heavy_partial + 36 * (size_t)item, acc);
        }
    }
}
// One warp per work item that is the FIRST chunk of its bucket: sums the bucket's partial sums (its items are contiguous in the list: the
// bucket's chunk count follows from its point range) into the bucket's slot.
__global__ void __launch_bounds__(128) k_big_heavy_fold(const uint32_t *__restrict__ heavy_count, const big_heavy_item_t *__restrict__ heavy_items,
                                                        uint32_t heavy_cap, const uint32_t *__restrict__ start, uint32_t nb,
                                                        const uint32_t *__restrict__ heavy_partial, uint32_t *__restrict__ buckets_jac) {
    const uint32_t count = *heavy_count < heavy_cap ? *heavy_count : heavy_cap;
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (uint32_t item = blockIdx.x * wpb + (threadIdx.x >> 5); item < count; item += gridDim.x * wpb) {
        const big_heavy_item_t it = heavy_items[item];
        if (!(it.first & 0x80000000u)) continue;  // warp-uniform
        const uint32_t w = it.slot / nb, b = it.slot % nb;
        const uint32_t *st = start + (size_t)w * (nb + 1);
        const uint32_t nch = (st[b + 1] - st[b] + BIG_CHUNK - 1) / BIG_CHUNK;
        g1j acc;
        g1j_set_inf(acc);
#pragma unroll 1
        for (uint32_t j = lane; j < ((nch + 31) & ~31u); j += 32) {
            g1j q;
            g1j_set_inf(q);
            if (j < nch) g1j_load(q, heavy_partial + 36 * (size_t)(item + j));
            g1j_add(acc, acc, q);
        }
#pragma unroll 1
        for (int d = 16; d >= 1; d >>= 1) {
            g1j o;
            shfl_down_g1j(o, acc, d, 32);
            g1j_add(acc, acc, o);
        }
        if (lane == 0) g1j_store(buckets_jac + 36 * (size_t)it.slot, acc);
    }
}

// One level of the hierarchical bucket reduction.  Every node carries A = sum of its buckets and Bv = sum of (local index) * bucket.
// A parent of g children (child k spans 2^shift buckets):  A = sum_k A_k ,  Bv = sum_k Bv_k + 2^shift * sum_k k * A_k,
// the last sum by the running-sum trick.  One thread per parent; Bin == nullptr at the leaves (Bv = 0).
__global__ void __launch_bounds__(128) k_big_reduce_level(const uint32_t *__restrict__ Ain, const uint32_t *__restrict__ Bin, uint32_t n_out,
                                                          uint32_t g, int shift, uint32_t *__restrict__ Aout, uint32_t *__restrict__ Bout) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_out) return;
    g1j run, wsum, bsum;
    g1j_set_inf(run);
    g1j_set_inf(wsum);
    g1j_set_inf(bsum);
#pragma unroll 1
    for (int k = (int)g - 1; k >= 0; k--) {
        size_t idx = (size_t)q * g + k;
        g1j a;
        g1j_load(a, Ain + 36 * idx);
        g1j_add(run, run, a);
        // second addition of the step: the running sum into the weighted sum, or (last step / leaves) nothing
        if (k >= 1) g1j_add(wsum, wsum, run);
    }
    if (Bin) {
#pragma unroll 1
        for (uint32_t k = 0; k < g; k++) {
            g1j bb;
            g1j_load(bb, Bin + 36 * ((size_t)q * g + k));
            g1j_add(bsum, bsum, bb);
        }
    }
#pragma unroll 1
    for (int s2 = 0; s2 < shift; s2++) g1j_dbl(wsum, wsum);
    g1j_add(bsum, bsum, wsum);
    g1j_store(Aout + 36 * (size_t)q, run);
    g1j_store(Bout + 36 * (size_t)q, bsum);
}

// First level over AFFINE bucket sums (k_batchaff.cu leaves them so; x = y = 0 is the empty bucket): the running sum takes mixed additions.
// QUAD: a quad of lanes per node, as below (levels of few nodes).
template <bool QUAD>
__global__ void __launch_bounds__(128) k_big_reduce_leaf_affine(const uint32_t *__restrict__ bucket_aff, uint32_t n_out, uint32_t g,
                                                                uint32_t *__restrict__ Aout, uint32_t *__restrict__ Bout) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, q = QUAD ? gid >> 2 : gid;
    const int s = (int)(threadIdx.x & 3);
    const bool valid = q < n_out;
    if (!QUAD && !valid) return;
    g1j run, wsum;
    g1j_set_inf(run);
    g1j_set_inf(wsum);
#pragma unroll 1
    for (int k = (int)g - 1; k >= 0; k--) {
        g1a a;
        g1a_set_inf(a);
        if (valid) g1a_load(a, bucket_aff + 24 * ((size_t)q * g + k));
        if (QUAD) {
            g1j aj;
            g1j_from_affine(aj, a);
            g1j_add_quad(run, run, aj, s);
            if (k >= 1) g1j_add_quad(wsum, wsum, run, s);
        } else {
            g1j_add_mixed(run, run, a);
            if (k >= 1) g1j_add(wsum, wsum, run);
        }
    }
    if (valid && (!QUAD || s == 0)) {
        g1j_store(Aout + 36 * (size_t)q, run);
        g1j_store(Bout + 36 * (size_t)q, wsum);
    }
}

// k_big_reduce_level with a quad of lanes per output node (g1_quad.cuh): the 2 g - 1 dependent full additions of a node run at 5 product
// latencies each instead of 16.  For the upper levels of the hierarchy, which are a handful of threads deep in a latency chain.
__global__ void __launch_bounds__(128) k_big_reduce_level_quad(const uint32_t *__restrict__ Ain, const uint32_t *__restrict__ Bin, uint32_t n_out,
                                                               uint32_t g, int shift, uint32_t *__restrict__ Aout, uint32_t *__restrict__ Bout) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, q = gid >> 2;
    const int s = (int)(threadIdx.x & 3);
    const bool valid = q < n_out;  // idle quads of the last warp run along with points at infinity (full-mask shuffles inside)
    g1j run, wsum, bsum;
    g1j_set_inf(run);
    g1j_set_inf(wsum);
    g1j_set_inf(bsum);
#pragma unroll 1
    for (int k = (int)g - 1; k >= 0; k--) {
        g1j a;
        g1j_set_inf(a);
        if (valid) g1j_load(a, Ain + 36 * ((size_t)q * g + k));
        g1j_add_quad(run, run, a, s);
        if (k >= 1) g1j_add_quad(wsum, wsum, run, s);
    }
    if (Bin) {
#pragma unroll 1
        for (uint32_t k = 0; k < g; k++) {
            g1j bb;
            g1j_set_inf(bb);
            if (valid) g1j_load(bb, Bin + 36 * ((size_t)q * g + k));
            g1j_add_quad(bsum, bsum, bb, s);
        }
    }
#pragma unroll 1
    for (int s2 = 0; s2 < shift; s2++) g1j_dbl_quad(wsum, s);
    g1j_add_quad(bsum, bsum, wsum, s);
    if (valid && s == 0) {
        g1j_store(Aout + 36 * (size_t)q, run);
        g1j_store(Bout + 36 * (size_t)q, bsum);
    }
}

// total = sum_w 2^(c w) * (Bv_w + A_w)   (weight of bucket b is b + 1); one thread, Horner from the top window down
__global__ void __launch_bounds__(32) k_big_horner(const uint32_t *__restrict__ A, const uint32_t *__restrict__ Bv, int nwin, int c, uint32_t *__restrict__ out_jac) {
    // one warp; its eight quads all run the same chain (g1_quad.cuh: 3 product latencies per doubling instead of 7), quad 0 stores
    const int s = (int)(threadIdx.x & 3);
    g1j total;
    g1j_set_inf(total);
#pragma unroll 1
    for (int w = nwin - 1; w >= 0; w--) {
        if (w != nwin - 1) {
#pragma unroll 1
            for (int k = 0; k < c; k++) g1j_dbl_quad(total, s);
        }
        g1j a, b;
        g1j_load(a, A + 36 * (size_t)w);
        g1j_load(b, Bv + 36 * (size_t)w);
        g1j_add_quad(total, total, a, s);
        g1j_add_quad(total, total, b, s);
    }
    if (threadIdx.x == 0) g1j_store(out_jac, total);
}

// ---- the upper part of the bucket reduction as PLAIN sums.  After the 16-ary levels above a window is N nodes (A_k, Bv_k), k < N, and its
// total is  T = sum_k (A_k + Bv_k) + 2^shift * sum_k k A_k .  Writing k in binary,  sum_k k A_k = sum_j 2^j S_j  with  S_j = sum over the k
// that have bit j set of A_k : log2 N + 1 plain sums per window, each a tree -- ~14 dependent quad additions instead of the ~50 per level
// (x 3 levels) of the running-sum hierarchy, whose last levels are a handful of threads deep in a latency chain.
//   k_big_bitsums   one CTA of 64 quads per (window, sum): strided partial sums, folded by shuffles and shared memory
//   k_big_horner2   one warp per window: T_w by Horner over the bits; then warp 0: Horner over the windows
constexpr int BITSUM_PARTS = 4;  // CTAs per sum (4 warps each: one per scheduler -- more warps on an SM only share its multiply pipe)
__global__ void __launch_bounds__(128) k_big_bitsums(const uint32_t *__restrict__ A, const uint32_t *__restrict__ Bv, uint32_t N, int nbits,
                                                     uint32_t *__restrict__ sums /* [nwin][nbits + 1][BITSUM_PARTS] */) {
    __shared__ uint32_t sm[4 * 36];
    const int sidx = blockIdx.x, w = blockIdx.y, part = blockIdx.z;
    const int s = (int)(threadIdx.x & 3);
    const uint32_t quad = (threadIdx.x >> 2) + 32 * part, nquads = 32 * BITSUM_PARTS;
    const uint32_t *Aw = A + 36 * (size_t)w * N, *Bw = Bv + 36 * (size_t)w * N;
    const uint32_t count = sidx < nbits ? N / 2 : 2 * N;
    g1j acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (uint32_t e0 = 0; e0 < count; e0 += nquads) {
        const uint32_t e = e0 + quad;
        g1j q;
        g1j_set_inf(q);
        if (e < count) {
            if (sidx < nbits) {  // the e-th index with bit sidx set
                const uint32_t k = ((e >> sidx) << (sidx + 1)) | (1u << sidx) | (e & ((1u << sidx) - 1u));
                g1j_load(q, Aw + 36 * (size_t)k);
            } else {
                g1j_load(q, e < N ? Aw + 36 * (size_t)e : Bw + 36 * (size_t)(e - N));
            }
        }
        g1j_add_quad(acc, acc, q, s);
    }
#pragma unroll 1
    for (int d = 4; d < 32; d <<= 1) {  // the 8 quads of a warp
        g1j o;
        shfl_down_g1j(o, acc, d, 32);
        g1j_add_quad(acc, acc, o, s);
    }
    if ((threadIdx.x & 31) == 0) g1j_store(sm + 36 * (threadIdx.x >> 5), acc);
    __syncthreads();
    if (threadIdx.x < 32) {  // the 4 warps: quads 0..3 hold them, the others infinity
        g1j_set_inf(acc);
        if (threadIdx.x < 16) g1j_load(acc, sm + 36 * (threadIdx.x >> 2));
#pragma unroll 1
        for (int d = 4; d < 16; d <<= 1) {
            g1j o;
            shfl_down_g1j(o, acc, d, 32);
            g1j_add_quad(acc, acc, o, s);
        }
        if (threadIdx.x == 0) g1j_store(sums + 36 * (((size_t)w * (nbits + 1) + sidx) * BITSUM_PARTS + part), acc);
    }
}
__global__ void __launch_bounds__(512) k_big_horner2(const uint32_t *__restrict__ sums, int nbits, int shift, int nwin, int c, uint32_t *__restrict__ out_jac) {
    __shared__ uint32_t sm[16 * 36];
    extern __shared__ uint32_t smj[];  // [nwin][nbits + 1] merged sums
    const int s = (int)(threadIdx.x & 3), w = (int)(threadIdx.x >> 5), quad = (int)((threadIdx.x >> 2) & 7);
    {
        // merge the parts of every sum: quad q of warp w takes sums q, q + 8, ... of window w (every quad runs the same number of steps)
        const uint32_t *S = sums + 36 * (size_t)w * (nbits + 1) * BITSUM_PARTS;
        uint32_t *M = smj + 36 * (size_t)w * (nbits + 1);
#pragma unroll 1
        for (int j0 = 0; j0 <= nbits; j0 += 8) {
            const int j = j0 + quad;
            g1j t;
            g1j_set_inf(t);
#pragma unroll 1
            for (int pt = 0; pt < BITSUM_PARTS; pt++) {
                g1j q;
                g1j_set_inf(q);
                if (j <= nbits) g1j_load(q, S + 36 * ((size_t)j * BITSUM_PARTS + pt));
                g1j_add_quad(t, t, q, s);
            }
            if (j <= nbits && s == 0) g1j_store(M + 36 * (size_t)j, t);
        }
        __syncwarp();
        // every quad of warp w runs the same chain
        g1j t;
        g1j_set_inf(t);
#pragma unroll 1
        for (int j = nbits - 1; j >= 0; j--) {
            g1j_dbl_quad(t, s);
            g1j q;
            g1j_load(q, M + 36 * (size_t)j);
            g1j_add_quad(t, t, q, s);
        }
#pragma unroll 1
        for (int k = 0; k < shift; k++) g1j_dbl_quad(t, s);
        g1j u;
        g1j_load(u, M + 36 * (size_t)nbits);
        g1j_add_quad(t, t, u, s);
        if ((threadIdx.x & 31) == 0) g1j_store(sm + 36 * w, t);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        g1j total;
        g1j_set_inf(total);
#pragma unroll 1
        for (int ww = nwin - 1; ww >= 0; ww--) {
            if (ww != nwin - 1) {
#pragma unroll 1
                for (int k = 0; k < c; k++) g1j_dbl_quad(total, s);
            }
            g1j a;
            g1j_load(a, sm + 36 * ww);
            g1j_add_quad(total, total, a, s);
        }
        if (threadIdx.x == 0) g1j_store(out_jac, total);
    }
}

// top window fold, one step: out[j * nbt + b] = sum over r = j, j + nw, j + 2 nw, ... < sp of in[r * nbt + b]   (warp (b, j)).
// Called twice (sp -> nw partial sums per bucket, then nw -> 1); the last call (nw == 1) also pads out[nbt .. nb) with infinity.
__global__ void __launch_bounds__(128) k_big_fold_top(const uint32_t *__restrict__ in, uint32_t nbt, uint32_t sp, uint32_t nw, uint32_t pad_to,
                                                      uint32_t *__restrict__ out) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = nbt * nw > pad_to ? nbt * nw : pad_to;
    if (warp >= n_warps) return;
    g1j acc;
    g1j_set_inf(acc);
    if (warp < nbt * nw) {
        const uint32_t b = warp % nbt, jj = warp / nbt;
        const uint32_t cnt = (sp - jj + nw - 1) / nw;  // elements of this warp
#pragma unroll 1
        for (uint32_t it = lane; it < ((cnt + 31) & ~31u); it += 32) {
            g1j q;
            g1j_set_inf(q);
            if (it < cnt) g1j_load(q, in + 36 * ((size_t)(jj + (size_t)it * nw) * nbt + b));
            g1j_add(acc, acc, q);
        }
#pragma unroll 1
        for (int d = 16; d >= 1; d >>= 1) {
            g1j o;
            shfl_down_g1j(o, acc, d, 32);
            g1j_add(acc, acc, o);
        }
    }
    if (lane == 0) g1j_store(out + 36 * (size_t)warp, acc);
}

size_t big_msm_sort_temp_bytes(uint32_t n2, int nwin, int c) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (size_t)n2 * nwin, 0, c + 4);
    return bytes;
}

cudaError_t launch_big_digits(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, uint32_t n, int c, int nwin, uint32_t *keys,
                              uint32_t *vals, uint32_t *bx) {
    k_big_digits<<<(n + 255) / 256, 256, 0, st>>>(pts, scalars, n, c, nwin, keys, vals, bx);
    return cudaGetLastError();
}
cudaError_t launch_big_sort(cudaStream_t st, void *temp, size_t temp_bytes, const uint32_t *keys, uint32_t *keys_out, const uint32_t *vals,
                            uint32_t *vals_out, uint32_t n2, int nwin, int c, const uint32_t *seg_offsets /* nwin + 1 */) {
    (void)seg_offsets;
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, vals, vals_out, (size_t)n2 * nwin, 0, c + 4, st);
}
cudaError_t launch_big_offsets(cudaStream_t st, const uint32_t *keys_sorted, uint32_t n2, int nwin, uint32_t nb, int c, uint32_t *start) {
    k_big_offsets<<<dim3((nb + 1 + 255) / 256, (unsigned)nwin), 256, 0, st>>>(keys_sorted, n2, nwin, nb, c, start);
    return cudaGetLastError();
}
size_t big_order_temp_bytes(size_t slots) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                              slots, 0, 12);
    return bytes;
}
// order_ws: 4 arrays of `slots` words (keys, values, sorted keys, sorted values = the order), then the sort's temporary storage
cudaError_t launch_big_order(cudaStream_t st, const uint32_t *start, int nwin, uint32_t nb, uint32_t sp_top, uint32_t *order_ws, size_t temp_bytes) {
    const size_t slots = (size_t)nwin * nb;
    uint32_t *keys = order_ws, *vals = order_ws + slots, *keys2 = order_ws + 2 * slots, *vals2 = order_ws + 3 * slots;
    k_big_loads<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(start, nwin, nb, sp_top, keys, vals);
    return cub::DeviceRadixSort::SortPairsDescending(order_ws + 4 * slots, temp_bytes, keys, keys2, vals, vals2, slots, 0, 12, st);
}
size_t big_heavy_capacity(uint32_t n2, int nwin) {  // worst case: every bucket's last chunk is partial, at most items / BIG_HEAVY heavy buckets
    const size_t items = (size_t)n2 * nwin;
    return items / BIG_CHUNK + items / BIG_HEAVY + 16;
}
size_t big_heavy_bytes(uint32_t n2, int nwin) {  // counter (256 B) + work list + partial sums
    const size_t cap = big_heavy_capacity(n2, nwin);
    return 256 + ((cap * sizeof(big_heavy_item_t) + 255) & ~size_t(255)) + cap * 144;
}
cudaError_t launch_big_accumulate(cudaStream_t st, const uint32_t *pts, const uint32_t *vals_sorted, const uint32_t *start, uint32_t n2, int nwin,
                                  uint32_t nb, uint32_t sp_top, const uint32_t *order, uint32_t *buckets_jac, void *heavy_ws) {
    size_t total = (size_t)nwin * nb;
    const size_t cap = big_heavy_capacity(n2, nwin);
    uint32_t *count = reinterpret_cast<uint32_t *>(heavy_ws);
    big_heavy_item_t *items = reinterpret_cast<big_heavy_item_t *>(reinterpret_cast<uint8_t *>(heavy_ws) + 256);
    uint32_t *partial = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(heavy_ws) + 256 + ((cap * sizeof(big_heavy_item_t) + 255) & ~size_t(255)));
    cudaError_t e = cudaMemsetAsync(count, 0, 4, st);
    if (e != cudaSuccess) return e;
    k_big_accumulate<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(pts, vals_sorted, start, n2, nwin, nb, sp_top, order, buckets_jac, count, items,
                                                                      (uint32_t)cap);
    // skewed inputs only: with uniform scalars the list is empty and both kernels return at once (grid-stride over the list)
    const unsigned grid = (unsigned)(cap < 148 * 12 ? cap : 148 * 12);
    k_big_heavy<<<grid, 128, 0, st>>>(pts, vals_sorted, n2, nb, count, items, (uint32_t)cap, partial);
    k_big_heavy_fold<<<148, 128, 0, st>>>(count, items, (uint32_t)cap, start, nb, partial, buckets_jac);
    return cudaGetLastError();
}
cudaError_t launch_big_reduce_level(cudaStream_t st, const uint32_t *Ain, const uint32_t *Bin, uint32_t n_out, uint32_t g, int shift, uint32_t *Aout,
                                    uint32_t *Bout) {
    // a level of few nodes is a latency chain: a quad per node; the wide first levels keep one thread per node (same time, a quarter of the lanes)
    if (n_out <= 32768) k_big_reduce_level_quad<<<(unsigned)(((size_t)n_out * 4 + 127) / 128), 128, 0, st>>>(Ain, Bin, n_out, g, shift, Aout, Bout);
    else k_big_reduce_level<<<(n_out + 127) / 128, 128, 0, st>>>(Ain, Bin, n_out, g, shift, Aout, Bout);
    return cudaGetLastError();
}
cudaError_t launch_big_reduce_leaf_affine(cudaStream_t st, const uint32_t *bucket_aff, uint32_t n_out, uint32_t g, uint32_t *Aout, uint32_t *Bout) {
    if (n_out <= 32768) k_big_reduce_leaf_affine<true><<<(unsigned)(((size_t)n_out * 4 + 127) / 128), 128, 0, st>>>(bucket_aff, n_out, g, Aout, Bout);
    else k_big_reduce_leaf_affine<false><<<(n_out + 127) / 128, 128, 0, st>>>(bucket_aff, n_out, g, Aout, Bout);
    return cudaGetLastError();
}
cudaError_t launch_big_horner(cudaStream_t st, const uint32_t *A, const uint32_t *Bv, int nwin, int c, uint32_t *out_jac) {
    k_big_horner<<<1, 32, 0, st>>>(A, Bv, nwin, c, out_jac);
    return cudaGetLastError();
}
// window totals and the final Horner from N nodes per window (N a power of two >= 2, nwin <= 16); sums: 4 nwin (log2 N + 1) Jacobian points of scratch
cudaError_t launch_big_bitsum_horner(cudaStream_t st, const uint32_t *A, const uint32_t *Bv, uint32_t N, int shift, int nwin, int c, uint32_t *sums,
                                     uint32_t *out_jac) {
    int nbits = 0;
    while ((1u << nbits) < N) nbits++;
    k_big_bitsums<<<dim3((unsigned)nbits + 1, (unsigned)nwin, BITSUM_PARTS), 128, 0, st>>>(A, Bv, N, nbits, sums);
    k_big_horner2<<<1, 32 * nwin, (size_t)nwin * (nbits + 1) * 144, st>>>(sums, nbits, shift, nwin, c, out_jac);
    return cudaGetLastError();
}
cudaError_t launch_big_fold_top(cudaStream_t st, const uint32_t *in, uint32_t nbt, uint32_t sp, uint32_t nw, uint32_t pad_to, uint32_t *out) {
    uint32_t n_warps = nbt * nw > pad_to ? nbt * nw : pad_to;
    k_big_fold_top<<<(n_warps * 32 + 127) / 128, 128, 0, st>>>(in, nbt, sp, nw, pad_to, out);
    return cudaGetLastError();
}
}  // namespace cdp
