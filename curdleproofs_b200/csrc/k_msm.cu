// Compiled once per window width: -DMSM_C=2..6 (see Makefile); each object exports launch_msm_buckets_c<MSM_C>.
#include "launch.h"
#include "msm_common.cuh"

#ifndef MSM_C
#error "compile with -DMSM_C=<2..6>"
#endif

namespace cdp {

// The small-MSM path -- `util::msm` (/root/reference/src/util.rs:19-22) for a batch of independent MSMs -- runs in three kernels:
//
// k_msm_digits   one thread per (segment, pair): coalesced 128-bit scalar loads, GLV split (2n half-width "points": 2j -> P_j
//                with k1, 2j+1 -> phi(P_j) with k2), signed radix-2^c recoding; digit rows (int8, one row per window) go to global
//                memory once per MSM.
// k_msm_buckets  one WARP per (MSM, group of 64 / 2^(c-1) windows); warps are independent (per-warp shared memory, __syncwarp only),
//                so every MSM uses exactly ceil(NWIN / WPW) warps whatever the CTA shape.  Lane (w, b) owns bucket b of window w:
//   phase 0  the warp's digit rows are staged global -> shared memory with 16-byte loads
//   phase 1-2  counting sort of the point ids (uint16 id | sign bit) into one list per (window, bucket) slot, 64 slots per warp
//   phase 3  the slots are ranked by load; lane l takes the l-th largest and the l-th smallest
//   phase 4  each lane adds the points of its two slots (mixed adds) and writes the sums to [msm][window][slot]
// k_msm_combine  (k_misc.cu) one thread per (MSM, bucket index): Horner over the windows FIRST (sum_w 2^(cw) B_{w,b}: the doublings run
//                on 2^(c-1) lanes in parallel), THEN one weighted reduction sum_b (b+1) S_b per MSM -- instead of one reduction per window.
// Handles infinity bases and zero scalars (digits 0) and all-equal scalars (one long list, still correct).
template <int C>
__global__ void __launch_bounds__(128) k_msm_digits(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ scalars,
                                                    const msm_seg_t *__restrict__ segs, int8_t *__restrict__ dig, uint32_t rowstride,
                                                    uint32_t *__restrict__ bx, uint32_t nmax) {
    constexpr int NB = 1 << (C - 1);
    constexpr int NWIN = (130 + C - 1) / C;
    const msm_seg_t seg = segs[blockIdx.x];
    const uint32_t n_plain = seg.n;
    const uint32_t n = seg.n + (seg.extra ? 1u : 0u);  // the optional extra base is logically element n_plain
    const uint32_t *PX = seg.extra ? pts + 24 * (size_t)(seg.extra - 1) : nullptr;
    const uint32_t *P = pts + 24 * (size_t)seg.pts_off;
    const uint32_t *S = scalars + 8 * (size_t)seg.scalars_off;
    int8_t *digits = dig + (size_t)blockIdx.x * NWIN * rowstride;
    for (uint32_t j = blockIdx.y * blockDim.x + threadIdx.x; j < n; j += blockDim.x * gridDim.y) {
        uint32_t k[8];
        const uint4 *sp = reinterpret_cast<const uint4 *>(S + 8 * (size_t)j);
        uint4 a = sp[0], b = sp[1];
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
        // infinity base: contributes nothing
        const uint4 *pp = reinterpret_cast<const uint4 *>(j < n_plain ? P + 24 * (size_t)j : PX);
        uint32_t nz = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            uint4 v = pp[q];
            nz |= v.x | v.y | v.z | v.w;
        }
        {   // x-coordinate of phi(P) = (beta x, y), once per base instead of once per bucket addition
            fp x;
            fp_load(x, reinterpret_cast<const uint32_t *>(pp));
            fp_mul_beta(x, x);
            fp_store(bx + 12 * ((size_t)blockIdx.x * nmax + j), x);
        }
        glv_t g;
        glv_split(g, k);
        if (nz == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) g.k1[q] = g.k2[q] = 0;
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t carry = 0;
#pragma unroll
            for (int w = 0; w < NWIN; w++) {
                const int bit = w * C, li = bit >> 5, sh = bit & 31;
                uint32_t v = 0;
                if (li < 4) v = (half ? g.k2[li] : g.k1[li]) >> sh;
                if (sh + C > 32 && li + 1 < 4) v |= (half ? g.k2[li + 1] : g.k1[li + 1]) << (32 - sh);
                v = (v & ((1u << C) - 1)) + carry;
                carry = (v + NB) >> C;  // recentre to [-NB, NB)
                int d = (int)v - (int)(carry << C);
                digits[(size_t)w * rowstride + 2 * j + half] = (int8_t)d;
            }
        }
    }
}

// dynamic shared memory, per warp: int8 digits[WPW][dstride] ; uint16 lists[WPW][2*nmax] ; uint32 cnt[64], cur[64] ; uint8 order[64]
template <int C, int OCC>
__global__ void __launch_bounds__(128, OCC)
    k_msm_buckets(const uint32_t *__restrict__ pts, const msm_seg_t *__restrict__ segs, const int8_t *__restrict__ dig, uint32_t rowstride,
                  uint32_t *__restrict__ bucket_sums /* [msm][nwin][NB] jacobian */, uint32_t nmax, uint32_t n_msm,
                  const uint32_t *__restrict__ bx /* [msm][nmax] beta * x */) {
    constexpr int NB = 1 << (C - 1);
    constexpr int NWIN = (130 + C - 1) / C;
    constexpr int SLOTS = 64;                       // (window, bucket) slots per warp: two per lane
    constexpr int WPW = SLOTS / NB;                 // windows per warp
    constexpr int G = (NWIN + WPW - 1) / WPW;       // warps per MSM
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane_id = threadIdx.x & 31;
    const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (unit >= n_msm * G) return;  // whole warps leave; nothing below synchronises across warps
    const uint32_t msm = unit / G;
    const int w0 = (int)(unit - msm * G) * WPW;
    const msm_seg_t seg = segs[msm];
    const uint32_t n_plain = seg.n;
    const uint32_t n = seg.n + (seg.extra ? 1u : 0u), n2 = 2 * n;
    const uint32_t *PX = seg.extra ? pts + 24 * (size_t)(seg.extra - 1) : nullptr;
    const uint32_t dstride = ((2 * nmax + 15) & ~15u) + 16;  // 16-byte aligned rows; the extra 16 bytes spread the rows over the banks
    const size_t per_warp = msm_smem_per_warp(C, nmax);
    int8_t *digits = reinterpret_cast<int8_t *>(smem + (threadIdx.x >> 5) * per_warp);
    uint16_t *lists = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(digits) + (size_t)WPW * dstride);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(lists) + (size_t)WPW * 2 * nmax * sizeof(uint16_t));
    uint32_t *cur = cnt + SLOTS;
    uint8_t *order = reinterpret_cast<uint8_t *>(cur + SLOTS);
    const uint32_t *P = pts + 24 * (size_t)seg.pts_off;

    // ---- phase 0: stage this warp's digit rows
    {
        const int8_t *gd = dig + (size_t)msm * NWIN * rowstride;
        const uint32_t vec_per_row = (n2 + 15) / 16;
        for (uint32_t t = lane_id; t < WPW * vec_per_row; t += 32) {
            uint32_t wl = t / vec_per_row, v = t % vec_per_row;
            if (w0 + (int)wl < NWIN)
                reinterpret_cast<uint4 *>(digits + (size_t)wl * dstride)[v] =
                    reinterpret_cast<const uint4 *>(gd + (size_t)(w0 + wl) * rowstride)[v];
        }
        cnt[lane_id] = 0;
        cnt[lane_id + 32] = 0;
    }
    __syncwarp();
    // The top window only holds the few leftover bits (plus the recoding carry): NBT distinct non-zero digits.  One list per digit
    // would leave NB - NBT slots empty and put 2n/NBT points into each of the others, so in the top window every bucket is spread over
    // SP = NB / NBT slots (by point index); k_msm_combine adds the SP partial sums of a bucket before it starts.
    constexpr int TB = 128 - C * (NWIN - 1);
    constexpr int NBT = TB > 0 ? (1 << TB) : 1;
    constexpr int SP = NB / NBT >= 1 ? NB / NBT : 1;
    static_assert(NBT <= NB, "top-window digits must fit the slots");
    // ---- phases 1-2: counting sort of the point ids by slot, the whole warp working on one window row at a time: every lane takes
    //      every 32nd digit (shared-memory atomics on the row's NB counters).  List order inside a slot is arbitrary; its SUM does not
    //      depend on it.
#pragma unroll 1
    for (int r = 0; r < WPW; r++) {
        if (w0 + r >= NWIN) break;
        const bool row_top = (w0 + r == NWIN - 1);
        const int8_t *row = digits + r * dstride;
        for (uint32_t p = lane_id; p < n2; p += 32) {
            int d = row[p];
            int ad = d < 0 ? -d : d;
            if (ad) atomicAdd(&cnt[r * NB + (row_top ? (ad - 1) * SP + (int)((p >> 1) & (SP - 1)) : ad - 1)], 1u);
        }
    }
    __syncwarp();
    // list offsets: exclusive scan of the counts inside every window row (rows are NB-aligned groups of slots; NB <= 32)
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int sl = lane_id + 32 * half;
        const uint32_t c0 = cnt[sl];
        uint32_t off = c0;
#pragma unroll
        for (int d = 1; d < NB; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, off, d, NB);
            if ((sl & (NB - 1)) >= d) off += o;
        }
        cur[sl] = off - c0;
    }
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < WPW; r++) {
        if (w0 + r >= NWIN) break;
        const bool row_top = (w0 + r == NWIN - 1);
        const int8_t *row = digits + r * dstride;
        uint16_t *lst = lists + (size_t)r * (2 * nmax);
        // plain points (even ids) first, phi-points (odd ids) second: in phase 4 the lanes of a warp then reach the extra
        // multiplication by beta at about the same step instead of diverging on it at every step
#pragma unroll 1
        for (uint32_t par = 0; par < 2; par++) {
            for (uint32_t p = 2 * lane_id + par; p < n2; p += 64) {
                int d = row[p];
                int ad = d < 0 ? -d : d;
                if (ad) {
                    uint32_t pos = atomicAdd(&cur[r * NB + (row_top ? (ad - 1) * SP + (int)((p >> 1) & (SP - 1)) : ad - 1)], 1u);
                    lst[pos] = (uint16_t)(p | (d < 0 ? 0x8000u : 0u));
                }
            }
            __syncwarp();
        }
    }
    // ---- phase 3: balance.  Slot loads are binomial (+-35% around the mean at these sizes) and a warp runs as long as its busiest lane,
    //      so the 64 slots are ranked by load and lane l takes the l-th largest and the l-th smallest: every lane ends up within a few
    //      points of two mean loads (measured before: 21 of 32 lanes active on average, profiles/r01_ncu_msm_buckets_v3.txt).
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int sl = lane_id + 32 * half;
        const uint32_t c0 = cnt[sl];
        int rank = 0;
#pragma unroll 8
        for (int t = 0; t < SLOTS; t++) {
            const uint32_t ct = cnt[t];
            rank += (ct > c0 || (ct == c0 && t < sl)) ? 1 : 0;
        }
        order[rank] = (uint8_t)sl;
    }
    __syncwarp();
    // ---- phase 4: every lane sums its two slots (different lanes, different points => no serialisation; mixed additions) and writes
    //      each sum to its place in [msm][window][slot].  ONE loop over both lists -- the warp runs max_l (load_A + load_B) steps, not
    //      max_l load_A + max_l load_B -- with the accumulator flushed where the first list ends.
    const int slA = order[lane_id], slB = order[SLOTS - 1 - lane_id];
    const int rA = slA / NB, rB = slB / NB;
    const bool okA = w0 + rA < NWIN, okB = w0 + rB < NWIN;
    const uint32_t cA = okA ? cnt[slA] : 0u, cB = okB ? cnt[slB] : 0u;
    const uint16_t *lstA = lists + (size_t)rA * (2 * nmax) + (cur[slA] - cnt[slA]);
    const uint16_t *lstB = lists + (size_t)rB * (2 * nmax) + (cur[slB] - cnt[slB]);
    uint32_t *outA = bucket_sums + 36 * (((size_t)msm * NWIN + (w0 + rA)) * NB + (slA & (NB - 1)));
    uint32_t *outB = bucket_sums + 36 * (((size_t)msm * NWIN + (w0 + rB)) * NB + (slB & (NB - 1)));
    // the running sum in XYZZ coordinates (8M + 2S per point instead of 7M + 4S), written as the Jacobian point (X ZZ, Y ZZZ, ZZ)
    g1x acc;
    g1x_set_inf(acc);
    auto flush = [&](uint32_t *dst) {
        g1j j;
        g1x_to_jac(j, acc);
        g1j_store(dst, j);
    };
#pragma unroll 1
    for (uint32_t e = 0; e < cA + cB; e++) {
        if (e == cA) {  // first list done (an empty slot A is written here too, at e = 0, as infinity)
            if (okA) flush(outA);
            g1x_set_inf(acc);
        }
        const uint32_t id = e < cA ? lstA[e] : lstB[e - cA];
        const uint32_t p = id & 0x7FFFu;
        g1a q;
        g1a_load(q, (p >> 1) < n_plain ? P + 24 * (size_t)(p >> 1) : PX);
        if (p & 1) fp_load(q.x, bx + 12 * ((size_t)msm * nmax + (p >> 1)));  // phi(P): beta * x from the digit kernel
        if (id & 0x8000u) fp_neg(q.y, q.y);
        g1x_add_mixed(acc, acc, q);
    }
    if (cB == 0) {  // the loop never crossed into list B: acc is still slot A's sum (or infinity), slot B is empty
        if (okA) flush(outA);
        g1x_set_inf(acc);
        if (okB) flush(outB);
    } else if (okB) {
        flush(outB);
    }
}

#define CDP_CAT2(a, b) a##b
#define CDP_CAT(a, b) CDP_CAT2(a, b)
cudaError_t CDP_CAT(launch_msm_buckets_c, MSM_C)(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, const msm_seg_t *segs,
                                                 uint32_t count, uint32_t nmax, int8_t *dig, uint32_t *bucket_sums) {
    constexpr int C = MSM_C;
    const uint32_t rowstride = msm_dig_rowstride(nmax);
    dim3 dgrid(count, nmax > 512 ? (nmax + 511) / 512 : 1);
    uint32_t *bx = reinterpret_cast<uint32_t *>(dig + msm_dig_rows_bytes(C, nmax, count));
    k_msm_digits<C><<<dgrid, 128, 0, st>>>(pts, scalars, segs, dig, rowstride, bx, nmax);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t smem = msm_smem_bytes(C, nmax);
    const int occ = tuned_occupancy("CDP_OCC_BUCKETS", 3);
    auto kern = occ == 5 ? k_msm_buckets<C, 5> : occ == 4 ? k_msm_buckets<C, 4> : k_msm_buckets<C, 3>;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const uint64_t units = (uint64_t)count * msm_groups_for(C);
    kern<<<(unsigned)((units + 3) / 4), 128, smem, st>>>(pts, segs, dig, rowstride, bucket_sums, nmax, count, bx);
    return cudaGetLastError();
}

}  // namespace cdp
