// Fixed-base MSM: `util::msm` (/root/reference/src/util.rs:19-22) for the call sites whose bases are CRS points
// (crs.vec_G, vec_H, H, G_t, G_u -- /root/reference/src/crs.rs:19-34), which never change between proofs:
//   A, C, B_c, B_d, B_a, the four GroupCommitment T_1 points, and -- once the IPA / SameMSM round MSMs are written over the
//   ORIGINAL bases with the fold challenges moved into the scalars (the reference's own verifier identity,
//   src/inner_product_argument.rs:202-250) -- every L/R cross term over G, G' = u o G and G_with_blinders.
//
// B200-first layout: with 180 GB of HBM per GPU the whole digit table fits on the device,
//   table[(base * nw + w) * nd + (d - 1)] = d * 2^(c w) * B_base   (affine, 96 B; c = 16: 16 windows x 32768 digits = 50 MB per base;
//   entries 96 bytes apart, or 128 -- one DRAM line per gathered entry -- with CDP_FIXED_STRIDE=128)
// so one (scalar, base) pair costs nw = 16 mixed additions -- gathered 96-byte reads, no buckets, no doublings, no window
// combine -- instead of ~52 bucket additions plus the bucket reduction of the variable-base kernel.
#include <algorithm>

#include "launch.h"
#include "msm_common.cuh"
#include "batch_affine.cuh"

namespace cdp {

// ------------------------------------------------------------------------------------------------ table construction
// jac_out[base * nw + w] = 2^(c w) * B_base : one thread per base, c doublings between windows.
__global__ void __launch_bounds__(64) k_fixed_pow(const uint32_t *__restrict__ bases, uint32_t n_bases, int c, int nw, uint32_t *__restrict__ jac_out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_bases) return;
    g1a P;
    g1a_load(P, bases + 24 * (size_t)j);
    g1j acc;
    g1j_from_affine(acc, P);
#pragma unroll 1
    for (int w = 0; w < nw; w++) {
        g1j_store(jac_out + 36 * ((size_t)j * nw + w), acc);
#pragma unroll 1
        for (int k = 0; k < c; k++) g1j_dbl(acc, acc);
    }
}

// table[chain * nd + 0] = aff[chain]   (digit 1 of every (base, window) chain)
__global__ void __launch_bounds__(256) k_fixed_seed(const uint32_t *__restrict__ aff, uint32_t chains, uint32_t nd, uint32_t es, uint32_t *__restrict__ table) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ch = t / 6, q = t % 6;
    if (ch >= chains) return;
    reinterpret_cast<uint4 *>(table + (size_t)es * ch * nd)[q] = reinterpret_cast<const uint4 *>(aff + 24 * (size_t)ch)[q];
}

// One doubling level of every chain: entries [0, half) hold 1Q .. half*Q; this writes (half + e + 1) Q = (e + 1) Q + half Q for
// e < half.  Affine + affine with Montgomery's simultaneous inversion over CH entries per thread (one Fp inversion per CH points).
// e == half - 1 is the doubling (half Q + half Q).  Bases are prime-order points (or infinity: the whole chain stays all-zero), so no
// other coincidence of x-coordinates can occur below the group order.
constexpr int FIXED_CH = 16;
__global__ void __launch_bounds__(128) k_fixed_level(uint32_t *__restrict__ table, uint32_t chains, uint32_t nd, uint32_t half, uint32_t es) {
    const uint32_t parts = (half + FIXED_CH - 1) / FIXED_CH;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ch = t / parts, part = t - ch * parts;
    if (ch >= chains) return;
    const uint32_t e0 = part * FIXED_CH;
    const uint32_t cnt = half - e0 < (uint32_t)FIXED_CH ? half - e0 : (uint32_t)FIXED_CH;
    uint32_t *T = table + (size_t)es * ch * nd;  // entries are es words apart (24, or 32: one 128-byte line each)
    g1a Bp;
    g1a_load(Bp, T + (size_t)es * (half - 1));
    const bool binf = g1a_is_inf(Bp);
    fp prefix[FIXED_CH];
    fp run;
    fp_set_one(run);
#pragma unroll 1
    for (uint32_t i = 0; i < cnt; i++) {
        prefix[i] = run;
        fp den;
        if (binf) {
            fp_set_one(den);
        } else if (e0 + i == half - 1) {
            fp_dbl(den, Bp.y);
        } else {
            fp ax;
            fp_load(ax, T + (size_t)es * (e0 + i));
            fp_sub(den, Bp.x, ax);
        }
        fp_mul(run, run, den);
    }
    fp inv;
    fp_inv(inv, run);
#pragma unroll 1
    for (int i = (int)cnt - 1; i >= 0; i--) {
        g1a A, R;
        g1a_load(A, T + (size_t)es * (e0 + i));
        if (binf) {
            g1a_set_inf(R);
        } else {
            const bool dbl = (e0 + i == half - 1);
            fp den, num, lam, dinv;
            if (dbl) {
                fp_dbl(den, Bp.y);
                fp_sqr(num, Bp.x);
                fp_dbl(lam, num);
                fp_add(num, lam, num);  // 3 x^2
            } else {
                fp_sub(den, Bp.x, A.x);
                fp_sub(num, Bp.y, A.y);
            }
            fp_mul(dinv, inv, prefix[i]);
            fp_mul(inv, inv, den);
            fp_mul(lam, num, dinv);
            fp_sqr(R.x, lam);
            fp_sub(R.x, R.x, A.x);
            fp_sub(R.x, R.x, Bp.x);
            fp_sub(R.y, A.x, R.x);
            fp_mul(R.y, lam, R.y);
            fp_sub(R.y, R.y, A.y);
        }
        g1a_store(T + (size_t)es * (half + e0 + i), R);
    }
}

// ------------------------------------------------------------------------------------------------ the MSM
// Signed digit w of a canonical scalar k < 2^255: with k' = k + sum_{w < nw-1} 2^(c w + c - 1), digit_w = window_w(k') - 2^(c-1) in
// [-2^(c-1), 2^(c-1)) for w < nw - 1 and the top digit = k' >> c(nw-1) in [0, 2^(c-1)]: sum_w digit_w 2^(c w) = k.
__device__ __forceinline__ int fixed_digit(const uint32_t *__restrict__ sp, uint32_t w, const fixed_kparams_t &kp) {
    const uint4 a = reinterpret_cast<const uint4 *>(sp)[0], b = reinterpret_cast<const uint4 *>(sp)[1];
    uint32_t k[9];
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(k[0]), "=r"(k[1]), "=r"(k[2]), "=r"(k[3]), "=r"(k[4]), "=r"(k[5]), "=r"(k[6]), "=r"(k[7]), "=r"(k[8])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(kp.recode[0]), "r"(kp.recode[1]),
          "r"(kp.recode[2]), "r"(kp.recode[3]), "r"(kp.recode[4]), "r"(kp.recode[5]), "r"(kp.recode[6]), "r"(kp.recode[7]));
    const uint32_t bit = (uint32_t)kp.c * w, wi = bit >> 5, sh = bit & 31;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (uint32_t t = 0; t < 9; t++) {
        if (t == wi) lo = k[t];
        if (t == wi + 1) hi = k[t];
    }
    const uint32_t v = __funnelshift_r(lo, hi, sh);
    if (w + 1 < (uint32_t)kp.nw) return (int)(v & ((1u << kp.c) - 1u)) - (int)(kp.nd);
    return (int)(v < kp.nd ? v : kp.nd);  // canonical scalars never exceed nd here; the clamp only keeps a non-canonical one in bounds
}

// Table entry of item q = (pair q / nw, window q % nw) of a segment: nullptr when its digit is zero, else the entry's address and the digit's sign.
__device__ __forceinline__ const uint32_t *fixed_locate(const uint32_t *__restrict__ table, const uint32_t *__restrict__ scalars, const fixed_seg_t &seg,
                                                        const fixed_kparams_t &kp, uint32_t q, bool &neg) {
    const uint32_t i = q / (uint32_t)kp.nw, w = q - i * (uint32_t)kp.nw;
    uint32_t bidx, sidx2;
    if (i < seg.n) {
        uint32_t j = i;
        if (seg.sel_h) {
            const uint32_t lo = i & (seg.sel_h - 1);
            j = ((i - lo) << 1) | lo | seg.sel_val;
        }
        sidx2 = seg.scalars_off + j;
        const uint32_t pj = seg.pos_off + j * (seg.pos_stride ? seg.pos_stride : 1u);
        bidx = seg.base_off + pj + (pj >= seg.remap_from ? seg.remap_delta : 0u);
    } else {
        bidx = seg.extra_base - 1;
        sidx2 = seg.scalars_off + seg.extra_scalar;
    }
    const int d = fixed_digit(scalars + 8 * (size_t)sidx2, w, kp);
    if (d == 0) return nullptr;
    neg = d < 0;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    return table + (size_t)kp.es * (((size_t)bidx * kp.nw + w) * kp.nd + (ad - 1));
}

// One warp per segment.  Work item q = (pair i, window w) = (q / nw, q % nw); lane l takes items l, l + 32, ...; every item is one
// gathered 96-byte table read and one mixed addition into the lane's Jacobian accumulator; the next item's point is fetched before the
// current addition is issued, so the HBM gather latency (~1 us) hides behind ~3300 integer instructions.  The 32 partial sums are
// folded with warp shuffles (5 full additions).
template <int OCC, bool EXPANDED = false>
__global__ void __launch_bounds__(128, OCC) k_fixed_msm(const uint32_t *__restrict__ table, const uint32_t *__restrict__ scalars,
                                                       const fixed_seg_t *__restrict__ segs, uint32_t count, const fixed_kparams_t kp,
                                                       const uint32_t *__restrict__ var_pts, uint32_t *__restrict__ out_jac, uint32_t lg,
                                                       const uint32_t *__restrict__ arr = nullptr, uint32_t arr_n = 0) {
    // G = 2^lg lanes per segment (32: a warp per segment).  With many segments in flight fewer lanes per segment mean longer per-lane chains
    // and fewer fold steps: the 5 full additions of a 32-lane fold are ~12 % of a 129-pair segment's work, the 3 of an 8-lane fold ~2 %.
    const uint32_t G = 1u << lg, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sidx = gtid >> lg, lane = gtid & (G - 1);
    const bool active = sidx < count;  // lanes of an idle group still take part in the warp-wide shuffles below
    fixed_seg_t seg;
    if (active) {
        seg = segs[sidx];
    } else {
        seg.n = 0; seg.extra_base = 0; seg.addv_n = 0; seg.base_off = 0; seg.scalars_off = 0; seg.sel_h = 0; seg.sel_val = 0;
        seg.remap_from = 0xFFFFFFFFu; seg.remap_delta = 0; seg.extra_scalar = 0; seg.out_idx = 0; seg.addv_off = 0; seg.pos_off = 0; seg.pos_stride = 0;
    }
    // arr: the segment's table points were already summed down to arr_n partial sums per segment by the batched affine rounds below
    const uint32_t items = !active ? 0u : arr ? arr_n : (seg.n + (seg.extra_base ? 1u : 0u)) * (uint32_t)kp.nw;
    // returns true and the (un-negated) point of item q when there is one (non-zero digit / partial sum not at infinity)
    auto fetch = [&](uint32_t q, g1a &P, bool &neg) -> bool {
        if (arr) {
            g1a_load(P, arr + 24 * ((size_t)sidx * arr_n + q));
            neg = false;
            return !g1a_is_inf(P);
        }
        const uint32_t *e = fixed_locate(table, scalars, seg, kp, q, neg);
        if (!e) return false;
        g1a_load(P, e);
        return true;
    };
    g1j acc;
    g1j_set_inf(acc);
    g1a cur;
    bool cur_neg = false, have = false;
    uint32_t q = lane;
    while (q < items && !have) {
        have = fetch(q, cur, cur_neg);
        q += G;
    }
    if (EXPANDED) {
#pragma unroll 1
        while (have) {
            g1a nxt;
            bool nxt_neg = false, hn = false;
            while (q < items && !hn) {
                hn = fetch(q, nxt, nxt_neg);
                q += G;
            }
            if (cur_neg) fp_neg(cur.y, cur.y);
            g1j_add_mixed_expanded(acc, acc, cur);
            cur = nxt;
            cur_neg = nxt_neg;
            have = hn;
        }
    } else {  // the lane's running sum in XYZZ coordinates (8M + 2S per table point), Jacobian again for the fold across lanes
        g1x ax;
        g1x_set_inf(ax);
#pragma unroll 1
        while (have) {
            g1a nxt;
            bool nxt_neg = false, hn = false;
            while (q < items && !hn) {
                hn = fetch(q, nxt, nxt_neg);
                q += G;
            }
            if (cur_neg) fp_neg(cur.y, cur.y);
            g1x_add_mixed(ax, ax, cur);
            cur = nxt;
            cur_neg = nxt_neg;
            have = hn;
        }
        g1x_to_jac(acc, ax);
    }
#pragma unroll 1
    for (uint32_t a = lane; a < seg.addv_n; a += G) {  // plain (coefficient 1) device-resident points of the sum
        g1a_load(cur, var_pts + 24 * ((size_t)seg.addv_off + a));
        g1j_add_mixed(acc, acc, cur);
    }
#pragma unroll 1
    for (int d = (int)(G >> 1); d >= 1; d >>= 1) {
        g1j o;
        shfl_down_g1j(o, acc, d, (int)G);
        g1j_add(acc, acc, o);
    }
    if (active && lane == 0) g1j_store(out_jac + 36 * (size_t)seg.out_idx, acc);
}

// ---- the same kernel with the table points staged through shared memory by the bulk-copy engine (cp.async.bulk, SASS: UBLKCP) ----
// Every thread owns two 96-byte slots and two mbarriers: the copy of the NEXT item's table entry is issued (one instruction, no destination
// registers) before the current addition starts and is waited for only when its addition begins, so the gathered point does not occupy 24
// registers across a ~3300-instruction addition as the register-prefetch version above does.  Slots are 112 bytes apart (conflict-free
// 16-byte shared loads for 8 consecutive lanes).
constexpr uint32_t FIXED_SLOT_BYTES = 112;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_issue(uint32_t dst, const void *src, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(96u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(96u), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void bulk_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(mbar), "r"(parity)
                     : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();  // a lost copy must not hang the GPU
    } while (!ok);
}
template <int OCC>
__global__ void __launch_bounds__(128, OCC) k_fixed_msm_bulk(const uint32_t *__restrict__ table, const uint32_t *__restrict__ scalars,
                                                            const fixed_seg_t *__restrict__ segs, uint32_t count, const fixed_kparams_t kp,
                                                            const uint32_t *__restrict__ var_pts, uint32_t *__restrict__ out_jac) {
    extern __shared__ __align__(16) uint8_t fixed_smem[];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    // per thread: slots [2][112 B] then, behind all slots, mbarriers [2]
    uint8_t *slot0 = fixed_smem + (size_t)threadIdx.x * 2 * FIXED_SLOT_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(fixed_smem + (size_t)blockDim.x * 2 * FIXED_SLOT_BYTES) + 2 * threadIdx.x;
    const uint32_t s_slot = smem_u32(slot0), s_bar = smem_u32(bars);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_bar + 8));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp >= count) return;
    const fixed_seg_t seg = segs[warp];
    const uint32_t items = (seg.n + (seg.extra_base ? 1u : 0u)) * (uint32_t)kp.nw;
    auto locate = [&](uint32_t q, bool &neg) -> const uint32_t * { return fixed_locate(table, scalars, seg, kp, q, neg); };
    g1j acc;
    g1j_set_inf(acc);
    bool cur_neg = false, have = false;
    uint32_t q = lane, buf = 0, parity = 0;  // parity bit b = phase of buffer b's barrier
    while (q < items && !have) {
        const uint32_t *src = locate(q, cur_neg);
        q += 32;
        if (src) {
            bulk_issue(s_slot, src, s_bar);
            have = true;
        }
    }
#pragma unroll 1
    while (have) {
        bool nxt_neg = false, hn = false;
        while (q < items && !hn) {
            const uint32_t *src = locate(q, nxt_neg);
            q += 32;
            if (src) {
                bulk_issue(s_slot + (buf ^ 1) * FIXED_SLOT_BYTES, src, s_bar + (buf ^ 1) * 8);
                hn = true;
            }
        }
        bulk_wait(s_bar + buf * 8, (parity >> buf) & 1);
        parity ^= 1u << buf;
        g1a cur;
        {
            const uint4 *sp = reinterpret_cast<const uint4 *>(slot0 + buf * FIXED_SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint4 a = sp[k], b = sp[3 + k];
                cur.x.v[4 * k] = a.x; cur.x.v[4 * k + 1] = a.y; cur.x.v[4 * k + 2] = a.z; cur.x.v[4 * k + 3] = a.w;
                cur.y.v[4 * k] = b.x; cur.y.v[4 * k + 1] = b.y; cur.y.v[4 * k + 2] = b.z; cur.y.v[4 * k + 3] = b.w;
            }
        }
        if (cur_neg) fp_neg(cur.y, cur.y);
        g1j_add_mixed(acc, acc, cur);
        buf ^= 1;
        cur_neg = nxt_neg;
        have = hn;
    }
    if (lane < seg.addv_n) {  // plain (coefficient 1) device-resident points of the sum
        g1a cur;
        g1a_load(cur, var_pts + 24 * ((size_t)seg.addv_off + lane));
        g1j_add_mixed(acc, acc, cur);
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        g1j o;
        shfl_down_g1j(o, acc, d, 32);
        g1j_add(acc, acc, o);
    }
    if (lane == 0) g1j_store(out_jac + 36 * (size_t)seg.out_idx, acc);
}

// ---- the sum of a segment's table points as a tree of batched affine additions (batch_affine.cuh: 5M + 1S each instead of the 8M + 2S of
// the XYZZ accumulator above).  Every segment owns `ipad` item slots (its items, padded with "absent" to a multiple of 2^rounds); round 0
// adds items 2t and 2t + 1 of a segment straight from the table, round r the neighbours of round r - 1's output -- flat arrays, halving, no
// index structure at all -- and k_fixed_msm (array source) sums the ipad / 2^rounds partial sums left per segment, plus the plain points.
struct fixed_ba_src {
    const uint32_t *table, *scalars;
    const fixed_seg_t *segs;
    fixed_kparams_t kp;
    uint32_t half_ipad;  // jobs per segment
    uint32_t *out;
    __device__ __forceinline__ uint32_t *stash(uint32_t) const { return nullptr; }
    struct ref {
        const uint32_t *p, *q;  // table entries (nullptr: absent); the sign bits ride in the destination index
        uint32_t d;
    };
    __device__ __forceinline__ ref resolve(uint32_t job) const {
        const uint32_t sidx = job / half_ipad, t = job - sidx * half_ipad;
        const fixed_seg_t seg = segs[sidx];
        const uint32_t items = (seg.n + (seg.extra_base ? 1u : 0u)) * (uint32_t)kp.nw;
        bool n0 = false, n1 = false;
        ref r;
        r.p = 2 * t < items ? fixed_locate(table, scalars, seg, kp, 2 * t, n0) : nullptr;
        r.q = 2 * t + 1 < items ? fixed_locate(table, scalars, seg, kp, 2 * t + 1, n1) : nullptr;
        r.d = job | (n0 ? 0x80000000u : 0u) | (n1 ? 0x40000000u : 0u);
        return r;
    }
    __device__ __forceinline__ uint32_t *dst(const ref &r) const { return out + 24 * (size_t)(r.d & 0x3FFFFFFFu); }
    __device__ __forceinline__ void prefetch_x(const ref &r) const {
        if (r.p) prefetch_fp(r.p);
        if (r.q) prefetch_fp(r.q);
    }
    __device__ __forceinline__ void prefetch(const ref &r) const {
        if (r.p) prefetch_g1a(r.p);
        if (r.q) prefetch_g1a(r.q);
    }
    __device__ __forceinline__ void load_x(const ref &r, fp &px, fp &qx) const {
        fp_set_zero(px);
        fp_set_zero(qx);
        if (r.p) fp_load(px, r.p);
        if (r.q) fp_load(qx, r.q);
    }
    __device__ __forceinline__ void load(const ref &r, g1a &P, g1a &Q) const {
        g1a_set_inf(P);
        g1a_set_inf(Q);
        if (r.p) {
            g1a_load(P, r.p);
            if (r.d & 0x80000000u) fp_neg(P.y, P.y);
        }
        if (r.q) {
            g1a_load(Q, r.q);
            if (r.d & 0x40000000u) fp_neg(Q.y, Q.y);
        }
    }
};
struct fixed_ba_arr_src {
    const uint32_t *in;
    uint32_t *out;
    typedef uint32_t ref;
    __device__ __forceinline__ uint32_t *stash(uint32_t) const { return nullptr; }
    __device__ __forceinline__ ref resolve(uint32_t q) const { return q; }
    __device__ __forceinline__ uint32_t *dst(ref q) const { return out + 24 * (size_t)q; }
    __device__ __forceinline__ void prefetch_x(ref) const {}
    __device__ __forceinline__ void prefetch(ref) const {}
    __device__ __forceinline__ void load_x(ref q, fp &px, fp &qx) const {
        fp_load(px, in + 48 * (size_t)q);
        fp_load(qx, in + 48 * (size_t)q + 24);
    }
    __device__ __forceinline__ void load(ref q, g1a &P, g1a &Q) const {
        g1a_load(P, in + 48 * (size_t)q);
        g1a_load(Q, in + 48 * (size_t)q + 24);
    }
};
__global__ void __launch_bounds__(128, 4) k_fixed_ba_first(fixed_ba_src src, uint32_t total, uint32_t K, uint32_t pf) {
    const uint32_t T = (total + K - 1) / K, tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= T) return;
    ba_run<fixed_ba_src, 3>(src, tid, T, total, K, pf);
}
__global__ void __launch_bounds__(128, 4) k_fixed_ba_next(fixed_ba_arr_src src, uint32_t total, uint32_t K) {
    const uint32_t T = (total + K - 1) / K, tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= T) return;
    ba_run<fixed_ba_arr_src, 3>(src, tid, T, total, K);
}

cudaError_t launch_fixed_pow(cudaStream_t st, const uint32_t *bases_affine, uint32_t n_bases, int c, int nw, uint32_t *jac_out) {
    k_fixed_pow<<<(n_bases + 63) / 64, 64, 0, st>>>(bases_affine, n_bases, c, nw, jac_out);
    return cudaGetLastError();
}
cudaError_t launch_fixed_seed(cudaStream_t st, const uint32_t *aff, uint32_t chains, uint32_t nd, uint32_t es, uint32_t *table) {
    k_fixed_seed<<<(chains * 6 + 255) / 256, 256, 0, st>>>(aff, chains, nd, es, table);
    return cudaGetLastError();
}
cudaError_t launch_fixed_level(cudaStream_t st, uint32_t *table, uint32_t chains, uint32_t nd, uint32_t half, uint32_t es) {
    const uint64_t parts = (half + FIXED_CH - 1) / FIXED_CH, threads = parts * chains;
    k_fixed_level<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(table, chains, nd, half, es);
    return cudaGetLastError();
}
cudaError_t launch_fixed_msm(cudaStream_t st, const uint32_t *table, const uint32_t *scalars, const fixed_seg_t *segs, uint32_t count,
                             const fixed_kparams_t &kp, const uint32_t *var_pts, uint32_t *out_jac, int lanes_per_seg) {
    if (count == 0) return cudaSuccess;
    const int occ = tuned_occupancy("CDP_OCC_FIXED", 3);
    static const int bulk = [] { const char *e = getenv("CDP_FIXED_BULK"); return e ? atoi(e) : 0; }();
    if (bulk) {  // table points staged through shared memory by cp.async.bulk (see k_fixed_msm_bulk); a warp per segment
        const size_t smem = 128 * 2 * FIXED_SLOT_BYTES + 128 * 2 * 8;
        if (occ == 5) k_fixed_msm_bulk<5><<<(count + 3) / 4, 128, smem, st>>>(table, scalars, segs, count, kp, var_pts, out_jac);
        else if (occ == 4) k_fixed_msm_bulk<4><<<(count + 3) / 4, 128, smem, st>>>(table, scalars, segs, count, kp, var_pts, out_jac);
        else k_fixed_msm_bulk<3><<<(count + 3) / 4, 128, smem, st>>>(table, scalars, segs, count, kp, var_pts, out_jac);
        return cudaGetLastError();
    }
    static const int forced_lanes = [] { const char *e = getenv("CDP_FIXED_LANES"); return e ? atoi(e) : 0; }();
    if (forced_lanes == 8 || forced_lanes == 16 || forced_lanes == 32) lanes_per_seg = forced_lanes;
    const uint32_t lg = lanes_per_seg == 8 ? 3u : lanes_per_seg == 16 ? 4u : 5u;
    const unsigned blocks = (unsigned)((((size_t)count << lg) + 127) / 128);
    static const int expanded = [] { const char *e = getenv("CDP_FIXED_EXPANDED"); return e ? atoi(e) : 0; }();
    if (expanded) {  // the hot loop's mixed addition with its field products expanded in place (g1j_add_mixed_expanded)
        if (occ == 4) k_fixed_msm<4, true><<<blocks, 128, 0, st>>>(table, scalars, segs, count, kp, var_pts, out_jac, lg);
        else k_fixed_msm<3, true><<<blocks, 128, 0, st>>>(table, scalars, segs, count, kp, var_pts, out_jac, lg);
        return cudaGetLastError();
    }
    if (occ == 5) k_fixed_msm<5><<<blocks, 128, 0, st>>>(table, scalars, segs, count, kp, var_pts, out_jac, lg);
    else if (occ == 4) k_fixed_msm<4><<<blocks, 128, 0, st>>>(table, scalars, segs, count, kp, var_pts, out_jac, lg);
    else k_fixed_msm<3><<<blocks, 128, 0, st>>>(table, scalars, segs, count, kp, var_pts, out_jac, lg);
    return cudaGetLastError();
}

// scratch words needed by launch_fixed_msm_ba: round 0 writes count * ipad / 2 points, round 1 half of that (ping-pong from there on)
uint32_t fixed_ba_ipad(uint32_t max_pairs, int nw, int rounds) { return ((max_pairs * (uint32_t)nw + (1u << rounds) - 1) >> rounds) << rounds; }
size_t fixed_ba_scratch_bytes(uint32_t count, uint32_t max_pairs, int nw, int rounds) {
    const size_t ipad = fixed_ba_ipad(max_pairs, nw, rounds);
    return ((size_t)count * ipad / 2 + (size_t)count * ipad / 4) * 96;
}
cudaError_t launch_fixed_msm_ba(cudaStream_t st, const uint32_t *table, const uint32_t *scalars, const fixed_seg_t *segs, uint32_t count,
                                const fixed_kparams_t &kp, const uint32_t *var_pts, uint32_t *out_jac, uint32_t max_pairs, int rounds, uint32_t *scratch,
                                uint32_t t_target, uint32_t kmax) {
    if (count == 0) return cudaSuccess;
    const uint32_t ipad = fixed_ba_ipad(max_pairs, kp.nw, rounds);
    uint32_t *buf[2] = {scratch, scratch + (size_t)count * (ipad / 2) * 24};
    // all threads of a round do the same work: K packs the round into w full waves of t_target resident threads, w the fewest with K <= kmax
    auto k_for = [&](uint32_t pairs) {
        const uint64_t w = std::max<uint64_t>(1, ((uint64_t)pairs + (uint64_t)kmax * t_target - 1) / ((uint64_t)kmax * t_target));
        return std::max(4u, (uint32_t)((pairs + w * t_target - 1) / (w * t_target)));
    };
    uint32_t pairs = count * (ipad / 2);
    {
        const uint32_t K = k_for(pairs), T = (pairs + K - 1) / K;
        static const uint32_t pf = getenv("CDP_FIXED_TREE_PF") ? (uint32_t)atoi(getenv("CDP_FIXED_TREE_PF")) : 0u;
        k_fixed_ba_first<<<(T + 127) / 128, 128, 0, st>>>(fixed_ba_src{table, scalars, segs, kp, ipad / 2, buf[0]}, pairs, K, pf);
    }
    for (int r = 1; r < rounds; r++) {
        pairs >>= 1;
        const uint32_t K = k_for(pairs), T = (pairs + K - 1) / K;
        k_fixed_ba_next<<<(T + 127) / 128, 128, 0, st>>>(fixed_ba_arr_src{buf[(r + 1) & 1], buf[r & 1]}, pairs, K);
    }
    // ipad >> rounds partial sums per segment, plus its plain points, by 8 lanes per segment
    const uint32_t left = ipad >> rounds, lg = left > 256 ? 4u : 3u;
    const unsigned blocks = (unsigned)((((size_t)count << lg) + 127) / 128);
    k_fixed_msm<3><<<blocks, 128, 0, st>>>(table, scalars, segs, count, kp, var_pts, out_jac, lg, buf[(rounds + 1) & 1], left);
    return cudaGetLastError();
}

}  // namespace cdp
