// Batched affine additions: K independent sums R_i = P_i + Q_i of AFFINE points per thread with ONE field inversion (Montgomery's trick).
//
//   pass 1 (i = 0 .. K-1):  den_i = x(Q_i) - x(P_i)            prefix_i = den_0 * ... * den_i            1M
//   inv = 1 / prefix_{K-1}                                      binary Euclid on the integer ALU (fp_inv), shared by the K additions
//   pass 2 (i = K-1 .. 0):  1/den_i = inv * prefix_{i-1} ;  inv *= den_i                                  2M
//                           l = (y(Q) - y(P)) / den ;  x3 = l^2 - x(P) - x(Q) ;  y3 = l (x(P) - x3) - y(P)   2M + 1S
//
// 5M + 1S = 1 662 wide multiply-adds per addition against 8M + 2S = 2 748 for the XYZZ mixed addition the accumulators use (g1.cuh) and
// 11M + 5S for a full Jacobian addition: the bucket sums of a large `util::msm` (/root/reference/src/util.rs:19-22) are plain sums of many
// affine points, i.e. a tree of such additions.  prefix_i is parked in the first 48 bytes of the slot R_i will be written to, so the trick
// needs no memory of its own.  Every special case of the group law is decided per pair and keeps the chain intact (its denominator is 1):
// P or Q at infinity (x = y = 0), P = Q (tangent: 3x^2 / 2y), P = -Q (infinity out).
#pragma once
#include "g1.cuh"

namespace cdp {

enum { BA_GEN = 0, BA_DBL = 1, BA_COPY_P = 2, BA_COPY_Q = 3, BA_INF = 4 };

// kind of the pair and its denominator (Montgomery form; one for the kinds that need no division)
__device__ __forceinline__ int ba_classify(fp &den, const g1a &P, const g1a &Q) {
    const bool pinf = g1a_is_inf(P), qinf = g1a_is_inf(Q);
    if (pinf || qinf) {
        fp_set_one(den);
        return pinf ? BA_COPY_Q : BA_COPY_P;
    }
    if (fp_eq(P.x, Q.x)) {
        if (fp_eq(P.y, Q.y) && !fp_is_zero(P.y)) {
            fp_dbl(den, P.y);
            return BA_DBL;
        }
        fp_set_one(den);
        return BA_INF;
    }
    fp_sub(den, Q.x, P.x);
    return BA_GEN;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the sectors of a 16-byte aligned object of 48 / 96 bytes
__device__ __forceinline__ void prefetch_fp(const uint32_t *p) {
    prefetch_l2(p);
    prefetch_l2(p + 8);
}
__device__ __forceinline__ void prefetch_g1a(const uint32_t *p) {
    prefetch_l2(p);
    prefetch_l2(p + 8);
    prefetch_l2(p + 16);
    prefetch_l2(p + 23);
}

// Src: the K jobs of a thread.
//   ref resolve(uint32_t q)                      job q's operand / destination references (one independent load)
//   void prefetch_x(ref), prefetch(ref)          pull the x coordinates / both points and the destination's first 48 bytes towards L2
//   void load_x(ref, fp &px, fp &qx)             x coordinates of the two points (pass 1 reads nothing else for an ordinary pair)
//   void load(ref, g1a &P, g1a &Q)               both points
//   uint32_t *dst(ref)                           where R goes (96 bytes, 16-byte aligned)
// Thread `tid` of `T` takes jobs q = tid + i T < total, i < K.  The references of job i + 2 are fetched while job i is computed and its operands
// prefetched one job ahead: the loop is a chain of dependent loads (job -> operands) per ~3 000 instructions, and 12 warps per SM do not hide
// two DRAM latencies per job by themselves.
// INL: which products are expanded in place instead of called (the call marshals 36 registers through IMAD.MOV, which shares the multiply
// pipe): bit 0 the prefix product of pass 1, bit 1 the two products of the trick in pass 2, bit 2 the three of the addition itself
template <int INL, bool SQR = false>
__device__ __forceinline__ void ba_mul(fp &r, const fp &a, const fp &b) {
    if (INL) {
        if (SQR) fp_sqr_rw(r, a);
        else fp_mul_eo(r, a, b);
    } else {
        if (SQR) fp_sqr(r, a);
        else fp_mul(r, a, b);
    }
}
// STASH: pass 1 reads both points in full and parks them in src.stash(q) (192 bytes, streamed); pass 2 reads them back from there instead of
// gathering them a second time -- for sources whose operands are random 128-byte lines (the caller's bases, the digit table)
template <class Src, int INL = 0, bool STASH = false>
__device__ __forceinline__ void ba_run(const Src &src, uint32_t tid, uint32_t T, uint32_t total, uint32_t K, uint32_t pf = 0) {
    if (tid >= total) return;
    uint32_t cnt = (total - tid + T - 1) / T;
    if (cnt > K) cnt = K;
    typedef typename Src::ref ref;
    fp acc;
    fp_set_one(acc);
    {
        ref r0 = src.resolve(tid), r1 = src.resolve(cnt > 1 ? tid + T : tid);
        if (pf & 1) src.prefetch_x(r1);
#pragma unroll 1
        for (uint32_t i = 0; i < cnt; i++) {
            const ref r2 = src.resolve(i + 2 < cnt ? tid + (i + 2) * T : tid);
            fp den;
            if (STASH) {
                g1a P, Q;
                src.load(r0, P, Q);
                ba_classify(den, P, Q);
                uint32_t *sp = src.stash(tid + i * T);
                g1a_store(sp, P);
                g1a_store(sp + 24, Q);
            } else {
                fp px, qx;
                src.load_x(r0, px, qx);
                if (fp_is_zero(px) || fp_is_zero(qx) || fp_eq(px, qx)) {  // rare: infinity (x = 0 is necessary), tangent, cancellation
                    g1a P, Q;
                    src.load(r0, P, Q);
                    ba_classify(den, P, Q);
                } else {
                    fp_sub(den, qx, px);
                }
            }
            ba_mul<INL & 1>(acc, acc, den);
            fp_store(src.dst(r0), acc);
            if (pf & 1) src.prefetch_x(r2);
            r0 = r1;
            r1 = r2;
        }
    }
    fp inv;
    fp_inv(inv, acc);
    ref r0 = src.resolve(tid + (cnt - 1) * T), r1 = src.resolve(cnt > 1 ? tid + (cnt - 2) * T : tid);
    if (pf & 2) src.prefetch(r1);
#pragma unroll 1
    for (uint32_t i = cnt; i-- > 0;) {
        const ref r2 = src.resolve(i >= 2 ? tid + (i - 2) * T : tid);
        g1a P, Q;
        if (STASH) {
            const uint32_t *sp = src.stash(tid + i * T);
            g1a_load(P, sp);
            g1a_load(Q, sp + 24);
        } else {
            src.load(r0, P, Q);
        }
        fp den;
        const int kind = ba_classify(den, P, Q);
        fp dinv = inv;
        if (i) {
            fp prev;
            fp_load(prev, src.dst(r1));
            ba_mul<INL & 2>(dinv, inv, prev);
            ba_mul<INL & 2>(inv, inv, den);
        }
        g1a R;
        if (kind == BA_GEN || kind == BA_DBL) {
            fp num, l, t;
            if (kind == BA_GEN) {
                fp_sub(num, Q.y, P.y);
            } else {
                fp_sqr(t, P.x);
                fp_dbl(num, t);
                fp_add(num, num, t);
            }
            ba_mul<INL & 4>(l, num, dinv);
            ba_mul<INL & 4, true>(R.x, l, l);
            fp_sub(R.x, R.x, P.x);
            fp_sub(R.x, R.x, Q.x);
            fp_sub(t, P.x, R.x);
            ba_mul<INL & 4>(R.y, l, t);
            fp_sub(R.y, R.y, P.y);
        } else if (kind == BA_COPY_P) {
            R = P;
        } else if (kind == BA_COPY_Q) {
            R = Q;
        } else {
            g1a_set_inf(R);
        }
        g1a_store(src.dst(r0), R);
        if (pf & 2) src.prefetch(r2);
        r0 = r1;
        r1 = r2;
    }
}

}  // namespace cdp
