// Group law with the independent field products of a formula spread over the four lanes of a quad.
//
// The window combine of a small MSM (k_msm_combine) and the Horner pass of the large one are chains of ~130 dependent doublings and a few
// dozen additions on a handful of threads: pure latency (a field product takes a lone warp ~1 900 cycles, the carry chains cannot be
// overlapped any further), ~1.4 ms per MSM.  But inside one doubling only 3 of the 7 products depend on each other, inside a full addition
// 5 of 16.  Here the four lanes 4g .. 4g+3 of a warp hold the SAME point(s); at every dependency level lane s computes product s (all lanes
// run the same fp_mul on their own operands, so the instruction stream stays uniform) and the results are exchanged with quad-wide shuffles:
// 3 product latencies per doubling instead of 7, 5 per addition instead of 16.
// Every lane of the warp must call these functions together (full-mask shuffles); quads without work carry the point at infinity along.
#pragma once
#include "g1.cuh"

namespace cdp {

__device__ __forceinline__ void quad_get(fp &dst, const fp &src, int from) {
#pragma unroll
    for (int i = 0; i < 12; i++) dst.v[i] = __shfl_sync(0xffffffffu, src.v[i], from, 4);
}

// p <- 2p (dbl-2009-l, as g1j_dbl); valid for infinity (Z3 = 2 Y Z = 0)
__device__ __forceinline__ void g1j_dbl_quad(g1j &p, int s) {
    fp a, b, r, A, B, YZ, C, T, F, E, D, t;
    // level 1: X^2 | Y^2 | Y Z | (lane 3 repeats lane 2)
    a = s == 0 ? p.X : p.Y;
    b = s == 0 ? p.X : (s == 1 ? p.Y : p.Z);
    fp_mul(r, a, b);
    quad_get(A, r, 0);
    quad_get(B, r, 1);
    quad_get(YZ, r, 2);
    // level 2: B^2 | (X + B)^2 | (3A)^2
    fp_dbl(E, A);
    fp_add(E, E, A);
    fp_add(t, p.X, B);
    a = s == 0 ? B : (s == 1 ? t : E);
    fp_sqr(r, a);
    quad_get(C, r, 0);
    quad_get(T, r, 1);
    quad_get(F, r, 2);
    fp_sub(T, T, A);
    fp_sub(T, T, C);
    fp_dbl(D, T);
    fp_dbl(t, D);
    fp_sub(p.X, F, t);
    // level 3 (every lane the same product)
    fp_sub(t, D, p.X);
    fp_mul(t, E, t);
    fp_dbl(C, C);
    fp_dbl(C, C);
    fp_dbl(C, C);
    fp_sub(p.Y, t, C);
    fp_dbl(p.Z, YZ);
}

// r <- p + q, both Jacobian (add-2007-bl, as g1j_add); all special cases handled.  The formula is evaluated unconditionally (the shuffles
// must not diverge) and the result selected at the end; the doubling case (p == q) is a warp-uniform branch.
__device__ __forceinline__ void g1j_add_quad(g1j &r, const g1j &p, const g1j &q, int s) {
    const bool pinf = g1j_is_inf(p), qinf = g1j_is_inf(q);
    fp a, b, m, Z1Z1, Z2Z2, Y1Z2, Y2Z1, U1, U2, S1, S2, H, rr, I, ZZ, R2, J, V, Z3, X3, Y3, t;
    // level 1: Z1^2 | Z2^2 | Y1 Z2 | Y2 Z1
    a = s == 0 ? p.Z : (s == 1 ? q.Z : (s == 2 ? p.Y : q.Y));
    b = s == 0 ? p.Z : (s == 1 ? q.Z : (s == 2 ? q.Z : p.Z));
    fp_mul(m, a, b);
    quad_get(Z1Z1, m, 0);
    quad_get(Z2Z2, m, 1);
    quad_get(Y1Z2, m, 2);
    quad_get(Y2Z1, m, 3);
    // level 2: U1 = X1 Z2Z2 | U2 = X2 Z1Z1 | S1 = Y1 Z2 Z2Z2 | S2 = Y2 Z1 Z1Z1
    a = s == 0 ? p.X : (s == 1 ? q.X : (s == 2 ? Y1Z2 : Y2Z1));
    b = (s == 0 || s == 2) ? Z2Z2 : Z1Z1;
    fp_mul(m, a, b);
    quad_get(U1, m, 0);
    quad_get(U2, m, 1);
    quad_get(S1, m, 2);
    quad_get(S2, m, 3);
    fp_sub(H, U2, U1);
    fp_sub(rr, S2, S1);
    const bool same_x = fp_is_zero(H), same_y = fp_is_zero(rr);
    fp_dbl(rr, rr);
    // level 3: I = (2H)^2 | (Z1 + Z2)^2 | rr^2
    fp_dbl(t, H);
    fp_add(ZZ, p.Z, q.Z);
    a = s == 0 ? t : (s == 1 ? ZZ : rr);
    fp_sqr(m, a);
    quad_get(I, m, 0);
    quad_get(ZZ, m, 1);
    quad_get(R2, m, 2);
    fp_sub(ZZ, ZZ, Z1Z1);
    fp_sub(ZZ, ZZ, Z2Z2);
    // level 4: J = H I | V = U1 I | Z3 = ((Z1 + Z2)^2 - Z1Z1 - Z2Z2) H
    a = s == 0 ? H : (s == 1 ? U1 : ZZ);
    b = s == 2 ? H : I;
    if (s == 3) { a = H; b = I; }
    fp_mul(m, a, b);
    quad_get(J, m, 0);
    quad_get(V, m, 1);
    quad_get(Z3, m, 2);
    fp_sub(X3, R2, J);
    fp_sub(X3, X3, V);
    fp_sub(X3, X3, V);
    // level 5: rr (V - X3) | S1 J
    fp_sub(t, V, X3);
    a = (s & 1) == 0 ? rr : S1;
    b = (s & 1) == 0 ? t : J;
    fp_mul(m, a, b);
    quad_get(Y3, m, 0);
    quad_get(t, m, 1);
    fp_dbl(t, t);
    fp_sub(Y3, Y3, t);
    // p == q (neither at infinity): the doubling, computed by every quad of the warp when any quad needs it
    const bool need_dbl = !pinf && !qinf && same_x && same_y;
    g1j dbl = p;
    if (__any_sync(0xffffffffu, need_dbl)) g1j_dbl_quad(dbl, s);
    if (qinf) {
        r = p;
    } else if (pinf) {
        r = q;
    } else if (same_x) {
        if (same_y) r = dbl;
        else g1j_set_inf(r);
    } else {
        r.X = X3;
        r.Y = Y3;
        r.Z = Z3;
    }
}

}  // namespace cdp
