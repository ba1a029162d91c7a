// BLS12-381 G1 group law on device (y^2 = x^3 + 4), Jacobian coordinates, a = 0.
//
// Device counterpart of ark-ec's short-Weierstrass `Affine` / `Projective` arithmetic that the reference reaches
// through `util::msm` (/root/reference/src/util.rs:19-29), `Mul<Fr>` and `into_affine()` in the fold loops
// (/root/reference/src/inner_product_argument.rs:174-179, src/same_multiscalar_argument.rs:126-131,
// src/grand_product_argument.rs:92-102).  Every output that the reference ever observes is a canonical group
// element (affine / compressed), so only the group law has to agree -- not the projective representative.
//
// Layouts (include/cdp_msm.h): affine = x || y, 2 x 48 B Montgomery, infinity <=> x == y == 0 (not on the curve);
// Jacobian = X || Y || Z, infinity <=> Z == 0.  Handles the edge cases the reference feeds the path
// (SURVEY.md D9): infinity bases, zero scalars, P + P, P + (-P).
#pragma once
#include "fp381.cuh"

namespace cdp {

struct g1a {
    fp x, y;
};
struct g1j {
    fp X, Y, Z;
};

__device__ __forceinline__ bool g1a_is_inf(const g1a &p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
__device__ __forceinline__ bool g1j_is_inf(const g1j &p) { return fp_is_zero(p.Z); }
__device__ __forceinline__ void g1j_set_inf(g1j &p) {
    fp_set_one(p.X);
    fp_set_one(p.Y);
    fp_set_zero(p.Z);
}
__device__ __forceinline__ void g1a_set_inf(g1a &p) {
    fp_set_zero(p.x);
    fp_set_zero(p.y);
}
__device__ __forceinline__ void g1j_from_affine(g1j &r, const g1a &p) {
    if (g1a_is_inf(p)) {
        g1j_set_inf(r);
    } else {
        r.X = p.x;
        r.Y = p.y;
        fp_set_one(r.Z);
    }
}
__device__ __forceinline__ void g1a_load(g1a &p, const uint32_t *src) {
    fp_load(p.x, src);
    fp_load(p.y, src + 12);
}
__device__ __forceinline__ void g1a_store(uint32_t *dst, const g1a &p) {
    fp_store(dst, p.x);
    fp_store(dst + 12, p.y);
}
__device__ __forceinline__ void g1j_load(g1j &p, const uint32_t *src) {
    fp_load(p.X, src);
    fp_load(p.Y, src + 12);
    fp_load(p.Z, src + 24);
}
__device__ __forceinline__ void g1j_store(uint32_t *dst, const g1j &p) {
    fp_store(dst, p.X);
    fp_store(dst + 12, p.Y);
    fp_store(dst + 24, p.Z);
}

// r = 2p   (2M + 5S; valid for infinity as well: Z3 = 2*Y*Z = 0)
__device__ __forceinline__ void g1j_dbl(g1j &r, const g1j &p) {
    fp A, B, C, D, E, F, t;
    fp_sqr(A, p.X);
    fp_sqr(B, p.Y);
    fp_sqr(C, B);
    fp_add(t, p.X, B);
    fp_sqr(t, t);
    fp_sub(t, t, A);
    fp_sub(t, t, C);
    fp_dbl(D, t);
    fp_dbl(E, A);
    fp_add(E, E, A);
    fp_sqr(F, E);
    fp_mul(t, p.Y, p.Z);
    fp_dbl(r.Z, t);
    fp_dbl(t, D);
    fp_sub(r.X, F, t);
    fp_sub(t, D, r.X);
    fp_mul(t, E, t);
    fp_dbl(C, C);
    fp_dbl(C, C);
    fp_dbl(C, C);
    fp_sub(r.Y, t, C);
}

// Out-of-line doubling for the (rare) P == Q branch inside the addition formulas: keeps the inlined hot code small.
static __device__ __noinline__ void g1j_dbl_outlined(g1j *r, const g1j *p) {
    g1j t = *p;
    g1j_dbl(t, t);
    *r = t;
}

// r = p + q, q affine (7M + 4S); all special cases handled
__device__ __forceinline__ void g1j_add_mixed(g1j &r, const g1j &p, const g1a &q) {
    if (g1a_is_inf(q)) {
        r = p;
        return;
    }
    if (g1j_is_inf(p)) {
        r.X = q.x;
        r.Y = q.y;
        fp_set_one(r.Z);
        return;
    }
    fp Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    fp_sqr(Z1Z1, p.Z);
    fp_mul(U2, q.x, Z1Z1);
    fp_mul(S2, q.y, p.Z);
    fp_mul(S2, S2, Z1Z1);
    fp_sub(H, U2, p.X);
    fp_sub(rr, S2, p.Y);
    if (fp_is_zero(H)) {
        if (fp_is_zero(rr)) {
            g1j_dbl_outlined(&r, &p);
        } else {
            g1j_set_inf(r);
        }
        return;
    }
    fp_dbl(rr, rr);
    fp_sqr(HH, H);
    fp_dbl(I, HH);
    fp_dbl(I, I);
    fp_mul(J, H, I);
    fp_mul(V, p.X, I);
    fp X3, Y3, Z3;
    fp_sqr(X3, rr);
    fp_sub(X3, X3, J);
    fp_sub(X3, X3, V);
    fp_sub(X3, X3, V);
    fp_sub(t, V, X3);
    fp_mul(Y3, rr, t);
    fp_mul(t, p.Y, J);
    fp_dbl(t, t);
    fp_sub(Y3, Y3, t);
    fp_add(Z3, p.Z, H);
    fp_sqr(Z3, Z3);
    fp_sub(Z3, Z3, Z1Z1);
    fp_sub(Z3, Z3, HH);
    r.X = X3;
    r.Y = Y3;
    r.Z = Z3;
}

// ---- XYZZ accumulators: (X, Y, ZZ, ZZZ) with x = X / ZZ, y = Y / ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ = 0.
// The bucket / table accumulation loops add affine points into one running sum: in XYZZ that is 8M + 2S (madd-2008-s) instead of the Jacobian
// 7M + 4S, with far less add / double glue between the products: 2.95 instead of 2.59 G additions/s in the chain micro-benchmark
// (tools/madd_bench.py, +14 %), for 12 more live registers.  (X ZZ, Y ZZZ, ZZ) is the same point in Jacobian coordinates.
struct g1x {
    fp X, Y, ZZ, ZZZ;
};
__device__ __forceinline__ void g1x_set_inf(g1x &p) {
    fp_set_one(p.X);
    fp_set_one(p.Y);
    fp_set_zero(p.ZZ);
    fp_set_zero(p.ZZZ);
}
__device__ __forceinline__ bool g1x_is_inf(const g1x &p) { return fp_is_zero(p.ZZ); }
// 2q for an affine q != infinity (mdbl-2008-s-1); the rare p == q branch of the addition below
static __device__ __noinline__ void g1x_mdbl_outlined(g1x *r, const g1a *q) {
    fp U, V, W, S, M, t;
    fp_dbl(U, q->y);
    fp_sqr(V, U);
    fp_mul(W, U, V);
    fp_mul(S, q->x, V);
    fp_sqr(M, q->x);
    fp_dbl(t, M);
    fp_add(M, t, M);
    fp_sqr(r->X, M);
    fp_sub(r->X, r->X, S);
    fp_sub(r->X, r->X, S);
    fp_sub(t, S, r->X);
    fp_mul(t, M, t);
    fp_mul(S, W, q->y);
    fp_sub(r->Y, t, S);
    r->ZZ = V;
    r->ZZZ = W;
}
// r = p + q, q affine; all special cases handled
__device__ __forceinline__ void g1x_add_mixed(g1x &r, const g1x &p, const g1a &q) {
    if (g1a_is_inf(q)) {
        r = p;
        return;
    }
    if (g1x_is_inf(p)) {
        r.X = q.x;
        r.Y = q.y;
        fp_set_one(r.ZZ);
        fp_set_one(r.ZZZ);
        return;
    }
    fp U2, S2, P, R, PP, PPP, Q, t;
    fp_mul(U2, q.x, p.ZZ);
    fp_mul(S2, q.y, p.ZZZ);
    fp_sub(P, U2, p.X);
    fp_sub(R, S2, p.Y);
    if (fp_is_zero(P)) {
        if (fp_is_zero(R)) {
            g1x_mdbl_outlined(&r, &q);
        } else {
            g1x_set_inf(r);
        }
        return;
    }
    fp_sqr(PP, P);
    fp_mul(PPP, P, PP);
    fp_mul(Q, p.X, PP);
    fp X3;
    fp_sqr(X3, R);
    fp_sub(X3, X3, PPP);
    fp_sub(X3, X3, Q);
    fp_sub(X3, X3, Q);
    fp_sub(t, Q, X3);
    fp_mul(t, R, t);
    fp_mul(S2, p.Y, PPP);
    fp_sub(r.Y, t, S2);
    fp_mul(r.ZZ, p.ZZ, PP);
    fp_mul(r.ZZZ, p.ZZZ, PPP);
    r.X = X3;
}
__device__ __forceinline__ void g1x_to_jac(g1j &r, const g1x &p) {
    fp_mul(r.X, p.X, p.ZZ);
    fp_mul(r.Y, p.Y, p.ZZZ);
    r.Z = p.ZZ;
}

// The same mixed addition with the field products expanded in place (fp_mul_eo / fp_sqr_rw inlined) instead of called: ~55 KB of
// straight-line SASS, but none of the ~36 register moves per call that the by-value ABI of fp_mul_fn / fp_sqr_fn costs -- ptxas emits those
// as IMAD.MOV.U32, i.e. on the same fmaheavy pipe that the 2904 IMAD.WIDE of the addition already keep ~70 % busy
// (profiles/r02_ncu_fixed_pipes.txt).  For the one hot loop of a kernel only; everything else keeps the compact called form.
__device__ __forceinline__ void g1j_add_mixed_expanded(g1j &r, const g1j &p, const g1a &q) {
    if (g1a_is_inf(q)) {
        r = p;
        return;
    }
    if (g1j_is_inf(p)) {
        r.X = q.x;
        r.Y = q.y;
        fp_set_one(r.Z);
        return;
    }
    fp Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    fp_sqr_rw(Z1Z1, p.Z);
    fp_mul_eo(U2, q.x, Z1Z1);
    fp_mul_eo(S2, q.y, p.Z);
    fp_mul_eo(S2, S2, Z1Z1);
    fp_sub(H, U2, p.X);
    fp_sub(rr, S2, p.Y);
    if (fp_is_zero(H)) {
        if (fp_is_zero(rr)) {
            g1j_dbl_outlined(&r, &p);
        } else {
            g1j_set_inf(r);
        }
        return;
    }
    fp_dbl(rr, rr);
    fp_sqr_rw(HH, H);
    fp_dbl(I, HH);
    fp_dbl(I, I);
    fp_mul_eo(J, H, I);
    fp_mul_eo(V, p.X, I);
    fp X3, Y3, Z3;
    fp_sqr_rw(X3, rr);
    fp_sub(X3, X3, J);
    fp_sub(X3, X3, V);
    fp_sub(X3, X3, V);
    fp_sub(t, V, X3);
    fp_mul_eo(Y3, rr, t);
    fp_mul_eo(t, p.Y, J);
    fp_dbl(t, t);
    fp_sub(Y3, Y3, t);
    fp_add(Z3, p.Z, H);
    fp_sqr_rw(Z3, Z3);
    fp_sub(Z3, Z3, Z1Z1);
    fp_sub(Z3, Z3, HH);
    r.X = X3;
    r.Y = Y3;
    r.Z = Z3;
}

// r = p + q, both Jacobian (11M + 5S); all special cases handled
__device__ __forceinline__ void g1j_add(g1j &r, const g1j &p, const g1j &q) {
    if (g1j_is_inf(q)) {
        r = p;
        return;
    }
    if (g1j_is_inf(p)) {
        r = q;
        return;
    }
    fp Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
    fp_sqr(Z1Z1, p.Z);
    fp_sqr(Z2Z2, q.Z);
    fp_mul(U1, p.X, Z2Z2);
    fp_mul(U2, q.X, Z1Z1);
    fp_mul(S1, p.Y, q.Z);
    fp_mul(S1, S1, Z2Z2);
    fp_mul(S2, q.Y, p.Z);
    fp_mul(S2, S2, Z1Z1);
    fp_sub(H, U2, U1);
    fp_sub(rr, S2, S1);
    if (fp_is_zero(H)) {
        if (fp_is_zero(rr)) {
            g1j_dbl_outlined(&r, &p);
        } else {
            g1j_set_inf(r);
        }
        return;
    }
    fp_dbl(rr, rr);
    fp_dbl(I, H);
    fp_sqr(I, I);
    fp_mul(J, H, I);
    fp_mul(V, U1, I);
    fp X3, Y3, Z3;
    fp_sqr(X3, rr);
    fp_sub(X3, X3, J);
    fp_sub(X3, X3, V);
    fp_sub(X3, X3, V);
    fp_sub(t, V, X3);
    fp_mul(Y3, rr, t);
    fp_mul(t, S1, J);
    fp_dbl(t, t);
    fp_sub(Y3, Y3, t);
    fp_add(Z3, p.Z, q.Z);
    fp_sqr(Z3, Z3);
    fp_sub(Z3, Z3, Z1Z1);
    fp_sub(Z3, Z3, Z2Z2);
    fp_mul(Z3, Z3, H);
    r.X = X3;
    r.Y = Y3;
    r.Z = Z3;
}

}  // namespace cdp
