// Verifier scalar preparation on the device -- SURVEY.md section 8(f) rank 3.  For every proof of a batch the coefficient of EVERY
// base of its accumulated check is computed here from the O(log n) Fiat-Shamir challenges, instead of on the host:
//   * the verification scalars s_i = prod_{j : bit (m-1-j) of i set} gamma_j and their inverses
//     (`get_verification_scalars_bitstring`, /root/reference/src/util.rs:40-64; used at src/inner_product_argument.rs:202-250 and
//     src/same_multiscalar_argument.rs:242-259), the GrandProduct rescaling u_i = beta^-(i+1) (src/grand_product_argument.rs:92-102);
//   * the `a * x_i` products of the eight `MsmAccumulator::accumulate_check` calls (src/msm_accumulator.rs:37-52) with the
//     accumulator's HashMap replaced by the verifier's fixed slot table, plus the four SameScalar equalities
//     (src/same_scalar_argument.rs:127-136) with their own random factors.
// Input per proof: 27 + 4 m Montgomery scalars (random factors, challenges, the proof's seven scalars) and vec_a; output: the
// 5 ell + 8 + (proof points) canonical 32-byte scalars the MSM kernels read.  The host keeps the transcript (hashing) only.
// One CTA per proof, one thread per vector index; Fr in 8 x 32-bit limbs with 64-bit products (a few thousand products per proof:
// nothing here is throughput-critical).  Integer work only.
#ifndef CDP_VCOEFFS_HOST_HARNESS  // tests/host/vcoeffs_check.cpp compiles this file with g++ to check it on the CPU
#include "launch.h"
#endif
#include "fr256.cuh"

namespace cdp {



// One thread of the CTA that serves proof `pr`.  chal: [B][vch][8], vec_a: [B][ell][8] canonical.  Outputs, canonical: the CRS slots
// (< n + 5) go to out_crs[pr][n + 5], the per-proof slots (R, S, T, U, M, proof points) to out_var[pr * vw + slot] -- the same index the
// point of that slot has in the verifier's per-lane base array, so that ONE msm over the whole array is the merged check of the batch --
// and the 14 scalars of the exact SameScalar form to out_ex[pr][14].
__device__ __forceinline__ void vcoef_thread(uint32_t pr, uint32_t tid, uint32_t nthreads, const uint32_t *chal, const uint32_t *vec_a,
                                             const vcoef_params_t P, uint32_t *out_crs, uint32_t *out_var, uint32_t *out_ex) {
    using namespace vcoef;
    const uint32_t ell = P.ell, n = P.n, m = P.m;
    const uint32_t *ch = chal + 8 * (size_t)pr * P.vch;
    auto C = [&](uint32_t k) { return fr_load(ch + 8 * k); };
    const uint32_t *gam = ch + 8 * CH_VEC, *gam_inv = gam + 8 * m, *gam2 = gam_inv + 8 * m, *gam2_inv = gam2 + 8 * m;
    const uint32_t cH = n, cGt = n + 1, cGu = n + 2, cGsum = n + 3, cHsum = n + 4;
    // proof point indices in serialisation order (host/verifier.cpp ProofLayout)
    const uint32_t L_A = 0, L_T1 = 1, L_T2 = 2, L_U1 = 3, L_U2 = 4, L_R = 5, L_S = 6, L_B = 7, L_C = 8, L_Bc = 9, L_Bd = 10;
    const uint32_t L_LC = 11, L_RC = L_LC + m, L_LD = L_RC + m, L_RD = L_LD + m, L_A1 = L_RD + m, L_A2 = L_A1 + 1, L_B1 = L_A2 + 1,
                   L_B2 = L_B1 + 1, L_Ba = L_B2 + 1, L_Bt = L_Ba + 1, L_Bu = L_Bt + 1, L_LA = L_Bu + 1, L_LT = L_LA + m, L_LU = L_LT + m,
                   L_RA = L_LU + m, L_RT = L_RA + m, L_RU = L_RT + m;
    const bool joint = P.exact_eq == 0;  // the SameScalar equalities join the accumulated check with rho[8..11]
    const uint32_t crs_n = n + 5;
    auto put = [&](uint32_t slot, const fr_t &x) {
        uint32_t *d = slot < crs_n ? out_crs + 8 * ((size_t)pr * crs_n + slot)
                                   : slot < P.big_n ? out_var + 8 * ((size_t)pr * P.vw + slot) : out_ex + 8 * ((size_t)pr * 14 + (slot - P.big_n));
        fr_store_canonical(d, x);
    };
    const fr_t xf = C(CH_X);
    const fr_t x4 = fr_mul(C(CH_RHO + 3), xf), x5 = fr_mul(C(CH_RHO + 4), xf), x6 = fr_mul(C(CH_RHO + 5), xf);
    const fr_t t2 = fr_mul(C(CH_RHO + 2), C(CH_ALPHA_I));

    // ---- per-index part: CRS bases G | Hvec (slot i), R_i, S_i, T_i, U_i
    {
        const fr_t t0 = fr_mul(C(CH_RHO + 0), C(CH_BETA_SP)), t1 = fr_mul(C(CH_RHO + 1), C(CH_C)), t3 = fr_mul(C(CH_RHO + 2), C(CH_D));
        const fr_t binv = C(CH_BETA_INV);
        const fr_t g_sum = fr_neg(fr_mul(t2, binv)), h_sum = fr_mul(t2, C(CH_ALPHA_G));  // coefficients of sum(G), sum(Hvec): onto every G_i / Hvec_i
        const fr_t r6 = C(CH_RHO + 6), r7 = C(CH_RHO + 7);
        for (uint32_t i = tid; i < n; i += nthreads) {
            const fr_t s_ipa = s_value(gam, m, i), sinv_ipa = s_value(gam_inv, m, i), s_sm = s_value(gam2, m, i);
            const fr_t u = fr_pow_u32(binv, (i < ell ? i : ell) + 1);
            fr_t cf = i < ell ? fr_sub(g_sum, t0) : h_sum;
            cf = fr_sub(cf, fr_mul(t1, s_ipa));
            cf = fr_sub(cf, fr_mul(fr_mul(t3, sinv_ipa), u));
            if (i < ell + 2) cf = fr_sub(cf, fr_mul(x4, s_sm));
            put(i, cf);
            if (i < ell) {
                const fr_t a = fr_to_mont(fr_load(vec_a + 8 * ((size_t)pr * ell + i)));
                put(P.o_R + i, fr_neg(fr_mul(r6, a)));
                put(P.o_S + i, fr_neg(fr_mul(r7, a)));
                put(P.o_T + i, fr_neg(fr_mul(x5, s_sm)));
                put(P.o_U + i, fr_neg(fr_mul(x6, s_sm)));
            }
        }
    }
    // ---- the 10 m round points: thread k takes L/R number k of every argument
    for (uint32_t k = tid; k < m; k += nthreads) {
        const fr_t g = fr_load(gam + 8 * k), gi = fr_load(gam_inv + 8 * k), g2 = fr_load(gam2 + 8 * k), g2i = fr_load(gam2_inv + 8 * k);
        const fr_t r1 = C(CH_RHO + 1), r2 = C(CH_RHO + 2), r3 = C(CH_RHO + 3), r4 = C(CH_RHO + 4), r5 = C(CH_RHO + 5);
        put(P.o_P + L_LC + k, fr_mul(r1, g)); put(P.o_P + L_RC + k, fr_mul(r1, gi));
        put(P.o_P + L_LD + k, fr_mul(r2, g)); put(P.o_P + L_RD + k, fr_mul(r2, gi));
        put(P.o_P + L_LA + k, fr_mul(r3, g2)); put(P.o_P + L_RA + k, fr_mul(r3, g2i));
        put(P.o_P + L_LT + k, fr_mul(r4, g2)); put(P.o_P + L_RT + k, fr_mul(r4, g2i));
        put(P.o_P + L_LU + k, fr_mul(r5, g2)); put(P.o_P + L_RU + k, fr_mul(r5, g2i));
    }
    // ---- the remaining single slots (one thread)
    if (tid == (m < nthreads ? m : 0)) {
        const fr_t r0 = C(CH_RHO + 0), r1 = C(CH_RHO + 1), r2 = C(CH_RHO + 2), r3 = C(CH_RHO + 3), r4 = C(CH_RHO + 4), r5 = C(CH_RHO + 5),
                   r6 = C(CH_RHO + 6), r7 = C(CH_RHO + 7), r8 = C(CH_RHO + 8), r9 = C(CH_RHO + 9), r10 = C(CH_RHO + 10), r11 = C(CH_RHO + 11);
        const fr_t alpha_i = C(CH_ALPHA_I), beta_i = C(CH_BETA_I), alpha_sm = C(CH_ALPHA_SM), alpha_ss = C(CH_ALPHA_SS);
        const fr_t zk = C(CH_ZK), zt = C(CH_ZT), zu = C(CH_ZU);
        const fr_t s2 = s_value(gam2, m, ell + 2), s3 = s_value(gam2, m, ell + 3);
        const fr_t a4 = fr_mul(r3, alpha_sm);
        // H: IPA first check, the blinder slots of T (ell + 2) and U (ell + 3), SameScalar
        fr_t h = fr_sub(fr_mul(fr_mul(fr_mul(alpha_i, alpha_i), C(CH_Z)), beta_i), fr_mul(fr_mul(C(CH_C), C(CH_D)), beta_i));
        h = fr_mul(r1, h);
        h = fr_sub(h, fr_mul(x5, s2));
        h = fr_sub(h, fr_mul(x6, s3));
        fr_t gt = fr_neg(fr_mul(x4, s2)), gu = fr_neg(fr_mul(x4, s3));
        fr_t cT1 = a4, cU1 = a4, cT2 = fr_mul(r4, alpha_sm), cU2 = fr_mul(r5, alpha_sm), cR = r6, cS = r7;
        fr_t cA1 = fr_zero(), cA2 = fr_zero(), cB1 = fr_zero(), cB2 = fr_zero();
        if (joint) {
            h = fr_sub(h, fr_mul(r9, zt));
            h = fr_sub(h, fr_mul(r11, zu));
            gt = fr_sub(gt, fr_mul(r8, zt));
            gu = fr_sub(gu, fr_mul(r10, zu));
            cA1 = r8; cT1 = fr_add(cT1, fr_mul(r8, alpha_ss));
            cA2 = r9; cT2 = fr_add(cT2, fr_mul(r9, alpha_ss)); cR = fr_sub(cR, fr_mul(r9, zk));
            cB1 = r10; cU1 = fr_add(cU1, fr_mul(r10, alpha_ss));
            cB2 = r11; cU2 = fr_add(cU2, fr_mul(r11, alpha_ss)); cS = fr_sub(cS, fr_mul(r11, zk));
        }
        put(cH, h); put(cGt, gt); put(cGu, gu);
        put(cGsum, fr_zero()); put(cHsum, fr_zero());
        put(P.o_M, fr_neg(fr_mul(r0, C(CH_ALPHA_SP))));
        put(P.o_P + L_B, fr_add(r0, t2));
        put(P.o_P + L_A, fr_sub(a4, r0));
        put(P.o_P + L_Bc, r1); put(P.o_P + L_C, fr_mul(r1, alpha_i)); put(P.o_P + L_Bd, r2);
        put(P.o_P + L_Ba, r3); put(P.o_P + L_Bt, r4); put(P.o_P + L_Bu, r5);
        put(P.o_P + L_T1, cT1); put(P.o_P + L_U1, cU1); put(P.o_P + L_T2, cT2); put(P.o_P + L_U2, cU2);
        put(P.o_P + L_R, cR); put(P.o_P + L_S, cS);
        put(P.o_P + L_A1, cA1); put(P.o_P + L_A2, cA2); put(P.o_P + L_B1, cB1); put(P.o_P + L_B2, cB2);
        // scalars of the exact form of the SameScalar equalities (CDP_VERIFY_EXACT_EQ=1)
        const fr_t one = fr_one(), nzt = fr_neg(zt), nzk = fr_neg(zk), nzu = fr_neg(zu);
        const fr_t e[14] = {one, alpha_ss, nzt, one, alpha_ss, nzk, nzt, one, alpha_ss, nzu, one, alpha_ss, nzk, nzu};
        for (uint32_t k = 0; k < 14; k++) put(P.big_n + k, e[k]);
    }
}

__global__ void __launch_bounds__(256, 2) k_verify_coeffs(const uint32_t *__restrict__ chal, const uint32_t *__restrict__ vec_a,
                                                       const vcoef_params_t P, uint32_t *__restrict__ out_crs, uint32_t *__restrict__ out_var,
                                                       uint32_t *__restrict__ out_ex) {
    vcoef_thread(blockIdx.x, threadIdx.x, blockDim.x, chal, vec_a, P, out_crs, out_var, out_ex);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Prover side: the scalars of one folding round, expanded on the device.  The batched prover writes the round MSMs of the IPA / SameMSM
// arguments over the ORIGINAL bases (/root/reference/src/inner_product_argument.rs:158-161, src/same_multiscalar_argument.rs:107-112 with
// the folded bases G^(k)_i = sum_{j = i mod n_k} w(j) G_j): scalar j is   w(prefix of j) * v[side of j],
// where w depends only on the bits of j above the round's split h (the product of the earlier rounds' challenges on those bits) and v is
// the current folded vector of 2h entries.  The host therefore only sends Q = n / 2h prefix weights and the 2h vector entries; this kernel
// forms the n (or 2n) products -- the host used to do these n products per vector per round itself, which was the bulk of its field work.
//   mode 0 (IPA):     compact = Wc[Q] canonical | c[2h] Montgomery | Wd[Q] Montgomery | d[2h] Montgomery | ipL | ipR (canonical)
//                     out[j] = Wc[q] c[sel j],  out[n] = ipL, out[n+1] = ipR,  out[n+2+j] = Wd[q] u[j] d[sel j]   (u canonical, per proof: G' = u o G)
//   mode 1 (SameMSM): compact = Ws[Q] canonical | x[2h] Montgomery;   out[j] = Ws[q] x[sel j],  out[n + i] = x[i] (canonical), i < 2h
//   q = j / 2h,  sel j = (j mod h) when bit h of j is set, h + (j mod h) otherwise.   A Montgomery product of a canonical and a
//   Montgomery-form value is the canonical product, so every output is already the byte form the MSM kernels read.
__device__ __forceinline__ void round_expand_thread(uint32_t pr, uint32_t tid, uint32_t nthreads, const uint32_t *compact, const uint32_t *ucan,
                                                    const round_expand_params_t P, uint32_t *out) {
    using namespace vcoef;
    const uint32_t n = P.n, h = P.h, Q = n / (2 * h);
    const uint32_t *cb = compact + 8 * (size_t)pr * P.cpp;
    uint32_t *sc = out + 8 * (size_t)pr * P.spp;
    auto put_raw = [&](uint32_t slot, const fr_t &x) { for (int k = 0; k < 8; k++) sc[8 * slot + k] = x.v[k]; };
    if (P.mode == 0) {
        const uint32_t *Wc = cb, *cv = cb + 8 * Q, *Wd = cv + 8 * 2 * h, *dv = Wd + 8 * Q, *ip = dv + 8 * 2 * h;
        const uint32_t *u = ucan + 8 * (size_t)pr * n;
        for (uint32_t j = tid; j < n; j += nthreads) {
            const uint32_t q = j / (2 * h), i = j & (h - 1), sel = (j & h) ? i : h + i;
            put_raw(j, fr_mul(fr_load(Wc + 8 * q), fr_load(cv + 8 * sel)));
            put_raw(n + 2 + j, fr_mul(fr_mul(fr_load(Wd + 8 * q), fr_load(u + 8 * j)), fr_load(dv + 8 * sel)));
        }
        if (tid == 0) { put_raw(n, fr_load(ip)); put_raw(n + 1, fr_load(ip + 8)); }
    } else {
        const uint32_t *Ws = cb, *xv = cb + 8 * Q;
        for (uint32_t j = tid; j < n; j += nthreads) {
            const uint32_t q = j / (2 * h), i = j & (h - 1), sel = (j & h) ? i : h + i;
            put_raw(j, fr_mul(fr_load(Ws + 8 * q), fr_load(xv + 8 * sel)));
        }
        for (uint32_t i = tid; i < 2 * h; i += nthreads) fr_store_canonical(sc + 8 * (n + i), fr_load(xv + 8 * i));
    }
}
__global__ void __launch_bounds__(256, 2) k_round_expand(const uint32_t *__restrict__ compact, const uint32_t *__restrict__ ucan,
                                                         const round_expand_params_t P, uint32_t *__restrict__ out) {
    round_expand_thread(blockIdx.x, threadIdx.x, blockDim.x, compact, ucan, P, out);
}

// out[i] = sum over rows of in[row * stride + i] (mod r), canonical scalars in and out: the CRS coefficients of all proofs of a sub-batch
// added up for the merged check (the accumulator's `entry(base) += a * x_i` of msm_accumulator.rs:47-51 across proofs).  One warp per column:
// the lanes take every 32nd row, then a shuffle reduction.
#ifndef CDP_VCOEFFS_HOST_HARNESS
__global__ void __launch_bounds__(128) k_sum_scalars(const uint32_t *__restrict__ in, uint32_t stride, uint32_t cols, uint32_t rows,
                                                     uint32_t *__restrict__ out) {
    using namespace vcoef;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= cols) return;
    fr_t acc = fr_zero();
    for (uint32_t r = lane; r < rows; r += 32) acc = fr_add(acc, fr_load(in + 8 * ((size_t)r * stride + i)));
    for (int d = 16; d >= 1; d >>= 1) {
        fr_t o;
        for (int k = 0; k < 8; k++) o.v[k] = __shfl_down_sync(0xffffffffu, acc.v[k], d);
        acc = fr_add(acc, o);
    }
    if (lane == 0)
        for (int k = 0; k < 8; k++) out[8 * (size_t)i + k] = acc.v[k];
}
#endif

#ifndef CDP_VCOEFFS_HOST_HARNESS
cudaError_t launch_verify_coeffs(cudaStream_t st, const uint32_t *chal, const uint32_t *vec_a, const vcoef_params_t &P, uint32_t batch,
                                 uint32_t *out_crs, uint32_t *out_var, uint32_t *out_ex) {
    k_verify_coeffs<<<batch, 256, 0, st>>>(chal, vec_a, P, out_crs, out_var, out_ex);
    return cudaGetLastError();
}
cudaError_t launch_round_expand(cudaStream_t st, const uint32_t *compact, const uint32_t *ucan, const round_expand_params_t &P, uint32_t batch,
                                uint32_t *out) {
    k_round_expand<<<batch, 256, 0, st>>>(compact, ucan, P, out);
    return cudaGetLastError();
}
cudaError_t launch_sum_scalars(cudaStream_t st, const uint32_t *in, uint32_t stride, uint32_t cols, uint32_t rows, uint32_t *out) {
    k_sum_scalars<<<(cols * 32 + 127) / 128, 128, 0, st>>>(in, stride, cols, rows, out);
    return cudaGetLastError();
}
#endif

}  // namespace cdp
