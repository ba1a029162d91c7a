// Test helper (CPU): compiles the DEVICE verifier-scalar code (curdleproofs_b200/csrc/k_vcoeffs.cu) as plain C++ and compares every
// output scalar with a straightforward host restatement of the same algebra -- the eight `accumulate_check` calls of
// `CurdleproofsProof::verify` (/root/reference/src/curdleproofs.rs:283-296 and the argument verifiers they reach) plus the four
// SameScalar equalities -- written with the product's host Fr (host/fr.hpp) in the vector-at-a-time style the host driver used before
// this step moved to the GPU.  Random challenges; ell = 4, 12, 124, 252; both SameScalar modes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CDP_VCOEFFS_HOST_HARNESS
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
struct dim3_t { unsigned x; };
static dim3_t blockIdx{0}, blockDim{1}, threadIdx{0};
namespace cdp {
struct vcoef_params_t { uint32_t ell, n, m, big_n, vw, o_R, o_S, o_T, o_U, o_M, o_P, exact_eq, vch; };
struct round_expand_params_t { uint32_t n, h, spp, cpp, mode; };
}
#include "../../curdleproofs_b200/csrc/k_vcoeffs.cu"
#include "../../curdleproofs_b200/host/fr.hpp"

using namespace cdp_host;

static uint64_t rng_state = 0x243F6A8885A308D3ULL;
static uint64_t next64() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static Fr rand_fr() {
    uint64_t a[4] = {next64(), next64(), next64(), next64() >> 2};
    return Fr::raw(a) * Fr::raw(Fr::R2);  // some field element, Montgomery form
}
static void s_vector(std::vector<Fr> &s, const Fr *gam, size_t m) {
    size_t n = (size_t)1 << m;
    s.resize(n);
    s[0] = Fr::one();
    for (size_t j = m; j-- > 0;) {
        size_t bit = (size_t)1 << (m - 1 - j);
        for (size_t i = 0; i < bit; i++) s[bit + i] = s[i] * gam[j];
    }
}

static int run(size_t ell, bool exact_eq) {
    const size_t n = ell + 4;
    size_t m = 0;
    while (((size_t)1 << m) < n) m++;
    const size_t LC = 11, RC = LC + m, LD = RC + m, RD = LD + m, A1 = RD + m, A2 = A1 + 1, B1 = A2 + 1, B2 = B1 + 1, Ba = B2 + 1, Bt = Ba + 1,
                 Bu = Bt + 1, LA = Bu + 1, LT = LA + m, LU = LT + m, RA = LU + m, RT = RA + m, RU = RT + m, np = RU + m;
    const size_t LA_ = 0, T1 = 1, T2 = 2, U1 = 3, U2 = 4, R = 5, S = 6, Bp = 7, Cp = 8, Bc = 9, Bd = 10;
    const size_t crs_n = n + 5, oR = crs_n, oS = oR + ell, oT = oS + ell, oU = oT + ell, oM = oU + ell, oP = oM + 1, big_n = oP + np, scal_pp = big_n + 14;
    const size_t vw = big_n - crs_n + 20;  // the verifier's per-proof block width: per-proof slots + 20 gathered copies
    const size_t vch = 27 + 4 * m;
    const size_t B = 3;
    std::vector<Fr> ch(B * vch), va(B * ell);
    for (auto &x : ch) x = rand_fr();
    for (auto &x : va) x = rand_fr();
    std::vector<uint8_t> va_bytes(B * ell * 32), want(B * scal_pp * 32, 0), got(B * scal_pp * 32, 0xEE);
    std::vector<uint8_t> g_crs(B * crs_n * 32), g_var((B * vw + crs_n) * 32), g_ex(B * 14 * 32);
    for (size_t i = 0; i < B * ell; i++) va[i].to_bytes(&va_bytes[32 * i]);
    // ---- reference
    for (size_t pr = 0; pr < B; pr++) {
        const Fr *c = &ch[pr * vch], *rho = c;
        const Fr alpha_sp = c[12], beta_sp = c[13], alpha_g = c[14], beta_inv = c[15], alpha_i = c[16], beta_i = c[17], z = c[18], cf_c = c[19],
                 cf_d = c[20], xf = c[21], alpha_sm = c[22], alpha_ss = c[23], z_k = c[24], z_t = c[25], z_u = c[26];
        const Fr *gam = c + 27, *gam_inv = gam + m, *gam2 = gam_inv + m, *gam2_inv = gam2 + m;
        std::vector<Fr> s_ipa, sinv_ipa, s_sm, u(n);
        s_vector(s_ipa, gam, m); s_vector(sinv_ipa, gam_inv, m); s_vector(s_sm, gam2, m);
        Fr pw = beta_inv;
        for (size_t i = 0; i < ell; i++) { u[i] = pw; pw *= beta_inv; }
        for (size_t i = 0; i < 4; i++) u[ell + i] = pw;
        std::vector<Fr> cf(big_n, Fr::zero());
        const size_t cG = 0, cH = n, cGt = n + 1, cGu = n + 2, cGsum = n + 3, cHsum = n + 4;
        cf[oP + Bp] += rho[0]; cf[oP + LA_] -= rho[0]; cf[oM] -= rho[0] * alpha_sp;
        Fr t0 = rho[0] * beta_sp;
        for (size_t i = 0; i < ell; i++) cf[cG + i] -= t0;
        cf[oP + Bc] += rho[1]; cf[oP + Cp] += rho[1] * alpha_i;
        cf[cH] += rho[1] * (alpha_i * alpha_i * z * beta_i - cf_c * cf_d * beta_i);
        for (size_t k = 0; k < m; k++) { cf[oP + LC + k] += rho[1] * gam[k]; cf[oP + RC + k] += rho[1] * gam_inv[k]; }
        Fr t1 = rho[1] * cf_c;
        for (size_t i = 0; i < n; i++) cf[cG + i] -= t1 * s_ipa[i];
        cf[oP + Bd] += rho[2];
        Fr t2 = rho[2] * alpha_i;
        cf[oP + Bp] += t2; cf[cGsum] -= t2 * beta_inv; cf[cHsum] += t2 * alpha_g;
        for (size_t k = 0; k < m; k++) { cf[oP + LD + k] += rho[2] * gam[k]; cf[oP + RD + k] += rho[2] * gam_inv[k]; }
        Fr t3 = rho[2] * cf_d;
        for (size_t i = 0; i < n; i++) cf[cG + i] -= t3 * sinv_ipa[i] * u[i];
        const size_t Bx[3] = {Ba, Bt, Bu}, Lx[3] = {LA, LT, LU}, Rx[3] = {RA, RT, RU};
        for (int q = 0; q < 3; q++) {
            const Fr &r = rho[3 + q];
            cf[oP + Bx[q]] += r;
            for (size_t k = 0; k < m; k++) { cf[oP + Lx[q] + k] += r * gam2[k]; cf[oP + Rx[q] + k] += r * gam2_inv[k]; }
        }
        Fr a4 = rho[3] * alpha_sm;
        cf[oP + LA_] += a4; cf[oP + T1] += a4; cf[oP + U1] += a4;
        cf[oP + T2] += rho[4] * alpha_sm;
        cf[oP + U2] += rho[5] * alpha_sm;
        Fr x4 = rho[3] * xf, x5 = rho[4] * xf, x6 = rho[5] * xf;
        for (size_t i = 0; i < ell + 2; i++) cf[cG + i] -= x4 * s_sm[i];
        cf[cGt] -= x4 * s_sm[ell + 2]; cf[cGu] -= x4 * s_sm[ell + 3];
        for (size_t i = 0; i < ell; i++) { cf[oT + i] -= x5 * s_sm[i]; cf[oU + i] -= x6 * s_sm[i]; }
        cf[cH] -= x5 * s_sm[ell + 2];
        cf[cH] -= x6 * s_sm[ell + 3];
        cf[oP + R] += rho[6]; cf[oP + S] += rho[7];
        for (size_t i = 0; i < ell; i++) { cf[oR + i] -= rho[6] * va[pr * ell + i]; cf[oS + i] -= rho[7] * va[pr * ell + i]; }
        if (!exact_eq) {
            const Fr r8 = rho[8], r9 = rho[9], r10 = rho[10], r11 = rho[11];
            cf[oP + A1] += r8; cf[oP + T1] += r8 * alpha_ss; cf[cGt] -= r8 * z_t;
            cf[oP + A2] += r9; cf[oP + T2] += r9 * alpha_ss; cf[oP + R] -= r9 * z_k; cf[cH] -= r9 * z_t;
            cf[oP + B1] += r10; cf[oP + U1] += r10 * alpha_ss; cf[cGu] -= r10 * z_u;
            cf[oP + B2] += r11; cf[oP + U2] += r11 * alpha_ss; cf[oP + S] -= r11 * z_k; cf[cH] -= r11 * z_u;
        }
        for (size_t i = 0; i < ell; i++) cf[cG + i] += cf[cGsum];
        for (size_t i = ell; i < n; i++) cf[cG + i] += cf[cHsum];
        cf[cGsum] = cf[cHsum] = Fr::zero();
        uint8_t *sc = &want[pr * scal_pp * 32];
        for (size_t i = 0; i < big_n; i++) cf[i].to_bytes(sc + 32 * i);
        const Fr one = Fr::one();
        const Fr e[14] = {one, alpha_ss, z_t.neg(), one, alpha_ss, z_k.neg(), z_t.neg(), one, alpha_ss, z_u.neg(), one, alpha_ss, z_k.neg(), z_u.neg()};
        for (int i = 0; i < 14; i++) e[i].to_bytes(sc + 32 * (big_n + i));
    }
    // ---- device code on the CPU: 7 "threads" per proof (so that the strided loops and the single-slot thread are both exercised)
    cdp::vcoef_params_t P = {(uint32_t)ell, (uint32_t)n, (uint32_t)m, (uint32_t)big_n, (uint32_t)vw, (uint32_t)oR, (uint32_t)oS, (uint32_t)oT,
                             (uint32_t)oU, (uint32_t)oM, (uint32_t)oP, exact_eq ? 1u : 0u, (uint32_t)vch};
    for (uint32_t nthreads : {1u, 7u, 300u}) {
        std::fill(g_crs.begin(), g_crs.end(), 0xEE); std::fill(g_var.begin(), g_var.end(), 0xEE); std::fill(g_ex.begin(), g_ex.end(), 0xEE);
        for (uint32_t pr = 0; pr < B; pr++)
            for (uint32_t t = 0; t < nthreads; t++)
                cdp::vcoef_thread(pr, t, nthreads, reinterpret_cast<const uint32_t *>(ch.data()), reinterpret_cast<const uint32_t *>(va_bytes.data()), P,
                                  reinterpret_cast<uint32_t *>(g_crs.data()), reinterpret_cast<uint32_t *>(g_var.data()), reinterpret_cast<uint32_t *>(g_ex.data()));
        // back into one row per proof; every byte of the per-proof window outside [crs_n, big_n) must be untouched
        for (size_t pr = 0; pr < B; pr++) {
            uint8_t *row = &got[pr * scal_pp * 32];
            memcpy(row, &g_crs[pr * crs_n * 32], crs_n * 32);
            memcpy(row + crs_n * 32, &g_var[(pr * vw + crs_n) * 32], (big_n - crs_n) * 32);
            memcpy(row + big_n * 32, &g_ex[pr * 14 * 32], 14 * 32);
            for (size_t k = big_n; k < crs_n + vw; k++)
                for (int q = 0; q < 32; q++)
                    if (g_var[(pr * vw + k) * 32 + q] != 0xEE) { printf("ell=%zu: gathered-copy slot %zu of proof %zu was written\n", ell, k, pr); return 1; }
        }
        if (got != want) {
            for (size_t i = 0; i < got.size() / 32; i++)
                if (memcmp(&got[32 * i], &want[32 * i], 32)) { printf("ell=%zu exact=%d threads=%u: first mismatch at scalar %zu (slot %zu of proof %zu)\n", ell, (int)exact_eq, nthreads, i, i % scal_pp, i / scal_pp); break; }
            return 1;
        }
    }
    printf("ell=%zu exact_eq=%d ok (%zu scalars per proof)\n", ell, (int)exact_eq, scal_pp);
    return 0;
}

// The prover's round scalars: the host loop the batched prover used (per-index fold weights, n products per vector per round) against the
// device expansion from prefix weights (k_round_expand / round_expand_thread), over all log2(n) rounds of both arguments.
static Fr canonical_value(const Fr &x) { uint64_t c[4]; x.to_canonical(c); return Fr::raw(c); }
static int run_expand(size_t n) {
    size_t m = 0;
    while (((size_t)1 << m) < n) m++;
    const uint64_t one_c[4] = {1, 0, 0, 0};
    std::vector<Fr> c(n), d(n), x(n), u(n), wG(n, Fr::raw(one_c)), wGp(n), wS(n, Fr::raw(one_c));
    for (size_t i = 0; i < n; i++) { c[i] = rand_fr(); d[i] = rand_fr(); x[i] = rand_fr(); u[i] = rand_fr(); wGp[i] = canonical_value(u[i]); }
    std::vector<uint8_t> ucan(n * 32);
    for (size_t i = 0; i < n; i++) memcpy(&ucan[32 * i], wGp[i].v, 32);
    std::vector<Fr> Wc(1, Fr::raw(one_c)), Wd(1, Fr::one()), Ws(1, Fr::raw(one_c));
    for (size_t k = 0; k < m; k++) {
        const size_t h = n >> (k + 1), Q = n / (2 * h);
        // ---- reference: the per-index loops (prover.cpp before this step moved to the device)
        std::vector<uint8_t> want_ipa((2 * n + 2) * 32), want_sm((n + 2 * h) * 32);
        Fr ipL = rand_fr(), ipR = rand_fr();
        const Fr *cL = c.data(), *cR = c.data() + h, *dL = d.data(), *dR = d.data() + h;
        for (size_t j = 0; j < n; j++) {
            const size_t i = j & (h - 1);
            const bool hi = (j & h) != 0;
            Fr a = wG[j] * (hi ? cL[i] : cR[i]), b = wGp[j] * (hi ? dL[i] : dR[i]);
            memcpy(&want_ipa[32 * j], a.v, 32); memcpy(&want_ipa[32 * (n + 2 + j)], b.v, 32);
            Fr e = wS[j] * x[(j & h) ? (j & (h - 1)) : h + (j & (h - 1))];
            memcpy(&want_sm[32 * j], e.v, 32);
        }
        ipL.to_bytes(&want_ipa[32 * n]); ipR.to_bytes(&want_ipa[32 * (n + 1)]);
        for (size_t i = 0; i < 2 * h; i++) x[i].to_bytes(&want_sm[32 * (n + i)]);
        // ---- device expansion from the compact blocks
        std::vector<uint8_t> cmp_ipa((2 * Q + 4 * h + 2) * 32), cmp_sm((Q + 2 * h) * 32), got_ipa((2 * n + 2) * 32, 0xEE), got_sm((n + 2 * h) * 32, 0xEE);
        uint8_t *w = cmp_ipa.data();
        for (size_t q = 0; q < Q; q++, w += 32) memcpy(w, Wc[q].v, 32);
        for (size_t i = 0; i < 2 * h; i++, w += 32) memcpy(w, c[i].v, 32);
        for (size_t q = 0; q < Q; q++, w += 32) memcpy(w, Wd[q].v, 32);
        for (size_t i = 0; i < 2 * h; i++, w += 32) memcpy(w, d[i].v, 32);
        ipL.to_bytes(w); ipR.to_bytes(w + 32);
        w = cmp_sm.data();
        for (size_t q = 0; q < Q; q++, w += 32) memcpy(w, Ws[q].v, 32);
        for (size_t i = 0; i < 2 * h; i++, w += 32) memcpy(w, x[i].v, 32);
        cdp::round_expand_params_t P0 = {(uint32_t)n, (uint32_t)h, (uint32_t)(2 * n + 2), (uint32_t)(2 * Q + 4 * h + 2), 0};
        cdp::round_expand_params_t P1 = {(uint32_t)n, (uint32_t)h, (uint32_t)(n + 2 * h), (uint32_t)(Q + 2 * h), 1};
        for (uint32_t t = 0; t < 5; t++) {
            cdp::round_expand_thread(0, t, 5, reinterpret_cast<const uint32_t *>(cmp_ipa.data()), reinterpret_cast<const uint32_t *>(ucan.data()), P0,
                                     reinterpret_cast<uint32_t *>(got_ipa.data()));
            cdp::round_expand_thread(0, t, 5, reinterpret_cast<const uint32_t *>(cmp_sm.data()), nullptr, P1, reinterpret_cast<uint32_t *>(got_sm.data()));
        }
        if (got_ipa != want_ipa || got_sm != want_sm) { printf("expand n=%zu round %zu: mismatch (ipa %d, sm %d)\n", n, k, (int)(got_ipa != want_ipa), (int)(got_sm != want_sm)); return 1; }
        // ---- the folds: vectors, per-index weights (reference) and prefix weights (new)
        Fr gamma = rand_fr(), gamma_inv = gamma.inverse(), g2 = rand_fr(), g2_inv = g2.inverse();
        for (size_t i = 0; i < h; i++) { c[i] += gamma_inv * c[h + i]; d[i] += gamma * d[h + i]; x[i] += g2_inv * x[h + i]; }
        for (size_t j = 0; j < n; j++)
            if (j & h) { wG[j] *= gamma; wGp[j] *= gamma_inv; wS[j] *= g2; }
        std::vector<Fr> Wc2(2 * Q), Wd2(2 * Q), Ws2(2 * Q);
        for (size_t q = 0; q < Q; q++) {
            Wc2[2 * q] = Wc[q]; Wc2[2 * q + 1] = Wc[q] * gamma;
            Wd2[2 * q] = Wd[q]; Wd2[2 * q + 1] = Wd[q] * gamma_inv;
            Ws2[2 * q] = Ws[q]; Ws2[2 * q + 1] = Ws[q] * g2;
        }
        Wc.swap(Wc2); Wd.swap(Wd2); Ws.swap(Ws2);
    }
    printf("expand n=%zu ok (%zu rounds)\n", n, m);
    return 0;
}

int main() {
    if (run_expand(8) | run_expand(16) | run_expand(256)) return 1;

    int rc = 0;
    for (size_t ell : {4, 12, 124, 252})
        for (bool ex : {false, true}) rc |= run(ell, ex);
    return rc;
}
