// TEST INFRASTRUCTURE ONLY -- never built into, linked by or loaded from the product (curdleproofs_b200/).
//
// A CPU stand-in for the subset of the C ABI of include/cdp_msm.h that the batched prover's host driver (curdleproofs_b200/host/prover.cpp)
// calls, so that the driver's orchestration AND the device-side prover code (csrc/k_prove.cu, csrc/k_transcript.cu, csrc/k_vcoeffs.cu's
// round expansion -- the real sources, compiled as plain C++) can be checked in a container without a GPU:
//   * "device memory" is host memory, the stream is synchronous;
//   * every group operation (MSM segments, folds, normalisation, encodings) is evaluated with the CPU oracle (oracle/liboracle.so),
//     directly from the segment / job descriptors, i.e. from the scalars the code under test produced;
//   * the transcript / scalar kernels run one emulated CTA (or thread) per proof.
// tests/host/prove_dev_check.cpp links this with prover.cpp and compares whole proofs with the oracle's prover, byte for byte
// (including the reference's 4496-byte golden vector, /root/reference/src/whisk.rs:455).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CDP_TRANSCRIPT_HOST_HARNESS
#define CDP_VCOEFFS_HOST_HARNESS
#define CDP_PROVE_HOST_HARNESS
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
struct dim3_t { unsigned x; };
static thread_local dim3_t blockIdx{0}, blockDim{1}, threadIdx{0};  // the prover's lanes are host threads
namespace cdp {
static const uint32_t FR_R[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
struct round_expand_params_t { uint32_t n, h, spp, cpp, mode; };
struct vcoef_params_t { uint32_t ell, n, m, big_n, vw, o_R, o_S, o_T, o_U, o_M, o_P, exact_eq, vch; };
}
#include "../../include/cdp_msm.h"
#include "../../curdleproofs_b200/csrc/k_transcript.cu"
#include "../../curdleproofs_b200/csrc/k_vcoeffs.cu"
#include "../../curdleproofs_b200/csrc/k_prove.cu"
#include "../../curdleproofs_b200/host/rng.hpp"

extern "C" {
int oracle_msm(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t out_jac[144], int threads);
int oracle_scalar_mul_batch(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out_affine);
int oracle_fold(const uint8_t *L, const uint8_t *R, const uint8_t gamma[32], size_t n, uint8_t *out_affine);
int oracle_normalize_batch(const uint8_t *jac, size_t n, uint8_t *out_affine);
int oracle_compress(const uint8_t *affine, size_t n, uint8_t *out48);
int oracle_decompress(const uint8_t *in48, size_t n, uint8_t *out_affine, int check_subgroup);
}

struct cdp_ctx {
    std::string err = "ok";
    uint64_t launches = 0;
};
struct cdp_fixed_table {
    std::vector<uint8_t> bases;
    size_t n = 0;
};

extern "C" {

int cdp_ctx_create(cdp_ctx **out, int, void *) { *out = new cdp_ctx(); return CDP_OK; }
void cdp_ctx_destroy(cdp_ctx *c) { delete c; }
const char *cdp_last_error(const cdp_ctx *c) { return c ? c->err.c_str() : "null"; }
uint64_t cdp_launch_count(const cdp_ctx *c) { return c ? c->launches : 0; }
int cdp_ctx_device(const cdp_ctx *) { return 0; }
int cdp_sync(cdp_ctx *) { return CDP_OK; }
void *cdp_dev_alloc(cdp_ctx *, size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void cdp_dev_free(cdp_ctx *, void *p) { free(p); }
void *cdp_host_alloc(cdp_ctx *, size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void cdp_host_free(cdp_ctx *, void *p) { free(p); }
int cdp_h2d(cdp_ctx *, void *d, const void *h, size_t n) { memcpy(d, h, n); return CDP_OK; }
int cdp_d2h(cdp_ctx *, void *h, const void *d, size_t n) { memcpy(h, d, n); return CDP_OK; }
int cdp_dev_zero(cdp_ctx *, void *d, size_t n) { memset(d, 0, n); return CDP_OK; }
int cdp_host_is_pinned(const void *) { return 0; }
int cdp_h2d_2d(cdp_ctx *, void *d, size_t dp, const void *h, size_t sp, size_t w, size_t rows) {
    for (size_t r = 0; r < rows; r++) memcpy((uint8_t *)d + r * dp, (const uint8_t *)h + r * sp, w);
    return CDP_OK;
}

int cdp_msm(cdp_ctx *, const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t out[144]) { return oracle_msm(pts, sc, n, out, 1); }
int cdp_normalize_batch(cdp_ctx *, const uint8_t *jac, size_t n, uint8_t *out) { return oracle_normalize_batch(jac, n, out); }
int cdp_compress_batch(cdp_ctx *, const uint8_t *jac, size_t n, uint8_t *out) {
    std::vector<uint8_t> aff(96 * n);
    oracle_normalize_batch(jac, n, aff.data());
    return oracle_compress(aff.data(), n, out);
}
int cdp_decompress_batch(cdp_ctx *, const uint8_t *comp, size_t n, uint8_t *out, uint8_t *status) {
    int rc = CDP_OK;
    for (size_t i = 0; i < n; i++) {
        const int bad = oracle_decompress(comp + 48 * i, 1, out + 96 * i, 1);
        if (status) status[i] = bad ? 1 : 0;
        if (bad) rc = CDP_ERR_NOT_ON_CURVE;
    }
    return rc;
}
int cdp_scalar_mul_batch(cdp_ctx *, const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t *out) { return oracle_scalar_mul_batch(pts, sc, n, out); }

int cdp_fixed_table_create(cdp_ctx *, const uint8_t *pts, size_t n, int, cdp_fixed_table **out) {
    cdp_fixed_table *t = new cdp_fixed_table();
    t->bases.assign(pts, pts + 96 * n);
    t->n = n;
    *out = t;
    return CDP_OK;
}
void cdp_fixed_table_destroy(cdp_ctx *, cdp_fixed_table *t) { delete t; }
size_t cdp_fixed_table_bytes(const cdp_fixed_table *t) { return t ? t->bases.size() : 0; }
size_t cdp_fixed_table_bases(const cdp_fixed_table *t) { return t ? t->n : 0; }
int cdp_msm_fixed(cdp_ctx *, const cdp_fixed_table *t, size_t base_off, const uint8_t *sc, size_t n, uint8_t out[144]) {
    return oracle_msm(t->bases.data() + 96 * base_off, sc, n, out, 1);
}
// the pair selection of a fixed-base segment, as csrc/k_fixed.cu's `fetch` reads it
int cdp_msm_fixed_batch_dev(cdp_ctx *c, const cdp_fixed_table *t, const uint8_t *sc, const cdp_fixed_seg *segs, size_t count, size_t,
                            const uint8_t *var_pts, uint8_t *out_jac) {
    c->launches++;
    static const uint8_t one[32] = {1};
    for (size_t s = 0; s < count; s++) {
        const cdp_fixed_seg &g = segs[s];
        std::vector<uint8_t> P, S;
        auto push = [&](const uint8_t *pt, const uint8_t *k) { P.insert(P.end(), pt, pt + 96); S.insert(S.end(), k, k + 32); };
        for (uint32_t i = 0; i < g.n; i++) {
            uint32_t j = i;
            if (g.sel_h) {
                const uint32_t lo = i & (g.sel_h - 1);
                j = ((i - lo) << 1) | lo | g.sel_val;
            }
            const uint32_t pj = g.pos_off + j * (g.pos_stride ? g.pos_stride : 1u);
            const uint32_t b = g.base_off + pj + (pj >= g.remap_from ? g.remap_delta : 0u);
            if (b >= t->n) { c->err = "mock: base out of range"; return CDP_ERR_INVALID_ARG; }
            push(t->bases.data() + 96 * (size_t)b, sc + 32 * ((size_t)g.scalars_off + j));
        }
        if (g.extra_base) push(t->bases.data() + 96 * (size_t)(g.extra_base - 1), sc + 32 * ((size_t)g.scalars_off + g.extra_scalar));
        for (uint32_t a = 0; a < g.addv_n; a++) push(var_pts + 96 * ((size_t)g.addv_off + a), one);
        oracle_msm(P.data(), S.data(), P.size() / 96, out_jac + 144 * (size_t)g.out_idx, 1);
    }
    return CDP_OK;
}
int cdp_msm_fixed_batch_dev_lanes(cdp_ctx *c, const cdp_fixed_table *t, const uint8_t *sc, const cdp_fixed_seg *segs, size_t count, size_t pairs,
                                  const uint8_t *var_pts, uint8_t *out_jac, int) {
    return cdp_msm_fixed_batch_dev(c, t, sc, segs, count, pairs, var_pts, out_jac);
}
int cdp_msm_fixed_batch_dev_tree(cdp_ctx *c, const cdp_fixed_table *t, const uint8_t *sc, const cdp_fixed_seg *segs, size_t count, size_t pairs,
                                 const uint8_t *var_pts, uint8_t *out_jac, size_t) {
    return cdp_msm_fixed_batch_dev(c, t, sc, segs, count, pairs, var_pts, out_jac);
}
int cdp_msm_batch_dev(cdp_ctx *c, const uint8_t *pts, const uint8_t *sc, const cdp_msm_seg *segs, size_t count, size_t, size_t, uint8_t *out_jac) {
    c->launches++;
    for (size_t s = 0; s < count; s++) {
        const cdp_msm_seg &g = segs[s];
        std::vector<uint8_t> P(pts + 96 * (size_t)g.pts_off, pts + 96 * ((size_t)g.pts_off + g.n)), S(sc + 32 * (size_t)g.scalars_off, sc + 32 * ((size_t)g.scalars_off + g.n));
        if (g.extra) {
            P.insert(P.end(), pts + 96 * (size_t)(g.extra - 1), pts + 96 * (size_t)g.extra);
            S.insert(S.end(), sc + 32 * ((size_t)g.scalars_off + g.n), sc + 32 * ((size_t)g.scalars_off + g.n + 1));
        }
        oracle_msm(P.data(), S.data(), P.size() / 96, out_jac + 144 * s, 1);
    }
    return CDP_OK;
}
int cdp_normalize_dev(cdp_ctx *c, const uint8_t *jac, size_t n, uint8_t *out_aff, uint8_t *out_comp) {
    c->launches++;
    std::vector<uint8_t> aff(96 * n);
    oracle_normalize_batch(jac, n, aff.data());
    if (out_aff) memcpy(out_aff, aff.data(), 96 * n);
    if (out_comp) oracle_compress(aff.data(), n, out_comp);
    return CDP_OK;
}
int cdp_smul_jobs_dev(cdp_ctx *c, uint8_t *pts, const uint8_t *sc, const cdp_smul_job *jobs, size_t n_jobs, size_t epj) {
    c->launches++;
    std::vector<uint8_t> res(96 * n_jobs * epj);
    for (size_t j = 0; j < n_jobs; j++)
        for (size_t e = 0; e < epj; e++) {
            const cdp_smul_job &b = jobs[j];
            const uint8_t *k = sc + 32 * ((size_t)b.scalar_off + e * b.scalar_stride), *src = pts + 96 * ((size_t)b.src_off + e);
            uint8_t *o = res.data() + 96 * (j * epj + e);
            if (b.add_off != CDP_NONE) oracle_fold(pts + 96 * ((size_t)b.add_off + e), src, k, 1, o);
            else oracle_scalar_mul_batch(src, k, 1, o);
        }
    for (size_t j = 0; j < n_jobs; j++) memcpy(pts + 96 * (size_t)jobs[j].out_off, res.data() + 96 * j * epj, 96 * epj);
    return CDP_OK;
}
int cdp_gather_dev(cdp_ctx *c, uint8_t *pts, const uint8_t *src, const uint32_t *si, const uint32_t *di, size_t n) {
    c->launches++;
    for (size_t i = 0; i < n; i++) memmove(pts + 96 * (size_t)di[i], src + 96 * (size_t)si[i], 96);
    return CDP_OK;
}
int cdp_compress_affine_dev(cdp_ctx *c, const uint8_t *pts, const uint32_t *idx, size_t n, uint8_t *out) {
    c->launches++;
    for (size_t i = 0; i < n; i++) oracle_compress(pts + 96 * (size_t)(idx ? idx[i] : i), 1, out + 48 * i);
    return CDP_OK;
}
int cdp_transcript_open_dev(cdp_ctx *c, const uint8_t *cv, const uint8_t *cm, size_t ell, size_t B, uint8_t *va, uint8_t *st) {
    c->launches++;
    blockDim.x = 1; threadIdx.x = 0;
    for (size_t pr = 0; pr < B; pr++) {
        blockIdx.x = (unsigned)pr;
        cdp::k_transcript_open(cv, cm, (uint32_t)ell, (uint32_t)B, va, reinterpret_cast<uint64_t *>(st));
    }
    return CDP_OK;
}
int cdp_round_expand_dev(cdp_ctx *c, const uint8_t *cmp, const uint8_t *ucan, size_t n, size_t h, size_t spp, size_t cpp, int mode, size_t B, uint8_t *out) {
    c->launches++;
    const cdp::round_expand_params_t P = {(uint32_t)n, (uint32_t)h, (uint32_t)spp, (uint32_t)cpp, (uint32_t)mode};
    for (size_t pr = 0; pr < B; pr++)
        cdp::round_expand_thread((uint32_t)pr, 0, 1, reinterpret_cast<const uint32_t *>(cmp), reinterpret_cast<const uint32_t *>(ucan), P, reinterpret_cast<uint32_t *>(out));
    return CDP_OK;
}
size_t cdp_prove_work_scalars(size_t ell) { return 11 * (ell + 4) + 64; }
size_t cdp_prove_random_scalars(size_t ell) { return 3 * (ell + 4) + 11; }
// the device generator (k_prove_random) stood in for by the host's StdRng on the same key and stream position
int cdp_prove_random_dev(cdp_ctx *c, const uint8_t *keys, const uint64_t *skip, size_t batch, size_t ell, uint8_t *out) {
    c->launches++;
    const size_t n = ell + 4, nrnd = 3 * n + 11;
    for (size_t pr = 0; pr < batch; pr++) {
        cdp_host::StdRng rng(cdp_host::StdRng::from_key_t{}, keys + 32 * pr);
        if (skip) rng.skip_words(skip[pr]);
        uint64_t *w = reinterpret_cast<uint64_t *>(out + pr * nrnd * 32);
        auto draw = [&](size_t slot) { cdp_host::Fr x = rng.fr_rand(); memcpy(w + 4 * slot, x.v, 32); };
        for (size_t i = 0; i < 6 + 2 * n - 2; i++) draw(i);
        memset(w + 4 * (6 + 2 * n - 2), 0, 64);
        for (size_t i = 0; i < 5 + n; i++) draw(6 + 2 * n + i);
    }
    return CDP_OK;
}
int cdp_prove_stage_dev(cdp_ctx *c, const cdp_prove_dev *P, int stage, unsigned round) {
    c->launches++;
    for (uint32_t pr = 0; pr < P->batch; pr++) {
        cdp::prove::harness_proof_index = pr;
        cdp::k_prove_stage(*P, stage, round);
    }
    return CDP_OK;
}

// the division-step and the binary Euclidean inversion of csrc/fr256.cuh against the Fermat ladder: edge values and `count` pseudo-random ones; returns the number of mismatches
int mock_check_fr_inverse(int count) {
    using namespace cdp::vcoef;
    int bad = 0;
    uint64_t s = 0x9E3779B97F4A7C15ULL;
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    for (int t = 0; t < count + 6; t++) {
        fr_t x = fr_zero();
        if (t == 0) x.v[0] = 1;
        else if (t == 1) x.v[0] = 2;
        else if (t == 2) { for (int i = 0; i < 8; i++) x.v[i] = fr_mod(i); x.v[0] -= 1; }
        else if (t == 3) x = fr_one();
        else if (t == 4) x.v[7] = 0x40000000u;
        else if (t == 5) { /* zero */ }
        else { for (int i = 0; i < 8; i++) x.v[i] = next(); x.v[7] &= 0x3FFFFFFFu; }
        const fr_t a = fr_inverse(x), b = fr_inverse_euclid(x), c = fr_inverse_safegcd(x);
        bool same = true;
        for (int i = 0; i < 8; i++) same = same && a.v[i] == b.v[i] && a.v[i] == c.v[i];
        if (t != 5) {
            const fr_t one = fr_mul(x, b), want = fr_one();
            for (int i = 0; i < 8; i++) same = same && one.v[i] == want.v[i];
        }
        bad += !same;
    }
    return bad;
}

}  // extern "C"
