// Batched Curdleproofs prover: host driver above the C ABI of include/cdp_msm.h.
//
// Restates `CurdleproofsProof::new` (/root/reference/src/curdleproofs.rs:59-184) and the provers it calls
//   SamePermutationProof::new   src/same_permutation_argument.rs:40-99
//   GrandProductProof::new      src/grand_product_argument.rs:43-166
//   InnerProductProof::new      src/inner_product_argument.rs:98-198
//   SameScalarProof::new        src/same_scalar_argument.rs:39-84
//   SameMultiscalarProof::new   src/same_multiscalar_argument.rs:54-150
// for B independent shuffles advancing in lock-step.  GPU-first restructuring (results are the same group elements):
//   * every point that the reference appends to the transcript or writes into the proof is expressed as ONE msm over
//     device-resident affine bases (the CRS, the instance vectors, the folded working vectors) with host-computed
//     scalars -- e.g. B = A + alpha*M + beta*sum(G) is msm(G|H|M, (a+beta)|r_a|alpha); L_C = msm(G_R, c_L) + ip*H is one
//     msm with H as an extra base.  The host therefore never touches a curve point: it receives 48-byte encodings.
//   * all prover randomness is drawn up front in the reference's order (the draws do not depend on the transcript).
//   * every MSM whose bases are CRS points goes through the fixed-base digit table (cdp_fixed_table, 12 GiB in HBM at
//     ell = 252).  That includes the IPA round MSMs over G and G' = u o G and the SameMSM round MSMs over G_with_blinders:
//     the round-k folded base is G^(k)_i = sum_{j = i mod n_k} w_k(j) G_j with w_k(j) = prod_l gamma_l^{bit_l(j)} (the
//     identity the reference's verifier uses, src/inner_product_argument.rs:202-250), so msm(G^(k)_R, c_L) is an MSM over
//     the ORIGINAL bases with scalars c_L[j mod h] * w_k(j) -- the G, G' and G_with_blinders vectors are never folded
//     (3(n-1) + n of the reference's 5(n-1) + n fold scalar-muls disappear); only T and U, which are per-proof, are.
//     Those n scalars per vector are an outer product (prefix weight x folded vector entry): the host uploads the factors and
//     cdp_round_expand_dev forms the products on the device (CDP_PROVE_HOST_EXPAND=1: on the host).
//   * per round: one batched MSM launch (+ one fixed-base launch) over all proofs, one normalise+compress, one D2H,
//     host transcripts in parallel over proofs, one batched fold launch for T and U.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/cdp_prover.h"
#pragma GCC visibility pop
#include "crs_table.hpp"
#include "merlin.hpp"
#include "rng.hpp"

using namespace cdp_host;

namespace {

constexpr size_t NBL = 4;  // N_BLINDERS, src/lib.rs:35

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct SubLaunch {
    size_t K = 0;      // segments per proof
    size_t max_n = 0;  // largest segment (including the extra base)
    size_t pairs_per_proof = 0;
    std::vector<cdp_msm_seg> segs;  // max_batch * K, proof-major
    cdp_msm_seg *d_segs = nullptr;
    bool fixed = false;               // segments over the CRS digit table
    bool uniform = false;             // fixed: all segments (nearly) equally long -> 16 lanes per segment when many are in flight
    std::vector<cdp_fixed_seg> fsegs;
    cdp_fixed_seg *d_fsegs = nullptr;
};
struct MsmStage {
    size_t scalars_per_proof = 0;
    std::vector<SubLaunch> subs;
    std::vector<std::pair<int, int>> where;  // logical output q -> (sub, k)
    size_t aff_region = (size_t)-1;          // when set: the stage's normalised points are also stored (affine) at this point index
    size_t outputs() const { return where.size(); }
};
struct FoldStage {
    size_t J = 0, epj = 0, scalars_per_proof = 0;
    std::vector<cdp_smul_job> jobs;
    cdp_smul_job *d_jobs = nullptr;
};

// per-proof host state
struct ProofState {
    std::unique_ptr<Transcript> tr;
    std::vector<Fr> vec_a, a_perm, factors, c, d, r_c, r_d, u, x, r_sm;
    // fold weights of the original bases: G (gamma), G' (u, gamma^-1), G_with_blinders (gamma).  Kept as CANONICAL integers in the Fr
    // container: a Montgomery product of a canonical value and a Montgomery-form value is the canonical product, so `weight * scalar`
    // is already the byte form the MSM wants (no separate conversion), and `weight * gamma` stays canonical.
    std::vector<Fr> wG, wGp, wS;
    // the same weights per PREFIX of the index (the bits above the current split), for the device-side expansion: WcP, WsP canonical, WdP Montgomery
    std::vector<Fr> WcP, WdP, WsP;
    Fr a_bl[2], c_bl[4], r_t, r_u, r_a, r_b, r_k, k, m_bl[4], b_bl[4], rb_alpha[4];
    Fr alpha_sp, beta_sp, gprod_result, alpha_g, beta_g, r_p, z, alpha_i, beta_i;
    Fr z_k, z_t, z_u, c_final, d_final, x_final;
    std::vector<uint32_t> perm;
    // compressed points collected for the proof
    uint8_t pts1[15][48];                       // stage-1 outputs
    uint8_t B[48], C[48], D[48], B_c[48], B_d[48];
    uint8_t M_comp[48];
    std::vector<uint8_t> ipa_rounds, sm_rounds;  // m * 4 * 48, m * 6 * 48
};

// proof-level names of the 15 points produced before the IPA rounds
enum ProofPt { O_A = 0, O_R, O_S, O_T1, O_T2, O_U1, O_U2, O_A1, O_A2, O_B1, O_B2, O_AP, O_BA, O_BT, O_BU };
// stage 1: the six full-size MSMs (their affine results are kept in HBM for stage 2) and four single scalar-muls
enum Stage1Out { S1_A = 0, S1_R, S1_S, S1_BA, S1_BT, S1_BU, S1_T1, S1_U1, S1_A1, S1_B1, S1_COUNT };
// stage 2: short combinations of points that already exist (B plus the second halves of the four GroupCommitments and A')
enum Stage2Out { S2_B = 0, S2_T2, S2_A2, S2_U2, S2_B2, S2_AP, S2_COUNT };
// per-proof scratch block X of gathered affine points feeding the short MSMs
enum XSlot { X_A = 0, X_M, X_GSUM, X_R, X_H1, X_S, X_H2, X_A2, X_GT, X_GU, X_B, X_GSUM2, X_HSUM, X_COUNT };

}  // namespace

struct Lane {
    cdp_ctx *ctx = nullptr;
    const cdp_fixed_table *table = nullptr;  // digit table of the CRS points (shared by all lanes of a prover)
    size_t ell = 0, n = 0, m = 0, max_batch = 0;
    int threads = 1;
    std::string err = "ok";
    double timing[4] = {0, 0, 0, 0};
    uint64_t h2d_bytes = 0, d2h_bytes = 0;  // of the last cdp_prove_batch call

    // device point array: [CRS block | per-proof working blocks]
    size_t crs_n = 0, PW = 0;
    size_t o_T = 0, o_U = 0, o_R = 0, o_S = 0, o_X = 0;
    size_t reg1 = 0, reg2 = 0;  // affine outputs of stage 1 / stage 2 (in launch order), kept for the next stage
    uint32_t *d_x1src = nullptr, *d_x1dst = nullptr, *d_x2src = nullptr, *d_x2dst = nullptr;
    size_t xtab_B = 0;          // batch size the x1/x2 gather tables were built for
    uint8_t *d_pts = nullptr;
    uint8_t *d_in = nullptr;       // staging for the instance vectors of a batch (R,S,T,U: 4*ell per proof) + M affine
    uint8_t *d_Mjac = nullptr;
    uint32_t *d_gsrc = nullptr, *d_gdst = nullptr;      // working-vector assembly from the CRS block
    uint32_t *d_isrc = nullptr, *d_idst = nullptr;      // working-vector assembly from the instance staging
    size_t g_count_per_proof = 0, i_count_per_proof = 0;
    uint32_t *d_cidx = nullptr;                         // indices of the points compressed for the transcript opening
    uint8_t *d_scal = nullptr, *d_fscal = nullptr;
    // device-side expansion of the round scalars (cdp_round_expand_dev): compact per-proof blocks (prefix weights + folded vectors), u canonical
    uint8_t *d_cmp = nullptr, *h_cmp = nullptr, *d_ucan = nullptr, *h_ucan = nullptr;
    bool dev_expand = true;  // CDP_PROVE_HOST_EXPAND=1: the n products per vector per round on the host (the older path)
    uint8_t *d_veca = nullptr, *d_tstate = nullptr, *h_veca = nullptr, *h_tstate = nullptr;  // device-side transcript opening
    uint8_t *d_jac = nullptr, *d_comp = nullptr;
    // device-side prover (cdp_prove_stage_dev): the transcript and the Fr algebra of every step run on the GPU, the host only stages inputs
    // and draws the randomness.  CDP_PROVE_HOST_TRANSCRIPT=1 keeps the older host path (same proofs).
    bool dev_prove = true;
    uint8_t *d_comp0 = nullptr;    // encodings of the instance (4 ell per proof, then one M per proof): read again by same_msm_step1
    uint8_t *d_compH = nullptr;    // encoding of crs.H
    uint32_t *d_perm = nullptr, *h_perm = nullptr;
    uint8_t *d_keys = nullptr, *h_keys = nullptr;  // per proof: 32-byte ChaCha12 key | u64 stream position (cdp_prove_random_dev)
    uint8_t *d_wit = nullptr, *h_wit = nullptr, *d_rnd = nullptr, *h_rnd = nullptr, *d_work = nullptr, *d_side = nullptr, *d_proofs = nullptr,
            *h_proofs = nullptr;
    uint8_t *h_scal = nullptr, *h_fscal = nullptr, *h_comp = nullptr, *h_in = nullptr;
    size_t max_scalars_pp = 0, max_out_pp = 0;
    size_t staged_batch = 0;
    uint8_t H_comp[48];

    MsmStage st1, st2, st3, st4;
    std::vector<MsmStage> st_ipa, st_sm;
    std::vector<FoldStage> f_sm;  // T, U folds (and, from the switch round on, the materialised G_with_blinders)
    // switch round k0 (device-side prover only): rounds k >= k0 run over MATERIALISED folded bases (st_ipa_mat / st_sm_mat produce them, n >> k0
    // points per vector, kept affine in regM / regS), see cdp_prove_dev::switch_round; k0 = m: never
    size_t k0 = 0, regM = 0, regS = 0;
    MsmStage st_ipa_mat, st_sm_mat;
    std::vector<FoldStage> f_ipa;

    std::vector<ProofState> ps;
};

namespace {

int perr(Lane *p, int code, const std::string &msg) {
    p->err = msg;
    return code;
}
#define PTRY(expr)                                                                                     \
    do {                                                                                               \
        int rc__ = (expr);                                                                             \
        if (rc__ != CDP_OK) return perr(p, rc__, std::string(#expr) + ": " + cdp_last_error(p->ctx));  \
    } while (0)

template <class F>
void parallel_for(int threads, size_t n, F f) {
    if (threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    size_t nt = std::min<size_t>((size_t)threads, n);
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (size_t t = 0; t < nt; t++)
        pool.emplace_back([=]() {
            for (size_t i = t; i < n; i += nt) f(i);
        });
    for (auto &th : pool) th.join();
}

// contiguous ranges, one per thread: lets a thread batch work (simultaneous inversion) over its own proofs
template <class F>
void parallel_chunks(int threads, size_t n, F f) {
    size_t nt = std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n));
    if (nt <= 1) {
        f((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (size_t t = 0; t < nt; t++)
        pool.emplace_back([=]() { f(n * t / nt, n * (t + 1) / nt); });
    for (auto &th : pool) th.join();
}
// xs[i] <- xs[i]^-1 for all i with one field inversion (Montgomery's trick); no element may be zero
void batch_inverse(std::vector<Fr> &xs) {
    const size_t n = xs.size();
    if (n == 0) return;
    std::vector<Fr> pre(n);
    Fr acc = Fr::one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc *= xs[i]; }
    Fr inv = acc.inverse();
    for (size_t i = n; i-- > 0;) {
        Fr t = inv * pre[i];
        inv *= xs[i];
        xs[i] = t;
    }
}

void put_fr(uint8_t *dst, const Fr &x) { x.to_bytes(dst); }
// values held as canonical integers inside the Fr container (the fold weights): see ProofState
void put_canonical(uint8_t *dst, const Fr &x) { memcpy(dst, x.v, 32); }
Fr fr_canonical_one() { uint64_t one[4] = {1, 0, 0, 0}; return Fr::raw(one); }
Fr fr_to_canonical_value(const Fr &x) { uint64_t c[4]; x.to_canonical(c); return Fr::raw(c); }

// ---- stage table construction --------------------------------------------------------------------------------
struct SegSpec {
    size_t pts_rel = 0;    // variable-base: offset inside the proof block, or absolute when `absolute`
    bool absolute = false;
    size_t scal_rel = 0;   // offset inside the proof's scalar block
    size_t n = 0;
    long extra_abs = -1;   // variable-base: absolute index of the extra base, -1 none
    // fixed-base (CRS digit table): pts_rel = first table base of the range; see cdp_fixed_seg for the rest
    bool fixed = false;
    size_t sel_h = 0, sel_val = 0, remap_from = 0xFFFFFFFFu, remap_delta = 0;
    long fextra_base = -1;
    size_t fextra_scalar = 0;
    size_t addv_rel = 0, addv_n = 0;  // device-resident points of the proof block added with coefficient 1
    size_t pos_off = 0, pos_stride = 0;  // fixed-base: strided walk over the bases (cdp_fixed_seg)
    // variable-base: points of a stage region (the affine outputs of an earlier stage, K_region per proof) instead of the proof block
    long region_base = -1;
    size_t region_stride = 0;
};
SegSpec fix_seg(size_t base_off, size_t scal_rel, size_t n) {
    SegSpec s;
    s.pts_rel = base_off; s.scal_rel = scal_rel; s.n = n; s.fixed = true;
    return s;
}
// the positions j of a 0..2*cnt-1 range (every other run of h) with (j & h) == val
SegSpec fix_half(size_t base_off, size_t scal_rel, size_t cnt, size_t h, size_t val) {
    SegSpec s = fix_seg(base_off, scal_rel, cnt);
    s.sel_h = h; s.sel_val = val;
    return s;
}
void build_stage(Lane *p, MsmStage &st, const std::vector<SegSpec> &specs, size_t scalars_pp) {
    st.scalars_per_proof = scalars_pp;
    // variable-base: two size classes keep the tiny (1-3 point) MSMs out of the big-CTA launch; fixed-base: one more sub-launch
    auto eff = [](const SegSpec &s) { return s.n + ((s.fixed ? s.fextra_base : s.extra_abs) >= 0 ? 1 : 0); };
    size_t big = 0;
    bool any_var = false, any_fixed = false;
    for (auto &s : specs) {
        if (s.fixed) { any_fixed = true; continue; }
        any_var = true;
        big = std::max(big, eff(s));
    }
    std::vector<int> cls(specs.size(), 0);
    bool split = false;
    if (big >= 12) {
        for (size_t i = 0; i < specs.size(); i++)
            if (!specs[i].fixed && eff(specs[i]) * 8 <= big && eff(specs[i]) < 12) { cls[i] = 1; split = true; }
    }
    const int n_var = any_var ? (split ? 2 : 1) : 0;
    // fixed-base: the long, equally long segments (>= 64 pairs: what the tree path of cdp_msm_fixed_batch_dev_tree wants) apart from the short ones
    size_t flo = (size_t)-1, fhi = 0, n_long = 0, n_fixed = 0;
    for (auto &s : specs) {
        if (!s.fixed) continue;
        n_fixed++;
        if (eff(s) >= 64) { n_long++; flo = std::min(flo, eff(s)); fhi = std::max(fhi, eff(s)); }
    }
    const bool fsplit = n_long > 0 && n_long < n_fixed && fhi <= flo + 1;
    for (size_t i = 0; i < specs.size(); i++)
        if (specs[i].fixed) cls[i] = n_var + (fsplit && eff(specs[i]) < 64 ? 1 : 0);
    st.subs.assign(n_var + (any_fixed ? (fsplit ? 2 : 1) : 0), SubLaunch());
    if (any_fixed) st.subs[n_var].fixed = true;
    if (fsplit) st.subs[n_var + 1].fixed = true;
    st.where.resize(specs.size());
    for (size_t i = 0; i < specs.size(); i++) {
        SubLaunch &sl = st.subs[cls[i]];
        st.where[i] = {cls[i], (int)sl.K};
        sl.K++;
        sl.max_n = std::max(sl.max_n, eff(specs[i]));
        sl.pairs_per_proof += eff(specs[i]);
    }
    for (size_t c = 0; c < st.subs.size(); c++) {
        SubLaunch &sl = st.subs[c];
        if (sl.fixed) {
            size_t lo = (size_t)-1, hi = 0;
            for (size_t i = 0; i < specs.size(); i++)
                if (cls[i] == (int)c) { lo = std::min(lo, eff(specs[i])); hi = std::max(hi, eff(specs[i])); }
            sl.uniform = hi <= lo + 1 && lo >= 16;
        }
        if (sl.fixed) sl.fsegs.resize(p->max_batch * sl.K);
        else sl.segs.resize(p->max_batch * sl.K);
        for (size_t pr = 0; pr < p->max_batch; pr++) {
            size_t bp = p->crs_n + pr * p->PW;
            for (size_t i = 0; i < specs.size(); i++) {
                if (cls[i] != (int)c) continue;
                const SegSpec &sp = specs[i];
                const size_t slot = pr * sl.K + st.where[i].second;
                if (sl.fixed) {
                    cdp_fixed_seg &fs = sl.fsegs[slot];
                    memset(&fs, 0, sizeof fs);
                    fs.base_off = (uint32_t)sp.pts_rel;
                    fs.scalars_off = (uint32_t)(pr * scalars_pp + sp.scal_rel);
                    fs.n = (uint32_t)sp.n;
                    fs.sel_h = (uint32_t)sp.sel_h; fs.sel_val = (uint32_t)sp.sel_val;
                    fs.remap_from = (uint32_t)sp.remap_from; fs.remap_delta = (uint32_t)sp.remap_delta;
                    fs.extra_base = sp.fextra_base >= 0 ? (uint32_t)(sp.fextra_base + 1) : 0;
                    fs.extra_scalar = (uint32_t)sp.fextra_scalar;
                    fs.out_idx = (uint32_t)slot;
                    fs.addv_off = (uint32_t)(bp + sp.addv_rel); fs.addv_n = (uint32_t)sp.addv_n;
                    fs.pos_off = (uint32_t)sp.pos_off; fs.pos_stride = (uint32_t)sp.pos_stride;
                } else {
                    cdp_msm_seg &sg = sl.segs[slot];
                    sg.pts_off = (uint32_t)(sp.region_base >= 0 ? (size_t)sp.region_base + pr * sp.region_stride + sp.pts_rel
                                            : sp.absolute ? sp.pts_rel : bp + sp.pts_rel);
                    sg.scalars_off = (uint32_t)(pr * scalars_pp + sp.scal_rel);
                    sg.n = (uint32_t)sp.n;
                    sg.extra = sp.extra_abs >= 0 ? (uint32_t)(sp.extra_abs + 1) : 0;
                }
            }
        }
    }
    p->max_scalars_pp = std::max(p->max_scalars_pp, scalars_pp);
    p->max_out_pp = std::max(p->max_out_pp, specs.size());
}
struct JobSpec {
    size_t src_rel, add_rel, out_rel;
    bool has_add;
    size_t scal_rel, stride;
    long region_base = -1;      // offsets relative to a stage region (region_base + pr * region_stride) instead of the proof block
    size_t region_stride = 0;
};
void build_fold(Lane *p, FoldStage &fs, const std::vector<JobSpec> &specs, size_t epj, size_t scalars_pp) {
    fs.J = specs.size();
    fs.epj = epj;
    fs.scalars_per_proof = scalars_pp;
    fs.jobs.resize(p->max_batch * fs.J);
    for (size_t pr = 0; pr < p->max_batch; pr++) {
        for (size_t j = 0; j < fs.J; j++) {
            const size_t bp = specs[j].region_base >= 0 ? (size_t)specs[j].region_base + pr * specs[j].region_stride : p->crs_n + pr * p->PW;
            cdp_smul_job &jb = fs.jobs[pr * fs.J + j];
            memset(&jb, 0, sizeof jb);
            jb.src_off = (uint32_t)(bp + specs[j].src_rel);
            jb.add_off = specs[j].has_add ? (uint32_t)(bp + specs[j].add_rel) : CDP_NONE;
            jb.out_off = (uint32_t)(bp + specs[j].out_rel);
            jb.scalar_off = (uint32_t)(pr * scalars_pp + specs[j].scal_rel);
            jb.scalar_stride = (uint32_t)specs[j].stride;
        }
    }
}

int upload_tables(Lane *p) {
    auto up_stage = [&](MsmStage &st) -> int {
        for (auto &sl : st.subs) {
            const void *src = sl.fixed ? (const void *)sl.fsegs.data() : (const void *)sl.segs.data();
            size_t bytes = sl.fixed ? sl.fsegs.size() * sizeof(cdp_fixed_seg) : sl.segs.size() * sizeof(cdp_msm_seg);
            void *d = cdp_dev_alloc(p->ctx, bytes);
            if (!d) return CDP_ERR_CUDA;
            if (sl.fixed) sl.d_fsegs = (cdp_fixed_seg *)d;
            else sl.d_segs = (cdp_msm_seg *)d;
            int rc = cdp_h2d(p->ctx, d, src, bytes);
            if (rc) return rc;
        }
        return CDP_OK;
    };
    auto up_fold = [&](FoldStage &fs) -> int {
        size_t bytes = fs.jobs.size() * sizeof(cdp_smul_job);
        fs.d_jobs = (cdp_smul_job *)cdp_dev_alloc(p->ctx, bytes);
        if (!fs.d_jobs) return CDP_ERR_CUDA;
        return cdp_h2d(p->ctx, fs.d_jobs, fs.jobs.data(), bytes);
    };
    PTRY(up_stage(p->st1)); PTRY(up_stage(p->st2)); PTRY(up_stage(p->st3)); PTRY(up_stage(p->st4));
    for (auto &s : p->st_ipa) PTRY(up_stage(s));
    for (auto &s : p->st_sm) PTRY(up_stage(s));
    for (auto &f : p->f_sm) PTRY(up_fold(f));
    if (p->k0 < p->m) {
        PTRY(up_stage(p->st_ipa_mat)); PTRY(up_stage(p->st_sm_mat));
        for (auto &f : p->f_ipa)
            if (f.J) PTRY(up_fold(f));
    }
    PTRY(cdp_sync(p->ctx));  // the host vectors are pageable: make sure the copies are done before they can move
    return CDP_OK;
}

// lanes of a warp per fixed-base segment: 16 for a launch of many equally long segments (measured, 4096 proofs over 8 lanes: 731 vs 745 ms
// per step; 8 lanes per segment: 756 ms), a whole warp otherwise
int fixed_lanes(const SubLaunch &sl, size_t B) { return sl.uniform && B * sl.K >= 2048 ? 16 : 32; }
// a launch of many equally long segments sums its table points as a tree of batched affine additions (cdp_msm_fixed_batch_dev_tree);
// CDP_FIXED_TREE=0 keeps the lane kernel, CDP_FIXED_TREE_MIN_SEGS / _MIN_PAIRS move the thresholds
bool fixed_tree(const SubLaunch &sl, size_t B) {
    static const int on = [] { const char *e = getenv("CDP_FIXED_TREE"); return e ? atoi(e) : 1; }();
    static const size_t min_segs = [] { const char *e = getenv("CDP_FIXED_TREE_MIN_SEGS"); return e ? (size_t)atol(e) : (size_t)512; }();
    static const size_t min_pairs = [] { const char *e = getenv("CDP_FIXED_TREE_MIN_PAIRS"); return e ? (size_t)atol(e) : (size_t)64; }();
    return on && sl.uniform && sl.max_n >= min_pairs && B * sl.K >= min_segs;
}
int launch_fixed_sub(Lane *p, const SubLaunch &sl, size_t B, size_t out_off) {
    if (fixed_tree(sl, B))
        return cdp_msm_fixed_batch_dev_tree(p->ctx, p->table, p->d_scal, sl.d_fsegs, B * sl.K, B * sl.pairs_per_proof, p->d_pts, p->d_jac + out_off * 144, sl.max_n);
    return cdp_msm_fixed_batch_dev_lanes(p->ctx, p->table, p->d_scal, sl.d_fsegs, B * sl.K, B * sl.pairs_per_proof, p->d_pts, p->d_jac + out_off * 144, fixed_lanes(sl, B));
}

// ---- running a stage -----------------------------------------------------------------------------------------
// scalars for all proofs are already in p->h_scal (proof-major, st.scalars_per_proof each)
struct ExpandSpec {
    int mode;            // 0 = IPA round (c and d scalars), 1 = SameMSM round (x scalars + the folded vector)
    size_t n, h, cpp;    // vector length, split, compact scalars per proof (in p->h_cmp)
};
int run_msm_stage(Lane *p, MsmStage &st, size_t B, double &t_wait, double &t_copy, const ExpandSpec *ex = nullptr) {
    double t0 = now_ms();
    if (ex) {  // the stage's scalars are formed on the device from the compact blocks in p->h_cmp
        PTRY(cdp_h2d(p->ctx, p->d_cmp, p->h_cmp, B * ex->cpp * 32));
        p->h2d_bytes += B * ex->cpp * 32;
        PTRY(cdp_round_expand_dev(p->ctx, p->d_cmp, p->d_ucan, ex->n, ex->h, st.scalars_per_proof, ex->cpp, ex->mode, B, p->d_scal));
    } else {
        PTRY(cdp_h2d(p->ctx, p->d_scal, p->h_scal, B * st.scalars_per_proof * 32));
        p->h2d_bytes += B * st.scalars_per_proof * 32;
    }
    size_t out_off = 0;
    for (auto &sl : st.subs) {
        if (sl.fixed) PTRY(launch_fixed_sub(p, sl, B, out_off));
        else PTRY(cdp_msm_batch_dev(p->ctx, p->d_pts, p->d_scal, sl.d_segs, B * sl.K, sl.max_n, B * sl.pairs_per_proof, p->d_jac + out_off * 144));
        out_off += B * sl.K;
    }
    PTRY(cdp_normalize_dev(p->ctx, p->d_jac, out_off, st.aff_region != (size_t)-1 ? p->d_pts + st.aff_region * 96 : nullptr, p->d_comp));
    PTRY(cdp_d2h(p->ctx, p->h_comp, p->d_comp, out_off * 48));
    p->d2h_bytes += out_off * 48;
    double t1 = now_ms();
    PTRY(cdp_sync(p->ctx));
    double t2 = now_ms();
    t_copy += t1 - t0;
    t_wait += t2 - t1;
    return CDP_OK;
}
// compressed output q of proof pr after run_msm_stage
const uint8_t *stage_out(const Lane *p, const MsmStage &st, size_t B, size_t pr, size_t q) {
    size_t base = 0;
    for (int s = 0; s < st.where[q].first; s++) base += B * st.subs[s].K;
    const SubLaunch &sl = st.subs[st.where[q].first];
    return p->h_comp + 48 * (base + pr * sl.K + st.where[q].second);
}
// index of output q of proof pr among the stage's outputs (launch order) for a batch of B
size_t stage_out_index(const MsmStage &st, size_t B, size_t pr, size_t q) {
    size_t base = 0;
    for (int s = 0; s < st.where[q].first; s++) base += B * st.subs[s].K;
    return base + pr * st.subs[st.where[q].first].K + st.where[q].second;
}
int run_fold_stage(Lane *p, FoldStage &fs, size_t B) {
    PTRY(cdp_h2d(p->ctx, p->d_fscal, p->h_fscal, B * fs.scalars_per_proof * 32));
    p->h2d_bytes += B * fs.scalars_per_proof * 32;
    PTRY(cdp_smul_jobs_dev(p->ctx, p->d_pts, p->d_fscal, fs.d_jobs, B * fs.J, fs.epj));
    return CDP_OK;
}

}  // namespace

extern "C" size_t cdp_proof_size(size_t ell) {
    size_t n = ell + NBL, m = 0;
    while (((size_t)1 << m) < n) m++;
    return 1088 + 480 * m;
}

static void lane_destroy(Lane *p) {
    if (!p) return;
    cdp_ctx *c = p->ctx;
    auto free_stage = [&](MsmStage &st) { for (auto &sl : st.subs) { cdp_dev_free(c, sl.d_segs); cdp_dev_free(c, sl.d_fsegs); } };
    free_stage(p->st1); free_stage(p->st2); free_stage(p->st3); free_stage(p->st4);
    for (auto &s : p->st_ipa) free_stage(s);
    for (auto &s : p->st_sm) free_stage(s);
    for (auto &f : p->f_sm) cdp_dev_free(c, f.d_jobs);
    free_stage(p->st_ipa_mat); free_stage(p->st_sm_mat);
    for (auto &f : p->f_ipa) cdp_dev_free(c, f.d_jobs);
    for (void *d : {(void *)p->d_pts, (void *)p->d_in, (void *)p->d_Mjac, (void *)p->d_gsrc, (void *)p->d_gdst, (void *)p->d_isrc,
                    (void *)p->d_idst, (void *)p->d_cidx, (void *)p->d_x1src, (void *)p->d_x1dst, (void *)p->d_x2src, (void *)p->d_x2dst, (void *)p->d_scal, (void *)p->d_fscal, (void *)p->d_cmp, (void *)p->d_ucan, (void *)p->d_jac, (void *)p->d_comp, (void *)p->d_veca, (void *)p->d_tstate, (void *)p->d_comp0, (void *)p->d_compH,
                    (void *)p->d_perm, (void *)p->d_wit, (void *)p->d_rnd, (void *)p->d_work, (void *)p->d_side, (void *)p->d_proofs})
        cdp_dev_free(c, d);
    for (void *h : {(void *)p->h_scal, (void *)p->h_fscal, (void *)p->h_cmp, (void *)p->h_ucan, (void *)p->h_comp, (void *)p->h_in, (void *)p->h_veca, (void *)p->h_tstate, (void *)p->h_perm, (void *)p->h_wit, (void *)p->h_rnd, (void *)p->h_keys, (void *)p->h_proofs}) cdp_host_free(c, h);
    delete p;
}

static int lane_create(Lane **out, cdp_ctx *ctx, const cdp_fixed_table *table, size_t ell, const uint8_t *crs_points, size_t max_batch,
                       int host_threads) {
    if (!out || !ctx || !table || !crs_points || max_batch == 0 || ell < 4) return CDP_ERR_INVALID_ARG;
    *out = nullptr;
    size_t n = ell + NBL, m = 0;
    while (((size_t)1 << m) < n) m++;
    if (((size_t)1 << m) != n) return CDP_ERR_INVALID_ARG;  // n must be a power of two (src/inner_product_argument.rs:116)
    if (n + 1 > 2048) return CDP_ERR_TOO_LARGE;
    Lane *p = new Lane();
    p->ctx = ctx; p->table = table; p->ell = ell; p->n = n; p->m = m; p->max_batch = max_batch;
    p->threads = std::max(1, host_threads);
    // ---- device point layout
    // CRS block (also the base order of the digit table, whose first ell + 7 entries are crs_points): G | Hvec | H | G_t | G_u | sum(G) | sum(Hvec)
    const size_t cH = n, cGt = n + 1, cGu = n + 2, cGsum = n + 3, cHsum = n + 4;
    p->crs_n = n + 5;
    // per-proof block: T_with_blinders | U_with_blinders (folded in place) | vec_R | vec_S | X (gathered short-MSM inputs)
    p->o_T = 0; p->o_U = n; p->o_R = 2 * n; p->o_S = 2 * n + ell; p->o_X = 2 * n + 2 * ell;
    p->PW = 2 * n + 2 * ell + X_COUNT;
    p->reg1 = p->crs_n + max_batch * p->PW;
    p->reg2 = p->reg1 + max_batch * S1_COUNT;
    // switch round: CDP_PROVE_SWITCH_LEN = L materialises the folded bases when the vectors are down to L entries and runs the remaining
    // rounds like the reference's (small variable-base MSMs + folds).  OFF by default: measured at ell = 252, 4096 proofs per step
    // (profiles/r02_switch_round.txt): never 731.9 ms, L = 8: 735.6, 16: 745.5, 32: 812.9, 64: 892.0 -- a table pair costs 16 additions
    // whatever the round, while h-point variable-base MSMs, folds and their normalisations are latency-bound launches.  Only the
    // device-side prover knows the folded form.
    {
        bool dev_prove = true;
        if (const char *e = getenv("CDP_PROVE_HOST_TRANSCRIPT")) dev_prove = atoi(e) == 0;
        size_t sw_len = 0;
        if (const char *e = getenv("CDP_PROVE_SWITCH_LEN")) sw_len = (size_t)atoll(e);
        p->k0 = m;
        if (dev_prove && sw_len >= 2 && (sw_len & (sw_len - 1)) == 0 && sw_len * 2 <= n) {
            size_t k0 = 0;
            while ((n >> k0) > sw_len) k0++;
            if (k0 >= 1 && k0 < m) p->k0 = k0;
        }
    }
    const size_t k0 = p->k0, ns = n >> std::min(k0, m);  // ns: entries of a materialised vector
    p->regM = p->reg2 + max_batch * S2_COUNT;            // per proof: G^(k0) (ns) | G'^(k0) (ns)
    p->regS = p->regM + max_batch * 2 * ns;              // per proof: G_with_blinders^(k0) (ns)
    size_t total_pts = p->regS + max_batch * ns;
    if (total_pts >= ((size_t)1 << 31)) { delete p; return CDP_ERR_TOO_LARGE; }
    // G_with_blinders = G | H_0 | H_1 | G_t | G_u (curdleproofs.rs:134-140) in table bases: positions >= ell + 2 skip H_2, H_3, H
    const size_t gs_from = ell + 2, gs_delta = cGt - (ell + 2);

    // ---- stage tables
    {   // stage 1: everything that depends only on vec_a and the prover's own randomness
        // scalars: a_perm|r_a' (n) | vec_a (ell) | r_sm (n) | r_t | r_u | r_a | r_b
        const size_t sa = n, sR = n + ell, sx = 2 * n + ell;
        std::vector<SegSpec> v(S1_COUNT);
        v[S1_A] = fix_seg(0, 0, n);                    // A = msm(G|Hvec, a_perm | r_a')            curdleproofs.rs:93
        v[S1_R] = {p->o_R, false, sa, ell, -1};        // R = msm(vec_R, a)                         :112
        v[S1_S] = {p->o_S, false, sa, ell, -1};        // S = msm(vec_S, a)                         :113
        v[S1_BA] = fix_seg(0, sR, n);                  // B_a = msm(G_with_blinders, r)             same_multiscalar_argument.rs:80
        v[S1_BA].remap_from = gs_from; v[S1_BA].remap_delta = gs_delta;
        v[S1_BT] = {p->o_T, false, sR, n, -1};         // B_t, B_u                                  :81-82
        v[S1_BU] = {p->o_U, false, sR, n, -1};
        v[S1_T1] = fix_seg(cGt, sx, 1);                // cm_T.T_1 = r_t G_t                        :115 / commitments.rs:50
        v[S1_U1] = fix_seg(cGu, sx + 1, 1);            // cm_U.T_1 = r_u G_u                        :116
        v[S1_A1] = fix_seg(cGt, sx + 2, 1);            // cm_A.T_1 = r_a G_t                        same_scalar_argument.rs:60
        v[S1_B1] = fix_seg(cGu, sx + 3, 1);            // cm_B.T_1 = r_b G_u                        :61
        build_stage(p, p->st1, v, sx + 4);
        p->st1.aff_region = p->reg1;
    }
    {   // stage 2: values that are short combinations of points already in HBM (gathered into the proof's X block)
        std::vector<SegSpec> v(S2_COUNT);
        v[S2_B] = {p->o_X + X_A, false, 0, 3, -1};     // B = A + alpha M + beta sum(G)             same_permutation_argument.rs:75-76
        v[S2_T2] = {p->o_X + X_R, false, 3, 2, -1};    // cm_T.T_2 = k R + r_t H                    curdleproofs.rs:115 / commitments.rs:51
        v[S2_A2] = {p->o_X + X_R, false, 5, 2, -1};    // cm_A.T_2 = r_k R + r_a H                  same_scalar_argument.rs:60
        v[S2_U2] = {p->o_X + X_S, false, 7, 2, -1};    // cm_U.T_2 = k S + r_u H
        v[S2_B2] = {p->o_X + X_S, false, 9, 2, -1};    // cm_B.T_2 = r_k S + r_b H
        v[S2_AP] = fix_seg(cGt, 12, 2);                // A' = A + cm_T.T_1 + cm_U.T_1 = A + r_t G_t + r_u G_u   curdleproofs.rs:131
        v[S2_AP].addv_rel = p->o_X + X_A2; v[S2_AP].addv_n = 1;
        build_stage(p, p->st2, v, 14);
        p->st2.aff_region = p->reg2;
    }
    build_stage(p, p->st3, {fix_seg(0, 0, n)}, n);                                 // C     grand_product_argument.rs:76
    SegSpec segD = fix_seg(cGsum, 1, 2);                                           // D = B - beta^-1 sum(G) + alpha sum(Hvec)   grand_product_argument.rs:132
    segD.addv_rel = p->o_X + X_B; segD.addv_n = 1;
    build_stage(p, p->st4, {segD,
                            fix_seg(0, 3, n),                                      // B_c   inner_product_argument.rs:126
                            fix_seg(0, 3 + n, n)},                                 // B_d = msm(G', r_d) = msm(G|Hvec, r_d o u)   :127
                2 * n + 3);
    p->st_ipa.resize(m); p->st_sm.resize(m); p->f_sm.resize(m); p->f_ipa.resize(m);
    const size_t SPPF = 2 * n + 2 * ns + 2;  // scalars per proof of the IPA emissions from the switch on: materialisation (2n) | one folded round (<= 2 ns + 2)
    if (k0 < m) {
        // materialisation: entry i of the folded vector = the Q = 2^k0 original bases i, i + ns, i + 2 ns, ... with the prefix weights (k_prove.cu)
        const size_t Q = (size_t)1 << k0;
        std::vector<SegSpec> vi(2 * ns), vs(ns);
        for (size_t i = 0; i < ns; i++) {
            vi[i] = fix_seg(0, i * Q, Q); vi[i].pos_off = i; vi[i].pos_stride = ns;                    // G^(k0)_i
            vi[ns + i] = fix_seg(0, n + i * Q, Q); vi[ns + i].pos_off = i; vi[ns + i].pos_stride = ns;  // G'^(k0)_i = sum Wd[q] u_j G_j
            vs[i] = fix_seg(0, i * Q, Q); vs[i].pos_off = i; vs[i].pos_stride = ns;                    // G_with_blinders^(k0)_i
            vs[i].remap_from = gs_from; vs[i].remap_delta = gs_delta;
        }
        build_stage(p, p->st_ipa_mat, vi, SPPF);
        p->st_ipa_mat.aff_region = p->regM;
        build_stage(p, p->st_sm_mat, vs, n + ns);
        p->st_sm_mat.aff_region = p->regS;
    }
    for (size_t k = 0; k < m; k++) {
        size_t h = n >> (k + 1);
        if (k < k0) {
            // inner_product_argument.rs:158-161 over the original bases; scalars: cw (n) | ipL | ipR | dw (n)
            //   cw[j] = w(j) c[j mod h] (bit h of j set) or w(j) c[h + j mod h] (clear);  dw[j] = w'(j) d[h + j mod h] (clear) or w'(j) d[j mod h] (set)
            SegSpec LC = fix_half(0, 0, n / 2, h, h), LD = fix_half(0, n + 2, n / 2, h, 0), RC = fix_half(0, 0, n / 2, h, 0),
                    RD = fix_half(0, n + 2, n / 2, h, h);
            LC.fextra_base = (long)cH; LC.fextra_scalar = n;          // L_C = msm(G_R, c_L) + <c_L,d_R> H
            RC.fextra_base = (long)cH; RC.fextra_scalar = n + 1;      // R_C = msm(G_L, c_R) + <c_R,d_L> H
            build_stage(p, p->st_ipa[k], {LC, LD, RC, RD}, 2 * n + 2);
            // same_multiscalar_argument.rs:107-112; scalars: xw (n, G_with_blinders over the original bases) | x_L | x_R (folded T, U)
            SegSpec LA = fix_half(0, 0, n / 2, h, h), RA = fix_half(0, 0, n / 2, h, 0);
            LA.remap_from = RA.remap_from = gs_from; LA.remap_delta = RA.remap_delta = gs_delta;
            build_stage(p, p->st_sm[k], {LA, {p->o_T + h, false, n, h, -1}, {p->o_U + h, false, n, h, -1},
                                         RA, {p->o_T, false, n + h, h, -1}, {p->o_U, false, n + h, h, -1}},
                        n + 2 * h);
            build_fold(p, p->f_sm[k], {{p->o_T + h, p->o_T, p->o_T, true, 0, 0}, {p->o_U + h, p->o_U, p->o_U, true, 0, 0}},   // :128-129
                       h, 1);
        } else {
            // the reference's own form over the materialised vectors (regM: G (ns) | G' (ns); regS: G_with_blinders (ns)), all variable-base:
            //   scalars at 2n: c_L (h) | ipL | c_R (h) | ipR | d (2h)
            auto reg_seg = [&](size_t region, size_t stride, size_t rel, size_t scal_rel, size_t cnt, long extra) {
                SegSpec sgm;
                sgm.pts_rel = rel; sgm.scal_rel = scal_rel; sgm.n = cnt; sgm.extra_abs = extra;
                sgm.region_base = (long)region; sgm.region_stride = stride;
                return sgm;
            };
            const size_t F = 2 * n;
            build_stage(p, p->st_ipa[k], {reg_seg(p->regM, 2 * ns, h, F, h, (long)cH),                      // L_C = msm(G_R, c_L) + ipL H
                                          reg_seg(p->regM, 2 * ns, ns, F + 2 * h + 2 + h, h, -1),           // L_D = msm(G'_L, d_R)
                                          reg_seg(p->regM, 2 * ns, 0, F + h + 1, h, (long)cH),              // R_C = msm(G_L, c_R) + ipR H
                                          reg_seg(p->regM, 2 * ns, ns + h, F + 2 * h + 2, h, -1)},          // R_D = msm(G'_R, d_L)
                        SPPF);
            build_stage(p, p->st_sm[k], {reg_seg(p->regS, ns, h, n, h, -1), {p->o_T + h, false, n, h, -1}, {p->o_U + h, false, n, h, -1},
                                         reg_seg(p->regS, ns, 0, n + h, h, -1), {p->o_T, false, n + h, h, -1}, {p->o_U, false, n + h, h, -1}},
                        n + 2 * h);
            JobSpec jg = {h, 0, 0, true, 0, 0}, jgp = {ns + h, ns, ns, true, 1, 0}, js = {h, 0, 0, true, 0, 0};
            jg.region_base = jgp.region_base = (long)p->regM; jg.region_stride = jgp.region_stride = 2 * ns;
            js.region_base = (long)p->regS; js.region_stride = ns;
            if (h >= 1) build_fold(p, p->f_ipa[k], {jg, jgp}, h, 2);                                        // G_L + gamma G_R, G'_L + gamma^-1 G'_R  (:177-178)
            build_fold(p, p->f_sm[k], {{p->o_T + h, p->o_T, p->o_T, true, 0, 0}, {p->o_U + h, p->o_U, p->o_U, true, 0, 0}, js}, h, 1);
        }
    }

    // ---- device buffers
    bool ok = true;
    auto dalloc = [&](size_t bytes) { void *d = cdp_dev_alloc(ctx, bytes); ok = ok && d; return d; };
    auto halloc = [&](size_t bytes) { void *h = cdp_host_alloc(ctx, bytes); ok = ok && h; return h; };
    p->d_pts = (uint8_t *)dalloc((total_pts + 1) * 96);  // +1: an all-zero point (infinity) used by the gather tables
    size_t in_pp = 4 * ell + 1;  // R,S,T,U of every proof (proof-major), then the affine M of every proof
    p->d_in = (uint8_t *)dalloc(max_batch * in_pp * 96);
    p->d_Mjac = (uint8_t *)dalloc(max_batch * 144);
    p->h_in = (uint8_t *)halloc(max_batch * (4 * ell * 96 + 144));
    size_t max_out = std::max(p->max_out_pp, in_pp);
    p->d_scal = (uint8_t *)dalloc(max_batch * p->max_scalars_pp * 32);
    p->h_scal = (uint8_t *)halloc(max_batch * p->max_scalars_pp * 32);
    p->d_fscal = (uint8_t *)dalloc(max_batch * n * 32);
    p->h_fscal = (uint8_t *)halloc(max_batch * n * 32);
    p->d_cmp = (uint8_t *)dalloc(max_batch * (2 * n + 4) * 32); p->h_cmp = (uint8_t *)halloc(max_batch * (2 * n + 4) * 32);
    p->d_ucan = (uint8_t *)dalloc(max_batch * n * 32); p->h_ucan = (uint8_t *)halloc(max_batch * n * 32);
    if (const char *e = getenv("CDP_PROVE_HOST_EXPAND")) p->dev_expand = atoi(e) == 0;
    p->d_veca = (uint8_t *)dalloc(max_batch * ell * 32); p->h_veca = (uint8_t *)halloc(max_batch * ell * 32);
    p->d_tstate = (uint8_t *)dalloc(max_batch * CDP_TRANSCRIPT_STATE_BYTES); p->h_tstate = (uint8_t *)halloc(max_batch * CDP_TRANSCRIPT_STATE_BYTES);
    p->d_jac = (uint8_t *)dalloc(max_batch * p->max_out_pp * 144);
    if (const char *e = getenv("CDP_PROVE_HOST_TRANSCRIPT")) p->dev_prove = atoi(e) == 0;
    p->d_comp0 = (uint8_t *)dalloc(max_batch * in_pp * 48);
    p->d_compH = (uint8_t *)dalloc(48);
    if (p->dev_prove) {
        const size_t psz = cdp_proof_size(ell), nrnd = cdp_prove_random_scalars(ell);
        p->d_perm = (uint32_t *)dalloc(max_batch * ell * 4); p->h_perm = (uint32_t *)halloc(max_batch * ell * 4);
        p->d_wit = (uint8_t *)dalloc(max_batch * 5 * 32); p->h_wit = (uint8_t *)halloc(max_batch * 5 * 32);
        p->d_rnd = (uint8_t *)dalloc(max_batch * nrnd * 32); p->h_rnd = (uint8_t *)halloc(max_batch * nrnd * 32);
        p->d_keys = (uint8_t *)dalloc(max_batch * 40); p->h_keys = (uint8_t *)halloc(max_batch * 40);
        p->d_work = (uint8_t *)dalloc(max_batch * cdp_prove_work_scalars(ell) * 32);
        p->d_side = (uint8_t *)dalloc(max_batch * 96);
        p->d_proofs = (uint8_t *)dalloc(max_batch * psz); p->h_proofs = (uint8_t *)halloc(max_batch * psz);
    }
    p->d_comp = (uint8_t *)dalloc(max_batch * max_out * 48);
    p->h_comp = (uint8_t *)halloc(max_batch * max_out * 48);
    // gather tables.  From the CRS block: the blinder slots of T / U (H) and the constant entries of the X block
    std::vector<uint32_t> gsrc, gdst, isrc, idst, cidx;
    const uint32_t INF_SRC = (uint32_t)(total_pts);  // one all-zero point appended after the array (the point at infinity)
    for (size_t pr = 0; pr < max_batch; pr++) {
        size_t bp = p->crs_n + pr * p->PW;
        // vec_T_with_blinders = T | inf inf H inf ; vec_U_with_blinders = U | inf inf inf H   (curdleproofs.rs:142-155)
        const uint32_t tb[4] = {INF_SRC, INF_SRC, (uint32_t)cH, INF_SRC}, ub[4] = {INF_SRC, INF_SRC, INF_SRC, (uint32_t)cH};
        for (int i = 0; i < 4; i++) {
            gsrc.push_back(tb[i]); gdst.push_back((uint32_t)(bp + p->o_T + ell + i));
            gsrc.push_back(ub[i]); gdst.push_back((uint32_t)(bp + p->o_U + ell + i));
        }
        {   // constant entries of the X block
            const std::pair<int, size_t> cx[] = {{X_GSUM, cGsum}, {X_H1, cH}, {X_H2, cH}, {X_GT, cGt}, {X_GU, cGu}, {X_GSUM2, cGsum}, {X_HSUM, cHsum}};
            for (auto &e : cx) { gsrc.push_back((uint32_t)e.second); gdst.push_back((uint32_t)(bp + p->o_X + e.first)); }
        }
        if (pr == 0) p->g_count_per_proof = gsrc.size();
        // from the instance staging: [R | S | T | U] of this proof, M in the block behind all proofs
        size_t ib = pr * 4 * ell;
        const size_t dsts[4] = {p->o_R, p->o_S, p->o_T, p->o_U};
        for (int v = 0; v < 4; v++)
            for (size_t i = 0; i < ell; i++) { isrc.push_back((uint32_t)(ib + v * ell + i)); idst.push_back((uint32_t)(bp + dsts[v] + i)); }
        isrc.push_back((uint32_t)(max_batch * 4 * ell + pr)); idst.push_back((uint32_t)(bp + p->o_X + X_M));
        if (pr == 0) p->i_count_per_proof = isrc.size();
    }
    // after stage 1: A, R, S (affine) into the X block; after stage 2: B.  The source indices depend on the batch size: built per call.
    p->d_x1src = (uint32_t *)dalloc(max_batch * 4 * 4); p->d_x1dst = (uint32_t *)dalloc(max_batch * 4 * 4);
    p->d_x2src = (uint32_t *)dalloc(max_batch * 4); p->d_x2dst = (uint32_t *)dalloc(max_batch * 4);
    p->d_gsrc = (uint32_t *)dalloc(gsrc.size() * 4); p->d_gdst = (uint32_t *)dalloc(gdst.size() * 4);
    p->d_isrc = (uint32_t *)dalloc(isrc.size() * 4); p->d_idst = (uint32_t *)dalloc(idst.size() * 4);
    if (!ok) { p->err = "allocation failed"; lane_destroy(p); return CDP_ERR_CUDA; }
    int rc = CDP_OK;
    // CRS block + the trailing all-zero point
    std::vector<uint8_t> zero(96, 0);
    rc |= cdp_h2d(ctx, p->d_pts, crs_points, (ell + 9) * 96);  // crs_points here = the CRS followed by sum(G), sum(Hvec) (cdp_prover_create_lanes)
    rc |= cdp_h2d(ctx, p->d_pts + total_pts * 96, zero.data(), 96);
    rc |= cdp_h2d(ctx, p->d_gsrc, gsrc.data(), gsrc.size() * 4);
    rc |= cdp_h2d(ctx, p->d_gdst, gdst.data(), gdst.size() * 4);
    rc |= cdp_h2d(ctx, p->d_isrc, isrc.data(), isrc.size() * 4);
    rc |= cdp_h2d(ctx, p->d_idst, idst.data(), idst.size() * 4);
    // compressed H for the blinder slots of the same_msm transcript message
    rc |= cdp_compress_affine_dev(ctx, p->d_pts + cH * 96, nullptr, 1, p->d_compH);
    rc |= cdp_d2h(ctx, p->h_comp, p->d_compH, 48);
    rc |= cdp_sync(ctx);
    if (rc) { p->err = std::string("setup: ") + cdp_last_error(ctx); lane_destroy(p); return CDP_ERR_CUDA; }
    memcpy(p->H_comp, p->h_comp, 48);
    if (upload_tables(p) != CDP_OK) { lane_destroy(p); return CDP_ERR_CUDA; }
    p->ps.resize(max_batch);
    *out = p;
    return CDP_OK;
}

// the prover's `rng` for proof pr (cdp_prove_inputs): a 32-byte ChaCha12 key, the reference's test-vector u64 seed, or a fresh key from the OS
static bool make_rng(const cdp_prove_inputs *in, size_t pr, StdRng &rng) {
    if (in->rng_key) rng = StdRng(StdRng::from_key_t{}, in->rng_key + 32 * pr);
    else if (in->rng_seed) rng = StdRng(in->rng_seed[pr]);
    else {
        uint8_t key[32];
        if (!StdRng::os_key(key)) return false;
        rng = StdRng(StdRng::from_key_t{}, key);
    }
    if (in->rng_skip_words) rng.skip_words(in->rng_skip_words[pr]);
    return true;
}

namespace {
uint32_t out_map_entry(const MsmStage &st, size_t q) {
    size_t base = 0;
    for (int s = 0; s < st.where[q].first; s++) base += st.subs[s].K;
    return (uint32_t)((base << 16) | (st.subs[st.where[q].first].K << 8) | (size_t)st.where[q].second);
}
// every launch of one MSM stage (its scalars are already in d_scal), then normalise + compress into d_comp; nothing returns to the host
int enqueue_msm_stage(Lane *p, MsmStage &st, size_t B) {
    size_t out_off = 0;
    for (auto &sl : st.subs) {
        if (sl.fixed) PTRY(launch_fixed_sub(p, sl, B, out_off));
        else PTRY(cdp_msm_batch_dev(p->ctx, p->d_pts, p->d_scal, sl.d_segs, B * sl.K, sl.max_n, B * sl.pairs_per_proof, p->d_jac + out_off * 144));
        out_off += B * sl.K;
    }
    PTRY(cdp_normalize_dev(p->ctx, p->d_jac, out_off, st.aff_region != (size_t)-1 ? p->d_pts + st.aff_region * 96 : nullptr, p->d_comp));
    return CDP_OK;
}
// one step of the device-side transcript + scalar algebra: reads the outputs of `consumed`, writes the scalars of the next stage
int enqueue_prove_stage(Lane *p, cdp_prove_dev &P, int stage, unsigned round, const MsmStage *consumed, size_t emitted_spp) {
    memset(P.out_map, 0, sizeof P.out_map);
    if (consumed)
        for (size_t q = 0; q < consumed->outputs(); q++) P.out_map[q] = out_map_entry(*consumed, q);
    P.scalars_per_proof = (uint32_t)emitted_spp;
    PTRY(cdp_prove_stage_dev(p->ctx, &P, stage, round));
    return CDP_OK;
}
}  // namespace

// `CurdleproofsProof::new` for a sub-batch with the whole protocol on the device: after the instance, the witnesses and the prover's
// randomness are in HBM, the lane is ONE stream of kernels (MSM launches alternating with cdp_prove_stage_dev steps) and one
// synchronisation; the host's only arithmetic is the ChaCha12 stream of the randomness.
static int lane_prove_device(Lane *p, size_t B, const cdp_prove_inputs *in, uint8_t *proofs_out, double t_start, double t_copy) {
    const size_t ell = p->ell, n = p->n, m = p->m;
    const int T = p->threads;
    const size_t proof_size = cdp_proof_size(ell), nrnd = cdp_prove_random_scalars(ell);
    double t0 = now_ms(), t_host = 0, t_wait = 0;
    // witnesses and randomness (all draws in the reference's order; they do not depend on the transcript).  The generator runs on the device
    // (cdp_prove_random_dev: the host hands over each proof's ChaCha12 key and stream position); CDP_PROVE_HOST_RNG=1 draws here instead
    static const bool host_rng = [] { const char *e = getenv("CDP_PROVE_HOST_RNG"); return e && atoi(e) != 0; }();
    std::vector<int> rng_ok(B, 1);
    parallel_for(T, B, [&](size_t pr) {
        memcpy(p->h_perm + pr * ell, in->permutation + pr * ell, ell * 4);
        memcpy(p->h_wit + pr * 160, in->k + 32 * pr, 32);
        memcpy(p->h_wit + pr * 160 + 32, in->vec_m_blinders + 128 * pr, 128);
        StdRng rng(0);
        if (!host_rng) {  // key and position only
            if (in->rng_key) memcpy(p->h_keys + 32 * pr, in->rng_key + 32 * pr, 32);
            else if (in->rng_seed) { rng = StdRng(in->rng_seed[pr]); memcpy(p->h_keys + 32 * pr, rng.key(), 32); }
            else if (!StdRng::os_key(p->h_keys + 32 * pr)) { rng_ok[pr] = 0; return; }
            const uint64_t skip = in->rng_skip_words ? in->rng_skip_words[pr] : 0;
            memcpy(p->h_keys + 32 * B + 8 * pr, &skip, 8);
            return;
        }
        if (!make_rng(in, pr, rng)) { rng_ok[pr] = 0; return; }
        uint64_t *w = reinterpret_cast<uint64_t *>(p->h_rnd + pr * nrnd * 32);
        auto draw = [&](size_t slot) { Fr x = rng.fr_rand(); memcpy(w + 4 * slot, x.v, 32); };
        for (size_t i = 0; i < 6; i++) draw(i);                       // r_a[2] (curdleproofs.rs:86), r_c[4] (grand_product_argument.rs:75)
        for (size_t i = 0; i < n; i++) draw(6 + i);                   // inner_product_argument.rs:46
        for (size_t i = 0; i + 2 < n; i++) draw(6 + n + i);           // :47
        memset(w + 4 * (6 + 2 * n - 2), 0, 64);                       // the two entries solved for on the device (:53-77)
        for (size_t i = 0; i < 5; i++) draw(6 + 2 * n + i);           // r_t r_u (curdleproofs.rs:110-111), r_a r_b r_k (same_scalar_argument.rs:56-58)
        for (size_t i = 0; i < n; i++) draw(11 + 2 * n + i);          // same_multiscalar_argument.rs:78
    });
    for (size_t pr = 0; pr < B; pr++)
        if (!rng_ok[pr]) return perr(p, CDP_ERR_INVALID_ARG, "cdp_prove_batch: no entropy source for the prover's randomness");
    t_host += now_ms() - t0;
    t0 = now_ms();
    PTRY(cdp_h2d(p->ctx, p->d_perm, p->h_perm, B * ell * 4));
    PTRY(cdp_h2d(p->ctx, p->d_wit, p->h_wit, B * 160));
    if (host_rng) {
        PTRY(cdp_h2d(p->ctx, p->d_rnd, p->h_rnd, B * nrnd * 32));
        p->h2d_bytes += B * (ell * 4 + 160 + nrnd * 32);
    } else {
        PTRY(cdp_h2d(p->ctx, p->d_keys, p->h_keys, B * 40));
        PTRY(cdp_prove_random_dev(p->ctx, p->d_keys, reinterpret_cast<const uint64_t *>(p->d_keys + 32 * B), B, ell, p->d_rnd));
        p->h2d_bytes += B * (ell * 4 + 160 + 40);
    }

    cdp_prove_dev P;
    memset(&P, 0, sizeof P);
    P.ell = (uint32_t)ell; P.m = (uint32_t)m; P.batch = (uint32_t)B; P.proof_bytes = (uint32_t)proof_size;
    P.d_state = p->d_tstate; P.d_vec_a = p->d_veca; P.d_perm = p->d_perm; P.d_witness = p->d_wit; P.d_random = p->d_rnd; P.d_work = p->d_work;
    P.d_comp0_vecs = p->d_comp0; P.d_comp0_M = p->d_comp0 + B * 4 * ell * 48; P.d_comp_H = p->d_compH; P.d_comp = p->d_comp;
    P.d_side = p->d_side; P.d_proofs = p->d_proofs; P.d_scalars = p->d_scal; P.d_fold_scalars = p->d_fscal;
    P.switch_round = (uint32_t)p->k0;

    if (int rc = enqueue_prove_stage(p, P, CDP_PS_S1, 0, nullptr, p->st1.scalars_per_proof)) return rc;
    if (int rc = enqueue_msm_stage(p, p->st1, B)) return rc;
    if (int rc = enqueue_prove_stage(p, P, CDP_PS_SAMEPERM, 0, &p->st1, p->st2.scalars_per_proof)) return rc;
    PTRY(cdp_gather_dev(p->ctx, p->d_pts, p->d_pts, p->d_x1src, p->d_x1dst, B * 4));  // A, R, S (affine, from stage 1) -> X block
    if (int rc = enqueue_msm_stage(p, p->st2, B)) return rc;
    if (int rc = enqueue_prove_stage(p, P, CDP_PS_GPROD1, 0, &p->st2, p->st3.scalars_per_proof)) return rc;
    if (int rc = enqueue_msm_stage(p, p->st3, B)) return rc;
    if (int rc = enqueue_prove_stage(p, P, CDP_PS_GPROD2, 0, &p->st3, p->st4.scalars_per_proof)) return rc;
    PTRY(cdp_gather_dev(p->ctx, p->d_pts, p->d_pts, p->d_x2src, p->d_x2dst, B));      // B (affine, from stage 2) -> X block
    if (int rc = enqueue_msm_stage(p, p->st4, B)) return rc;
    if (int rc = enqueue_prove_stage(p, P, CDP_PS_IPA0, 0, &p->st4, p->st_ipa[0].scalars_per_proof)) return rc;
    const size_t k0 = p->k0;
    for (size_t k = 0; k < m; k++) {
        if (k == k0)  // G^(k0), G'^(k0) into regM
            if (int rc = enqueue_msm_stage(p, p->st_ipa_mat, B)) return rc;
        if (int rc = enqueue_msm_stage(p, p->st_ipa[k], B)) return rc;
        const size_t next_spp = k + 1 >= m ? p->st_sm[0].scalars_per_proof : k + 1 >= k0 ? p->st_ipa[k0].scalars_per_proof : p->st_ipa[k + 1].scalars_per_proof;
        if (int rc = enqueue_prove_stage(p, P, CDP_PS_IPA_ROUND, (unsigned)k, &p->st_ipa[k], next_spp)) return rc;
        if (k >= k0 && (n >> (k + 1)) > 1)
            PTRY(cdp_smul_jobs_dev(p->ctx, p->d_pts, p->d_fscal, p->f_ipa[k].d_jobs, B * p->f_ipa[k].J, p->f_ipa[k].epj));  // G, G' folds (:177-178)
    }
    for (size_t k = 0; k < m; k++) {
        if (k == k0)  // G_with_blinders^(k0) into regS
            if (int rc = enqueue_msm_stage(p, p->st_sm_mat, B)) return rc;
        if (int rc = enqueue_msm_stage(p, p->st_sm[k], B)) return rc;
        const size_t next_spp = p->st_sm[k + 1 < m ? k + 1 : k].scalars_per_proof;
        if (int rc = enqueue_prove_stage(p, P, CDP_PS_SM_ROUND, (unsigned)k, &p->st_sm[k], next_spp)) return rc;
        if ((n >> (k + 1)) > 1) PTRY(cdp_smul_jobs_dev(p->ctx, p->d_pts, p->d_fscal, p->f_sm[k].d_jobs, B * p->f_sm[k].J, p->f_sm[k].epj));  // T, U (, G) folds (:128-130)
    }
    PTRY(cdp_d2h(p->ctx, p->h_proofs, p->d_proofs, B * proof_size));
    p->d2h_bytes += B * proof_size;
    t_copy += now_ms() - t0;
    t0 = now_ms();
    PTRY(cdp_sync(p->ctx));
    t_wait += now_ms() - t0;
    t0 = now_ms();
    memcpy(proofs_out, p->h_proofs, B * proof_size);
    t_host += now_ms() - t0;
    p->timing[0] = now_ms() - t_start;
    p->timing[1] = t_host;
    p->timing[2] = t_wait;
    p->timing[3] = t_copy;
    return CDP_OK;
}

static int lane_prove(Lane *p, size_t B, const cdp_prove_inputs *in, uint8_t *proofs_out) {
    if (!p) return CDP_ERR_INVALID_ARG;
    if (!in || !proofs_out || B > p->max_batch) return perr(p, CDP_ERR_INVALID_ARG, "cdp_prove_batch: bad argument");
    if (B == 0) return CDP_OK;
    const bool resident = in->vec_R == nullptr;  // instance vectors of the previous call are still staged in HBM
    if (resident && p->staged_batch < B) return perr(p, CDP_ERR_INVALID_ARG, "cdp_prove_batch: no resident instance batch of that size");
    p->h2d_bytes = p->d2h_bytes = 0;
    const size_t ell = p->ell, n = p->n, m = p->m;
    const int T = p->threads;
    double t_start = now_ms(), t_host = 0, t_wait = 0, t_copy = 0, t0;
    const size_t proof_size = cdp_proof_size(ell);
    const size_t Moff = p->max_batch * 4 * ell;  // index of the first affine M in d_in

    // ---- stage 0: instance to the device, working vectors assembled, transcript openings compressed
    t0 = now_ms();
    if (!resident) {
        const uint8_t *vecs[4] = {in->vec_R, in->vec_S, in->vec_T, in->vec_U};
        bool pinned = true;
        for (const uint8_t *v : vecs) pinned = pinned && cdp_host_is_pinned(v);
        memcpy(p->h_in + B * 4 * ell * 96, in->M, B * 144);
        if (pinned) {  // page-locked caller buffers: four strided DMAs straight into the proof-major layout, no staging pass
            for (int v = 0; v < 4; v++) PTRY(cdp_h2d_2d(p->ctx, p->d_in + v * ell * 96, 4 * ell * 96, vecs[v], ell * 96, ell * 96, B));
        } else {
            parallel_for(T, B, [&](size_t pr) {
                uint8_t *dst = p->h_in + pr * 4 * ell * 96;
                for (int v = 0; v < 4; v++) memcpy(dst + v * ell * 96, vecs[v] + pr * ell * 96, ell * 96);
            });
            PTRY(cdp_h2d(p->ctx, p->d_in, p->h_in, B * 4 * ell * 96));
        }
        PTRY(cdp_h2d(p->ctx, p->d_Mjac, p->h_in + B * 4 * ell * 96, B * 144));
        p->h2d_bytes += B * (4 * ell * 96 + 144);
        PTRY(cdp_normalize_dev(p->ctx, p->d_Mjac, B, p->d_in + Moff * 96, nullptr));  // M.into_affine()
        p->staged_batch = B;
    }
    if (p->xtab_B != B) {  // where stage 1 / stage 2 leave A, R, S / B (affine) depends on the batch size
        std::vector<uint32_t> x1s, x1d, x2s, x2d;
        for (size_t pr = 0; pr < B; pr++) {
            size_t bp = p->crs_n + pr * p->PW;
            const std::pair<int, int> e1[] = {{X_A, S1_A}, {X_R, S1_R}, {X_S, S1_S}, {X_A2, S1_A}};
            for (auto &e : e1) {
                x1s.push_back((uint32_t)(p->reg1 + stage_out_index(p->st1, B, pr, e.second)));
                x1d.push_back((uint32_t)(bp + p->o_X + e.first));
            }
            x2s.push_back((uint32_t)(p->reg2 + stage_out_index(p->st2, B, pr, S2_B)));
            x2d.push_back((uint32_t)(bp + p->o_X + X_B));
        }
        PTRY(cdp_h2d(p->ctx, p->d_x1src, x1s.data(), x1s.size() * 4)); PTRY(cdp_h2d(p->ctx, p->d_x1dst, x1d.data(), x1d.size() * 4));
        PTRY(cdp_h2d(p->ctx, p->d_x2src, x2s.data(), x2s.size() * 4)); PTRY(cdp_h2d(p->ctx, p->d_x2dst, x2d.data(), x2d.size() * 4));
        PTRY(cdp_sync(p->ctx));  // pageable sources
        p->xtab_B = B;
    }
    PTRY(cdp_gather_dev(p->ctx, p->d_pts, p->d_pts, p->d_gsrc, p->d_gdst, B * p->g_count_per_proof));
    PTRY(cdp_gather_dev(p->ctx, p->d_pts, p->d_in, p->d_isrc, p->d_idst, B * p->i_count_per_proof));
    PTRY(cdp_compress_affine_dev(p->ctx, p->d_in, nullptr, B * 4 * ell, p->d_comp0));
    PTRY(cdp_compress_affine_dev(p->ctx, p->d_in + Moff * 96, nullptr, B, p->d_comp0 + B * 4 * ell * 48));
    // transcript opening (R, S, T, U, M -> vec_a) hashed on the device
    PTRY(cdp_transcript_open_dev(p->ctx, p->d_comp0, p->d_comp0 + B * 4 * ell * 48, ell, B, p->d_veca, p->d_tstate));
    if (p->dev_prove) return lane_prove_device(p, B, in, proofs_out, t_start, t_copy + now_ms() - t0);
    // older path: the host continues the transcripts from the returned STROBE states
    PTRY(cdp_d2h(p->ctx, p->h_comp, p->d_comp0, (B * 4 * ell + B) * 48));
    PTRY(cdp_d2h(p->ctx, p->h_veca, p->d_veca, B * ell * 32));
    PTRY(cdp_d2h(p->ctx, p->h_tstate, p->d_tstate, B * CDP_TRANSCRIPT_STATE_BYTES));
    p->d2h_bytes += (B * 4 * ell + B) * 48 + B * ell * 32 + B * CDP_TRANSCRIPT_STATE_BYTES;
    t_copy += now_ms() - t0;
    t0 = now_ms();
    PTRY(cdp_sync(p->ctx));
    t_wait += now_ms() - t0;

    // ---- transcript opening, randomness, stage-1 scalars (curdleproofs.rs:78-93,110-116; same_scalar_argument.rs:56-61;
    //      same_multiscalar_argument.rs:78-82)
    t0 = now_ms();
    std::vector<uint8_t> tu_comp(B * 2 * n * 48);  // vec_T_with_blinders | vec_U_with_blinders encodings, kept for same_msm_step1
    std::atomic<bool> rng_fail{false};
    parallel_for(T, B, [&](size_t pr) {
        ProofState &s = p->ps[pr];
        const uint8_t *cmp = p->h_comp + pr * 4 * ell * 48;
        memcpy(s.M_comp, p->h_comp + (B * 4 * ell + pr) * 48, 48);
        s.tr.reset(new Transcript(reinterpret_cast<const uint64_t *>(p->h_tstate + pr * CDP_TRANSCRIPT_STATE_BYTES)));
        uint8_t *tu = tu_comp.data() + pr * 2 * n * 48;
        uint8_t inf[48] = {0xC0};
        memcpy(tu, cmp + 2 * ell * 48, ell * 48);
        memcpy(tu + ell * 48, inf, 48); memcpy(tu + (ell + 1) * 48, inf, 48); memcpy(tu + (ell + 2) * 48, p->H_comp, 48); memcpy(tu + (ell + 3) * 48, inf, 48);
        uint8_t *uu = tu + n * 48;
        memcpy(uu, cmp + 3 * ell * 48, ell * 48);
        memcpy(uu + ell * 48, inf, 48); memcpy(uu + (ell + 1) * 48, inf, 48); memcpy(uu + (ell + 2) * 48, inf, 48); memcpy(uu + (ell + 3) * 48, p->H_comp, 48);
        s.vec_a.resize(ell);
        for (size_t i = 0; i < ell; i++) Fr::from_bytes(p->h_veca + (pr * ell + i) * 32, s.vec_a[i]);
        // witnesses
        s.perm.assign(in->permutation + pr * ell, in->permutation + (pr + 1) * ell);
        Fr::from_bytes(in->k + 32 * pr, s.k);
        for (int i = 0; i < 4; i++) Fr::from_bytes(in->vec_m_blinders + 32 * (4 * pr + i), s.m_bl[i]);
        // all prover randomness, in the reference's draw order
        StdRng rng(0);
        if (!make_rng(in, pr, rng)) { rng_fail = true; return; }
        s.a_bl[0] = rng.fr_rand(); s.a_bl[1] = rng.fr_rand();                 // curdleproofs.rs:86
        for (int i = 0; i < 4; i++) s.c_bl[i] = rng.fr_rand();                  // grand_product_argument.rs:75
        s.r_c.resize(n); s.r_d.resize(n);
        for (size_t i = 0; i < n; i++) s.r_c[i] = rng.fr_rand();                // inner_product_argument.rs:46
        for (size_t i = 0; i + 2 < n; i++) s.r_d[i] = rng.fr_rand();            // :47
        s.r_t = rng.fr_rand(); s.r_u = rng.fr_rand();                           // curdleproofs.rs:110-111
        s.r_a = rng.fr_rand(); s.r_b = rng.fr_rand(); s.r_k = rng.fr_rand();    // same_scalar_argument.rs:56-58
        s.r_sm.resize(n);
        for (size_t i = 0; i < n; i++) s.r_sm[i] = rng.fr_rand();               // same_multiscalar_argument.rs:78
        s.a_perm.resize(ell);
        for (size_t i = 0; i < ell; i++) s.a_perm[i] = s.vec_a[s.perm[i]];
        // stage-1 scalar block (layout fixed in lane_create)
        uint8_t *sc = p->h_scal + pr * p->st1.scalars_per_proof * 32;
        size_t o = 0;
        for (size_t i = 0; i < ell; i++) put_fr(sc + 32 * (o++), s.a_perm[i]);                 // A
        put_fr(sc + 32 * (o++), s.a_bl[0]); put_fr(sc + 32 * (o++), s.a_bl[1]);
        memset(sc + 32 * o, 0, 64); o += 2;
        for (size_t i = 0; i < ell; i++) put_fr(sc + 32 * (o++), s.vec_a[i]);                  // R, S
        for (size_t i = 0; i < n; i++) put_fr(sc + 32 * (o++), s.r_sm[i]);                     // B_a, B_t, B_u
        put_fr(sc + 32 * (o++), s.r_t); put_fr(sc + 32 * (o++), s.r_u);                        // cm_T.T_1, cm_U.T_1
        put_fr(sc + 32 * (o++), s.r_a); put_fr(sc + 32 * (o++), s.r_b);                        // cm_A.T_1, cm_B.T_1
        s.ipa_rounds.resize(m * 4 * 48);
        s.sm_rounds.resize(m * 6 * 48);
    });
    t_host += now_ms() - t0;
    if (rng_fail) return perr(p, CDP_ERR_INVALID_ARG, "cdp_prove_batch: no entropy source for the prover's randomness");
    if (int rc = run_msm_stage(p, p->st1, B, t_wait, t_copy)) return rc;

    // ---- same_perm (same_permutation_argument.rs:60-82) -> stage 2: B
    t0 = now_ms();
    parallel_for(T, B, [&](size_t pr) {
        ProofState &s = p->ps[pr];
        static const int map1[S1_COUNT] = {O_A, O_R, O_S, O_BA, O_BT, O_BU, O_T1, O_U1, O_A1, O_B1};
        for (int q = 0; q < S1_COUNT; q++) memcpy(s.pts1[map1[q]], stage_out(p, p->st1, B, pr, q), 48);
        s.tr->append_point("same_perm_step1", s.pts1[O_A]);
        s.tr->append_point("same_perm_step1", s.M_comp);
        s.tr->append_fr_vec("same_perm_step1", s.vec_a.data(), ell);
        s.alpha_sp = s.tr->challenge("same_perm_alpha");
        s.beta_sp = s.tr->challenge("same_perm_beta");
        s.factors.resize(ell);
        s.gprod_result = Fr::one();
        for (size_t i = 0; i < ell; i++) {
            s.factors[i] = s.a_perm[i] + Fr::from_u64(s.perm[i]) * s.alpha_sp + s.beta_sp;
            s.gprod_result *= s.factors[i];
        }
        const Fr r_a_prime[4] = {s.a_bl[0], s.a_bl[1], Fr::zero(), Fr::zero()};
        for (int i = 0; i < 4; i++) s.b_bl[i] = r_a_prime[i] + s.alpha_sp * s.m_bl[i];
        // stage-2 scalars: B = 1 A + alpha M + beta sum(G) | T_2 = k R + r_t H | A_2 = r_k R + r_a H | U_2 | B_2 | A' = 1 A + r_t G_t + r_u G_u
        uint8_t *sc = p->h_scal + pr * p->st2.scalars_per_proof * 32;
        const Fr one = Fr::one();
        const Fr v2[14] = {one, s.alpha_sp, s.beta_sp, s.k, s.r_t, s.r_k, s.r_a, s.k, s.r_u, s.r_k, s.r_b, one, s.r_t, s.r_u};
        for (int i = 0; i < 14; i++) put_fr(sc + 32 * i, v2[i]);
    });
    t_host += now_ms() - t0;
    PTRY(cdp_gather_dev(p->ctx, p->d_pts, p->d_pts, p->d_x1src, p->d_x1dst, B * 4));  // A, R, S (affine, from stage 1) -> X block
    if (int rc = run_msm_stage(p, p->st2, B, t_wait, t_copy)) return rc;

    // ---- gprod step 1-2 (grand_product_argument.rs:63-83) -> stage 3: C
    t0 = now_ms();
    parallel_for(T, B, [&](size_t pr) {
        ProofState &s = p->ps[pr];
        memcpy(s.B, stage_out(p, p->st2, B, pr, S2_B), 48);
        memcpy(s.pts1[O_T2], stage_out(p, p->st2, B, pr, S2_T2), 48);
        memcpy(s.pts1[O_A2], stage_out(p, p->st2, B, pr, S2_A2), 48);
        memcpy(s.pts1[O_U2], stage_out(p, p->st2, B, pr, S2_U2), 48);
        memcpy(s.pts1[O_B2], stage_out(p, p->st2, B, pr, S2_B2), 48);
        memcpy(s.pts1[O_AP], stage_out(p, p->st2, B, pr, S2_AP), 48);
        s.tr->append_point("gprod_step1", s.B);
        s.tr->append_fr("gprod_step1", s.gprod_result);
        s.alpha_g = s.tr->challenge("gprod_alpha");
        s.c.assign(n, Fr::zero());
        s.c[0] = Fr::one();
        for (size_t i = 0; i + 1 < ell; i++) s.c[i + 1] = s.c[i] * s.factors[i];
        for (int i = 0; i < 4; i++) s.c[ell + i] = s.c_bl[i];
        for (int i = 0; i < 4; i++) s.rb_alpha[i] = s.b_bl[i] + s.alpha_g;
        s.r_p = inner_product(s.rb_alpha, s.c_bl, 4);
        uint8_t *sc = p->h_scal + pr * p->st3.scalars_per_proof * 32;
        for (size_t i = 0; i < n; i++) put_fr(sc + 32 * i, s.c[i]);
    });
    t_host += now_ms() - t0;
    if (int rc = run_msm_stage(p, p->st3, B, t_wait, t_copy)) return rc;

    // ---- gprod step 3-4 (grand_product_argument.rs:85-147) + IPA step 1 (inner_product_argument.rs:124-127) -> stage 4: D, B_c, B_d
    t0 = now_ms();
    parallel_chunks(T, B, [&](size_t lo, size_t hi) {
        // the field inversions of all proofs of the chunk share one exponentiation: {beta, c[n-2]} first, then the blinder denominator
        std::vector<Fr> inv1(2 * (hi - lo)), inv2(hi - lo);
        for (size_t pr = lo; pr < hi; pr++) {
            ProofState &s = p->ps[pr];
            memcpy(s.C, stage_out(p, p->st3, B, pr, 0), 48);
            s.tr->append_point("gprod_step2", s.C);
            s.tr->append_fr("gprod_step2", s.r_p);
            s.beta_g = s.tr->challenge("gprod_beta");
            inv1[2 * (pr - lo)] = s.beta_g;
            inv1[2 * (pr - lo) + 1] = s.c[n - 2];
        }
        batch_inverse(inv1);
        std::vector<Fr> omega(hi - lo), delta(hi - lo);
        for (size_t pr = lo; pr < hi; pr++) {
            ProofState &s = p->ps[pr];
            const Fr beta = s.beta_g, beta_inv = inv1[2 * (pr - lo)], inv_c = inv1[2 * (pr - lo) + 1];
            s.u.resize(n);
            Fr pw = beta_inv;
            for (size_t i = 0; i < ell; i++) { s.u[i] = pw; pw *= beta_inv; }     // beta^-(i+1)
            for (size_t i = 0; i < 4; i++) s.u[ell + i] = pw;                       // beta^-(ell+1)
            s.d.assign(n, Fr::zero());
            Fr pb = beta, p1 = Fr::one();
            for (size_t i = 0; i < ell; i++) {                                      // d_i = b_i beta^(i+1) - beta^i
                s.d[i] = s.factors[i] * pb - p1;
                p1 = pb;
                pb *= beta;
            }
            const Fr beta_l = p1, beta_l1 = pb;                                     // beta^ell, beta^(ell+1)
            for (int i = 0; i < 4; i++) s.d[ell + i] = beta_l1 * s.rb_alpha[i];
            s.z = s.r_p * beta_l1 + s.gprod_result * beta_l - Fr::one();            // inner_prod
            // generate_ipa_blinders: the two left-out blinders solve <r_c,d> + <r_d,c> = 0 and <r_c,r_d> = 0   (:53-77)
            const std::vector<Fr> &c = s.c, &d = s.d;
            std::vector<Fr> &r = s.r_c, &z = s.r_d;
            omega[pr - lo] = inner_product(r.data(), d.data(), n) + inner_product(z.data(), c.data(), n - 2);
            delta[pr - lo] = inner_product(r.data(), z.data(), n - 2);
            inv2[pr - lo] = r[n - 2].neg() * inv_c * c[n - 1] + r[n - 1];
        }
        batch_inverse(inv2);
        for (size_t pr = lo; pr < hi; pr++) {
            ProofState &s = p->ps[pr];
            const Fr beta_inv = inv1[2 * (pr - lo)], inv_c = inv1[2 * (pr - lo) + 1];
            const std::vector<Fr> &c = s.c;
            std::vector<Fr> &r = s.r_c, &z = s.r_d;
            Fr last_z = (r[n - 2] * inv_c * omega[pr - lo] - delta[pr - lo]) * inv2[pr - lo];
            Fr pen_z = inv_c.neg() * (last_z * c[n - 1] + omega[pr - lo]);
            z[n - 2] = pen_z;
            z[n - 1] = last_z;
            uint8_t *sc = p->h_scal + pr * p->st4.scalars_per_proof * 32;
            // D = 1 B - beta^-1 sum(G) + alpha_g sum(Hvec)   (the reference's verifier uses the same identity, grand_product_argument.rs:223)
            put_fr(sc, Fr::one()); put_fr(sc + 32, beta_inv.neg()); put_fr(sc + 64, s.alpha_g);
            for (size_t i = 0; i < n; i++) put_fr(sc + 32 * (3 + i), s.r_c[i]);                    // B_c = msm(G|Hvec, r_c)
            for (size_t i = 0; i < n; i++) put_fr(sc + 32 * (3 + n + i), s.r_d[i] * s.u[i]);       // B_d = msm(G', r_d)
        }
    });
    t_host += now_ms() - t0;
    t0 = now_ms();
    PTRY(cdp_gather_dev(p->ctx, p->d_pts, p->d_pts, p->d_x2src, p->d_x2dst, B));  // B (affine, from stage 2) -> X block
    t_copy += now_ms() - t0;
    if (int rc = run_msm_stage(p, p->st4, B, t_wait, t_copy)) return rc;

    // ---- IPA step 1 transcript + rewrite of c, d (inner_product_argument.rs:129-140), then the rounds (:150-186)
    t0 = now_ms();
    parallel_for(T, B, [&](size_t pr) {
        ProofState &s = p->ps[pr];
        memcpy(s.D, stage_out(p, p->st4, B, pr, 0), 48);
        memcpy(s.B_c, stage_out(p, p->st4, B, pr, 1), 48);
        memcpy(s.B_d, stage_out(p, p->st4, B, pr, 2), 48);
        s.tr->append_point("ipa_step1", s.C);
        s.tr->append_point("ipa_step1", s.D);
        s.tr->append_fr("ipa_step1", s.z);
        s.tr->append_point("ipa_step1", s.B_c);
        s.tr->append_point("ipa_step1", s.B_d);
        s.alpha_i = s.tr->challenge("ipa_alpha");
        s.beta_i = s.tr->challenge("ipa_beta");
        for (size_t i = 0; i < n; i++) {
            s.c[i] = s.r_c[i] + s.alpha_i * s.c[i];
            s.d[i] = s.r_d[i] + s.alpha_i * s.d[i];
        }
        // fold weights of the original bases: G^(k)_i = sum_{j = i mod n_k} wG[j] G_j,  G'^(k)_i = sum wGp[j] G_j  (G' = u o G, :92-102 of gprod)
        if (p->dev_expand) {
            s.WcP.assign(1, fr_canonical_one());
            s.WdP.assign(1, Fr::one());
            uint8_t *uc = p->h_ucan + pr * n * 32;
            for (size_t j = 0; j < n; j++) s.u[j].to_bytes(uc + 32 * j);  // u canonical, resident on the device for all rounds
        } else {
            s.wG.assign(n, fr_canonical_one());
            s.wGp.resize(n);
            for (size_t j = 0; j < n; j++) s.wGp[j] = fr_to_canonical_value(s.u[j]);
        }
    });
    t_host += now_ms() - t0;
    if (p->dev_expand) {
        PTRY(cdp_h2d(p->ctx, p->d_ucan, p->h_ucan, B * n * 32));
        p->h2d_bytes += B * n * 32;
    }
    for (size_t k = 0; k < m; k++) {
        const size_t h = n >> (k + 1);
        MsmStage &st = p->st_ipa[k];
        t0 = now_ms();
        parallel_for(T, B, [&](size_t pr) {
            ProofState &s = p->ps[pr];
            const Fr *cL = s.c.data(), *cR = s.c.data() + h, *dL = s.d.data(), *dR = s.d.data() + h;
            if (p->dev_expand) {  // compact block: Wc[Q] | c[2h] | Wd[Q] | d[2h] | ipL | ipR  (cdp_round_expand_dev, mode 0)
                const size_t Q = n / (2 * h);
                uint8_t *w = p->h_cmp + pr * (2 * Q + 4 * h + 2) * 32;
                memcpy(w, s.WcP.data(), Q * 32); w += Q * 32;
                memcpy(w, s.c.data(), 2 * h * 32); w += 2 * h * 32;
                memcpy(w, s.WdP.data(), Q * 32); w += Q * 32;
                memcpy(w, s.d.data(), 2 * h * 32); w += 2 * h * 32;
                put_fr(w, s.beta_i * inner_product(cL, dR, h));
                put_fr(w + 32, s.beta_i * inner_product(cR, dL, h));
                return;
            }
            uint8_t *sc = p->h_scal + pr * st.scalars_per_proof * 32;
            // msm(G_R, c_L), msm(G_L, c_R), msm(G'_L, d_R), msm(G'_R, d_L) (:158-161) over the original bases
            for (size_t j = 0; j < n; j++) {
                const size_t i = j & (h - 1);
                const bool hi = (j & h) != 0;
                put_canonical(sc + 32 * j, s.wG[j] * (hi ? cL[i] : cR[i]));
                put_canonical(sc + 32 * (n + 2 + j), s.wGp[j] * (hi ? dL[i] : dR[i]));
            }
            put_fr(sc + 32 * n, s.beta_i * inner_product(cL, dR, h));        // H = beta crs_H ; L_C += <c_L,d_R> H
            put_fr(sc + 32 * (n + 1), s.beta_i * inner_product(cR, dL, h));  //                   R_C += <c_R,d_L> H
        });
        t_host += now_ms() - t0;
        {
            const ExpandSpec ex = {0, n, h, 2 * (n / (2 * h)) + 4 * h + 2};
            if (int rc = run_msm_stage(p, st, B, t_wait, t_copy, p->dev_expand ? &ex : nullptr)) return rc;
        }
        t0 = now_ms();
        parallel_chunks(T, B, [&](size_t lo, size_t hi) {
            std::vector<Fr> gam(hi - lo), ginv;
            for (size_t pr = lo; pr < hi; pr++) {
                ProofState &s = p->ps[pr];
                uint8_t *rp = s.ipa_rounds.data() + k * 4 * 48;
                for (int q = 0; q < 4; q++) {  // L_C, L_D, R_C, R_D
                    memcpy(rp + 48 * q, stage_out(p, st, B, pr, q), 48);
                    s.tr->append_point("ipa_loop", rp + 48 * q);
                }
                gam[pr - lo] = s.tr->challenge("ipa_gamma");
            }
            ginv = gam;
            batch_inverse(ginv);  // one inversion for the chunk's gamma^-1 (:171)
            for (size_t pr = lo; pr < hi; pr++) {
                ProofState &s = p->ps[pr];
                const Fr gamma = gam[pr - lo], gamma_inv = ginv[pr - lo];
                for (size_t i = 0; i < h; i++) {
                    s.c[i] += gamma_inv * s.c[h + i];
                    s.d[i] += gamma * s.d[h + i];
                }
                // G_L += gamma G_R, G'_L += gamma^-1 G'_R (:177-178), as weights on the original bases
                if (h > 1 && p->dev_expand) {  // per prefix: the new low bit of the prefix is bit h of the index
                    const size_t Q = n / (2 * h);
                    std::vector<Fr> wc(2 * Q), wd(2 * Q);
                    for (size_t q = 0; q < Q; q++) {
                        wc[2 * q] = s.WcP[q]; wc[2 * q + 1] = s.WcP[q] * gamma;
                        wd[2 * q] = s.WdP[q]; wd[2 * q + 1] = s.WdP[q] * gamma_inv;
                    }
                    s.WcP.swap(wc); s.WdP.swap(wd);
                } else if (h > 1)
                    for (size_t j = 0; j < n; j++)
                        if (j & h) { s.wG[j] *= gamma; s.wGp[j] *= gamma_inv; }
            }
        });
        t_host += now_ms() - t0;
    }

    // ---- same_scalar (same_scalar_argument.rs:64-75) and same_msm step 1 (same_multiscalar_argument.rs:84-91)
    t0 = now_ms();
    parallel_for(T, B, [&](size_t pr) {
        ProofState &s = p->ps[pr];
        s.c_final = s.c[0];
        s.d_final = s.d[0];
        static const int order[10] = {O_R, O_S, O_T1, O_T2, O_U1, O_U2, O_A1, O_A2, O_B1, O_B2};
        for (int q = 0; q < 10; q++) s.tr->append_point("sameexp_points", s.pts1[order[q]]);
        Fr alpha = s.tr->challenge("same_scalar_alpha");
        s.z_k = s.r_k + s.k * alpha;
        s.z_t = s.r_a + s.r_t * alpha;
        s.z_u = s.r_b + s.r_u * alpha;
        s.tr->append_point("same_msm_step1", s.pts1[O_AP]);
        s.tr->append_point("same_msm_step1", s.pts1[O_T2]);
        s.tr->append_point("same_msm_step1", s.pts1[O_U2]);
        const uint8_t *tu = tu_comp.data() + pr * 2 * n * 48;
        s.tr->append_point_vec("same_msm_step1", tu, n);
        s.tr->append_point_vec("same_msm_step1", tu + n * 48, n);
        s.tr->append_point("same_msm_step1", s.pts1[O_BA]);
        s.tr->append_point("same_msm_step1", s.pts1[O_BT]);
        s.tr->append_point("same_msm_step1", s.pts1[O_BU]);
        Fr a_sm = s.tr->challenge("same_msm_alpha");
        s.x.resize(n);
        for (size_t i = 0; i < ell; i++) s.x[i] = s.r_sm[i] + a_sm * s.a_perm[i];
        const Fr tail[4] = {s.a_bl[0], s.a_bl[1], s.r_t, s.r_u};  // vec_a_with_blinders, curdleproofs.rs:157-160
        for (int i = 0; i < 4; i++) s.x[ell + i] = s.r_sm[ell + i] + a_sm * tail[i];
        if (p->dev_expand) s.WsP.assign(1, fr_canonical_one());
        else s.wS.assign(n, fr_canonical_one());
    });
    t_host += now_ms() - t0;
    for (size_t k = 0; k < m; k++) {
        const size_t h = n >> (k + 1);
        MsmStage &st = p->st_sm[k];
        t0 = now_ms();
        parallel_for(T, B, [&](size_t pr) {
            ProofState &s = p->ps[pr];
            uint8_t *sc = p->h_scal + pr * st.scalars_per_proof * 32;
            // msm(G_R, x_L), msm(G_L, x_R) (:107,:110) over the original G_with_blinders; T, U use the folded vectors
            if (p->dev_expand) {  // compact block: Ws[Q] | x[2h]  (cdp_round_expand_dev, mode 1)
                const size_t Q = n / (2 * h);
                uint8_t *w = p->h_cmp + pr * (Q + 2 * h) * 32;
                memcpy(w, s.WsP.data(), Q * 32);
                memcpy(w + Q * 32, s.x.data(), 2 * h * 32);
                return;
            }
            for (size_t j = 0; j < n; j++) put_canonical(sc + 32 * j, s.wS[j] * s.x[(j & h) ? (j & (h - 1)) : h + (j & (h - 1))]);
            for (size_t i = 0; i < 2 * h; i++) put_fr(sc + 32 * (n + i), s.x[i]);
        });
        t_host += now_ms() - t0;
        {
            const ExpandSpec ex = {1, n, h, n / (2 * h) + 2 * h};
            if (int rc = run_msm_stage(p, st, B, t_wait, t_copy, p->dev_expand ? &ex : nullptr)) return rc;
        }
        t0 = now_ms();
        parallel_chunks(T, B, [&](size_t lo, size_t hi) {
            std::vector<Fr> gam(hi - lo), ginv;
            for (size_t pr = lo; pr < hi; pr++) {
                ProofState &s = p->ps[pr];
                uint8_t *rp = s.sm_rounds.data() + k * 6 * 48;
                for (int q = 0; q < 6; q++) {  // L_A, L_T, L_U, R_A, R_T, R_U
                    memcpy(rp + 48 * q, stage_out(p, st, B, pr, q), 48);
                    s.tr->append_point("same_msm_loop", rp + 48 * q);
                }
                gam[pr - lo] = s.tr->challenge("same_msm_gamma");
            }
            ginv = gam;
            batch_inverse(ginv);
            for (size_t pr = lo; pr < hi; pr++) {
                ProofState &s = p->ps[pr];
                const Fr gamma = gam[pr - lo], gamma_inv = ginv[pr - lo];
                for (size_t i = 0; i < h; i++) s.x[i] += gamma_inv * s.x[h + i];
                put_fr(p->h_fscal + pr * 32, gamma);
                if (h > 1 && p->dev_expand) {
                    const size_t Q = n / (2 * h);
                    std::vector<Fr> ws(2 * Q);
                    for (size_t q = 0; q < Q; q++) { ws[2 * q] = s.WsP[q]; ws[2 * q + 1] = s.WsP[q] * gamma; }
                    s.WsP.swap(ws);
                } else if (h > 1)
                    for (size_t j = 0; j < n; j++)
                        if (j & h) s.wS[j] *= gamma;   // G_L += gamma G_R (:130)
            }
        });
        t_host += now_ms() - t0;
        if (h > 1) {
            t0 = now_ms();
            if (int rc = run_fold_stage(p, p->f_sm[k], B)) return rc;
            t_copy += now_ms() - t0;
        }
    }

    // ---- serialise (curdleproofs.rs:300-310 and the per-argument serialisers)
    t0 = now_ms();
    parallel_for(T, B, [&](size_t pr) {
        ProofState &s = p->ps[pr];
        s.x_final = s.x[0];
        uint8_t *w = proofs_out + pr * proof_size;
        auto pt = [&](const uint8_t *c) { memcpy(w, c, 48); w += 48; };
        auto fr = [&](const Fr &x) { x.to_bytes(w); w += 32; };
        pt(s.pts1[O_A]); pt(s.pts1[O_T1]); pt(s.pts1[O_T2]); pt(s.pts1[O_U1]); pt(s.pts1[O_U2]); pt(s.pts1[O_R]); pt(s.pts1[O_S]);
        pt(s.B); pt(s.C); fr(s.r_p);
        pt(s.B_c); pt(s.B_d);
        static const int ipa_order[4] = {0, 2, 1, 3};  // serialised as vec_L_C, vec_R_C, vec_L_D, vec_R_D; rounds hold L_C, L_D, R_C, R_D
        for (int v = 0; v < 4; v++)
            for (size_t k = 0; k < m; k++) pt(s.ipa_rounds.data() + (k * 4 + ipa_order[v]) * 48);
        fr(s.c_final); fr(s.d_final);
        pt(s.pts1[O_A1]); pt(s.pts1[O_A2]); pt(s.pts1[O_B1]); pt(s.pts1[O_B2]);
        fr(s.z_k); fr(s.z_t); fr(s.z_u);
        pt(s.pts1[O_BA]); pt(s.pts1[O_BT]); pt(s.pts1[O_BU]);
        for (int v = 0; v < 6; v++)  // vec_L_A, vec_L_T, vec_L_U, vec_R_A, vec_R_T, vec_R_U
            for (size_t k = 0; k < m; k++) pt(s.sm_rounds.data() + (k * 6 + v) * 48);
        fr(s.x_final);
    });
    t_host += now_ms() - t0;
    p->timing[0] = now_ms() - t_start;
    p->timing[1] = t_host;
    p->timing[2] = t_wait;
    p->timing[3] = t_copy;
    return CDP_OK;
}

// =================================================================================================== public C ABI
// The prover runs `lanes` independent sub-batches concurrently, each on its own CUDA stream (its own cdp_ctx) driven by its
// own host thread: while one lane hashes transcripts on the host, or sits in a latency-bound tail of a small launch, the
// others keep the SMs busy.
struct cdp_prover {
    cdp_ctx *ctx0 = nullptr;  // the caller's context (lane 0): used by the whisk wrappers for their own launches
    SharedCrsTable *shared = nullptr;  // digit table of the CRS points: shared (read-only) by all lanes and by every prover / verifier of the process over the same CRS
    const cdp_fixed_table *table = nullptr;
    std::vector<Lane *> lanes;
    std::vector<cdp_ctx *> owned;
    std::vector<size_t> last_split;
    size_t ell = 0, max_batch = 0;
    bool serial = false;  // run the lanes one after the other (cdp_prover_set_serial): per-kernel device times without overlap
    std::string err = "ok";
    double timing[4] = {0, 0, 0, 0};
    uint64_t traffic[2] = {0, 0};
};

extern "C" const char *cdp_prover_last_error(const cdp_prover *p) { return p ? p->err.c_str() : "null prover"; }
extern "C" void cdp_prover_last_timing(const cdp_prover *p, double out_ms[4]) {
    for (int i = 0; i < 4; i++) out_ms[i] = p ? p->timing[i] : 0.0;
}
extern "C" void cdp_prover_last_traffic(const cdp_prover *p, uint64_t out_bytes[2]) {
    out_bytes[0] = p ? p->traffic[0] : 0;
    out_bytes[1] = p ? p->traffic[1] : 0;
}
extern "C" void cdp_prover_destroy(cdp_prover *p) {
    if (!p) return;
    for (Lane *l : p->lanes) lane_destroy(l);
    crs_table_release(p->shared);
    for (cdp_ctx *c : p->owned) cdp_ctx_destroy(c);
    delete p;
}
extern "C" int cdp_prover_create(cdp_prover **out, cdp_ctx *ctx, size_t ell, const uint8_t *crs_points, size_t max_batch, int host_threads) {
    return cdp_prover_create_lanes(out, ctx, ell, crs_points, max_batch, host_threads, 0);
}
extern "C" int cdp_prover_create_lanes(cdp_prover **out, cdp_ctx *ctx, size_t ell, const uint8_t *crs_points, size_t max_batch, int host_threads,
                                       int lanes) {
    if (!out || !ctx || !crs_points || max_batch == 0) return CDP_ERR_INVALID_ARG;
    *out = nullptr;
    {   // checked before the (multi-GiB) digit table is built: ell >= N_BLINDERS, ell + 4 a power of two (src/inner_product_argument.rs:116)
        const size_t n = ell + NBL;
        if (ell < 4 || (n & (n - 1)) != 0) return CDP_ERR_INVALID_ARG;
        if (n + 1 > 2048) return CDP_ERR_TOO_LARGE;
    }
    int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    if (host_threads <= 0) host_threads = hw;
    // With the whole protocol on the device a lane has no host work to hide, and every kernel of a step is more efficient the larger its launch
    // (additions per inversion and full waves in the tree path, fewer latency-bound tails): measured at ell = 252, 1 / 2 / 4 / 8 lanes take
    // 112 / 113 / 115 / 116 ms at 512 proofs, 178 / 181 / 197 / 204 ms at 1024, 314 / 315 / 332 / 375 ms at 2048 and 581 / 575 / 598 / 639 ms
    // at 4096.  Two lanes keep the copies and the staging of one half overlapped with the kernels of the other.
    if (lanes <= 0) lanes = max_batch >= 256 ? 2 : 1;
    lanes = (int)std::min<size_t>((size_t)lanes, max_batch);
    cdp_prover *p = new cdp_prover();
    p->ctx0 = ctx;
    p->ell = ell;
    p->max_batch = max_batch;
    size_t per_lane = (max_batch + lanes - 1) / lanes;
    // rounded up: lanes alternate between host phases and waiting for the GPU, so a mild oversubscription of the cores costs nothing, while one
    // thread per lane does (measured: 8 / 16 / 32 / 64 threads for 8 lanes on 16 cores: 875 / 767 / 767 / 797 ms per 4096 proofs)
    int threads_per_lane = std::max(1, (host_threads + lanes - 1) / lanes);
    // CRS digit table: G | Hvec | H | G_t | G_u | sum(G) | sum(Hvec) (the last two are the reference's crs.G_sum / crs.H_sum,
    // src/crs.rs:46-47), one per (device, CRS) in the process (crs_table.hpp).  CDP_FIXED_BITS overrides the window width (default 16: 50 MB per base)
    if (int rc = crs_table_acquire(ctx, ell, crs_points, &p->shared)) { delete p; return rc; }
    p->table = p->shared->table;
    const std::vector<uint8_t> &crs_ext = p->shared->crs_ext;
    for (int i = 0; i < lanes; i++) {
        cdp_ctx *c = ctx;
        if (i > 0) {
            if (cdp_ctx_create(&c, cdp_ctx_device(ctx), nullptr) != CDP_OK) { cdp_prover_destroy(p); return CDP_ERR_CUDA; }
            p->owned.push_back(c);
        }
        Lane *l = nullptr;
        int rc = lane_create(&l, c, p->table, ell, crs_ext.data(), per_lane, threads_per_lane);
        if (rc != CDP_OK) { cdp_prover_destroy(p); return rc; }
        p->lanes.push_back(l);
    }
    *out = p;
    return CDP_OK;
}
extern "C" int cdp_prover_lane_count(const cdp_prover *p) { return p ? (int)p->lanes.size() : 0; }
extern "C" void cdp_prover_set_serial(cdp_prover *p, int on) { if (p) p->serial = on != 0; }
extern "C" size_t cdp_prover_table_bytes(const cdp_prover *p) { return p && p->table ? cdp_fixed_table_bytes(p->table) : 0; }
extern "C" cdp_ctx *cdp_prover_lane_ctx(const cdp_prover *p, int lane) {
    return (p && lane >= 0 && lane < (int)p->lanes.size()) ? p->lanes[lane]->ctx : nullptr;
}

extern "C" int cdp_prove_batch(cdp_prover *p, size_t B, const cdp_prove_inputs *in, uint8_t *proofs_out) {
    if (!p) return CDP_ERR_INVALID_ARG;
    if (!in || !proofs_out || B == 0 || B > p->max_batch) { p->err = "cdp_prove_batch: bad argument"; return CDP_ERR_INVALID_ARG; }
    const size_t L = p->lanes.size(), ell = p->ell, psz = cdp_proof_size(ell);
    if (!in->permutation || !in->k || !in->vec_m_blinders) { p->err = "cdp_prove_batch: missing witness"; return CDP_ERR_INVALID_ARG; }
    if (in->vec_R && (!in->vec_S || !in->vec_T || !in->vec_U || !in->M)) { p->err = "cdp_prove_batch: missing instance vector"; return CDP_ERR_INVALID_ARG; }
    {   // witnesses: every permutation a bijection of 0..ell-1 (the reference indexes with it and panics out of range), scalars canonical
        std::vector<uint8_t> seen(ell);
        for (size_t pr = 0; pr < B; pr++) {
            std::fill(seen.begin(), seen.end(), 0);
            for (size_t i = 0; i < ell; i++) {
                const uint32_t v = in->permutation[pr * ell + i];
                if (v >= ell || seen[v]) { p->err = "cdp_prove_batch: permutation " + std::to_string(pr) + " is not a bijection of 0..ell-1"; return CDP_ERR_INVALID_ARG; }
                seen[v] = 1;
            }
            uint64_t w[4];
            for (int j = 0; j < 5; j++) {
                memcpy(w, j == 0 ? in->k + 32 * pr : in->vec_m_blinders + 128 * pr + 32 * (j - 1), 32);
                if (Fr::geq_mod(w)) { p->err = "cdp_prove_batch: witness scalar of proof " + std::to_string(pr) + " is not canonical"; return CDP_ERR_INVALID_ARG; }
            }
        }
    }
    // contiguous split, as even as possible
    std::vector<size_t> off(L + 1, 0);
    for (size_t i = 0; i < L; i++) off[i + 1] = off[i] + (B / L + (i < B % L ? 1 : 0));
    if (!in->vec_R && p->last_split != off) { p->err = "cdp_prove_batch: resident mode needs the same batch size as the staged call"; return CDP_ERR_INVALID_ARG; }
    p->last_split = off;
    std::vector<int> rcs(L, CDP_OK);
    double t0 = now_ms();
    auto run = [&](size_t i) {
        size_t o = off[i], cnt = off[i + 1] - off[i];
        if (cnt == 0) return;
        cdp_prove_inputs sub = *in;
        if (in->vec_R) {
            sub.vec_R = in->vec_R + o * ell * 96; sub.vec_S = in->vec_S + o * ell * 96;
            sub.vec_T = in->vec_T + o * ell * 96; sub.vec_U = in->vec_U + o * ell * 96;
            sub.M = in->M + o * 144;
        }
        sub.permutation = in->permutation + o * ell;
        sub.k = in->k + o * 32;
        sub.vec_m_blinders = in->vec_m_blinders + o * 128;
        sub.rng_seed = in->rng_seed ? in->rng_seed + o : nullptr;
        sub.rng_skip_words = in->rng_skip_words ? in->rng_skip_words + o : nullptr;
        sub.rng_key = in->rng_key ? in->rng_key + 32 * o : nullptr;
        rcs[i] = lane_prove(p->lanes[i], cnt, &sub, proofs_out + o * psz);
    };
    if (L == 1 || p->serial) {
        for (size_t i = 0; i < L; i++) run(i);
    } else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < L; i++) th.emplace_back(run, i);
        for (auto &t : th) t.join();
    }
    p->timing[0] = now_ms() - t0;
    p->timing[1] = p->timing[2] = p->timing[3] = 0;
    p->traffic[0] = p->traffic[1] = 0;
    for (size_t i = 0; i < L; i++) {
        if (rcs[i] != CDP_OK) { p->err = "lane " + std::to_string(i) + ": " + p->lanes[i]->err; return rcs[i]; }
        for (int k = 1; k < 4; k++) p->timing[k] = std::max(p->timing[k], p->lanes[i]->timing[k]);
        p->traffic[0] += p->lanes[i]->h2d_bytes;
        p->traffic[1] += p->lanes[i]->d2h_bytes;
    }
    return CDP_OK;
}

// =================================================================================================== whisk byte-level API
// generate_whisk_shuffle_proof (/root/reference/src/whisk.rs:144-179) for a batch: witnesses from the caller's rng stream, shuffled
// trackers and the permutation commitment on the GPU (decompression, k * R / k * S, M through the CRS digit table), the proof through
// cdp_prove_batch with the same stream handed on.
namespace {
// Montgomery one (R mod p): Z of an affine point lifted to Jacobian
const uint64_t FP_ONE_MONT[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL,
                                 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
void affine_to_jac(uint8_t out[144], const uint8_t aff[96]) {
    bool inf = true;
    for (int i = 0; i < 96; i++) inf = inf && aff[i] == 0;
    memcpy(out, aff, 96);
    if (inf) memset(out + 96, 0, 48);
    else memcpy(out + 96, FP_ONE_MONT, 48);
}
}  // namespace

extern "C" size_t cdp_whisk_shuffle_proof_size(size_t ell) { return 48 + cdp_proof_size(ell); }

extern "C" int cdp_whisk_generate_shuffle_proofs(cdp_prover *p, size_t B, const uint8_t *pre_trackers, const uint64_t *rng_seed,
                                                 const uint64_t *rng_skip_words, uint8_t *post_out, uint8_t *proofs_out) {
    if (!p) return CDP_ERR_INVALID_ARG;
    if (!pre_trackers || !post_out || !proofs_out || B == 0 || B > p->max_batch) { p->err = "cdp_whisk_generate_shuffle_proofs: bad argument"; return CDP_ERR_INVALID_ARG; }
    // rng_seed == NULL: every shuffle gets a fresh 256-bit ChaCha12 key from the operating system (what a production caller wants; the
    // u64 seeds reproduce the reference's test vectors and carry 64 bits of entropy at most)
    std::vector<uint8_t> keys;
    if (!rng_seed) {
        keys.resize(32 * B);
        for (size_t b = 0; b < B; b++)
            if (!StdRng::os_key(keys.data() + 32 * b)) { p->err = "cdp_whisk_generate_shuffle_proofs: no entropy source"; return CDP_ERR_INVALID_ARG; }
    }
    cdp_ctx *ctx = p->ctx0;
    const size_t ell = p->ell, n = ell + NBL, psz = cdp_proof_size(ell), np = B * ell;
    auto fail = [&](int rc, const char *what) { p->err = std::string(what) + ": " + cdp_last_error(ctx); return rc; };
    // witnesses, in the reference's draw order (whisk.rs:153-154, util.rs:102)
    std::vector<uint32_t> perm(np);
    std::vector<uint8_t> kk(32 * B), mbl(128 * B);
    std::vector<uint64_t> skip(B);
    for (size_t b = 0; b < B; b++) {
        StdRng rng = rng_seed ? StdRng(rng_seed[b]) : StdRng(StdRng::from_key_t{}, keys.data() + 32 * b);
        if (rng_skip_words) rng.skip_words(rng_skip_words[b]);
        uint32_t *pm = perm.data() + b * ell;
        for (size_t i = 0; i < ell; i++) pm[i] = (uint32_t)i;
        rng.shuffle_u32(pm, ell);
        rng.fr_rand().to_bytes(kk.data() + 32 * b);
        for (int i = 0; i < 4; i++) rng.fr_rand().to_bytes(mbl.data() + 128 * b + 32 * i);
        skip[b] = rng.words_consumed();  // CurdleproofsProof::new continues the same stream
    }
    // unzip_trackers: 2 * ell encodings per shuffle, r_G / k_r_G interleaved (whisk.rs:265-277)
    std::vector<uint8_t> aff(2 * np * 96), status(2 * np);
    int rc = cdp_decompress_batch(ctx, pre_trackers, 2 * np, aff.data(), status.data());
    if (rc != CDP_OK) return fail(rc, "unzip_trackers");
    // vec_T = sigma(k * vec_R), vec_U = sigma(k * vec_S)   (util.rs:94-97)
    std::vector<uint8_t> ks(2 * np * 32), kaff(2 * np * 96);
    for (size_t b = 0; b < B; b++)
        for (size_t i = 0; i < 2 * ell; i++) memcpy(ks.data() + (b * 2 * ell + i) * 32, kk.data() + 32 * b, 32);
    rc = cdp_scalar_mul_batch(ctx, aff.data(), ks.data(), 2 * np, kaff.data());
    if (rc != CDP_OK) return fail(rc, "k * trackers");
    std::vector<uint8_t> R(np * 96), S(np * 96), T(np * 96), U(np * 96), M(B * 144);
    for (size_t b = 0; b < B; b++)
        for (size_t i = 0; i < ell; i++) {
            const size_t o = b * ell + i, src = b * ell + perm[o];
            memcpy(R.data() + 96 * o, aff.data() + 96 * (2 * o), 96);
            memcpy(S.data() + 96 * o, aff.data() + 96 * (2 * o + 1), 96);
            memcpy(T.data() + 96 * o, kaff.data() + 96 * (2 * src), 96);
            memcpy(U.data() + 96 * o, kaff.data() + 96 * (2 * src + 1), 96);
        }
    // M = msm(vec_G, sigma as field elements) + msm(vec_H, m_blinders)   (util.rs:99-103): one MSM over the first n table bases
    std::vector<uint8_t> msc(32 * n);
    for (size_t b = 0; b < B; b++) {
        memset(msc.data(), 0, msc.size());
        for (size_t i = 0; i < ell; i++) memcpy(msc.data() + 32 * i, &perm[b * ell + i], 4);
        memcpy(msc.data() + 32 * ell, mbl.data() + 128 * b, 128);
        rc = cdp_msm_fixed(ctx, p->table, 0, msc.data(), n, M.data() + 144 * b);
        if (rc != CDP_OK) return fail(rc, "permutation commitment");
    }
    std::vector<uint8_t> proofs(B * psz);
    cdp_prove_inputs in;
    in.vec_R = R.data(); in.vec_S = S.data(); in.vec_T = T.data(); in.vec_U = U.data(); in.M = M.data();
    in.permutation = perm.data(); in.k = kk.data(); in.vec_m_blinders = mbl.data(); in.rng_seed = rng_seed; in.rng_skip_words = skip.data();
    in.rng_key = rng_seed ? nullptr : keys.data();
    rc = cdp_prove_batch(p, B, &in, proofs.data());
    if (rc != CDP_OK) return rc;
    // zip_trackers + serialisation: compress T, U (interleaved) and M
    std::vector<uint8_t> jac((2 * np + B) * 144), comp((2 * np + B) * 48);
    for (size_t o = 0; o < np; o++) {
        affine_to_jac(jac.data() + 144 * (2 * o), T.data() + 96 * o);
        affine_to_jac(jac.data() + 144 * (2 * o + 1), U.data() + 96 * o);
    }
    memcpy(jac.data() + 144 * 2 * np, M.data(), 144 * B);
    rc = cdp_compress_batch(ctx, jac.data(), 2 * np + B, comp.data());
    if (rc != CDP_OK) return fail(rc, "zip_trackers");
    memcpy(post_out, comp.data(), 2 * np * 48);
    for (size_t b = 0; b < B; b++) {
        memcpy(proofs_out + b * (48 + psz), comp.data() + 48 * (2 * np + b), 48);
        memcpy(proofs_out + b * (48 + psz) + 48, proofs.data() + b * psz, psz);
    }
    return CDP_OK;
}
