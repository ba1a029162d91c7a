// sm_100a kernels of the curdleproofs MSM / fold hot path.
//
//   k_smul_add        out[i] = A[i] + s[idx(i)] * P[i]        fold loops (inner_product_argument.rs:174-179,
//                                                             same_multiscalar_argument.rs:126-131), the CRS rescale
//                                                             (grand_product_argument.rs:92-102), shuffling (util.rs:94-95)
//   k_normalize       Jacobian -> affine (+ 48-byte encoding) `into_affine()` / `normalize_batch` (util.rs:27),
//                                                             `serialize_compressed` (transcript.rs:29-33)
//   k_msm_buckets     bucket phase of a batch of independent  `util::msm` (util.rs:19-22) at every call site of
//                     small MSMs, one CTA per MSM               CurdleproofsProof::new / verify
//   k_msm_combine     window combine (Horner) per MSM
//
// All of it is 32-bit integer multiply-add work on the FMA pipe (IMAD.WIDE.U32); no tensor cores.
// Host-side launchers, one per translation unit (the kernels are compiled in parallel, without -rdc).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#pragma GCC visibility push(default)
#include "../../include/cdp_msm.h"
#pragma GCC visibility pop

namespace cdp {

// resident CTAs per SM a kernel is compiled for (register budget 65536 / (128 * occ)); env override for tuning runs
inline int tuned_occupancy(const char *env, int dflt) {
    const char *e = getenv(env);
    int v = e ? atoi(e) : dflt;
    return v >= 3 && v <= 5 ? v : dflt;
}


struct msm_seg_t {
    uint32_t pts_off;      // first base, in points, into the launch's base array
    uint32_t scalars_off;  // first scalar, in scalars
    uint32_t n;
    uint32_t extra;        // 0 = none; otherwise 1 + index of one more base, logically element n, whose scalar is scalars[scalars_off + n]
};

struct smul_job_t {
    uint32_t src_off;        // first point to multiply, in points, into the launch's point array
    uint32_t add_off;        // first point to add (the fold's L half), or 0xFFFFFFFF for none
    uint32_t out_off;        // where the affine result goes
    uint32_t scalar_off;     // first scalar
    uint32_t scalar_stride;  // 0: one scalar for the whole job (fold), 1: one scalar per element
    uint32_t pad[3];
};

cudaError_t launch_smul_jobs(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, const smul_job_t *jobs, uint32_t n_jobs,
                             uint32_t elems_per_job, uint32_t *out_jac);
cudaError_t launch_normalize(cudaStream_t st, int chunk, const uint32_t *jac, uint32_t *out_affine, uint8_t *out_comp, uint32_t n,
                             const smul_job_t *jobs, uint32_t elems_per_job);
cudaError_t launch_compress_affine(cudaStream_t st, const uint32_t *pts, const uint32_t *idx, uint8_t *out_comp, uint32_t n);
cudaError_t launch_decompress(cudaStream_t st, const uint8_t *comp, const uint32_t *dst_idx, uint32_t *out_affine, uint8_t *status, uint32_t n);
struct round_expand_params_t {
    uint32_t n, h, spp, cpp, mode;  // vector length, split of this round, scalars per proof in the output / in the compact block, 0 = IPA, 1 = SameMSM
};
cudaError_t launch_round_expand(cudaStream_t st, const uint32_t *compact, const uint32_t *ucan, const round_expand_params_t &P, uint32_t batch,
                                uint32_t *out);
cudaError_t launch_sum_scalars(cudaStream_t st, const uint32_t *in, uint32_t stride, uint32_t cols, uint32_t rows, uint32_t *out);
cudaError_t launch_verify_transcript_a(cudaStream_t st, const uint8_t *pcomp, const uint8_t *pscal, const uint8_t *comp_M, const uint8_t *vec_a,
                                       uint32_t ell, uint32_t np, uint32_t vch, uint32_t B, uint64_t *state, uint32_t *chal, uint32_t *tmp,
                                       uint32_t *stage_scal, const uint8_t *comp_vecs, uint8_t *flags);
cudaError_t launch_verify_transcript_b(cudaStream_t st, const uint8_t *pcomp, const uint8_t *pscal, const uint8_t *da_comp, const uint8_t *comp_vecs,
                                       const uint8_t *H_comp, uint32_t ell, uint32_t m, uint32_t np, uint32_t vch, uint32_t B, uint64_t *state,
                                       uint32_t *chal, const uint32_t *tmp);
// k_vcoeffs.cu: the verifier's scalar preparation (slot layout of one proof's accumulated check; host/verifier.cpp builds it)
struct vcoef_params_t {
    uint32_t ell, n, m, big_n, vw, o_R, o_S, o_T, o_U, o_M, o_P, exact_eq, vch;
};
cudaError_t launch_verify_coeffs(cudaStream_t st, const uint32_t *chal, const uint32_t *vec_a, const vcoef_params_t &P, uint32_t batch, uint32_t *out_crs,
                                 uint32_t *out_var, uint32_t *out_ex);
cudaError_t launch_transcript_open(cudaStream_t st, const uint8_t *comp_vecs, const uint8_t *comp_M, uint32_t ell, uint32_t B, uint8_t *vec_a_out,
                                   uint64_t *state_out);
// k_prove.cu: one step of the prover's transcript + scalar algebra for a batch, one CTA per proof
cudaError_t launch_prove_random(cudaStream_t st, const uint32_t *keys, const uint64_t *skip_words, uint32_t ell, uint32_t batch, uint32_t *out);
cudaError_t launch_prove_stage(cudaStream_t st, const cdp_prove_dev &P, int stage, uint32_t round);
cudaError_t launch_gather_points(cudaStream_t st, uint32_t *pts, const uint32_t *src, const uint32_t *src_idx, const uint32_t *dst_idx, uint32_t n);
// window sums of `count` MSM segments; c in 2..6 selects the kernel instantiation
// `dig`: scratch for the digit rows, msm_dig_bytes(c, nmax, count) bytes
cudaError_t launch_msm_buckets(cudaStream_t st, int c, const uint32_t *pts, const uint32_t *scalars, const msm_seg_t *segs, uint32_t count,
                               uint32_t nmax, int8_t *dig, uint32_t *win_sums);
#define CDP_DECL_MSM(C)                                                                                                         \
    cudaError_t launch_msm_buckets_c##C(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, const msm_seg_t *segs, \
                                        uint32_t count, uint32_t nmax, int8_t *dig, uint32_t *win_sums);
CDP_DECL_MSM(2) CDP_DECL_MSM(3) CDP_DECL_MSM(4) CDP_DECL_MSM(5) CDP_DECL_MSM(6)
#undef CDP_DECL_MSM
cudaError_t launch_msm_combine(cudaStream_t st, const uint32_t *bucket_sums, uint32_t *out_jac, uint32_t n_msm, int c, int nwin);
cudaError_t launch_msm_reduce_quad(cudaStream_t st, const uint32_t *S, uint32_t *out_jac, uint32_t n_msm, int c);
uint32_t horner_groups_max();
size_t horner_groups_scratch_bytes(uint32_t n_msm, int c, int nwin);
cudaError_t launch_msm_horner_quad(cudaStream_t st, const uint32_t *bucket_sums, uint32_t *S_out, uint32_t n_msm, int c, int nwin,
                                   uint32_t *scratch = nullptr /* horner_groups_scratch_bytes, for n_msm <= horner_groups_max() */);
cudaError_t launch_sum_groups(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n_out, uint32_t per_out, uint32_t group_stride);
cudaError_t launch_sum_groups2(cudaStream_t st, const uint32_t *A, uint32_t per_a, uint32_t sa, uint32_t ga, const uint32_t *B, uint32_t per_b,
                               uint32_t sb, uint32_t gb, uint32_t *out, uint32_t n_out);

// large Pippenger MSM (k_bigmsm.cu)
size_t big_msm_sort_temp_bytes(uint32_t n2, int nwin, int c);
cudaError_t launch_big_digits(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, uint32_t n, int c, int nwin, uint32_t *keys, uint32_t *vals,
                              uint32_t *bx /* beta x per base, or nullptr */);
cudaError_t launch_big_sort(cudaStream_t st, void *temp, size_t temp_bytes, const uint32_t *keys, uint32_t *keys_out, const uint32_t *vals,
                            uint32_t *vals_out, uint32_t n2, int nwin, int c, const uint32_t *seg_offsets);
cudaError_t launch_big_offsets(cudaStream_t st, const uint32_t *keys_sorted, uint32_t n2, int nwin, uint32_t nb, int c, uint32_t *start);
size_t big_heavy_bytes(uint32_t n2, int nwin);
size_t big_order_temp_bytes(size_t slots);
cudaError_t launch_big_order(cudaStream_t st, const uint32_t *start, int nwin, uint32_t nb, uint32_t sp_top, uint32_t *order_ws, size_t temp_bytes);
cudaError_t launch_big_accumulate(cudaStream_t st, const uint32_t *pts, const uint32_t *vals_sorted, const uint32_t *start, uint32_t n2, int nwin,
                                  uint32_t nb, uint32_t sp_top, const uint32_t *order, uint32_t *buckets_jac, void *heavy_ws);
cudaError_t launch_big_reduce_level(cudaStream_t st, const uint32_t *Ain, const uint32_t *Bin, uint32_t n_out, uint32_t g, int shift, uint32_t *Aout,
                                    uint32_t *Bout);
cudaError_t launch_big_horner(cudaStream_t st, const uint32_t *A, const uint32_t *Bv, int nwin, int c, uint32_t *out_jac);
cudaError_t launch_big_bitsum_horner(cudaStream_t st, const uint32_t *A, const uint32_t *Bv, uint32_t N, int shift, int nwin, int c, uint32_t *sums,
                                     uint32_t *out_jac);
cudaError_t launch_big_fold_top(cudaStream_t st, const uint32_t *in, uint32_t nbt, uint32_t sp, uint32_t nw, uint32_t pad_to, uint32_t *out);
cudaError_t launch_big_reduce_leaf_affine(cudaStream_t st, const uint32_t *bucket_aff, uint32_t n_out, uint32_t g, uint32_t *Aout, uint32_t *Bout);
// bucket sums by rounds of batched affine additions (k_batchaff.cu)
size_t ba_scan_temp_bytes(size_t n);
constexpr int BA_STATS_ROUNDS = 32;  // k_ba_init writes (additions, elements kept, list length) for up to this many rounds
cudaError_t launch_ba_init(cudaStream_t st, const uint32_t *start, uint32_t n2, int nwin, uint32_t nb, const uint32_t *pts, const uint32_t *bx,
                           const uint32_t *vals, uint32_t *act, void *scan_in, uint32_t *bucket_aff, uint32_t *stats);
cudaError_t launch_ba_round(cudaStream_t st, bool first, void *scan_tmp, size_t scan_tmp_bytes, void *scan_in, void *scan_out, uint32_t list_len,
                            uint32_t pairs, uint32_t K, const uint32_t *act, uint32_t *nact, size_t act_stride, void *nscan_in, const uint32_t *pts,
                            const uint32_t *bx, const uint32_t *vals, const uint32_t *in, uint32_t *out, uint32_t *bucket_aff, void *jobs,
                            uint32_t *stash /* round 0 only: 192 B per pair, or nullptr */);
cudaError_t launch_ba_finish(cudaStream_t st, bool first, const uint32_t *act, size_t act_stride, uint32_t list_len, const uint32_t *pts, const uint32_t *bx,
                             const uint32_t *vals, const uint32_t *in, uint32_t *bucket_aff);
cudaError_t launch_bench_ba(cudaStream_t st, uint32_t *buf, uint32_t T, uint32_t K, bool fill, int inl);

// fixed-base MSM over a precomputed digit table (k_fixed.cu).  Table layout: affine points, 96 B each,
//   table[(base * nw + w) * nd + (d - 1)] = d * 2^(c w) * B_base      d = 1 .. nd = 2^(c-1),  w < nw = ceil(256 / c)
struct fixed_seg_t {
    uint32_t base_off;      // first base of the segment's range, in bases of the table
    uint32_t scalars_off;   // first scalar of the range
    uint32_t n;             // pairs the segment sums (without the extra one)
    uint32_t sel_h;         // 0: the range is 0..n-1.  Power of two h: only range positions j with (j & h) == sel_val take part
    uint32_t sel_val;       //    (n of them; position of the i-th one = i with bit h inserted) -- the L / R halves of a folded vector
    uint32_t remap_from;    // range positions j >= remap_from use base j + remap_delta (a base list with a gap)
    uint32_t remap_delta;
    uint32_t extra_base;    // 0 = none, else 1 + table base index of one more pair ...
    uint32_t extra_scalar;  // ... whose scalar is scalars[scalars_off + extra_scalar]
    uint32_t out_idx;       // result goes to out_jac[out_idx]
    uint32_t addv_off;      // plus addv_n (<= 32) device-resident affine points var_pts[addv_off ..] added as they are: short sums such as
    uint32_t addv_n;        // D = B - beta^-1 sum(G) + alpha sum(Hvec) or A' = A + T_1 + U_1 stay one launch
    uint32_t pos_off;       // range position j walks the bases at pos_off + j * pos_stride (0 = 1) instead of j; its scalar stays scalars[scalars_off + j]
    uint32_t pos_stride;
};
struct fixed_kparams_t {
    int c, nw;
    uint32_t nd;
    uint32_t es;         // distance between table entries in 32-bit words: 24 (packed) or 32 (one 128-byte line per entry)
    uint32_t recode[8];  // sum over w < nw-1 of 2^(c w + c - 1): added to the scalar, turns unsigned windows into signed digits
};
cudaError_t launch_fixed_pow(cudaStream_t st, const uint32_t *bases_affine, uint32_t n_bases, int c, int nw, uint32_t *jac_out);
cudaError_t launch_fixed_seed(cudaStream_t st, const uint32_t *aff, uint32_t chains, uint32_t nd, uint32_t es, uint32_t *table);
cudaError_t launch_fixed_level(cudaStream_t st, uint32_t *table, uint32_t chains, uint32_t nd, uint32_t half, uint32_t es);
cudaError_t launch_fixed_msm(cudaStream_t st, const uint32_t *table, const uint32_t *scalars, const fixed_seg_t *segs, uint32_t count,
                             const fixed_kparams_t &kp, const uint32_t *var_pts, uint32_t *out_jac, int lanes_per_seg = 32);
// the same sums as a tree of batched affine additions: `rounds` halving rounds, then the lane kernel over what is left (k_fixed.cu)
size_t fixed_ba_scratch_bytes(uint32_t count, uint32_t max_pairs, int nw, int rounds);
cudaError_t launch_fixed_msm_ba(cudaStream_t st, const uint32_t *table, const uint32_t *scalars, const fixed_seg_t *segs, uint32_t count,
                                const fixed_kparams_t &kp, const uint32_t *var_pts, uint32_t *out_jac, uint32_t max_pairs, int rounds, uint32_t *scratch,
                                uint32_t t_target, uint32_t kmax);

// which: 0 = raw IMAD.WIDE chains (128 multiply-adds / thread / iteration), 1 = Fp mul chain, 2 = Fp sqr chain (1 / thread / iteration)
cudaError_t launch_bench(cudaStream_t st, int which, uint32_t *out, int blocks, int threads, int iters);

// geometry shared by host and device
constexpr int msm_nwin_for(int c) { return (130 + c - 1) / c; }
// the bucket kernel gives every warp 64 / 2^(c-1) windows of one MSM (64 slots, two per lane); CTAs are 4 independent warps
constexpr int msm_nb_for(int c) { return 1 << (c - 1); }
constexpr int msm_wpw_for(int c) { return 64 / msm_nb_for(c); }
constexpr int msm_groups_for(int c) { return (msm_nwin_for(c) + msm_wpw_for(c) - 1) / msm_wpw_for(c); }
inline uint32_t msm_dig_rowstride(size_t nmax) { return (uint32_t)((2 * nmax + 15) & ~size_t(15)); }
// scratch of one small-MSM launch: the digit rows, then beta * x of every base (48 B each)
inline size_t msm_dig_rows_bytes(int c, size_t nmax, size_t count) { return (count * (size_t)msm_nwin_for(c) * msm_dig_rowstride(nmax) + 255) & ~size_t(255); }
inline size_t msm_dig_bytes(int c, size_t nmax, size_t count) { return msm_dig_rows_bytes(c, nmax, count) + count * nmax * 48; }
__host__ __device__ inline size_t msm_smem_per_warp(int c, size_t nmax) {
    size_t wpw = 64 >> (c - 1), dstride = ((2 * nmax + 15) & ~size_t(15)) + 16;
    return (wpw * dstride + wpw * 2 * nmax * sizeof(uint16_t) + 128 * sizeof(uint32_t) + 64 + 15) & ~size_t(15);
}
inline size_t msm_smem_bytes(int c, size_t nmax) { return 4 * msm_smem_per_warp(c, nmax); }
// bucket sums of `count` MSMs: [msm][window][bucket] Jacobian points
inline size_t msm_bucket_sums_bytes(int c, size_t count) { return count * (size_t)msm_nwin_for(c) * msm_nb_for(c) * 144; }

}  // namespace cdp
