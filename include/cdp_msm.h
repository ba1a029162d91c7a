/* cdp_msm.h -- C ABI of the B200 (sm_100a) engine for the Curdleproofs MSM / fold hot path.
 *
 * This is the drop-in boundary: every entry point replaces one reference-side operation on the path
 * BASELINE.json's north_star names, and is exactly what a Rust `extern "C"` block in the reference crate would bind
 * (see INTEGRATION.md for the stub).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Byte layouts (all little-endian 64-bit limbs, i.e. the in-memory form arkworks already holds):
 *   Fp        48 B   6 x u64, MONTGOMERY form, R = 2^384        (= ark-ff Fp384 backing array)
 *   scalar    32 B   4 x u64, CANONICAL integer in [0, r)        (= `Fr::into_bigint()`, what ark-ec's MSM consumes)
 *   affine    96 B   x || y ; the point at infinity is the all-zero encoding ((0,0) is not on y^2 = x^3 + 4)
 *   jacobian 144 B   X || Y || Z (x = X/Z^2, y = Y/Z^3) ; infinity <=> Z == 0   (= ark-ec `Projective`)
 *   compressed 48 B  big-endian x with ZCash flag bits (0x80 compressed, 0x40 infinity, 0x20 y > -y)
 *                    (= `serialize_compressed`, pinned by the KAT at /root/reference/src/whisk.rs:363-368)
 *
 * Error behaviour: the reference panics on length mismatch (`assert_eq!` src/util.rs:20,26); a C ABI cannot carry
 * slice lengths, so the caller passes one `n` for both arrays and every function returns a status code instead of
 * aborting.  0 = CDP_OK.  Nothing here ever falls back to a CPU implementation: without a usable CUDA device
 * `cdp_ctx_create` fails with CDP_ERR_CUDA.
 *
 * Threading: a context owns one CUDA stream and its scratch buffers; calls on one context are serialised by the
 * caller.  Use one context per host thread / per GPU.
 */
#ifndef CDP_MSM_H
#define CDP_MSM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDP_OK 0
#define CDP_ERR_INVALID_ARG 1
#define CDP_ERR_CUDA 2
#define CDP_ERR_TOO_LARGE 3
#define CDP_ERR_NOT_ON_CURVE 4
#define CDP_ERR_NCCL 5

#define CDP_FP_BYTES 48
#define CDP_SCALAR_BYTES 32
#define CDP_AFFINE_BYTES 96
#define CDP_JACOBIAN_BYTES 144
#define CDP_COMPRESSED_BYTES 48

typedef struct cdp_ctx cdp_ctx;

/* ------------------------------------------------------------------ context */
/* Create a context on CUDA device `device_id`.  `stream` may be NULL (the context creates its own stream) or an
 * existing `cudaStream_t` cast to void* (e.g. torch.cuda.current_stream().cuda_stream) that all work is queued on. */
int cdp_ctx_create(cdp_ctx **out, int device_id, void *stream);
void cdp_ctx_destroy(cdp_ctx *ctx);
/* Human-readable description of the last error on this context (never NULL). */
const char *cdp_last_error(const cdp_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's `gpu_launches`). */
uint64_t cdp_launch_count(const cdp_ctx *ctx);
/* CUDA device ordinal the context was created on. */
int cdp_ctx_device(const cdp_ctx *ctx);
/* Block until everything queued on the context's stream has finished. */
int cdp_sync(cdp_ctx *ctx);

/* ------------------------------------------------------------------ host-buffer entry points (the drop-ins) */
/* Replaces `util::msm(points: &[G1Affine], scalars: &[Fr]) -> G1Projective`      /root/reference/src/util.rs:19-22
 * out = sum_i scalars[i] * points[i].  n == 0 yields infinity. */
int cdp_msm(cdp_ctx *ctx, const uint8_t *affine_pts, const uint8_t *scalars, size_t n, uint8_t out_jac[CDP_JACOBIAN_BYTES]);

/* Replaces `util::msm_from_projective(points: &[G1Projective], scalars)`          /root/reference/src/util.rs:25-29
 * (batch-normalise, then MSM). */
int cdp_msm_from_projective(cdp_ctx *ctx, const uint8_t *jac_pts, const uint8_t *scalars, size_t n,
                            uint8_t out_jac[CDP_JACOBIAN_BYTES]);

/* Many independent MSMs in one launch -- what a prover round issues (4 per IPA round, src/inner_product_argument.rs:
 * 158-161; 6 per SameMSM round, src/same_multiscalar_argument.rs:107-112), across a batch of proofs. */
typedef struct {
    const uint8_t *affine_pts; /* n points  */
    const uint8_t *scalars;    /* n scalars */
    size_t n;
} cdp_msm_desc;
int cdp_msm_batch(cdp_ctx *ctx, const cdp_msm_desc *descs, size_t count, uint8_t *out_jac /* count * 144 B */);

/* The fold step, which the reference writes as an inline loop:
 *   L[i] = (L[i] + R[i].mul(gamma)).into_affine()      src/inner_product_argument.rs:177-178
 *                                                      src/same_multiscalar_argument.rs:128-130
 * out_affine may alias L. */
int cdp_fold(cdp_ctx *ctx, const uint8_t *L_affine, const uint8_t *R_affine, const uint8_t gamma[CDP_SCALAR_BYTES], size_t n,
             uint8_t *out_affine);

/* out[i] = (scalars[i] * pts[i]).into_affine()        src/grand_product_argument.rs:92-102 (CRS rescale),
 *                                                      src/util.rs:94-95 (shuffling), `GroupCommitment` src/commitments.rs:50-51 */
int cdp_scalar_mul_batch(cdp_ctx *ctx, const uint8_t *affine_pts, const uint8_t *scalars, size_t n, uint8_t *out_affine);

/* `G1Projective::normalize_batch` / `into_affine()`    src/util.rs:27 */
int cdp_normalize_batch(cdp_ctx *ctx, const uint8_t *jac_pts, size_t n, uint8_t *out_affine);

/* Jacobian points -> 48-byte compressed encodings (`serialize_compressed`, src/transcript.rs:29-33, src/util.rs:125-133) */
int cdp_compress_batch(cdp_ctx *ctx, const uint8_t *jac_pts, size_t n, uint8_t *out_compressed);

/* 48-byte encodings -> affine points (`G1Affine::deserialize_compressed`, src/whisk.rs:313-315; every point of
 * `CurdleproofsProof::deserialize`, src/curdleproofs.rs:312-323) with ark-serialize's validation: canonical x, on the
 * curve, in the prime-order subgroup.  status[i]: 0 ok, 1 malformed, 2 not on the curve, 3 not in the subgroup (the
 * output is then the all-zero point).  Returns CDP_ERR_NOT_ON_CURVE when any status is non-zero. */
int cdp_decompress_batch(cdp_ctx *ctx, const uint8_t *compressed, size_t n, uint8_t *out_affine, uint8_t *status);

/* ------------------------------------------------------------------ device-resident entry points
 * Same operations on buffers that already live in HBM (bases stay resident across the rounds of a proof batch).
 * All `d_*` arguments are device pointers on the context's device; work is queued on the context's stream and is
 * asynchronous -- call cdp_sync() (or synchronise the stream you supplied) before reading results. */
void *cdp_dev_alloc(cdp_ctx *ctx, size_t bytes);
void cdp_dev_free(cdp_ctx *ctx, void *d_ptr);
int cdp_h2d(cdp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int cdp_d2h(cdp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int cdp_dev_zero(cdp_ctx *ctx, void *d_ptr, size_t bytes); /* asynchronous memset(0) on the context's stream */
/* Strided copy (cudaMemcpy2DAsync): `height` rows of `width` bytes, rows spitch / dpitch bytes apart.  With a page-locked source it is a
 * direct DMA from the caller's buffer: the batched prover / verifier use it to lay the caller's vec_R | vec_S | vec_T | vec_U out proof-major
 * in HBM without a staging pass when cdp_host_is_pinned says the buffers are page-locked (cdp_host_alloc, cudaHostRegister, ...). */
int cdp_h2d_2d(cdp_ctx *ctx, void *d_dst, size_t dpitch, const void *h_src, size_t spitch, size_t width, size_t height);
int cdp_host_is_pinned(const void *h_ptr);
/* Pinned host memory for the asynchronous copies above. */
void *cdp_host_alloc(cdp_ctx *ctx, size_t bytes);
void cdp_host_free(cdp_ctx *ctx, void *h_ptr);

/* One MSM of any size over device-resident bases and scalars (n >= 1); asynchronous, result = one Jacobian point. */
int cdp_msm_dev(cdp_ctx *ctx, const uint8_t *d_affine_pts, const uint8_t *d_scalars, size_t n, uint8_t *d_out_jac);

/* One MSM (cdp_msm, cdp_msm_dev) takes the chunked small-MSM kernels below 2^16 pairs and the sort-based large Pippenger from there on
 * (the crossover measured on a B200).  This moves the crossover for one context (n_pairs >= 2048; 0 restores the default): tuning, and the
 * tests that must reach the large path at sizes the CPU oracle finishes in seconds. */
int cdp_set_big_msm_min(cdp_ctx *ctx, size_t n_pairs);
/* Inside the large Pippenger the bucket sums are rounds of batched affine additions (5M + 1S per addition, one shared inversion per
 * thread) from 2^19 pairs on, and one thread per bucket with XYZZ additions below.  Same purpose and argument rules as above. */
int cdp_set_big_ba_min(cdp_ctx *ctx, size_t n_pairs);

/* d_out_jac = sum of `count` Jacobian points.  The multi-GPU combine of a base-range-sharded MSM: every rank all-gathers the
 * 144-byte partial sums (NCCL has no elliptic-curve reduction op) and adds them locally. */
int cdp_sum_jacobian_dev(cdp_ctx *ctx, const uint8_t *d_jac_in, size_t count, uint8_t *d_out_jac);
/* d_out_jac[g] = sum_{s < per_out} d_jac_in[s * group_stride + g] for g < n_out: adds the partial sums of many MSMs at once (the
 * chunk sums and the fixed-base part of each proof's accumulated check, /root/reference/src/msm_accumulator.rs:55-68). */
int cdp_sum_groups_dev(cdp_ctx *ctx, const uint8_t *d_jac_in, size_t n_out, size_t per_out, size_t group_stride, uint8_t *d_out_jac);
/* Two sources with their own strides: d_out_jac[g] = sum_{s < per_a} d_a[s * sa + g * ga] + sum_{s < per_b} d_b[s * sb + g * gb]
 * (indices in points).  d_out_jac must not overlap the inputs. */
int cdp_sum_groups2_dev(cdp_ctx *ctx, const uint8_t *d_a, size_t per_a, size_t sa, size_t ga, const uint8_t *d_b, size_t per_b, size_t sb,
                        size_t gb, size_t n_out, uint8_t *d_out_jac);

/* A batch of MSMs over device-resident bases and scalars.  Segment i computes
 *   sum_{j < n} d_scalars[scalars_off + j] * d_pts[pts_off + j]   (offsets in elements, not bytes). */
typedef struct {
    uint32_t pts_off;
    uint32_t scalars_off;
    uint32_t n;
    uint32_t extra; /* 0 = none; else 1 + index (into d_affine_pts) of one more base, logically element n, whose scalar is
                       d_scalars[scalars_off + n].  Lets `msm(G_R, c_L) + ip * H` (src/inner_product_argument.rs:158) be ONE msm. */
} cdp_msm_seg;
/* `d_segs` is a DEVICE array of `count` cdp_msm_seg; `max_n` >= every segment's n.  Results: count Jacobian points. */
/* `total_pairs` = sum of the segments' lengths (incl. extras); used for profiling accounting only, may be 0. */
int cdp_msm_batch_dev(cdp_ctx *ctx, const uint8_t *d_affine_pts, const uint8_t *d_scalars, const cdp_msm_seg *d_segs, size_t count,
                      size_t max_n, size_t total_pairs, uint8_t *d_out_jac);

/* ------------------------------------------------------------------ fixed-base MSM (CRS bases)
 * `util::msm` (/root/reference/src/util.rs:19-22) for call sites whose `points` are CRS elements (crs.vec_G, vec_H, H, G_t, G_u,
 * /root/reference/src/crs.rs:19-34): they are identical for every proof, so the engine keeps a digit table of them in HBM
 *   table[base][w][d] = d * 2^(c w) * B_base,  d = 1 .. 2^(c-1),  w < ceil(256 / c)        (c = 16: 50 MB per base, c = 19: 352 MB)
 * and one (scalar, base) pair costs ceil(256 / c) gathered mixed additions.  Bases must be prime-order points or infinity;
 * scalars canonical (< r).  Results are the same group elements as cdp_msm over the same points.
 * A table is immutable after creation and may be used from any context of the same device concurrently. */
typedef struct cdp_fixed_table cdp_fixed_table;
int cdp_fixed_table_create(cdp_ctx *ctx, const uint8_t *affine_pts /* host */, size_t n_bases,
                           int window_bits /* 2..20; 0 = default 16 */,
                           cdp_fixed_table **out);
void cdp_fixed_table_destroy(cdp_ctx *ctx, cdp_fixed_table *t);
size_t cdp_fixed_table_bytes(const cdp_fixed_table *t);
size_t cdp_fixed_table_bases(const cdp_fixed_table *t);
/* sum_{j < n} scalars[j] * B[base_off + j] with host scalars: the drop-in for one `util::msm(&crs.vec_G[a..b], scalars)`. */
int cdp_msm_fixed(cdp_ctx *ctx, const cdp_fixed_table *t, size_t base_off, const uint8_t *scalars, size_t n, uint8_t out_jac[CDP_JACOBIAN_BYTES]);
/* A batch of fixed-base MSMs, device scalars.  Segment = the pairs (d_scalars[scalars_off + j], B[base_off + j + gap(j)]) over the
 * range positions j it selects:
 *   sel_h == 0: j = 0 .. n-1;   sel_h = power of two h: the n positions with (j & h) == sel_val (sel_val is 0 or h) -- the L / R
 *   half pattern of a vector folded down to 2h entries, which is how the IPA / SameMSM round MSMs over *folded* bases
 *   (src/inner_product_argument.rs:158-161, src/same_multiscalar_argument.rs:107-112) are written over the original ones;
 *   gap(j) = remap_delta for j >= remap_from (a base list that skips some table entries), else 0;
 *   pos_off / pos_stride: see the struct;
 *   extra_base != 0 adds the pair (d_scalars[scalars_off + extra_scalar], B[extra_base - 1]) -- the `+ ip * H` term;
 *   addv_n (<= 32) != 0 adds the device-resident affine points d_var_pts[addv_off .. addv_off + addv_n) with coefficient 1, so
 *   short sums like D = B - beta^-1 G_sum + alpha H_sum (src/grand_product_argument.rs:223) are one segment.
 * The result of segment i goes to d_out_jac[out_idx].  d_var_pts may be NULL when no segment has addv_n. */
typedef struct {
    uint32_t base_off, scalars_off, n, sel_h, sel_val, remap_from, remap_delta, extra_base, extra_scalar, out_idx;
    uint32_t addv_off, addv_n;
    uint32_t pos_off, pos_stride; /* the pair of range position j uses base position pos_off + j * pos_stride (pos_stride 0 = 1) in place of j
                                     -- also in the remap test -- while its scalar stays d_scalars[scalars_off + j]: a strided walk over the
                                     bases, e.g. the Q bases i, i + n', i + 2n', ... that fold into entry i of a vector folded down to n' entries */
} cdp_fixed_seg;
int cdp_msm_fixed_batch_dev(cdp_ctx *ctx, const cdp_fixed_table *t, const uint8_t *d_scalars, const cdp_fixed_seg *d_segs, size_t count,
                            size_t total_pairs, const uint8_t *d_var_pts, uint8_t *d_out_jac);
/* The same with 8, 16 or 32 (= the default of the call above) lanes of a warp working on one segment: fewer lanes per segment when many
 * equally long segments are in flight (longer per-lane addition chains, fewer fold steps per segment). */
int cdp_msm_fixed_batch_dev_lanes(cdp_ctx *ctx, const cdp_fixed_table *t, const uint8_t *d_scalars, const cdp_fixed_seg *d_segs, size_t count,
                                  size_t total_pairs, const uint8_t *d_var_pts, uint8_t *d_out_jac, int lanes_per_segment);
/* The same batch with every segment's table points summed as a TREE of batched affine additions (5M + 1S each, one shared inversion per
 * thread, against 8M + 2S for the lane accumulators): a few halving rounds over all segments at once, then the lane kernel over the partial
 * sums that are left.  For launches of many long segments; `max_pairs_per_segment` (the extra pair included) sizes the scratch and the
 * padding -- segments may be shorter.  Short segments fall back to cdp_msm_fixed_batch_dev_lanes.  Same results, bit for bit. */
int cdp_msm_fixed_batch_dev_tree(cdp_ctx *ctx, const cdp_fixed_table *t, const uint8_t *d_scalars, const cdp_fixed_seg *d_segs, size_t count,
                                 size_t total_pairs, const uint8_t *d_var_pts, uint8_t *d_out_jac, size_t max_pairs_per_segment);

/* Batched scalar multiplication / fold over device-resident points.  For every job j and element e < elems_per_job:
 *   d_pts[out_off + e] = ( (add_off != CDP_NONE ? d_pts[add_off + e] : O)
 *                          + d_scalars[scalar_off + e * scalar_stride] * d_pts[src_off + e] ).into_affine()
 * One job = one vector of one proof.  With src = R half, add = out = L half and scalar_stride = 0 this is the fold loop
 * (src/inner_product_argument.rs:174-179, src/same_multiscalar_argument.rs:126-131) for a whole batch of proofs in one
 * launch; with add = CDP_NONE and scalar_stride = 1 it is the CRS rescale (src/grand_product_argument.rs:92-102).
 * `d_jobs` is a DEVICE array.  out ranges may alias add ranges (results are staged before they are written). */
#define CDP_NONE 0xFFFFFFFFu
typedef struct {
    uint32_t src_off, add_off, out_off, scalar_off, scalar_stride;
    uint32_t reserved[3];
} cdp_smul_job;
int cdp_smul_jobs_dev(cdp_ctx *ctx, uint8_t *d_pts, const uint8_t *d_scalars, const cdp_smul_job *d_jobs, size_t n_jobs,
                      size_t elems_per_job);

/* d_pts[d_dst_idx[i]] = d_src[d_src_idx[i]] for i < n (96-byte points): assembles per-proof working vectors. */
int cdp_gather_dev(cdp_ctx *ctx, uint8_t *d_pts, const uint8_t *d_src, const uint32_t *d_src_idx, const uint32_t *d_dst_idx, size_t n);

/* Device-resident cdp_decompress_batch; point i is written to d_out_affine[d_dst_index ? d_dst_index[i] : i]. */
int cdp_decompress_dev(cdp_ctx *ctx, const uint8_t *d_compressed, const uint32_t *d_dst_index, size_t n, uint8_t *d_out_affine,
                       uint8_t *d_status);

/* Affine points (d_pts[d_index[i]], or d_pts[i] when d_index is NULL) -> 48-byte encodings (`serialize_compressed` of the
 * instance vectors, src/curdleproofs.rs:81). */
int cdp_compress_affine_dev(cdp_ctx *ctx, const uint8_t *d_pts, const uint32_t *d_index, size_t n, uint8_t *d_out_compressed);

/* Opening of the Fiat-Shamir transcript on the device, one thread per proof (merlin 3.0.0 / STROBE-128 / Keccak-f[1600]):
 *   Transcript::new(b"curdleproofs"); append_list(b"curdleproofs_step1", [vec_R, vec_S, vec_T, vec_U]); append(.., M);
 *   vec_a = get_and_append_challenges(b"curdleproofs_vec_a", ell)          /root/reference/src/curdleproofs.rs:78-83, :213-225
 * d_comp_vecs: batch x 4 x ell 48-byte encodings (proof-major: R | S | T | U), d_comp_M: batch encodings.
 * Outputs: d_vec_a = batch x ell canonical 32-byte scalars; d_state = batch x 208 bytes: the 200-byte STROBE state, then pos, pos_begin
 * (one byte each, then zero padding) -- the host continues the same transcript from it. */
#define CDP_TRANSCRIPT_STATE_BYTES 208
int cdp_transcript_open_dev(cdp_ctx *ctx, const uint8_t *d_comp_vecs, const uint8_t *d_comp_M, size_t ell, size_t batch, uint8_t *d_vec_a,
                            uint8_t *d_state);

/* The rest of the VERIFIER's transcript on the device, one thread per proof, continuing the state cdp_transcript_open_dev left
 * (`CurdleproofsProof::verify`, /root/reference/src/curdleproofs.rs:226-296, and the `verify` of the five arguments it calls: they only append
 * proof points / scalars and draw challenges).  Two steps, because the transcript needs D and A' from the GPU in between:
 *   _a: same_perm and gprod up to the gprod beta; writes challenge-block entries 12..15 (cdp_verify_coeffs_dev), d_tmp (batch x 2 scalars: the
 *       grand product and beta, for _b), d_stage_scalars (batch x 6 canonical: {1, -beta^-1, alpha_g, 1, 1, 1}, the scalars of
 *       D = B - beta^-1 sum(G) + alpha_g sum(Hvec) and A' = A + T_1 + U_1) and d_flags (batch bytes: 1 when vec_T[0] is the identity);
 *   _b: z, IPA, SameScalar, SameMSM; inverts the round challenges; completes entries 16..26 and the four challenge vectors.
 * d_proof_points: batch x (18 + 10 m) encodings, the proof's points in serialisation order; d_proof_scalars: batch x 7 canonical scalars
 * (r_p, c_final, d_final, z_k, z_t, z_u, x_final); d_comp_vecs / d_comp_M / d_vec_a: as in cdp_transcript_open_dev; d_comp_DA: batch x 2
 * encodings (D, A'); d_comp_H: the encoding of crs.H; d_state: batch x 208 bytes, in/out. */
int cdp_verify_transcript_a_dev(cdp_ctx *ctx, const uint8_t *d_proof_points, const uint8_t *d_proof_scalars, const uint8_t *d_comp_vecs,
                                const uint8_t *d_comp_M, const uint8_t *d_vec_a, size_t ell, size_t batch, uint8_t *d_state, uint8_t *d_challenges,
                                uint8_t *d_tmp, uint8_t *d_stage_scalars, uint8_t *d_flags);
int cdp_verify_transcript_b_dev(cdp_ctx *ctx, const uint8_t *d_proof_points, const uint8_t *d_proof_scalars, const uint8_t *d_comp_vecs,
                                const uint8_t *d_comp_DA, const uint8_t *d_comp_H, size_t ell, size_t batch, uint8_t *d_state, uint8_t *d_challenges,
                                const uint8_t *d_tmp);

/* Verifier scalar preparation on the device: from a proof's Fiat-Shamir challenges to the coefficient of every
 * base of its accumulated check -- the verification scalars s_i / 1/s_i of `get_verification_scalars_bitstring`
 * (/root/reference/src/util.rs:40-64, used at src/inner_product_argument.rs:202-250 and src/same_multiscalar_argument.rs:242-259),
 * the GrandProduct rescaling beta^-(i+1) (src/grand_product_argument.rs:92-102) and the `a * x_i` products of the
 * `MsmAccumulator::accumulate_check` calls (src/msm_accumulator.rs:37-52), with the accumulator's HashMap replaced by fixed slots.
 * d_challenges: batch x vch scalars, 4 x u64 MONTGOMERY form (R = 2^256) as the host's Fr keeps them:
 *   [0..11] the random factors of the 8 checks + 4 SameScalar equalities, [12] same_perm alpha, [13] same_perm beta, [14] gprod alpha,
 *   [15] gprod beta^-1, [16] ipa alpha, [17] ipa beta, [18] z, [19] c_final, [20] d_final, [21] x_final, [22] same_msm alpha,
 *   [23] same_scalar alpha, [24] z_k, [25] z_t, [26] z_u, then ipa gamma[m], ipa gamma^-1[m], same_msm gamma[m], same_msm gamma^-1[m].
 * d_vec_a: batch x ell canonical scalars (as cdp_transcript_open_dev leaves them).
 * Outputs (canonical scalars): d_crs_scalars[batch][n + 5] = G | Hvec | H | G_t | G_u | (sum(G), sum(Hvec): zero, folded into the G_i /
 *   Hvec_i); d_var_scalars[pr * vw + slot] for the per-proof slots n + 5 <= slot < big_n: R, S, T, U (ell each) at o_R .. o_U, M at o_M, the
 *   proof's points in serialisation order at o_P -- the index the point of that slot has in the verifier's base array, so that one MSM
 *   over the whole array is the merged check of a batch; d_exact_scalars[batch][14] = the exact form of the SameScalar equalities.
 *   exact_eq != 0 keeps those four equalities out of the accumulated check. */
typedef struct {
    uint32_t ell, n, m, big_n, vw, o_R, o_S, o_T, o_U, o_M, o_P, exact_eq, vch /* = 27 + 4 m */;
} cdp_vcoef_params;
int cdp_verify_coeffs_dev(cdp_ctx *ctx, const uint8_t *d_challenges, const uint8_t *d_vec_a, const cdp_vcoef_params *params, size_t batch,
                          uint8_t *d_crs_scalars, uint8_t *d_var_scalars, uint8_t *d_exact_scalars);
/* Prover: the scalars of one folding round expanded on the device.  The batched prover writes the round MSMs of the IPA / SameMSM arguments
 * (/root/reference/src/inner_product_argument.rs:158-161, src/same_multiscalar_argument.rs:107-112) over the ORIGINAL bases, so scalar j of a
 * round with split h is  w(j / 2h) * v[j mod h (+ h when bit h of j is clear)]: the host sends the n / 2h prefix weights and the 2h entries of
 * the folded vector, this forms the products.  mode 0 (IPA): compact = Wc[Q] canonical | c[2h] Montgomery | Wd[Q] Montgomery | d[2h] Montgomery
 * | ipL | ipR canonical; out = c-scalars[n] | ipL | ipR | d-scalars[n] with d-scalar j = Wd[q] u[j] d[..] (d_u_canonical: batch x n, the
 * GrandProduct rescaling beta^-(j+1) of src/grand_product_argument.rs:92-102).  mode 1 (SameMSM): compact = Ws[Q] canonical | x[2h]
 * Montgomery; out = x-scalars[n] | x[2h] canonical.  All outputs are canonical 32-byte scalars. */
int cdp_round_expand_dev(cdp_ctx *ctx, const uint8_t *d_compact, const uint8_t *d_u_canonical, size_t n, size_t h, size_t scalars_per_proof,
                         size_t compact_per_proof, int mode, size_t batch, uint8_t *d_scalars_out);

/* ------------------------------------------------------------------ the prover's transcript + scalar algebra on the device
 * SURVEY.md 8(f) rank 1, prover half.  Between two group launches `CurdleproofsProof::new` (/root/reference/src/curdleproofs.rs:59-184) appends
 * the new points to the merlin transcript, draws challenges (/root/reference/src/transcript.rs:28-61) and does O(n) arithmetic in Fr to form
 * the next launch's scalars.  cdp_prove_stage_dev runs one such step for a whole batch, one CTA per proof (csrc/k_prove.cu), so that a batch
 * goes from cdp_transcript_open_dev to the serialised proof bytes without a host round trip:
 *   CDP_PS_S1         witness vectors and blinders (curdleproofs.rs:86-116)                 -> scalars of A, R, S, B_a, B_t, B_u, T_1, U_1, A_1, B_1
 *   CDP_PS_SAMEPERM   same_permutation_argument.rs:60-82                                    -> scalars of B, T_2, A_2, U_2, B_2, A'
 *   CDP_PS_GPROD1     grand_product_argument.rs:63-83                                       -> scalars of C
 *   CDP_PS_GPROD2     grand_product_argument.rs:85-147, inner_product_argument.rs:53-77     -> scalars of D, B_c, B_d
 *   CDP_PS_IPA0       inner_product_argument.rs:129-140                                     -> scalars of IPA round 0
 *   CDP_PS_IPA_ROUND  inner_product_argument.rs:150-186 for `round`; the last round goes on through same_scalar_argument.rs:64-75 and
 *                     same_multiscalar_argument.rs:84-91                                    -> scalars of IPA round + 1 / SameMSM round 0
 *   CDP_PS_SM_ROUND   same_multiscalar_argument.rs:99-136 for `round`                       -> fold scalar gamma + scalars of SameMSM round + 1
 * Scalar layouts per stage are those of host/prover.cpp's stage tables (round MSMs over the ORIGINAL CRS bases).  Every stage also writes
 * the points it consumed, and the proof's scalars, into d_proofs (`CurdleproofsProof::serialize` layout). */
enum { CDP_PS_S1 = 0, CDP_PS_SAMEPERM, CDP_PS_GPROD1, CDP_PS_GPROD2, CDP_PS_IPA0, CDP_PS_IPA_ROUND, CDP_PS_SM_ROUND };
typedef struct {
    uint32_t ell, m, batch;          /* m = log2(ell + 4) */
    uint32_t proof_bytes;            /* cdp_proof_size(ell): stride of d_proofs */
    uint32_t scalars_per_proof;      /* stride (in scalars) of d_scalars for the stage being EMITTED */
    uint32_t switch_round;           /* k0 in 1 .. m: rounds k < k0 of the IPA / SameMSM arguments are written over the ORIGINAL CRS bases (scalars =
                                        prefix weight x folded vector entry); the step before round k0 also emits the scalars that MATERIALISE the
                                        folded bases (n each for G, G' -- and for G_with_blinders in the SameMSM argument --, entry (i, q) at i 2^k0 + q),
                                        and rounds k >= k0 run over those folded vectors like the reference's: scalars at offset 2n
                                        (c_L | ipL | c_R | ipR | d) resp. n (x), the fold challenges (gamma, gamma^-1 | gamma) in d_fold_scalars.
                                        k0 = m: never switch */
    uint32_t out_map[12];            /* where output q of the CONSUMED stage sits in d_comp, per proof pr: encoding index
                                        batch * (e >> 16) + pr * ((e >> 8) & 255) + (e & 255) */
    uint8_t *d_state;                /* batch x 208 B, the STROBE states cdp_transcript_open_dev left; in/out */
    const uint8_t *d_vec_a;          /* batch x ell canonical scalars (cdp_transcript_open_dev) */
    const uint32_t *d_perm;          /* batch x ell: the permutation witness */
    const uint8_t *d_witness;        /* batch x 5 canonical scalars: k, vec_m_blinders[4] */
    uint8_t *d_random;               /* batch x cdp_prove_random_scalars(ell) raw `Fr::rand` outputs (Montgomery representations), draw order:
                                        r_a[2] (curdleproofs.rs:86), r_c[4] (grand_product_argument.rs:75), ipa r_c[n], r_d[n] (the last two
                                        are solved for on the device, inner_product_argument.rs:46-77), r_t, r_u (curdleproofs.rs:110-111),
                                        r_a, r_b, r_k (same_scalar_argument.rs:56-58), same_msm r[n] (same_multiscalar_argument.rs:78) */
    uint8_t *d_work;                 /* batch x cdp_prove_work_scalars(ell) x 32 B of per-proof state, opaque */
    const uint8_t *d_comp0_vecs;     /* batch x 4 x ell encodings of the instance (R | S | T | U), d_comp0_M: batch encodings of M */
    const uint8_t *d_comp0_M;
    const uint8_t *d_comp_H;         /* the encoding of crs.H */
    const uint8_t *d_comp;           /* the consumed stage's output encodings (cdp_normalize_dev) */
    uint8_t *d_side;                 /* batch x 2 encodings kept for later transcript messages: A', D */
    uint8_t *d_proofs;               /* batch x proof_bytes */
    uint8_t *d_scalars;              /* out: batch x scalars_per_proof canonical scalars */
    uint8_t *d_fold_scalars;         /* out: batch x 2 canonical scalars: gamma (| gamma^-1 in IPA rounds >= switch_round) of the fold launches */
} cdp_prove_dev;
size_t cdp_prove_work_scalars(size_t ell);
size_t cdp_prove_random_scalars(size_t ell);
int cdp_prove_stage_dev(cdp_ctx *ctx, const cdp_prove_dev *params, int stage, unsigned round);
/* The prover's `rng` on the device (`rng: &mut impl RngCore`, /root/reference/src/curdleproofs.rs:74, as rand 0.8's StdRng = ChaCha12): fills
 * d_random (batch x cdp_prove_random_scalars(ell) x 32 B, the layout cdp_prove_dev.d_random documents) with the `Fr::rand` draws of
 * `CurdleproofsProof::new` in the reference's order, proof i from the keystream of the 32-byte key d_keys[32 i ..] starting at 32-bit word
 * d_skip_words[i] of it (NULL = 0: a fresh generator).  Same values as the host generator (host/rng.hpp) for the same key and position. */
int cdp_prove_random_dev(cdp_ctx *ctx, const uint8_t *d_keys, const uint64_t *d_skip_words, size_t batch, size_t ell, uint8_t *d_random);

/* d_out[i] = sum over r < rows of d_scalars[r * row_stride + i] (mod r), canonical 32-byte scalars, i < cols: the coefficients that several
 * proofs put on the same CRS base, added up for the merged check (`*entry += a * x_i`, /root/reference/src/msm_accumulator.rs:47-51). */
int cdp_sum_scalars_dev(cdp_ctx *ctx, const uint8_t *d_scalars, size_t row_stride, size_t cols, size_t rows, uint8_t *d_out);

/* Jacobian -> affine and/or compressed (either output may be NULL). d_out_affine may alias nothing in d_jac. */
int cdp_normalize_dev(cdp_ctx *ctx, const uint8_t *d_jac, size_t n, uint8_t *d_out_affine, uint8_t *d_out_compressed);

/* ------------------------------------------------------------------ multi-GPU: one large MSM sharded by base range
 * `util::msm` (/root/reference/src/util.rs:19-22) over N = sum of n_local pairs, rank r holding a contiguous range of the
 * bases and their scalars.  Every rank runs the large Pippenger on its shard; the only collective is ONE ncclAllGather of the 144-byte
 * partial sums (NCCL has no elliptic-curve reduction op, so the "allreduce" of partial G1 sums is all-gather + local add); every rank ends
 * with the full sum.  NCCL (>= 2.x, `libnccl.so.2`, or the path in $CDP_NCCL_LIB) is loaded at the first use, so single-GPU callers do not
 * depend on it; every NCCL failure is reported as CDP_ERR_NCCL with ncclGetErrorString in cdp_last_error.
 *
 * One process per GPU (the layout bench.py and `torch.distributed.run` use): rank 0 calls cdp_comm_unique_id and distributes the 128 bytes
 * out of band (exactly like ncclGetUniqueId), then every rank calls cdp_comm_create with its own context.
 * One process driving several GPUs: cdp_comm_create_all over one context per device, and cdp_msm_sharded_group to issue the collective
 * for all of them (ncclGroupStart / End around the per-device calls). */
typedef struct cdp_comm cdp_comm;
#define CDP_COMM_ID_BYTES 128
int cdp_comm_unique_id(uint8_t out_id[CDP_COMM_ID_BYTES]);
int cdp_comm_create(cdp_comm **out, cdp_ctx *ctx, const uint8_t id[CDP_COMM_ID_BYTES], int n_ranks, int rank);
int cdp_comm_create_all(cdp_comm **out /* n */, cdp_ctx *const *ctxs, int n);
void cdp_comm_destroy(cdp_comm *comm);
int cdp_comm_rank(const cdp_comm *comm);
int cdp_comm_size(const cdp_comm *comm);
cdp_ctx *cdp_comm_ctx(const cdp_comm *comm); /* the context the communicator is bound to */
const char *cdp_comm_last_error(const cdp_comm *comm);
/* The contiguous base range [lo, hi) of `rank` for an MSM of n pairs, as even as possible (the partition every caller should use). */
void cdp_shard_range(size_t n, int rank, int n_ranks, size_t *lo, size_t *hi);
/* d_out_jac (144 B, on every rank) = sum over all ranks of msm(shard).  n_local may be 0.  Asynchronous on the context's stream: the local
 * MSM, the all-gather and the final addition are ordered on that one stream. */
int cdp_msm_sharded_dev(cdp_comm *comm, const uint8_t *d_affine_pts_shard, const uint8_t *d_scalars_shard, size_t n_local, uint8_t *d_out_jac);
/* The same for the n communicators of cdp_comm_create_all (arrays of n device pointers, one per communicator / device). */
int cdp_msm_sharded_group(cdp_comm *const *comms, int n, const uint8_t *const *d_affine_pts_shard, const uint8_t *const *d_scalars_shard,
                          const size_t *n_local, uint8_t *const *d_out_jac);
/* Partial sums that were computed elsewhere (the merged accumulated check of a verifier's sub-batches, /root/reference/src/msm_accumulator.rs:55-68):
 * d_out_jac = sum over ranks of d_partial_jac (one Jacobian point per rank).  d_partial_jac and d_out_jac may alias. */
int cdp_allreduce_jacobian_dev(cdp_comm *comm, const uint8_t *d_partial_jac, uint8_t *d_out_jac);

/* ------------------------------------------------------------------ per-kernel profiling
 * When enabled, every kernel launch of the context is bracketed by CUDA events on the context's stream and the elapsed
 * device time is accumulated per kernel kind.  `units` accumulates the work items each launch processed: (scalar, point)
 * pairs for the MSM bucket and fixed-base kernels, MSMs for the combine kernel, elements for smul / normalise.  bench.py derives the
 * roofline figures from these. */
#define CDP_PROFILE_MSM_BUCKETS 0
#define CDP_PROFILE_MSM_COMBINE 1
#define CDP_PROFILE_SMUL 2
#define CDP_PROFILE_NORMALIZE 3
#define CDP_PROFILE_OTHER 4
#define CDP_PROFILE_MSM_FIXED 5
#define CDP_PROFILE_PROVE_STAGE 6 /* cdp_prove_stage_dev: transcript + Fr algebra of the prover */
#define CDP_PROFILE_TRANSCRIPT 7  /* transcript opening / verifier transcript kernels */
#define CDP_PROFILE_KINDS 8
int cdp_profile_enable(cdp_ctx *ctx, int on);
int cdp_profile_reset(cdp_ctx *ctx);
int cdp_profile_read(cdp_ctx *ctx, double ms[CDP_PROFILE_KINDS], uint64_t launches[CDP_PROFILE_KINDS], uint64_t units[CDP_PROFILE_KINDS]);

/* ------------------------------------------------------------------ diagnostics
 * Integer-pipe micro-benchmarks used by bench.py for the roofline denominators.
 * which = 0: independent IMAD.WIDE.U32 chains (128 multiply-adds per thread per iteration);
 * which = 1 / 2: dependent chain of Fp Montgomery multiplications / squarings (one per thread per iteration).
 * Writes the kernel time (CUDA events on the context's stream) to *ms_out. */
int cdp_bench_kernel(cdp_ctx *ctx, int which, int blocks, int threads, int iters, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* CDP_MSM_H */
