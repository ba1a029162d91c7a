/* cdp_prover.h -- C ABI of the batched Curdleproofs prover / verifier host driver.
 *
 * The driver restates `CurdleproofsProof::new` (/root/reference/src/curdleproofs.rs:59-184) and, through it, the
 * SamePermutation / GrandProduct / InnerProduct / SameScalar / SameMultiscalar provers, for a BATCH of independent
 * shuffles that advance in lock-step through the Fiat-Shamir rounds.  All group arithmetic (every MSM, every fold,
 * every normalisation / compression) runs on the GPU through include/cdp_msm.h; the host keeps only the transcript
 * (merlin), the prover's RNG stream and O(n) scalar-field bookkeeping.  Proofs are byte-identical to the reference's
 * for the same inputs and the same RNG stream.
 *
 * Layouts are those of cdp_msm.h: affine 96 B (Montgomery), jacobian 144 B, scalars 32 B canonical little-endian.
 */
#ifndef CDP_PROVER_H
#define CDP_PROVER_H
#include <stddef.h>
#include <stdint.h>

#include "cdp_msm.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdp_prover cdp_prover;

/* Serialised proof size, 1088 + 480 * log2(ell + 4) bytes (`CurdleproofsProof::serialize`, src/curdleproofs.rs:300-310). */
size_t cdp_proof_size(size_t ell);

/* `crs_points`: ell + 7 affine points in `CurdleproofsCrs::from_points` order (src/crs.rs:37-58):
 * vec_G (ell) | vec_H (4) | H | G_t | G_u.  ell + 4 must be a power of two (src/inner_product_argument.rs:116).
 * `ctx` must outlive the prover.  `host_threads` <= 0 selects all host cores. */
int cdp_prover_create(cdp_prover **out, cdp_ctx *ctx, size_t ell, const uint8_t *crs_points, size_t max_batch, int host_threads);
/* Same, with an explicit number of concurrent lanes.  A lane is an independent sub-batch pipeline with its own CUDA stream
 * (lane 0 uses `ctx`, the others create private contexts on the same device) and its own host threads, so that the copies, the
 * staging and the latency-bound launches of one lane overlap with GPU work of the others.  lanes <= 0 picks a default (2 for
 * max_batch >= 256, else 1: with the whole protocol on the device larger launches beat more overlap).  cdp_prover_create is
 * cdp_prover_create_lanes with lanes = 0. */
int cdp_prover_create_lanes(cdp_prover **out, cdp_ctx *ctx, size_t ell, const uint8_t *crs_points, size_t max_batch, int host_threads,
                            int lanes);
int cdp_prover_lane_count(const cdp_prover *p);
/* Diagnostics: run the lanes of the following cdp_prove_batch calls one after the other instead of concurrently, so that the per-kernel
 * device times of cdp_profile_* are those of kernels running alone (bench.py's roofline pass). */
void cdp_prover_set_serial(cdp_prover *p, int on);
/* Size of the CRS digit table the prover uses.  The table is shared by every prover / verifier of the process created over the same CRS
 * on the same device (reference-counted), so a verifier next to a prover costs no second table. */
size_t cdp_prover_table_bytes(const cdp_prover *p);
/* The cdp_ctx a lane runs on (for cdp_profile_* / cdp_launch_count accounting across lanes). */
cdp_ctx *cdp_prover_lane_ctx(const cdp_prover *p, int lane);
void cdp_prover_destroy(cdp_prover *p);
const char *cdp_prover_last_error(const cdp_prover *p);

/* Inputs of `batch` independent shuffles, proof-major.  Argument names follow CurdleproofsProof::new. */
typedef struct {
    const uint8_t *vec_R;          /* batch * ell affine   */
    const uint8_t *vec_S;          /* batch * ell affine   */
    const uint8_t *vec_T;          /* batch * ell affine   */
    const uint8_t *vec_U;          /* batch * ell affine   */
    const uint8_t *M;              /* batch jacobian       */
    const uint32_t *permutation;   /* batch * ell          */
    const uint8_t *k;              /* batch scalars        */
    const uint8_t *vec_m_blinders; /* batch * 4 scalars    */
    const uint64_t *rng_seed;      /* batch: the prover's `rng` is StdRng::seed_from_u64(rng_seed[i]) ...            */
    const uint64_t *rng_skip_words;/* ... advanced by this many u32 outputs first (NULL = 0): lets a caller that already
                                      consumed part of the stream (src/whisk.rs:153-154, src/util.rs:102) hand it over. */
    const uint8_t *rng_key;        /* batch * 32 bytes, or NULL.  When set it replaces rng_seed: the prover's `rng` is
                                      StdRng::from_seed(rng_key[i]) (a full 256-bit ChaCha12 key, e.g. from the caller's OsRng).
                                      rng_seed carries only 64 bits of entropy -- a proof made from a guessable seed leaks its witness to
                                      anyone who replays the seed -- and is meant for reproducing test vectors.  With rng_seed == NULL and
                                      rng_key == NULL every proof draws a fresh key from the operating system's entropy source. */
} cdp_prove_inputs;

/* proofs_out: batch * cdp_proof_size(ell) bytes.  Host buffers in, host buffers out (copies are inside the call).
 * If in->vec_R is NULL the instance vectors (R, S, T, U, M) staged in HBM by the previous call are reused -- the
 * "inputs already resident" mode; witnesses and rng fields are still read from `in`.
 * Witnesses are validated before any work starts: every permutation must be a bijection of 0..ell-1 and k / vec_m_blinders canonical
 * (< r); otherwise CDP_ERR_INVALID_ARG (the reference panics on an out-of-range index). */
int cdp_prove_batch(cdp_prover *p, size_t batch, const cdp_prove_inputs *in, uint8_t *proofs_out);

/* Timing breakdown of the last cdp_prove_batch call, milliseconds: [0] total, [1] host transcript/scalar work,
 * [2] waiting on the GPU (stream synchronisation), [3] H2D/D2H staging issue time. */
void cdp_prover_last_timing(const cdp_prover *p, double out_ms[4]);

/* ------------------------------------------------------------------ batched verifier
 * `CurdleproofsProof::deserialize` + `verify` (src/curdleproofs.rs:197-323) for `batch` independent proofs.
 * Proof points are decompressed and subgroup-checked on the GPU; the eight accumulated checks of a proof become one MSM
 * over [CRS | R | S | T | U | M | proof points] compared with the identity (the reference's MsmAccumulator, without the
 * HashMap); SameScalar's four point equalities join that check with their own random factors (CDP_VERIFY_EXACT_EQ=1: four exact MSMs).
 * The transcript (cdp_transcript_open_dev, cdp_verify_transcript_{a,b}_dev; CDP_VERIFY_HOST_TRANSCRIPT=1 keeps the per-round part on the
 * host) and the coefficients (cdp_verify_coeffs_dev) are computed on the device.  By default a lane first runs the MERGED check of its whole
 * sub-batch -- the per-proof bases of all its proofs in ONE large MSM plus the summed CRS parts, every check already carrying its own
 * random factor -- and accepts all of them when that is the identity; otherwise (or when a proof of the sub-batch is malformed) every
 * proof is decided by its own accumulated MSM, so the verdicts never depend on the mode.  CDP_VERIFY_MERGE=0 disables the merged check. */
typedef struct cdp_verifier cdp_verifier;
int cdp_verifier_create(cdp_verifier **out, cdp_ctx *ctx, size_t ell, const uint8_t *crs_points, size_t max_batch, int host_threads, int lanes);
void cdp_verifier_destroy(cdp_verifier *v);
const char *cdp_verifier_last_error(const cdp_verifier *v);
/* Lanes (concurrent sub-batch pipelines; a batch is split over them contiguously, as evenly as possible). */
int cdp_verifier_lane_count(const cdp_verifier *v);
/* Timing of the last cdp_verify_batch call, max over lanes, in ms: total, host compute (transcripts, coefficients), waiting for the GPU. */
void cdp_verifier_last_timing(const cdp_verifier *v, double out_ms[3]);
/* Since creation: [0] lane sub-batches accepted by the merged check, [1] lane sub-batches that fell back to proof-by-proof checks. */
void cdp_verifier_merge_stats(const cdp_verifier *v, uint64_t out[2]);
typedef struct {
    const uint8_t *vec_R;     /* batch * ell affine */
    const uint8_t *vec_S;
    const uint8_t *vec_T;
    const uint8_t *vec_U;
    const uint8_t *M;         /* batch jacobian */
    const uint8_t *proofs;    /* batch * cdp_proof_size(ell) bytes, `CurdleproofsProof::serialize` format */
    const uint64_t *rng_seed; /* batch: seeds of the random factors of the accumulated checks (msm_accumulator.rs:44; the `rng` argument of
                                 `verify`), or NULL: fresh seeds from the OS entropy source.  They must be unpredictable to the prover. */
} cdp_verify_inputs;
/* result[i]: 1 = Ok(()), 0 = Err(VerificationError), 2 = the proof does not deserialise (bad encoding / not in the subgroup) */
int cdp_verify_batch(cdp_verifier *v, size_t batch, const cdp_verify_inputs *in, uint8_t *result);
/* The accumulated check of a batch that is spread over the GPUs of a node (`MsmAccumulator::verify`,
 * /root/reference/src/msm_accumulator.rs:55-68): every rank verifies ITS proofs (its share of the bases of the one large accumulated MSM),
 * the merged sums of all sub-batches of all ranks are added with ONE all-gather of the 144-byte partial sums (cdp_allreduce_jacobian_dev),
 * and the identity accepts every proof of every rank at once; otherwise -- or for sub-batches holding a malformed proof -- each rank
 * decides its own proofs exactly as cdp_verify_batch does, so the verdicts never depend on the mode.  Collective: every rank of `comm`
 * must call it the same number of times.  `comm` must have been created on the verifier's context. */
int cdp_verify_batch_sharded(cdp_verifier *v, cdp_comm *comm, size_t batch, const cdp_verify_inputs *in, uint8_t *result);
/* Since creation: [0] sharded calls accepted by the cross-rank sum, [1] sharded calls that fell back to local decisions. */
void cdp_verifier_global_stats(const cdp_verifier *v, uint64_t out[2]);

/* ------------------------------------------------------------------ Whisk byte-level API
 * `generate_whisk_shuffle_proof` / `is_valid_whisk_shuffle_proof` (/root/reference/src/whisk.rs:106-179) for a batch of shuffles.
 * A tracker is 96 bytes: r_G || k_r_G, two 48-byte compressed points (src/whisk.rs:38-44); a whisk shuffle proof is
 * compress(M) || CurdleproofsProof::serialize = 48 + cdp_proof_size(ell) bytes (4496 at ell = 124, src/whisk.rs:23).
 * The reference fixes ell = 124 at compile time (src/whisk.rs:28-29); here ell is the prover's / verifier's.
 *
 * generate: for shuffle b, with rng = StdRng::seed_from_u64(rng_seed[b]) advanced by rng_skip_words[b] 32-bit words (NULL = 0):
 *   permutation.shuffle(rng); k = Fr::rand(rng); (vec_R, vec_S) = unzip(pre_trackers);
 *   (vec_T, vec_U, M, m_blinders) = shuffle_permute_and_commit_input(..)  (src/util.rs:83-106);  proof = CurdleproofsProof::new(.., rng)
 * post_trackers_out: batch * ell trackers (vec_T[i] || vec_U[i]); proofs_out: batch * cdp_whisk_shuffle_proof_size(ell) bytes.
 * Returns CDP_ERR_NOT_ON_CURVE when a tracker does not deserialise (the reference's SerializationError). */
size_t cdp_whisk_shuffle_proof_size(size_t ell);
int cdp_whisk_generate_shuffle_proofs(cdp_prover *p, size_t batch, const uint8_t *pre_trackers, const uint64_t *rng_seed,
                                      const uint64_t *rng_skip_words, uint8_t *post_trackers_out, uint8_t *proofs_out);
/* verify: result[b] = 1 valid, 0 invalid, 2 a tracker or the proof does not deserialise (the reference returns Err there). */
int cdp_whisk_verify_shuffle_proofs(cdp_verifier *v, size_t batch, const uint8_t *pre_trackers, const uint8_t *post_trackers,
                                    const uint8_t *proofs, const uint64_t *rng_seed, uint8_t *result);

/* Tracker opening proofs (src/whisk.rs:183-263): knowledge of k with k_r_G = k * r_G and k_commitment = k * G.
 * trackers: batch * 96 bytes; k: batch canonical 32-byte scalars; proofs: batch * 128 bytes (A || B || s, src/whisk.rs:24-25).
 * generate: blinder = Fr::rand(StdRng::seed_from_u64(rng_seed[b]) advanced by rng_skip_words[b] words); verify: result[b] = 1 / 0 / 2
 * (valid / invalid / an encoding does not deserialise). */
int cdp_whisk_generate_tracker_proofs(cdp_ctx *ctx, size_t batch, const uint8_t *trackers, const uint8_t *k, const uint64_t *rng_seed,
                                      const uint64_t *rng_skip_words, uint8_t *proofs_out);
int cdp_whisk_verify_tracker_proofs(cdp_ctx *ctx, size_t batch, const uint8_t *trackers, const uint8_t *k_commitments, const uint8_t *proofs,
                                    uint8_t *result);

/* Host<->device bytes moved by the last cdp_prove_batch call: [0] host-to-device, [1] device-to-host. */
void cdp_prover_last_traffic(const cdp_prover *p, uint64_t out_bytes[2]);

#ifdef __cplusplus
}
#endif
#endif /* CDP_PROVER_H */
