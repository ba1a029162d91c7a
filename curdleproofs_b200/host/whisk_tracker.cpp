// Whisk tracker opening proofs for a batch: `generate_whisk_tracker_proof` / `is_valid_whisk_tracker_proof`
// (/root/reference/src/whisk.rs:183-263): a Chaum-Pedersen proof of the k with k_r_G = k * r_G and k_commitment = k * G.
// Host: rng, transcript, Fr algebra; GPU (through the C ABI): decompression, the scalar multiplications, normalisation.
#include <cstring>
#include <string>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/cdp_prover.h"
#pragma GCC visibility pop
#include "merlin.hpp"
#include "rng.hpp"

using namespace cdp_host;

namespace {
const uint64_t FP_ONE_MONT[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL,
                                 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
// the standard generator, ZCash encoding (pinned by the reference's KAT, src/whisk.rs:363-368)
const uint8_t G1_GEN_COMP[48] = {0x97, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f,
                                 0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58,
                                 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};
void affine_to_jac(uint8_t out[144], const uint8_t aff[96]) {
    bool inf = true;
    for (int i = 0; i < 96; i++) inf = inf && aff[i] == 0;
    memcpy(out, aff, 96);
    if (inf) memset(out + 96, 0, 48);
    else memcpy(out + 96, FP_ONE_MONT, 48);
}
// transcript of both functions: append_list(b"tracker_opening_proof", [k_G, G, k_r_G, r_G, A, B]) then the challenge
Fr tracker_challenge(const uint8_t *k_G, const uint8_t *k_r_G, const uint8_t *r_G, const uint8_t *A, const uint8_t *Bp) {
    Transcript tr("whisk_opening_proof");
    const uint8_t *items[6] = {k_G, G1_GEN_COMP, k_r_G, r_G, A, Bp};
    for (const uint8_t *it : items) tr.append_point("tracker_opening_proof", it);
    return tr.challenge("tracker_opening_proof_challenge");
}
}  // namespace

// trackers: batch * 96 bytes (r_G || k_r_G); k: batch * 32 canonical scalars; out: batch * 128 bytes (A || B || s)
extern "C" int cdp_whisk_generate_tracker_proofs(cdp_ctx *ctx, size_t B, const uint8_t *trackers, const uint8_t *k, const uint64_t *rng_seed,
                                                 const uint64_t *rng_skip_words, uint8_t *proofs_out) {
    if (!ctx || !trackers || !k || !rng_seed || !proofs_out) return CDP_ERR_INVALID_ARG;
    if (B == 0) return CDP_OK;
    // points: [r_G, k_r_G] per tracker, then the generator
    std::vector<uint8_t> comp((2 * B + 1) * 48), aff((2 * B + 1) * 96), status(2 * B + 1);
    memcpy(comp.data(), trackers, 96 * B);
    memcpy(comp.data() + 96 * B, G1_GEN_COMP, 48);
    int rc = cdp_decompress_batch(ctx, comp.data(), 2 * B + 1, aff.data(), status.data());
    if (rc != CDP_OK) return rc;
    const uint8_t *G = aff.data() + 96 * 2 * B;
    std::vector<Fr> kf(B), blinder(B);
    std::vector<uint8_t> pts(3 * B * 96), sc(3 * B * 32), res(3 * B * 96);
    for (size_t b = 0; b < B; b++) {
        if (!Fr::from_bytes(k + 32 * b, kf[b])) return CDP_ERR_INVALID_ARG;
        StdRng rng(rng_seed[b]);
        if (rng_skip_words) rng.skip_words(rng_skip_words[b]);
        blinder[b] = rng.fr_rand();
        // k_G = k * G ; A = blinder * G ; B = blinder * r_G
        memcpy(pts.data() + 96 * (3 * b), G, 96);
        memcpy(pts.data() + 96 * (3 * b + 1), G, 96);
        memcpy(pts.data() + 96 * (3 * b + 2), aff.data() + 96 * (2 * b), 96);
        memcpy(sc.data() + 32 * (3 * b), k + 32 * b, 32);
        blinder[b].to_bytes(sc.data() + 32 * (3 * b + 1));
        blinder[b].to_bytes(sc.data() + 32 * (3 * b + 2));
    }
    rc = cdp_scalar_mul_batch(ctx, pts.data(), sc.data(), 3 * B, res.data());
    if (rc != CDP_OK) return rc;
    std::vector<uint8_t> jac(3 * B * 144), cmp(3 * B * 48);
    for (size_t i = 0; i < 3 * B; i++) affine_to_jac(jac.data() + 144 * i, res.data() + 96 * i);
    rc = cdp_compress_batch(ctx, jac.data(), 3 * B, cmp.data());
    if (rc != CDP_OK) return rc;
    for (size_t b = 0; b < B; b++) {
        const uint8_t *k_G = cmp.data() + 48 * (3 * b), *A = k_G + 48, *Bp = k_G + 96;
        Fr c = tracker_challenge(k_G, trackers + 96 * b + 48, trackers + 96 * b, A, Bp);
        Fr s = blinder[b] - c * kf[b];
        memcpy(proofs_out + 128 * b, A, 48);
        memcpy(proofs_out + 128 * b + 48, Bp, 48);
        s.to_bytes(proofs_out + 128 * b + 96);
    }
    return CDP_OK;
}

// result[b]: 1 valid, 0 invalid, 2 an encoding does not deserialise (the reference returns Err there)
extern "C" int cdp_whisk_verify_tracker_proofs(cdp_ctx *ctx, size_t B, const uint8_t *trackers, const uint8_t *k_commitments,
                                               const uint8_t *proofs, uint8_t *result) {
    if (!ctx || !trackers || !k_commitments || !proofs || !result) return CDP_ERR_INVALID_ARG;
    if (B == 0) return CDP_OK;
    // per proof: r_G, k_r_G, k_G, A, B ; then the generator
    std::vector<uint8_t> comp((5 * B + 1) * 48), aff((5 * B + 1) * 96), status(5 * B + 1);
    for (size_t b = 0; b < B; b++) {
        memcpy(comp.data() + 48 * (5 * b), trackers + 96 * b, 96);
        memcpy(comp.data() + 48 * (5 * b + 2), k_commitments + 48 * b, 48);
        memcpy(comp.data() + 48 * (5 * b + 3), proofs + 128 * b, 96);
    }
    memcpy(comp.data() + 48 * 5 * B, G1_GEN_COMP, 48);
    int rc = cdp_decompress_batch(ctx, comp.data(), 5 * B + 1, aff.data(), status.data());
    if (rc != CDP_OK && rc != CDP_ERR_NOT_ON_CURVE) return rc;
    const uint8_t *G = aff.data() + 96 * 5 * B;
    // A' = s G + c k_G ; B' = s r_G + c k_r_G   as 2-point MSMs
    std::vector<uint8_t> pts(4 * B * 96), sc(4 * B * 32), out(2 * B * 144), outa(2 * B * 96);
    std::vector<cdp_msm_desc> descs(2 * B);
    std::vector<uint8_t> bad(B, 0);
    for (size_t b = 0; b < B; b++) {
        for (int i = 0; i < 5; i++) bad[b] |= status[5 * b + i];
        Fr s;
        if (!Fr::from_bytes(proofs + 128 * b + 96, s)) bad[b] = 1;
        Fr c = tracker_challenge(k_commitments + 48 * b, trackers + 96 * b + 48, trackers + 96 * b, proofs + 128 * b, proofs + 128 * b + 48);
        uint8_t *p = pts.data() + 96 * 4 * b, *q = sc.data() + 32 * 4 * b;
        memcpy(p, G, 96); memcpy(p + 96, aff.data() + 96 * (5 * b + 2), 96);                                       // G, k_G
        memcpy(p + 192, aff.data() + 96 * (5 * b), 96); memcpy(p + 288, aff.data() + 96 * (5 * b + 1), 96);       // r_G, k_r_G
        memcpy(q, proofs + 128 * b + 96, 32); c.to_bytes(q + 32); memcpy(q + 64, proofs + 128 * b + 96, 32); c.to_bytes(q + 96);
        if (bad[b]) memset(q, 0, 128);
        descs[2 * b] = {p, q, 2};
        descs[2 * b + 1] = {p + 192, q + 64, 2};
    }
    rc = cdp_msm_batch(ctx, descs.data(), 2 * B, out.data());
    if (rc != CDP_OK) return rc;
    rc = cdp_normalize_batch(ctx, out.data(), 2 * B, outa.data());
    if (rc != CDP_OK) return rc;
    for (size_t b = 0; b < B; b++) {
        bool ok = memcmp(outa.data() + 96 * (2 * b), aff.data() + 96 * (5 * b + 3), 96) == 0 &&
                  memcmp(outa.data() + 96 * (2 * b + 1), aff.data() + 96 * (5 * b + 4), 96) == 0;
        result[b] = bad[b] ? 2 : (ok ? 1 : 0);
    }
    return CDP_OK;
}
