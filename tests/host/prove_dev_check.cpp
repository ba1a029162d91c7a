// Test helper (CPU): the batched prover's host driver (curdleproofs_b200/host/prover.cpp) and the device-side prover code it drives
// (csrc/k_prove.cu, k_transcript.cu, the round expansion of k_vcoeffs.cu -- compiled as plain C++ by tests/host/cpu_engine_mock.cpp, with
// the CPU oracle standing in for the group kernels) against the oracle's restatement of `CurdleproofsProof::new`
// (/root/reference/src/curdleproofs.rs:59-184): whole proofs, byte for byte.
//   prove_dev_check random <ell> <batch> <lanes>   random instances, seeds 1.., both rng forms; device path AND the host-transcript path
//   prove_dev_check golden                        the reference's seed-0 whisk shuffle (ell = 124, src/whisk.rs:416-456): the proof part of
//                                                 the 4496-byte golden vector (src/whisk.rs:455), which tests/test_oracle_golden.py pins
//   prove_dev_check inverse                       the device code's single-thread Euclidean Fr inversion against the Fermat ladder
//   prove_dev_check badinput                      witness validation: out-of-range / repeated permutation entries, non-canonical scalars
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cdp_prover.h"

extern "C" {
int oracle_generate_crs_points(size_t ell, uint8_t *out_pts);
int oracle_random_instance(size_t ell, const uint8_t *crs_pts, uint64_t seed, int fast_points, uint8_t *vec_R, uint8_t *vec_S, uint8_t *vec_T,
                           uint8_t *vec_U, uint8_t M_jac[144], uint32_t *perm, uint8_t k_out[32], uint8_t *m_blinders, int threads);
int oracle_prove(size_t ell, const uint8_t *crs_pts, const uint8_t *vec_R, const uint8_t *vec_S, const uint8_t *vec_T, const uint8_t *vec_U,
                 const uint8_t M_jac[144], const uint32_t *perm, const uint8_t k_bytes[32], const uint8_t *m_blinders, uint64_t rng_seed,
                 uint8_t *proof_out, int threads);
int oracle_whisk_shuffle_proof_seed0(size_t ell, uint8_t *proof_out, uint8_t *inst_out, int *verified, int threads);
void oracle_stdrng_seed_bytes(uint64_t seed, uint8_t out[32]);
int mock_check_fr_inverse(int count);  // tests/host/cpu_engine_mock.cpp
}

struct Batch {
    size_t ell, B;
    std::vector<uint8_t> crs, R, S, T, U, M, k, mb;
    std::vector<uint32_t> perm;
    std::vector<uint64_t> seeds;
};

static void make_batch(Batch &b, size_t ell, size_t B) {
    b.ell = ell; b.B = B;
    b.crs.resize((ell + 7) * 96);
    oracle_generate_crs_points(ell, b.crs.data());
    b.R.resize(B * ell * 96); b.S.resize(B * ell * 96); b.T.resize(B * ell * 96); b.U.resize(B * ell * 96);
    b.M.resize(B * 144); b.k.resize(B * 32); b.mb.resize(B * 128); b.perm.resize(B * ell); b.seeds.resize(B);
    for (size_t i = 0; i < B; i++) {
        oracle_random_instance(ell, b.crs.data(), 100 + i, 1, &b.R[i * ell * 96], &b.S[i * ell * 96], &b.T[i * ell * 96], &b.U[i * ell * 96], &b.M[i * 144],
                               &b.perm[i * ell], &b.k[i * 32], &b.mb[i * 128], 1);
        b.seeds[i] = 1 + i;
    }
}

static int prove(const Batch &b, int lanes, bool host_transcript, bool use_keys, std::vector<uint8_t> &out, const uint64_t *skip = nullptr) {
    setenv("CDP_PROVE_HOST_TRANSCRIPT", host_transcript ? "1" : "0", 1);
    cdp_ctx *ctx = nullptr;
    cdp_ctx_create(&ctx, 0, nullptr);
    cdp_prover *p = nullptr;
    int rc = cdp_prover_create_lanes(&p, ctx, b.ell, b.crs.data(), b.B, 2, lanes);
    if (rc) { printf("create failed %d\n", rc); return rc; }
    std::vector<uint8_t> keys(32 * b.B);
    for (size_t i = 0; i < b.B; i++) oracle_stdrng_seed_bytes(b.seeds[i], &keys[32 * i]);  // seed_from_u64's key: the same stream through rng_key
    cdp_prove_inputs in;
    memset(&in, 0, sizeof in);
    in.vec_R = b.R.data(); in.vec_S = b.S.data(); in.vec_T = b.T.data(); in.vec_U = b.U.data(); in.M = b.M.data();
    in.permutation = b.perm.data(); in.k = b.k.data(); in.vec_m_blinders = b.mb.data();
    in.rng_seed = use_keys ? nullptr : b.seeds.data(); in.rng_key = use_keys ? keys.data() : nullptr; in.rng_skip_words = skip;
    out.assign(b.B * cdp_proof_size(b.ell), 0);
    rc = cdp_prove_batch(p, b.B, &in, out.data());
    if (rc) printf("cdp_prove_batch failed %d: %s\n", rc, cdp_prover_last_error(p));
    cdp_prover_destroy(p);
    cdp_ctx_destroy(ctx);
    return rc;
}

static int first_diff(const uint8_t *a, const uint8_t *b, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (a[i] != b[i]) return (int)i;
    return -1;
}

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "random";
    if (!strcmp(mode, "random")) {
        const size_t ell = argc > 2 ? atoi(argv[2]) : 4, B = argc > 3 ? atoi(argv[3]) : 3;
        const int lanes = argc > 4 ? atoi(argv[4]) : 1;
        Batch b;
        make_batch(b, ell, B);
        const size_t psz = cdp_proof_size(ell);
        std::vector<uint8_t> want(B * psz), dev, host, devk;
        for (size_t i = 0; i < B; i++)
            oracle_prove(ell, b.crs.data(), &b.R[i * ell * 96], &b.S[i * ell * 96], &b.T[i * ell * 96], &b.U[i * ell * 96], &b.M[i * 144], &b.perm[i * ell],
                         &b.k[i * 32], &b.mb[i * 128], b.seeds[i], &want[i * psz], 1);
        if (prove(b, lanes, false, false, dev) || prove(b, lanes, true, false, host) || prove(b, lanes, false, true, devk)) return 1;
        int bad = 0;
        for (size_t i = 0; i < B; i++) {
            const int d1 = first_diff(&dev[i * psz], &want[i * psz], psz), d2 = first_diff(&host[i * psz], &want[i * psz], psz),
                      d3 = first_diff(&devk[i * psz], &want[i * psz], psz);
            if (d1 >= 0 || d2 >= 0 || d3 >= 0) {
                printf("MISMATCH ell=%zu proof %zu: device path first diff at byte %d, host path %d, device path with rng_key %d (of %zu)\n", ell, i, d1, d2, d3, psz);
                bad = 1;
            }
        }
        if (!bad) printf("ell=%zu batch=%zu lanes=%d ok : device-side prover, host-transcript prover and the oracle agree on every proof byte\n", ell, B, lanes);
        return bad;
    }
    if (!strcmp(mode, "golden")) {
        const size_t ell = 124, psz = cdp_proof_size(ell);
        std::vector<uint8_t> gold(48 + psz), inst(384 * ell + 144 + 4 * ell + 32 + 128 + 16);
        int verified = 0;
        oracle_whisk_shuffle_proof_seed0(ell, gold.data(), inst.data(), &verified, 1);
        Batch b;
        b.ell = ell; b.B = 1;
        b.crs.resize((ell + 7) * 96);
        oracle_generate_crs_points(ell, b.crs.data());
        const uint8_t *w = inst.data();
        b.R.assign(w, w + 96 * ell); b.S.assign(w + 96 * ell, w + 192 * ell); b.T.assign(w + 192 * ell, w + 288 * ell); b.U.assign(w + 288 * ell, w + 384 * ell);
        b.M.assign(w + 384 * ell, w + 384 * ell + 144);
        w += 384 * ell + 144;
        b.perm.resize(ell);
        memcpy(b.perm.data(), w, 4 * ell); w += 4 * ell;
        b.k.assign(w, w + 32); w += 32;
        b.mb.assign(w, w + 128); w += 128;
        uint64_t skip;
        memcpy(&skip, w, 8);
        b.seeds.assign(1, 0);
        std::vector<uint8_t> dev;
        if (prove(b, 1, false, false, dev, &skip)) return 1;
        const int d = first_diff(dev.data(), gold.data() + 48, psz);
        if (d >= 0 || !verified) { printf("MISMATCH golden: first diff at byte %d of %zu (oracle verified: %d)\n", d, psz, verified); return 1; }
        printf("golden ok : the proof part of the seed-0 whisk shuffle proof (ell = 124, %zu bytes) through the device-side prover\n", psz);
        return 0;
    }
    if (!strcmp(mode, "badinput")) {
        Batch b;
        make_batch(b, 4, 2);
        std::vector<uint8_t> out;
        int bad = 0;
        auto expect_reject = [&](const char *what) {
            const int rc = prove(b, 1, false, false, out);
            if (rc != CDP_ERR_INVALID_ARG) { printf("MISMATCH badinput: %s accepted (rc %d)\n", what, rc); bad = 1; }
        };
        Batch good = b;
        b.perm[5] = 4; expect_reject("out-of-range permutation entry"); b = good;
        b.perm[1] = b.perm[0]; expect_reject("repeated permutation entry"); b = good;
        memset(&b.k[32], 0xFF, 32); expect_reject("non-canonical k"); b = good;
        memset(&b.mb[128 + 64], 0xFF, 32); expect_reject("non-canonical blinder"); b = good;
        if (prove(b, 1, false, false, out) != CDP_OK) { printf("MISMATCH badinput: the untouched batch was rejected\n"); bad = 1; }
        if (!bad) printf("badinput ok : malformed witnesses are refused with CDP_ERR_INVALID_ARG before any work\n");
        return bad;
    }
    if (!strcmp(mode, "inverse")) {
        const int bad = mock_check_fr_inverse(2000);
        printf(bad ? "MISMATCH inverse: %d values\n" : "inverse ok : Euclidean and Fermat inversions agree on edge values and 2000 random ones (%d mismatches)\n", bad);
        return bad != 0;
    }
    printf("unknown mode %s\n", mode);
    return 2;
}
