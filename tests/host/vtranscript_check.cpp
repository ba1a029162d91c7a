// Test helper (CPU): compiles the DEVICE verifier-transcript kernels (k_verify_transcript_a / _b in curdleproofs_b200/csrc/k_transcript.cu)
// as plain C++ and runs them next to the product's host transcript (host/merlin.hpp + host/fr.hpp) following
// `CurdleproofsProof::verify` (/root/reference/src/curdleproofs.rs:226-296): every challenge, the inverted round challenges, z, the six
// stage scalars and the final STROBE state must agree.  Inputs are random bytes (the transcript does not care whether an encoding is a
// curve point); ell = 12 and 252.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CDP_TRANSCRIPT_HOST_HARNESS
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
struct dim3_t { unsigned x; };
static dim3_t blockIdx{0}, blockDim{1}, threadIdx{0};
namespace cdp {
static const uint32_t FR_R[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
}
#include "../../curdleproofs_b200/csrc/k_transcript.cu"
#include "../../curdleproofs_b200/host/merlin.hpp"

using namespace cdp_host;

static uint64_t rs = 0x9E3779B97F4A7C15ULL;
static uint64_t next64() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
static void rand_bytes(uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) p[i] = (uint8_t)(next64() >> 32); }
static void rand_canonical_fr(uint8_t *p) {
    Fr x;
    do { rand_bytes(p, 32); p[31] &= 0x3F; } while (!Fr::from_bytes(p, x));
}
static void batch_inverse(Fr *xs, size_t n) {
    std::vector<Fr> pre(n);
    Fr acc = Fr::one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc *= xs[i]; }
    Fr inv = acc.inverse();
    for (size_t i = n; i-- > 0;) { Fr t = inv * pre[i]; inv *= xs[i]; xs[i] = t; }
}

static int run(uint32_t ell) {
    const uint32_t n = ell + 4;
    uint32_t m = 0;
    while ((1u << m) < n) m++;
    const uint32_t LC = 11, RC = LC + m, LD = RC + m, RD = LD + m, A1 = RD + m, A2 = A1 + 1, B1 = A2 + 1, B2 = B1 + 1, Ba = B2 + 1, Bt = Ba + 1,
                   Bu = Bt + 1, LA = Bu + 1, LT = LA + m, LU = LT + m, RA = LU + m, RT = RA + m, RU = RT + m, np = RU + m;
    const uint32_t vch = 27 + 4 * m, B = 3;
    std::vector<uint8_t> pcomp(B * np * 48), pscal(B * 7 * 32), compM(B * 48), va(B * ell * 32), da(B * 2 * 48), vecs(B * 4 * ell * 48), H(48);
    rand_bytes(pcomp.data(), pcomp.size()); rand_bytes(compM.data(), compM.size()); rand_bytes(da.data(), da.size());
    rand_bytes(vecs.data(), vecs.size()); rand_bytes(H.data(), 48);
    for (size_t i = 0; i < B * 7; i++) rand_canonical_fr(&pscal[32 * i]);
    for (size_t i = 0; i < B * ell; i++) rand_canonical_fr(&va[32 * i]);
    // starting states: a host transcript after a few messages (so that pos / pos_begin are not trivial)
    std::vector<uint64_t> st0(B * 26), st(B * 26);
    std::vector<Transcript> host;
    for (uint32_t pr = 0; pr < B; pr++) {
        Transcript t("curdleproofs");
        t.append_message("curdleproofs_step1", pcomp.data(), 100 + 37 * pr);
        host.push_back(t);
    }
    // the host transcript cannot export its state, so the device side replays the same prefix with its own primitives
    for (uint32_t pr = 0; pr < B; pr++) {
        cdp::strobe_t s;
        for (int i = 0; i < 25; i++) s.st[i] = 0;
        const uint8_t init[6] = {1, (uint8_t)(cdp::STROBE_R + 2), 1, 0, 1, 96};
        for (uint32_t i = 0; i < 6; i++) cdp::st_xor_byte(s, i, init[i]);
        for (uint32_t i = 0; i < 12; i++) cdp::st_xor_byte(s, 6 + i, cdp::L_STROBE[i]);
        cdp::keccak_f1600(s.st);
        s.pos = 0; s.pos_begin = 0;
        cdp::strobe_meta_ad(s, cdp::L_MERLIN, 11, false);
        cdp::merlin_append(s, cdp::L_DOMSEP, 7, 0, 0, cdp::L_PROTO, 12);
        cdp::merlin_append(s, cdp::L_STEP1, 18, 0, 0, pcomp.data(), 100 + 37 * pr);
        cdp::state_store(&st0[pr * 26], s);
    }
    st = st0;
    std::vector<uint32_t> chal(B * vch * 8, 0xABABABABu), tmp(B * 2 * 8), stage(B * 6 * 8);
    std::vector<uint8_t> flags(B, 9);
    vecs[((1 * 4 + 2) * ell) * 48] |= 0x40; vecs[((0 * 4 + 2) * ell) * 48] &= ~0x40; vecs[((2 * 4 + 2) * ell) * 48] &= ~0x40;  // only proof 1 has T[0] = identity
    for (uint32_t pr = 0; pr < B; pr++) {
        blockIdx.x = pr;
        cdp::k_verify_transcript_a(pcomp.data(), pscal.data(), compM.data(), va.data(), ell, np, vch, B, st.data(), chal.data(), tmp.data(), stage.data(), vecs.data(), flags.data());
    }
    for (uint32_t pr = 0; pr < B; pr++) {
        blockIdx.x = pr;
        cdp::k_verify_transcript_b(pcomp.data(), pscal.data(), da.data(), vecs.data(), H.data(), ell, m, np, vch, B, st.data(), chal.data(), tmp.data());
    }
    // ---- host reference
    int bad = 0;
    for (uint32_t pr = 0; pr < B; pr++) {
        Transcript &tr = host[pr];
        const uint8_t *pc = &pcomp[pr * np * 48];
        auto PS = [&](int k) { Fr x; Fr::from_bytes(&pscal[(pr * 7 + k) * 32], x); return x; };
        std::vector<Fr> a(ell);
        for (uint32_t i = 0; i < ell; i++) Fr::from_bytes(&va[(pr * ell + i) * 32], a[i]);
        tr.append_point("same_perm_step1", pc);
        tr.append_point("same_perm_step1", &compM[pr * 48]);
        tr.append_fr_vec("same_perm_step1", a.data(), ell);
        Fr alpha_sp = tr.challenge("same_perm_alpha"), beta_sp = tr.challenge("same_perm_beta");
        Fr gprod = Fr::one();
        for (uint32_t i = 0; i < ell; i++) gprod *= a[i] + Fr::from_u64(i) * alpha_sp + beta_sp;
        tr.append_point("gprod_step1", pc + 48 * 7);
        tr.append_fr("gprod_step1", gprod);
        Fr alpha_g = tr.challenge("gprod_alpha");
        tr.append_point("gprod_step2", pc + 48 * 8);
        tr.append_fr("gprod_step2", PS(0));
        Fr beta = tr.challenge("gprod_beta"), beta_inv = beta.inverse();
        Fr beta_l = beta.pow_u64(ell);
        Fr z = PS(0) * beta_l * beta + gprod * beta_l - Fr::one();
        tr.append_point("ipa_step1", pc + 48 * 8);
        tr.append_point("ipa_step1", &da[(pr * 2) * 48]);
        tr.append_fr("ipa_step1", z);
        tr.append_point("ipa_step1", pc + 48 * 9);
        tr.append_point("ipa_step1", pc + 48 * 10);
        Fr alpha_i = tr.challenge("ipa_alpha"), beta_i = tr.challenge("ipa_beta");
        std::vector<Fr> gam(m), gam_inv, gam2(m), gam2_inv;
        for (uint32_t k = 0; k < m; k++) {
            tr.append_point("ipa_loop", pc + 48 * (LC + k)); tr.append_point("ipa_loop", pc + 48 * (LD + k));
            tr.append_point("ipa_loop", pc + 48 * (RC + k)); tr.append_point("ipa_loop", pc + 48 * (RD + k));
            gam[k] = tr.challenge("ipa_gamma");
        }
        gam_inv = gam; batch_inverse(gam_inv.data(), m);
        const uint32_t ss[10] = {5, 6, 1, 2, 3, 4, A1, A2, B1, B2};
        for (int q = 0; q < 10; q++) tr.append_point("sameexp_points", pc + 48 * ss[q]);
        Fr alpha_ss = tr.challenge("same_scalar_alpha");
        tr.append_point("same_msm_step1", &da[(pr * 2 + 1) * 48]);
        tr.append_point("same_msm_step1", pc + 48 * 2);
        tr.append_point("same_msm_step1", pc + 48 * 4);
        std::vector<uint8_t> tu(2 * n * 48);
        uint8_t inf[48] = {0xC0};
        memcpy(&tu[0], &vecs[((pr * 4 + 2) * ell) * 48], ell * 48);
        memcpy(&tu[ell * 48], inf, 48); memcpy(&tu[(ell + 1) * 48], inf, 48); memcpy(&tu[(ell + 2) * 48], H.data(), 48); memcpy(&tu[(ell + 3) * 48], inf, 48);
        memcpy(&tu[n * 48], &vecs[((pr * 4 + 3) * ell) * 48], ell * 48);
        memcpy(&tu[(n + ell) * 48], inf, 48); memcpy(&tu[(n + ell + 1) * 48], inf, 48); memcpy(&tu[(n + ell + 2) * 48], inf, 48); memcpy(&tu[(n + ell + 3) * 48], H.data(), 48);
        tr.append_point_vec("same_msm_step1", &tu[0], n);
        tr.append_point_vec("same_msm_step1", &tu[n * 48], n);
        tr.append_point("same_msm_step1", pc + 48 * Ba);
        tr.append_point("same_msm_step1", pc + 48 * Bt);
        tr.append_point("same_msm_step1", pc + 48 * Bu);
        Fr alpha_sm = tr.challenge("same_msm_alpha");
        for (uint32_t k = 0; k < m; k++) {
            const uint32_t o[6] = {LA, LT, LU, RA, RT, RU};
            for (int q = 0; q < 6; q++) tr.append_point("same_msm_loop", pc + 48 * (o[q] + k));
            gam2[k] = tr.challenge("same_msm_gamma");
        }
        gam2_inv = gam2; batch_inverse(gam2_inv.data(), m);
        // expected block (entries 0..11 stay untouched)
        std::vector<Fr> want(vch);
        want[12] = alpha_sp; want[13] = beta_sp; want[14] = alpha_g; want[15] = beta_inv; want[16] = alpha_i; want[17] = beta_i; want[18] = z;
        want[19] = PS(1); want[20] = PS(2); want[21] = PS(6); want[22] = alpha_sm; want[23] = alpha_ss; want[24] = PS(3); want[25] = PS(4); want[26] = PS(5);
        for (uint32_t k = 0; k < m; k++) { want[27 + k] = gam[k]; want[27 + m + k] = gam_inv[k]; want[27 + 2 * m + k] = gam2[k]; want[27 + 3 * m + k] = gam2_inv[k]; }
        for (uint32_t k = 0; k < vch; k++) {
            const uint32_t *got = &chal[(pr * vch + k) * 8];
            if (k < 12) { for (int q = 0; q < 8; q++) if (got[q] != 0xABABABABu) { printf("ell=%u proof %u: entry %u was written\n", ell, pr, k); bad = 1; } }
            else if (memcmp(got, want[k].v, 32)) { printf("ell=%u proof %u: challenge entry %u differs\n", ell, pr, k); bad = 1; }
        }
        uint8_t sc[6 * 32];
        const Fr one = Fr::one();
        one.to_bytes(sc); beta_inv.neg().to_bytes(sc + 32); alpha_g.to_bytes(sc + 64); one.to_bytes(sc + 96); one.to_bytes(sc + 128); one.to_bytes(sc + 160);
        if (memcmp(sc, &stage[pr * 6 * 8], 6 * 32)) { printf("ell=%u proof %u: stage scalars differ\n", ell, pr); bad = 1; }
        // the transcripts must still be in step: one more challenge from both
        uint8_t hb[32], db[32];
        tr.challenge_bytes("tail", hb, 32);
        cdp::strobe_t s;
        cdp::state_load(s, &st[pr * 26]);
        const uint8_t lbl[] = "tail";
        cdp::merlin_challenge(s, lbl, 4, db, 32);
        if (memcmp(hb, db, 32)) { printf("ell=%u proof %u: final transcript state differs\n", ell, pr); bad = 1; }
    }
    if (flags[0] != 0 || flags[1] != 1 || flags[2] != 0) { printf("ell=%u: identity flags wrong\n", ell); bad = 1; }
    if (!bad) printf("ell=%u ok (%u challenge scalars per proof)\n", ell, vch - 12);
    return bad;
}

int main() { return run(12) | run(252) | run(124); }
