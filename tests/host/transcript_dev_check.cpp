// Test helper (CPU): compiles the DEVICE transcript-opening code (curdleproofs_b200/csrc/k_transcript.cu) as plain C++ and runs it
// next to the product's host transcript (merlin.hpp) on the same encodings: vec_a and the continuation of the transcript must agree.
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

static void hex(const uint8_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) printf("%02x", b[i]);
    printf("\n");
}
// `dump ell B`: prints vec_a and the 208-byte states of the DEVICE code (run on the CPU) for the LCG inputs tests/test_gpu_transcript.py
// also generates, so the GPU run of the same code can be compared byte for byte
static int dump(uint32_t ell, uint32_t B) {
    std::vector<uint8_t> buf(B * 4 * ell * 48), M(B * 48), va(B * ell * 32);
    std::vector<uint64_t> st(B * 26);
    uint32_t x = ell * 7919u + B;
    for (auto &b : buf) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
    for (auto &b : M) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
    for (uint32_t pr = 0; pr < B; pr++) {
        blockIdx.x = pr;
        cdp::k_transcript_open(buf.data(), M.data(), ell, B, va.data(), st.data());
    }
    hex(va.data(), va.size());
    hex(reinterpret_cast<const uint8_t *>(st.data()), st.size() * 8);
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 4 && !strcmp(argv[1], "dump")) return dump((uint32_t)atoi(argv[2]), (uint32_t)atoi(argv[3]));
    int bad = 0;
    for (uint32_t ell : {4u, 5u, 12u, 28u, 124u, 252u}) {
        const uint32_t B = 3;
        std::vector<uint8_t> buf(B * 4 * ell * 48 + 7), M(B * 48 + 3), va(B * ell * 32);
        std::vector<uint64_t> st(B * 26);
        uint32_t x = ell;
        for (auto &b : buf) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
        for (auto &b : M) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
        for (uint32_t off : {0u, 3u}) {  // aligned and misaligned sources
            const uint8_t *vecs = buf.data() + off, *Mp = M.data() + (off ? 1 : 0);
            for (uint32_t pr = 0; pr < B; pr++) {
                blockIdx.x = pr;
                cdp::k_transcript_open(vecs, Mp, ell, B, va.data(), st.data());
                Transcript tr("curdleproofs");
                for (int v = 0; v < 4; v++) tr.append_point_vec("curdleproofs_step1", vecs + ((size_t)pr * 4 + v) * ell * 48, ell);
                tr.append_point("curdleproofs_step1", Mp + pr * 48);
                for (uint32_t i = 0; i < ell; i++) {
                    Fr c = tr.challenge("curdleproofs_vec_a");
                    uint8_t b[32];
                    c.to_bytes(b);
                    if (memcmp(b, va.data() + ((size_t)pr * ell + i) * 32, 32)) { bad++; break; }
                }
                Transcript tr2(st.data() + pr * 26);
                uint8_t o1[64], o2[64];
                tr.challenge_bytes("next", o1, 64);
                tr2.challenge_bytes("next", o2, 64);
                if (memcmp(o1, o2, 64)) bad++;
            }
        }
        printf("ell=%u %s\n", ell, bad ? "MISMATCH" : "ok");
    }
    return bad ? 1 : 0;
}
