// Test helper (CPU): runs a fixed script through the PRODUCT's host-side transcript and Fr code
// (curdleproofs_b200/host/merlin.hpp, fr.hpp) and prints every observable byte as hex.  tests/test_host_logic.py
// compares the output with the same script run through the oracle's C implementation (host_script_oracle.c).
#include <cstdio>
#include <vector>

#include "../../curdleproofs_b200/host/merlin.hpp"
using namespace cdp_host;

static void hex(const uint8_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) printf("%02x", b[i]);
    printf("\n");
}
int main() {
    Transcript tr("curdleproofs");
    const size_t lens[] = {0, 1, 47, 48, 165, 166, 167, 331, 332, 333, 1000, 12104, 24576};
    uint32_t x = 12345;
    for (size_t L : lens) {
        std::vector<uint8_t> m(L);
        for (auto &b : m) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
        tr.append_message("curdleproofs_step1", m.data(), L);
        uint8_t out[64];
        tr.challenge_bytes("ch", out, 64);
        hex(out, 64);
        uint8_t big[400];
        tr.challenge_bytes("big", big, 400);  // squeeze across rate boundaries
        hex(big, 400);
    }
    Fr acc = Fr::one();
    for (int i = 0; i < 300; i++) {
        Fr c = tr.challenge("curdleproofs_vec_a");
        acc = acc * c + c;
        uint8_t b[32];
        c.to_bytes(b);
        hex(b, 32);
    }
    uint8_t b[32];
    acc.to_bytes(b); hex(b, 32);
    acc.inverse().to_bytes(b); hex(b, 32);
    (acc - acc * acc).neg().to_bytes(b); hex(b, 32);
    Fr::from_u64(0xdeadbeefULL).to_bytes(b); hex(b, 32);
    return 0;
}
