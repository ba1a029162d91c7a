// Test helper (CPU): the Barrett GLV split of the device code (curdleproofs_b200/csrc/glv_split.cuh, plain C++) against schoolbook long
// division: k = k2 * lambda + k1 with 0 <= k1 < lambda for edge values and pseudo-random scalars below r.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../curdleproofs_b200/csrc/glv_split.cuh"
typedef unsigned __int128 u128;
static const uint32_t R[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
static const u128 LAM = ((u128)0xac45a4010001a402ULL << 64) | 0x00000000ffffffffULL;
static bool below_r(const uint32_t *k) {
    for (int i = 7; i >= 0; i--) { if (k[i] < R[i]) return true; if (k[i] > R[i]) return false; }
    return false;
}
int main() {
    uint64_t s = 0x243F6A8885A308D3ULL;
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    int bad = 0, n = 0;
    for (int t = 0; t < 200000; t++) {
        uint32_t k[8] = {0};
        if (t == 1) k[0] = 1;
        else if (t == 2) { memcpy(k, R, 32); k[0] -= 1; }                                  // r - 1
        else if (t == 3) { k[0] = 0xffffffffu; k[2] = 0x0001a402u; k[3] = 0xac45a401u; }    // lambda
        else if (t == 4) { k[0] = 0xfffffffeu; k[2] = 0x0001a402u; k[3] = 0xac45a401u; }    // lambda - 1
        else if (t == 5) { k[1] = 1; k[2] = 0x0001a402u; k[3] = 0xac45a401u; }              // lambda + 1
        else if (t >= 6) { for (int i = 0; i < 8; i++) k[i] = next(); k[7] &= 0x7fffffffu; if (t % 5 == 0) for (int i = t % 8; i < 8; i++) k[i] = 0; }
        if (!below_r(k)) continue;
        n++;
        cdp::glv_t g;
        cdp::glv_split(g, k);
        // k2 * lambda + k1 == k (256-bit check by limbs) and k1 < lambda
        u128 k1 = 0, k2 = 0;
        for (int i = 3; i >= 0; i--) { k1 = (k1 << 32) | g.k1[i]; k2 = (k2 << 32) | g.k2[i]; }
        // product k2 * LAM as 256 bits
        uint64_t a[2] = {(uint64_t)k2, (uint64_t)(k2 >> 64)}, b[2] = {(uint64_t)LAM, (uint64_t)(LAM >> 64)};
        uint64_t p[4] = {0, 0, 0, 0};
        for (int i = 0; i < 2; i++) {
            u128 carry = 0;
            for (int j = 0; j < 2; j++) {
                u128 v = (u128)a[i] * b[j] + p[i + j] + carry;
                p[i + j] = (uint64_t)v;
                carry = v >> 64;
            }
            p[i + 2] += (uint64_t)carry;
        }
        u128 c = (u128)p[0] + (uint64_t)k1;
        p[0] = (uint64_t)c;
        c = (u128)p[1] + (uint64_t)(k1 >> 64) + (uint64_t)(c >> 64);
        p[1] = (uint64_t)c;
        c = (u128)p[2] + (uint64_t)(c >> 64);
        p[2] = (uint64_t)c;
        p[3] += (uint64_t)(c >> 64);
        bool ok = k1 < LAM;
        for (int i = 0; i < 4; i++) ok = ok && p[i] == (((uint64_t)k[2 * i + 1] << 32) | k[2 * i]);
        bad += !ok;
    }
    printf(bad ? "MISMATCH glv split: %d of %d\n" : "glv split ok : %d of %d values (bad %d)\n", bad ? bad : n, n, bad);
    return bad != 0;
}
