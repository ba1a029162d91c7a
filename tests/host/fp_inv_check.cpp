// Test helper (CPU): the Fp inversions of the device code (curdleproofs_b200/csrc/fp_inv_euclid.cuh: binary Euclid; fp_inv_safegcd.cuh:
// Bernstein-Yang division steps; both plain C++) against each other and against the
// oracle's field: for edge values and pseudo-random ones, inverse_int(a) * a == 1 (mod p) and inverse_int(a) < p.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../curdleproofs_b200/csrc/fp_inv_euclid.cuh"
#include "../../curdleproofs_b200/csrc/fp_inv_safegcd.cuh"
extern "C" {
void oracle_fp_from_canon(const uint8_t in[48], uint8_t out[48]);
void oracle_fp_to_canon(const uint8_t in[48], uint8_t out[48]);
void oracle_fp_mul(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]);
}
static const uint32_t P[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                               0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
static bool less_than_p(const uint32_t *a) {
    for (int i = 11; i >= 0; i--) { if (a[i] < P[i]) return true; if (a[i] > P[i]) return false; }
    return false;
}
int main() {
    uint64_t s = 0x9E3779B97F4A7C15ULL;
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    int bad = 0, n = 0;
    for (int t = 0; t < 20000; t++) {
        uint32_t a[12] = {0};
        if (t == 0) a[0] = 1;
        else if (t == 1) a[0] = 2;
        else if (t == 2) { memcpy(a, P, 48); a[0] -= 1; }                 // p - 1
        else if (t == 3) { memcpy(a, P, 48); a[0] -= 2; }
        else if (t == 4) a[11] = 0x10000000u;                             // a power of two
        else if (t == 5) { /* zero */ }
        else if (t == 6) { a[0] = 0xffffffffu; a[1] = 0x3fffffffu; }
        else if (t < 40) a[(t - 7) / 3] = 1u << (10 * ((t - 7) % 3) + 1);           // small powers of two across the limbs
        else { for (int i = 0; i < 12; i++) a[i] = next(); a[11] &= 0x0fffffffu; if (t % 7 == 0) a[0] &= ~0xffu; if (t % 11 == 0) for (int i = t % 12; i < 12; i++) a[i] = 0; }
        if (!less_than_p(a)) continue;
        uint32_t inv[12];
        cdp::euclid::inverse_int(inv, a, [](bool d) { return d; });
        n++;
        uint32_t inv2[12];
        cdp::safegcd::inverse_int(inv2, a, [](bool d) { return d; });
        bool ok = less_than_p(inv) && memcmp(inv, inv2, 48) == 0;
        bool zero = true;
        for (int i = 0; i < 12; i++) zero = zero && a[i] == 0;
        if (zero) {
            for (int i = 0; i < 12; i++) ok = ok && inv[i] == 0;
        } else {
            uint8_t am[48], im[48], pm[48], pc[48];
            oracle_fp_from_canon((const uint8_t *)a, am);
            oracle_fp_from_canon((const uint8_t *)inv, im);
            oracle_fp_mul(am, im, pm);
            oracle_fp_to_canon(pm, pc);
            uint32_t one[12] = {1};
            ok = ok && memcmp(pc, one, 48) == 0;
        }
        if (!ok && bad < 8) { printf("t=%d a=", t); for (int i = 11; i >= 0; i--) printf("%08x", a[i]); printf("\n"); }
        bad += !ok;
    }
    printf(bad ? "MISMATCH fp inverse: %d of %d\n" : "fp inverse ok : %d of %d values (bad %d)\n", bad ? bad : n, n, bad);
    return bad != 0;
}
