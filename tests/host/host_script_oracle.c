/* Test helper (CPU): the script of host_script_product.cpp run through the ORACLE's C transcript / Fr code. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/field.h"
#include "../../oracle/merlin.h"

static void hex(const uint8_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) printf("%02x", b[i]);
    printf("\n");
}
static void fr_bytes(uint8_t out[32], const fr_t *a) {
    uint64_t c[4];
    fr_to_canon(c, a);
    memcpy(out, c, 32);
}
/* get_and_append_challenge, /root/reference/src/transcript.rs:41-54 */
static void challenge(transcript_t *t, const char *label, fr_t *out) {
    for (;;) {
        uint8_t buf[64];
        uint64_t c[4];
        transcript_challenge_bytes(t, label, buf, 64);
        buf[31] &= 0x7F;
        memcpy(c, buf, 32);
        int ge = 1;
        for (int i = 3; i >= 0; i--) {
            if (c[i] > FR_R[i]) { ge = 1; break; }
            if (c[i] < FR_R[i]) { ge = 0; break; }
        }
        if (ge || (c[0] | c[1] | c[2] | c[3]) == 0) continue;
        fr_from_canon(out, c);
        transcript_append_message(t, label, buf, 32);
        return;
    }
}
int main(void) {
    transcript_t tr;
    transcript_init(&tr, "curdleproofs");
    const size_t lens[] = {0, 1, 47, 48, 165, 166, 167, 331, 332, 333, 1000, 12104, 24576};
    uint32_t x = 12345;
    for (size_t li = 0; li < sizeof lens / sizeof lens[0]; li++) {
        size_t L = lens[li];
        uint8_t *m = (uint8_t *)malloc(L + 1);
        for (size_t i = 0; i < L; i++) { x = x * 1664525u + 1013904223u; m[i] = (uint8_t)(x >> 24); }
        transcript_append_message(&tr, "curdleproofs_step1", m, L);
        free(m);
        uint8_t out[64], big[400];
        transcript_challenge_bytes(&tr, "ch", out, 64);
        hex(out, 64);
        transcript_challenge_bytes(&tr, "big", big, 400);
        hex(big, 400);
    }
    fr_t acc;
    fr_from_u64(&acc, 1);
    uint8_t b[32];
    for (int i = 0; i < 300; i++) {
        fr_t c;
        challenge(&tr, "curdleproofs_vec_a", &c);
        fr_mul(&acc, &acc, &c);
        fr_add(&acc, &acc, &c);
        fr_bytes(b, &c); hex(b, 32);
    }
    fr_bytes(b, &acc); hex(b, 32);
    fr_t t, u;
    fr_inv(&t, &acc); fr_bytes(b, &t); hex(b, 32);
    fr_mul(&t, &acc, &acc); fr_sub(&t, &acc, &t); fr_neg(&u, &t); fr_bytes(b, &u); hex(b, 32);
    uint64_t c[4] = {0xdeadbeefULL, 0, 0, 0};
    fr_from_canon(&t, c); fr_bytes(b, &t); hex(b, 32);
    return 0;
}
