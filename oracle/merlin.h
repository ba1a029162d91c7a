/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Fiat-Shamir transcript: Keccak-f[1600] -> STROBE-128 (the subset merlin uses) -> merlin
 * `Transcript`.  The reference depends on `merlin = "3.0.0"` (`/root/reference/Cargo.toml:22`;
 * not vendored under /root/reference), used at `src/transcript.rs:28-61`,
 * `src/curdleproofs.rs:78,213`.  This restates the published construction:
 *   STROBE-128/1600: rate R = 166, ops AD / meta-AD / PRF (KEY unused here);
 *   Transcript::new(label)        = Strobe128::new("Merlin v1.0"); append_message("dom-sep", label)
 *   append_message(label, msg)    = meta_AD(label) ; meta_AD(le32(len), more) ; AD(msg)
 *   challenge_bytes(label, out)   = meta_AD(label) ; meta_AD(le32(len), more) ; PRF(out)
 * Pinned by merlin's own published test vector (tests/test_oracle_kat.py) and by the two
 * golden proofs of the reference (src/whisk.rs:401,455), which exercise every label.
 */
#ifndef CDP_ORACLE_MERLIN_H
#define CDP_ORACLE_MERLIN_H
#include <stdint.h>
#include <string.h>

static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static void keccak_f1600(uint64_t st[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63));
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = KECCAK_PIL[i];
            uint64_t b = st[j];
            st[j] = (t << KECCAK_ROT[i]) | (t >> (64 - KECCAK_ROT[i]));
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KECCAK_RC[round];
    }
}

#define STROBE_R 166
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };

typedef struct {
    union { uint64_t w[25]; uint8_t b[200]; } st; /* little-endian host assumed */
    uint8_t pos, pos_begin, cur_flags;
} strobe_t;

static void strobe_run_f(strobe_t *s) {
    s->st.b[s->pos] ^= s->pos_begin;
    s->st.b[s->pos + 1] ^= 0x04;
    s->st.b[STROBE_R + 1] ^= 0x80;
    keccak_f1600(s->st.w);
    s->pos = 0; s->pos_begin = 0;
}
static void strobe_absorb(strobe_t *s, const uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        s->st.b[s->pos++] ^= d[i];
        if (s->pos == STROBE_R) strobe_run_f(s);
    }
}
static void strobe_squeeze(strobe_t *s, uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        d[i] = s->st.b[s->pos]; s->st.b[s->pos] = 0; s->pos++;
        if (s->pos == STROBE_R) strobe_run_f(s);
    }
}
static void strobe_begin_op(strobe_t *s, uint8_t flags, int more) {
    if (more) return; /* caller guarantees same flags */
    uint8_t old_begin = s->pos_begin;
    s->pos_begin = s->pos + 1;
    s->cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    strobe_absorb(s, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && s->pos != 0) strobe_run_f(s);
}
static void strobe_meta_ad(strobe_t *s, const uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_M | FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_ad(strobe_t *s, const uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_prf(strobe_t *s, uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, more); strobe_squeeze(s, d, n); }
static void strobe_init(strobe_t *s, const uint8_t *label, size_t n) {
    memset(s, 0, sizeof *s);
    static const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(s->st.b, hdr, 6);
    memcpy(s->st.b + 6, "STROBEv1.0.2", 12);
    keccak_f1600(s->st.w);
    strobe_meta_ad(s, label, n, 0);
}

typedef struct { strobe_t s; } transcript_t;

static void transcript_append_message(transcript_t *t, const char *label, const uint8_t *msg, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(&t->s, (const uint8_t *)label, strlen(label), 0);
    strobe_meta_ad(&t->s, len, 4, 1);
    strobe_ad(&t->s, msg, n, 0);
}
static void transcript_challenge_bytes(transcript_t *t, const char *label, uint8_t *out, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(&t->s, (const uint8_t *)label, strlen(label), 0);
    strobe_meta_ad(&t->s, len, 4, 1);
    strobe_prf(&t->s, out, n, 0);
}
static void transcript_init(transcript_t *t, const char *label) {
    strobe_init(&t->s, (const uint8_t *)"Merlin v1.0", 11);
    transcript_append_message(t, "dom-sep", (const uint8_t *)label, strlen(label));
}
#endif
