// Host-side Fiat-Shamir transcript: merlin 3.0.0 over STROBE-128 over Keccak-f[1600], plus the reference's
// `CurdleproofsTranscript` helpers (/root/reference/src/transcript.rs:28-61).  Sequential ~KB-sized hashing per round,
// so it stays on the host cores (SURVEY.md section 8f lists a device-side transcript as the next widening step).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "fr.hpp"

namespace cdp_host {

class Keccak {
   public:
    static void permute(uint64_t a[25]) {
        static const uint64_t RC[24] = {0x1ULL, 0x8082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x808bULL, 0x80000001ULL,
                                        0x8000000080008081ULL, 0x8000000000008009ULL, 0x8aULL, 0x88ULL, 0x80008009ULL, 0x8000000aULL,
                                        0x8000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                        0x8000000000008002ULL, 0x8000000000000080ULL, 0x800aULL, 0x800000008000000aULL,
                                        0x8000000080008081ULL, 0x8000000000008080ULL, 0x80000001ULL, 0x8000000080008008ULL};
        // rho offsets indexed [x + 5y]
        static const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        for (int round = 0; round < 24; round++) {
            uint64_t c[5], d[5], b[25];
            for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
            for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
            for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
            // rho + pi: B[y, 2x+3y] = rot(A[x, y])
            for (int x = 0; x < 5; x++)
                for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(a[x + 5 * y], RHO[x + 5 * y]);
            for (int y = 0; y < 5; y++)
                for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
            a[0] ^= RC[round];
        }
    }

   private:
    static uint64_t rotl(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }
};

class Strobe128 {
   public:
    explicit Strobe128(const char *protocol_label) {
        memset(st_.b, 0, 200);
        const uint8_t init[6] = {1, RATE + 2, 1, 0, 1, 96};
        memcpy(st_.b, init, 6);
        memcpy(st_.b + 6, "STROBEv1.0.2", 12);
        Keccak::permute(st_.w);
        meta_ad(reinterpret_cast<const uint8_t *>(protocol_label), strlen(protocol_label), false);
    }
    void meta_ad(const uint8_t *data, size_t n, bool more) { begin(FLAG_M | FLAG_A, more); absorb(data, n); }
    void ad(const uint8_t *data, size_t n, bool more) { begin(FLAG_A, more); absorb(data, n); }
    void prf(uint8_t *out, size_t n, bool more) { begin(FLAG_I | FLAG_A | FLAG_C, more); squeeze(out, n); }

   private:
    static constexpr int RATE = 166;
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    union { uint64_t w[25]; uint8_t b[200]; } st_;
    uint8_t pos_ = 0, pos_begin_ = 0;

    void run_f() {
        st_.b[pos_] ^= pos_begin_;
        st_.b[pos_ + 1] ^= 0x04;
        st_.b[RATE + 1] ^= 0x80;
        Keccak::permute(st_.w);
        pos_ = 0;
        pos_begin_ = 0;
    }
    void absorb(const uint8_t *d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            st_.b[pos_++] ^= d[i];
            if (pos_ == RATE) run_f();
        }
    }
    void squeeze(uint8_t *d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            d[i] = st_.b[pos_];
            st_.b[pos_++] = 0;
            if (pos_ == RATE) run_f();
        }
    }
    void begin(uint8_t flags, bool more) {
        if (more) return;
        uint8_t hdr[2] = {pos_begin_, flags};
        pos_begin_ = pos_ + 1;
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
    }
};

// merlin::Transcript + the CurdleproofsTranscript trait
class Transcript {
   public:
    explicit Transcript(const char *label) : s_("Merlin v1.0") { append_message("dom-sep", reinterpret_cast<const uint8_t *>(label), strlen(label)); }
    void append_message(const char *label, const uint8_t *msg, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s_.meta_ad(reinterpret_cast<const uint8_t *>(label), strlen(label), false);
        s_.meta_ad(len, 4, true);
        s_.ad(msg, n, false);
    }
    void challenge_bytes(const char *label, uint8_t *out, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s_.meta_ad(reinterpret_cast<const uint8_t *>(label), strlen(label), false);
        s_.meta_ad(len, 4, true);
        s_.prf(out, n, false);
    }
    // `append(label, &G1)` : 48-byte compressed encoding (src/transcript.rs:29-33)
    void append_point(const char *label, const uint8_t comp[48]) { append_message(label, comp, 48); }
    void append_fr(const char *label, const Fr &x) {
        uint8_t b[32];
        x.to_bytes(b);
        append_message(label, b, 32);
    }
    // `append(label, &Vec<G1Affine>)`: u64-LE length, then the elements (ark-serialize)
    void append_point_vec(const char *label, const uint8_t *comp, size_t count) {
        buf_.resize(8 + 48 * count);
        uint64_t len = count;
        memcpy(buf_.data(), &len, 8);
        memcpy(buf_.data() + 8, comp, 48 * count);
        append_message(label, buf_.data(), buf_.size());
    }
    void append_fr_vec(const char *label, const Fr *v, size_t count) {
        buf_.resize(8 + 32 * count);
        uint64_t len = count;
        memcpy(buf_.data(), &len, 8);
        for (size_t i = 0; i < count; i++) v[i].to_bytes(buf_.data() + 8 + 32 * i);
        append_message(label, buf_.data(), buf_.size());
    }
    // get_and_append_challenge (src/transcript.rs:41-54): 64 challenge bytes, first 32 with bit 255 cleared must be a
    // non-zero value < r, otherwise draw again; the accepted challenge is fed back
    Fr challenge(const char *label) {
        for (;;) {
            uint8_t buf[64];
            challenge_bytes(label, buf, 64);
            buf[31] &= 0x7F;
            Fr e;
            if (!Fr::from_bytes(buf, e) || e.is_zero()) continue;
            append_fr(label, e);
            return e;
        }
    }

   private:
    Strobe128 s_;
    std::vector<uint8_t> buf_;
};

}  // namespace cdp_host
