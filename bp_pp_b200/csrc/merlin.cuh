// Keccak-f[1600] / STROBE-128 / Merlin 3.0.0 transcript, host+device, byte-compatible with the
// `merlin` crate the reference drives at src/transcript.rs:6-14, src/wnla.rs:88-94,162-168,
// src/circuit.rs:155-164,189-191,347-355,472-474, src/range_proof/reciprocal.rs:99-100,114-115.
// Construction restated from SURVEY Appendix D (merlin 3.0.0 / strobe-rs semantics).
// One transcript per proof lives in one thread; the state is 25 x u64 lanes + 3 bytes.
#pragma once
#include <stdint.h>
#include "fe.cuh"

namespace bppp {

BPPP_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

BPPP_HD void keccak_f1600(uint64_t a[25]) {
    const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
        0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
        0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    uint64_t s[25];
#pragma unroll
    for (int i = 0; i < 25; i++) s[i] = a[i];
#pragma unroll 1
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20];
        uint64_t c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21];
        uint64_t c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22];
        uint64_t c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23];
        uint64_t c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
        uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1);
        uint64_t d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
#pragma unroll
        for (int y = 0; y < 25; y += 5) { s[y] ^= d0; s[y + 1] ^= d1; s[y + 2] ^= d2; s[y + 3] ^= d3; s[y + 4] ^= d4; }
        // rho + pi:  b[y + 5*((2x+3y)%5)] = rotl(s[x+5y], r[x][y])
        uint64_t b[25];
        b[0] = s[0];
        b[10] = rotl64(s[1], 1);   b[20] = rotl64(s[2], 62);  b[5] = rotl64(s[3], 28);   b[15] = rotl64(s[4], 27);
        b[16] = rotl64(s[5], 36);  b[1] = rotl64(s[6], 44);   b[11] = rotl64(s[7], 6);   b[21] = rotl64(s[8], 55);
        b[6] = rotl64(s[9], 20);   b[7] = rotl64(s[10], 3);   b[17] = rotl64(s[11], 10); b[2] = rotl64(s[12], 43);
        b[12] = rotl64(s[13], 25); b[22] = rotl64(s[14], 39); b[23] = rotl64(s[15], 41); b[8] = rotl64(s[16], 45);
        b[18] = rotl64(s[17], 15); b[3] = rotl64(s[18], 21);  b[13] = rotl64(s[19], 8);  b[14] = rotl64(s[20], 18);
        b[24] = rotl64(s[21], 2);  b[9] = rotl64(s[22], 61);  b[19] = rotl64(s[23], 56); b[4] = rotl64(s[24], 14);
#pragma unroll
        for (int y = 0; y < 25; y += 5) {
            s[y] = b[y] ^ (~b[y + 1] & b[y + 2]);
            s[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]);
            s[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
            s[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y]);
            s[y + 4] = b[y + 4] ^ (~b[y] & b[y + 1]);
        }
        s[0] ^= RC[rnd];
    }
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = s[i];
}

struct Merlin {
    uint64_t st[25];
    uint32_t pos, pos_begin, cur_flags;
    uint32_t _pad;
};
static constexpr uint32_t STROBE_R = 166;

BPPP_HD void strobe_xor_byte(Merlin &m, uint32_t pos, uint8_t b) { m.st[pos >> 3] ^= (uint64_t)b << (8 * (pos & 7)); }
BPPP_HD uint8_t strobe_get_byte(const Merlin &m, uint32_t pos) { return (uint8_t)(m.st[pos >> 3] >> (8 * (pos & 7))); }
BPPP_HD void strobe_clear_byte(Merlin &m, uint32_t pos) { m.st[pos >> 3] &= ~((uint64_t)0xFF << (8 * (pos & 7))); }

BPPP_HD void strobe_run_f(Merlin &m) {
    strobe_xor_byte(m, m.pos, (uint8_t)m.pos_begin);
    strobe_xor_byte(m, m.pos + 1, 0x04);
    strobe_xor_byte(m, STROBE_R + 1, 0x80);
    keccak_f1600(m.st);
    m.pos = 0; m.pos_begin = 0;
}
BPPP_HD void strobe_absorb(Merlin &m, const uint8_t *d, uint32_t n) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        strobe_xor_byte(m, m.pos, d[i]);
        if (++m.pos == STROBE_R) strobe_run_f(m);
    }
}
BPPP_HD void strobe_squeeze(Merlin &m, uint8_t *d, uint32_t n) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        d[i] = strobe_get_byte(m, m.pos);
        strobe_clear_byte(m, m.pos);
        if (++m.pos == STROBE_R) strobe_run_f(m);
    }
}
BPPP_HD void strobe_begin_op(Merlin &m, uint32_t flags, bool more) {
    if (more) return;
    uint8_t hdr[2] = {(uint8_t)m.pos_begin, (uint8_t)flags};
    m.pos_begin = m.pos + 1;
    m.cur_flags = flags;
    strobe_absorb(m, hdr, 2);
    if ((flags & (4u | 32u)) && m.pos != 0) strobe_run_f(m);
}
BPPP_HD void strobe_meta_ad(Merlin &m, const uint8_t *d, uint32_t n, bool more) { strobe_begin_op(m, 16u | 2u, more); strobe_absorb(m, d, n); }
BPPP_HD void strobe_ad(Merlin &m, const uint8_t *d, uint32_t n, bool more) { strobe_begin_op(m, 2u, more); strobe_absorb(m, d, n); }
BPPP_HD void strobe_prf(Merlin &m, uint8_t *d, uint32_t n) { strobe_begin_op(m, 1u | 2u | 4u, false); strobe_squeeze(m, d, n); }

// Transcript::append_message(label, msg)
BPPP_HD void merlin_append(Merlin &m, const char *label, uint32_t label_len, const uint8_t *msg, uint32_t n) {
    uint8_t l4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(m, (const uint8_t *)label, label_len, false);
    strobe_meta_ad(m, l4, 4, true);
    strobe_ad(m, msg, n, false);
}
// Transcript::append_u64
BPPP_HD void merlin_append_u64(Merlin &m, const char *label, uint32_t label_len, uint64_t x) {
    uint8_t b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    merlin_append(m, label, label_len, b, 8);
}
// Transcript::challenge_bytes
BPPP_HD void merlin_challenge(Merlin &m, const char *label, uint32_t label_len, uint8_t *out, uint32_t n) {
    uint8_t l4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(m, (const uint8_t *)label, label_len, false);
    strobe_meta_ad(m, l4, 4, true);
    strobe_prf(m, out, n);
}
// Transcript::new(label)
BPPP_HD void merlin_init(Merlin &m, const uint8_t *label, uint32_t label_len) {
#pragma unroll
    for (int i = 0; i < 25; i++) m.st[i] = 0;
    m.pos = 0; m.pos_begin = 0; m.cur_flags = 0; m._pad = 0;
    const uint8_t hdr[18] = {1, STROBE_R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    for (uint32_t i = 0; i < 18; i++) strobe_xor_byte(m, i, hdr[i]);
    keccak_f1600(m.st);
    const char proto[] = "Merlin v1.0";
    strobe_meta_ad(m, (const uint8_t *)proto, 11, false);
    merlin_append(m, "dom-sep", 7, label, label_len);
}

#define BPPP_LBL(s) s, (uint32_t)(sizeof(s) - 1)

}  // namespace bppp
