// secp256k1 scalar field F_n (k256::Scalar on the device): 8 x 32-bit limbs, little-endian,
// always canonical (< n).  Products are 32x32->64 multiply-accumulates (IMAD.WIDE.U32); the
// reduction folds the high half with 2^256 = NC (mod n), NC = 2^256 - n (129 bits).
//
// Replaces k256::Scalar::{add,sub,mul,invert,from_repr,generate_biased} used at
// reference src/util.rs:28-99, src/circuit.rs:161-524, src/wnla.rs:135-175,
// src/range_proof/reciprocal.rs:117-121,179-183.
#pragma once
#include "fe.cuh"

namespace bppp {

struct Sc {
    uint32_t v[8];
};

#if defined(__CUDA_ARCH__)
#define BPPP_CONST_ARRAY __device__ __constant__ static const
#else
#define BPPP_CONST_ARRAY static const
#endif

// n, little-endian words
#define BPPP_N0 0xD0364141u
#define BPPP_N1 0xBFD25E8Cu
#define BPPP_N2 0xAF48A03Bu
#define BPPP_N3 0xBAAEDCE6u
#define BPPP_N4 0xFFFFFFFEu
#define BPPP_N5 0xFFFFFFFFu
#define BPPP_N6 0xFFFFFFFFu
#define BPPP_N7 0xFFFFFFFFu
// NC = 2^256 - n = 0x1_45512319_50B75FC4_402DA173_2FC9BEBF
#define BPPP_NC0 0x2FC9BEBFu
#define BPPP_NC1 0x402DA173u
#define BPPP_NC2 0x50B75FC4u
#define BPPP_NC3 0x45512319u
// NC4 = 1

BPPP_HD uint32_t sc_n_word(int i) {
    switch (i) {
        case 0: return BPPP_N0; case 1: return BPPP_N1; case 2: return BPPP_N2; case 3: return BPPP_N3;
        case 4: return BPPP_N4; default: return 0xFFFFFFFFu;
    }
}
BPPP_HD uint32_t sc_nc_word(int i) {
    switch (i) {
        case 0: return BPPP_NC0; case 1: return BPPP_NC1; case 2: return BPPP_NC2; case 3: return BPPP_NC3;
        default: return 1u;
    }
}

BPPP_HD Sc sc_zero() { Sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
BPPP_HD Sc sc_from_u64(uint64_t x) { Sc r = sc_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }
BPPP_HD Sc sc_one() { return sc_from_u64(1); }
BPPP_HD bool sc_is_zero(const Sc &a) { uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) m |= a.v[i]; return m == 0; }
BPPP_HD bool sc_eq(const Sc &a, const Sc &b) { uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) m |= a.v[i] ^ b.v[i]; return m == 0; }

// a >= n ?
BPPP_HD bool sc_words_ge_n(const uint32_t a[8]) {
    bool ge = true, decided = false;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        uint32_t nw = sc_n_word(i);
        if (!decided && a[i] != nw) { ge = a[i] > nw; decided = true; }
    }
    return ge;
}
// r = a - n (mod 2^256)
BPPP_HD void sc_words_sub_n(uint32_t r[8], const uint32_t a[8]) {
    uint64_t c = 0;  // add NC, drop the carry: a - n = a + NC - 2^256
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + (i < 4 ? sc_nc_word(i) : (i == 4 ? 1u : 0u)); r[i] = (uint32_t)c; c >>= 32; }
}

BPPP_HD Sc sc_add(const Sc &a, const Sc &b) {
    uint32_t t[8]; uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    Sc r;
    if (c || sc_words_ge_n(t)) sc_words_sub_n(r.v, t);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = t[i];
    }
    return r;
}
BPPP_HD Sc sc_neg(const Sc &a) {
    Sc r; int64_t c = 0;
    bool z = sc_is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)sc_n_word(i) - (int64_t)a.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    if (z) r = sc_zero();
    return r;
}
BPPP_HD Sc sc_sub(const Sc &a, const Sc &b) {
    Sc r; int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    if (c) {  // borrow: add n
        uint64_t cc = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { cc += (uint64_t)r.v[i] + sc_n_word(i); r.v[i] = (uint32_t)cc; cc >>= 32; }
    }
    return r;
}

// t[0..15] (512-bit LE) mod n
BPPP_HD Sc sc_reduce512(const uint32_t t[16]) {
    // stage 1: r1 = lo + hi * NC   (hi 8 limbs) -> 13 limbs
    uint32_t r1[13];
    {
        uint32_t prod[12];
#pragma unroll
        for (int i = 0; i < 12; i++) prod[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {   // hi * (NC0..NC3), schoolbook rows
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                c += (uint64_t)t[8 + i] * sc_nc_word(j) + prod[i + j];
                prod[i + j] = (uint32_t)c; c >>= 32;
            }
            prod[i + 4] = (uint32_t)c;
        }
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {   // lo + prod + (hi << 128)   (NC4 = 1)
            c += (uint64_t)prod[i] + (i < 8 ? t[i] : 0u) + (i >= 4 ? t[4 + i] : 0u);
            r1[i] = (uint32_t)c; c >>= 32;
        }
        r1[12] = (uint32_t)c;
    }
    // stage 2: r2 = lo1 + hi1 * NC (hi1 = r1[8..12], 5 limbs) -> 10 limbs
    uint32_t r2[10];
    {
        uint32_t prod[9];
#pragma unroll
        for (int i = 0; i < 9; i++) prod[i] = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                c += (uint64_t)r1[8 + i] * sc_nc_word(j) + prod[i + j];
                prod[i + j] = (uint32_t)c; c >>= 32;
            }
            prod[i + 4] = (uint32_t)c;
        }
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            c += (uint64_t)prod[i] + (i < 8 ? r1[i] : 0u) + (i >= 4 ? r1[4 + i] : 0u);
            r2[i] = (uint32_t)c; c >>= 32;
        }
        r2[9] = (uint32_t)c;
    }
    // stage 3: r3 = lo2 + hi2 * NC, hi2 = r2[8] (r2[9] == 0 by the size bound) -> 9 limbs, r3[8] in {0,1}
    uint32_t r3[9];
    {
        uint32_t h = r2[8];
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { c += (uint64_t)h * sc_nc_word(j) + r2[j]; r3[j] = (uint32_t)c; c >>= 32; }
        c += (uint64_t)r2[4] + h; r3[4] = (uint32_t)c; c >>= 32;
#pragma unroll
        for (int j = 5; j < 8; j++) { c += r2[j]; r3[j] = (uint32_t)c; c >>= 32; }
        r3[8] = (uint32_t)c;
    }
    // stage 4: fold the last bit, then one conditional subtraction
    uint32_t r4[8];
    if (r3[8]) sc_words_sub_n(r4, r3);  // r3 - n = r3 + NC - 2^256: drops bit 256, adds NC
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) r4[i] = r3[i];
    }
    Sc r;
    if (sc_words_ge_n(r4)) sc_words_sub_n(r.v, r4);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = r4[i];
    }
    return r;
}

BPPP_HD Sc sc_mul(const Sc &a, const Sc &b) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[i] * b.v[j] + t[i + j];
            t[i + j] = (uint32_t)c; c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
    return sc_reduce512(t);
}
BPPP_HD Sc sc_sqr(const Sc &a) { return sc_mul(a, a); }
BPPP_HD Sc sc_dbl(const Sc &a) { return sc_add(a, a); }

// a^(n-2), 4-bit fixed window.  Caller handles a == 0 (the reference panics there).
BPPP_HD Sc sc_inv(const Sc &a) {
    Sc tab[16];
    tab[0] = sc_one(); tab[1] = a;
#pragma unroll 1
    for (int i = 2; i < 16; i++) tab[i] = sc_mul(tab[i - 1], a);
    // n - 2, little-endian words
    const uint32_t e[8] = {BPPP_N0 - 2u, BPPP_N1, BPPP_N2, BPPP_N3, BPPP_N4, BPPP_N5, BPPP_N6, BPPP_N7};
    Sc acc = sc_one();
#pragma unroll 1
    for (int w = 63; w >= 0; w--) {
        if (w != 63) { acc = sc_sqr(acc); acc = sc_sqr(acc); acc = sc_sqr(acc); acc = sc_sqr(acc); }
        uint32_t d = (e[w >> 3] >> (4 * (w & 7))) & 15u;
        acc = sc_mul(acc, tab[d]);
    }
    return acc;
}

// Scalar::from_repr: 32 bytes big-endian; returns false when >= n
BPPP_HD bool sc_from_be32(Sc &r, const uint8_t *b) { be32_to_words(r.v, b); return !sc_words_ge_n(r.v); }
BPPP_HD void sc_to_be32(uint8_t *b, const Sc &a) { words_to_be32(b, a.v); }
// Scalar::generate_biased: 64 bytes big-endian mod n [recalled convention, SURVEY App. D]
BPPP_HD Sc sc_from_wide_be64(const uint8_t *b) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint8_t *p = b + 4 * (15 - i);
        t[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
    return sc_reduce512(t);
}

}  // namespace bppp
