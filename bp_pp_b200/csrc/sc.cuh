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

#if defined(__CUDA_ARCH__)
// r[0..8) = a + b + cin, returns the carry out (one self-contained carry chain; cin in {0, 1})
BPPP_D uint32_t add_chain8(uint32_t *r, const uint32_t *a, const uint32_t *b, uint32_t cin) {
    uint32_t c;
    asm volatile("add.cc.u32 %9, %9, 0xFFFFFFFF;\n\t"
                 "addc.cc.u32 %0, %10, %18;\n\t addc.cc.u32 %1, %11, %19;\n\t addc.cc.u32 %2, %12, %20;\n\t addc.cc.u32 %3, %13, %21;\n\t"
                 "addc.cc.u32 %4, %14, %22;\n\t addc.cc.u32 %5, %15, %23;\n\t addc.cc.u32 %6, %16, %24;\n\t addc.cc.u32 %7, %17, %25;\n\t addc.u32 %8, 0, 0;"
                 : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c), "+r"(cin)
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                   "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}
BPPP_D uint32_t add_chain4(uint32_t *r, const uint32_t *a, const uint32_t *b, uint32_t cin) {
    uint32_t c;
    asm volatile("add.cc.u32 %5, %5, 0xFFFFFFFF;\n\t"
                 "addc.cc.u32 %0, %6, %10;\n\t addc.cc.u32 %1, %7, %11;\n\t addc.cc.u32 %2, %8, %12;\n\t addc.cc.u32 %3, %9, %13;\n\t addc.u32 %4, 0, 0;"
                 : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(c), "+r"(cin)
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    return c;
}
// t[0..15] (512-bit LE) mod n on the device: the same four folds as the portable version below, with the products on
// IMAD.WIDE carry chains (fe.cuh) -- 56 wide MADs and ~110 adds instead of ~280 instructions of 64-bit C arithmetic.
BPPP_D void sc_reduce512_dev(uint32_t out[8], const uint32_t t[16]) {
    const uint32_t *H = t + 8;
    const uint32_t zero4[4] = {0, 0, 0, 0};
    // stage 1: P = H * NC_low (12 limbs), R1 = L + P + (H << 128) (13 limbs)
    uint32_t E[13], O[12];
    E[8] = 0; E[12] = 0; O[11] = 0;
    fe_mul4(E, H[0], H[2], H[4], H[6], BPPP_NC0);
    fe_mul4(O, H[1], H[3], H[5], H[7], BPPP_NC0);
    fe_chain4_carry(O + 0, H[0], H[2], H[4], H[6], BPPP_NC1, O[8]);
    fe_chain4_top(E + 2, H[1], H[3], H[5], H[7], BPPP_NC1, E[9]);
    fe_chain4_carry(E + 2, H[0], H[2], H[4], H[6], BPPP_NC2, E[10]);
    fe_chain4_top(O + 2, H[1], H[3], H[5], H[7], BPPP_NC2, O[9]);
    fe_chain4_carry(O + 2, H[0], H[2], H[4], H[6], BPPP_NC3, O[10]);
    fe_chain4_top(E + 4, H[1], H[3], H[5], H[7], BPPP_NC3, E[11]);
    uint32_t P[13], U[13], R1[13], c;
    P[0] = E[0];
    c = add_chain8(P + 1, E + 1, O, 0);
    c = add_chain4(P + 9, E + 9, O + 8, c);            // P[12] stays 0: H * NC_low < 2^384
    U[0] = t[0]; U[1] = t[1]; U[2] = t[2]; U[3] = t[3];
    c = add_chain4(U + 4, t + 4, H, 0);
    U[12] = add_chain4(U + 8, H + 4, zero4, c);
    c = add_chain8(R1, U, P, 0);
    c = add_chain4(R1 + 8, U + 8, P + 8, c);
    R1[12] = U[12] + c;
    // stage 2: H1 = R1[8..13) (130 bits), P2 = H1 * NC_low (< 2^258), R2 = R1[0..8) + P2 + (H1 << 128) (9 limbs)
    const uint32_t *h = R1 + 8;
    uint32_t E2[10], O2[10];
#pragma unroll
    for (int k = 0; k < 10; k++) { E2[k] = 0; O2[k] = 0; }
    fe_chain3(E2 + 0, BPPP_NC0, h[0], h[2], h[4]); fe_chain2(O2 + 0, BPPP_NC0, h[1], h[3]);
    fe_chain3(O2 + 0, BPPP_NC1, h[0], h[2], h[4]); fe_chain2(E2 + 2, BPPP_NC1, h[1], h[3]);
    fe_chain3(E2 + 2, BPPP_NC2, h[0], h[2], h[4]); fe_chain2(O2 + 2, BPPP_NC2, h[1], h[3]);
    fe_chain3(O2 + 2, BPPP_NC3, h[0], h[2], h[4]); fe_chain2(E2 + 4, BPPP_NC3, h[1], h[3]);
    uint32_t P2[9], U2[9], R2[9];
    P2[0] = E2[0];
    add_chain8(P2 + 1, E2 + 1, O2, 0);                 // P2[9] would be 0
    U2[0] = R1[0]; U2[1] = R1[1]; U2[2] = R1[2]; U2[3] = R1[3];
    c = add_chain4(U2 + 4, R1 + 4, h, 0);
    U2[8] = h[4] + c;
    c = add_chain8(R2, U2, P2, 0);
    R2[8] = U2[8] + P2[8] + c;
    // stage 3: R3 = R2[0..8) + hh * NC, hh = R2[8] (a few bits) -> 9 limbs, R3[8] in {0, 1}
    uint32_t hh = R2[8], r8;
    asm volatile("mad.lo.cc.u32 %0, %9, %10, %0;\n\t madc.hi.cc.u32 %1, %9, %10, %1;\n\t madc.lo.cc.u32 %2, %9, %11, %2;\n\t madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
                 "addc.cc.u32 %4, %4, %9;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.cc.u32 %7, %7, 0;\n\t addc.u32 %8, 0, 0;"
                 : "+r"(R2[0]), "+r"(R2[1]), "+r"(R2[2]), "+r"(R2[3]), "+r"(R2[4]), "+r"(R2[5]), "+r"(R2[6]), "+r"(R2[7]), "=r"(r8)
                 : "r"(hh), "r"(BPPP_NC0), "r"(BPPP_NC2));
    asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t madc.lo.cc.u32 %2, %8, %10, %2;\n\t madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                 "addc.cc.u32 %4, %4, 0;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.u32 %7, %7, 0;"
                 : "+r"(R2[1]), "+r"(R2[2]), "+r"(R2[3]), "+r"(R2[4]), "+r"(R2[5]), "+r"(R2[6]), "+r"(R2[7]), "+r"(r8)
                 : "r"(hh), "r"(BPPP_NC1), "r"(BPPP_NC3));
    // stage 4: value = r8 2^256 + R2 < 2n: subtract n when r8 is set or R2 >= n (R2 + NC carries out)
    uint32_t tt[8], cc;
    asm volatile("add.cc.u32 %0, %9, %17;\n\t addc.cc.u32 %1, %10, %18;\n\t addc.cc.u32 %2, %11, %19;\n\t addc.cc.u32 %3, %12, %20;\n\t"
                 "addc.cc.u32 %4, %13, 1;\n\t addc.cc.u32 %5, %14, 0;\n\t addc.cc.u32 %6, %15, 0;\n\t addc.cc.u32 %7, %16, 0;\n\t addc.u32 %8, 0, 0;"
                 : "=&r"(tt[0]), "=&r"(tt[1]), "=&r"(tt[2]), "=&r"(tt[3]), "=&r"(tt[4]), "=&r"(tt[5]), "=&r"(tt[6]), "=&r"(tt[7]), "=&r"(cc)
                 : "r"(R2[0]), "r"(R2[1]), "r"(R2[2]), "r"(R2[3]), "r"(R2[4]), "r"(R2[5]), "r"(R2[6]), "r"(R2[7]),
                   "r"(BPPP_NC0), "r"(BPPP_NC1), "r"(BPPP_NC2), "r"(BPPP_NC3));
    bool take = (r8 | cc) != 0;
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = take ? tt[i] : R2[i];
}
#endif

// t[0..15] (512-bit LE) mod n
BPPP_HD Sc sc_reduce512(const uint32_t t[16]) {
#if defined(__CUDA_ARCH__)
    Sc rd;
    sc_reduce512_dev(rd.v, t);
    return rd;
#else
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
#endif
}

BPPP_HD Sc sc_mul(const Sc &a, const Sc &b) {
    uint32_t t[16];
    wide_mul8(t, a.v, b.v);
    return sc_reduce512(t);
}
BPPP_HD Sc sc_sqr(const Sc &a) {
    uint32_t t[16];
    wide_sqr8(t, a.v);
    return sc_reduce512(t);
}
BPPP_HD Sc sc_dbl(const Sc &a) { return sc_add(a, a); }

// a^(n-2), 4-bit fixed window.  Kept as the cross-check of sc_inv (tests/hostemu).
BPPP_HD Sc sc_inv_fermat(const Sc &a) {
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

// a^-1 mod n by safegcd division steps (modinv.cuh): ~14 k instructions against ~86 k for the power.  Scalars are
// canonical.  Caller handles a == 0 (the reference panics there; this returns 0).
#if defined(__CUDACC__)
static __device__ __noinline__ Sc sc_inv_gcd_call(Sc a) {
    Sc r;
    mi_modinv_words<MIModN>(r.v, a.v);
    return r;
}
#endif
BPPP_HD Sc sc_inv(const Sc &a) {
#if defined(BPPP_INV_FERMAT)
    return sc_inv_fermat(a);
#elif defined(__CUDA_ARCH__)
    return sc_inv_gcd_call(a);
#else
    Sc r;
    mi_modinv_words<MIModN>(r.v, a.v);
    return r;
#endif
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
