// secp256k1 base field F_p, p = 2^256 - 2^32 - 977, for sm_100a.
//
// Representation: 10 limbs of 26 bits held in 32-bit registers, lazily reduced.
//   value = sum n[i] * 2^(26 i)   (any representative mod p)
// "Magnitude" m bounds the limbs: n[i] <= 2m(2^26-1) for i<9, n[9] <= 2m(2^22-1).
// fe_mul/fe_sqr accept magnitudes with ma*mb <= 64 and return magnitude 1; add/negate/mul_int
// are carry-free limb-wise operations that only grow the magnitude.  The product columns are
// 64-bit sums of 32x32->64 multiply-accumulates (IMAD.WIDE.U32 on the FMA pipe, 64-bit
// accumulate for free) so there is no carry chain on the ALU pipe inside the product.
//
// This replaces k256::FieldElement (not in /root/reference; Cargo.lock:411-414) on the device.
// The internal representation is unobservable; only canonical bytes leave the device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BPPP_HD __host__ __device__ __forceinline__
#define BPPP_D __device__ __forceinline__
#else
#define BPPP_HD inline
#define BPPP_D inline
#endif

// Host-side verification build (tests/hostemu): every Fe carries exact upper bounds on its limbs
// (lim for n[0..8], lim9 for n[9]) and every operation asserts its preconditions on them, so the
// lazy-reduction discipline of the point formulas is machine-checked on each test run.
#if defined(BPPP_VERIFY_MAG)
#include <assert.h>
#define BPPP_MAG_FIELD uint64_t lim, lim9;
#define BPPP_SET_MAG(r, m) ((r).lim = 2ull * (uint64_t)(m) * 0x3FFFFFFull, (r).lim9 = 2ull * (uint64_t)(m) * 0x3FFFFFull)
#define BPPP_SET_LIM(r, l, l9) ((r).lim = (l), (r).lim9 = (l9))
#define BPPP_ASSERT(c) assert(c)
#else
#define BPPP_MAG_FIELD
#define BPPP_SET_MAG(r, m) ((void)0)
#define BPPP_SET_LIM(r, l, l9) ((void)0)
#define BPPP_ASSERT(c) ((void)0)
#endif

// A translation unit may define BPPP_FE_NOINLINE: fe_mul / fe_sqr then compile to real device functions taking
// their operands BY VALUE (nvcc passes the 10-word structs in registers, ~15 MOVs per call, no stack traffic).
// Fully inlined point formulas are ~100 KB of SASS per ladder step and thrash the instruction cache
// (ncu on k_v_var2: stall_no_instruction 5.4 warps per issue); with calls the hot loop is < 10 KB.

namespace bppp {

struct Fe {
    uint32_t n[10];
    BPPP_MAG_FIELD
};

static constexpr uint32_t FE_M26 = 0x3FFFFFFu;
static constexpr uint32_t FE_M22 = 0x3FFFFFu;
// 2^260 = R0 + R1 * 2^26 (mod p);  2^256 = 977 + 64 * 2^26 (mod p)
static constexpr uint32_t FE_R0 = 0x3D10u;
static constexpr uint32_t FE_R1 = 0x400u;

#if defined(BPPP_VERIFY_MAG)
inline void fe_check(const Fe &a) {
    for (int i = 0; i < 9; i++) assert((uint64_t)a.n[i] <= a.lim);
    assert((uint64_t)a.n[9] <= a.lim9);
    assert(a.lim <= 0xFFFFFFFFull && a.lim9 <= 0xFFFFFFFFull);   // bounds themselves must fit a 32-bit limb
}
#else
BPPP_HD void fe_check(const Fe &) {}
#endif

BPPP_HD Fe fe_zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.n[i] = 0;
    BPPP_SET_MAG(r, 0);
    return r;
}
BPPP_HD Fe fe_from_u32(uint32_t x) {  // x < 2^26
    Fe r = fe_zero();
    r.n[0] = x;
    BPPP_SET_MAG(r, 1);
    return r;
}
BPPP_HD Fe fe_one() { return fe_from_u32(1); }

// 8 x u32 little-endian words (a canonical or any 256-bit integer) -> Fe, magnitude 1
BPPP_HD Fe fe_from_words(const uint32_t w[8]) {
    Fe r;
    r.n[0] = w[0] & FE_M26;
    r.n[1] = ((w[0] >> 26) | (w[1] << 6)) & FE_M26;
    r.n[2] = ((w[1] >> 20) | (w[2] << 12)) & FE_M26;
    r.n[3] = ((w[2] >> 14) | (w[3] << 18)) & FE_M26;
    r.n[4] = ((w[3] >> 8) | (w[4] << 24)) & FE_M26;
    r.n[5] = (w[4] >> 2) & FE_M26;
    r.n[6] = ((w[4] >> 28) | (w[5] << 4)) & FE_M26;
    r.n[7] = ((w[5] >> 22) | (w[6] << 10)) & FE_M26;
    r.n[8] = ((w[6] >> 16) | (w[7] << 16)) & FE_M26;
    r.n[9] = w[7] >> 10;
    BPPP_SET_MAG(r, 1);
    return r;
}

// weak normalisation: any magnitude <= 32 -> magnitude 1 (limbs < 2^26 except a small excess in n[0..1])
BPPP_HD Fe fe_normalize_weak(const Fe &a) {
    fe_check(a);
#if defined(BPPP_VERIFY_MAG)
    assert(a.lim + (a.lim9 >> 22) * 977ull + 64 <= 0xFFFFFFFFull);   // no 32-bit overflow in the carry pass
#endif
    Fe r;
    uint32_t x = a.n[9] >> 22;
    uint32_t t9 = a.n[9] & FE_M22;
    uint32_t t = a.n[0] + x * 977u;
    uint32_t c = t >> 26; r.n[0] = t & FE_M26;
    t = a.n[1] + (x << 6) + c; c = t >> 26; r.n[1] = t & FE_M26;
#pragma unroll
    for (int i = 2; i < 9; i++) { t = a.n[i] + c; c = t >> 26; r.n[i] = t & FE_M26; }
    r.n[9] = t9 + c;
    BPPP_SET_LIM(r, 0x3FFFFFFull, 0x3FFFFFull + 64);
    fe_check(r);
    return r;
}

// full normalisation to the canonical representative in [0, p)
BPPP_HD Fe fe_normalize(const Fe &a) {
    Fe r = fe_normalize_weak(a);
    // after the weak pass: value < 2^256 + small.  One more fold of bit 256 and a final conditional -p.
    uint32_t x = r.n[9] >> 22;
    r.n[9] &= FE_M22;
    uint32_t t = r.n[0] + x * 977u, c = t >> 26; r.n[0] = t & FE_M26;
    t = r.n[1] + (x << 6) + c; c = t >> 26; r.n[1] = t & FE_M26;
#pragma unroll
    for (int i = 2; i < 10; i++) { t = r.n[i] + c; c = t >> 26; r.n[i] = t & FE_M26; }
    // now r < 2^256 (n[9] may be exactly 2^22 only if everything below overflowed; fold once more)
    x = r.n[9] >> 22;
    r.n[9] &= FE_M22;
    t = r.n[0] + x * 977u; c = t >> 26; r.n[0] = t & FE_M26;
    t = r.n[1] + (x << 6) + c; c = t >> 26; r.n[1] = t & FE_M26;
#pragma unroll
    for (int i = 2; i < 10; i++) { t = r.n[i] + c; c = t >> 26; r.n[i] = t & FE_M26; }
    // r in [0, 2^256): subtract p if r >= p.  r >= p  <=>  n[9]==M22, n[2..8]==M26, and (n[1],n[0]) >= (0x3FFFFBF, 0x3FFFC2F)
    uint32_t m = r.n[9] ^ FE_M22;
#pragma unroll
    for (int i = 2; i < 9; i++) m |= r.n[i] ^ FE_M26;
    bool ge = (m == 0) && ((r.n[1] > 0x3FFFFBFu) || (r.n[1] == 0x3FFFFBFu && r.n[0] >= 0x3FFFC2Fu));
    if (ge) {
        // r - p = r + (2^32 + 977) - 2^256
        t = r.n[0] + 977u; c = t >> 26; r.n[0] = t & FE_M26;
        t = r.n[1] + 64u + c; c = t >> 26; r.n[1] = t & FE_M26;
#pragma unroll
        for (int i = 2; i < 9; i++) { t = r.n[i] + c; c = t >> 26; r.n[i] = t & FE_M26; }
        r.n[9] = (r.n[9] + c) & FE_M22;
    }
    BPPP_SET_MAG(r, 1);
    return r;
}

// canonical Fe -> 8 x u32 little-endian words
BPPP_HD void fe_to_words(uint32_t w[8], const Fe &a_canonical) {
    const uint32_t *n = a_canonical.n;
    w[0] = n[0] | (n[1] << 26);
    w[1] = (n[1] >> 6) | (n[2] << 20);
    w[2] = (n[2] >> 12) | (n[3] << 14);
    w[3] = (n[3] >> 18) | (n[4] << 8);
    w[4] = (n[4] >> 24) | (n[5] << 2) | (n[6] << 28);
    w[5] = (n[6] >> 4) | (n[7] << 22);
    w[6] = (n[7] >> 10) | (n[8] << 16);
    w[7] = (n[8] >> 16) | (n[9] << 10);
}

BPPP_HD bool fe_is_zero_canonical(const Fe &a) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) m |= a.n[i];
    return m == 0;
}
BPPP_HD bool fe_is_zero(const Fe &a) { return fe_is_zero_canonical(fe_normalize(a)); }
// value == 0 (mod p) for a lazily reduced element, without the full normalisation: after one weak pass the value is
// below 2^256 + 2^235, so it is a multiple of p only if it is exactly 0 or exactly p
BPPP_HD bool fe_normalizes_to_zero(const Fe &a) {
    Fe t = fe_normalize_weak(a);
    uint32_t z = t.n[0] | t.n[1] | t.n[9];
    uint32_t pm = (t.n[0] ^ 0x3FFFC2Fu) | (t.n[1] ^ 0x3FFFFBFu) | (t.n[9] ^ FE_M22);
#pragma unroll
    for (int i = 2; i < 9; i++) { z |= t.n[i]; pm |= t.n[i] ^ FE_M26; }
    return z == 0 || pm == 0;
}
BPPP_HD bool fe_equal_canonical(const Fe &a, const Fe &b) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) m |= a.n[i] ^ b.n[i];
    return m == 0;
}
BPPP_HD bool fe_is_odd_canonical(const Fe &a) { return a.n[0] & 1u; }

BPPP_HD Fe fe_add(const Fe &a, const Fe &b) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.n[i] = a.n[i] + b.n[i];
#if defined(BPPP_VERIFY_MAG)
    r.lim = a.lim + b.lim; r.lim9 = a.lim9 + b.lim9;
#endif
    fe_check(r);
    return r;
}
// -a for a of magnitude <= m; result magnitude m+1
BPPP_HD Fe fe_negate(const Fe &a, int m) {
    Fe r;
    const uint32_t k = 2u * (uint32_t)(m + 1);
#if defined(BPPP_VERIFY_MAG)
    assert(a.lim <= (uint64_t)k * 0x3FFFC2Full && a.lim9 <= (uint64_t)k * 0x3FFFFFull && k <= 64);
#endif
    r.n[0] = 0x3FFFC2Fu * k - a.n[0];
    r.n[1] = 0x3FFFFBFu * k - a.n[1];
#pragma unroll
    for (int i = 2; i < 9; i++) r.n[i] = FE_M26 * k - a.n[i];
    r.n[9] = FE_M22 * k - a.n[9];
    BPPP_SET_LIM(r, (uint64_t)k * 0x3FFFFFFull, (uint64_t)k * 0x3FFFFFull);
    fe_check(r);
    return r;
}
// a - b for b of magnitude <= mb; result magnitude mag(a) + mb + 1
BPPP_HD Fe fe_sub(const Fe &a, const Fe &b, int mb) { return fe_add(a, fe_negate(b, mb)); }
BPPP_HD Fe fe_mul_int(const Fe &a, uint32_t k) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.n[i] = a.n[i] * k;
#if defined(BPPP_VERIFY_MAG)
    r.lim = a.lim * k; r.lim9 = a.lim9 * k;
#endif
    fe_check(r);
    return r;
}
BPPP_HD Fe fe_cmov(const Fe &a, const Fe &b, bool take_b) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.n[i] = take_b ? b.n[i] : a.n[i];
#if defined(BPPP_VERIFY_MAG)
    r.lim = a.lim > b.lim ? a.lim : b.lim; r.lim9 = a.lim9 > b.lim9 ? a.lim9 : b.lim9;
#endif
    return r;
}

// Reduce 19 product columns c[0..18] (each < 2^63.9) to limbs below 2^27 + 2^12 (top limb < 2^22).
// Carry-save instead of carry-propagate: every 64-bit column is cut into 26 + 26 + 12 bit pieces that are
// re-assembled into lazy limbs with independent 3-input adds, so the only serial carry chain left is three
// steps long (the previous two 10-step chains made the kernels wait on fixed-latency dependencies, ncu
// "stall_wait" 2.2 per issue).  2^260 = R0 + R1 2^26 and 2^256 = 977 + 64 2^26 (mod p).
BPPP_HD Fe fe_reduce_columns(uint64_t c[19]) {
    // high columns 10..18 -> lazy limbs H[0..10] at weights 2^(26 (10 + j))
    uint32_t lo[9], mid[9], hi[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        uint64_t v = c[10 + k];
        lo[k] = (uint32_t)v & FE_M26;
        mid[k] = (uint32_t)(v >> 26) & FE_M26;
        hi[k] = (uint32_t)(v >> 52);
    }
    uint32_t H[11];
    H[0] = lo[0];
    H[1] = lo[1] + mid[0];
#pragma unroll
    for (int j = 2; j < 9; j++) H[j] = lo[j] + mid[j - 1] + hi[j - 2];
    H[9] = mid[8] + hi[7];
    H[10] = hi[8];
    // fold into the low columns (two more columns appear at positions 10 and 11)
    uint64_t low[12];
#pragma unroll
    for (int k = 0; k < 10; k++) low[k] = c[k];
    low[10] = 0; low[11] = 0;
#pragma unroll
    for (int j = 0; j < 11; j++) {
        low[j] += (uint64_t)H[j] * FE_R0;
        low[j + 1] += (uint64_t)H[j] * FE_R1;
    }
    // low columns -> lazy limbs L[0..11]  (low[10] < 2^38, low[11] < 2^23, so L[12], L[13] vanish)
    uint32_t lo2[12], mid2[12], hi2[10];
#pragma unroll
    for (int k = 0; k < 12; k++) {
        uint64_t v = low[k];
        lo2[k] = (uint32_t)v & FE_M26;
        mid2[k] = (uint32_t)(v >> 26) & FE_M26;
        if (k < 10) hi2[k] = (uint32_t)(v >> 52);
    }
    uint32_t L[12];
    L[0] = lo2[0];
    L[1] = lo2[1] + mid2[0];
#pragma unroll
    for (int j = 2; j < 12; j++) L[j] = lo2[j] + mid2[j - 1] + hi2[j - 2];
    // top: limbs 10, 11 and the bits of limb 9 above 2^22 wrap around once more
    uint32_t x = L[9] >> 22;
    uint64_t e0 = (uint64_t)L[0] + (uint64_t)L[10] * FE_R0 + (uint64_t)x * 977u;
    uint64_t e1 = (uint64_t)L[1] + (uint64_t)L[10] * FE_R1 + (uint64_t)L[11] * FE_R0 + (uint64_t)(x << 6);
    uint64_t e2 = (uint64_t)L[2] + (uint64_t)L[11] * FE_R1;
    Fe r;
    r.n[0] = (uint32_t)e0 & FE_M26; e1 += e0 >> 26;
    r.n[1] = (uint32_t)e1 & FE_M26; e2 += e1 >> 26;
    r.n[2] = (uint32_t)e2 & FE_M26;
    r.n[3] = L[3] + (uint32_t)(e2 >> 26);
#pragma unroll
    for (int k = 4; k < 9; k++) r.n[k] = L[k];
    r.n[9] = L[9] & FE_M22;
    BPPP_SET_LIM(r, (1ull << 27) + (1ull << 12) + 128, 0x3FFFFFull);
    fe_check(r);
    return r;
}

BPPP_HD Fe fe_mul_inl(const Fe &a, const Fe &b) {
    fe_check(a); fe_check(b);
#if defined(BPPP_VERIFY_MAG)
    assert((unsigned __int128)10 * a.lim * b.lim < ((unsigned __int128)15 << 60));   // columns < 2^63.9: room for the folds
#endif
    uint64_t c[19];
#pragma unroll
    for (int k = 0; k < 19; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
#pragma unroll
        for (int j = 0; j < 10; j++) c[i + j] += (uint64_t)a.n[i] * b.n[j];
    }
    return fe_reduce_columns(c);
}

BPPP_HD Fe fe_sqr_inl(const Fe &a) {
    fe_check(a);
#if defined(BPPP_VERIFY_MAG)
    assert(a.lim < (1ull << 31) && (unsigned __int128)10 * a.lim * a.lim < ((unsigned __int128)15 << 60));
#endif
    uint64_t c[19];
#pragma unroll
    for (int k = 0; k < 19; k++) c[k] = 0;
    uint32_t d[10];
#pragma unroll
    for (int i = 0; i < 10; i++) d[i] = a.n[i] << 1;   // magnitude <= 8 => limbs < 2^30, doubled < 2^31
#pragma unroll
    for (int i = 0; i < 10; i++) {
        c[2 * i] += (uint64_t)a.n[i] * a.n[i];
#pragma unroll
        for (int j = i + 1; j < 10; j++) c[i + j] += (uint64_t)d[i] * a.n[j];
    }
    return fe_reduce_columns(c);
}

#if defined(__CUDACC__) && defined(BPPP_FE_NOINLINE)
static __device__ __noinline__ Fe fe_mul_call(Fe a, Fe b) { return fe_mul_inl(a, b); }
static __device__ __noinline__ Fe fe_sqr_call(Fe a) { return fe_sqr_inl(a); }
__host__ __device__ __forceinline__ Fe fe_mul(const Fe &a, const Fe &b) {
#if defined(__CUDA_ARCH__)
    return fe_mul_call(a, b);
#else
    return fe_mul_inl(a, b);
#endif
}
__host__ __device__ __forceinline__ Fe fe_sqr(const Fe &a) {
#if defined(__CUDA_ARCH__)
    return fe_sqr_call(a);
#else
    return fe_sqr_inl(a);
#endif
}
#else
BPPP_HD Fe fe_mul(const Fe &a, const Fe &b) { return fe_mul_inl(a, b); }
BPPP_HD Fe fe_sqr(const Fe &a) { return fe_sqr_inl(a); }
#endif

BPPP_HD Fe fe_sqr_n(Fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe_sqr(a);
    return a;
}

// a^(p-2): 255 squarings + 15 multiplications.  a of magnitude <= 8.  fe_inv(0) = 0.
BPPP_HD Fe fe_inv(const Fe &a) {
    Fe x2 = fe_mul(fe_sqr(a), a);
    Fe x3 = fe_mul(fe_sqr(x2), a);
    Fe x6 = fe_mul(fe_sqr_n(x3, 3), x3);
    Fe x9 = fe_mul(fe_sqr_n(x6, 3), x3);
    Fe x11 = fe_mul(fe_sqr_n(x9, 2), x2);
    Fe x22 = fe_mul(fe_sqr_n(x11, 11), x11);
    Fe x44 = fe_mul(fe_sqr_n(x22, 22), x22);
    Fe x88 = fe_mul(fe_sqr_n(x44, 44), x44);
    Fe x176 = fe_mul(fe_sqr_n(x88, 88), x88);
    Fe x220 = fe_mul(fe_sqr_n(x176, 44), x44);
    Fe x223 = fe_mul(fe_sqr_n(x220, 3), x3);
    Fe t = fe_mul(fe_sqr_n(x223, 23), x22);
    t = fe_mul(fe_sqr_n(t, 5), a);
    t = fe_mul(fe_sqr_n(t, 3), x2);
    t = fe_mul(fe_sqr_n(t, 2), a);
    return t;
}

// candidate square root a^((p+1)/4); caller checks r^2 == a
BPPP_HD Fe fe_sqrt_candidate(const Fe &a) {
    Fe x2 = fe_mul(fe_sqr(a), a);
    Fe x3 = fe_mul(fe_sqr(x2), a);
    Fe x6 = fe_mul(fe_sqr_n(x3, 3), x3);
    Fe x9 = fe_mul(fe_sqr_n(x6, 3), x3);
    Fe x11 = fe_mul(fe_sqr_n(x9, 2), x2);
    Fe x22 = fe_mul(fe_sqr_n(x11, 11), x11);
    Fe x44 = fe_mul(fe_sqr_n(x22, 22), x22);
    Fe x88 = fe_mul(fe_sqr_n(x44, 44), x44);
    Fe x176 = fe_mul(fe_sqr_n(x88, 88), x88);
    Fe x220 = fe_mul(fe_sqr_n(x176, 44), x44);
    Fe x223 = fe_mul(fe_sqr_n(x220, 3), x3);
    Fe t = fe_mul(fe_sqr_n(x223, 23), x22);
    t = fe_mul(fe_sqr_n(t, 6), x2);
    t = fe_sqr_n(t, 2);
    return t;
}

// 32-byte big-endian <-> words helpers (byte order conversions only)
BPPP_HD void be32_to_words(uint32_t w[8], const uint8_t *b) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t *p = b + 4 * (7 - i);
        w[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
}
BPPP_HD void words_to_be32(uint8_t *b, const uint32_t w[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint8_t *p = b + 4 * (7 - i);
        p[0] = (uint8_t)(w[i] >> 24); p[1] = (uint8_t)(w[i] >> 16); p[2] = (uint8_t)(w[i] >> 8); p[3] = (uint8_t)w[i];
    }
}
// words (LE) >= p ?
BPPP_HD bool words_ge_p(const uint32_t w[8]) {
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 2; i < 8; i++) m &= w[i];
    if (m != 0xFFFFFFFFu) return false;
    if (w[1] != 0xFFFFFFFEu) return w[1] == 0xFFFFFFFFu;
    return w[0] >= 0xFFFFFC2Fu;
}

}  // namespace bppp
