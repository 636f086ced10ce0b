// secp256k1 base field F_p, p = 2^256 - 2^32 - 977, for sm_100a.
//
// Representation: 8 saturated 32-bit limbs, little-endian, value = sum v[i] * 2^(32 i) < 2^256 -- ANY representative
// mod p in that range (p .. 2^256-1 is allowed); only fe_normalize produces the canonical one.  Every operation returns
// a value below 2^256, so there is no magnitude bookkeeping: the `m` arguments of fe_negate / fe_sub are kept for source
// compatibility and ignored.
//
// Multiplication is the 8x8 schoolbook product on IMAD.WIDE.U32 with hardware carry chains (PTX mad.lo.cc / madc.hi.cc
// pairs, which ptxas fuses into one IMAD.WIDE.U32[.X] with a carry predicate): the products of one row that start at
// even limb positions tile an accumulator without overlap, the odd ones a second accumulator, so a row is two
// independent carry chains of four wide MADs.  64 wide MADs + 9 for the reduction (2^256 = 2^32 + 977 mod p) and ~65
// adds: 138 SASS instructions per fe_mul, against 298 (113 wide MADs) for the earlier 10 x 26-bit lazy representation;
// measured 82-105 G fe_mul/s on a B200 against 46.6 G (profiles/r1_fe_representation.txt).
//
// Host builds (tests/hostemu) compile the portable branches below; they mirror the device steps one to one (same folds,
// same rare-carry handling) and assert the bounds each step relies on when BPPP_VERIFY_MAG is defined.
//
// This replaces k256::FieldElement (not in /root/reference; Cargo.lock:411-414) on the device.
// The internal representation is unobservable; only canonical bytes leave the device.
#pragma once
#include <stdint.h>
#include "modinv.cuh"

#if defined(__CUDACC__)
#define BPPP_HD __host__ __device__ __forceinline__
#define BPPP_D __device__ __forceinline__
#else
#define BPPP_HD inline
#define BPPP_D inline
#endif

#if defined(BPPP_VERIFY_MAG)
#include <assert.h>
#define BPPP_ASSERT(c) assert(c)
#else
#define BPPP_ASSERT(c) ((void)0)
#endif
// magnitude bookkeeping of the earlier lazy representation: nothing to track any more
#define BPPP_SET_MAG(r, m) ((void)0)
#define BPPP_SET_LIM(r, l, l9) ((void)0)

// A translation unit may define BPPP_FE_NOINLINE: fe_mul / fe_sqr then compile to real device functions taking
// their operands BY VALUE (nvcc passes the 8-word structs in registers, no stack traffic).  Fully inlined point
// formulas thrash the instruction cache (ncu on k_v_var2: stall_no_instruction 5.4 warps per issue); with calls
// the hot loop is a few KB.

namespace bppp {

static constexpr int FE_W = 8;          // 32-bit words per field element (registers and workspace)
static constexpr int PT_W = 3 * FE_W;   // projective / Jacobian point

struct Fe {
    uint32_t v[8];
};

static constexpr uint32_t FE_C0 = 977u;           // 2^256 = 2^32 + 977 (mod p)
static constexpr uint32_t FE_P0 = 0xFFFFFC2Fu;    // p = {P0, P1, ~0 x 6}
static constexpr uint32_t FE_P1 = 0xFFFFFFFEu;

BPPP_HD void fe_check(const Fe &) {}

BPPP_HD Fe fe_zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
BPPP_HD Fe fe_from_u32(uint32_t x) {
    Fe r = fe_zero();
    r.v[0] = x;
    return r;
}
BPPP_HD Fe fe_one() { return fe_from_u32(1); }

// 8 x u32 little-endian words (any 256-bit integer) -> Fe
BPPP_HD Fe fe_from_words(const uint32_t w[8]) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    return r;
}
BPPP_HD Fe fe_normalize_weak(const Fe &a) { return a; }

// r = a + b * (2^32 + 977) for b in {0, 1}; returns the carry out of 2^256
BPPP_HD uint32_t fe_add_fold(Fe &r, uint32_t b) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %9, 977, %0;\n\t addc.cc.u32 %1, %1, %9;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.cc.u32 %7, %7, 0;\n\t addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c)
        : "r"(b));
#else
    uint64_t t = (uint64_t)r.v[0] + (uint64_t)b * FE_C0; r.v[0] = (uint32_t)t;
    t = (t >> 32) + r.v[1] + b; r.v[1] = (uint32_t)t;
    for (int i = 2; i < 8; i++) { t = (t >> 32) + r.v[i]; r.v[i] = (uint32_t)t; }
    c = (uint32_t)(t >> 32);
#endif
    return c;
}
// after a wrap past 2^256 the residue is below 2^68: adding 2^32 + 977 once more cannot carry past limb 2
BPPP_HD void fe_add_fold_low(Fe &r, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %3, 977, %0;\n\t addc.cc.u32 %1, %1, %3;\n\t addc.u32 %2, %2, 0;" : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]) : "r"(b));
#else
    if (b) { BPPP_ASSERT((r.v[3] | r.v[4] | r.v[5] | r.v[6] | r.v[7]) == 0 && r.v[2] < 16); }
    uint64_t t = (uint64_t)r.v[0] + (uint64_t)b * FE_C0; r.v[0] = (uint32_t)t;
    t = (t >> 32) + r.v[1] + b; r.v[1] = (uint32_t)t;
    t = (t >> 32) + r.v[2]; r.v[2] = (uint32_t)t;
    BPPP_ASSERT((t >> 32) == 0);
#endif
}


// r += k * (2^32 + 977) over all eight limbs, returns the carry out of 2^256 (k * 977 < 2^32)
BPPP_HD uint32_t fe_add_fold_k(Fe &r, uint32_t k) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %9, 977, %0;\n\t addc.cc.u32 %1, %1, %9;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.cc.u32 %7, %7, 0;\n\t addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c)
        : "r"(k));
#else
    uint64_t t = (uint64_t)r.v[0] + (uint64_t)k * FE_C0; r.v[0] = (uint32_t)t;
    t = (t >> 32) + r.v[1] + k; r.v[1] = (uint32_t)t;
    for (int i = 2; i < 8; i++) { t = (t >> 32) + r.v[i]; r.v[i] = (uint32_t)t; }
    c = (uint32_t)(t >> 32);
#endif
    return c;
}
// r -= k * (2^32 + 977), k in {0, 1}, straight-line: full borrow chain, then the second subtraction a wrap needs
BPPP_HD void fe_sub_fold_full(Fe &r, uint32_t k) {
#if defined(__CUDA_ARCH__)
    uint32_t k0 = k * FE_C0, bw2;
    asm volatile("sub.cc.u32 %0, %0, %9;\n\t subc.cc.u32 %1, %1, %10;\n\t subc.cc.u32 %2, %2, 0;\n\t subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t subc.cc.u32 %5, %5, 0;\n\t subc.cc.u32 %6, %6, 0;\n\t subc.cc.u32 %7, %7, 0;\n\t subc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(bw2)
        : "r"(k0), "r"(k));
    bw2 &= 1u;
    uint32_t k2 = bw2 * FE_C0;
    asm volatile("sub.cc.u32 %0, %0, %3;\n\t subc.cc.u32 %1, %1, %4;\n\t subc.u32 %2, %2, 0;" : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]) : "r"(k2), "r"(bw2));
#else
    int64_t t = (int64_t)r.v[0] - (int64_t)(k * FE_C0); r.v[0] = (uint32_t)t;
    t = (int64_t)r.v[1] - k + (t >> 32); r.v[1] = (uint32_t)t;
    for (int i = 2; i < 8; i++) { t = (int64_t)r.v[i] + (t >> 32); r.v[i] = (uint32_t)t; }
    uint32_t bw2 = (uint32_t)((t >> 32) & 1);
    if (bw2) { BPPP_ASSERT((r.v[3] & r.v[4] & r.v[5] & r.v[6] & r.v[7]) == 0xFFFFFFFFu && r.v[2] >= 0xFFFFFFF0u); }
    t = (int64_t)r.v[0] - (int64_t)(bw2 * FE_C0); r.v[0] = (uint32_t)t;
    t = (int64_t)r.v[1] - bw2 + (t >> 32); r.v[1] = (uint32_t)t;
    t = (int64_t)r.v[2] + (t >> 32); r.v[2] = (uint32_t)t;
    BPPP_ASSERT((t >> 32) == 0);
#endif
}

#if defined(__CUDACC__)
#define BPPP_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define BPPP_HD_NOINLINE static inline
#endif
// Out-of-line tails of the short folds below (taken with probability ~2^-32 per operation).
// A carry of 1 enters limb 3; if it runs off the top the value wrapped past 2^256 and is tiny: add 2^32 + 977 once more.
BPPP_HD_NOINLINE Fe fe_carry_slow(Fe r) {
    uint64_t t = 1;
    for (int i = 3; i < 8; i++) { t += r.v[i]; r.v[i] = (uint32_t)t; t >>= 32; }
    fe_add_fold_low(r, (uint32_t)t);
    return r;
}
// A borrow of 1 enters limb 3; if it runs off the top the value was below 2^32 + 977 before the subtraction and is now
// just under 2^256: subtract 2^32 + 977 once more (only limbs 0..2 change).
BPPP_HD_NOINLINE Fe fe_borrow_slow(Fe r) {
    int64_t t = -1;
    for (int i = 3; i < 8; i++) { t += r.v[i]; r.v[i] = (uint32_t)t; t >>= 32; }
    if (t) {
        BPPP_ASSERT((r.v[3] & r.v[4] & r.v[5] & r.v[6] & r.v[7]) == 0xFFFFFFFFu);
        int64_t u = (int64_t)r.v[0] - FE_C0; r.v[0] = (uint32_t)u;
        u = (int64_t)r.v[1] - 1 + (u >> 32); r.v[1] = (uint32_t)u;
        u = (int64_t)r.v[2] + (u >> 32); r.v[2] = (uint32_t)u;
        BPPP_ASSERT((u >> 32) == 0);
    }
    return r;
}
// r += k (2^32 + 977) for a small k (k * 977 < 2^32).  BR = true: three limbs and a rarely taken branch (best inside
// the call-based field functions); BR = false: straight-line full-width chains (best where several field operations are
// inlined into one block and ptxas interleaves them -- a branch would end the scheduling region).
template <bool BR>
BPPP_HD void fe_fold_add(Fe &r, uint32_t k) {
    if (!BR) {
        uint32_t c2 = fe_add_fold_k(r, k);
        fe_add_fold_low(r, c2);
        return;
    }
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %4, 977, %0;\n\t addc.cc.u32 %1, %1, %4;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;"
                 : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "=r"(c) : "r"(k));
#else
    uint64_t t = (uint64_t)r.v[0] + (uint64_t)k * FE_C0; r.v[0] = (uint32_t)t;
    t = (t >> 32) + r.v[1] + k; r.v[1] = (uint32_t)t;
    t = (t >> 32) + r.v[2]; r.v[2] = (uint32_t)t;
    c = (uint32_t)(t >> 32);
#endif
    if (__builtin_expect(c != 0, 0)) r = fe_carry_slow(r);
}
// r -= k (2^32 + 977) for k in {0, 1}
template <bool BR>
BPPP_HD void fe_fold_sub(Fe &r, uint32_t k) {
    if (!BR) { fe_sub_fold_full(r, k); return; }
    uint32_t b;
#if defined(__CUDA_ARCH__)
    uint32_t k0 = k * FE_C0;
    asm volatile("sub.cc.u32 %0, %0, %4;\n\t subc.cc.u32 %1, %1, %5;\n\t subc.cc.u32 %2, %2, 0;\n\t subc.u32 %3, 0, 0;"
                 : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "=r"(b) : "r"(k0), "r"(k));
#else
    int64_t t = (int64_t)r.v[0] - (int64_t)(k * FE_C0); r.v[0] = (uint32_t)t;
    t = (int64_t)r.v[1] - k + (t >> 32); r.v[1] = (uint32_t)t;
    t = (int64_t)r.v[2] + (t >> 32); r.v[2] = (uint32_t)t;
    b = (uint32_t)((t >> 32) & 1);
#endif
    if (__builtin_expect(b != 0, 0)) r = fe_borrow_slow(r);
}

// canonical representative in [0, p):  v >= p  <=>  v + (2^32 + 977) carries out of 2^256, and then v - p is that sum
BPPP_HD Fe fe_normalize(const Fe &a) {
    Fe t = a;
    uint32_t c = fe_add_fold(t, 1u);
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? t.v[i] : a.v[i];
    return r;
}

// canonical Fe -> 8 x u32 little-endian words
BPPP_HD void fe_to_words(uint32_t w[8], const Fe &a_canonical) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = a_canonical.v[i];
}

BPPP_HD bool fe_is_zero_canonical(const Fe &a) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) m |= a.v[i];
    return m == 0;
}
// value == 0 (mod p): the only multiples of p below 2^256 are 0 and p
BPPP_HD bool fe_normalizes_to_zero(const Fe &a) {
    uint32_t z = a.v[0] | a.v[1], pm = (a.v[0] ^ FE_P0) | (a.v[1] ^ FE_P1);
#pragma unroll
    for (int i = 2; i < 8; i++) { z |= a.v[i]; pm |= ~a.v[i]; }
    return z == 0 || pm == 0;
}
BPPP_HD bool fe_is_zero(const Fe &a) { return fe_normalizes_to_zero(a); }
BPPP_HD bool fe_equal_canonical(const Fe &a, const Fe &b) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) m |= a.v[i] ^ b.v[i];
    return m == 0;
}
BPPP_HD bool fe_is_odd_canonical(const Fe &a) { return a.v[0] & 1u; }

template <bool BR>
BPPP_HD Fe fe_add_t(const Fe &a, const Fe &b) {
    Fe r;
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm volatile("add.cc.u32 %0, %9, %17;\n\t addc.cc.u32 %1, %10, %18;\n\t addc.cc.u32 %2, %11, %19;\n\t addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t addc.cc.u32 %5, %14, %22;\n\t addc.cc.u32 %6, %15, %23;\n\t addc.cc.u32 %7, %16, %24;\n\t addc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) { t = (t >> 32) + a.v[i] + b.v[i]; r.v[i] = (uint32_t)t; }
    c = (uint32_t)(t >> 32);
#endif
    fe_fold_add<BR>(r, c);    // carried: a + b - 2^256 + (2^32 + 977)
    return r;
}
BPPP_HD Fe fe_add(const Fe &a, const Fe &b) { return fe_add_t<true>(a, b); }

// a - b (the magnitude argument of the lazy representation is ignored)
template <bool BR>
BPPP_HD Fe fe_sub_t(const Fe &a, const Fe &b) {
    Fe r;
    uint32_t bw;   // 0 or 1
#if defined(__CUDA_ARCH__)
    asm volatile("sub.cc.u32 %0, %9, %17;\n\t subc.cc.u32 %1, %10, %18;\n\t subc.cc.u32 %2, %11, %19;\n\t subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t subc.cc.u32 %5, %14, %22;\n\t subc.cc.u32 %6, %15, %23;\n\t subc.cc.u32 %7, %16, %24;\n\t subc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    bw &= 1u;   // subc yields 0 - 0 - borrow = 0 or 0xFFFFFFFF
#else
    int64_t t = 0;
    for (int i = 0; i < 8; i++) { t = (int64_t)a.v[i] - b.v[i] + (t >> 32); r.v[i] = (uint32_t)t; }
    bw = (uint32_t)((t >> 32) & 1);
#endif
    fe_fold_sub<BR>(r, bw);   // borrowed: a - b + 2^256 - (2^32 + 977) = a - b + p
    return r;
}
BPPP_HD Fe fe_sub(const Fe &a, const Fe &b, int = 0) { return fe_sub_t<true>(a, b); }
BPPP_HD Fe fe_negate(const Fe &a, int = 0) { return fe_sub(fe_zero(), a); }

// a * k for a small constant k (k * 977 must fit 32 bits); powers of two are funnel shifts
template <bool BR>
BPPP_HD Fe fe_mul_int_t(const Fe &a, uint32_t k) {
    Fe r;
    uint32_t top;
    if (k == 2 || k == 4 || k == 8) {
        const int sh = k == 2 ? 1 : (k == 4 ? 2 : 3);
        top = a.v[7] >> (32 - sh);
#pragma unroll
        for (int i = 7; i > 0; i--) r.v[i] = (a.v[i] << sh) | (a.v[i - 1] >> (32 - sh));
        r.v[0] = a.v[0] << sh;
        fe_fold_add<BR>(r, top);
        return r;
    }
#if defined(__CUDA_ARCH__)
    asm volatile("mul.lo.u32 %0, %9, %17;\n\t mul.hi.u32 %1, %9, %17;\n\t mul.lo.u32 %2, %11, %17;\n\t mul.hi.u32 %3, %11, %17;\n\t"
        "mul.lo.u32 %4, %13, %17;\n\t mul.hi.u32 %5, %13, %17;\n\t mul.lo.u32 %6, %15, %17;\n\t mul.hi.u32 %7, %15, %17;\n\t"
        "mad.lo.cc.u32 %1, %10, %17, %1;\n\t madc.hi.cc.u32 %2, %10, %17, %2;\n\t madc.lo.cc.u32 %3, %12, %17, %3;\n\t madc.hi.cc.u32 %4, %12, %17, %4;\n\t"
        "madc.lo.cc.u32 %5, %14, %17, %5;\n\t madc.hi.cc.u32 %6, %14, %17, %6;\n\t madc.lo.cc.u32 %7, %16, %17, %7;\n\t madc.hi.u32 %8, %16, %17, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(top)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(k));
#else
    BPPP_ASSERT((uint64_t)k * FE_C0 <= 0xFFFFFFFFull);
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) { t = (t >> 32) + (uint64_t)a.v[i] * k; r.v[i] = (uint32_t)t; }
    top = (uint32_t)(t >> 32);
#endif
    fe_fold_add<BR>(r, top);  // top < k
    return r;
}
BPPP_HD Fe fe_mul_int(const Fe &a, uint32_t k) { return fe_mul_int_t<true>(a, k); }
BPPP_HD Fe fe_cmov(const Fe &a, const Fe &b, bool take_b) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = take_b ? b.v[i] : a.v[i];
    return r;
}

// ---- 512-bit product -> Fe ------------------------------------------------------------------------------------
// T = L + 2^256 H  ->  L + H * (2^32 + 977): first fold leaves a 34-bit overflow V, second fold a possible single
// carry whose residue is below 2^68.
template <bool BR>
BPPP_HD Fe fe_reduce512(uint32_t T[16]) {
    Fe r;
#if defined(__CUDA_ARCH__)
    uint32_t R8, R9;
    // T[0..8) += {T8, T10, T12, T14} * 977 at even positions, carry -> R8
    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "=r"(R8)
        : "r"(T[8]), "r"(T[10]), "r"(T[12]), "r"(T[14]), "r"(FE_C0));
    // T[1..8), R8 += {T9, T11, T13, T15} * 977 at odd positions (R8 <= 1 + 976: no carry out)
    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(R8)
        : "r"(T[9]), "r"(T[11]), "r"(T[13]), "r"(T[15]), "r"(FE_C0));
    // + H << 32
    asm volatile("add.cc.u32 %0, %0, %9;\n\t addc.cc.u32 %1, %1, %10;\n\t addc.cc.u32 %2, %2, %11;\n\t addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t addc.cc.u32 %5, %5, %14;\n\t addc.cc.u32 %6, %6, %15;\n\t addc.cc.u32 %7, %7, %16;\n\t addc.u32 %8, 0, 0;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(R8), "=r"(R9)
        : "r"(T[8]), "r"(T[9]), "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
    // W = V * (2^32 + 977), V = R8 + R9 2^32 < 2^34
    uint32_t W0, W1, W2;
    asm volatile("mul.lo.u32 %0, %3, 977;\n\t mul.hi.u32 %1, %3, 977;\n\t mad.lo.u32 %1, %4, 977, %1;\n\t add.cc.u32 %1, %1, %3;\n\t addc.u32 %2, %4, 0;"
        : "=&r"(W0), "=&r"(W1), "=&r"(W2) : "r"(R8), "r"(R9));
    uint32_t c2;
    if (BR) {
        asm volatile("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, %6;\n\t addc.u32 %3, 0, 0;"
            : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "=r"(c2) : "r"(W0), "r"(W1), "r"(W2));
#pragma unroll
        for (int k = 0; k < 8; k++) r.v[k] = T[k];
        if (__builtin_expect(c2 != 0, 0)) r = fe_carry_slow(r);
    } else {
        asm volatile("add.cc.u32 %0, %0, %9;\n\t addc.cc.u32 %1, %1, %10;\n\t addc.cc.u32 %2, %2, %11;\n\t addc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.cc.u32 %7, %7, 0;\n\t addc.u32 %8, 0, 0;"
            : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "=r"(c2)
            : "r"(W0), "r"(W1), "r"(W2));
#pragma unroll
        for (int k = 0; k < 8; k++) r.v[k] = T[k];
        fe_add_fold_low(r, c2);
    }
#else
    // same steps with 64-bit temporaries
    uint64_t t = 0, R[10];
    for (int k = 0; k < 8; k++) R[k] = T[k];
    R[8] = R[9] = 0;
    uint64_t carry = 0;
    for (int k = 0; k < 8; k++) { t = R[k] + (uint64_t)T[8 + k] * FE_C0 + carry; R[k] = (uint32_t)t; carry = t >> 32; }
    R[8] = carry;                                         // <= 977
    BPPP_ASSERT(R[8] <= 977);
    carry = 0;
    for (int k = 1; k < 9; k++) { t = R[k] + T[7 + k] + carry; R[k] = (uint32_t)t; carry = t >> 32; }
    R[9] = carry;
    BPPP_ASSERT(R[9] <= 1);
    uint64_t V = R[8] | (R[9] << 32);
    uint64_t W0 = (uint32_t)(V * FE_C0), Whi = (V * FE_C0) >> 32;     // V * 977 < 2^44
    uint64_t W1 = Whi + (uint32_t)V, W2 = (V >> 32) + (W1 >> 32);
    W1 = (uint32_t)W1;
    BPPP_ASSERT(W2 < 8);
    t = R[0] + W0; r.v[0] = (uint32_t)t;
    t = (t >> 32) + R[1] + W1; r.v[1] = (uint32_t)t;
    t = (t >> 32) + R[2] + W2; r.v[2] = (uint32_t)t;
    for (int k = 3; k < 8; k++) r.v[k] = (uint32_t)R[k];
    if (t >> 32) r = fe_carry_slow(r);
#endif
    return r;
}

#if defined(__CUDA_ARCH__)
// acc[0..8) += {a0..a3} * b as four adjacent 64-bit products, carry out into a limb that was still zero
BPPP_D void fe_chain4_carry(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b, uint32_t &top) {
    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(top)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// same, but the last product is the first one at its position: its high limb was zero and nothing can carry out
BPPP_D void fe_chain4_top(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b, uint32_t &top) {
    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "=r"(top)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
BPPP_D void fe_mul4(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]), "=&r"(acc[6]), "=&r"(acc[7])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// n adjacent 64-bit products acc[0..2n) += a * {b0..}, carry added into acc[2n] (which may already hold carries)
BPPP_D void fe_chain1(uint32_t *acc, uint32_t a, uint32_t b0) {
    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t addc.u32 %2, %2, 0;" : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(a), "r"(b0));
}
BPPP_D void fe_chain2(uint32_t *acc, uint32_t a, uint32_t b0, uint32_t b1) {
    asm volatile("mad.lo.cc.u32 %0, %5, %6, %0;\n\t madc.hi.cc.u32 %1, %5, %6, %1;\n\t madc.lo.cc.u32 %2, %5, %7, %2;\n\t madc.hi.cc.u32 %3, %5, %7, %3;\n\t addc.u32 %4, %4, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]) : "r"(a), "r"(b0), "r"(b1));
}
BPPP_D void fe_chain3(uint32_t *acc, uint32_t a, uint32_t b0, uint32_t b1, uint32_t b2) {
    asm volatile("mad.lo.cc.u32 %0, %7, %8, %0;\n\t madc.hi.cc.u32 %1, %7, %8, %1;\n\t madc.lo.cc.u32 %2, %7, %9, %2;\n\t madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %7, %10, %4;\n\t madc.hi.cc.u32 %5, %7, %10, %5;\n\t addc.u32 %6, %6, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]) : "r"(a), "r"(b0), "r"(b1), "r"(b2));
}
// T[1..16) = E[1..16) + O[0..15): two single-block chains, the carry handed over in a register
BPPP_D void fe_merge_even_odd(uint32_t *T, const uint32_t *E, const uint32_t *O) {
    uint32_t c;
    asm volatile("add.cc.u32 %0, %9, %17;\n\t addc.cc.u32 %1, %10, %18;\n\t addc.cc.u32 %2, %11, %19;\n\t addc.cc.u32 %3, %12, %20;\n\t"
                 "addc.cc.u32 %4, %13, %21;\n\t addc.cc.u32 %5, %14, %22;\n\t addc.cc.u32 %6, %15, %23;\n\t addc.cc.u32 %7, %16, %24;\n\t addc.u32 %8, 0, 0;"
                 : "=&r"(T[1]), "=&r"(T[2]), "=&r"(T[3]), "=&r"(T[4]), "=&r"(T[5]), "=&r"(T[6]), "=&r"(T[7]), "=&r"(T[8]), "=&r"(c)
                 : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]),
                   "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
    asm volatile("add.cc.u32 %7, %7, 0xFFFFFFFF;\n\t"      // carry flag <- c
                 "addc.cc.u32 %0, %8, %15;\n\t addc.cc.u32 %1, %9, %16;\n\t addc.cc.u32 %2, %10, %17;\n\t addc.cc.u32 %3, %11, %18;\n\t"
                 "addc.cc.u32 %4, %12, %19;\n\t addc.cc.u32 %5, %13, %20;\n\t addc.u32 %6, %14, %21;"
                 : "=&r"(T[9]), "=&r"(T[10]), "=&r"(T[11]), "=&r"(T[12]), "=&r"(T[13]), "=&r"(T[14]), "=&r"(T[15]), "+r"(c)
                 : "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
                   "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
}
#endif

// T[0..16) = a * b (8 x 8 limbs).  Shared with the scalar field (sc.cuh).
BPPP_HD void wide_mul8(uint32_t T[16], const uint32_t a[8], const uint32_t b[8]) {
#if defined(__CUDA_ARCH__)
    // E: products starting at even limb positions (index = position); O: odd positions (index = position - 1)
    uint32_t E[16], O[16];
    E[8] = 0;
    fe_mul4(E, a[0], a[2], a[4], a[6], b[0]);
    fe_mul4(O, a[1], a[3], a[5], a[7], b[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        if (i & 1) {
            fe_chain4_carry(O + i - 1, a[0], a[2], a[4], a[6], b[i], O[i + 7]);
            fe_chain4_top(E + i + 1, a[1], a[3], a[5], a[7], b[i], E[i + 8]);
        } else {
            fe_chain4_carry(E + i, a[0], a[2], a[4], a[6], b[i], E[i + 8]);
            fe_chain4_top(O + i, a[1], a[3], a[5], a[7], b[i], O[i + 7]);
        }
    }
    T[0] = E[0];
    fe_merge_even_odd(T, E, O);
#else
    for (int k = 0; k < 16; k++) T[k] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) {
            uint64_t t = (uint64_t)T[i + j] + (uint64_t)a[j] * b[i] + carry;
            T[i + j] = (uint32_t)t; carry = t >> 32;
        }
        T[i + 8] = (uint32_t)carry;
    }
#endif
}
// T[0..16) = a^2
BPPP_HD void wide_sqr8(uint32_t T[16], const uint32_t v[8]) {
#if defined(__CUDA_ARCH__)
    // off-diagonal products a_i a_j (i < j) once, into the even / odd position accumulators
    uint32_t E[16], O[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { E[k] = 0; O[k] = 0; }
    // row 0
    fe_mul4(O, v[1], v[3], v[5], v[7], v[0]);                            // positions 1 3 5 7
    asm volatile("mul.lo.u32 %0, %6, %9;\n\t mul.hi.u32 %1, %6, %9;\n\t mul.lo.u32 %2, %7, %9;\n\t mul.hi.u32 %3, %7, %9;\n\t mul.lo.u32 %4, %8, %9;\n\t mul.hi.u32 %5, %8, %9;"
        : "=&r"(E[2]), "=&r"(E[3]), "=&r"(E[4]), "=&r"(E[5]), "=&r"(E[6]), "=&r"(E[7]) : "r"(v[2]), "r"(v[4]), "r"(v[6]), "r"(v[0]));   // positions 2 4 6
    // row 1
    fe_chain3(O + 2, v[1], v[2], v[4], v[6]);                            // 3 5 7
    fe_chain3(E + 4, v[1], v[3], v[5], v[7]);                            // 4 6 8
    // row 2
    fe_chain3(O + 4, v[2], v[3], v[5], v[7]);                            // 5 7 9
    fe_chain2(E + 6, v[2], v[4], v[6]);                                  // 6 8
    // row 3
    fe_chain2(O + 6, v[3], v[4], v[6]);                                  // 7 9
    fe_chain2(E + 8, v[3], v[5], v[7]);                                  // 8 10
    // row 4
    fe_chain2(O + 8, v[4], v[5], v[7]);                                  // 9 11
    fe_chain1(E + 10, v[4], v[6]);                                       // 10
    // row 5
    fe_chain1(O + 10, v[5], v[6]);                                       // 11
    fe_chain1(E + 12, v[5], v[7]);                                       // 12
    // row 6
    fe_chain1(O + 12, v[6], v[7]);                                       // 13
    // S = E + (O << 32), then T = 2 S + sum a_i^2 2^(64 i)
    T[0] = 0;
    fe_merge_even_odd(T, E, O);
    // T = 2 T (S < 2^511: nothing shifts out)
    asm volatile("add.cc.u32 %0, %0, %0;\n\t addc.cc.u32 %1, %1, %1;\n\t addc.cc.u32 %2, %2, %2;\n\t addc.cc.u32 %3, %3, %3;\n\t addc.cc.u32 %4, %4, %4;\n\t"
                 "addc.cc.u32 %5, %5, %5;\n\t addc.cc.u32 %6, %6, %6;\n\t addc.cc.u32 %7, %7, %7;\n\t addc.cc.u32 %8, %8, %8;\n\t addc.cc.u32 %9, %9, %9;\n\t"
                 "addc.cc.u32 %10, %10, %10;\n\t addc.cc.u32 %11, %11, %11;\n\t addc.cc.u32 %12, %12, %12;\n\t addc.cc.u32 %13, %13, %13;\n\t addc.u32 %14, %14, %14;"
                 : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]),
                   "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15]));
    // T += sum a_k^2 2^(64 k): one chain over all 16 limbs
    asm volatile("mad.lo.cc.u32 %0, %16, %16, %0;\n\t madc.hi.cc.u32 %1, %16, %16, %1;\n\t madc.lo.cc.u32 %2, %17, %17, %2;\n\t madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
                 "madc.lo.cc.u32 %4, %18, %18, %4;\n\t madc.hi.cc.u32 %5, %18, %18, %5;\n\t madc.lo.cc.u32 %6, %19, %19, %6;\n\t madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
                 "madc.lo.cc.u32 %8, %20, %20, %8;\n\t madc.hi.cc.u32 %9, %20, %20, %9;\n\t madc.lo.cc.u32 %10, %21, %21, %10;\n\t madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
                 "madc.lo.cc.u32 %12, %22, %22, %12;\n\t madc.hi.cc.u32 %13, %22, %22, %13;\n\t madc.lo.cc.u32 %14, %23, %23, %14;\n\t madc.hi.u32 %15, %23, %23, %15;"
                 : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]),
                   "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
                 : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
#else
    wide_mul8(T, v, v);
#endif
}

template <bool BR>
BPPP_HD Fe fe_mul_inl_t(const Fe &a, const Fe &b) {
    uint32_t T[16];
    wide_mul8(T, a.v, b.v);
    return fe_reduce512<BR>(T);
}
template <bool BR>
BPPP_HD Fe fe_sqr_inl_t(const Fe &a) {
    uint32_t T[16];
    wide_sqr8(T, a.v);
    return fe_reduce512<BR>(T);
}
BPPP_HD Fe fe_mul_inl(const Fe &a, const Fe &b) { return fe_mul_inl_t<true>(a, b); }
BPPP_HD Fe fe_sqr_inl(const Fe &a) { return fe_sqr_inl_t<true>(a); }

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

#if defined(__CUDACC__) && defined(BPPP_FE_NOINLINE)
// the long squaring runs of fe_inv / fe_sqrt_candidate as one call: no per-squaring argument marshalling
static __device__ __noinline__ Fe fe_sqr_n_call(Fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe_sqr_inl_t<false>(a);
    return a;
}
#endif
BPPP_HD Fe fe_sqr_n(Fe a, int n) {
#if defined(__CUDA_ARCH__) && defined(BPPP_FE_NOINLINE)
    return fe_sqr_n_call(a, n);
#else
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe_sqr(a);
    return a;
#endif
}

// a^(p-2): 255 squarings + 15 multiplications.  fe_inv_fermat(0) = 0.  Kept as the cross-check of fe_inv (tests/hostemu).
BPPP_HD Fe fe_inv_fermat(const Fe &a) {
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

// a^-1 by safegcd division steps (modinv.cuh): ~14 k instructions against ~36 k for the power.  fe_inv(0) = 0.
#if defined(__CUDACC__)
static __device__ __noinline__ Fe fe_inv_gcd_call(Fe a) {
    Fe c = fe_normalize(a), r;
    mi_modinv_words<MIModP>(r.v, c.v);
    return r;
}
#endif
BPPP_HD Fe fe_inv(const Fe &a) {
#if defined(BPPP_INV_FERMAT)
    return fe_inv_fermat(a);
#elif defined(__CUDA_ARCH__)
    return fe_inv_gcd_call(a);          // one copy per translation unit: every caller is a cold, once-per-thread site
#else
    Fe c = fe_normalize(a), r;
    mi_modinv_words<MIModP>(r.v, c.v);
    return r;
#endif
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
