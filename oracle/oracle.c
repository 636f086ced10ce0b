/*
 * CPU oracle #2 (plain C) -- TEST INFRASTRUCTURE AND CPU BASELINE ONLY.
 *
 * A restatement in C of the reference's algorithm (distributed-lab/bp-pp 0.1.1)
 * executed the way the reference executes it: one full scalar multiplication per
 * MSM term (src/util.rs:46-60), dense circuit matrices (src/circuit.rs:584-653),
 * generator folding and full re-commit in every WNLA round (src/wnla.rs:170-186).
 * Each function cites the reference file:line it follows (relative to
 * /root/reference).  The curve/field/transcript arithmetic of the un-vendored
 * crates k256 0.13.3 and merlin 3.0.0 is restated from their published
 * algorithms.
 *
 * PARITY STATUS: parity unpinned against real k256 (see oracle/bppp_ref.py header):
 * pinned against the Python-int oracle, OpenSSL and the Merlin conformance vector.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Nothing under bp_pp_b200/ links or calls it.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint8_t u8;

/* ------------------------------------------------------------------ */
/* 256-bit helpers                                                      */
/* ------------------------------------------------------------------ */
typedef struct { u64 v[4]; } u256; /* little-endian limbs */

static const u256 FP = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const u256 FN = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
/* 2^256 - n */
static const u64 NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 0x1ULL};
#define PC 0x1000003D1ULL /* 2^256 - p */

static int u256_is_zero(const u256 *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int u256_eq(const u256 *a, const u256 *b) {
    return ((a->v[0] ^ b->v[0]) | (a->v[1] ^ b->v[1]) | (a->v[2] ^ b->v[2]) | (a->v[3] ^ b->v[3])) == 0;
}
static int u256_geq(const u256 *a, const u256 *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->v[i] > b->v[i]) return 1;
        if (a->v[i] < b->v[i]) return 0;
    }
    return 1;
}
static u64 u256_add(u256 *r, const u256 *a, const u256 *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static u64 u256_sub(u256 *r, const u256 *a, const u256 *b) {
    u64 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->v[i] - b->v[i] - borrow;
        r->v[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static void u256_from_be(u256 *r, const u8 *b) {
    for (int i = 0; i < 4; i++) {
        u64 w = 0;
        for (int j = 0; j < 8; j++) w = (w << 8) | b[8 * (3 - i) + j];
        r->v[i] = w;
    }
}
static void u256_to_be(u8 *b, const u256 *a) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (u8)(a->v[i] >> (56 - 8 * j));
}
static void mul_wide(u64 t[8], const u256 *a, const u256 *b) {
    memset(t, 0, 8 * sizeof(u64));
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->v[i] * b->v[j] + t[i + j];
            t[i + j] = (u64)c; c >>= 64;
        }
        t[i + 4] = (u64)c;
    }
}

/* ------------------------------------------------------------------ */
/* base field mod p (k256 FieldElement): always canonical here          */
/* ------------------------------------------------------------------ */
typedef u256 fe;
/* Elements are kept weakly reduced (any representative below 2^256); fe_canon gives the canonical one.
 * 2^256 = PC (mod p), so a carry/borrow out of the top limb is folded back as +/- PC. */
static inline void fe_canon(fe *r) { if (u256_geq(r, &FP)) u256_sub(r, r, &FP); }
static inline void fe_add_small(fe *r, u64 x) { /* r += x, folding carries */
    while (x) {
        u128 c = (u128)r->v[0] + x; r->v[0] = (u64)c; c >>= 64;
        c += r->v[1]; r->v[1] = (u64)c; c >>= 64;
        c += r->v[2]; r->v[2] = (u64)c; c >>= 64;
        c += r->v[3]; r->v[3] = (u64)c; c >>= 64;
        x = (u64)c * PC;
    }
}
static inline void fe_norm(fe *r, u64 carry) { if (carry) fe_add_small(r, carry * PC); fe_canon(r); }
static inline u64 addc(u64 a, u64 b, u64 cin, u64 *cout) { u64 s = a + b; u64 c1 = s < a; u64 t = s + cin; *cout = c1 | (t < s); return t; }
static inline u64 subb(u64 a, u64 b, u64 bin, u64 *bout) { u64 d = a - b; u64 b1 = a < b; u64 t = d - bin; *bout = b1 | (d < bin); return t; }
static inline void fe_add(fe *r, const fe *a, const fe *b) {
    u64 c, r0 = addc(a->v[0], b->v[0], 0, &c), r1 = addc(a->v[1], b->v[1], c, &c), r2 = addc(a->v[2], b->v[2], c, &c), r3 = addc(a->v[3], b->v[3], c, &c);
    /* fold the carry: + PC; a second carry can only happen when the sum wrapped to < PC, so one more fold terminates */
    u64 k = (0 - c) & PC;
    r0 = addc(r0, k, 0, &c); r1 = addc(r1, 0, c, &c); r2 = addc(r2, 0, c, &c); r3 = addc(r3, 0, c, &c);
    r0 += (0 - c) & PC;
    r->v[0] = r0; r->v[1] = r1; r->v[2] = r2; r->v[3] = r3;
}
static inline void fe_sub(fe *r, const fe *a, const fe *b) {
    u64 bw, r0 = subb(a->v[0], b->v[0], 0, &bw), r1 = subb(a->v[1], b->v[1], bw, &bw), r2 = subb(a->v[2], b->v[2], bw, &bw), r3 = subb(a->v[3], b->v[3], bw, &bw);
    /* a - b + 2^256 = a - b + PC (mod p): take PC off again; a second borrow means the value was < PC before, so once more */
    u64 k = (0 - bw) & PC;
    r0 = subb(r0, k, 0, &bw); r1 = subb(r1, 0, bw, &bw); r2 = subb(r2, 0, bw, &bw); r3 = subb(r3, 0, bw, &bw);
    r0 -= (0 - bw) & PC;
    r->v[0] = r0; r->v[1] = r1; r->v[2] = r2; r->v[3] = r3;
}
static inline void fe_neg(fe *r, const fe *a) { fe z = {{0, 0, 0, 0}}; fe_sub(r, &z, a); }
static inline void fe_mul(fe *r, const fe *a, const fe *b) {
    const u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3];
    u64 t[8]; u128 c;
    c = (u128)a0 * b->v[0]; t[0] = (u64)c; c >>= 64;
    c += (u128)a0 * b->v[1]; t[1] = (u64)c; c >>= 64;
    c += (u128)a0 * b->v[2]; t[2] = (u64)c; c >>= 64;
    c += (u128)a0 * b->v[3]; t[3] = (u64)c; t[4] = (u64)(c >> 64);
    c = (u128)a1 * b->v[0] + t[1]; t[1] = (u64)c; c >>= 64;
    c += (u128)a1 * b->v[1] + t[2]; t[2] = (u64)c; c >>= 64;
    c += (u128)a1 * b->v[2] + t[3]; t[3] = (u64)c; c >>= 64;
    c += (u128)a1 * b->v[3] + t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
    c = (u128)a2 * b->v[0] + t[2]; t[2] = (u64)c; c >>= 64;
    c += (u128)a2 * b->v[1] + t[3]; t[3] = (u64)c; c >>= 64;
    c += (u128)a2 * b->v[2] + t[4]; t[4] = (u64)c; c >>= 64;
    c += (u128)a2 * b->v[3] + t[5]; t[5] = (u64)c; t[6] = (u64)(c >> 64);
    c = (u128)a3 * b->v[0] + t[3]; t[3] = (u64)c; c >>= 64;
    c += (u128)a3 * b->v[1] + t[4]; t[4] = (u64)c; c >>= 64;
    c += (u128)a3 * b->v[2] + t[5]; t[5] = (u64)c; c >>= 64;
    c += (u128)a3 * b->v[3] + t[6]; t[6] = (u64)c; t[7] = (u64)(c >> 64);
    /* fold the high half: hi * (2^32 + 977) + lo */
    c = (u128)t[4] * PC + t[0]; r->v[0] = (u64)c; c >>= 64;
    c += (u128)t[5] * PC + t[1]; r->v[1] = (u64)c; c >>= 64;
    c += (u128)t[6] * PC + t[2]; r->v[2] = (u64)c; c >>= 64;
    c += (u128)t[7] * PC + t[3]; r->v[3] = (u64)c; c >>= 64;
    if (c) {                                   /* c < 2^34: c * PC < 2^67 spans two limbs */
        u128 x = c * PC;
        u128 d = (u128)r->v[0] + (u64)x; r->v[0] = (u64)d; d >>= 64;
        d += (u128)r->v[1] + (u64)(x >> 64); r->v[1] = (u64)d; d >>= 64;
        d += r->v[2]; r->v[2] = (u64)d; d >>= 64;
        d += r->v[3]; r->v[3] = (u64)d; d >>= 64;
        if (d) fe_add_small(r, PC);
    }
}
static inline void fe_sqr(fe *r, const fe *a) { fe_mul(r, a, a); }
static inline void fe_mul_small(fe *r, const fe *a, u64 k) {
    u128 c = (u128)a->v[0] * k; r->v[0] = (u64)c; c >>= 64;
    c += (u128)a->v[1] * k; r->v[1] = (u64)c; c >>= 64;
    c += (u128)a->v[2] * k; r->v[2] = (u64)c; c >>= 64;
    c += (u128)a->v[3] * k; r->v[3] = (u64)c; c >>= 64;
    if (c) fe_add_small(r, (u64)c * PC);
}
static void fe_pow(fe *r, const fe *a, const u256 *e) {
    fe acc = {{1, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) {
        fe_sqr(&acc, &acc);
        if ((e->v[i / 64] >> (i % 64)) & 1) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fe_inv(fe *r, const fe *a) {
    u256 e = FP; e.v[0] -= 2;
    fe_pow(r, a, &e);
}
static int fe_sqrt(fe *r, const fe *a) {
    /* p = 3 mod 4: a^((p+1)/4) */
    u256 e = {{0xFFFFFFFFBFFFFF0CULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFFFFULL}};
    fe s, chk;
    fe_pow(&s, a, &e);
    fe_sqr(&chk, &s);
    fe_canon(&s); fe_canon(&chk);
    *r = s;
    fe ac = *a; fe_canon(&ac);
    return u256_eq(&chk, &ac);
}

/* ------------------------------------------------------------------ */
/* scalar field mod n (k256 Scalar): always canonical                   */
/* ------------------------------------------------------------------ */
typedef u256 sc;
static const sc SC_ZERO = {{0, 0, 0, 0}};
static const sc SC_ONE = {{1, 0, 0, 0}};
static void sc_add(sc *r, const sc *a, const sc *b) {
    u64 c = u256_add(r, a, b);
    if (c || u256_geq(r, &FN)) u256_sub(r, r, &FN);
}
static void sc_sub(sc *r, const sc *a, const sc *b) { if (u256_sub(r, a, b)) u256_add(r, r, &FN); }
static void sc_neg(sc *r, const sc *a) { sc_sub(r, &SC_ZERO, a); }
/* reduce a little-endian nlimbs-limb integer mod n */
static void sc_reduce_wide(sc *r, const u64 *t, int nlimbs) {
    u64 cur[9]; int n = nlimbs;
    memset(cur, 0, sizeof cur);
    memcpy(cur, t, nlimbs * sizeof(u64));
    while (n > 4) {
        int any = 0;
        for (int i = 4; i < n; i++) any |= cur[i] != 0;
        if (!any) break;
        /* cur = lo + hi * NC */
        u64 nxt[9]; memset(nxt, 0, sizeof nxt);
        memcpy(nxt, cur, 4 * sizeof(u64));
        int hn = n - 4;
        for (int i = 0; i < hn; i++) {
            u128 c = 0;
            for (int j = 0; j < 3; j++) {
                c += (u128)cur[4 + i] * NC[j] + nxt[i + j];
                nxt[i + j] = (u64)c; c >>= 64;
            }
            for (int k = i + 3; c && k < 9; k++) { c += nxt[k]; nxt[k] = (u64)c; c >>= 64; }
        }
        memcpy(cur, nxt, sizeof cur);
        n = 9;
        while (n > 4 && cur[n - 1] == 0) n--;
    }
    sc x = {{cur[0], cur[1], cur[2], cur[3]}};
    while (u256_geq(&x, &FN)) u256_sub(&x, &x, &FN);
    *r = x;
}
static inline void mul_wide4(u64 t[8], const u256 *a, const u256 *b) {
    const u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3];
    u128 c;
    c = (u128)a0 * b->v[0]; t[0] = (u64)c; c >>= 64;
    c += (u128)a0 * b->v[1]; t[1] = (u64)c; c >>= 64;
    c += (u128)a0 * b->v[2]; t[2] = (u64)c; c >>= 64;
    c += (u128)a0 * b->v[3]; t[3] = (u64)c; t[4] = (u64)(c >> 64);
    c = (u128)a1 * b->v[0] + t[1]; t[1] = (u64)c; c >>= 64;
    c += (u128)a1 * b->v[1] + t[2]; t[2] = (u64)c; c >>= 64;
    c += (u128)a1 * b->v[2] + t[3]; t[3] = (u64)c; c >>= 64;
    c += (u128)a1 * b->v[3] + t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
    c = (u128)a2 * b->v[0] + t[2]; t[2] = (u64)c; c >>= 64;
    c += (u128)a2 * b->v[1] + t[3]; t[3] = (u64)c; c >>= 64;
    c += (u128)a2 * b->v[2] + t[4]; t[4] = (u64)c; c >>= 64;
    c += (u128)a2 * b->v[3] + t[5]; t[5] = (u64)c; t[6] = (u64)(c >> 64);
    c = (u128)a3 * b->v[0] + t[3]; t[3] = (u64)c; c >>= 64;
    c += (u128)a3 * b->v[1] + t[4]; t[4] = (u64)c; c >>= 64;
    c += (u128)a3 * b->v[2] + t[5]; t[5] = (u64)c; c >>= 64;
    c += (u128)a3 * b->v[3] + t[6]; t[6] = (u64)c; t[7] = (u64)(c >> 64);
}
/* 512 -> 256 bits mod n with 2^256 = NC (mod n), NC = NC[0] + NC[1] 2^64 + 2^128 */
static void sc_mul(sc *r, const sc *a, const sc *b) {
    u64 t[8];
    mul_wide4(t, a, b);
    /* fold 1: lo + hi * NC -> 7 limbs */
    u64 m[7]; u128 c;
    c = (u128)t[4] * NC[0] + t[0]; m[0] = (u64)c; c >>= 64;
    c += (u128)t[5] * NC[0] + t[1]; u64 x1 = (u64)c; c >>= 64;
    c += (u128)t[6] * NC[0] + t[2]; u64 x2 = (u64)c; c >>= 64;
    c += (u128)t[7] * NC[0] + t[3]; u64 x3 = (u64)c; u64 x4 = (u64)(c >> 64);
    c = (u128)t[4] * NC[1] + x1; m[1] = (u64)c; c >>= 64;
    c += (u128)t[5] * NC[1] + x2; x2 = (u64)c; c >>= 64;
    c += (u128)t[6] * NC[1] + x3; x3 = (u64)c; c >>= 64;
    c += (u128)t[7] * NC[1] + x4; x4 = (u64)c; u64 x5 = (u64)(c >> 64);
    c = (u128)t[4] + x2; m[2] = (u64)c; c >>= 64;          /* + hi << 128 */
    c += (u128)t[5] + x3; m[3] = (u64)c; c >>= 64;
    c += (u128)t[6] + x4; m[4] = (u64)c; c >>= 64;
    c += (u128)t[7] + x5; m[5] = (u64)c; m[6] = (u64)(c >> 64);
    sc_reduce_wide(r, m, 7);
}
static void sc_from_u64(sc *r, u64 x) { r->v[0] = x; r->v[1] = r->v[2] = r->v[3] = 0; }
static void sc_pow_u64(sc *r, const sc *a, u64 e) { /* util.rs:97-99 pow_vartime */
    sc acc = SC_ONE;
    for (int i = 63; i >= 0; i--) {
        sc_mul(&acc, &acc, &acc);
        if ((e >> i) & 1) sc_mul(&acc, &acc, a);
    }
    *r = acc;
}
/* Scalar::invert().unwrap() / invert_vartime(): returns 0 (-> panic) on zero.  Variable time, as k256's invert_vartime is.
 * Kaliski's almost-Montgomery inverse: phase 1 is a binary GCD whose cofactors r, s are only ever shifted and added (no
 * halving mod n per bit, which made the earlier binary extended Euclid ~12 us against k256's ~5 us), giving
 * a^-1 2^k mod n with 256 <= k <= 512; phase 2 multiplies by the precomputed 2^-k.  ~3 us on the bench hosts. */
static inline void u256_shr1(u256 *a, u64 top) {
    a->v[0] = (a->v[0] >> 1) | (a->v[1] << 63); a->v[1] = (a->v[1] >> 1) | (a->v[2] << 63);
    a->v[2] = (a->v[2] >> 1) | (a->v[3] << 63); a->v[3] = (a->v[3] >> 1) | (top << 63);
}
static inline void sc_half(sc *x) { /* x / 2 mod n */
    if (x->v[0] & 1) { u64 c = u256_add(x, x, &FN); u256_shr1(x, c); } else u256_shr1(x, 0);
}
typedef struct { u64 v[5]; } u320;       /* cofactors stay below 2 n < 2^257 */
static inline void u320_shl1(u320 *a) {
    a->v[4] = (a->v[4] << 1) | (a->v[3] >> 63); a->v[3] = (a->v[3] << 1) | (a->v[2] >> 63);
    a->v[2] = (a->v[2] << 1) | (a->v[1] >> 63); a->v[1] = (a->v[1] << 1) | (a->v[0] >> 63); a->v[0] <<= 1;
}
static inline void u320_add(u320 *r, const u320 *a, const u320 *b) {
    u128 c = 0;
    for (int i = 0; i < 5; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (u64)c; c >>= 64; }
}
static sc SC_INV2K[257];                 /* 2^-(256 + i) mod n */
static int sc_inv2k_ready = 0;
static void sc_inv2k_init(void) {
    sc x = SC_ONE;
    for (int i = 0; i < 256; i++) sc_half(&x);
    for (int i = 0; i <= 256; i++) { SC_INV2K[i] = x; sc_half(&x); }
    __atomic_store_n(&sc_inv2k_ready, 1, __ATOMIC_RELEASE);
}
static int sc_inv(sc *r, const sc *a) {
    if (u256_is_zero(a)) return 0;
    if (!__atomic_load_n(&sc_inv2k_ready, __ATOMIC_ACQUIRE)) sc_inv2k_init();    /* idempotent: racing threads write the same values */
    u256 u = FN, v = *a;
    u320 rr = {{0, 0, 0, 0, 0}}, ss = {{1, 0, 0, 0, 0}};
    int k = 0;
    while (!u256_is_zero(&v)) {
        if (!(u.v[0] & 1)) { u256_shr1(&u, 0); u320_shl1(&ss); }
        else if (!(v.v[0] & 1)) { u256_shr1(&v, 0); u320_shl1(&rr); }
        else if (!u256_geq(&v, &u)) { u256_sub(&u, &u, &v); u256_shr1(&u, 0); u320_add(&rr, &rr, &ss); u320_shl1(&ss); }
        else { u256_sub(&v, &v, &u); u256_shr1(&v, 0); u320_add(&ss, &ss, &rr); u320_shl1(&rr); }
        k++;
    }
    /* rr < 2 n:  a^-1 2^k = n - (rr mod n) */
    u256 x = {{rr.v[0], rr.v[1], rr.v[2], rr.v[3]}};
    if (rr.v[4] || u256_geq(&x, &FN)) u256_sub(&x, &x, &FN);
    u256_sub(&x, &FN, &x);
    sc_mul(r, &x, &SC_INV2K[k - 256]);
    return 1;
}
/* Scalar::generate_biased: 64 bytes big-endian mod n [recalled] */
static void sc_from_wide_be(sc *r, const u8 *b) {
    u64 t[8];
    for (int i = 0; i < 8; i++) {
        u64 w = 0;
        for (int j = 0; j < 8; j++) w = (w << 8) | b[8 * (7 - i) + j];
        t[i] = w;
    }
    sc_reduce_wide(r, t, 8);
}
/* Scalar::from_repr: 0 if >= n */
static int sc_from_repr(sc *r, const u8 *b) { u256_from_be(r, b); return !u256_geq(r, &FN); }

/* ------------------------------------------------------------------ */
/* group: projective (X:Y:Z), complete formulas (Renes-Costello-Batina) */
/* ------------------------------------------------------------------ */
typedef struct { fe x, y, z; } pt;
static const pt PT_IDENTITY = {{{0, 0, 0, 0}}, {{1, 0, 0, 0}}, {{0, 0, 0, 0}}};

static void pt_add(pt *r, const pt *p, const pt *q) {
    fe t0, t1, t2, t3, t4, t5, x3, y3, z3;
    fe_mul(&t0, &p->x, &q->x); fe_mul(&t1, &p->y, &q->y); fe_mul(&t2, &p->z, &q->z);
    fe_add(&t3, &p->x, &p->y); fe_add(&t4, &q->x, &q->y); fe_mul(&t3, &t3, &t4);
    fe_add(&t4, &t0, &t1); fe_sub(&t3, &t3, &t4);                 /* t3 = X1Y2 + X2Y1 */
    fe_add(&t4, &p->y, &p->z); fe_add(&t5, &q->y, &q->z); fe_mul(&t4, &t4, &t5);
    fe_add(&t5, &t1, &t2); fe_sub(&t4, &t4, &t5);                 /* t4 = Y1Z2 + Y2Z1 */
    fe_add(&x3, &p->x, &p->z); fe_add(&y3, &q->x, &q->z); fe_mul(&x3, &x3, &y3);
    fe_add(&y3, &t0, &t2); fe_sub(&y3, &x3, &y3);                 /* y3 = X1Z2 + X2Z1 */
    fe_add(&x3, &t0, &t0); fe_add(&t0, &x3, &t0);                 /* t0 = 3 X1X2 */
    fe_mul_small(&t2, &t2, 21);                                   /* b3 Z1Z2 */
    fe_add(&z3, &t1, &t2); fe_sub(&t1, &t1, &t2);
    fe_mul_small(&y3, &y3, 21);
    fe_mul(&x3, &t4, &y3); fe_mul(&t2, &t3, &t1); fe_sub(&x3, &t2, &x3);
    fe_mul(&y3, &y3, &t0); fe_mul(&t1, &t1, &z3); fe_add(&y3, &t1, &y3);
    fe_mul(&t0, &t0, &t3); fe_mul(&z3, &z3, &t4); fe_add(&z3, &z3, &t0);
    r->x = x3; r->y = y3; r->z = z3;
}
static void pt_double(pt *r, const pt *p) {
    fe t0, t1, t2, x3, y3, z3;
    fe_sqr(&t0, &p->y); fe_add(&z3, &t0, &t0); fe_add(&z3, &z3, &z3); fe_add(&z3, &z3, &z3);
    fe_mul(&t1, &p->y, &p->z); fe_sqr(&t2, &p->z); fe_mul_small(&t2, &t2, 21);
    fe_mul(&x3, &t2, &z3); fe_add(&y3, &t0, &t2); fe_mul(&z3, &t1, &z3);
    fe_add(&t1, &t2, &t2); fe_add(&t2, &t1, &t2); fe_sub(&t0, &t0, &t2);
    fe_mul(&y3, &t0, &y3); fe_add(&y3, &x3, &y3);
    fe_mul(&t1, &p->x, &p->y); fe_mul(&x3, &t0, &t1); fe_add(&x3, &x3, &x3);
    r->x = x3; r->y = y3; r->z = z3;
}
static void pt_neg(pt *r, const pt *p) { r->x = p->x; fe_neg(&r->y, &p->y); r->z = p->z; }
static void pt_sub(pt *r, const pt *p, const pt *q) { pt nq; pt_neg(&nq, q); pt_add(r, p, &nq); }
static int pt_is_identity(const pt *p) { fe z = p->z; fe_canon(&z); return u256_is_zero(&z); }
/* ProjectivePoint::eq */
static int pt_eq(const pt *p, const pt *q) {
    fe a, b, c, d;
    fe_mul(&a, &p->x, &q->z); fe_mul(&b, &q->x, &p->z);
    fe_mul(&c, &p->y, &q->z); fe_mul(&d, &q->y, &p->z);
    fe_canon(&a); fe_canon(&b); fe_canon(&c); fe_canon(&d);
    return u256_eq(&a, &b) && u256_eq(&c, &d);
}
/* ProjectivePoint * Scalar: one full multiplication per call, as the reference does (no sharing between the terms of an
 * MSM).  Like k256 it uses the GLV endomorphism lambda (x, y) = (beta x, y): k = k1 + k2 lambda with |k1|, |k2| < 2^128,
 * signed 4-bit windows over the two halves, 128 doublings. */
static const sc GLV_G1 = {{0xE893209A45DBB031ULL, 0x3DAA8A1471E8CA7FULL, 0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL}};
static const sc GLV_G2 = {{0x1571B4AE8AC47F71ULL, 0x221208AC9DF506C6ULL, 0x6F547FA90ABFE4C4ULL, 0xE4437ED6010E8828ULL}};
static const sc GLV_MB1 = {{0x6F547FA90ABFE4C3ULL, 0xE4437ED6010E8828ULL, 0, 0}};
static const sc GLV_MB2 = {{0xD765CDA83DB1562CULL, 0x8A280AC50774346DULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
static const sc GLV_LAMBDA = {{0xDF02967C1B23BD72ULL, 0x122E22EA20816678ULL, 0xA5261C028812645AULL, 0x5363AD4CC05C30E0ULL}};
static const fe GLV_BETA = {{0xC1396C28719501EEULL, 0x9CF0497512F58995ULL, 0x6E64479EAC3434E9ULL, 0x7AE96A2B657C0710ULL}};
static void mul_shift384(sc *r, const sc *k, const sc *g) {
    u64 t[8];
    mul_wide4(t, k, g);
    u128 c = (u128)t[6] + (t[5] >> 63);
    r->v[0] = (u64)c; c >>= 64; c += t[7]; r->v[1] = (u64)c; r->v[2] = (u64)(c >> 64); r->v[3] = 0;
}
static void pt_mul(pt *r, const pt *p, const sc *k) {
    sc c1, c2, k1, k2, t;
    mul_shift384(&c1, k, &GLV_G1); mul_shift384(&c2, k, &GLV_G2);
    sc_mul(&c1, &c1, &GLV_MB1); sc_mul(&c2, &c2, &GLV_MB2); sc_add(&k2, &c1, &c2);
    sc_mul(&t, &k2, &GLV_LAMBDA); sc_sub(&k1, k, &t);
    int neg1 = (k1.v[2] | k1.v[3]) != 0, neg2 = (k2.v[2] | k2.v[3]) != 0;
    if (neg1) sc_neg(&k1, &k1);
    if (neg2) sc_neg(&k2, &k2);
    pt tab[8], ltab[8];   /* 1P .. 8P and lambda * them */
    tab[0] = *p; pt_double(&tab[1], p); pt_add(&tab[2], &tab[1], p); pt_double(&tab[3], &tab[1]);
    pt_add(&tab[4], &tab[3], p); pt_double(&tab[5], &tab[2]); pt_add(&tab[6], &tab[5], p); pt_double(&tab[7], &tab[3]);
    for (int i = 0; i < 8; i++) { ltab[i] = tab[i]; fe_mul(&ltab[i].x, &ltab[i].x, &GLV_BETA); }
    /* signed digits: (k + 0x88..8) nibbles minus 8, plus an unsigned 33rd digit */
    u64 d1[3], d2[3]; u128 c;
    c = (u128)k1.v[0] + 0x8888888888888888ULL; d1[0] = (u64)c; c >>= 64; c += (u128)k1.v[1] + 0x8888888888888888ULL; d1[1] = (u64)c; d1[2] = (u64)(c >> 64);
    c = (u128)k2.v[0] + 0x8888888888888888ULL; d2[0] = (u64)c; c >>= 64; c += (u128)k2.v[1] + 0x8888888888888888ULL; d2[1] = (u64)c; d2[2] = (u64)(c >> 64);
    pt acc = PT_IDENTITY, e;
    for (int w = 32; w >= 0; w--) {
        if (w != 32) { pt_double(&acc, &acc); pt_double(&acc, &acc); pt_double(&acc, &acc); pt_double(&acc, &acc); }
        int a = (int)((d1[w / 16] >> (4 * (w % 16))) & 15), b = (int)((d2[w / 16] >> (4 * (w % 16))) & 15);
        if (w < 32) { a -= 8; b -= 8; }
        if (neg1) a = -a;
        if (neg2) b = -b;
        if (a) { e = tab[(a < 0 ? -a : a) - 1]; if (a < 0) fe_neg(&e.y, &e.y); pt_add(&acc, &acc, &e); }
        if (b) { e = ltab[(b < 0 ? -b : b) - 1]; if (b < 0) fe_neg(&e.y, &e.y); pt_add(&acc, &acc, &e); }
    }
    *r = acc;
}
/* to_affine + to_bytes: SEC1 compressed, identity = 33 zero bytes */
static void pt_to_bytes(u8 out[33], const pt *p) {
    if (pt_is_identity(p)) { memset(out, 0, 33); return; }
    fe zi, x, y;
    fe_inv(&zi, &p->z); fe_mul(&x, &p->x, &zi); fe_mul(&y, &p->y, &zi);
    fe_canon(&x); fe_canon(&y);
    out[0] = 2 + (u8)(y.v[0] & 1);
    u256_to_be(out + 1, &x);
}
/* returns 1 ok, 0 malformed */
static int pt_from_bytes(pt *r, const u8 in[33]) {
    int allzero = 1;
    for (int i = 0; i < 33; i++) allzero &= in[i] == 0;
    if (allzero) { *r = PT_IDENTITY; return 1; }
    if (in[0] != 2 && in[0] != 3) return 0;
    fe x, y, y2, seven = {{7, 0, 0, 0}};
    u256_from_be(&x, in + 1);
    if (u256_geq(&x, &FP)) return 0;
    fe_sqr(&y2, &x); fe_mul(&y2, &y2, &x); fe_add(&y2, &y2, &seven);
    if (!fe_sqrt(&y, &y2)) return 0;
    fe_canon(&y);
    if ((y.v[0] & 1) != (u64)(in[0] & 1)) fe_neg(&y, &y);
    r->x = x; r->y = y; r->z = SC_ONE;
    return 1;
}
/* 64-byte affine x||y big-endian; all-zero = identity. 1 ok / 0 not on curve */
static int pt_from_xy(pt *r, const u8 in[64]) {
    int allzero = 1;
    for (int i = 0; i < 64; i++) allzero &= in[i] == 0;
    if (allzero) { *r = PT_IDENTITY; return 1; }
    fe x, y, l, rr, seven = {{7, 0, 0, 0}};
    u256_from_be(&x, in); u256_from_be(&y, in + 32);
    if (u256_geq(&x, &FP) || u256_geq(&y, &FP)) return 0;
    fe_sqr(&l, &y); fe_sqr(&rr, &x); fe_mul(&rr, &rr, &x); fe_add(&rr, &rr, &seven);
    fe_canon(&l); fe_canon(&rr);
    if (!u256_eq(&l, &rr)) return 0;
    r->x = x; r->y = y; r->z = SC_ONE;
    return 1;
}
static void pt_to_xy(u8 out[64], const pt *p) {
    if (pt_is_identity(p)) { memset(out, 0, 64); return; }
    fe zi, x, y;
    fe_inv(&zi, &p->z); fe_mul(&x, &p->x, &zi); fe_mul(&y, &p->y, &zi);
    fe_canon(&x); fe_canon(&y);
    u256_to_be(out, &x); u256_to_be(out + 32, &y);
}

/* ------------------------------------------------------------------ */
/* Keccak-f[1600] / STROBE-128 / Merlin 3.0.0        */
/* ------------------------------------------------------------------ */
static const u64 KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static u64 rol64(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(u64 a[25]) {
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= KRC[rnd];
    }
}
#define STROBE_R 166
typedef struct { u8 st[200]; int pos, pos_begin, cur_flags; } merlin_t;
static void strobe_run_f(merlin_t *s) {
    s->st[s->pos] ^= (u8)s->pos_begin;
    s->st[s->pos + 1] ^= 0x04;
    s->st[STROBE_R + 1] ^= 0x80;
    u64 lanes[25];
    for (int i = 0; i < 25; i++) { u64 w = 0; for (int j = 7; j >= 0; j--) w = (w << 8) | s->st[8 * i + j]; lanes[i] = w; }
    keccak_f(lanes);
    for (int i = 0; i < 25; i++) for (int j = 0; j < 8; j++) s->st[8 * i + j] = (u8)(lanes[i] >> (8 * j));
    s->pos = 0; s->pos_begin = 0;
}
static void strobe_absorb(merlin_t *s, const u8 *d, size_t n) {
    for (size_t i = 0; i < n; i++) { s->st[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_squeeze(merlin_t *s, u8 *d, size_t n) {
    for (size_t i = 0; i < n; i++) { d[i] = s->st[s->pos]; s->st[s->pos++] = 0; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_begin_op(merlin_t *s, int flags, int more) {
    if (more) return;
    u8 hdr[2] = {(u8)s->pos_begin, (u8)flags};
    s->pos_begin = s->pos + 1;
    s->cur_flags = flags;
    strobe_absorb(s, hdr, 2);
    if ((flags & (4 | 32)) && s->pos != 0) strobe_run_f(s);
}
static void strobe_meta_ad(merlin_t *s, const u8 *d, size_t n, int more) { strobe_begin_op(s, 16 | 2, more); strobe_absorb(s, d, n); }
static void strobe_ad(merlin_t *s, const u8 *d, size_t n, int more) { strobe_begin_op(s, 2, more); strobe_absorb(s, d, n); }
static void strobe_prf(merlin_t *s, u8 *d, size_t n) { strobe_begin_op(s, 1 | 2 | 4, 0); strobe_squeeze(s, d, n); }
static void le32(u8 b[4], uint32_t x) { b[0] = (u8)x; b[1] = (u8)(x >> 8); b[2] = (u8)(x >> 16); b[3] = (u8)(x >> 24); }
static void merlin_append(merlin_t *t, const char *label, const u8 *m, size_t n) {
    u8 l4[4]; le32(l4, (uint32_t)n);
    strobe_meta_ad(t, (const u8 *)label, strlen(label), 0);
    strobe_meta_ad(t, l4, 4, 1);
    strobe_ad(t, m, n, 0);
}
static void merlin_init(merlin_t *t, const u8 *label, size_t n) {
    memset(t, 0, sizeof *t);
    static const u8 hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(t->st, hdr, 6); memcpy(t->st + 6, "STROBEv1.0.2", 12);
    u64 lanes[25];
    for (int i = 0; i < 25; i++) { u64 w = 0; for (int j = 7; j >= 0; j--) w = (w << 8) | t->st[8 * i + j]; lanes[i] = w; }
    keccak_f(lanes);
    for (int i = 0; i < 25; i++) for (int j = 0; j < 8; j++) t->st[8 * i + j] = (u8)(lanes[i] >> (8 * j));
    strobe_meta_ad(t, (const u8 *)"Merlin v1.0", 11, 0);
    merlin_append(t, "dom-sep", label, n);
}
static void merlin_append_u64(merlin_t *t, const char *label, u64 x) {
    u8 b[8]; for (int i = 0; i < 8; i++) b[i] = (u8)(x >> (8 * i));
    merlin_append(t, label, b, 8);
}
static void merlin_challenge(merlin_t *t, const char *label, u8 *out, size_t n) {
    u8 l4[4]; le32(l4, (uint32_t)n);
    strobe_meta_ad(t, (const u8 *)label, strlen(label), 0);
    strobe_meta_ad(t, l4, 4, 1);
    strobe_prf(t, out, n);
}

/* "panic" flag: the reference unwraps; we record and bail out with a status code */
enum { ORACLE_OK = 0, ORACLE_PANIC_INVERT_ZERO = -1, ORACLE_PANIC_CHALLENGE_RANGE = -2, ORACLE_BAD_POINT = -3,
       ORACLE_BAD_SCALAR = -4, ORACLE_BAD_ARG = -5 };
typedef struct arena { u8 *base; size_t used, cap; int status; } arena;
static void *aalloc(arena *A, size_t n) {
    n = (n + 15) & ~(size_t)15;
    if (A->used + n > A->cap) {
        /* grow: chain not needed -- allocate generously up front; abort on overflow */
        abort();
    }
    void *p = A->base + A->used; A->used += n; return p;
}

/* transcript.rs:6-8 */
static void app_point(const char *label, const pt *p, merlin_t *t) { u8 b[33]; pt_to_bytes(b, p); merlin_append(t, label, b, 33); }
/* transcript.rs:10-14 */
static sc get_challenge(const char *label, merlin_t *t, arena *A) {
    u8 b[32]; sc r;
    merlin_challenge(t, label, b, 32);
    if (!sc_from_repr(&r, b)) { if (!A->status) A->status = ORACLE_PANIC_CHALLENGE_RANGE; r = SC_ZERO; }
    return r;
}
static sc inv_or_panic(const sc *a, arena *A) {
    sc r = SC_ZERO;
    if (!sc_inv(&r, a)) { if (!A->status) A->status = ORACLE_PANIC_INVERT_ZERO; }
    return r;
}

/* ------------------------------------------------------------------ */
/* util.rs                                                              */
/* ------------------------------------------------------------------ */
typedef struct { sc *v; size_t n; } scv;
typedef struct { pt *v; size_t n; } ptv;
static scv scv_new(arena *A, size_t n) { scv r = {(sc *)aalloc(A, (n ? n : 1) * sizeof(sc)), n}; for (size_t i = 0; i < n; i++) r.v[i] = SC_ZERO; return r; }
static ptv ptv_new(arena *A, size_t n) { ptv r = {(pt *)aalloc(A, (n ? n : 1) * sizeof(pt)), n}; for (size_t i = 0; i < n; i++) r.v[i] = PT_IDENTITY; return r; }
static size_t zmax(size_t a, size_t b) { return a > b ? a : b; }
static sc scv_get(scv a, size_t i) { return i < a.n ? a.v[i] : SC_ZERO; }           /* vector_extend, util.rs:24-26 */
static pt ptv_get(ptv a, size_t i) { return i < a.n ? a.v[i] : PT_IDENTITY; }
static scv scv_slice(scv a, size_t from, size_t to) { scv r = {a.v + from, to - from}; return r; }
static ptv ptv_slice(ptv a, size_t from, size_t to) { ptv r = {a.v + from, to - from}; return r; }
static scv scv_concat(arena *A, scv a, scv b) { scv r = scv_new(A, a.n + b.n); memcpy(r.v, a.v, a.n * sizeof(sc)); memcpy(r.v + a.n, b.v, b.n * sizeof(sc)); return r; }
static ptv ptv_concat(arena *A, ptv a, ptv b) { ptv r = ptv_new(A, a.n + b.n); memcpy(r.v, a.v, a.n * sizeof(pt)); memcpy(r.v + a.n, b.v, b.n * sizeof(pt)); return r; }
/* util.rs:7-22 */
static void reduce_sc(arena *A, scv v, scv *r0, scv *r1) {
    *r0 = scv_new(A, (v.n + 1) / 2); *r1 = scv_new(A, v.n / 2);
    for (size_t i = 0; i < v.n; i++) { if (i % 2 == 0) r0->v[i / 2] = v.v[i]; else r1->v[i / 2] = v.v[i]; }
}
static void reduce_pt(arena *A, ptv v, ptv *r0, ptv *r1) {
    *r0 = ptv_new(A, (v.n + 1) / 2); *r1 = ptv_new(A, v.n / 2);
    for (size_t i = 0; i < v.n; i++) { if (i % 2 == 0) r0->v[i / 2] = v.v[i]; else r1->v[i / 2] = v.v[i]; }
}
/* util.rs:28-44 */
static sc weight_vector_mul_sc(scv a, scv b, const sc *w) {
    sc exp = SC_ONE, res = SC_ZERO, t, ai, bi;
    size_t m = zmax(a.n, b.n);
    for (size_t i = 0; i < m; i++) {
        sc_mul(&exp, &exp, w);
        ai = scv_get(a, i); bi = scv_get(b, i);
        sc_mul(&t, &bi, &exp); sc_mul(&t, &ai, &t); sc_add(&res, &res, &t);
    }
    return res;
}
/* util.rs:46-60 */
static sc vector_mul_sc(scv a, scv b) {
    sc res = SC_ZERO, t, ai, bi;
    size_t m = zmax(a.n, b.n);
    for (size_t i = 0; i < m; i++) { ai = scv_get(a, i); bi = scv_get(b, i); sc_mul(&t, &ai, &bi); sc_add(&res, &res, &t); }
    return res;
}
static pt vector_mul_pt(ptv a, scv b) { /* the naive MSM */
    pt res = PT_IDENTITY, t, ai; sc bi;
    size_t m = zmax(a.n, b.n);
    for (size_t i = 0; i < m; i++) { ai = ptv_get(a, i); bi = scv_get(b, i); pt_mul(&t, &ai, &bi); pt_add(&res, &res, &t); }
    return res;
}
/* util.rs:62-67 */
static scv vector_mul_on_scalar_sc(arena *A, scv a, const sc *s) { scv r = scv_new(A, a.n); for (size_t i = 0; i < a.n; i++) sc_mul(&r.v[i], &a.v[i], s); return r; }
static ptv vector_mul_on_scalar_pt(arena *A, ptv a, const sc *s) { ptv r = ptv_new(A, a.n); for (size_t i = 0; i < a.n; i++) pt_mul(&r.v[i], &a.v[i], s); return r; }
/* util.rs:69-85 */
static scv vector_add_sc(arena *A, scv a, scv b) { size_t m = zmax(a.n, b.n); scv r = scv_new(A, m); for (size_t i = 0; i < m; i++) { sc x = scv_get(a, i), y = scv_get(b, i); sc_add(&r.v[i], &x, &y); } return r; }
static scv vector_sub_sc(arena *A, scv a, scv b) { size_t m = zmax(a.n, b.n); scv r = scv_new(A, m); for (size_t i = 0; i < m; i++) { sc x = scv_get(a, i), y = scv_get(b, i); sc_sub(&r.v[i], &x, &y); } return r; }
static ptv vector_add_pt(arena *A, ptv a, ptv b) { size_t m = zmax(a.n, b.n); ptv r = ptv_new(A, m); for (size_t i = 0; i < m; i++) { pt x = ptv_get(a, i), y = ptv_get(b, i); pt_add(&r.v[i], &x, &y); } return r; }
/* util.rs:87-95 */
static scv e_pow(arena *A, const sc *v, size_t n) { scv r = scv_new(A, n); sc buf = SC_ONE; for (size_t i = 0; i < n; i++) { r.v[i] = buf; sc_mul(&buf, &buf, v); } return r; }
/* util.rs:111-116 */
static scv vector_tensor_mul(arena *A, scv a, scv b) {
    scv r = scv_new(A, a.n * b.n);
    for (size_t j = 0; j < b.n; j++) for (size_t i = 0; i < a.n; i++) sc_mul(&r.v[j * a.n + i], &a.v[i], &b.v[j]);
    return r;
}
/* dense row-major matrix */
typedef struct { sc *v; size_t rows, cols; } scm;
static scm scm_new(arena *A, size_t rows, size_t cols) { scm m = {(sc *)aalloc(A, ((rows * cols) != 0 ? rows * cols : 1) * sizeof(sc)), rows, cols}; for (size_t i = 0; i < rows * cols; i++) m.v[i] = SC_ZERO; return m; }
/* util.rs:118-132 */
static scm diag_inv(arena *A, const sc *x, size_t n) {
    sc xi = inv_or_panic(x, A), val = SC_ONE;
    scm m = scm_new(A, n, n);
    for (size_t i = 0; i < n; i++) { sc_mul(&val, &val, &xi); m.v[i * n + i] = val; }
    return m;
}
/* util.rs:134-142: out[j] = vector_mul(a, column_j(m)) with zero-extension */
static scv vector_mul_on_matrix(arena *A, scv a, scm m) {
    scv r = scv_new(A, m.cols);
    for (size_t j = 0; j < m.cols; j++) {
        sc res = SC_ZERO, t;
        size_t mm = zmax(a.n, m.rows);
        for (size_t i = 0; i < mm; i++) {
            sc ai = scv_get(a, i), bi = i < m.rows ? m.v[i * m.cols + j] : SC_ZERO;
            sc_mul(&t, &ai, &bi); sc_add(&res, &res, &t);
        }
        r.v[j] = res;
    }
    return r;
}
/* util.rs:153-155 */
static sc minus_sc(const sc *v) { sc m1, r; sc_sub(&m1, &SC_ZERO, &SC_ONE); sc_mul(&r, v, &m1); return r; }

/* ------------------------------------------------------------------ */
/* wnla.rs                                                              */
/* ------------------------------------------------------------------ */
typedef struct { pt g; ptv g_vec, h_vec; scv c; sc rho, mu; } wnla_t;   /* wnla.rs:12-19 */
typedef struct { ptv r, x; scv l, n; } wnla_proof_t;                    /* wnla.rs:25-30 */

/* wnla.rs:66-72 */
static pt wnla_commit(const wnla_t *w, scv l, scv n) {
    sc v = vector_mul_sc(w->c, l), t = weight_vector_mul_sc(n, n, &w->mu);
    sc_add(&v, &v, &t);
    pt r, a;
    pt_mul(&r, &w->g, &v);
    a = vector_mul_pt(w->h_vec, l); pt_add(&r, &r, &a);
    a = vector_mul_pt(w->g_vec, n); pt_add(&r, &r, &a);
    return r;
}
/* wnla.rs:75-121 */
static int wnla_verify(arena *A, const wnla_t *w, const pt *commitment, merlin_t *t, wnla_proof_t proof) {
    if (proof.x.n != proof.r.n) return 0;
    if (proof.x.n == 0) { pt c = wnla_commit(w, proof.l, proof.n); return pt_eq(commitment, &c); }
    scv c0, c1; ptv g0, g1, h0, h1;
    reduce_sc(A, w->c, &c0, &c1); reduce_pt(A, w->g_vec, &g0, &g1); reduce_pt(A, w->h_vec, &h0, &h1);
    const pt *xl = &proof.x.v[proof.x.n - 1], *rl = &proof.r.v[proof.r.n - 1];
    app_point("wnla_com", commitment, t); app_point("wnla_x", xl, t); app_point("wnla_r", rl, t);
    merlin_append_u64(t, "l.sz", (u64)w->h_vec.n); merlin_append_u64(t, "n.sz", (u64)w->g_vec.n);
    sc y = get_challenge("wnla_challenge", t, A);
    wnla_t w2;
    w2.g = w->g;
    w2.h_vec = vector_add_pt(A, h0, vector_mul_on_scalar_pt(A, h1, &y));
    w2.g_vec = vector_add_pt(A, vector_mul_on_scalar_pt(A, g0, &w->rho), vector_mul_on_scalar_pt(A, g1, &y));
    w2.c = vector_add_sc(A, c0, vector_mul_on_scalar_sc(A, c1, &y));
    sc y2; sc_mul(&y2, &y, &y); sc_sub(&y2, &y2, &SC_ONE);
    pt com_, a;
    pt_mul(&a, xl, &y); pt_add(&com_, commitment, &a);
    pt_mul(&a, rl, &y2); pt_add(&com_, &com_, &a);
    w2.rho = w->mu; sc_mul(&w2.mu, &w->mu, &w->mu);
    wnla_proof_t p2 = proof; p2.r.n--; p2.x.n--;
    return wnla_verify(A, &w2, &com_, t, p2);
}
/* wnla.rs:125-190.  out->r / out->x must have capacity for all rounds; filled innermost-first */
static void wnla_prove(arena *A, const wnla_t *w, const pt *commitment, merlin_t *t, scv l, scv n, wnla_proof_t *out) {
    if (l.n + n.n < 6) { out->l = l; out->n = n; out->r.n = 0; out->x.n = 0; return; }
    sc rho_inv = inv_or_panic(&w->rho, A);
    scv c0, c1, l0, l1, n0, n1; ptv g0, g1, h0, h1;
    reduce_sc(A, w->c, &c0, &c1); reduce_sc(A, l, &l0, &l1); reduce_sc(A, n, &n0, &n1);
    reduce_pt(A, w->g_vec, &g0, &g1); reduce_pt(A, w->h_vec, &h0, &h1);
    sc mu2; sc_mul(&mu2, &w->mu, &w->mu);
    sc two; sc_from_u64(&two, 2);
    sc vx = weight_vector_mul_sc(n0, n1, &mu2), tmp;
    sc_mul(&tmp, &rho_inv, &two); sc_mul(&vx, &vx, &tmp);
    tmp = vector_mul_sc(c0, l1); sc_add(&vx, &vx, &tmp);
    tmp = vector_mul_sc(c1, l0); sc_add(&vx, &vx, &tmp);
    sc vr = weight_vector_mul_sc(n1, n1, &mu2);
    tmp = vector_mul_sc(c1, l1); sc_add(&vr, &vr, &tmp);
    pt x, r, a;
    pt_mul(&x, &w->g, &vx);
    a = vector_mul_pt(h0, l1); pt_add(&x, &x, &a);
    a = vector_mul_pt(h1, l0); pt_add(&x, &x, &a);
    a = vector_mul_pt(g0, vector_mul_on_scalar_sc(A, n1, &w->rho)); pt_add(&x, &x, &a);
    a = vector_mul_pt(g1, vector_mul_on_scalar_sc(A, n0, &rho_inv)); pt_add(&x, &x, &a);
    pt_mul(&r, &w->g, &vr);
    a = vector_mul_pt(h1, l1); pt_add(&r, &r, &a);
    a = vector_mul_pt(g1, n1); pt_add(&r, &r, &a);
    app_point("wnla_com", commitment, t); app_point("wnla_x", &x, t); app_point("wnla_r", &r, t);
    merlin_append_u64(t, "l.sz", (u64)l.n); merlin_append_u64(t, "n.sz", (u64)n.n);
    sc y = get_challenge("wnla_challenge", t, A);
    wnla_t w2;
    w2.g = w->g;
    w2.h_vec = vector_add_pt(A, h0, vector_mul_on_scalar_pt(A, h1, &y));
    w2.g_vec = vector_add_pt(A, vector_mul_on_scalar_pt(A, g0, &w->rho), vector_mul_on_scalar_pt(A, g1, &y));
    w2.c = vector_add_sc(A, c0, vector_mul_on_scalar_sc(A, c1, &y));
    scv l_ = vector_add_sc(A, l0, vector_mul_on_scalar_sc(A, l1, &y));
    scv n_ = vector_add_sc(A, vector_mul_on_scalar_sc(A, n0, &rho_inv), vector_mul_on_scalar_sc(A, n1, &y));
    w2.rho = w->mu; w2.mu = mu2;
    pt com2 = wnla_commit(&w2, l_, n_);
    wnla_prove(A, &w2, &com2, t, l_, n_, out);
    out->r.v[out->r.n++] = r;
    out->x.v[out->x.n++] = x;
}

/* ------------------------------------------------------------------ */
/* circuit.rs                                                           */
/* ------------------------------------------------------------------ */
enum { PT_LO = 0, PT_LL = 1, PT_LR = 2, PT_NO = 3 }; /* circuit.rs:15-20 */
typedef struct {
    size_t dim_nm, dim_no, k, dim_nl, dim_nv, dim_nw;
    pt g; ptv g_vec, h_vec;
    scm W_m, W_l; scv a_m, a_l;
    int f_l, f_m;
    ptv g_vec_, h_vec_;
    /* partition(typ, j) -> index or -1, tabulated for j < part_n */
    const int32_t *part[4]; size_t part_n;
} circuit_t; /* circuit.rs:95-139 */
typedef struct { pt c_l, c_r, c_o, c_s; ptv r, x; scv l, n; } circuit_proof_t; /* circuit.rs:24-33 */
typedef struct { scv *v; scv s_v; scv w_l, w_r, w_o; } circuit_witness_t;      /* circuit.rs:80-91 */

static int part_get(const circuit_t *c, int typ, size_t j) { return j < c->part_n ? c->part[typ][j] : -1; }
/* circuit.rs:146-151 */
static pt circuit_commit(const circuit_t *c, scv v, const sc *s) {
    pt r, a;
    pt_mul(&r, &c->g, &v.v[0]);
    pt_mul(&a, &c->h_vec.v[0], s); pt_add(&r, &r, &a);
    a = vector_mul_pt(ptv_slice(c->h_vec, 9, c->h_vec.n), scv_slice(v, 1, v.n)); pt_add(&r, &r, &a);
    return r;
}
/* circuit.rs:559-570 */
static sc linear_comb_coef(const circuit_t *c, size_t i, const sc *lambda, const sc *mu) {
    sc coef = SC_ZERO, t;
    if (c->f_l) { sc_pow_u64(&t, lambda, (u64)(c->dim_nv * i)); sc_add(&coef, &coef, &t); }
    if (c->f_m) { sc_pow_u64(&t, mu, (u64)(c->dim_nv * i + 1)); sc_add(&coef, &coef, &t); }
    return coef;
}
/* circuit.rs:572-582 */
static scv collect_cl0(arena *A, const circuit_t *c, const sc *lambda, const sc *mu) {
    scv c_l0 = scv_new(A, c->dim_nv - 1);
    if (c->f_l) { scv ev = e_pow(A, lambda, c->dim_nv); c_l0 = scv_slice(ev, 1, ev.n); }
    if (c->f_m) { scv ev = e_pow(A, mu, c->dim_nv); c_l0 = vector_sub_sc(A, c_l0, vector_mul_on_scalar_sc(A, scv_slice(ev, 1, ev.n), mu)); }
    return c_l0;
}
/* circuit.rs:601-614 */
static scv collect_lambda(arena *A, const circuit_t *c, const sc *lambda, const sc *mu) {
    scv lv = e_pow(A, lambda, c->dim_nl);
    if (c->f_l && c->f_m) {
        sc pm, pl; sc_pow_u64(&pm, mu, (u64)c->dim_nv); sc_pow_u64(&pl, lambda, (u64)c->dim_nv);
        scv t1 = vector_tensor_mul(A, vector_mul_on_scalar_sc(A, e_pow(A, lambda, c->dim_nv), mu), e_pow(A, &pm, c->k));
        scv t2 = vector_tensor_mul(A, e_pow(A, mu, c->dim_nv), e_pow(A, &pl, c->k));
        lv = vector_sub_sc(A, lv, vector_add_sc(A, t1, t2));
    }
    return lv;
}
static scm sub_cols(arena *A, scm W, size_t rows, size_t from, size_t to) {
    scm m = scm_new(A, rows, to - from);
    for (size_t i = 0; i < rows; i++) for (size_t j = from; j < to; j++) m.v[i * m.cols + (j - from)] = W.v[i * W.cols + j];
    return m;
}
static scm map_f(arena *A, const circuit_t *c, size_t isz, size_t jsz, int typ, scm W_x) { /* circuit.rs:628-638 */
    scm m = scm_new(A, isz, jsz);
    for (size_t i = 0; i < isz; i++) for (size_t j = 0; j < jsz; j++) {
        int j_ = part_get(c, typ, j);
        m.v[i * jsz + j] = j_ >= 0 ? W_x.v[i * W_x.cols + (size_t)j_] : SC_ZERO;
    }
    return m;
}
/* circuit.rs:584-599 (+ :616-653) */
static void collect_c(arena *A, const circuit_t *c, scv lambda_vec, scv mu_vec, const sc *mu,
                      scv *c_nL, scv *c_nR, scv *c_nO, scv *c_lL, scv *c_lR, scv *c_lO) {
    size_t nm = c->dim_nm;
    scm M_lnL = sub_cols(A, c->W_l, c->dim_nl, 0, nm), M_mnL = sub_cols(A, c->W_m, c->dim_nm, 0, nm);
    scm M_lnR = sub_cols(A, c->W_l, c->dim_nl, nm, 2 * nm), M_mnR = sub_cols(A, c->W_m, c->dim_nm, nm, 2 * nm);
    scm W_lO = sub_cols(A, c->W_l, c->dim_nl, 2 * nm, c->W_l.cols), W_mO = sub_cols(A, c->W_m, c->dim_nm, 2 * nm, c->W_m.cols);
    scm M_lnO = map_f(A, c, c->dim_nl, c->dim_nm, PT_NO, W_lO), M_llL = map_f(A, c, c->dim_nl, c->dim_nv, PT_LL, W_lO);
    scm M_llR = map_f(A, c, c->dim_nl, c->dim_nv, PT_LR, W_lO), M_llO = map_f(A, c, c->dim_nl, c->dim_nv, PT_LO, W_lO);
    scm M_mnO = map_f(A, c, c->dim_nm, c->dim_nm, PT_NO, W_mO), M_mlL = map_f(A, c, c->dim_nm, c->dim_nv, PT_LL, W_mO);
    scm M_mlR = map_f(A, c, c->dim_nm, c->dim_nv, PT_LR, W_mO), M_mlO = map_f(A, c, c->dim_nm, c->dim_nv, PT_LO, W_mO);
    scm mdi = diag_inv(A, mu, c->dim_nm);
#define VM(a, m) vector_mul_on_matrix(A, a, m)
    *c_nL = VM(vector_sub_sc(A, VM(lambda_vec, M_lnL), VM(mu_vec, M_mnL)), mdi);
    *c_nR = VM(vector_sub_sc(A, VM(lambda_vec, M_lnR), VM(mu_vec, M_mnR)), mdi);
    *c_nO = VM(vector_sub_sc(A, VM(lambda_vec, M_lnO), VM(mu_vec, M_mnO)), mdi);
    *c_lL = vector_sub_sc(A, VM(lambda_vec, M_llL), VM(mu_vec, M_mlL));
    *c_lR = vector_sub_sc(A, VM(lambda_vec, M_llR), VM(mu_vec, M_mlR));
    *c_lO = vector_sub_sc(A, VM(lambda_vec, M_llO), VM(mu_vec, M_mlO));
#undef VM
}
static scv make_cr_tau(arena *A, const sc *tau, const sc *tau_inv, const sc *tau2, const sc *tau3, const sc *beta) {
    scv r = scv_new(A, 9); sc t;
    r.v[0] = SC_ONE;
    sc_mul(&r.v[1], tau_inv, beta); sc_mul(&r.v[2], tau, beta); sc_mul(&r.v[3], tau2, beta); sc_mul(&r.v[4], tau3, beta);
    sc_mul(&t, tau, tau3); sc_mul(&r.v[5], &t, beta);
    sc_mul(&t, tau2, tau3); sc_mul(&r.v[6], &t, beta);
    sc_mul(&t, tau3, tau3); sc_mul(&r.v[7], &t, beta);
    sc_mul(&t, tau3, tau3); sc_mul(&t, &t, tau); sc_mul(&r.v[8], &t, beta);
    return r;
}
static scv pad_to(arena *A, scv a, size_t n) { if (a.n >= n) return a; scv r = scv_new(A, n); memcpy(r.v, a.v, a.n * sizeof(sc)); return r; }

/* circuit.rs:154-256 */
static int circuit_verify(arena *A, const circuit_t *c, ptv v, merlin_t *t, circuit_proof_t proof) {
    app_point("commitment_cl", &proof.c_l, t); app_point("commitment_cr", &proof.c_r, t); app_point("commitment_co", &proof.c_o, t);
    for (size_t i = 0; i < v.n; i++) app_point("commitment_v", &v.v[i], t);
    sc rho = get_challenge("circuit_rho", t, A), lambda = get_challenge("circuit_lambda", t, A);
    sc beta = get_challenge("circuit_beta", t, A), delta = get_challenge("circuit_delta", t, A);
    sc mu; sc_mul(&mu, &rho, &rho);
    scv lambda_vec = collect_lambda(A, c, &lambda, &mu);
    scv mu_vec = vector_mul_on_scalar_sc(A, e_pow(A, &mu, c->dim_nm), &mu);
    scv c_nL, c_nR, c_nO, c_lL, c_lR, c_lO;
    collect_c(A, c, lambda_vec, mu_vec, &mu, &c_nL, &c_nR, &c_nO, &c_lL, &c_lR, &c_lO);
    sc two; sc_from_u64(&two, 2);
    pt v_ = PT_IDENTITY, a;
    for (size_t i = 0; i < c->k; i++) { sc cf = linear_comb_coef(c, i, &lambda, &mu); pt_mul(&a, &v.v[i], &cf); pt_add(&v_, &v_, &a); }
    pt_mul(&v_, &v_, &two);
    app_point("commitment_cs", &proof.c_s, t);
    sc tau = get_challenge("circuit_tau", t, A);
    sc tau_inv = inv_or_panic(&tau, A), tau2, tau3;
    sc_mul(&tau2, &tau, &tau); sc_mul(&tau3, &tau2, &tau);
    sc delta_inv = inv_or_panic(&delta, A), t3d;
    sc_mul(&t3d, &tau3, &delta_inv);
    scv pn_tau = vector_mul_on_scalar_sc(A, c_nO, &t3d);
    pn_tau = vector_sub_sc(A, pn_tau, vector_mul_on_scalar_sc(A, c_nL, &tau2));
    pn_tau = vector_add_sc(A, pn_tau, vector_mul_on_scalar_sc(A, c_nR, &tau));
    sc ps_tau = weight_vector_mul_sc(pn_tau, pn_tau, &mu), tmp;
    tmp = vector_mul_sc(lambda_vec, c->a_l); sc_mul(&tmp, &tmp, &tau3); sc_mul(&tmp, &tmp, &two); sc_add(&ps_tau, &ps_tau, &tmp);
    tmp = vector_mul_sc(mu_vec, c->a_m); sc_mul(&tmp, &tmp, &tau3); sc_mul(&tmp, &tmp, &two); sc_sub(&ps_tau, &ps_tau, &tmp);
    pt ptp; pt_mul(&ptp, &c->g, &ps_tau); a = vector_mul_pt(c->g_vec, pn_tau); pt_add(&ptp, &ptp, &a);
    scv cr_tau = make_cr_tau(A, &tau, &tau_inv, &tau2, &tau3, &beta);
    scv c_l0 = collect_cl0(A, c, &lambda, &mu);
    scv cl_tau = vector_mul_on_scalar_sc(A, c_lO, &t3d);
    cl_tau = vector_sub_sc(A, cl_tau, vector_mul_on_scalar_sc(A, c_lL, &tau2));
    cl_tau = vector_add_sc(A, cl_tau, vector_mul_on_scalar_sc(A, c_lR, &tau));
    cl_tau = vector_mul_on_scalar_sc(A, cl_tau, &two);
    cl_tau = vector_sub_sc(A, cl_tau, c_l0);
    scv cc = scv_concat(A, cr_tau, cl_tau);
    pt com = ptp;
    pt_mul(&a, &proof.c_s, &tau_inv); pt_add(&com, &com, &a);
    pt_mul(&a, &proof.c_o, &delta); pt_sub(&com, &com, &a);
    pt_mul(&a, &proof.c_l, &tau); pt_add(&com, &com, &a);
    pt_mul(&a, &proof.c_r, &tau2); pt_sub(&com, &com, &a);
    pt_mul(&a, &v_, &tau3); pt_add(&com, &com, &a);
    cc = pad_to(A, cc, c->h_vec.n + c->h_vec_.n);
    wnla_t w = {c->g, ptv_concat(A, c->g_vec, c->g_vec_), ptv_concat(A, c->h_vec, c->h_vec_), cc, rho, mu};
    wnla_proof_t wp = {proof.r, proof.x, proof.l, proof.n};
    return wnla_verify(A, &w, &com, t, wp);
}

typedef struct { const u8 *data; size_t pos, len; int exhausted; } byterng;
static sc generate_biased(byterng *g) {
    sc r = SC_ZERO;
    if (g->pos + 64 > g->len) { g->exhausted = 1; return r; }
    sc_from_wide_be(&r, g->data + g->pos); g->pos += 64;
    return r;
}
static scv part_vec(arena *A, const circuit_t *c, int typ, size_t size, scv w_o) { /* circuit.rs:303-333 */
    scv r = scv_new(A, size);
    for (size_t j = 0; j < size; j++) { int i = part_get(c, typ, j); if (i >= 0) r.v[j] = w_o.v[i]; }
    return r;
}
/* circuit.rs:260-556.  proof->r / proof->x need capacity for all WNLA rounds */
static void circuit_prove(arena *A, const circuit_t *c, ptv v, circuit_witness_t wit, merlin_t *t, byterng *rng, circuit_proof_t *proof) {
    scv ro = scv_new(A, 9), rl = scv_new(A, 9), rr = scv_new(A, 9);
    static const int ro_i[7] = {0, 1, 2, 3, 5, 6, 7}, rl_i[6] = {0, 1, 2, 4, 5, 6}, rr_i[5] = {0, 1, 3, 4, 5};
    for (int i = 0; i < 7; i++) ro.v[ro_i[i]] = generate_biased(rng);
    for (int i = 0; i < 6; i++) rl.v[rl_i[i]] = generate_biased(rng);
    for (int i = 0; i < 5; i++) rr.v[rr_i[i]] = generate_biased(rng);
    scv nl = wit.w_l, nr = wit.w_r;
    scv no = part_vec(A, c, PT_NO, c->dim_nm, wit.w_o), lo = part_vec(A, c, PT_LO, c->dim_nv, wit.w_o);
    scv ll = part_vec(A, c, PT_LL, c->dim_nv, wit.w_o), lr = part_vec(A, c, PT_LR, c->dim_nv, wit.w_o);
    pt co = vector_mul_pt(c->h_vec, scv_concat(A, ro, lo)), cl = vector_mul_pt(c->h_vec, scv_concat(A, rl, ll));
    pt cr = vector_mul_pt(c->h_vec, scv_concat(A, rr, lr)), a;
    a = vector_mul_pt(c->g_vec, no); pt_add(&co, &co, &a);
    a = vector_mul_pt(c->g_vec, nl); pt_add(&cl, &cl, &a);
    a = vector_mul_pt(c->g_vec, nr); pt_add(&cr, &cr, &a);
    app_point("commitment_cl", &cl, t); app_point("commitment_cr", &cr, t); app_point("commitment_co", &co, t);
    for (size_t i = 0; i < v.n; i++) app_point("commitment_v", &v.v[i], t);
    sc rho = get_challenge("circuit_rho", t, A), lambda = get_challenge("circuit_lambda", t, A);
    sc beta = get_challenge("circuit_beta", t, A), delta = get_challenge("circuit_delta", t, A);
    sc mu; sc_mul(&mu, &rho, &rho);
    scv lambda_vec = collect_lambda(A, c, &lambda, &mu);
    scv mu_vec = vector_mul_on_scalar_sc(A, e_pow(A, &mu, c->dim_nm), &mu);
    scv c_nL, c_nR, c_nO, c_lL, c_lR, c_lO;
    collect_c(A, c, lambda_vec, mu_vec, &mu, &c_nL, &c_nR, &c_nO, &c_lL, &c_lR, &c_lO);
    scv ls = scv_new(A, c->dim_nv), ns = scv_new(A, c->dim_nm);
    for (size_t i = 0; i < c->dim_nv; i++) ls.v[i] = generate_biased(rng);
    for (size_t i = 0; i < c->dim_nm; i++) ns.v[i] = generate_biased(rng);
    sc two; sc_from_u64(&two, 2);
    sc v_0 = SC_ZERO, tmp;
    scv rv = scv_new(A, 9), v_1 = scv_new(A, c->dim_nv - 1);
    for (size_t i = 0; i < c->k; i++) {
        sc cf = linear_comb_coef(c, i, &lambda, &mu);
        sc_mul(&tmp, &wit.v[i].v[0], &cf); sc_add(&v_0, &v_0, &tmp);
        sc_mul(&tmp, &wit.s_v.v[i], &cf); sc_add(&rv.v[0], &rv.v[0], &tmp);
        v_1 = vector_add_sc(A, v_1, vector_mul_on_scalar_sc(A, scv_slice(wit.v[i], 1, wit.v[i].n), &cf));
    }
    sc_mul(&v_0, &v_0, &two); sc_mul(&rv.v[0], &rv.v[0], &two);
    v_1 = vector_mul_on_scalar_sc(A, v_1, &two);
    scv c_l0 = collect_cl0(A, c, &lambda, &mu);
    sc f_[8], delta2, delta_inv = inv_or_panic(&delta, A);
    sc_mul(&delta2, &delta, &delta);
#define WVM(a, b) weight_vector_mul_sc(a, b, &mu)
#define VMS(a, b) vector_mul_sc(a, b)
#define VA(a, b) vector_add_sc(A, a, b)
    sc s0, s1, s2;
    /* -2 (circuit.rs:406) */
    s0 = WVM(ns, ns); f_[0] = minus_sc(&s0);
    /* -1 (:409-410) */
    s0 = VMS(c_l0, ls); s1 = WVM(ns, no); sc_mul(&s2, &delta, &two); sc_mul(&s1, &s2, &s1); sc_add(&f_[1], &s0, &s1);
    /* 0 (:413-416) */
    s0 = VMS(c_lR, ls); sc_mul(&s0, &s0, &two); f_[2] = minus_sc(&s0);
    s0 = VMS(c_l0, lo); sc_mul(&s0, &s0, &delta); sc_sub(&f_[2], &f_[2], &s0);
    s0 = WVM(ns, VA(nl, c_nR)); sc_mul(&s0, &s0, &two); sc_sub(&f_[2], &f_[2], &s0);
    s0 = WVM(no, no); sc_mul(&s0, &s0, &delta2); sc_sub(&f_[2], &f_[2], &s0);
    /* 1 (:419-423) */
    s0 = VMS(c_lL, ls); sc_mul(&f_[3], &s0, &two);
    s0 = VMS(c_lR, lo); sc_mul(&s0, &s0, &delta); sc_mul(&s0, &s0, &two); sc_add(&f_[3], &f_[3], &s0);
    s0 = VMS(c_l0, ll); sc_add(&f_[3], &f_[3], &s0);
    s0 = WVM(ns, VA(nr, c_nL)); sc_mul(&s0, &s0, &two); sc_add(&f_[3], &f_[3], &s0);
    s0 = WVM(no, VA(nl, c_nR)); sc_mul(&s0, &s0, &two); sc_mul(&s0, &s0, &delta); sc_add(&f_[3], &f_[3], &s0);
    /* 2 (:426-433) */
    f_[4] = WVM(c_nR, c_nR);
    s0 = VMS(c_lO, ls); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); sc_sub(&f_[4], &f_[4], &s0);
    s0 = VMS(c_lL, lo); sc_mul(&s0, &s0, &delta); sc_mul(&s0, &s0, &two); sc_sub(&f_[4], &f_[4], &s0);
    s0 = VMS(c_lR, ll); sc_mul(&s0, &s0, &two); sc_sub(&f_[4], &f_[4], &s0);
    s0 = VMS(c_l0, lr); sc_sub(&f_[4], &f_[4], &s0);
    s0 = WVM(ns, c_nO); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); sc_sub(&f_[4], &f_[4], &s0);
    s0 = WVM(no, VA(nr, c_nL)); sc_mul(&s0, &s0, &delta); sc_mul(&s0, &s0, &two); sc_sub(&f_[4], &f_[4], &s0);
    s0 = WVM(VA(nl, c_nR), VA(nl, c_nR)); sc_sub(&f_[4], &f_[4], &s0);
    /* 4 (:438-444) */
    s0 = WVM(c_nO, c_nR); sc_mul(&s0, &s0, &delta_inv); sc_mul(&f_[5], &s0, &two);
    s0 = WVM(c_nL, c_nL); sc_add(&f_[5], &f_[5], &s0);
    s0 = VMS(c_lO, ll); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); sc_sub(&f_[5], &f_[5], &s0);
    s0 = VMS(c_lL, lr); sc_mul(&s0, &s0, &two); sc_sub(&f_[5], &f_[5], &s0);
    s0 = VMS(c_lR, v_1); sc_mul(&s0, &s0, &two); sc_sub(&f_[5], &f_[5], &s0);
    s0 = WVM(VA(nl, c_nR), c_nO); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); sc_sub(&f_[5], &f_[5], &s0);
    s0 = WVM(VA(nr, c_nL), VA(nr, c_nL)); sc_sub(&f_[5], &f_[5], &s0);
    /* 5 (:447-450) */
    s0 = WVM(c_nO, c_nL); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); f_[6] = minus_sc(&s0);
    s0 = VMS(c_nO, lr); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); sc_add(&f_[6], &f_[6], &s0);
    s0 = VMS(c_lL, v_1); sc_mul(&s0, &s0, &two); sc_add(&f_[6], &f_[6], &s0);
    s0 = WVM(VA(nr, c_nL), c_nO); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); sc_add(&f_[6], &f_[6], &s0);
    /* 6 (:453) */
    s0 = VMS(c_lO, v_1); sc_mul(&s0, &s0, &delta_inv); sc_mul(&s0, &s0, &two); f_[7] = minus_sc(&s0);
    sc beta_inv = inv_or_panic(&beta, A);
    scv rs = scv_new(A, 9);
    /* circuit.rs:457-467 */
    sc_mul(&s0, &ro.v[1], &delta); sc_mul(&s0, &s0, &beta); sc_add(&rs.v[0], &f_[1], &s0);
    sc_mul(&rs.v[1], &f_[0], &beta_inv);
    sc_mul(&s0, &ro.v[0], &delta); sc_add(&s0, &s0, &f_[2]); sc_mul(&s0, &s0, &beta_inv); sc_sub(&rs.v[2], &s0, &rl.v[1]);
    sc_sub(&s0, &f_[3], &rl.v[0]); sc_mul(&s0, &s0, &beta_inv); sc_mul(&s1, &ro.v[2], &delta); sc_add(&s1, &s1, &rr.v[1]); sc_add(&rs.v[3], &s0, &s1);
    sc_add(&s0, &f_[4], &rr.v[0]); sc_mul(&s0, &s0, &beta_inv); sc_mul(&s1, &ro.v[3], &delta); sc_sub(&s1, &s1, &rl.v[2]); sc_add(&rs.v[4], &s0, &s1);
    sc_mul(&s0, &rv.v[0], &beta_inv); rs.v[5] = minus_sc(&s0);
    sc_mul(&s0, &f_[5], &beta_inv); sc_mul(&s1, &ro.v[5], &delta); sc_add(&s0, &s0, &s1); sc_add(&s0, &s0, &rr.v[3]); sc_sub(&rs.v[6], &s0, &rl.v[4]);
    sc_mul(&s0, &f_[6], &beta_inv); sc_add(&s0, &s0, &rr.v[4]); sc_mul(&s1, &ro.v[6], &delta); sc_add(&s0, &s0, &s1); sc_sub(&rs.v[7], &s0, &rl.v[5]);
    sc_mul(&s0, &f_[7], &beta_inv); sc_mul(&s1, &ro.v[7], &delta); sc_add(&s0, &s0, &s1); sc_sub(&s0, &s0, &rl.v[6]); sc_add(&rs.v[8], &s0, &rr.v[5]);
    pt cs = vector_mul_pt(c->h_vec, scv_concat(A, rs, ls));
    a = vector_mul_pt(c->g_vec, ns); pt_add(&cs, &cs, &a);
    app_point("commitment_cs", &cs, t);
    sc tau = get_challenge("circuit_tau", t, A);
    sc tau_inv = inv_or_panic(&tau, A), tau2, tau3, t3d;
    sc_mul(&tau2, &tau, &tau); sc_mul(&tau3, &tau2, &tau); sc_mul(&t3d, &tau3, &delta_inv);
#define VMOS(a, s) vector_mul_on_scalar_sc(A, a, &(s))
    /* circuit.rs:479-483 */
    scv l = VMOS(scv_concat(A, rs, ls), tau_inv);
    l = vector_sub_sc(A, l, VMOS(scv_concat(A, ro, lo), delta));
    l = vector_add_sc(A, l, VMOS(scv_concat(A, rl, ll), tau));
    l = vector_sub_sc(A, l, VMOS(scv_concat(A, rr, lr), tau2));
    l = vector_add_sc(A, l, VMOS(scv_concat(A, rv, v_1), tau3));
    scv pn_tau = VMOS(c_nO, t3d);
    pn_tau = vector_sub_sc(A, pn_tau, VMOS(c_nL, tau2));
    pn_tau = vector_add_sc(A, pn_tau, VMOS(c_nR, tau));
    sc ps_tau = WVM(pn_tau, pn_tau);
    s0 = VMS(lambda_vec, c->a_l); sc_mul(&s0, &s0, &tau3); sc_mul(&s0, &s0, &two); sc_add(&ps_tau, &ps_tau, &s0);
    s0 = VMS(mu_vec, c->a_m); sc_mul(&s0, &s0, &tau3); sc_mul(&s0, &s0, &two); sc_sub(&ps_tau, &ps_tau, &s0);
    scv n_tau = VMOS(ns, tau_inv);
    n_tau = vector_sub_sc(A, n_tau, VMOS(no, delta));
    n_tau = vector_add_sc(A, n_tau, VMOS(nl, tau));
    n_tau = vector_sub_sc(A, n_tau, VMOS(nr, tau2));
    scv n = VA(pn_tau, n_tau);
    scv cr_tau = make_cr_tau(A, &tau, &tau_inv, &tau2, &tau3, &beta);
    scv cl_tau = VMOS(c_lO, t3d);
    cl_tau = vector_sub_sc(A, cl_tau, VMOS(c_lL, tau2));
    cl_tau = vector_add_sc(A, cl_tau, VMOS(c_lR, tau));
    cl_tau = VMOS(cl_tau, two);
    cl_tau = vector_sub_sc(A, cl_tau, c_l0);
    scv cc = scv_concat(A, cr_tau, cl_tau);
    sc vv; sc_mul(&vv, &tau3, &v_0); sc_add(&vv, &ps_tau, &vv);
    pt com; pt_mul(&com, &c->g, &vv);
    a = vector_mul_pt(c->h_vec, l); pt_add(&com, &com, &a);
    a = vector_mul_pt(c->g_vec, n); pt_add(&com, &com, &a);
    size_t hn = c->h_vec.n + c->h_vec_.n, gn = c->g_vec.n + c->g_vec_.n;
    if (l.n < hn) { size_t target = hn; cc = pad_to(A, cc, cc.n + (target - l.n)); l = pad_to(A, l, target); }
    n = pad_to(A, n, gn);
    wnla_t w = {c->g, ptv_concat(A, c->g_vec, c->g_vec_), ptv_concat(A, c->h_vec, c->h_vec_), cc, rho, mu};
    wnla_proof_t wp; wp.r = proof->r; wp.x = proof->x; wp.r.n = wp.x.n = 0;
    wnla_prove(A, &w, &com, t, l, n, &wp);
    proof->c_l = cl; proof->c_r = cr; proof->c_o = co; proof->c_s = cs;
    proof->r = wp.r; proof->x = wp.x; proof->l = wp.l; proof->n = wp.n;
#undef WVM
#undef VMS
#undef VA
#undef VMOS
}

/* ------------------------------------------------------------------ */
/* range_proof/reciprocal.rs                                            */
/* ------------------------------------------------------------------ */
typedef struct { size_t dim_nd, dim_np; pt g; ptv g_vec, h_vec, g_vec_, h_vec_; } reciprocal_t; /* reciprocal.rs:64-84 */

/* reciprocal.rs:150-214 */
static circuit_t make_circuit(arena *A, const reciprocal_t *p, const sc *e) {
    circuit_t c; memset(&c, 0, sizeof c);
    size_t nm = p->dim_nd, no = p->dim_np, nv = p->dim_nd + 1, nl = nv, nw = p->dim_nd * 2 + p->dim_np;
    c.dim_nm = nm; c.dim_no = no; c.k = 1; c.dim_nl = nl; c.dim_nv = nv; c.dim_nw = nw;
    c.a_m = scv_new(A, nm); for (size_t i = 0; i < nm; i++) c.a_m.v[i] = SC_ONE;
    c.W_m = scm_new(A, nm, nw);
    sc me = minus_sc(e);
    for (size_t i = 0; i < nm; i++) c.W_m.v[i * nw + i + nm] = me;
    c.a_l = scv_new(A, nl);
    sc base; sc_from_u64(&base, (u64)(uint32_t)p->dim_np);
    c.W_l = scm_new(A, nl, nw);
    for (size_t i = 0; i < nm; i++) { sc pw; sc_pow_u64(&pw, &base, (u64)i); c.W_l.v[0 * nw + i] = minus_sc(&pw); }
    for (size_t i = 0; i < nm; i++) for (size_t j = 0; j < nm; j++) c.W_l.v[(i + 1) * nw + j + nm] = SC_ONE;
    for (size_t i = 0; i < nm; i++) c.W_l.v[(i + 1) * nw + i + nm] = SC_ZERO;
    for (size_t i = 0; i < nm; i++)
        for (size_t j = 0; j < no; j++) { /* the reference recomputes the inversion per (i, j): reciprocal.rs:179-183 */
            sc jj, s, inv; sc_from_u64(&jj, (u64)(uint32_t)j); sc_add(&s, e, &jj);
            inv = inv_or_panic(&s, A);
            c.W_l.v[(i + 1) * nw + j + 2 * nm] = minus_sc(&inv);
        }
    size_t pn = nv > nm ? nv : nm;
    int32_t *tab = (int32_t *)aalloc(A, 4 * pn * sizeof(int32_t));
    for (int typ = 0; typ < 4; typ++) {
        for (size_t j = 0; j < pn; j++) tab[typ * pn + j] = (typ == PT_LL && j < p->dim_np) ? (int32_t)j : -1;
        c.part[typ] = tab + typ * pn;
    }
    c.part_n = pn;
    c.g = p->g; c.g_vec = p->g_vec; c.h_vec = p->h_vec; c.g_vec_ = p->g_vec_; c.h_vec_ = p->h_vec_;
    c.f_l = 1; c.f_m = 0;
    return c;
}
/* reciprocal.rs:88-95 */
static pt reciprocal_commit_value(const reciprocal_t *p, const sc *x, const sc *s) { pt r, a; pt_mul(&r, &p->g, x); pt_mul(&a, &p->h_vec.v[0], s); pt_add(&r, &r, &a); return r; }
static pt reciprocal_commit_poles(const reciprocal_t *p, scv r, const sc *s) { pt o, a; pt_mul(&o, &p->h_vec.v[0], s); a = vector_mul_pt(ptv_slice(p->h_vec, 9, p->h_vec.n), r); pt_add(&o, &o, &a); return o; }
/* reciprocal.rs:98-107 */
static int reciprocal_verify(arena *A, const reciprocal_t *p, const pt *commitment, circuit_proof_t cp, const pt *pr, merlin_t *t) {
    app_point("reciprocal_commitment", commitment, t);
    sc e = get_challenge("reciprocal_challenge", t, A);
    circuit_t c = make_circuit(A, p, &e);
    ptv v = ptv_new(A, 1); pt_add(&v.v[0], commitment, pr);
    return circuit_verify(A, &c, v, t, cp);
}
/* reciprocal.rs:110-146 */
static void reciprocal_prove(arena *A, const reciprocal_t *p, const pt *commitment, const sc *wx, const sc *ws, scv m, scv digits,
                             merlin_t *t, byterng *rng, circuit_proof_t *cp, pt *pr) {
    app_point("reciprocal_commitment", commitment, t);
    sc e = get_challenge("reciprocal_challenge", t, A);
    scv r = scv_new(A, p->dim_nd);
    for (size_t i = 0; i < p->dim_nd; i++) { sc s; sc_add(&s, &digits.v[i], &e); r.v[i] = inv_or_panic(&s, A); }
    sc r_blind = generate_biased(rng);
    *pr = reciprocal_commit_poles(p, r, &r_blind);
    scv v = scv_new(A, 1 + r.n); v.v[0] = *wx; memcpy(v.v + 1, r.v, r.n * sizeof(sc));
    circuit_t c = make_circuit(A, p, &e);
    circuit_witness_t cw; scv vs[1] = {v};
    cw.v = vs; cw.s_v = scv_new(A, 1); sc_add(&cw.s_v.v[0], ws, &r_blind);
    cw.w_l = digits; cw.w_r = r; cw.w_o = m;
    ptv cv = ptv_new(A, 1); cv.v[0] = circuit_commit(&c, v, &cw.s_v.v[0]);
    circuit_prove(A, &c, cv, cw, t, rng, cp);
}

/* ------------------------------------------------------------------ */
/* exported C API (ctypes)                                              */
/* ------------------------------------------------------------------ */
static arena arena_make(size_t cap) { arena A = {(u8 *)malloc(cap), 0, cap, 0}; return A; }
static int load_points33(ptv out, const u8 *in) { for (size_t i = 0; i < out.n; i++) if (!pt_from_bytes(&out.v[i], in + 33 * i)) return 0; return 1; }
static int load_scalars(scv out, const u8 *in) { for (size_t i = 0; i < out.n; i++) if (!sc_from_repr(&out.v[i], in + 32 * i)) return 0; return 1; }
static void store_points33(u8 *out, const pt *v, size_t n) { for (size_t i = 0; i < n; i++) pt_to_bytes(out + 33 * i, &v[i]); }
static void store_scalars(u8 *out, const sc *v, size_t n) { for (size_t i = 0; i < n; i++) u256_to_be(out + 32 * i, &v[i]); }

/* u64 protocol context: gens = g || g_vec[16] || h_vec[32], 49 x 64 B affine (x||y BE) */
typedef struct { pt g; pt g_vec[16]; pt h_vec[32]; } u64ctx;
static int u64ctx_load(u64ctx *c, const u8 *gens64) {
    if (!pt_from_xy(&c->g, gens64)) return 0;
    for (int i = 0; i < 16; i++) if (!pt_from_xy(&c->g_vec[i], gens64 + 64 * (1 + i))) return 0;
    for (int i = 0; i < 32; i++) if (!pt_from_xy(&c->h_vec[i], gens64 + 64 * (17 + i))) return 0;
    return 1;
}
static reciprocal_t u64_reciprocal(u64ctx *c) { /* u64_proof.rs:43-51 */
    reciprocal_t p;
    p.dim_nd = 16; p.dim_np = 16; p.g = c->g;
    p.g_vec.v = c->g_vec; p.g_vec.n = 16;
    p.h_vec.v = c->h_vec; p.h_vec.n = 26;
    p.g_vec_.v = c->g_vec; p.g_vec_.n = 0;
    p.h_vec_.v = c->h_vec + 26; p.h_vec_.n = 6;
    return p;
}

/* U64RangeProofProtocol::commit_value (u64_proof.rs:37-39) -> 33 B */
int oracle_u64_commit(const u8 *gens64, u64 x, const u8 *s32, u8 *out33) {
    u64ctx c; if (!u64ctx_load(&c, gens64)) return ORACLE_BAD_POINT;
    sc xs, s; sc_from_u64(&xs, x); if (!sc_from_repr(&s, s32)) return ORACLE_BAD_SCALAR;
    reciprocal_t p = u64_reciprocal(&c);
    pt v = reciprocal_commit_value(&p, &xs, &s);
    pt_to_bytes(out33, &v);
    return ORACLE_OK;
}

static int u64_prove_one(u64ctx *c, u64 x, const u8 *s32, const u8 *rng_bytes, size_t rng_len, const u8 *label, size_t label_len, u8 *out525) {
    arena A = arena_make(8u << 20);
    reciprocal_t p = u64_reciprocal(c);
    sc xs, s; sc_from_u64(&xs, x);
    if (!sc_from_repr(&s, s32)) { free(A.base); return ORACLE_BAD_SCALAR; }
    /* u64_to_hex / u64_to_hex_mapped (u64_proof.rs:84-102) */
    scv digits = scv_new(&A, 16), m = scv_new(&A, 16);
    u64 xx = x;
    for (int i = 0; i < 16; i++) { unsigned d = (unsigned)(xx % 16); sc_from_u64(&digits.v[i], d); sc_add(&m.v[d], &m.v[d], &SC_ONE); xx /= 16; }
    pt com = reciprocal_commit_value(&p, &xs, &s);
    merlin_t t; merlin_init(&t, label, label_len);
    byterng rng = {rng_bytes, 0, rng_len, 0};
    circuit_proof_t cp; pt pr;
    cp.r = ptv_new(&A, 8); cp.x = ptv_new(&A, 8);
    reciprocal_prove(&A, &p, &com, &xs, &s, m, digits, &t, &rng, &cp, &pr);
    int st = A.status;
    if (rng.exhausted) st = ORACLE_BAD_ARG;
    if (!st && !(cp.r.n == 4 && cp.x.n == 4 && cp.l.n == 2 && cp.n.n == 1)) st = ORACLE_BAD_ARG;
    if (!st) {
        u8 *o = out525;
        pt_to_bytes(o, &cp.c_l); pt_to_bytes(o + 33, &cp.c_r); pt_to_bytes(o + 66, &cp.c_o); pt_to_bytes(o + 99, &cp.c_s);
        store_points33(o + 132, cp.r.v, 4); store_points33(o + 264, cp.x.v, 4);
        store_scalars(o + 396, cp.l.v, 2); store_scalars(o + 460, cp.n.v, 1);
        pt_to_bytes(o + 492, &pr);
    }
    free(A.base);
    return st;
}
/* returns 1 true, 0 false, <0 status (malformed input / reference panic) */
static int u64_verify_one(u64ctx *c, const u8 *v33, const u8 *proof525, const u8 *label, size_t label_len) {
    arena A = arena_make(8u << 20);
    reciprocal_t p = u64_reciprocal(c);
    pt V, pr; circuit_proof_t cp;
    ptv pts = ptv_new(&A, 12);
    cp.l = scv_new(&A, 2); cp.n = scv_new(&A, 1);
    int ok = pt_from_bytes(&V, v33) && load_points33(pts, proof525) && pt_from_bytes(&pr, proof525 + 492);
    if (!ok) { free(A.base); return ORACLE_BAD_POINT; }
    if (!load_scalars(cp.l, proof525 + 396) || !load_scalars(cp.n, proof525 + 460)) { free(A.base); return ORACLE_BAD_SCALAR; }
    cp.c_l = pts.v[0]; cp.c_r = pts.v[1]; cp.c_o = pts.v[2]; cp.c_s = pts.v[3];
    cp.r = ptv_slice(pts, 4, 8); cp.x = ptv_slice(pts, 8, 12);
    merlin_t t; merlin_init(&t, label, label_len);
    int res = reciprocal_verify(&A, &p, &V, cp, &pr, &t);
    if (A.status) res = A.status;
    free(A.base);
    return res;
}

typedef struct {
    u64ctx *ctx; int mode; size_t begin, end;
    const u64 *xs; const u8 *blinds, *rngs, *commits, *proofs_in, *label; size_t label_len;
    u8 *proofs_out; int32_t *status;
} batch_job;
static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (size_t i = j->begin; i < j->end; i++) {
        if (j->mode == 0) j->status[i] = u64_prove_one(j->ctx, j->xs[i], j->blinds + 32 * i, j->rngs + 3328 * i, 3328, j->label, j->label_len, j->proofs_out + 525 * i);
        else j->status[i] = u64_verify_one(j->ctx, j->commits + 33 * i, j->proofs_in + 525 * i, j->label, j->label_len);
    }
    return NULL;
}
static int run_batch(batch_job proto, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto; jobs[t].begin = n * t / threads; jobs[t].end = n * (t + 1) / threads;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return ORACLE_OK;
}
/* prove N proofs: xs[N], blinds N*32, rngs N*3328 (52 draws x 64 B) -> proofs N*525, status[N] */
int oracle_u64_prove_batch(const u8 *gens64, size_t n, const u64 *xs, const u8 *blinds, const u8 *rngs, const u8 *label, size_t label_len,
                           u8 *proofs_out, int32_t *status, int threads) {
    u64ctx c; if (!u64ctx_load(&c, gens64)) return ORACLE_BAD_POINT;
    batch_job j; memset(&j, 0, sizeof j);
    j.ctx = &c; j.mode = 0; j.xs = xs; j.blinds = blinds; j.rngs = rngs; j.label = label; j.label_len = label_len; j.proofs_out = proofs_out; j.status = status;
    return run_batch(j, n, threads);
}
/* verify N proofs: commits N*33, proofs N*525 -> status[N] in {1,0,<0} */
int oracle_u64_verify_batch(const u8 *gens64, size_t n, const u8 *commits, const u8 *proofs, const u8 *label, size_t label_len,
                            int32_t *status, int threads) {
    u64ctx c; if (!u64ctx_load(&c, gens64)) return ORACLE_BAD_POINT;
    batch_job j; memset(&j, 0, sizeof j);
    j.ctx = &c; j.mode = 1; j.commits = commits; j.proofs_in = proofs; j.label = label; j.label_len = label_len; j.status = status;
    return run_batch(j, n, threads);
}

/* ---- generic WNLA (wnla.rs), points as 64-byte affine, scalars 32 B BE ---- */
static int load_points64(ptv out, const u8 *in) { for (size_t i = 0; i < out.n; i++) if (!pt_from_xy(&out.v[i], in + 64 * i)) return 0; return 1; }
static size_t wnla_rounds(size_t ln, size_t nn) { size_t r = 0; while (ln + nn >= 6) { ln = (ln + 1) / 2; nn = (nn + 1) / 2; r++; } return r; }
size_t oracle_wnla_rounds(size_t l_len, size_t n_len) { return wnla_rounds(l_len, n_len); }
static size_t arena_cap_for(size_t n) { return (64u << 20) + n * 4096; }

int oracle_wnla_commit(const u8 *g64, const u8 *gvec64, size_t gn, const u8 *hvec64, size_t hn, const u8 *c32, size_t cn,
                       const u8 *rho32, const u8 *mu32, const u8 *l32, size_t ln, const u8 *n32, size_t nn, u8 *out33) {
    arena A = arena_make(arena_cap_for(gn + hn + cn + ln + nn));
    wnla_t w; int st = ORACLE_OK;
    w.g_vec = ptv_new(&A, gn); w.h_vec = ptv_new(&A, hn); w.c = scv_new(&A, cn);
    scv l = scv_new(&A, ln), n = scv_new(&A, nn);
    if (!pt_from_xy(&w.g, g64) || !load_points64(w.g_vec, gvec64) || !load_points64(w.h_vec, hvec64)) st = ORACLE_BAD_POINT;
    else if (!load_scalars(w.c, c32) || !sc_from_repr(&w.rho, rho32) || !sc_from_repr(&w.mu, mu32) || !load_scalars(l, l32) || !load_scalars(n, n32)) st = ORACLE_BAD_SCALAR;
    if (!st) { pt c = wnla_commit(&w, l, n); pt_to_bytes(out33, &c); }
    free(A.base);
    return st;
}
/* out: r_out/x_out rounds*33 B (innermost first), l_out/n_out 32 B each; *_len out params */
int oracle_wnla_prove(const u8 *g64, const u8 *gvec64, size_t gn, const u8 *hvec64, size_t hn, const u8 *c32, size_t cn,
                      const u8 *rho32, const u8 *mu32, const u8 *commit33, const u8 *l32, size_t ln, const u8 *n32, size_t nn,
                      const u8 *label, size_t label_len,
                      u8 *r_out, u8 *x_out, size_t *rounds_out, u8 *l_out, size_t *l_out_len, u8 *n_out, size_t *n_out_len) {
    size_t tot = gn + hn + cn + ln + nn;
    arena A = arena_make(arena_cap_for(tot) + tot * 2048);
    wnla_t w; int st = ORACLE_OK; pt com;
    w.g_vec = ptv_new(&A, gn); w.h_vec = ptv_new(&A, hn); w.c = scv_new(&A, cn);
    scv l = scv_new(&A, ln), n = scv_new(&A, nn);
    if (!pt_from_xy(&w.g, g64) || !load_points64(w.g_vec, gvec64) || !load_points64(w.h_vec, hvec64) || !pt_from_bytes(&com, commit33)) st = ORACLE_BAD_POINT;
    else if (!load_scalars(w.c, c32) || !sc_from_repr(&w.rho, rho32) || !sc_from_repr(&w.mu, mu32) || !load_scalars(l, l32) || !load_scalars(n, n32)) st = ORACLE_BAD_SCALAR;
    if (!st) {
        merlin_t t; merlin_init(&t, label, label_len);
        wnla_proof_t wp; wp.r = ptv_new(&A, 72); wp.x = ptv_new(&A, 72); wp.r.n = wp.x.n = 0;
        wnla_prove(&A, &w, &com, &t, l, n, &wp);
        st = A.status;
        if (!st) {
            store_points33(r_out, wp.r.v, wp.r.n); store_points33(x_out, wp.x.v, wp.x.n);
            *rounds_out = wp.r.n;
            store_scalars(l_out, wp.l.v, wp.l.n); *l_out_len = wp.l.n;
            store_scalars(n_out, wp.n.v, wp.n.n); *n_out_len = wp.n.n;
        }
    }
    free(A.base);
    return st;
}
int oracle_wnla_verify(const u8 *g64, const u8 *gvec64, size_t gn, const u8 *hvec64, size_t hn, const u8 *c32, size_t cn,
                       const u8 *rho32, const u8 *mu32, const u8 *commit33,
                       const u8 *r33, size_t rn, const u8 *x33, size_t xn, const u8 *l32, size_t ln, const u8 *n32, size_t nn,
                       const u8 *label, size_t label_len) {
    size_t tot = gn + hn + cn + ln + nn;
    arena A = arena_make(arena_cap_for(tot) + tot * 2048);
    wnla_t w; int st = ORACLE_OK; pt com; wnla_proof_t wp;
    w.g_vec = ptv_new(&A, gn); w.h_vec = ptv_new(&A, hn); w.c = scv_new(&A, cn);
    wp.r = ptv_new(&A, rn); wp.x = ptv_new(&A, xn); wp.l = scv_new(&A, ln); wp.n = scv_new(&A, nn);
    if (!pt_from_xy(&w.g, g64) || !load_points64(w.g_vec, gvec64) || !load_points64(w.h_vec, hvec64) || !pt_from_bytes(&com, commit33) ||
        !load_points33(wp.r, r33) || !load_points33(wp.x, x33)) st = ORACLE_BAD_POINT;
    else if (!load_scalars(w.c, c32) || !sc_from_repr(&w.rho, rho32) || !sc_from_repr(&w.mu, mu32) || !load_scalars(wp.l, l32) || !load_scalars(wp.n, n32)) st = ORACLE_BAD_SCALAR;
    if (!st) {
        merlin_t t; merlin_init(&t, label, label_len);
        st = wnla_verify(&A, &w, &com, &t, wp);
        if (A.status) st = A.status;
    }
    free(A.base);
    return st;
}

/* ---- generic reciprocal range proof (reciprocal.rs) for arbitrary (dim_nd, dim_np) ---- */
/* gens: g(64) ; g_vec gn*64 ; h_vec hn*64 (= dim_nd+1+9) ; g_vec_ gn_*64 ; h_vec_ hn_*64.
 * digits: dim_nd small ints (< dim_np); x32/s32 scalars.  rng: (1 + 18 + dim_nv + dim_nm) * 64 bytes.
 * out record: c_l c_r c_o c_s | r[rounds] | x[rounds] | l[..] | n[..] | r  with lengths reported. */
int oracle_reciprocal_prove(size_t dim_nd, size_t dim_np, const u8 *g64, const u8 *gvec64, size_t gn, const u8 *hvec64, size_t hn,
                            const u8 *gvec2_64, size_t gn2, const u8 *hvec2_64, size_t hn2,
                            const u8 *x32, const u8 *s32, const uint32_t *digits, const u8 *rng_bytes, size_t rng_len,
                            const u8 *label, size_t label_len,
                            u8 *out, size_t out_cap, size_t *rounds_out, size_t *l_len_out, size_t *n_len_out, u8 *commit33_out) {
    size_t nw = 2 * dim_nd + dim_np;
    arena A = arena_make((256u << 20) + (dim_nd + 1) * nw * 32 * 24);
    reciprocal_t p; int st = ORACLE_OK;
    p.dim_nd = dim_nd; p.dim_np = dim_np;
    p.g_vec = ptv_new(&A, gn); p.h_vec = ptv_new(&A, hn); p.g_vec_ = ptv_new(&A, gn2); p.h_vec_ = ptv_new(&A, hn2);
    sc x, s;
    if (!pt_from_xy(&p.g, g64) || !load_points64(p.g_vec, gvec64) || !load_points64(p.h_vec, hvec64) ||
        !load_points64(p.g_vec_, gvec2_64) || !load_points64(p.h_vec_, hvec2_64)) st = ORACLE_BAD_POINT;
    else if (!sc_from_repr(&x, x32) || !sc_from_repr(&s, s32)) st = ORACLE_BAD_SCALAR;
    if (!st) {
        scv dg = scv_new(&A, dim_nd), m = scv_new(&A, dim_np);
        for (size_t i = 0; i < dim_nd; i++) {
            if (digits[i] >= dim_np) { st = ORACLE_BAD_ARG; break; }
            sc_from_u64(&dg.v[i], digits[i]); sc_add(&m.v[digits[i]], &m.v[digits[i]], &SC_ONE);
        }
        if (!st) {
            pt com = reciprocal_commit_value(&p, &x, &s);
            merlin_t t; merlin_init(&t, label, label_len);
            byterng rng = {rng_bytes, 0, rng_len, 0};
            circuit_proof_t cp; pt pr;
            cp.r = ptv_new(&A, 72); cp.x = ptv_new(&A, 72);
            reciprocal_prove(&A, &p, &com, &x, &s, m, dg, &t, &rng, &cp, &pr);
            st = A.status; if (rng.exhausted) st = ORACLE_BAD_ARG;
            size_t need = 33 * (5 + 2 * cp.r.n) + 32 * (cp.l.n + cp.n.n);
            if (!st && need > out_cap) st = ORACLE_BAD_ARG;
            if (!st) {
                u8 *o = out;
                pt_to_bytes(o, &cp.c_l); pt_to_bytes(o + 33, &cp.c_r); pt_to_bytes(o + 66, &cp.c_o); pt_to_bytes(o + 99, &cp.c_s); o += 132;
                store_points33(o, cp.r.v, cp.r.n); o += 33 * cp.r.n;
                store_points33(o, cp.x.v, cp.x.n); o += 33 * cp.x.n;
                store_scalars(o, cp.l.v, cp.l.n); o += 32 * cp.l.n;
                store_scalars(o, cp.n.v, cp.n.n); o += 32 * cp.n.n;
                pt_to_bytes(o, &pr);
                *rounds_out = cp.r.n; *l_len_out = cp.l.n; *n_len_out = cp.n.n;
                pt_to_bytes(commit33_out, &com);
            }
        }
    }
    free(A.base);
    return st;
}
int oracle_reciprocal_verify(size_t dim_nd, size_t dim_np, const u8 *g64, const u8 *gvec64, size_t gn, const u8 *hvec64, size_t hn,
                             const u8 *gvec2_64, size_t gn2, const u8 *hvec2_64, size_t hn2,
                             const u8 *commit33, const u8 *rec, size_t rounds_r, size_t rounds_x, size_t l_len, size_t n_len,
                             const u8 *label, size_t label_len) {
    size_t nw = 2 * dim_nd + dim_np;
    arena A = arena_make((256u << 20) + (dim_nd + 1) * nw * 32 * 24);
    reciprocal_t p; int st = ORACLE_OK;
    p.dim_nd = dim_nd; p.dim_np = dim_np;
    p.g_vec = ptv_new(&A, gn); p.h_vec = ptv_new(&A, hn); p.g_vec_ = ptv_new(&A, gn2); p.h_vec_ = ptv_new(&A, hn2);
    pt V, pr; circuit_proof_t cp;
    ptv head = ptv_new(&A, 4);
    cp.r = ptv_new(&A, rounds_r); cp.x = ptv_new(&A, rounds_x); cp.l = scv_new(&A, l_len); cp.n = scv_new(&A, n_len);
    const u8 *o = rec;
    if (!pt_from_xy(&p.g, g64) || !load_points64(p.g_vec, gvec64) || !load_points64(p.h_vec, hvec64) ||
        !load_points64(p.g_vec_, gvec2_64) || !load_points64(p.h_vec_, hvec2_64) || !pt_from_bytes(&V, commit33)) st = ORACLE_BAD_POINT;
    if (!st) {
        if (!load_points33(head, o)) st = ORACLE_BAD_POINT; o += 132;
        if (!st && !load_points33(cp.r, o)) st = ORACLE_BAD_POINT; o += 33 * rounds_r;
        if (!st && !load_points33(cp.x, o)) st = ORACLE_BAD_POINT; o += 33 * rounds_x;
        if (!st && !load_scalars(cp.l, o)) st = ORACLE_BAD_SCALAR; o += 32 * l_len;
        if (!st && !load_scalars(cp.n, o)) st = ORACLE_BAD_SCALAR; o += 32 * n_len;
        if (!st && !pt_from_bytes(&pr, o)) st = ORACLE_BAD_POINT;
    }
    if (!st) {
        cp.c_l = head.v[0]; cp.c_r = head.v[1]; cp.c_o = head.v[2]; cp.c_s = head.v[3];
        merlin_t t; merlin_init(&t, label, label_len);
        st = reciprocal_verify(&A, &p, &V, cp, &pr, &t);
        if (A.status) st = A.status;
    }
    free(A.base);
    return st;
}

/* ---- generic arithmetic circuit (circuit.rs) with dense W_m / W_l and tabulated partition ---- */
typedef struct {
    size_t dim_nm, dim_no, k, dim_nv;   /* dim_nl = dim_nv*k, dim_nw = 2*dim_nm + dim_no */
    int f_l, f_m;
    const u8 *g64, *gvec64, *hvec64, *gvec2_64, *hvec2_64; size_t gn, hn, gn2, hn2;
    const u8 *W_m32, *W_l32, *a_m32, *a_l32;       /* row-major scalars */
    const int32_t *part_lo, *part_ll, *part_lr, *part_no; size_t part_n;
} oracle_circuit_desc;
static int circuit_load(arena *A, circuit_t *c, const oracle_circuit_desc *d) {
    memset(c, 0, sizeof *c);
    c->dim_nm = d->dim_nm; c->dim_no = d->dim_no; c->k = d->k; c->dim_nv = d->dim_nv; c->dim_nl = d->dim_nv * d->k; c->dim_nw = 2 * d->dim_nm + d->dim_no;
    c->f_l = d->f_l; c->f_m = d->f_m;
    c->g_vec = ptv_new(A, d->gn); c->h_vec = ptv_new(A, d->hn); c->g_vec_ = ptv_new(A, d->gn2); c->h_vec_ = ptv_new(A, d->hn2);
    if (!pt_from_xy(&c->g, d->g64) || !load_points64(c->g_vec, d->gvec64) || !load_points64(c->h_vec, d->hvec64) ||
        !load_points64(c->g_vec_, d->gvec2_64) || !load_points64(c->h_vec_, d->hvec2_64)) return ORACLE_BAD_POINT;
    c->W_m = scm_new(A, c->dim_nm, c->dim_nw); c->W_l = scm_new(A, c->dim_nl, c->dim_nw);
    c->a_m = scv_new(A, c->dim_nm); c->a_l = scv_new(A, c->dim_nl);
    scv wm = {c->W_m.v, c->dim_nm * c->dim_nw}, wl = {c->W_l.v, c->dim_nl * c->dim_nw};
    if (!load_scalars(wm, d->W_m32) || !load_scalars(wl, d->W_l32) || !load_scalars(c->a_m, d->a_m32) || !load_scalars(c->a_l, d->a_l32)) return ORACLE_BAD_SCALAR;
    c->part[PT_LO] = d->part_lo; c->part[PT_LL] = d->part_ll; c->part[PT_LR] = d->part_lr; c->part[PT_NO] = d->part_no; c->part_n = d->part_n;
    return ORACLE_OK;
}
/* ArithmeticCircuit::commit for witness vector i */
int oracle_circuit_commit(const oracle_circuit_desc *d, const u8 *v32, const u8 *s32, u8 *out33) {
    arena A = arena_make((64u << 20) + (d->dim_nv * d->k + d->dim_nm) * (2 * d->dim_nm + d->dim_no) * 64);
    circuit_t c; int st = circuit_load(&A, &c, d);
    scv v = scv_new(&A, d->dim_nv); sc s;
    if (!st && (!load_scalars(v, v32) || !sc_from_repr(&s, s32))) st = ORACLE_BAD_SCALAR;
    if (!st) { pt r = circuit_commit(&c, v, &s); pt_to_bytes(out33, &r); }
    free(A.base);
    return st;
}
/* witness: v k*dim_nv scalars, s_v k, w_l dim_nm, w_r dim_nm, w_o dim_no.  rng: (18 + dim_nv + dim_nm)*64 B.
 * out: c_l c_r c_o c_s | r[rounds] | x[rounds] | l | n */
int oracle_circuit_prove(const oracle_circuit_desc *d, const u8 *commits33, const u8 *v32, const u8 *sv32, const u8 *wl32, const u8 *wr32, const u8 *wo32,
                         const u8 *rng_bytes, size_t rng_len, const u8 *label, size_t label_len,
                         u8 *out, size_t out_cap, size_t *rounds_out, size_t *l_len_out, size_t *n_len_out) {
    arena A = arena_make((256u << 20) + (d->dim_nv * d->k + d->dim_nm) * (2 * d->dim_nm + d->dim_no) * 32 * 24);
    circuit_t c; int st = circuit_load(&A, &c, d);
    ptv cv = ptv_new(&A, d->k);
    circuit_witness_t w;
    scv *vs = (scv *)aalloc(&A, sizeof(scv) * (d->k ? d->k : 1));
    w.v = vs; w.s_v = scv_new(&A, d->k); w.w_l = scv_new(&A, d->dim_nm); w.w_r = scv_new(&A, d->dim_nm); w.w_o = scv_new(&A, d->dim_no);
    if (!st && !load_points33(cv, commits33)) st = ORACLE_BAD_POINT;
    if (!st) {
        for (size_t i = 0; i < d->k; i++) { vs[i] = scv_new(&A, d->dim_nv); if (!load_scalars(vs[i], v32 + 32 * d->dim_nv * i)) st = ORACLE_BAD_SCALAR; }
        if (!load_scalars(w.s_v, sv32) || !load_scalars(w.w_l, wl32) || !load_scalars(w.w_r, wr32) || !load_scalars(w.w_o, wo32)) st = ORACLE_BAD_SCALAR;
    }
    if (!st) {
        merlin_t t; merlin_init(&t, label, label_len);
        byterng rng = {rng_bytes, 0, rng_len, 0};
        circuit_proof_t cp; cp.r = ptv_new(&A, 72); cp.x = ptv_new(&A, 72);
        circuit_prove(&A, &c, cv, w, &t, &rng, &cp);
        st = A.status; if (rng.exhausted) st = ORACLE_BAD_ARG;
        size_t need = 33 * (4 + 2 * cp.r.n) + 32 * (cp.l.n + cp.n.n);
        if (!st && need > out_cap) st = ORACLE_BAD_ARG;
        if (!st) {
            u8 *o = out;
            pt_to_bytes(o, &cp.c_l); pt_to_bytes(o + 33, &cp.c_r); pt_to_bytes(o + 66, &cp.c_o); pt_to_bytes(o + 99, &cp.c_s); o += 132;
            store_points33(o, cp.r.v, cp.r.n); o += 33 * cp.r.n;
            store_points33(o, cp.x.v, cp.x.n); o += 33 * cp.x.n;
            store_scalars(o, cp.l.v, cp.l.n); o += 32 * cp.l.n;
            store_scalars(o, cp.n.v, cp.n.n);
            *rounds_out = cp.r.n; *l_len_out = cp.l.n; *n_len_out = cp.n.n;
        }
    }
    free(A.base);
    return st;
}
int oracle_circuit_verify(const oracle_circuit_desc *d, const u8 *commits33, const u8 *rec, size_t rounds_r, size_t rounds_x, size_t l_len, size_t n_len,
                          const u8 *label, size_t label_len) {
    arena A = arena_make((256u << 20) + (d->dim_nv * d->k + d->dim_nm) * (2 * d->dim_nm + d->dim_no) * 32 * 24);
    circuit_t c; int st = circuit_load(&A, &c, d);
    ptv cv = ptv_new(&A, d->k), head = ptv_new(&A, 4);
    circuit_proof_t cp;
    cp.r = ptv_new(&A, rounds_r); cp.x = ptv_new(&A, rounds_x); cp.l = scv_new(&A, l_len); cp.n = scv_new(&A, n_len);
    const u8 *o = rec;
    if (!st && !load_points33(cv, commits33)) st = ORACLE_BAD_POINT;
    if (!st) {
        if (!load_points33(head, o)) st = ORACLE_BAD_POINT; o += 132;
        if (!st && !load_points33(cp.r, o)) st = ORACLE_BAD_POINT; o += 33 * rounds_r;
        if (!st && !load_points33(cp.x, o)) st = ORACLE_BAD_POINT; o += 33 * rounds_x;
        if (!st && !load_scalars(cp.l, o)) st = ORACLE_BAD_SCALAR; o += 32 * l_len;
        if (!st && !load_scalars(cp.n, o)) st = ORACLE_BAD_SCALAR;
    }
    if (!st) {
        cp.c_l = head.v[0]; cp.c_r = head.v[1]; cp.c_o = head.v[2]; cp.c_s = head.v[3];
        merlin_t t; merlin_init(&t, label, label_len);
        st = circuit_verify(&A, &c, cv, &t, cp);
        if (A.status) st = A.status;
    }
    free(A.base);
    return st;
}

/* ---- primitives for differential tests ---- */
/* naive MSM exactly as util.rs:46-60 (points 64 B affine, scalars 32 B) -> 33 B */
int oracle_msm(const u8 *pts64, const u8 *sc32, size_t n, u8 *out33) {
    pt acc = PT_IDENTITY, p, t; sc k;
    for (size_t i = 0; i < n; i++) {
        if (!pt_from_xy(&p, pts64 + 64 * i)) return ORACLE_BAD_POINT;
        if (!sc_from_repr(&k, sc32 + 32 * i)) return ORACLE_BAD_SCALAR;
        pt_mul(&t, &p, &k); pt_add(&acc, &acc, &t);
    }
    pt_to_bytes(out33, &acc);
    return ORACLE_OK;
}
int oracle_point_mul(const u8 *p64, const u8 *k32, u8 *out64) {
    pt p, r; sc k;
    if (!pt_from_xy(&p, p64)) return ORACLE_BAD_POINT;
    if (!sc_from_repr(&k, k32)) return ORACLE_BAD_SCALAR;
    pt_mul(&r, &p, &k); pt_to_xy(out64, &r);
    return ORACLE_OK;
}
int oracle_point_add(const u8 *p64, const u8 *q64, u8 *out64) {
    pt p, q, r;
    if (!pt_from_xy(&p, p64) || !pt_from_xy(&q, q64)) return ORACLE_BAD_POINT;
    pt_add(&r, &p, &q); pt_to_xy(out64, &r);
    return ORACLE_OK;
}
int oracle_point_decompress(const u8 *in33, u8 *out64) { pt p; if (!pt_from_bytes(&p, in33)) return ORACLE_BAD_POINT; pt_to_xy(out64, &p); return ORACLE_OK; }
int oracle_point_compress(const u8 *in64, u8 *out33) { pt p; if (!pt_from_xy(&p, in64)) return ORACLE_BAD_POINT; pt_to_bytes(out33, &p); return ORACLE_OK; }
void oracle_fe_mul(const u8 *a32, const u8 *b32, u8 *out32) { fe a, b, r; u256_from_be(&a, a32); u256_from_be(&b, b32); fe_norm(&a, 0); fe_norm(&b, 0); fe_mul(&r, &a, &b); fe_canon(&r); u256_to_be(out32, &r); }
void oracle_fe_inv(const u8 *a32, u8 *out32) { fe a, r; u256_from_be(&a, a32); fe_norm(&a, 0); fe_inv(&r, &a); fe_canon(&r); u256_to_be(out32, &r); }
void oracle_sc_mul(const u8 *a32, const u8 *b32, u8 *out32) { sc a, b, r; u64 t[4]; u256_from_be(&a, a32); u256_from_be(&b, b32); memcpy(t, a.v, 32); sc_reduce_wide(&a, t, 4); memcpy(t, b.v, 32); sc_reduce_wide(&b, t, 4); sc_mul(&r, &a, &b); u256_to_be(out32, &r); }
int oracle_sc_inv(const u8 *a32, u8 *out32) { sc a, r; if (!sc_from_repr(&a, a32)) return ORACLE_BAD_SCALAR; if (!sc_inv(&r, &a)) return ORACLE_PANIC_INVERT_ZERO; u256_to_be(out32, &r); return ORACLE_OK; }
void oracle_sc_from_wide(const u8 *in64, u8 *out32) { sc r; sc_from_wide_be(&r, in64); u256_to_be(out32, &r); }
/* Merlin: new(label); for each (lab, msg): append_message; challenge_bytes(chal_label, n) */
void oracle_merlin_simple(const u8 *label, size_t label_len, const char *mlabel, const u8 *msg, size_t msg_len, const char *clabel, u8 *out, size_t n) {
    merlin_t t; merlin_init(&t, label, label_len);
    merlin_append(&t, mlabel, msg, msg_len);
    merlin_challenge(&t, clabel, out, n);
}
void oracle_keccak_f(u64 *lanes25) { keccak_f(lanes25); }
/* time `iters` scalar multiplications (for normalising the CPU baseline against k256's ~25.7 us) */
void oracle_bench_point_mul(const u8 *p64, const u8 *k32, int iters, u8 *out64) {
    pt p, r; sc k; pt_from_xy(&p, p64); sc_from_repr(&k, k32);
    r = p;
    for (int i = 0; i < iters; i++) { pt_mul(&r, &r, &k); }
    pt_to_xy(out64, &r);
}
/* time `iters` chained scalar inversions (k256 invert_vartime: ~5.2 us on an M3 Pro core) */
void oracle_bench_sc_inv(const u8 *a32, int iters, u8 *out32) {
    sc a, one = SC_ONE; sc_from_repr(&a, a32);
    for (int i = 0; i < iters; i++) { sc_inv(&a, &a); sc_add(&a, &a, &one); if (u256_is_zero(&a)) a = one; }
    u256_to_be(out32, &a);
}
