// secp256k1 group law for sm_100a: homogeneous projective (X:Y:Z), complete formulas for
// y^2 = x^3 + 7 (a = 0, b3 = 21; Renes-Costello-Batina 2016) -- branch-free on every input,
// including P+P, P+(-P) and the identity (0:1:0), which tampered proofs do drive the verifier into.
//
// Replaces k256::ProjectivePoint::{add, sub, double, mul, eq, to_affine, to_bytes} as used at
// reference src/util.rs:40,56,66,75,84, src/wnla.rs:66-72,100-102, src/transcript.rs:6-8.
// The (X:Y:Z) representative is unobservable; results are compared/serialised in affine form.
#pragma once
#include "fe.cuh"
#include "sc.cuh"

namespace bppp {

struct Pt {   // projective; identity = (0 : 1 : 0)
    Fe x, y, z;
};
struct PtA {  // affine; (0, 0) is the sentinel for the identity (not on the curve)
    Fe x, y;
};

BPPP_HD Pt pt_identity() { Pt r; r.x = fe_zero(); r.y = fe_one(); r.z = fe_zero(); return r; }
BPPP_HD Pt pt_from_affine(const PtA &a, bool is_identity) {
    Pt r; r.x = a.x; r.y = a.y; r.z = fe_one();
    if (is_identity) r = pt_identity();
    return r;
}
BPPP_HD Pt pt_neg(const Pt &p) {
    Pt r = p; r.y = fe_normalize_weak(fe_negate(p.y, 1)); return r;
}
BPPP_HD Pt pt_cmov(const Pt &a, const Pt &b, bool take_b) {
    Pt r; r.x = fe_cmov(a.x, b.x, take_b); r.y = fe_cmov(a.y, b.y, take_b); r.z = fe_cmov(a.z, b.z, take_b); return r;
}

// Coordinates are any representatives below 2^256 (fe.cuh); the magnitude arguments of fe_negate / fe_sub are vestigial.

// P + Q, both projective.  12 M + 2 mul-by-21.
BPPP_HD Pt pt_add(const Pt &p, const Pt &q) {
    Fe t0 = fe_mul(p.x, q.x);                                  // X1X2
    Fe t1 = fe_mul(p.y, q.y);                                  // Y1Y2
    Fe t2 = fe_mul(p.z, q.z);                                  // Z1Z2
    Fe t3 = fe_mul(fe_add(p.x, p.y), fe_add(q.x, q.y));        // (X1+Y1)(X2+Y2)
    Fe t4 = fe_mul(fe_add(p.y, p.z), fe_add(q.y, q.z));        // (Y1+Z1)(Y2+Z2)
    Fe t5 = fe_mul(fe_add(p.x, p.z), fe_add(q.x, q.z));        // (X1+Z1)(X2+Z2)
    Fe xy = fe_sub(t3, fe_add(t0, t1), 2);   // X1Y2+X2Y1
    Fe yz = fe_sub(t4, fe_add(t1, t2), 2);   // Y1Z2+Y2Z1
    Fe xz = fe_sub(t5, fe_add(t0, t2), 2);   // X1Z2+X2Z1
    Fe x3 = fe_mul_int(t0, 3);   // 3 X1X2
    Fe bz = fe_normalize_weak(fe_mul_int(t2, 21));   // b3 Z1Z2
    Fe zp = fe_add(t1, bz);   // Y1Y2 + b3Z1Z2
    Fe zm = fe_sub(t1, bz, 1);   // Y1Y2 - b3Z1Z2
    Fe bxz = fe_normalize_weak(fe_mul_int(fe_normalize_weak(xz), 21));   // b3 (X1Z2+X2Z1)
    Pt r;
    r.x = fe_normalize_weak(fe_sub(fe_mul(xy, zm), fe_mul(yz, bxz), 1));
    r.y = fe_normalize_weak(fe_add(fe_mul(zm, zp), fe_mul(x3, bxz)));
    r.z = fe_normalize_weak(fe_add(fe_mul(yz, zp), fe_mul(x3, xy)));
    return r;
}

// P + Q with Q affine (Z2 = 1, Q != identity).  11 M.
BPPP_HD Pt pt_add_mixed(const Pt &p, const PtA &q) {
    Fe t0 = fe_mul(p.x, q.x);                                  // X1X2
    Fe t1 = fe_mul(p.y, q.y);                                  // Y1Y2
    Fe t3 = fe_mul(fe_add(p.x, p.y), fe_add(q.x, q.y));
    Fe xy = fe_sub(t3, fe_add(t0, t1), 2);   // X1Y2+X2Y1
    Fe yz = fe_add(fe_mul(q.y, p.z), p.y);   // Y2Z1+Y1
    Fe xz = fe_add(fe_mul(q.x, p.z), p.x);   // X2Z1+X1
    Fe x3 = fe_mul_int(t0, 3);
    Fe bz = fe_normalize_weak(fe_mul_int(p.z, 21));   // b3 Z1
    Fe zp = fe_add(t1, bz);
    Fe zm = fe_sub(t1, bz, 1);
    Fe bxz = fe_normalize_weak(fe_mul_int(fe_normalize_weak(xz), 21));
    Pt r;
    r.x = fe_normalize_weak(fe_sub(fe_mul(xy, zm), fe_mul(yz, bxz), 1));
    r.y = fe_normalize_weak(fe_add(fe_mul(zm, zp), fe_mul(x3, bxz)));
    r.z = fe_normalize_weak(fe_add(fe_mul(yz, zp), fe_mul(x3, xy)));
    return r;
}

// 2P.  6 M + 2 S.
BPPP_HD Pt pt_double(const Pt &p) {
    Fe yy = fe_sqr(p.y);                                       // Y^2
    Fe zz = fe_sqr(p.z);                                       // Z^2
    Fe yz = fe_mul(p.y, p.z);
    Fe xy = fe_mul(p.x, p.y);
    Fe bzz = fe_normalize_weak(fe_mul_int(zz, 21));   // b3 Z^2
    Fe y8 = fe_mul_int(yy, 8);   // 8 Y^2
    Fe t0 = fe_sub(yy, fe_mul_int(bzz, 3), 3);   // Y^2 - 9b Z^2
    Fe yp = fe_add(yy, bzz);   // Y^2 + 3b Z^2
    Pt r;
    r.x = fe_normalize_weak(fe_mul_int(fe_mul(t0, xy), 2));                    // 2 XY (Y^2 - 9bZ^2)
    r.y = fe_normalize_weak(fe_add(fe_mul(t0, yp), fe_mul(bzz, y8)));          // + 24 b Y^2 Z^2 = b3Z^2 * 8Y^2
    r.z = fe_mul(yz, y8);                                                      // 8 Y^3 Z
    return r;
}

BPPP_HD bool pt_is_identity(const Pt &p) { return fe_is_zero(p.z); }

// ProjectivePoint::eq -- X1 Z2 == X2 Z1 and Y1 Z2 == Y2 Z1
BPPP_HD bool pt_equal(const Pt &p, const Pt &q) {
    Fe a = fe_normalize(fe_sub(fe_mul(p.x, q.z), fe_mul(q.x, p.z), 1));
    Fe b = fe_normalize(fe_sub(fe_mul(p.y, q.z), fe_mul(q.y, p.z), 1));
    return fe_is_zero_canonical(a) && fe_is_zero_canonical(b);
}

// to_affine given zinv = 1/Z (0 for the identity).  Returns canonical coordinates.
BPPP_HD PtA pt_to_affine_with_zinv(const Pt &p, const Fe &zinv) {
    PtA r;
    r.x = fe_normalize(fe_mul(p.x, zinv));
    r.y = fe_normalize(fe_mul(p.y, zinv));
    return r;
}
// SEC1 compressed bytes (33): identity -> 33 zero bytes [recalled k256 convention]
BPPP_HD void pta_compress(uint8_t out[33], const PtA &a_canonical, bool is_identity) {
    uint32_t w[8];
    fe_to_words(w, a_canonical.x);
    out[0] = is_identity ? 0 : (uint8_t)(2u + (a_canonical.y.v[0] & 1u));
    words_to_be32(out + 1, w);
    if (is_identity) {
#pragma unroll
        for (int i = 1; i < 33; i++) out[i] = 0;
    }
}
// on-curve test for canonical affine coordinates
BPPP_HD bool pta_on_curve(const PtA &a) {
    Fe lhs = fe_sqr(a.y);
    Fe rhs = fe_add(fe_mul(fe_sqr(a.x), a.x), fe_from_u32(7));
    return fe_is_zero(fe_sub(lhs, rhs, 2));
}
// SEC1 compressed decode.  Returns 0 ok, 1 identity (all-zero), -1 malformed.
BPPP_HD int pta_decompress(PtA &r, const uint8_t in[33]) {
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 33; i++) any |= in[i];
    if (any == 0) { r.x = fe_zero(); r.y = fe_zero(); BPPP_SET_MAG(r.x, 1); BPPP_SET_MAG(r.y, 1); return 1; }
    if (in[0] != 2 && in[0] != 3) return -1;
    uint32_t w[8];
    be32_to_words(w, in + 1);
    if (words_ge_p(w)) return -1;
    Fe x = fe_from_words(w);
    Fe y2 = fe_add(fe_mul(fe_sqr(x), x), fe_from_u32(7));
    Fe y = fe_sqrt_candidate(y2);
    if (!fe_is_zero(fe_sub(fe_sqr(y), y2, 2))) return -1;
    y = fe_normalize(y);
    if ((y.v[0] & 1u) != (uint32_t)(in[0] & 1u)) y = fe_normalize(fe_negate(y, 1));
    r.x = x; r.y = y;
    return 0;
}
// 64-byte affine x||y big-endian.  Returns 0 ok, 1 identity (all-zero), -1 malformed / off-curve.
BPPP_HD int pta_from_xy64(PtA &r, const uint8_t in[64]) {
    uint32_t wx[8], wy[8];
    be32_to_words(wx, in); be32_to_words(wy, in + 32);
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) any |= wx[i] | wy[i];
    r.x = fe_from_words(wx); r.y = fe_from_words(wy);
    if (any == 0) return 1;
    if (words_ge_p(wx) || words_ge_p(wy)) return -1;
    if (!pta_on_curve(r)) return -1;
    return 0;
}
BPPP_HD void pta_to_xy64(uint8_t out[64], const PtA &a_canonical, bool is_identity) {
    uint32_t w[8];
    fe_to_words(w, a_canonical.x); words_to_be32(out, w);
    fe_to_words(w, a_canonical.y); words_to_be32(out + 32, w);
    if (is_identity) {
#pragma unroll
        for (int i = 0; i < 64; i++) out[i] = 0;
    }
}

// ---- XYZZ ("extended Jacobian") accumulator for sums of AFFINE points: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2 ----
// Mixed addition costs 8 M + 2 S (madd-2008-s) against 11 M for the complete projective formula, which is what the
// fixed-base and Pippenger accumulation loops are made of.  The formula is incomplete, so the exceptional inputs are
// handled explicitly and exactly (scalars chosen by an adversarial prover can steer an accumulator onto them):
// identity accumulator (flag), P1 == Q (affine doubling), P1 == -Q (identity).
// INL = true: field operations inlined into the point formula and branch-free (see fe_fold_add); false: the default
// call-based products and short-fold additions
template <bool INL> BPPP_HD Fe fe_mul_t(const Fe &a, const Fe &b) { if (INL) return fe_mul_inl_t<false>(a, b); return fe_mul(a, b); }
// fe_sub with the ignored magnitude argument of the formulas
template <bool BR> BPPP_HD Fe fe_subm_t(const Fe &a, const Fe &b, int = 0) { return fe_sub_t<BR>(a, b); }
template <bool INL> BPPP_HD Fe fe_sqr_t(const Fe &a) { if (INL) return fe_sqr_inl_t<false>(a); return fe_sqr(a); }
struct PtX {
    Fe x, y, zz, zzz;
    bool inf;
};
BPPP_HD PtX ptx_identity() { PtX r; r.x = fe_zero(); r.y = fe_zero(); r.zz = fe_zero(); r.zzz = fe_zero(); r.inf = true;
    BPPP_SET_MAG(r.x, 1); BPPP_SET_MAG(r.y, 1); BPPP_SET_MAG(r.zz, 1); BPPP_SET_MAG(r.zzz, 1); return r; }
// 2Q for affine Q (mdbl-2008-s-1)
template <bool INL> BPPP_HD PtX ptx_double_affine_t(const PtA &q) {
    PtX r;
    Fe U = fe_mul_int_t<!INL>(q.y, 2);   // 2y
    Fe V = fe_sqr_t<INL>(U);                                // 4y^2
    Fe W = fe_mul_t<INL>(U, V);                             // 8y^3
    Fe S = fe_mul_t<INL>(q.x, V);
    Fe M = fe_mul_int_t<!INL>(fe_sqr_t<INL>(q.x), 3);   // 3x^2
    r.x = fe_normalize_weak(fe_subm_t<!INL>(fe_sqr_t<INL>(M), fe_mul_int_t<!INL>(S, 2), 2));
    r.y = fe_normalize_weak(fe_subm_t<!INL>(fe_mul_t<INL>(M, fe_subm_t<!INL>(S, r.x, 1)), fe_mul_t<INL>(W, q.y), 1));
    r.zz = V; r.zzz = W;
    r.inf = fe_normalizes_to_zero(q.y);              // cannot happen on an odd-order curve; kept for exactness
    return r;
}
// P + Q, Q affine and not the identity
template <bool INL> BPPP_HD PtX ptx_add_mixed_t(const PtX &p, const PtA &q) {
    PtX r;
    if (p.inf) { r.x = q.x; r.y = q.y; r.zz = fe_one(); r.zzz = fe_one(); r.inf = false; return r; }
    Fe U2 = fe_mul_t<INL>(q.x, p.zz);
    Fe S2 = fe_mul_t<INL>(q.y, p.zzz);
    Fe P = fe_subm_t<!INL>(U2, p.x, 1);
    Fe R = fe_subm_t<!INL>(S2, p.y, 1);
    if (fe_normalizes_to_zero(P)) {                  // same x: P1 = +-Q
        if (fe_normalizes_to_zero(R)) return ptx_double_affine_t<INL>(q);
        return ptx_identity();
    }
    Fe PP = fe_sqr_t<INL>(P);
    Fe PPP = fe_mul_t<INL>(P, PP);
    Fe Q = fe_mul_t<INL>(p.x, PP);
    Fe RR = fe_sqr_t<INL>(R);
    r.x = fe_normalize_weak(fe_subm_t<!INL>(RR, fe_add_t<!INL>(PPP, fe_mul_int_t<!INL>(Q, 2)), 3));
    r.y = fe_normalize_weak(fe_subm_t<!INL>(fe_mul_t<INL>(R, fe_subm_t<!INL>(Q, r.x, 1)), fe_mul_t<INL>(p.y, PPP), 1));
    r.zz = fe_mul_t<INL>(p.zz, PP);
    r.zzz = fe_mul_t<INL>(p.zzz, PPP);
    r.inf = false;
    return r;
}
BPPP_HD PtX ptx_double_affine(const PtA &q) { return ptx_double_affine_t<false>(q); }
BPPP_HD PtX ptx_add_mixed(const PtX &p, const PtA &q) { return ptx_add_mixed_t<false>(p, q); }
// to homogeneous projective: (X ZZZ : Y ZZ : ZZ ZZZ)
BPPP_HD Pt ptx_to_pt(const PtX &p) {
    Pt r;
    r.x = fe_mul(p.x, p.zzz); r.y = fe_mul(p.y, p.zz); r.z = fe_mul(p.zz, p.zzz);
    if (p.inf) r = pt_identity();
    return r;
}

// ---- Jacobian accumulator (x = X/Z^2, y = Y/Z^3) for ladders over AFFINE per-proof tables ----
// doubling 2 M + 5 S (a = 0), mixed addition 8 M + 3 S, against 6 M + 2 S (+ two x21) and 12 M for the complete
// projective formulas.  Incomplete formulas: the identity (flag), P1 == Q (affine doubling) and P1 == -Q are handled
// explicitly and exactly -- tampered proofs do reach them (e.g. X_j == R_j).
struct PtJ {
    Fe x, y, z;
    bool inf;
};
BPPP_HD PtJ ptj_identity() { PtJ r; r.x = fe_zero(); r.y = fe_zero(); r.z = fe_zero(); r.inf = true;
    BPPP_SET_MAG(r.x, 1); BPPP_SET_MAG(r.y, 1); BPPP_SET_MAG(r.z, 1); return r; }
BPPP_HD PtJ ptj_from_affine(const PtA &a) { PtJ r; r.x = a.x; r.y = a.y; r.z = fe_one(); r.inf = false; return r; }
// 2P (dbl-2009-l).  Coordinates in: x, y weak-normalised, z limbs <= 4 * 2^26.  The curve has odd order, so Y != 0 unless P = O.
template <bool INL> BPPP_HD PtJ ptj_double_t(const PtJ &p) {
    PtJ r;
    Fe A = fe_sqr_t<INL>(p.x), B = fe_sqr_t<INL>(p.y), C = fe_sqr_t<INL>(B);
    Fe t = fe_sqr_t<INL>(fe_add_t<!INL>(p.x, B));
    Fe D = fe_normalize_weak(fe_mul_int_t<!INL>(fe_subm_t<!INL>(t, fe_add_t<!INL>(A, C), 2), 2));       // 2 ((X+B)^2 - A - C)
    Fe E = fe_mul_int_t<!INL>(A, 3);
    Fe F = fe_sqr_t<INL>(E);
    r.x = fe_normalize_weak(fe_subm_t<!INL>(F, fe_mul_int_t<!INL>(D, 2), 1));
    r.y = fe_normalize_weak(fe_subm_t<!INL>(fe_mul_t<INL>(E, fe_subm_t<!INL>(D, r.x, 1)), fe_mul_int_t<!INL>(C, 8), 8));
    r.z = fe_mul_int_t<!INL>(fe_mul_t<INL>(p.y, p.z), 2);
    r.inf = p.inf;
    return r;
}
// P + Q, Q affine and not the identity
template <bool INL> BPPP_HD PtJ ptj_add_mixed_t(const PtJ &p, const PtA &q) {
    if (p.inf) return ptj_from_affine(q);
    Fe Z1Z1 = fe_sqr_t<INL>(p.z);
    Fe U2 = fe_mul_t<INL>(q.x, Z1Z1);
    Fe S2 = fe_mul_t<INL>(fe_mul_t<INL>(q.y, p.z), Z1Z1);
    Fe H = fe_subm_t<!INL>(U2, p.x, 1);
    Fe Rh = fe_subm_t<!INL>(S2, p.y, 1);                       // (S2 - Y1)
    if (fe_normalizes_to_zero(H)) {                   // same x: P1 = +-Q
        if (fe_normalizes_to_zero(Rh)) return ptj_double_t<INL>(ptj_from_affine(q));
        return ptj_identity();
    }
    PtJ r;
    Fe HH = fe_sqr_t<INL>(H);
    Fe I = fe_mul_int_t<!INL>(HH, 4);
    Fe J = fe_mul_t<INL>(H, I);
    Fe rr = fe_mul_int_t<!INL>(Rh, 2);
    Fe V = fe_mul_t<INL>(p.x, I);
    r.x = fe_normalize_weak(fe_subm_t<!INL>(fe_sqr_t<INL>(rr), fe_add_t<!INL>(J, fe_mul_int_t<!INL>(V, 2)), 3));
    r.y = fe_normalize_weak(fe_subm_t<!INL>(fe_mul_t<INL>(rr, fe_subm_t<!INL>(V, r.x, 1)), fe_mul_int_t<!INL>(fe_mul_t<INL>(p.y, J), 2), 2));
    r.z = fe_mul_int_t<!INL>(fe_mul_t<INL>(p.z, H), 2);              // (Z1 + H)^2 - Z1Z1 - HH = 2 Z1 H
    r.inf = false;
    return r;
}
BPPP_HD PtJ ptj_double(const PtJ &p) { return ptj_double_t<false>(p); }
BPPP_HD PtJ ptj_add_mixed(const PtJ &p, const PtA &q) { return ptj_add_mixed_t<false>(p, q); }
// The accumulation loops call the *_hot forms.  With BPPP_PT_NOINLINE the call boundary moves from the field
// multiplication up to the point operation: one real function per formula with its 7-11 field products inlined, so the
// 16 + 8 argument / result moves of a by-value fe_mul call are paid once per point operation instead of once per
// product, and ptxas can interleave the carry chains of independent products.
// (BPPP_PT_NOINLINE selects all three; BPPP_PTJ_DBL_NOINLINE / BPPP_PTJ_ADD_NOINLINE / BPPP_PTX_ADD_NOINLINE one each.)
#if defined(BPPP_PT_NOINLINE)
#define BPPP_PTJ_DBL_NOINLINE 1
#define BPPP_PTJ_ADD_NOINLINE 1
#define BPPP_PTX_ADD_NOINLINE 1
#endif
#if defined(__CUDACC__) && defined(BPPP_PTJ_DBL_NOINLINE)
static __device__ __noinline__ PtJ ptj_double_call(PtJ p) { return ptj_double_t<true>(p); }
#endif
#if defined(__CUDACC__) && defined(BPPP_PTJ_ADD_NOINLINE)
static __device__ __noinline__ PtJ ptj_add_mixed_call(PtJ p, PtA q) { return ptj_add_mixed_t<true>(p, q); }
#endif
#if defined(__CUDACC__) && defined(BPPP_PTX_ADD_NOINLINE)
static __device__ __noinline__ PtX ptx_add_mixed_call(PtX p, PtA q) { return ptx_add_mixed_t<true>(p, q); }
#endif
BPPP_HD PtJ ptj_double_hot(const PtJ &p) {
#if defined(__CUDA_ARCH__) && defined(BPPP_PTJ_DBL_NOINLINE)
    return ptj_double_call(p);
#else
    return ptj_double(p);
#endif
}
BPPP_HD PtJ ptj_add_mixed_hot(const PtJ &p, const PtA &q) {
#if defined(__CUDA_ARCH__) && defined(BPPP_PTJ_ADD_NOINLINE)
    return ptj_add_mixed_call(p, q);
#else
    return ptj_add_mixed(p, q);
#endif
}
BPPP_HD PtX ptx_add_mixed_hot(const PtX &p, const PtA &q) {
#if defined(__CUDA_ARCH__) && defined(BPPP_PTX_ADD_NOINLINE)
    return ptx_add_mixed_call(p, q);
#else
    return ptx_add_mixed(p, q);
#endif
}
// to homogeneous projective (X Z : Y : Z^3)
BPPP_HD Pt ptj_to_pt(const PtJ &p) {
    Pt r;
    Fe zz = fe_sqr(p.z);
    r.x = fe_mul(p.x, p.z); r.y = fe_normalize_weak(p.y); r.z = fe_mul(zz, p.z);
    if (p.inf) r = pt_identity();
    return r;
}

// ---- variable-base scalar multiplication, signed 4-bit fixed windows ----
// k + C with C = sum_{i<64} 8*16^i gives unsigned nibbles d'_i; the signed digit is d'_i - 8 in [-8, 7]
// for i < 64, plus an unsigned top digit d'_64 in {0, 1}.  No data-dependent recoding.
struct Digits4 {
    uint32_t w[9];   // 65 nibbles of k + C
};
BPPP_HD Digits4 sc_signed_digits4(const Sc &k) {
    Digits4 d;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)k.v[i] + 0x88888888u; d.w[i] = (uint32_t)c; c >>= 32; }
    d.w[8] = (uint32_t)c;
    return d;
}
BPPP_HD int digits4_get(const Digits4 &d, int i) {   // i in 0..64 -> signed digit
    uint32_t nib = (d.w[i >> 3] >> (4 * (i & 7))) & 15u;
    return i < 64 ? (int)nib - 8 : (int)nib;
}

// table of 1P..8P for a projective P
struct PtTable8 {
    Pt m[8];
};
BPPP_HD void pt_table8_build(PtTable8 &t, const Pt &p) {
    t.m[0] = p;
    t.m[1] = pt_double(p);
    t.m[2] = pt_add(t.m[1], p);
    t.m[3] = pt_double(t.m[1]);
    t.m[4] = pt_add(t.m[3], p);
    t.m[5] = pt_double(t.m[2]);
    t.m[6] = pt_add(t.m[5], p);
    t.m[7] = pt_double(t.m[3]);
}
// signed digit lookup: returns d*P for d in [-8, 8], identity for 0
BPPP_HD Pt pt_table8_get(const PtTable8 &t, int d) {
    int a = d < 0 ? -d : d;
    Pt r = pt_identity();
    if (a != 0) {
        r = t.m[a - 1];
        if (d < 0) r = pt_neg(r);
    }
    return r;
}

// k*P (single point), used by the generic paths and tests
BPPP_HD Pt pt_mul(const Pt &p, const Sc &k) {
    PtTable8 tab;
    pt_table8_build(tab, p);
    Digits4 dg = sc_signed_digits4(k);
    Pt acc = pt_table8_get(tab, digits4_get(dg, 64));
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
        acc = pt_double(acc); acc = pt_double(acc); acc = pt_double(acc); acc = pt_double(acc);
        acc = pt_add(acc, pt_table8_get(tab, digits4_get(dg, i)));
    }
    return acc;
}

// ---- GLV endomorphism: lambda * (x, y) = (beta * x, y);  k = k1 + k2 * lambda (mod n), |k1|, |k2| < 2^128 ----
// Lattice basis from the extended Euclidean algorithm on (n, lambda); g1, g2 = round(2^384 b2 / n),
// round(2^384 (-b1) / n).  Constants and the < 2^128 bound were checked numerically (tests/test_hostemu.py).
struct GlvSplit {
    uint32_t k1[4], k2[4];   // magnitudes, little-endian words
    bool neg1, neg2;
};
// (k * g + 2^383) >> 384 for 256-bit k, g: the top 128 bits of the rounded 512-bit product
BPPP_HD Sc sc_mul_shift384(const Sc &k, const uint32_t g[8]) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)k.v[i] * g[j] + t[i + j];
            t[i + j] = (uint32_t)c; c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
    Sc r = sc_zero();
    uint64_t c = (uint64_t)(t[11] >> 31);
#pragma unroll
    for (int i = 0; i < 4; i++) { c += t[12 + i]; r.v[i] = (uint32_t)c; c >>= 32; }
    r.v[4] = (uint32_t)c;
    return r;
}
BPPP_HD GlvSplit glv_split(const Sc &k) {
    const uint32_t G1[8] = {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
    const uint32_t G2[8] = {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
    Sc mb1 = sc_zero(), mb2, lam;   // -b1, -b2 (mod n), lambda
    mb1.v[0] = 0x0ABFE4C3u; mb1.v[1] = 0x6F547FA9u; mb1.v[2] = 0x010E8828u; mb1.v[3] = 0xE4437ED6u;
    mb2.v[0] = 0x3DB1562Cu; mb2.v[1] = 0xD765CDA8u; mb2.v[2] = 0x0774346Du; mb2.v[3] = 0x8A280AC5u;
    mb2.v[4] = 0xFFFFFFFEu; mb2.v[5] = 0xFFFFFFFFu; mb2.v[6] = 0xFFFFFFFFu; mb2.v[7] = 0xFFFFFFFFu;
    lam.v[0] = 0x1B23BD72u; lam.v[1] = 0xDF02967Cu; lam.v[2] = 0x20816678u; lam.v[3] = 0x122E22EAu;
    lam.v[4] = 0x8812645Au; lam.v[5] = 0xA5261C02u; lam.v[6] = 0xC05C30E0u; lam.v[7] = 0x5363AD4Cu;
    Sc c1 = sc_mul_shift384(k, G1), c2 = sc_mul_shift384(k, G2);
    Sc k2 = sc_add(sc_mul(c1, mb1), sc_mul(c2, mb2));
    Sc k1 = sc_sub(k, sc_mul(k2, lam));
    GlvSplit r;
    r.neg1 = (k1.v[4] | k1.v[5] | k1.v[6] | k1.v[7]) != 0;   // canonical residue >= 2^128 means a negative half
    r.neg2 = (k2.v[4] | k2.v[5] | k2.v[6] | k2.v[7]) != 0;
    if (r.neg1) k1 = sc_neg(k1);
    if (r.neg2) k2 = sc_neg(k2);
#pragma unroll
    for (int i = 0; i < 4; i++) { r.k1[i] = k1.v[i]; r.k2[i] = k2.v[i]; }
    return r;
}
// signed 4-bit digits of a 128-bit magnitude: 33 nibbles of m + sum_{i<32} 8*16^i
struct Digits4h {
    uint32_t w[5];
};
BPPP_HD Digits4h half_signed_digits4(const uint32_t m[4]) {
    Digits4h d;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { c += (uint64_t)m[i] + 0x88888888u; d.w[i] = (uint32_t)c; c >>= 32; }
    d.w[4] = (uint32_t)c;
    return d;
}
BPPP_HD int digits4h_get(const Digits4h &d, int i) {   // i in 0..32
    uint32_t nib = (d.w[i >> 3] >> (4 * (i & 7))) & 15u;
    return i < 32 ? (int)nib - 8 : (int)nib;
}
BPPP_HD Fe fe_beta() {
    const uint32_t B[8] = {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu};
    return fe_from_words(B);
}

// joint sum_k ks[k] * P_k over NP points: 2*NP half-scalars of 128 bits, 128 shared doublings.
// tab[k] holds 1P..8P of point k; the lambda-half of a point reuses the same table with X scaled by beta.
template <int NP>
BPPP_HD Pt straus_glv(const PtTable8 *tab, const Sc *ks) {
    Digits4h dg[2 * NP];
    bool neg[2 * NP];
#pragma unroll 1
    for (int k = 0; k < NP; k++) {
        GlvSplit g = glv_split(ks[k]);
        dg[2 * k] = half_signed_digits4(g.k1); neg[2 * k] = g.neg1;
        dg[2 * k + 1] = half_signed_digits4(g.k2); neg[2 * k + 1] = g.neg2;
    }
    const Fe beta = fe_beta();
    Pt acc = pt_identity();
#pragma unroll 1
    for (int d = 32; d >= 0; d--) {
        if (d != 32) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = pt_double(acc);
        }
#pragma unroll 1
        for (int h = 0; h < 2 * NP; h++) {
            int sd = digits4h_get(dg[h], d);
            if (neg[h]) sd = -sd;
            Pt e = pt_table8_get(tab[h >> 1], sd);
            if (h & 1) e.x = fe_mul(e.x, beta);
            acc = pt_add(acc, e);
        }
    }
    return acc;
}
BPPP_HD Pt pt_mul_glv(const Pt &p, const Sc &k) {
    PtTable8 tab;
    pt_table8_build(tab, p);
    return straus_glv<1>(&tab, &k);
}

}  // namespace bppp
