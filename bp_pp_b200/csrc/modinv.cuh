// Modular inversion by the Bernstein-Yang "safegcd" division steps, for the two 256-bit moduli of secp256k1
// (base field p, group order n), host + device.
//
// Why: the Fermat inversions (fe.cuh:fe_inv_fermat, 255 squarings + 15 multiplications; sc.cuh:sc_inv_fermat, 252 + 78) cost
// 200-330 field multiplications each, i.e. 36-86 k instructions per thread.  600 division steps in batches of 30 on the
// low limbs, each batch followed by one 2x2-matrix update of the full-width (f, g) and (d, e) pairs, take ~14 k: the
// per-proof scalar inversion of the transcript phases and the one-per-thread inversions of every Montgomery batch
// (ws.cuh:batch_inv_strided, u64_verify.cuh:tables_*) get shorter.  Measured on a B200 (bppp_microbench): 1.08 G inversions/s
// for either modulus against 0.54 G/s (a^(p-2)) and 0.22 G/s (a^(n-2)), i.e. one inversion ~ 98 field multiplications.
//
// The algorithm (constant number of steps, no data-dependent branches, so a warp never diverges):
//   (f, g) = (M, x), (d, e) = (0, 1), zeta = -1;  invariant  d x = f, e x = g (mod M)
//   20 times: run 30 division steps on the low 30 bits of f, g, collecting them in a transition matrix t with
//             2^30 [f'; g'] = t [f; g];  apply t / 2^30 to (f, g) exactly and to (d, e) modulo M.
//   590 steps suffice for any 0 <= x < M < 2^256 (Bernstein-Yang 2019, with the half-delta start of Pornin / Wuille);
//   afterwards g = 0, f = +-gcd = +-1 and x^-1 = +-d.  x = 0 gives d = 0: the callers' inv(0) = 0 convention holds.
// Numbers are signed: 9 limbs of 30 bits, value = sum v[i] 2^(30 i), intermediate limbs in (-2^30, 2^30).
// This restates the published algorithm; the layout follows the well-known 30-bit-limb formulation for 32-bit machines.
// Replaces k256's FieldElement::invert / Scalar::invert on the device (k256 0.13.3, not in /root/reference).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BPPP_MI_HD __host__ __device__ __forceinline__
#else
#define BPPP_MI_HD inline
#endif

namespace bppp {

struct MI30 { int32_t v[9]; };
struct MITrans { int32_t u, v, q, r; };
static constexpr int32_t MI_M30 = 0x3FFFFFFF;

// modulus limbs and the modulus' inverse modulo 2^30 (M * inv30 = 1 mod 2^30)
struct MIModP {
    BPPP_MI_HD static int32_t m(int i) {
        return i == 0 ? 0x3FFFFC2F : i == 1 ? 0x3FFFFFFB : i < 8 ? 0x3FFFFFFF : 0xFFFF;
    }
    static constexpr uint32_t inv30 = 0x2DDACACFu;
};
struct MIModN {
    BPPP_MI_HD static int32_t m(int i) {
        return i == 0 ? 0x10364141 : i == 1 ? 0x3F497A33 : i == 2 ? 0x348A03BB : i == 3 ? 0x2BB739AB : i == 4 ? 0x3FFFFEBA : i < 8 ? 0x3FFFFFFF : 0xFFFF;
    }
    static constexpr uint32_t inv30 = 0x2A774EC1u;
};

// 8 x 32-bit little-endian words (value < 2^256) <-> 9 x 30-bit limbs
BPPP_MI_HD MI30 mi_from_words(const uint32_t w[8]) {
    MI30 r;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, word = bit >> 5, off = bit & 31;
        uint32_t lo = w[word] >> off;
        if (off > 2 && word + 1 < 8) lo |= w[word + 1] << (32 - off);
        r.v[i] = (int32_t)(lo & (uint32_t)MI_M30);
    }
    return r;
}
// limbs in [0, 2^30), value < 2^256
BPPP_MI_HD void mi_to_words(uint32_t w[8], const MI30 &a) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int bit = 32 * j, limb = bit / 30, off = bit % 30;
        uint32_t x = (uint32_t)a.v[limb] >> off;
        x |= (uint32_t)a.v[limb + 1] << (30 - off);
        if (60 - off < 32 && limb + 2 < 9) x |= (uint32_t)a.v[limb + 2] << (60 - off);
        w[j] = x;
    }
}

// 30 division steps on the low bits of f (odd) and g; returns the new zeta and the transition matrix
BPPP_MI_HD int32_t mi_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, MITrans &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 1
    for (int i = 0; i < 30; i++) {
        uint32_t mask1 = (uint32_t)(zeta >> 31);          // zeta < 0
        const uint32_t mask2 = 0u - (g & 1u);             // g odd
        const uint32_t x = (f ^ mask1) - mask1, y = (u ^ mask1) - mask1, z = (v ^ mask1) - mask1;   // conditionally negated f, u, v
        g += x & mask2; q += y & mask2; r += z & mask2;
        mask1 &= mask2;                                   // zeta < 0 and g odd: the pair is swapped
        zeta = (int32_t)((uint32_t)zeta ^ mask1) - 1;     // -zeta - 2 on a swap, zeta - 1 otherwise
        f += g & mask1; u += q & mask1; v += r & mask1;
        g >>= 1; u <<= 1; v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}
// (f, g) <- t (f, g) / 2^30, exact (the low 30 bits of both combinations are zero by construction)
BPPP_MI_HD void mi_update_fg(MI30 &f, MI30 &g, const MITrans &t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
    cf >>= 30; cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        const int64_t fi = f.v[i], gi = g.v[i];
        cf += u * fi + v * gi; cg += q * fi + r * gi;
        f.v[i - 1] = (int32_t)cf & MI_M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & MI_M30; cg >>= 30;
    }
    f.v[8] = (int32_t)cf; g.v[8] = (int32_t)cg;
}
// (d, e) <- t (d, e) / 2^30 (mod M); d, e stay in (-2 M, M): a multiple of M is added that clears the low 30 bits
template <class MOD>
BPPP_MI_HD void mi_update_de(MI30 &d, MI30 &e, const MITrans &t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;            // sign masks
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + r * e.v[0];
    md -= (int32_t)((MOD::inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)MI_M30);
    me -= (int32_t)((MOD::inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)MI_M30);
    cd += (int64_t)MOD::m(0) * md; ce += (int64_t)MOD::m(0) * me;
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        const int64_t di = d.v[i], ei = e.v[i];
        cd += u * di + v * ei; ce += q * di + r * ei;
        cd += (int64_t)MOD::m(i) * md; ce += (int64_t)MOD::m(i) * me;
        d.v[i - 1] = (int32_t)cd & MI_M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & MI_M30; ce >>= 30;
    }
    d.v[8] = (int32_t)cd; e.v[8] = (int32_t)ce;
}
// r in (-2 M, M) -> [0, M), negated first when sign < 0
template <class MOD>
BPPP_MI_HD void mi_normalize(MI30 &r, int32_t sign) {
    int32_t cond_add = r.v[8] >> 31;
    const int32_t cond_negate = sign >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        r.v[i] += MOD::m(i) & cond_add;
        r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= MI_M30; }
    cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += MOD::m(i) & cond_add;
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= MI_M30; }
}
// x^-1 mod M for canonical x (0 <= x < M), little-endian words in and out; 0 -> 0
template <class MOD>
BPPP_MI_HD void mi_modinv_words(uint32_t out[8], const uint32_t x[8]) {
    MI30 d, e, f, g = mi_from_words(x);
#pragma unroll
    for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; f.v[i] = MOD::m(i); }
    e.v[0] = 1;
    int32_t zeta = -1;
#pragma unroll 1
    for (int it = 0; it < 20; it++) {
        MITrans t;
        zeta = mi_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        mi_update_de<MOD>(d, e, t);
        mi_update_fg(f, g, t);
    }
    mi_normalize<MOD>(d, f.v[8]);
    mi_to_words(out, d);
}

}  // namespace bppp
