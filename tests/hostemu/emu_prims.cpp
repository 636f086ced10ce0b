// Host emulation of the DEVICE arithmetic headers (bp_pp_b200/csrc/*.cuh compiled as plain C++ with
// magnitude assertions on).  Test infrastructure only: lets the GPU-less CI exercise the exact
// field/scalar/point/transcript code the kernels run.  Never linked into libbppp.so.
#include <cstring>
#include "../../bp_pp_b200/csrc/ec.cuh"
#include "../../bp_pp_b200/csrc/merlin.cuh"
using namespace bppp;

static Fe fe_from_be(const uint8_t *b) { uint32_t w[8]; be32_to_words(w, b); return fe_from_words(w); }
static void fe_to_be(uint8_t *b, const Fe &a) { uint32_t w[8]; fe_to_words(w, fe_normalize(a)); words_to_be32(b, w); }

extern "C" {
void emu_fe_mul(const uint8_t *a, const uint8_t *b, uint8_t *r) { fe_to_be(r, fe_mul(fe_from_be(a), fe_from_be(b))); }
void emu_fe_sqr(const uint8_t *a, uint8_t *r) { fe_to_be(r, fe_sqr(fe_from_be(a))); }
void emu_fe_inv(const uint8_t *a, uint8_t *r) { fe_to_be(r, fe_inv(fe_from_be(a))); }
void emu_fe_inv_fermat(const uint8_t *a, uint8_t *r) { fe_to_be(r, fe_inv_fermat(fe_from_be(a))); }
void emu_sc_inv_fermat(const uint8_t *a, uint8_t *r) { Sc x; sc_from_be32(x, a); sc_to_be32(r, sc_inv_fermat(x)); }
// n inversions in one call (32 bytes big-endian each): field = 0 base field (any 256-bit input), 1 scalar field (canonical input)
void emu_inv_many(int field, size_t n, const uint8_t *a, uint8_t *r) {
    for (size_t i = 0; i < n; i++) {
        if (field == 0) fe_to_be(r + 32 * i, fe_inv(fe_from_be(a + 32 * i)));
        else { Sc x; sc_from_be32(x, a + 32 * i); sc_to_be32(r + 32 * i, sc_inv(x)); }
    }
}
void emu_fe_norm(const uint8_t *a, uint8_t *r) { fe_to_be(r, fe_from_be(a)); }
// (a*k1 + b*k2 - c) * d with lazy adds: exercises magnitude handling
void emu_fe_expr(const uint8_t *a, const uint8_t *b, const uint8_t *c, const uint8_t *d, uint32_t k1, uint32_t k2, uint8_t *r) {
    Fe t = fe_add(fe_mul_int(fe_from_be(a), k1), fe_mul_int(fe_from_be(b), k2));
    t = fe_sub(t, fe_from_be(c), 1);
    fe_to_be(r, fe_mul(t, fe_from_be(d)));
}
int emu_fe_sqrt(const uint8_t *a, uint8_t *r) {
    Fe x = fe_from_be(a), y = fe_sqrt_candidate(x);
    fe_to_be(r, y);
    return fe_is_zero(fe_sub(fe_sqr(y), x, 1)) ? 1 : 0;
}
void emu_sc_mul(const uint8_t *a, const uint8_t *b, uint8_t *r) { Sc x, y; sc_from_be32(x, a); sc_from_be32(y, b); sc_to_be32(r, sc_mul(x, y)); }
void emu_sc_add(const uint8_t *a, const uint8_t *b, uint8_t *r) { Sc x, y; sc_from_be32(x, a); sc_from_be32(y, b); sc_to_be32(r, sc_add(x, y)); }
void emu_sc_sub(const uint8_t *a, const uint8_t *b, uint8_t *r) { Sc x, y; sc_from_be32(x, a); sc_from_be32(y, b); sc_to_be32(r, sc_sub(x, y)); }
void emu_sc_neg(const uint8_t *a, uint8_t *r) { Sc x; sc_from_be32(x, a); sc_to_be32(r, sc_neg(x)); }
void emu_sc_inv(const uint8_t *a, uint8_t *r) { Sc x; sc_from_be32(x, a); sc_to_be32(r, sc_inv(x)); }
void emu_sc_wide(const uint8_t *a64, uint8_t *r) { sc_to_be32(r, sc_from_wide_be64(a64)); }
int emu_sc_from_repr(const uint8_t *a) { Sc x; return sc_from_be32(x, a) ? 1 : 0; }

static Pt load_pt(const uint8_t *xy, int *st) { PtA a; int s = pta_from_xy64(a, xy); *st = s; return pt_from_affine(a, s == 1); }
static void store_pt(uint8_t *out, const Pt &p) {
    bool id = pt_is_identity(p);
    Fe zi = fe_inv(p.z);
    PtA a = pt_to_affine_with_zinv(p, zi);
    pta_to_xy64(out, a, id);
}
int emu_pt_add(const uint8_t *p, const uint8_t *q, uint8_t *r) { int s1, s2; Pt a = load_pt(p, &s1), b = load_pt(q, &s2); if (s1 < 0 || s2 < 0) return -1; store_pt(r, pt_add(a, b)); return 0; }
int emu_pt_add_mixed(const uint8_t *p, const uint8_t *q, uint8_t *r) {
    int s1, s2; Pt a = load_pt(p, &s1); PtA b; s2 = pta_from_xy64(b, q); if (s1 < 0 || s2 != 0) return -1;
    store_pt(r, pt_add_mixed(a, b)); return 0;
}
// mixed add where the accumulator is a non-trivial projective representative (acc = k2*(p) scaled)
int emu_pt_add_mixed_proj(const uint8_t *p, const uint8_t *q, uint8_t *r) {
    int s1, s2; Pt a = load_pt(p, &s1); PtA b; s2 = pta_from_xy64(b, q); if (s1 < 0 || s2 != 0) return -1;
    Pt a3 = pt_add(pt_double(a), a);       // 3P projective
    store_pt(r, pt_add_mixed(a3, b)); return 0;
}
// sum of n affine points (64 B each, none the identity) through the XYZZ accumulator
int emu_ptx_sum(const uint8_t *pts, int n, uint8_t *r) {
    PtX acc = ptx_identity();
    for (int i = 0; i < n; i++) { PtA a; int s = pta_from_xy64(a, pts + 64 * i); if (s != 0) return -1; acc = ptx_add_mixed(acc, a); }
    store_pt(r, ptx_to_pt(acc)); return 0;
}
// Jacobian accumulator: k * base via double-and-add over the bits of k, then + sum of extra affine points
int emu_ptj_ladder(const uint8_t *base, const uint8_t *k32, const uint8_t *extra, int n_extra, uint8_t *r) {
    PtA b; if (pta_from_xy64(b, base) != 0) return -1;
    uint32_t w[8]; be32_to_words(w, k32);
    PtJ acc = ptj_identity();
    for (int bit = 255; bit >= 0; bit--) {
        acc = ptj_double(acc);
        if ((w[bit >> 5] >> (bit & 31)) & 1) acc = ptj_add_mixed(acc, b);
    }
    for (int i = 0; i < n_extra; i++) { PtA a; if (pta_from_xy64(a, extra + 64 * i) != 0) return -1; acc = ptj_add_mixed(acc, a); }
    store_pt(r, ptj_to_pt(acc)); return 0;
}
int emu_pt_double(const uint8_t *p, uint8_t *r) { int s; Pt a = load_pt(p, &s); if (s < 0) return -1; store_pt(r, pt_double(a)); return 0; }
int emu_pt_mul(const uint8_t *p, const uint8_t *k, uint8_t *r) { int s; Pt a = load_pt(p, &s); Sc kk; if (s < 0 || !sc_from_be32(kk, k)) return -1; store_pt(r, pt_mul(a, kk)); return 0; }
int emu_pt_mul_glv(const uint8_t *p, const uint8_t *k, uint8_t *r) { int s; Pt a = load_pt(p, &s); Sc kk; if (s < 0 || !sc_from_be32(kk, k)) return -1; store_pt(r, pt_mul_glv(a, kk)); return 0; }
// GLV split: out = k1 (16 B BE) || k2 (16 B BE) || neg1 || neg2
void emu_glv_split(const uint8_t *k, uint8_t *out) {
    Sc kk; sc_from_be32(kk, k);
    GlvSplit g = glv_split(kk);
    for (int i = 0; i < 4; i++) for (int b = 0; b < 4; b++) { out[15 - (4 * i + b)] = (uint8_t)(g.k1[i] >> (8 * b)); out[31 - (4 * i + b)] = (uint8_t)(g.k2[i] >> (8 * b)); }
    out[32] = g.neg1; out[33] = g.neg2;
}
int emu_pt_equal(const uint8_t *p, const uint8_t *q, const uint8_t *k) {
    // compare k*p (projective, non-trivial Z) against q
    int s1, s2; Pt a = load_pt(p, &s1), b = load_pt(q, &s2); Sc kk; sc_from_be32(kk, k);
    return pt_equal(pt_mul(a, kk), b) ? 1 : 0;
}
int emu_pt_decompress(const uint8_t *in33, uint8_t *out64) { PtA a; int s = pta_decompress(a, in33); if (s < 0) return s; pta_to_xy64(out64, a, s == 1); return s; }
int emu_pt_compress(const uint8_t *in64, uint8_t *out33) { PtA a; int s = pta_from_xy64(a, in64); if (s < 0) return s; a.x = fe_normalize(a.x); a.y = fe_normalize(a.y); pta_compress(out33, a, s == 1); return s; }
void emu_keccak(uint64_t *lanes) { keccak_f1600(lanes); }
void emu_merlin_simple(const uint8_t *label, uint32_t ll, const char *ml, uint32_t mll, const uint8_t *msg, uint32_t ml_n, const char *cl, uint32_t cll, uint8_t *out, uint32_t n) {
    Merlin m; merlin_init(m, label, ll);
    merlin_append(m, ml, mll, msg, ml_n);
    merlin_challenge(m, cl, cll, out, n);
}
// long transcript: several appends incl. u64 and multiple challenges, crossing the rate boundary
void emu_merlin_long(const uint8_t *label, uint32_t ll, const uint8_t *msg, uint32_t n_msg, uint32_t reps, uint8_t *out) {
    Merlin m; merlin_init(m, label, ll);
    for (uint32_t i = 0; i < reps; i++) {
        merlin_append(m, BPPP_LBL("wnla_com"), msg, n_msg);
        merlin_append_u64(m, BPPP_LBL("l.sz"), 32u >> (i & 3));
        merlin_challenge(m, BPPP_LBL("wnla_challenge"), out + 32 * i, 32);
    }
}
}
