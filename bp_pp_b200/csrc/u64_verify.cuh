// U64RangeProofProtocol::verify (reference src/range_proof/u64_proof.rs:42-54) for a batch of
// independent proofs run in lockstep: the per-proof logic of every phase between two multi-scalar
// multiplications, as host+device functions over the word-major workspace.  The call chain being
// restated is reciprocal.rs:98-107 -> circuit.rs:154-256 -> wnla.rs:75-121, with the reciprocal
// circuit's coefficient vectors in closed form (derivation in DESIGN.md):
//   c_nL[j] = -16^j mu^-(j+1)        c_nR[j] = (S - lambda^(j+1)) mu^-(j+1) + e     c_nO = 0
//   c_lL[j] = -S / (e + j) (j < 16)  c_lR = c_lO = 0    c_l0[j] = lambda^(j+1)       S = sum_{i=1..16} lambda^i
// Exact arithmetic => bit-identical to the dense-matrix evaluation of circuit.rs:584-653.
#pragma once
#include "ws.cuh"

namespace bppp {

// generator indices in the context: 0 = g, 1..16 = g_vec, 17..48 = h_vec (u64_proof.rs:19-28)
static constexpr int GEN_G = 0, GEN_GVEC = 1, GEN_HVEC = 17, NUM_GENS = 49;

// input point slots
enum { VP_V = 0, VP_CL = 1, VP_CR = 2, VP_CO = 3, VP_CS = 4, VP_R = 5 /*r[0..3]*/, VP_X = 9 /*x[0..3]*/, VP_RR = 13, VP_COUNT = 14 };

// word offsets of the per-proof verify record
struct VL {
    static constexpr int STATUS = 0;
    static constexpr int IDMASK = 1;                    // bit k: input point k is the identity; bit 14: V' is
    static constexpr int PT = 2;                        // 14 affine inputs x 16 words
    static constexpr int VPA = PT + 16 * VP_COUNT;      // V' = V + r, affine (16)
    static constexpr int L = VPA + 16;                  // l[2]
    static constexpr int N = L + 16;                    // n[1]
    static constexpr int MERLIN = N + 8;                // 51 words (+1 pad)
    static constexpr int VP = MERLIN + 52;              // V' projective (PT_W)
    static constexpr int ZINV = VP + PT_W;              // FE_W
    static constexpr int COM = ZINV + FE_W;             // running WNLA commitment, projective (PT_W)
    static constexpr int ACC = COM + PT_W;              // fixed-base MSM result (PT_W)
    static constexpr int C = ACC + PT_W;                // c vector, 32 scalars
    static constexpr int RHO = C + 256;
    static constexpr int MU = RHO + 8;
    static constexpr int Y = MU + 8;                    // y_0..y_3
    static constexpr int FS = Y + 32;                   // fixed-base scalars (<= 49)
    static constexpr int VS = FS + 8 * NUM_GENS;        // variable-base scalars (<= 5)
    // ladder tables: 13 points (c_l c_r c_o c_s r[0..3] x[0..3] V') x 8 multiples, 4 field elements per entry:
    // projective X Y Z + the running product of the batch inversion while being built (word-major like the rest of the record)
    static constexpr int TAB = VS + 40;
    static constexpr int TAB_POINTS = 13, TAB_ENTRIES = TAB_POINTS * 8, TAB_STRIDE = 4 * FE_W;
    // the finished tables, array-of-structures: proof i, entry e at TABA * n + (i * TAB_ENTRIES + e) * TABA_STRIDE words
    // = x[8] y[8] (beta x)[8].  A ladder step picks its entry by a per-proof digit, so one entry must be one contiguous
    // 64-byte read ([x|y] for the k1 halves, [y|beta x] for the k2 halves) instead of 16 words scattered over 16 lines.
    static constexpr int TABA = (TAB + TAB_ENTRIES * TAB_STRIDE + 3) & ~3;
    static constexpr int TABA_STRIDE = 3 * FE_W;
    static constexpr int WORDS = TABA + TAB_ENTRIES * TABA_STRIDE;
};
// A ladder-table region of a per-proof record: `entries` table entries, built word-major at word offset `tab`
// (TAB_STRIDE words each), finished array-of-structures at word offset `taba` (TABA_STRIDE words each).  The verifier's
// region holds 13 points, the prover's (u64_prove.cuh) the two points of a WNLA round.
struct TabRegion { int tab, taba, entries; };
static constexpr int TAB_STRIDE_W = 4 * FE_W, TABA_STRIDE_W = 3 * FE_W;
BPPP_HD TabRegion vtab_region() { TabRegion r; r.tab = VL::TAB; r.taba = VL::TABA; r.entries = VL::TAB_ENTRIES; return r; }
BPPP_HD uint32_t *tab_entry(const WS &w, const TabRegion &R, size_t i, int entry) {
    return w.p + (size_t)R.taba * w.n + ((size_t)i * R.entries + entry) * TABA_STRIDE_W;
}
// 8 words <-> Fe through 16-byte accesses (the pointers are 16-byte aligned: TABA and the per-proof record size are
// multiples of 4 words, TABA_STRIDE of 8)
BPPP_HD Fe vtab_ld_fe(const uint32_t *p) {
    Fe r;
#if defined(__CUDA_ARCH__)
    uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#else
    for (int k = 0; k < 8; k++) r.v[k] = p[k];
#endif
    return r;
}
BPPP_HD void vtab_st_fe(uint32_t *p, const Fe &a) {
#if defined(__CUDA_ARCH__)
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
#else
    for (int k = 0; k < 8; k++) p[k] = a.v[k];
#endif
}
// table point ids: input slot k (1..12) -> k - 1; V' -> 12
BPPP_HD int vtab_of_slot(int slot) { return slot - 1; }
static constexpr int VTAB_VP = 12;

enum { FMT_COMPRESSED = 0, FMT_AFFINE64 = 1 };
static constexpr int U64_PROOF_BYTES_COMPRESSED = 525;          // 13*33 + 3*32 (README.md:30-34)
static constexpr int U64_PROOF_BYTES_AFFINE = 13 * 64 + 96;     // 928

BPPP_HD void set_status(const WS &w, size_t i, int32_t st) {
    // sticky: the first error wins
    int32_t cur = (int32_t)ws_ld(w, i, VL::STATUS);
    if (cur >= 0) ws_st(w, i, VL::STATUS, (uint32_t)st);
}

// Phase 0a: decode ONE point of one proof (one thread per (proof, point): 14x the parallelism of a per-proof loop, and the
// square roots of SEC1 decompression are the bulk of this phase).  STATUS / IDMASK must be pre-set by u64v_load_init.
// Record order (reciprocal::SerializableProof, reciprocal.rs:37-41 / circuit.rs:37-46):
//   c_l c_r c_o c_s | r[0..3] | x[0..3] | l[0..1] | n[0] | r
BPPP_HD void u64v_decode_point_one(const WS &w, size_t i, int k, const uint8_t *commit, const uint8_t *proof, int fmt, uint32_t *idmask_bit, bool *bad) {
    const int psz = fmt == FMT_COMPRESSED ? 33 : 64;
    const uint8_t *src;
    if (k == VP_V) src = commit;
    else if (k == VP_RR) src = proof + 12 * psz + 96;
    else src = proof + (k - 1) * psz;       // slots 1..12 are the first 12 record points in order
    PtA a;
    int s = fmt == FMT_COMPRESSED ? pta_decompress(a, src) : pta_from_xy64(a, src);
    *bad = s < 0;
    if (s < 0) { a.x = fe_zero(); a.y = fe_zero(); BPPP_SET_MAG(a.x, 1); BPPP_SET_MAG(a.y, 1); s = 1; }
    *idmask_bit = s == 1 ? (1u << k) : 0u;
    a.x = fe_normalize(a.x); a.y = fe_normalize(a.y);
    ws_st_pta(w, i, VL::PT + 16 * k, a);
}
// Phase 0b: scalars, V' = V + r (projective), status
BPPP_HD void u64v_load_finish_one(const WS &w, size_t i, const uint8_t *proof, int fmt, uint32_t idmask, bool bad_point) {
    const int psz = fmt == FMT_COMPRESSED ? 33 : 64;
    int32_t status = bad_point ? (int32_t)ST_BAD_POINT : (int32_t)ST_TRUE;
    const uint8_t *sc_src = proof + 12 * psz;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        Sc s;
        if (!sc_from_be32(s, sc_src + 32 * k)) { if (status >= 0) status = ST_BAD_SCALAR; s = sc_zero(); }
        ws_st_sc(w, i, VL::L + 8 * k, s);
    }
    PtA vpt = ws_ld_pta(w, i, VL::PT + 16 * VP_V), rpt = ws_ld_pta(w, i, VL::PT + 16 * VP_RR);
    Pt vp = pt_add(pt_from_affine(vpt, idmask & (1u << VP_V)), pt_from_affine(rpt, idmask & (1u << VP_RR)));   // reciprocal.rs:104
    ws_st_pt(w, i, VL::VP, vp);
    ws_st(w, i, VL::STATUS, (uint32_t)status);
    ws_st(w, i, VL::IDMASK, idmask);
}
// whole phase 0 for one proof (host emulation and small batches)
BPPP_HD void u64v_load_one(const WS &w, size_t i, const uint8_t *commit, const uint8_t *proof, int fmt) {
    uint32_t idmask = 0; bool bad = false;
#pragma unroll 1
    for (int k = 0; k < VP_COUNT; k++) { uint32_t bit; bool b; u64v_decode_point_one(w, i, k, commit, proof, fmt, &bit, &b); idmask |= bit; bad |= b; }
    u64v_load_finish_one(w, i, proof, fmt, idmask, bad);
}

// to_affine of a stored projective point with its batch-inverted Z
BPPP_HD PtA ws_affine(const WS &w, size_t i, int pt_off, int zinv_off, bool &is_identity) {
    Pt p = ws_ld_pt(w, i, pt_off);
    Fe zi = ws_ld_fe(w, i, zinv_off);
    is_identity = fe_is_zero(zi);
    return pt_to_affine_with_zinv(p, zi);
}

// Phase 1: transcript up to tau, all challenge-derived scalars for the circuit part.
// reciprocal.rs:99-100; circuit.rs:155-239 (closed forms, see header).
BPPP_HD void u64v_phase1_one(const WS &w, size_t i, const Merlin &init, const uint8_t *ext = nullptr) {
    Tx m; tx_init(m, init, ext);      // ext: e, rho, lambda, beta, delta, tau from a caller-owned transcript
    uint32_t idmask = ws_ld(w, i, VL::IDMASK);
    bool bad = false, zero_inv = false;
    // reciprocal.rs:99-100
    PtA V = ws_ld_pta(w, i, VL::PT + 16 * VP_V);
    tx_point(m, BPPP_LBL("reciprocal_commitment"), V, idmask & (1u << VP_V));
    Sc e; bad |= !tx_challenge(m, BPPP_LBL("reciprocal_challenge"), e);
    // circuit.rs:155-164
    bool vp_id;
    PtA vpa = ws_affine(w, i, VL::VP, VL::ZINV, vp_id);
    ws_st_pta(w, i, VL::VPA, vpa);
    if (vp_id) idmask |= 1u << 14;
    ws_st(w, i, VL::IDMASK, idmask);
    tx_point(m, BPPP_LBL("commitment_cl"), ws_ld_pta(w, i, VL::PT + 16 * VP_CL), idmask & (1u << VP_CL));
    tx_point(m, BPPP_LBL("commitment_cr"), ws_ld_pta(w, i, VL::PT + 16 * VP_CR), idmask & (1u << VP_CR));
    tx_point(m, BPPP_LBL("commitment_co"), ws_ld_pta(w, i, VL::PT + 16 * VP_CO), idmask & (1u << VP_CO));
    tx_point(m, BPPP_LBL("commitment_v"), vpa, vp_id);
    Sc rho, lambda, beta, delta, tau;
    bad |= !tx_challenge(m, BPPP_LBL("circuit_rho"), rho);
    bad |= !tx_challenge(m, BPPP_LBL("circuit_lambda"), lambda);
    bad |= !tx_challenge(m, BPPP_LBL("circuit_beta"), beta);
    bad |= !tx_challenge(m, BPPP_LBL("circuit_delta"), delta);
    // circuit.rs:189-191
    tx_point(m, BPPP_LBL("commitment_cs"), ws_ld_pta(w, i, VL::PT + 16 * VP_CS), idmask & (1u << VP_CS));
    bad |= !tx_challenge(m, BPPP_LBL("circuit_tau"), tau);
    tx_store(m, w, i, VL::MERLIN);

    Sc mu = sc_sqr(rho);
    // One inversion for {mu, tau, e+0 .. e+15} (Montgomery's trick); delta only has to be non-zero
    // (circuit.rs:196 unwraps its inverse; c_nO = c_lO = 0 so the value is never used).
    Sc inv[18], pre[18];
    inv[0] = mu; inv[1] = tau;
#pragma unroll 1
    for (int j = 0; j < 16; j++) inv[2 + j] = sc_add(e, sc_from_u64((uint64_t)j));
    zero_inv |= sc_is_zero(delta);
    Sc run = sc_one();
#pragma unroll 1
    for (int k = 0; k < 18; k++) {
        if (sc_is_zero(inv[k])) { zero_inv = true; inv[k] = sc_one(); }
        pre[k] = run; run = sc_mul(run, inv[k]);
    }
    Sc rinv = sc_inv(run);
#pragma unroll 1
    for (int k = 17; k >= 0; k--) { Sc t = sc_mul(rinv, pre[k]); rinv = sc_mul(rinv, inv[k]); inv[k] = t; }
    Sc mu_inv = inv[0], tau_inv = inv[1];
    Sc tau2 = sc_sqr(tau), tau3 = sc_mul(tau2, tau);

    // S = sum_{k=1..16} lambda^k, lambda powers kept for c_nR and c_l0
    Sc lp[16];
    Sc S = sc_zero(), cur = sc_one();
#pragma unroll 1
    for (int k = 0; k < 16; k++) { cur = sc_mul(cur, lambda); lp[k] = cur; S = sc_add(S, cur); }

    // pn_tau, ps_tau (circuit.rs:198-204)
    Sc ps = sc_zero(), musum = sc_zero();
    Sc mip = sc_one(), mp = sc_one();       // mu^-(j+1), mu^(j+1)
    Sc p16 = sc_one(), sixteen = sc_from_u64(16);
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        mip = sc_mul(mip, mu_inv); mp = sc_mul(mp, mu);
        Sc t1 = sc_mul(sc_mul(p16, mip), tau2);                       // -c_nL[j] tau^2
        Sc t2 = sc_mul(sc_add(sc_mul(sc_sub(S, lp[j]), mip), e), tau);  // c_nR[j] tau
        Sc pn = sc_add(t1, t2);
        ws_st_sc(w, i, VL::FS + 8 * (1 + j), pn);
        ps = sc_add(ps, sc_mul(sc_sqr(pn), mp));
        musum = sc_add(musum, mp);
        p16 = sc_mul(p16, sixteen);
    }
    ps = sc_sub(ps, sc_dbl(sc_mul(tau3, musum)));      // a_l = 0, a_m = 1 (reciprocal.rs:159,164)
    ws_st_sc(w, i, VL::FS, ps);

    // c = cr_tau || cl_tau || 0  (circuit.rs:208-239)
    Sc bt = sc_mul(beta, tau);
    ws_st_sc(w, i, VL::C + 0, sc_one());
    ws_st_sc(w, i, VL::C + 8, sc_mul(tau_inv, beta));
#pragma unroll 1
    for (int k = 2; k < 9; k++) { ws_st_sc(w, i, VL::C + 8 * k, bt); bt = sc_mul(bt, tau); }
    Sc s2 = sc_dbl(sc_mul(tau2, S));
#pragma unroll 1
    for (int j = 0; j < 16; j++) ws_st_sc(w, i, VL::C + 8 * (9 + j), sc_sub(sc_mul(s2, inv[2 + j]), lp[j]));
#pragma unroll 1
    for (int k = 25; k < 32; k++) ws_st_sc(w, i, VL::C + 8 * k, sc_zero());

    // commitment = pt + tau^-1 c_s - delta c_o + tau c_l - tau^2 c_r + tau^3 (2 V')   (circuit.rs:182-187,230-235)
    ws_st_sc(w, i, VL::VS + 0, tau_inv);
    ws_st_sc(w, i, VL::VS + 8, sc_neg(delta));
    ws_st_sc(w, i, VL::VS + 16, tau);
    ws_st_sc(w, i, VL::VS + 24, sc_neg(tau2));
    ws_st_sc(w, i, VL::VS + 32, sc_dbl(tau3));
    ws_st_sc(w, i, VL::RHO, rho);
    ws_st_sc(w, i, VL::MU, mu);
    if (bad) set_status(w, i, ST_PANIC_CHALLENGE_RANGE);
    if (zero_inv) set_status(w, i, ST_PANIC_INVERT_ZERO);
}

// joint variable-base sum_k ks[k] * pts[k] + init: GLV halves, signed 4-bit windows, 128 shared doublings
// (complete projective formulas, per-thread tables; used by the prover's re-commit and the generic fold kernel)
template <int NP>
BPPP_HD Pt straus_var(const PtA *pts, const bool *ident, const Sc *ks, const Pt &init) {
    PtTable8 tab[NP];
#pragma unroll 1
    for (int k = 0; k < NP; k++) pt_table8_build(tab[k], pt_from_affine(pts[k], ident[k]));
    return pt_add(straus_glv<NP>(tab, ks), init);   // init is added last (it must not be doubled)
}

// ---- verifier ladders over AFFINE tables in the workspace (Jacobian accumulator, ec.cuh) ----
// Phase T1: 1P..8P of one of the 13 per-proof points, projective, into the table region
BPPP_HD void u64v_table_build_one(const WS &w, size_t i, int t) {
    uint32_t idmask = ws_ld(w, i, VL::IDMASK);
    PtA a; bool id;
    if (t == VTAB_VP) { a = ws_ld_pta(w, i, VL::VPA); id = idmask & (1u << 14); }
    else { a = ws_ld_pta(w, i, VL::PT + 16 * (t + 1)); id = idmask & (1u << (t + 1)); }
    PtTable8 tab;
    pt_table8_build(tab, pt_from_affine(a, id));
#pragma unroll 1
    for (int e = 0; e < 8; e++) ws_st_pt(w, i, VL::TAB + (t * 8 + e) * VL::TAB_STRIDE, tab.m[e]);
}
// Phases T2+T3 fused: Montgomery batch inversion of every table entry's Z over this thread's strided share of the
// (entry, proof) items, writing the affine entries (array-of-structures region) on the way back
BPPP_HD void tables_normalize_strided(const WS &w, const TabRegion &R, size_t t, size_t T) {
    const size_t total = (size_t)R.entries * w.n;
    const Fe beta = fe_beta();
    Fe run = fe_one();
#pragma unroll 1
    for (size_t idx = t; idx < total; idx += T) {
        size_t e = idx / w.n, i = idx - e * w.n;
        const int off = R.tab + (int)e * TAB_STRIDE_W;
        Fe z = ws_ld_fe(w, i, off + 2 * FE_W);
        ws_st_fe(w, i, off + PT_W, run);                  // prefix product before this item
        if (!fe_normalizes_to_zero(z)) run = fe_mul(run, z);
    }
    Fe rinv = fe_inv(run);
    size_t cnt = total > t ? (total - t + T - 1) / T : 0;
#pragma unroll 1
    for (size_t k = cnt; k-- > 0;) {
        size_t idx = t + k * T;
        size_t e = idx / w.n, i = idx - e * w.n;
        const int off = R.tab + (int)e * TAB_STRIDE_W;
        Pt p = ws_ld_pt(w, i, off);
        PtA a;
        if (fe_normalizes_to_zero(p.z)) { a.x = fe_zero(); a.y = fe_zero(); }       // identity -> zero sentinel
        else {
            Fe zi = fe_mul(rinv, ws_ld_fe(w, i, off + PT_W));
            rinv = fe_mul(rinv, p.z);
            a = pt_to_affine_with_zinv(p, zi);
        }
        uint32_t *dst = tab_entry(w, R, i, (int)e);
        vtab_st_fe(dst, a.x); vtab_st_fe(dst + FE_W, a.y);
        vtab_st_fe(dst + 2 * FE_W, fe_normalize(fe_mul(a.x, beta)));     // x of the endomorphism image (beta x, y)
    }
}
BPPP_HD void u64v_tables_normalize_strided(const WS &w, size_t t, size_t T) { tables_normalize_strided(w, vtab_region(), t, T); }

// ---- the same tables built directly in AFFINE coordinates, level by level ----
// kP for k = 2..8 by affine additions / doublings whose denominators are inverted together ACROSS proofs (Montgomery's
// trick over a thread's strided share of the (operation, point, proof) items of one level).  Per entry: 2 M + 1 S for the
// chord formula (+ 1 S for a tangent), 3 M for its share of the inversion, 1 M for beta x -- against 8-12 M for the
// complete projective formula plus 9 M to normalise the entry afterwards (T1 + T2/T3 above: 148 M per point, here ~60 M).
// Three dependent levels: {2P}, {3P = 2P + P, 4P = 2 (2P)}, {5P = 4P + P, 6P = 2 (3P), 7P = 4P + 3P, 8P = 2 (4P)}.
// The group has prime order, so for an on-curve P != O no denominator vanishes (x_a = x_b needs (a -+ b) P = O, y = 0 a
// point of order two).  A vanishing denominator can therefore only come from bytes a failed decode left behind: it is kept
// out of the thread's shared product (other proofs are unaffected) and yields the identity sentinel.
struct AffOp { int a, b, out; };                                  // out P = a P + b P;  b == 0: out P = 2 (a P)
BPPP_HD int aff_level_nops(int level) { return level == 1 ? 1 : level == 2 ? 2 : 4; }
BPPP_HD AffOp aff_level_op(int level, int k) {
    AffOp r;
    if (level == 1) { r.a = 1; r.b = 0; r.out = 2; }
    else if (level == 2) { r.a = 2; r.b = k == 0 ? 1 : 0; r.out = k == 0 ? 3 : 4; }
    else if (k == 0) { r.a = 4; r.b = 1; r.out = 5; }
    else if (k == 1) { r.a = 3; r.b = 0; r.out = 6; }
    else if (k == 2) { r.a = 4; r.b = 3; r.out = 7; }
    else { r.a = 4; r.b = 0; r.out = 8; }
    return r;
}
BPPP_HD bool pta_is_sentinel(const PtA &a) { return fe_is_zero_canonical(a.x) && fe_is_zero_canonical(a.y); }
BPPP_HD PtA vtab_ld_pta(const uint32_t *ent) { PtA a; a.x = vtab_ld_fe(ent); a.y = vtab_ld_fe(ent + FE_W); return a; }
BPPP_HD void vtab_st_entry(uint32_t *dst, const PtA &a_canonical, const Fe &beta) {
    vtab_st_fe(dst, a_canonical.x); vtab_st_fe(dst + FE_W, a_canonical.y);
    vtab_st_fe(dst + 2 * FE_W, fe_normalize(fe_mul(a_canonical.x, beta)));
}
// operands of item (op, point p, proof i); false when the result is the identity (P = O).  Level 1 reads P from the
// record through SRC (the table does not hold it yet), the other levels read finished entries of earlier levels.
template <class SRC>
BPPP_HD bool aff_operands(const WS &w, const TabRegion &R, int level, size_t i, int p, const AffOp &op, const SRC &src, PtA &A, PtA &B) {
    if (level == 1) {
        if (src(w, i, p, A)) return false;
        B = A;
        return !pta_is_sentinel(A);
    }
    A = vtab_ld_pta(tab_entry(w, R, i, p * 8 + op.a - 1));
    B = op.b ? vtab_ld_pta(tab_entry(w, R, i, p * 8 + op.b - 1)) : A;
    return !(pta_is_sentinel(A) || pta_is_sentinel(B));
}
BPPP_HD Fe aff_denominator(const AffOp &op, const PtA &A, const PtA &B) { return op.b ? fe_sub(B.x, A.x) : fe_mul_int(A.y, 2); }
template <class SRC>
BPPP_HD void tables_affine_level(const WS &w, const TabRegion &R, int level, size_t t, size_t T, const SRC &src) {
    const int points = R.entries / 8, nops = aff_level_nops(level);
    const size_t per_op = (size_t)points * w.n, total = per_op * (size_t)nops;
    const Fe beta = fe_beta();
    Fe run = fe_one();
#pragma unroll 1
    for (size_t idx = t; idx < total; idx += T) {
        const int k = (int)(idx / per_op); const size_t rem = idx - (size_t)k * per_op;
        const int p = (int)(rem / w.n); const size_t i = rem - (size_t)p * w.n;
        const AffOp op = aff_level_op(level, k);
        PtA A, B;
        const bool live = aff_operands(w, R, level, i, p, op, src, A, B);
        if (level == 1) {                                         // entry 0 is P itself
            if (!live) { A.x = fe_zero(); A.y = fe_zero(); }
            vtab_st_entry(tab_entry(w, R, i, p * 8), A, beta);
        }
        if (!live) continue;
        const Fe d = aff_denominator(op, A, B);
        if (fe_normalizes_to_zero(d)) continue;
        ws_st_fe(w, i, R.tab + (p * 8 + op.out - 1) * TAB_STRIDE_W + PT_W, run);      // prefix product before this item
        run = fe_mul(run, d);
    }
    Fe rinv = fe_inv(run);
    const size_t cnt = total > t ? (total - t + T - 1) / T : 0;
#pragma unroll 1
    for (size_t c = cnt; c-- > 0;) {
        const size_t idx = t + c * T;
        const int k = (int)(idx / per_op); const size_t rem = idx - (size_t)k * per_op;
        const int p = (int)(rem / w.n); const size_t i = rem - (size_t)p * w.n;
        const AffOp op = aff_level_op(level, k);
        uint32_t *dst = tab_entry(w, R, i, p * 8 + op.out - 1);
        PtA A, B, r;
        const bool live = aff_operands(w, R, level, i, p, op, src, A, B);
        Fe d = fe_zero();
        if (live) d = aff_denominator(op, A, B);
        if (!live || fe_normalizes_to_zero(d)) {
            r.x = fe_zero(); r.y = fe_zero();
            vtab_st_entry(dst, r, beta);
            continue;
        }
        const Fe inv = fe_mul(rinv, ws_ld_fe(w, i, R.tab + (p * 8 + op.out - 1) * TAB_STRIDE_W + PT_W));
        rinv = fe_mul(rinv, d);
        const Fe lam = fe_mul(op.b ? fe_sub(B.y, A.y) : fe_mul_int(fe_sqr(A.x), 3), inv);
        r.x = fe_sub(fe_sub(fe_sqr(lam), A.x), B.x);               // B = A for a doubling
        r.y = fe_sub(fe_mul(lam, fe_sub(A.x, r.x)), A.y);
        r.x = fe_normalize(r.x); r.y = fe_normalize(r.y);
        vtab_st_entry(dst, r, beta);
    }
}
// the verifier's 13 table points: input slots 1..12 and V' (u64v_table_build_one reads the same)
struct VTabSource {
    BPPP_HD bool operator()(const WS &w, size_t i, int t, PtA &a) const {
        const uint32_t idmask = ws_ld(w, i, VL::IDMASK);
        if (t == VTAB_VP) { a = ws_ld_pta(w, i, VL::VPA); return (idmask & (1u << 14)) != 0; }
        a = ws_ld_pta(w, i, VL::PT + 16 * (t + 1));
        return (idmask & (1u << (t + 1))) != 0;
    }
};
BPPP_HD void u64v_tables_affine_level(const WS &w, int level, size_t t, size_t T) { tables_affine_level(w, vtab_region(), level, t, T, VTabSource()); }
// One lane's share of sum_k ks[k] * P_{tids[k]} from the affine tables: the 2 NP GLV halves are dealt round-robin to
// `nlanes` lanes (half h belongs to lane h % nlanes); every lane runs the full 128-doubling chain over its own halves.
// nlanes = 1 is the whole sum.  More lanes shorten the dependent chain of one proof (fewer additions per lane) at the
// price of repeating the doublings: used when the batch alone cannot fill the GPU (engine_var.cu).
template <int NP>
BPPP_HD PtJ straus_tables_partial(const WS &w, const TabRegion &R, size_t i, const int *tids, const Sc *ks, int lane, int nlanes) {
    Digits4h dg[2 * NP];
    bool neg[2 * NP];
#pragma unroll 1
    for (int k = 0; k < NP; k++) {
        if (nlanes > 1 && (2 * k) % nlanes != lane && (2 * k + 1) % nlanes != lane) continue;
        GlvSplit g = glv_split(ks[k]);
        dg[2 * k] = half_signed_digits4(g.k1); neg[2 * k] = g.neg1;
        dg[2 * k + 1] = half_signed_digits4(g.k2); neg[2 * k + 1] = g.neg2;
    }
    PtJ acc = ptj_identity();
#pragma unroll 1
    for (int d = 32; d >= 0; d--) {
        if (d != 32) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = ptj_double_hot(acc);
        }
#pragma unroll 1
        for (int h = lane; h < 2 * NP; h += nlanes) {
            int sd = digits4h_get(dg[h], d);
            if (neg[h]) sd = -sd;
            if (sd == 0) continue;
            int a = sd < 0 ? -sd : sd;
            const uint32_t *ent = tab_entry(w, R, i, tids[h >> 1] * 8 + (a - 1));
            PtA q;
            q.x = vtab_ld_fe(ent + ((h & 1) ? 2 * FE_W : 0));                           // odd halves use (beta x, y)
            q.y = vtab_ld_fe(ent + FE_W);
            if (fe_is_zero_canonical(q.x) && fe_is_zero_canonical(q.y)) continue;       // identity point: nothing to add
            if (sd < 0) q.y = fe_normalize_weak(fe_negate(q.y, 1));
            acc = ptj_add_mixed_hot(acc, q);
        }
    }
    return acc;
}
// acc = sum_k ks[k] * P_{tids[k]} + init from the affine tables
template <int NP>
BPPP_HD Pt straus_tables(const WS &w, const TabRegion &R, size_t i, const int *tids, const Sc *ks, const Pt &init) {
    return pt_add(ptj_to_pt(straus_tables_partial<NP>(w, R, i, tids, ks, 0, 1)), init);
}

// Phase 2b: com_0 = ACC (fixed part) + tau^-1 c_s - delta c_o + tau c_l - tau^2 c_r + 2 tau^3 V'
BPPP_HD void u64v_var5_one(const WS &w, size_t i) {
    const int tids[5] = {vtab_of_slot(VP_CS), vtab_of_slot(VP_CO), vtab_of_slot(VP_CL), vtab_of_slot(VP_CR), VTAB_VP};
    Sc ks[5];
#pragma unroll 1
    for (int k = 0; k < 5; k++) ks[k] = ws_ld_sc(w, i, VL::VS + 8 * k);
    ws_st_pt(w, i, VL::COM, straus_tables<5>(w, vtab_region(), i, tids, ks, ws_ld_pt(w, i, VL::ACC)));
}

// one lane's partial sum of the same five terms (the caller reduces the lanes and adds ACC)
BPPP_HD Pt u64v_var5_partial(const WS &w, size_t i, int lane, int nlanes) {
    const int tids[5] = {vtab_of_slot(VP_CS), vtab_of_slot(VP_CO), vtab_of_slot(VP_CL), vtab_of_slot(VP_CR), VTAB_VP};
    Sc ks[5];
#pragma unroll 1
    for (int k = 0; k < 5; k++) ks[k] = ws_ld_sc(w, i, VL::VS + 8 * k);
    return ptj_to_pt(straus_tables_partial<5>(w, vtab_region(), i, tids, ks, lane, nlanes));
}

// WNLA round j = 0..3 (wnla.rs:84-102): transcript -> y_j, fold c, scalars for com' = com + y X + (y^2-1) R
BPPP_HD void u64v_round_one(const WS &w, size_t i, int j, const uint8_t *ext = nullptr) {
    Tx m; tx_load(m, w, i, VL::MERLIN, ext);      // ext: y_j
    uint32_t idmask = ws_ld(w, i, VL::IDMASK);
    bool com_id;
    PtA com = ws_affine(w, i, VL::COM, VL::ZINV, com_id);
    const int xs = VP_X + (3 - j), rs = VP_R + (3 - j);     // proof.x.last(), proof.r.last() (wnla.rs:89-90)
    tx_point(m, BPPP_LBL("wnla_com"), com, com_id);
    tx_point(m, BPPP_LBL("wnla_x"), ws_ld_pta(w, i, VL::PT + 16 * xs), idmask & (1u << xs));
    tx_point(m, BPPP_LBL("wnla_r"), ws_ld_pta(w, i, VL::PT + 16 * rs), idmask & (1u << rs));
    tx_u64(m, BPPP_LBL("l.sz"), (uint64_t)(32 >> j));    // |h_vec| (wnla.rs:91)
    tx_u64(m, BPPP_LBL("n.sz"), (uint64_t)(16 >> j));    // |g_vec| (wnla.rs:92)
    Sc y;
    if (!tx_challenge(m, BPPP_LBL("wnla_challenge"), y)) set_status(w, i, ST_PANIC_CHALLENGE_RANGE);
    tx_store(m, w, i, VL::MERLIN);
    ws_st_sc(w, i, VL::Y + 8 * j, y);
    const int half = (32 >> j) / 2;
#pragma unroll 1
    for (int k = 0; k < half; k++) {     // c' = c0 + y c1 (wnla.rs:98)
        Sc c0 = ws_ld_sc(w, i, VL::C + 8 * (2 * k)), c1 = ws_ld_sc(w, i, VL::C + 8 * (2 * k + 1));
        ws_st_sc(w, i, VL::C + 8 * k, sc_add(c0, sc_mul(y, c1)));
    }
    ws_st_sc(w, i, VL::VS + 0, y);
    ws_st_sc(w, i, VL::VS + 8, sc_sub(sc_sqr(y), sc_one()));
}

// com' = com + y X + (y^2 - 1) R  (wnla.rs:100-102)
BPPP_HD void u64v_var2_one(const WS &w, size_t i, int j) {
    const int tids[2] = {vtab_of_slot(VP_X + (3 - j)), vtab_of_slot(VP_R + (3 - j))};
    Sc ks[2] = {ws_ld_sc(w, i, VL::VS), ws_ld_sc(w, i, VL::VS + 8)};
    ws_st_pt(w, i, VL::COM, straus_tables<2>(w, vtab_region(), i, tids, ks, ws_ld_pt(w, i, VL::COM)));
}

BPPP_HD Pt u64v_var2_partial(const WS &w, size_t i, int j, int lane, int nlanes) {
    const int tids[2] = {vtab_of_slot(VP_X + (3 - j)), vtab_of_slot(VP_R + (3 - j))};
    Sc ks[2] = {ws_ld_sc(w, i, VL::VS), ws_ld_sc(w, i, VL::VS + 8)};
    return ptj_to_pt(straus_tables_partial<2>(w, vtab_region(), i, tids, ks, lane, nlanes));
}

// ---- the same ladders cut into SEGMENTS of consecutive windows (engine_var.cu:k_v_var_seg) ----
// A 65,536-proof batch is 2,048 one-warp ladders on 2,368 warp slots: the schedulers that received four of them finish a fifth
// later than those with three, which then idle.  Cut into segments that any warp may continue (state handed over through the
// free build rows of the table region), the ladders are re-dealt to whichever slot frees first and the SMs finish together.
// Scratch rows (word-major, relative to the region's build area): accumulator X Y Z + infinity flag, the GLV halves of the
// group's scalars (so that the split runs once), one row of synchronisation words.
static constexpr int LS_ACC = 0, LS_INF = 24, LS_NEG = 25, LS_GLV = 32, LS_SYNC = 80;
BPPP_HD uint32_t ls_ld(const WS &w, size_t i, int row) {
#if defined(__CUDA_ARCH__)
    return __ldcg(w.p + (size_t)row * w.n + i);         // written by another SM a moment ago: read past the L1
#else
    return w.p[(size_t)row * w.n + i];
#endif
}
BPPP_HD void ls_st(const WS &w, size_t i, int row, uint32_t v) { w.p[(size_t)row * w.n + i] = v; }
// windows d_hi .. d_lo (inclusive, descending) of the joint ladder over 2 NP half-scalars; same steps as straus_tables_partial
template <int NP>
BPPP_HD PtJ straus_tables_windows(const WS &w, const TabRegion &R, size_t i, const int *tids, const Digits4h *dg, const bool *neg, PtJ acc, int d_hi, int d_lo) {
#pragma unroll 1
    for (int d = d_hi; d >= d_lo; d--) {
        if (d != 32) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = ptj_double_hot(acc);
        }
#pragma unroll 1
        for (int h = 0; h < 2 * NP; h++) {
            int sd = digits4h_get(dg[h], d);
            if (neg[h]) sd = -sd;
            if (sd == 0) continue;
            int a = sd < 0 ? -sd : sd;
            const uint32_t *ent = tab_entry(w, R, i, tids[h >> 1] * 8 + (a - 1));
            PtA q;
            q.x = vtab_ld_fe(ent + ((h & 1) ? 2 * FE_W : 0));
            q.y = vtab_ld_fe(ent + FE_W);
            if (fe_is_zero_canonical(q.x) && fe_is_zero_canonical(q.y)) continue;
            if (sd < 0) q.y = fe_normalize_weak(fe_negate(q.y, 1));
            acc = ptj_add_mixed_hot(acc, q);
        }
    }
    return acc;
}
// segment `seg` of `nseg`: the 33 windows are dealt top-down in runs of 33 / nseg (+1).  Returns true when acc is the finished
// sum (last segment); otherwise acc has been parked in the scratch rows.  ks is read by segment 0 only.
template <int NP>
BPPP_HD bool straus_tables_seg(const WS &w, const TabRegion &R, size_t i, const int *tids, const Sc *ks, int seg, int nseg, PtJ &acc) {
    Digits4h dg[2 * NP];
    bool neg[2 * NP];
    const int sb = R.tab;
    if (seg == 0) {
        uint32_t negmask = 0;
#pragma unroll 1
        for (int k = 0; k < NP; k++) {
            GlvSplit g = glv_split(ks[k]);
            dg[2 * k] = half_signed_digits4(g.k1); neg[2 * k] = g.neg1;
            dg[2 * k + 1] = half_signed_digits4(g.k2); neg[2 * k + 1] = g.neg2;
            negmask |= (g.neg1 ? 1u : 0u) << (2 * k) | (g.neg2 ? 1u : 0u) << (2 * k + 1);
            if (nseg > 1) {
#pragma unroll
                for (int t = 0; t < 4; t++) { ls_st(w, i, sb + LS_GLV + 8 * k + t, g.k1[t]); ls_st(w, i, sb + LS_GLV + 8 * k + 4 + t, g.k2[t]); }
            }
        }
        if (nseg > 1) ls_st(w, i, sb + LS_NEG, negmask);
        acc = ptj_identity();
    } else {
        const uint32_t negmask = ls_ld(w, i, sb + LS_NEG);
#pragma unroll 1
        for (int h = 0; h < 2 * NP; h++) {
            uint32_t m[4];
#pragma unroll
            for (int t = 0; t < 4; t++) m[t] = ls_ld(w, i, sb + LS_GLV + 4 * h + t);
            dg[h] = half_signed_digits4(m); neg[h] = (negmask >> h) & 1u;
        }
#pragma unroll
        for (int t = 0; t < FE_W; t++) {
            acc.x.v[t] = ls_ld(w, i, sb + LS_ACC + t); acc.y.v[t] = ls_ld(w, i, sb + LS_ACC + FE_W + t); acc.z.v[t] = ls_ld(w, i, sb + LS_ACC + 2 * FE_W + t);
        }
        acc.inf = ls_ld(w, i, sb + LS_INF) != 0;
    }
    const int hi = 32 - (33 * seg) / nseg, lo = 32 - (33 * (seg + 1)) / nseg + 1;
    acc = straus_tables_windows<NP>(w, R, i, tids, dg, neg, acc, hi, lo);
    if (seg + 1 == nseg) return true;
#pragma unroll
    for (int t = 0; t < FE_W; t++) {
        ls_st(w, i, sb + LS_ACC + t, acc.x.v[t]); ls_st(w, i, sb + LS_ACC + FE_W + t, acc.y.v[t]); ls_st(w, i, sb + LS_ACC + 2 * FE_W + t, acc.z.v[t]);
    }
    ls_st(w, i, sb + LS_INF, acc.inf ? 1u : 0u);
    return false;
}
BPPP_HD void u64v_var5_seg(const WS &w, size_t i, int seg, int nseg) {
    const int tids[5] = {vtab_of_slot(VP_CS), vtab_of_slot(VP_CO), vtab_of_slot(VP_CL), vtab_of_slot(VP_CR), VTAB_VP};
    Sc ks[5];
    if (seg == 0) {
#pragma unroll 1
        for (int k = 0; k < 5; k++) ks[k] = ws_ld_sc(w, i, VL::VS + 8 * k);
    }
    PtJ acc;
    if (straus_tables_seg<5>(w, vtab_region(), i, tids, ks, seg, nseg, acc)) ws_st_pt(w, i, VL::COM, pt_add(ptj_to_pt(acc), ws_ld_pt(w, i, VL::ACC)));
}
BPPP_HD void u64v_var2_seg(const WS &w, size_t i, int j, int seg, int nseg) {
    const int tids[2] = {vtab_of_slot(VP_X + (3 - j)), vtab_of_slot(VP_R + (3 - j))};
    Sc ks[2];
    if (seg == 0) { ks[0] = ws_ld_sc(w, i, VL::VS); ks[1] = ws_ld_sc(w, i, VL::VS + 8); }
    PtJ acc;
    if (straus_tables_seg<2>(w, vtab_region(), i, tids, ks, seg, nseg, acc)) ws_st_pt(w, i, VL::COM, pt_add(ptj_to_pt(acc), ws_ld_pt(w, i, VL::COM)));
}

// Base case (wnla.rs:80-82): scalars of commit(l, n) over the ORIGINAL generators.  After 4 folds
//   h^(4)_s = sum_t (prod_k y_k^bit_k(t)) h_{16 s + t},  g^(4)_0 = sum_t (prod_k (bit_k(t) ? y_k : rho_k)) g_t
// with rho_0 = rho, rho_{k+1} = mu_k, mu_{k+1} = mu_k^2 (wnla.rs:96-97,108-109).
BPPP_HD void u64v_final_scalars_one(const WS &w, size_t i) {
    Sc y[4], rk[4];
#pragma unroll 1
    for (int k = 0; k < 4; k++) y[k] = ws_ld_sc(w, i, VL::Y + 8 * k);
    Sc mu = ws_ld_sc(w, i, VL::MU);
    rk[0] = ws_ld_sc(w, i, VL::RHO);
#pragma unroll 1
    for (int k = 1; k < 4; k++) { rk[k] = mu; mu = sc_sqr(mu); }
    mu = sc_sqr(mu);   // mu_4
    Sc l0 = ws_ld_sc(w, i, VL::L), l1 = ws_ld_sc(w, i, VL::L + 8), n0 = ws_ld_sc(w, i, VL::N);
    Sc c0 = ws_ld_sc(w, i, VL::C), c1 = ws_ld_sc(w, i, VL::C + 8);
    // v = <c, l> + |n|^2_mu (wnla.rs:67)
    Sc v = sc_add(sc_add(sc_mul(c0, l0), sc_mul(c1, l1)), sc_mul(sc_sqr(n0), mu));
    ws_st_sc(w, i, VL::FS, v);
    Sc yp[16], gp[16];
    yp[0] = sc_one(); gp[0] = sc_one();
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        int span = 1 << k;
#pragma unroll 1
        for (int t = 0; t < span; t++) {
            yp[t + span] = sc_mul(yp[t], y[k]);
            gp[t + span] = sc_mul(gp[t], y[k]);
            gp[t] = sc_mul(gp[t], rk[k]);
        }
    }
#pragma unroll 1
    for (int t = 0; t < 16; t++) {
        ws_st_sc(w, i, VL::FS + 8 * (GEN_GVEC + t), sc_mul(n0, gp[t]));
        ws_st_sc(w, i, VL::FS + 8 * (GEN_HVEC + t), sc_mul(l0, yp[t]));
        ws_st_sc(w, i, VL::FS + 8 * (GEN_HVEC + 16 + t), sc_mul(l1, yp[t]));
    }
}

// verdict: commitment == commit(l, n)  (wnla.rs:81)
BPPP_HD void u64v_verdict_one(const WS &w, size_t i) {
    Pt com = ws_ld_pt(w, i, VL::COM), f = ws_ld_pt(w, i, VL::ACC);
    int32_t st = (int32_t)ws_ld(w, i, VL::STATUS);
    if (st >= 0) ws_st(w, i, VL::STATUS, pt_equal(com, f) ? ST_TRUE : ST_FALSE);
}

}  // namespace bppp
