// U64RangeProofProtocol::prove (reference src/range_proof/u64_proof.rs:57-82) for a batch of independent
// witnesses in lockstep: per-proof phase logic between multi-scalar multiplications.  Call chain restated:
// reciprocal.rs:110-146 -> circuit.rs:260-556 -> wnla.rs:125-190.
//
// Exact algebraic restructuring (all identities over F_n / the group, so outputs are bit-identical):
//  * reciprocal-circuit coefficient vectors in closed form (see u64_verify.cuh header);
//    with them no = lo = lr = 0, c_nO = c_lR = c_lO = 0, ll = multiplicities || 0 (circuit.rs:303-333);
//  * circuit.commit(v, s + r_blind) == V + r_com (circuit.rs:146-151 vs reciprocal.rs:88-95);
//  * folded generators are never materialised: round-j X_j / R_j (wnla.rs:152-160) are multi-scalar
//    multiplications over the ORIGINAL 49 generators with scalars rescaled by
//      coefH_j(t) = prod_{k<j} y_k^bit_k(t),  coefG_j(t) = prod_{k<j} (bit_k(t) ? y_k : rho_k)
//    so every prover point is fixed-base (window tables);
//  * the per-round re-commit wnla.commit(l', n') (wnla.rs:186) equals C + y X + (y^2 - 1) R (wnla.rs:100-102).
#pragma once
#include "u64_verify.cuh"

namespace bppp {

static constexpr int U64_RNG_BYTES = 52 * 64;
static constexpr int U64_RNG_EARLY = 19;     // scalars drawn by the first transcript phase (u64p_phase1_one); the other 33 by the second

// prover point slots (projective in PL::PTS, 1/Z in PL::ZINV)
enum { PP_V = 0, PP_RCOM = 1, PP_CO = 2, PP_CL = 3, PP_CR = 4, PP_VP = 5, PP_CS = 6, PP_X = 7 /*X_0..X_3*/, PP_R = 11 /*R_0..R_3*/, PP_COM = 15, PP_COUNT = 16 };
// output record point order: c_l c_r c_o c_s r[0..3] x[0..3] r_com
enum { PO_CL = 0, PO_CR = 1, PO_CO = 2, PO_CS = 3, PO_R = 4, PO_X = 8, PO_RCOM = 12, PO_COUNT = 13 };
// random draws in reference order: index into PL::RND
enum { RD_RBLIND = 0, RD_RO = 1 /*7*/, RD_RL = 8 /*6*/, RD_RR = 14 /*5*/, RD_LS = 19 /*17*/, RD_NS = 36 /*16*/, RD_COUNT = 52 };

struct PL {
    static constexpr int STATUS = 0;
    static constexpr int X = 1;                               // u64 value, 2 words
    static constexpr int S = 3;                               // blinding
    static constexpr int MERLIN = S + 8;
    static constexpr int PTS = MERLIN + 52;
    static constexpr int COM = PTS + PT_W * PP_COM;
    static constexpr int ZINV = PTS + PT_W * PP_COUNT;
    static constexpr int OUT = ZINV + FE_W * PP_COUNT;          // 13 x (8 x-words + tag)
    static constexpr int RND = OUT + 9 * PO_COUNT;            // 52 reduced draws
    static constexpr int INVE = RND + 8 * RD_COUNT;           // 1/(e+j), j < 16
    static constexpr int E = INVE + 128;
    static constexpr int RHO = E + 8;
    static constexpr int RHOINV = RHO + 8;
    static constexpr int MU = RHOINV + 8;
    static constexpr int LAMBDA = MU + 8;
    static constexpr int BETA = LAMBDA + 8;
    static constexpr int DELTA = BETA + 8;
    static constexpr int RS9 = DELTA + 8;                     // rs[9]
    static constexpr int LV = RS9 + 72;                       // l, 32
    static constexpr int NV = LV + 256;                       // n, 16
    static constexpr int C = NV + 128;                        // c, 32
    static constexpr int Y = C + 256;                         // y_0..y_3
    static constexpr int VS = Y + 32;                         // y, y^2-1
    static constexpr int FS = VS + 16;                        // fixed-base scalars; stage 1 uses 83 (FS..XS)
    static constexpr int XS = FS + 8 * NUM_GENS;              // X_j scalars (49), directly after FS
    static constexpr int RS = XS + 8 * NUM_GENS;              // R_j scalars (25)
    // ladder tables of the round's two re-commit points X_j, R_j (u64_verify.cuh:TabRegion)
    static constexpr int TAB = (RS + 8 * 25 + 3) & ~3;
    static constexpr int TAB_ENTRIES = 16;
    static constexpr int TABA = TAB + TAB_ENTRIES * TAB_STRIDE_W;
    static constexpr int WORDS = TABA + TAB_ENTRIES * TABA_STRIDE_W;
};
BPPP_HD TabRegion ptab_region() { TabRegion r; r.tab = PL::TAB; r.taba = PL::TABA; r.entries = PL::TAB_ENTRIES; return r; }

BPPP_HD void pset_status(const WS &w, size_t i, int32_t st) {
    int32_t cur = (int32_t)ws_ld(w, i, PL::STATUS);
    if (cur >= 0) ws_st(w, i, PL::STATUS, (uint32_t)st);
}

// ---- term maps (host + device) ----
BPPP_HD void u64p_termmap_commit(int *gen) { gen[0] = GEN_G; gen[1] = GEN_HVEC; }
BPPP_HD int u64p_stage1_scalar_base(int k) { return k == 0 ? 0 : (k == 1 ? 17 : (k == 2 ? 24 : 62)); }
BPPP_HD int u64p_stage1_point(int k) { return k == 0 ? PP_RCOM : (k == 1 ? PP_CO : (k == 2 ? PP_CL : PP_CR)); }
BPPP_HD int u64p_stage1_norm_point(int k) { return k == 4 ? PP_VP : u64p_stage1_point(k); }
// non-zero blinder positions (circuit.rs:264-298)
BPPP_HD int ro_pos(int k) { const int p[7] = {0, 1, 2, 3, 5, 6, 7}; return p[k]; }
BPPP_HD int rl_pos(int k) { const int p[6] = {0, 1, 2, 4, 5, 6}; return p[k]; }
BPPP_HD int rr_pos(int k) { const int p[5] = {0, 1, 3, 4, 5}; return p[k]; }
BPPP_HD int u64p_termmap_stage1(int *gen, int k) {
    int n = 0;
    if (k == 0) {            // r_com = r_blind h_0 + <h[9..], r>   (reciprocal.rs:93-95)
        gen[n++] = GEN_HVEC;
        for (int j = 0; j < 16; j++) gen[n++] = GEN_HVEC + 9 + j;
    } else if (k == 1) {     // c_o = <h, ro || lo> + <g_vec, no>, lo = no = 0   (circuit.rs:335-337)
        for (int j = 0; j < 7; j++) gen[n++] = GEN_HVEC + ro_pos(j);
    } else if (k == 2) {     // c_l = <h, rl || ll> + <g_vec, nl>   (circuit.rs:339-341)
        for (int j = 0; j < 6; j++) gen[n++] = GEN_HVEC + rl_pos(j);
        for (int j = 0; j < 16; j++) gen[n++] = GEN_HVEC + 9 + j;
        for (int j = 0; j < 16; j++) gen[n++] = GEN_GVEC + j;
    } else {                 // c_r = <h, rr || lr> + <g_vec, nr>, lr = 0   (circuit.rs:343-345)
        for (int j = 0; j < 5; j++) gen[n++] = GEN_HVEC + rr_pos(j);
        for (int j = 0; j < 16; j++) gen[n++] = GEN_GVEC + j;
    }
    return n;
}
BPPP_HD void u64p_termmap_cs(int *gen) {   // c_s = <h[0..26), rs || ls> + <g_vec, ns>   (circuit.rs:469-470)
    for (int j = 0; j < 26; j++) gen[j] = GEN_HVEC + j;
    for (int j = 0; j < 16; j++) gen[26 + j] = GEN_GVEC + j;
}
BPPP_HD void u64p_termmap_c0(int *gen) {   // v g + <h[0..26), l> + <g_vec, n>   (circuit.rs:522-524)
    gen[0] = GEN_G;
    for (int j = 0; j < 26; j++) gen[1 + j] = GEN_HVEC + j;
    for (int j = 0; j < 16; j++) gen[27 + j] = GEN_GVEC + j;
}
BPPP_HD void u64p_termmap_r(int *gen, int j) {   // R_j touches only generators whose folded index is odd
    int n = 0;
    gen[n++] = GEN_G;
    for (int idx = 0; idx < 32; idx++) if ((idx >> j) & 1) gen[n++] = GEN_HVEC + idx;
    for (int idx = 0; idx < 16; idx++) if ((idx >> j) & 1) gen[n++] = GEN_GVEC + idx;
}

BPPP_HD uint64_t u64p_ld_x(const WS &w, size_t i) { return (uint64_t)ws_ld(w, i, PL::X) | ((uint64_t)ws_ld(w, i, PL::X + 1) << 32); }
BPPP_HD uint32_t u64_digit(uint64_t x, int k) { return (uint32_t)(x >> (4 * k)) & 15u; }       // u64_to_hex (u64_proof.rs:84-90)
BPPP_HD uint32_t u64_multiplicity(uint64_t x, uint32_t d) {                                       // u64_to_hex_mapped (:92-102)
    uint32_t c = 0;
#pragma unroll 1
    for (int k = 0; k < 16; k++) c += u64_digit(x, k) == d;
    return c;
}

BPPP_HD void u64p_load_one(const WS &w, size_t i, uint64_t x, const uint8_t *blind) {
    Sc s; int32_t st = ST_TRUE;
    if (!sc_from_be32(s, blind)) { st = ST_BAD_SCALAR; s = sc_zero(); }
    ws_st(w, i, PL::STATUS, (uint32_t)st);
    ws_st(w, i, PL::X, (uint32_t)x); ws_st(w, i, PL::X + 1, (uint32_t)(x >> 32));
    ws_st_sc(w, i, PL::S, s);
    ws_st_sc(w, i, PL::FS, sc_from_u64(x));     // commit_value: x g + s h_0 (reciprocal.rs:88-90)
    ws_st_sc(w, i, PL::FS + 8, s);
}

// affine + SEC1 record of prover point `slot`; also stashes the compressed form for output position `po` (or -1)
BPPP_HD PtA u64p_affine(const WS &w, size_t i, int slot, int po, bool &id) {
    PtA a = ws_affine(w, i, PL::PTS + PT_W * slot, PL::ZINV + FE_W * slot, id);
    if (po >= 0) {
        uint32_t xw[8];
        fe_to_words(xw, a.x);
#pragma unroll
        for (int k = 0; k < 8; k++) ws_st(w, i, PL::OUT + 9 * po + k, id ? 0u : xw[k]);
        ws_st(w, i, PL::OUT + 9 * po + 8, id ? 0u : (2u + (a.y.v[0] & 1u)));
    }
    return a;
}

// Phase 1 (reciprocal.rs:114-122; circuit.rs:264-345): e, reciprocals, blinders, scalars of r_com, c_o, c_l, c_r
BPPP_HD void u64p_phase1_one(const WS &w, size_t i, const Merlin &init, const uint8_t *rng, const uint8_t *ext = nullptr) {
    Tx m; tx_init(m, init, ext);      // ext: e
    bool id, bad = false, zero_inv = false;
    PtA V = u64p_affine(w, i, PP_V, -1, id);
    tx_point(m, BPPP_LBL("reciprocal_commitment"), V, id);
    Sc e; bad |= !tx_challenge(m, BPPP_LBL("reciprocal_challenge"), e);
    tx_store(m, w, i, PL::MERLIN);
    ws_st_sc(w, i, PL::E, e);
    // 1/(e + j), j < 16: the 16 distinct inverses behind both r_i = 1/(d_i + e) (reciprocal.rs:117-119)
    // and W_l's pole columns (reciprocal.rs:179-183).  One inversion (Montgomery's trick).
    Sc inv[16], pre[16];
    Sc run = sc_one();
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        inv[j] = sc_add(e, sc_from_u64((uint64_t)j));
        if (sc_is_zero(inv[j])) { zero_inv = true; inv[j] = sc_one(); }
        pre[j] = run; run = sc_mul(run, inv[j]);
    }
    Sc rinv = sc_inv(run);
#pragma unroll 1
    for (int j = 15; j >= 0; j--) { Sc t = sc_mul(rinv, pre[j]); rinv = sc_mul(rinv, inv[j]); inv[j] = t; ws_st_sc(w, i, PL::INVE + 8 * j, t); }
    // draws 1..19 (reciprocal.rs:121; circuit.rs:264-298)
#pragma unroll 1
    for (int k = 0; k < U64_RNG_EARLY; k++) ws_st_sc(w, i, PL::RND + 8 * k, sc_from_wide_be64(rng + 64 * k));
    uint64_t x = u64p_ld_x(w, i);
    // r_com scalars: [r_blind, r_0..r_15]
    int o = PL::FS;
    ws_st_sc(w, i, o, ws_ld_sc(w, i, PL::RND + 8 * RD_RBLIND)); o += 8;
#pragma unroll 1
    for (int k = 0; k < 16; k++) { ws_st_sc(w, i, o, inv[u64_digit(x, k)]); o += 8; }
    // c_o scalars: ro (7)
#pragma unroll 1
    for (int k = 0; k < 7; k++) { ws_st_sc(w, i, o, ws_ld_sc(w, i, PL::RND + 8 * (RD_RO + k))); o += 8; }
    // c_l scalars: rl (6), multiplicities (16), digits (16)
#pragma unroll 1
    for (int k = 0; k < 6; k++) { ws_st_sc(w, i, o, ws_ld_sc(w, i, PL::RND + 8 * (RD_RL + k))); o += 8; }
#pragma unroll 1
    for (int d = 0; d < 16; d++) { ws_st_sc(w, i, o, sc_from_u64(u64_multiplicity(x, (uint32_t)d))); o += 8; }
#pragma unroll 1
    for (int k = 0; k < 16; k++) { ws_st_sc(w, i, o, sc_from_u64(u64_digit(x, k))); o += 8; }
    // c_r scalars: rr (5), reciprocals (16)
#pragma unroll 1
    for (int k = 0; k < 5; k++) { ws_st_sc(w, i, o, ws_ld_sc(w, i, PL::RND + 8 * (RD_RR + k))); o += 8; }
#pragma unroll 1
    for (int k = 0; k < 16; k++) { ws_st_sc(w, i, o, inv[u64_digit(x, k)]); o += 8; }
    if (bad) pset_status(w, i, ST_PANIC_CHALLENGE_RANGE);
    if (zero_inv) pset_status(w, i, ST_PANIC_INVERT_ZERO);
}

// circuit_commitment = circuit.commit(v, s + r_blind) == V + r_com
BPPP_HD void u64p_vprime_one(const WS &w, size_t i) {
    ws_st_pt(w, i, PL::PTS + PT_W * PP_VP, pt_add(ws_ld_pt(w, i, PL::PTS + PT_W * PP_V), ws_ld_pt(w, i, PL::PTS + PT_W * PP_RCOM)));
}

struct CircuitCoefs {   // closed-form c_nL, c_nR, c_lL (see u64_verify.cuh header)
    Sc nL[16], nR[16], lL[16], lp[16] /*lambda^(j+1)*/, mp[16] /*mu^(j+1)*/;
};
BPPP_HD void u64_circuit_coefs(CircuitCoefs &cc, const WS &w, size_t i, const Sc &e, const Sc &lambda, const Sc &mu, const Sc &mu_inv) {
    Sc S = sc_zero(), cur = sc_one();
#pragma unroll 1
    for (int k = 0; k < 16; k++) { cur = sc_mul(cur, lambda); cc.lp[k] = cur; S = sc_add(S, cur); }
    Sc mip = sc_one(), mp = sc_one(), p16 = sc_one(), sixteen = sc_from_u64(16);
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        mip = sc_mul(mip, mu_inv); mp = sc_mul(mp, mu);
        cc.mp[j] = mp;
        cc.nL[j] = sc_neg(sc_mul(p16, mip));
        cc.nR[j] = sc_add(sc_mul(sc_sub(S, cc.lp[j]), mip), e);
        cc.lL[j] = sc_neg(sc_mul(S, ws_ld_sc(w, i, PL::INVE + 8 * j)));
        p16 = sc_mul(p16, sixteen);
    }
}

// Phase 2 (circuit.rs:347-470): rho, lambda, beta, delta; ls, ns; f_[0..8); rs; scalars of c_s
BPPP_HD void u64p_phase2_one(const WS &w, size_t i, const uint8_t *rng, const uint8_t *ext = nullptr) {
    Tx m; tx_load(m, w, i, PL::MERLIN, ext);      // ext: rho, lambda, beta, delta
    bool id, bad = false, zero_inv = false;
    PtA a;
    a = u64p_affine(w, i, PP_CL, PO_CL, id); tx_point(m, BPPP_LBL("commitment_cl"), a, id);
    a = u64p_affine(w, i, PP_CR, PO_CR, id); tx_point(m, BPPP_LBL("commitment_cr"), a, id);
    a = u64p_affine(w, i, PP_CO, PO_CO, id); tx_point(m, BPPP_LBL("commitment_co"), a, id);
    a = u64p_affine(w, i, PP_VP, -1, id);    tx_point(m, BPPP_LBL("commitment_v"), a, id);
    (void)u64p_affine(w, i, PP_RCOM, PO_RCOM, id);
    Sc rho, lambda, beta, delta;
    bad |= !tx_challenge(m, BPPP_LBL("circuit_rho"), rho);
    bad |= !tx_challenge(m, BPPP_LBL("circuit_lambda"), lambda);
    bad |= !tx_challenge(m, BPPP_LBL("circuit_beta"), beta);
    bad |= !tx_challenge(m, BPPP_LBL("circuit_delta"), delta);
    tx_store(m, w, i, PL::MERLIN);
    // draws 20..52: ls (17), ns (16)   (circuit.rs:371-372)
#pragma unroll 1
    for (int k = U64_RNG_EARLY; k < 52; k++) ws_st_sc(w, i, PL::RND + 8 * k, sc_from_wide_be64(rng + 64 * k));
    // inverses of rho and beta with one inversion; mu^-1 = rho^-2 (util.rs:119 inverts mu; circuit.rs:403,455 delta, beta)
    Sc mu = sc_sqr(rho);
    zero_inv |= sc_is_zero(rho) | sc_is_zero(beta) | sc_is_zero(delta);
    Sc rb = sc_mul(sc_is_zero(rho) ? sc_one() : rho, sc_is_zero(beta) ? sc_one() : beta);
    Sc rbi = sc_inv(rb);
    Sc rho_inv = sc_mul(rbi, beta), beta_inv = sc_mul(rbi, rho);
    Sc mu_inv = sc_sqr(rho_inv);
    Sc e = ws_ld_sc(w, i, PL::E);
    CircuitCoefs cc;
    u64_circuit_coefs(cc, w, i, e, lambda, mu, mu_inv);
    uint64_t x = u64p_ld_x(w, i);
    // f_ coefficients with the structural zeros removed (circuit.rs:399-453)
    Sc f0 = sc_zero(), f1 = sc_zero(), f2 = sc_zero(), f3a = sc_zero(), f3b = sc_zero(), f3c = sc_zero(), f4 = sc_zero(), f5 = sc_zero(), f6 = sc_zero();
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        Sc ns = ws_ld_sc(w, i, PL::RND + 8 * (RD_NS + j)), ls = ws_ld_sc(w, i, PL::RND + 8 * (RD_LS + j));
        Sc nl = sc_from_u64(u64_digit(x, j));                       // w_l = digits
        Sc nr = ws_ld_sc(w, i, PL::INVE + 8 * (int)u64_digit(x, j));  // w_r = reciprocals
        Sc nsw = sc_mul(ns, cc.mp[j]);                              // ns_j mu^(j+1)
        Sc a1 = sc_add(nl, cc.nR[j]), a2 = sc_add(nr, cc.nL[j]);
        f0 = sc_add(f0, sc_mul(ns, nsw));
        f1 = sc_add(f1, sc_mul(cc.lp[j], ls));
        f2 = sc_add(f2, sc_mul(nsw, a1));
        f3a = sc_add(f3a, sc_mul(cc.lL[j], ls));
        f3b = sc_add(f3b, sc_mul(cc.lp[j], sc_from_u64(u64_multiplicity(x, (uint32_t)j))));
        f3c = sc_add(f3c, sc_mul(nsw, a2));
        f4 = sc_add(f4, sc_mul(sc_sub(sc_sqr(cc.nR[j]), sc_sqr(a1)), cc.mp[j]));
        f5 = sc_add(f5, sc_mul(sc_sub(sc_sqr(cc.nL[j]), sc_sqr(a2)), cc.mp[j]));
        f6 = sc_add(f6, sc_mul(cc.lL[j], nr));
    }
    Sc f_[8];
    f_[0] = sc_neg(f0);
    f_[1] = f1;
    f_[2] = sc_neg(sc_dbl(f2));
    f_[3] = sc_add(sc_add(sc_dbl(f3a), f3b), sc_dbl(f3c));
    f_[4] = f4;
    f_[5] = f5;
    f_[6] = sc_dbl(sc_dbl(f6));        // 2 <c_lL, v_1>, v_1 = 2 r
    f_[7] = sc_zero();
    // rs (circuit.rs:457-467); ro/rl/rr zero entries dropped
    Sc ro[9], rl[9], rr[9];
#pragma unroll 1
    for (int k = 0; k < 9; k++) { ro[k] = sc_zero(); rl[k] = sc_zero(); rr[k] = sc_zero(); }
#pragma unroll 1
    for (int k = 0; k < 7; k++) ro[ro_pos(k)] = ws_ld_sc(w, i, PL::RND + 8 * (RD_RO + k));
#pragma unroll 1
    for (int k = 0; k < 6; k++) rl[rl_pos(k)] = ws_ld_sc(w, i, PL::RND + 8 * (RD_RL + k));
#pragma unroll 1
    for (int k = 0; k < 5; k++) rr[rr_pos(k)] = ws_ld_sc(w, i, PL::RND + 8 * (RD_RR + k));
    Sc rv0 = sc_dbl(sc_add(ws_ld_sc(w, i, PL::S), ws_ld_sc(w, i, PL::RND + 8 * RD_RBLIND)));   // 2 (s + r_blind)
    Sc rs[9];
    rs[0] = sc_add(f_[1], sc_mul(sc_mul(ro[1], delta), beta));
    rs[1] = sc_mul(f_[0], beta_inv);
    rs[2] = sc_sub(sc_mul(sc_add(sc_mul(ro[0], delta), f_[2]), beta_inv), rl[1]);
    rs[3] = sc_add(sc_mul(sc_sub(f_[3], rl[0]), beta_inv), sc_add(sc_mul(ro[2], delta), rr[1]));
    rs[4] = sc_add(sc_mul(sc_add(f_[4], rr[0]), beta_inv), sc_sub(sc_mul(ro[3], delta), rl[2]));
    rs[5] = sc_neg(sc_mul(rv0, beta_inv));
    rs[6] = sc_sub(sc_add(sc_add(sc_mul(f_[5], beta_inv), sc_mul(ro[5], delta)), rr[3]), rl[4]);
    rs[7] = sc_sub(sc_add(sc_add(sc_mul(f_[6], beta_inv), rr[4]), sc_mul(ro[6], delta)), rl[5]);
    rs[8] = sc_add(sc_sub(sc_add(sc_mul(f_[7], beta_inv), sc_mul(ro[7], delta)), rl[6]), rr[5]);
    // c_s scalars: rs (9) || ls (17) on h[0..26), ns (16) on g_vec
#pragma unroll 1
    for (int k = 0; k < 9; k++) { ws_st_sc(w, i, PL::RS9 + 8 * k, rs[k]); ws_st_sc(w, i, PL::FS + 8 * k, rs[k]); }
#pragma unroll 1
    for (int k = 0; k < 17; k++) ws_st_sc(w, i, PL::FS + 8 * (9 + k), ws_ld_sc(w, i, PL::RND + 8 * (RD_LS + k)));
#pragma unroll 1
    for (int k = 0; k < 16; k++) ws_st_sc(w, i, PL::FS + 8 * (26 + k), ws_ld_sc(w, i, PL::RND + 8 * (RD_NS + k)));
    ws_st_sc(w, i, PL::RHO, rho); ws_st_sc(w, i, PL::RHOINV, rho_inv); ws_st_sc(w, i, PL::MU, mu);
    ws_st_sc(w, i, PL::LAMBDA, lambda); ws_st_sc(w, i, PL::BETA, beta); ws_st_sc(w, i, PL::DELTA, delta);
    if (bad) pset_status(w, i, ST_PANIC_CHALLENGE_RANGE);
    if (zero_inv) pset_status(w, i, ST_PANIC_INVERT_ZERO);
}

// Scalars of X_j and R_j (wnla.rs:143-160) over the original generators, from the current l, n, c.
// rho_j = rho^(2^j), mu_j = mu^(2^j); y[k], k < j already drawn.
BPPP_HD void u64p_xr_scalars(const WS &w, size_t i, int j) {
    const int Lh = 32 >> j, Lg = 16 >> j, span = 1 << j;
    Sc rho_j = ws_ld_sc(w, i, PL::RHO), rho_inv_j = ws_ld_sc(w, i, PL::RHOINV), mu_j = ws_ld_sc(w, i, PL::MU);
    Sc rk[4];   // rho_k, k < j
    Sc yp[8], gp[8];
    yp[0] = sc_one(); gp[0] = sc_one();
#pragma unroll 1
    for (int k = 0; k < j; k++) {
        rk[k] = rho_j;
        Sc yk = ws_ld_sc(w, i, PL::Y + 8 * k);
        int sp = 1 << k;
#pragma unroll 1
        for (int t = 0; t < sp; t++) { yp[t + sp] = sc_mul(yp[t], yk); gp[t + sp] = sc_mul(gp[t], yk); gp[t] = sc_mul(gp[t], rho_j); }
        rho_j = sc_sqr(rho_j); rho_inv_j = sc_sqr(rho_inv_j); mu_j = sc_sqr(mu_j);
    }
    Sc mu2 = sc_sqr(mu_j);
    // vx = |n0, n1|_{mu2} * 2 rho^-1 + <c0, l1> + <c1, l0>;  vr = |n1|^2_{mu2} + <c1, l1>
    Sc wn = sc_zero(), wr = sc_zero(), pw = sc_one();
#pragma unroll 1
    for (int k = 0; k < Lg / 2; k++) {
        pw = sc_mul(pw, mu2);
        Sc n0 = ws_ld_sc(w, i, PL::NV + 8 * (2 * k)), n1 = ws_ld_sc(w, i, PL::NV + 8 * (2 * k + 1));
        Sc n1w = sc_mul(n1, pw);
        wn = sc_add(wn, sc_mul(n0, n1w));
        wr = sc_add(wr, sc_mul(n1, n1w));
    }
    Sc vx = sc_mul(wn, sc_dbl(rho_inv_j)), vr = wr;
#pragma unroll 1
    for (int k = 0; k < Lh / 2; k++) {
        Sc c0 = ws_ld_sc(w, i, PL::C + 8 * (2 * k)), c1 = ws_ld_sc(w, i, PL::C + 8 * (2 * k + 1));
        Sc l0 = ws_ld_sc(w, i, PL::LV + 8 * (2 * k)), l1 = ws_ld_sc(w, i, PL::LV + 8 * (2 * k + 1));
        vx = sc_add(vx, sc_add(sc_mul(c0, l1), sc_mul(c1, l0)));
        vr = sc_add(vr, sc_mul(c1, l1));
    }
    ws_st_sc(w, i, PL::XS, vx);
    ws_st_sc(w, i, PL::RS, vr);
    int ro = 1;   // running index into RS
#pragma unroll 1
    for (int mi = 0; mi < Lh; mi++) {          // h part: folded index mi, original index mi*span + t
        Sc lx = ws_ld_sc(w, i, PL::LV + 8 * (mi ^ 1)), lr = ws_ld_sc(w, i, PL::LV + 8 * mi);
#pragma unroll 1
        for (int t = 0; t < span; t++) {
            int idx = mi * span + t;
            ws_st_sc(w, i, PL::XS + 8 * (GEN_HVEC + idx), sc_mul(yp[t], lx));
            if (mi & 1) { ws_st_sc(w, i, PL::RS + 8 * ro, sc_mul(yp[t], lr)); ro++; }
        }
    }
#pragma unroll 1
    for (int mi = 0; mi < Lg; mi++) {          // g part
        Sc nx = sc_mul(ws_ld_sc(w, i, PL::NV + 8 * (mi ^ 1)), (mi & 1) ? rho_inv_j : rho_j);
        Sc nr = ws_ld_sc(w, i, PL::NV + 8 * mi);
#pragma unroll 1
        for (int t = 0; t < span; t++) {
            int idx = mi * span + t;
            ws_st_sc(w, i, PL::XS + 8 * (GEN_GVEC + idx), sc_mul(gp[t], nx));
            if (mi & 1) { ws_st_sc(w, i, PL::RS + 8 * ro, sc_mul(gp[t], nr)); ro++; }
        }
    }
}

// Phase 3 (circuit.rs:472-533): tau; l, n, c; scalars of the WNLA commitment C_0 and of X_0, R_0
BPPP_HD void u64p_phase3_one(const WS &w, size_t i, const uint8_t *ext = nullptr) {
    Tx m; tx_load(m, w, i, PL::MERLIN, ext);      // ext: tau
    bool id;
    PtA cs = u64p_affine(w, i, PP_CS, PO_CS, id);
    tx_point(m, BPPP_LBL("commitment_cs"), cs, id);
    Sc tau;
    if (!tx_challenge(m, BPPP_LBL("circuit_tau"), tau)) pset_status(w, i, ST_PANIC_CHALLENGE_RANGE);
    tx_store(m, w, i, PL::MERLIN);
    if (sc_is_zero(tau)) { pset_status(w, i, ST_PANIC_INVERT_ZERO); tau = sc_one(); }
    Sc tau_inv = sc_inv(tau), tau2 = sc_sqr(tau), tau3 = sc_mul(tau2, tau);
    Sc e = ws_ld_sc(w, i, PL::E), lambda = ws_ld_sc(w, i, PL::LAMBDA), mu = ws_ld_sc(w, i, PL::MU);
    Sc beta = ws_ld_sc(w, i, PL::BETA), delta = ws_ld_sc(w, i, PL::DELTA);
    Sc mu_inv = sc_sqr(ws_ld_sc(w, i, PL::RHOINV));
    CircuitCoefs cc;
    u64_circuit_coefs(cc, w, i, e, lambda, mu, mu_inv);
    uint64_t x = u64p_ld_x(w, i);
    // l (circuit.rs:479-483), first the 9 blinder slots, then the 17 witness slots, zero-padded to 32
    Sc ro[9], rl[9], rr[9];
#pragma unroll 1
    for (int k = 0; k < 9; k++) { ro[k] = sc_zero(); rl[k] = sc_zero(); rr[k] = sc_zero(); }
#pragma unroll 1
    for (int k = 0; k < 7; k++) ro[ro_pos(k)] = ws_ld_sc(w, i, PL::RND + 8 * (RD_RO + k));
#pragma unroll 1
    for (int k = 0; k < 6; k++) rl[rl_pos(k)] = ws_ld_sc(w, i, PL::RND + 8 * (RD_RL + k));
#pragma unroll 1
    for (int k = 0; k < 5; k++) rr[rr_pos(k)] = ws_ld_sc(w, i, PL::RND + 8 * (RD_RR + k));
    Sc rv0 = sc_dbl(sc_add(ws_ld_sc(w, i, PL::S), ws_ld_sc(w, i, PL::RND + 8 * RD_RBLIND)));
#pragma unroll 1
    for (int k = 0; k < 9; k++) {
        Sc v = sc_mul(ws_ld_sc(w, i, PL::RS9 + 8 * k), tau_inv);
        v = sc_sub(v, sc_mul(ro[k], delta));
        v = sc_add(v, sc_mul(rl[k], tau));
        v = sc_sub(v, sc_mul(rr[k], tau2));
        if (k == 0) v = sc_add(v, sc_mul(rv0, tau3));
        ws_st_sc(w, i, PL::LV + 8 * k, v);
    }
#pragma unroll 1
    for (int j = 0; j < 17; j++) {
        Sc v = sc_mul(ws_ld_sc(w, i, PL::RND + 8 * (RD_LS + j)), tau_inv);
        if (j < 16) {
            v = sc_add(v, sc_mul(sc_from_u64(u64_multiplicity(x, (uint32_t)j)), tau));                 // ll[j] tau
            v = sc_add(v, sc_mul(sc_dbl(ws_ld_sc(w, i, PL::INVE + 8 * (int)u64_digit(x, j))), tau3)); // v_1[j] tau^3, v_1 = 2 r
        }
        ws_st_sc(w, i, PL::LV + 8 * (9 + j), v);
    }
#pragma unroll 1
    for (int k = 26; k < 32; k++) ws_st_sc(w, i, PL::LV + 8 * k, sc_zero());
    // n = pn_tau + n_tau (circuit.rs:485-498), ps_tau (:489-491)
    Sc ps = sc_zero(), musum = sc_zero();
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        Sc pn = sc_add(sc_neg(sc_mul(cc.nL[j], tau2)), sc_mul(cc.nR[j], tau));
        ps = sc_add(ps, sc_mul(sc_sqr(pn), cc.mp[j]));
        musum = sc_add(musum, cc.mp[j]);
        Sc nt = sc_mul(ws_ld_sc(w, i, PL::RND + 8 * (RD_NS + j)), tau_inv);
        nt = sc_add(nt, sc_mul(sc_from_u64(u64_digit(x, j)), tau));
        nt = sc_sub(nt, sc_mul(ws_ld_sc(w, i, PL::INVE + 8 * (int)u64_digit(x, j)), tau2));
        ws_st_sc(w, i, PL::NV + 8 * j, sc_add(pn, nt));
    }
    ps = sc_sub(ps, sc_dbl(sc_mul(tau3, musum)));
    // c = cr_tau || cl_tau || 0 (circuit.rs:500-533)
    Sc bt = sc_mul(beta, tau);
    ws_st_sc(w, i, PL::C + 0, sc_one());
    ws_st_sc(w, i, PL::C + 8, sc_mul(tau_inv, beta));
#pragma unroll 1
    for (int k = 2; k < 9; k++) { ws_st_sc(w, i, PL::C + 8 * k, bt); bt = sc_mul(bt, tau); }
    Sc t22 = sc_dbl(tau2);
#pragma unroll 1
    for (int j = 0; j < 16; j++) ws_st_sc(w, i, PL::C + 8 * (9 + j), sc_sub(sc_neg(sc_mul(t22, cc.lL[j])), cc.lp[j]));
#pragma unroll 1
    for (int k = 25; k < 32; k++) ws_st_sc(w, i, PL::C + 8 * k, sc_zero());
    // C_0 scalars: v = ps_tau + tau^3 v_0 (v_0 = 2x) on g; l[0..26) on h; n on g_vec
    ws_st_sc(w, i, PL::FS, sc_add(ps, sc_mul(tau3, sc_dbl(sc_from_u64(x)))));
#pragma unroll 1
    for (int k = 0; k < 26; k++) ws_st_sc(w, i, PL::FS + 8 * (1 + k), ws_ld_sc(w, i, PL::LV + 8 * k));
#pragma unroll 1
    for (int k = 0; k < 16; k++) ws_st_sc(w, i, PL::FS + 8 * (27 + k), ws_ld_sc(w, i, PL::NV + 8 * k));
    u64p_xr_scalars(w, i, 0);
}

// WNLA round j (wnla.rs:162-175): transcript -> y_j; fold l, n, c; scalars of the next round's X, R
BPPP_HD void u64p_round_one(const WS &w, size_t i, int j, const uint8_t *ext = nullptr) {
    Tx m; tx_load(m, w, i, PL::MERLIN, ext);      // ext: y_j
    bool id;
    PtA a;
    a = u64p_affine(w, i, PP_COM, -1, id);           tx_point(m, BPPP_LBL("wnla_com"), a, id);
    a = u64p_affine(w, i, PP_X + j, PO_X + (3 - j), id); tx_point(m, BPPP_LBL("wnla_x"), a, id);
    a = u64p_affine(w, i, PP_R + j, PO_R + (3 - j), id); tx_point(m, BPPP_LBL("wnla_r"), a, id);
    const int Lh = 32 >> j, Lg = 16 >> j;
    tx_u64(m, BPPP_LBL("l.sz"), (uint64_t)Lh);    // l.len() (wnla.rs:165)
    tx_u64(m, BPPP_LBL("n.sz"), (uint64_t)Lg);
    Sc y;
    if (!tx_challenge(m, BPPP_LBL("wnla_challenge"), y)) pset_status(w, i, ST_PANIC_CHALLENGE_RANGE);
    tx_store(m, w, i, PL::MERLIN);
    ws_st_sc(w, i, PL::Y + 8 * j, y);
    Sc rho_inv_j = ws_ld_sc(w, i, PL::RHOINV);
#pragma unroll 1
    for (int k = 0; k < j; k++) rho_inv_j = sc_sqr(rho_inv_j);
#pragma unroll 1
    for (int k = 0; k < Lh / 2; k++) {
        Sc c0 = ws_ld_sc(w, i, PL::C + 8 * (2 * k)), c1 = ws_ld_sc(w, i, PL::C + 8 * (2 * k + 1));
        Sc l0 = ws_ld_sc(w, i, PL::LV + 8 * (2 * k)), l1 = ws_ld_sc(w, i, PL::LV + 8 * (2 * k + 1));
        ws_st_sc(w, i, PL::C + 8 * k, sc_add(c0, sc_mul(y, c1)));
        ws_st_sc(w, i, PL::LV + 8 * k, sc_add(l0, sc_mul(y, l1)));
    }
#pragma unroll 1
    for (int k = 0; k < Lg / 2; k++) {
        Sc n0 = ws_ld_sc(w, i, PL::NV + 8 * (2 * k)), n1 = ws_ld_sc(w, i, PL::NV + 8 * (2 * k + 1));
        ws_st_sc(w, i, PL::NV + 8 * k, sc_add(sc_mul(n0, rho_inv_j), sc_mul(y, n1)));
    }
    ws_st_sc(w, i, PL::VS, y);
    ws_st_sc(w, i, PL::VS + 8, sc_sub(sc_sqr(y), sc_one()));
    if (j < 3) u64p_xr_scalars(w, i, j + 1);
}

// next commitment C' = C + y X_j + (y^2 - 1) R_j  (== wnla'.commit(l', n'), wnla.rs:186)
// 1P..8P of X_j (t = 0) or R_j (t = 1), projective, into the prover's table region
BPPP_HD void u64p_table_build_one(const WS &w, size_t i, int j, int t) {
    bool id;
    const int slot = (t ? PP_R : PP_X) + j;
    PtA a = ws_affine(w, i, PL::PTS + PT_W * slot, PL::ZINV + FE_W * slot, id);
    PtTable8 tab;
    pt_table8_build(tab, pt_from_affine(a, id));
#pragma unroll 1
    for (int e = 0; e < 8; e++) ws_st_pt(w, i, PL::TAB + (t * 8 + e) * TAB_STRIDE_W, tab.m[e]);
}
// com_{j+1} = com_j + y X_j + (y^2 - 1) R_j from the normalised tables (same ladder as the verifier's rounds)
BPPP_HD void u64p_var2_one(const WS &w, size_t i, int j) {
    (void)j;
    const int tids[2] = {0, 1};
    Sc ks[2] = {ws_ld_sc(w, i, PL::VS), ws_ld_sc(w, i, PL::VS + 8)};
    Pt com = straus_tables<2>(w, ptab_region(), i, tids, ks, ws_ld_pt(w, i, PL::COM));
    ws_st_pt(w, i, PL::COM, com);
}

BPPP_HD Pt u64p_var2_partial(const WS &w, size_t i, int lane, int nlanes) {
    const int tids[2] = {0, 1};
    Sc ks[2] = {ws_ld_sc(w, i, PL::VS), ws_ld_sc(w, i, PL::VS + 8)};
    return ptj_to_pt(straus_tables_partial<2>(w, ptab_region(), i, tids, ks, lane, nlanes));
}

// 525-byte record: c_l c_r c_o c_s | r[0..4) | x[0..4) | l[0..2) | n[0] | r
BPPP_HD void u64p_output_one(const WS &w, size_t i, uint8_t *out) {
#pragma unroll 1
    for (int po = 0; po < PO_COUNT; po++) {
        uint8_t *dst = out + (po < 12 ? 33 * po : 33 * 12 + 96);
        uint32_t xw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) xw[k] = ws_ld(w, i, PL::OUT + 9 * po + k);
        dst[0] = (uint8_t)ws_ld(w, i, PL::OUT + 9 * po + 8);
        words_to_be32(dst + 1, xw);
    }
    sc_to_be32(out + 396, ws_ld_sc(w, i, PL::LV));
    sc_to_be32(out + 428, ws_ld_sc(w, i, PL::LV + 8));
    sc_to_be32(out + 460, ws_ld_sc(w, i, PL::NV));
}

}  // namespace bppp
