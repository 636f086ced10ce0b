// Host emulation of the u64 batch verify/prove DEVICE logic (bp_pp_b200/csrc/u64_*.cuh compiled as
// plain C++): every per-proof phase function the CUDA kernels call, driven by plain loops in the
// same order the engine launches the kernels.  Test infrastructure only; never part of libbppp.so.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../bp_pp_b200/csrc/u64_verify.cuh"
#ifdef BPPP_EMU_PROVE
#include "../../bp_pp_b200/csrc/u64_prove.cuh"
#endif
using namespace bppp;

struct EmuCtx {
    std::vector<uint4> tab;
    FixedTable T;
    PtA gens[NUM_GENS]; bool gen_id[NUM_GENS];
};

static int g_tab_affine = 1;      // ladder tables: 1 = affine levels (tables_affine_level), 0 = projective build + normalise

static int g_var_seg = 0;         // > 0: the verifier's ladders run as that many segments through the scratch rows (u64v_var*_seg)

extern "C" {
void emu_set_tab_affine(int on) { g_tab_affine = on; }
void emu_set_var_segments(int nseg) { g_var_seg = nseg; }
void *emu_ctx_create(const uint8_t *gens64, int W) {
    EmuCtx *c = new EmuCtx();
    bool sgn = W < 0;                          // negative: signed windows of |W| bits (ws.cuh:FixedTable)
    if (sgn) W = -W;
    fixed_table_shape(c->T, W, sgn);
    int nwin = c->T.nwin;
    size_t E = c->T.E;
    c->tab.assign((size_t)NUM_GENS * nwin * E * 4, uint4{0, 0, 0, 0});
    for (int g = 0; g < NUM_GENS; g++) {
        PtA a; int s = pta_from_xy64(a, gens64 + 64 * g);
        if (s < 0) { delete c; return nullptr; }
        c->gens[g] = a; c->gen_id[g] = s == 1;
        Pt base = pt_from_affine(a, s == 1);
        for (int w = 0; w < nwin; w++) {
            Pt cur = pt_identity();
            for (size_t d = 1; d <= E; d++) {
                cur = pt_add(cur, base);
                bool id = pt_is_identity(cur);
                PtA q = pt_to_affine_with_zinv(cur, fe_inv(cur.z));
                uint32_t x[8], y[8];
                fe_to_words(x, q.x); fe_to_words(y, q.y);
                if (id) { memset(x, 0, 32); memset(y, 0, 32); }
                size_t idx = ((size_t)(g * nwin + w) * E + (d - 1)) * 4;
                c->tab[idx] = uint4{x[0], x[1], x[2], x[3]}; c->tab[idx + 1] = uint4{x[4], x[5], x[6], x[7]};
                c->tab[idx + 2] = uint4{y[0], y[1], y[2], y[3]}; c->tab[idx + 3] = uint4{y[4], y[5], y[6], y[7]};
            }
            for (int k = 0; k < W; k++) base = pt_double(base);
        }
    }
    c->T.tab = c->tab.data(); c->T.ngens = NUM_GENS;
    return c;
}
void emu_ctx_destroy(void *p) { delete (EmuCtx *)p; }

static void emu_msm_fixed(EmuCtx *c, const WS &w, size_t n, int sc_off, const int *term_gen, int nterms, int out_off, int nlanes) {
    for (size_t i = 0; i < n; i++) {
        Pt acc = pt_identity();
        for (int lane = 0; lane < nlanes; lane++)
            acc = pt_add(acc, c->T.sgn ? msm_fixed_lane_signed(c->T, w, i, sc_off, term_gen, nterms, lane, nlanes) : msm_fixed_lane(c->T, w, i, sc_off, term_gen, nterms, lane, nlanes));
        ws_st_pt(w, i, out_off, acc);
    }
}
static void emu_batch_inv(const WS &w, size_t n, int in_off, int out_off) {
    size_t T = (n + 3) / 4;   // 4 items per emulated thread
    for (size_t t = 0; t < T; t++) batch_inv_strided(w, in_off, out_off, t, T, n);
}

int emu_u64_verify_batch(void *ctx, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt, const uint8_t *label, uint32_t label_len, int32_t *status) {
    EmuCtx *c = (EmuCtx *)ctx;
    std::vector<uint32_t> buf((size_t)VL::WORDS * n, 0);
    WS w{buf.data(), n};
    Merlin init; merlin_init(init, label, label_len);
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    for (size_t i = 0; i < n; i++) u64v_load_one(w, i, commits + csz * i, proofs + psz * i, fmt);
    emu_batch_inv(w, n, VL::VP + 2 * FE_W, VL::ZINV);
    for (size_t i = 0; i < n; i++) u64v_phase1_one(w, i, init);
    if (g_tab_affine) {
        for (int level = 1; level <= 3; level++) {
            size_t items = n * VL::TAB_POINTS * aff_level_nops(level), T = (items + 4) / 5;      // 5 items per emulated thread
            for (size_t t = 0; t < T; t++) u64v_tables_affine_level(w, level, t, T);
        }
    } else {
        for (size_t i = 0; i < n; i++) for (int t = 0; t < VL::TAB_POINTS; t++) u64v_table_build_one(w, i, t);
        size_t T = (n * VL::TAB_ENTRIES + 6) / 7;
        for (size_t t = 0; t < T; t++) u64v_tables_normalize_strided(w, t, T);
    }
    int tg17[17]; for (int t = 0; t < 17; t++) tg17[t] = t;
    emu_msm_fixed(c, w, n, VL::FS, tg17, 17, VL::ACC, 8);
    if (g_var_seg > 0) { for (int sg = 0; sg < g_var_seg; sg++) for (size_t i = 0; i < n; i++) u64v_var5_seg(w, i, sg, g_var_seg); }
    else for (size_t i = 0; i < n; i++) u64v_var5_one(w, i);
    for (int j = 0; j < 4; j++) {
        emu_batch_inv(w, n, VL::COM + 2 * FE_W, VL::ZINV);
        for (size_t i = 0; i < n; i++) u64v_round_one(w, i, j);
        if (g_var_seg > 0) { for (int sg = 0; sg < g_var_seg; sg++) for (size_t i = 0; i < n; i++) u64v_var2_seg(w, i, j, sg, g_var_seg); }
        else for (size_t i = 0; i < n; i++) u64v_var2_one(w, i, j);
    }
    for (size_t i = 0; i < n; i++) u64v_final_scalars_one(w, i);
    int tg49[49]; for (int t = 0; t < 49; t++) tg49[t] = t;
    emu_msm_fixed(c, w, n, VL::FS, tg49, 49, VL::ACC, 8);
    for (size_t i = 0; i < n; i++) { u64v_verdict_one(w, i); status[i] = (int32_t)ws_ld(w, i, VL::STATUS); }
    return 0;
}
// the finished ladder tables (array-of-structures region, 104 entries x 24 words per proof) of a batch after phase 1,
// built the way `affine` says: the two constructions must agree word for word
int emu_u64_verify_tables(size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt, int affine, uint32_t *out) {
    std::vector<uint32_t> buf((size_t)VL::WORDS * n, 0);
    WS w{buf.data(), n};
    Merlin init; merlin_init(init, (const uint8_t *)"t", 1);
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    for (size_t i = 0; i < n; i++) u64v_load_one(w, i, commits + csz * i, proofs + psz * i, fmt);
    emu_batch_inv(w, n, VL::VP + 2 * FE_W, VL::ZINV);
    for (size_t i = 0; i < n; i++) u64v_phase1_one(w, i, init);
    if (affine) {
        for (int level = 1; level <= 3; level++) {
            size_t items = n * VL::TAB_POINTS * aff_level_nops(level), T = (items + 2) / 3;
            for (size_t t = 0; t < T; t++) u64v_tables_affine_level(w, level, t, T);
        }
    } else {
        for (size_t i = 0; i < n; i++) for (int t = 0; t < VL::TAB_POINTS; t++) u64v_table_build_one(w, i, t);
        size_t T = (n * VL::TAB_ENTRIES + 6) / 7;
        for (size_t t = 0; t < T; t++) u64v_tables_normalize_strided(w, t, T);
    }
    memcpy(out, tab_entry(w, vtab_region(), 0, 0), sizeof(uint32_t) * n * VL::TAB_ENTRIES * VL::TABA_STRIDE);
    return 0;
}
// commit_value (engine_core.cu:bppp_u64_commit_batch): x g + s h_0 through the fixed-base lane code, 3 lanes
int emu_u64_commit_batch(void *ctx, size_t n, const uint64_t *xs, const uint8_t *blinds, uint8_t *out33) {
    EmuCtx *c = (EmuCtx *)ctx;
    std::vector<uint32_t> buf((size_t)VL::WORDS * n, 0);
    WS w{buf.data(), n};
    for (size_t i = 0; i < n; i++) {
        Sc s;
        if (!sc_from_be32(s, blinds + 32 * i)) return -1;
        ws_st_sc(w, i, VL::FS, sc_from_u64(xs[i]));
        ws_st_sc(w, i, VL::FS + 8, s);
    }
    int tg[2] = {GEN_G, GEN_HVEC};
    emu_msm_fixed(c, w, n, VL::FS, tg, 2, VL::ACC, 3);
    emu_batch_inv(w, n, VL::ACC + 2 * FE_W, VL::ZINV);
    for (size_t i = 0; i < n; i++) {
        bool id;
        PtA a = ws_affine(w, i, VL::ACC, VL::ZINV, id);
        pta_compress(out33 + 33 * i, a, id);
    }
    return 0;
}
}

#ifdef BPPP_EMU_PROVE
extern "C" int emu_u64_prove_batch(void *ctx, size_t n, const uint64_t *xs, const uint8_t *blinds, const uint8_t *rng, const uint8_t *label, uint32_t label_len,
                                   uint8_t *proofs, int32_t *status) {
    EmuCtx *c = (EmuCtx *)ctx;
    std::vector<uint32_t> buf((size_t)PL::WORDS * n, 0);
    WS w{buf.data(), n};
    Merlin init; merlin_init(init, label, label_len);
    int tm[NUM_GENS];
    for (size_t i = 0; i < n; i++) u64p_load_one(w, i, xs[i], blinds + 32 * i);
    u64p_termmap_commit(tm);
    emu_msm_fixed(c, w, n, PL::FS, tm, 2, PL::PTS + PT_W * PP_V, 8);
    emu_batch_inv(w, n, PL::PTS + PT_W * PP_V + 2 * FE_W, PL::ZINV + FE_W * PP_V);
    for (size_t i = 0; i < n; i++) u64p_phase1_one(w, i, init, rng + (size_t)U64_RNG_BYTES * i);
    for (int k = 0; k < 4; k++) {
        int nt = u64p_termmap_stage1(tm, k);
        emu_msm_fixed(c, w, n, PL::FS + 8 * u64p_stage1_scalar_base(k), tm, nt, PL::PTS + PT_W * u64p_stage1_point(k), 8);
    }
    for (size_t i = 0; i < n; i++) u64p_vprime_one(w, i);
    for (int k = 0; k < 5; k++) { int p = u64p_stage1_norm_point(k); emu_batch_inv(w, n, PL::PTS + PT_W * p + 2 * FE_W, PL::ZINV + FE_W * p); }
    for (size_t i = 0; i < n; i++) u64p_phase2_one(w, i, rng + (size_t)U64_RNG_BYTES * i);
    u64p_termmap_cs(tm);
    emu_msm_fixed(c, w, n, PL::FS, tm, 42, PL::PTS + PT_W * PP_CS, 8);
    emu_batch_inv(w, n, PL::PTS + PT_W * PP_CS + 2 * FE_W, PL::ZINV + FE_W * PP_CS);
    for (size_t i = 0; i < n; i++) u64p_phase3_one(w, i);
    u64p_termmap_c0(tm);
    emu_msm_fixed(c, w, n, PL::FS, tm, 43, PL::COM, 8);
    for (int j = 0; j < 4; j++) {
        int all[NUM_GENS]; for (int t = 0; t < NUM_GENS; t++) all[t] = t;
        emu_msm_fixed(c, w, n, PL::XS, all, NUM_GENS, PL::PTS + PT_W * (PP_X + j), 8);
        u64p_termmap_r(tm, j);
        emu_msm_fixed(c, w, n, PL::RS, tm, 25, PL::PTS + PT_W * (PP_R + j), 8);
        emu_batch_inv(w, n, PL::COM + 2 * FE_W, PL::ZINV + FE_W * PP_COM);
        emu_batch_inv(w, n, PL::PTS + PT_W * (PP_X + j) + 2 * FE_W, PL::ZINV + FE_W * (PP_X + j));
        emu_batch_inv(w, n, PL::PTS + PT_W * (PP_R + j) + 2 * FE_W, PL::ZINV + FE_W * (PP_R + j));
        for (size_t i = 0; i < n; i++) u64p_round_one(w, i, j);
        if (j < 3) {
            for (int t = 0; t < 2; t++) for (size_t i = 0; i < n; i++) u64p_table_build_one(w, i, j, t);
            { size_t T = (n * PL::TAB_ENTRIES + 4) / 5; for (size_t t = 0; t < T; t++) tables_normalize_strided(w, ptab_region(), t, T); }
            for (size_t i = 0; i < n; i++) u64p_var2_one(w, i, j);
        }
    }
    for (size_t i = 0; i < n; i++) { u64p_output_one(w, i, proofs + (size_t)U64_PROOF_BYTES_COMPRESSED * i); status[i] = (int32_t)ws_ld(w, i, PL::STATUS); }
    return 0;
}
#endif
