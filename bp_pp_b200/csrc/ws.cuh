// Batch workspace in HBM: word-major structure-of-arrays.  Word w of proof i lives at
// p[w * n + i], so a warp of 32 consecutive proofs reads/writes 128 contiguous bytes per word
// (coalesced) no matter which field of the per-proof record it touches.
#pragma once
#include "ec.cuh"
#include "merlin.cuh"

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace bppp {

struct WS {
    uint32_t *p;
    size_t n;   // proofs in this batch (row stride)
};

BPPP_HD uint32_t ws_ld(const WS &w, size_t i, int word) { return w.p[(size_t)word * w.n + i]; }
BPPP_HD void ws_st(const WS &w, size_t i, int word, uint32_t v) { w.p[(size_t)word * w.n + i] = v; }

BPPP_HD Sc ws_ld_sc(const WS &w, size_t i, int off) { Sc r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = ws_ld(w, i, off + k); return r; }
BPPP_HD void ws_st_sc(const WS &w, size_t i, int off, const Sc &a) {
#pragma unroll
    for (int k = 0; k < 8; k++) ws_st(w, i, off + k, a.v[k]); }
BPPP_HD Fe ws_ld_fe(const WS &w, size_t i, int off) { Fe r;
#pragma unroll
    for (int k = 0; k < FE_W; k++) r.v[k] = ws_ld(w, i, off + k);
    return r; }
BPPP_HD void ws_st_fe(const WS &w, size_t i, int off, const Fe &a) {
#pragma unroll
    for (int k = 0; k < FE_W; k++) ws_st(w, i, off + k, a.v[k]); }
BPPP_HD Pt ws_ld_pt(const WS &w, size_t i, int off) { Pt r; r.x = ws_ld_fe(w, i, off); r.y = ws_ld_fe(w, i, off + FE_W); r.z = ws_ld_fe(w, i, off + 2 * FE_W); return r; }
BPPP_HD void ws_st_pt(const WS &w, size_t i, int off, const Pt &a) { ws_st_fe(w, i, off, a.x); ws_st_fe(w, i, off + FE_W, a.y); ws_st_fe(w, i, off + 2 * FE_W, a.z); }
// affine point stored as 16 canonical words (x[8], y[8]); (0,0) = identity sentinel
BPPP_HD PtA ws_ld_pta(const WS &w, size_t i, int off) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = ws_ld(w, i, off + k); y[k] = ws_ld(w, i, off + 8 + k); }
    PtA r; r.x = fe_from_words(x); r.y = fe_from_words(y); return r;
}
BPPP_HD void ws_st_pta(const WS &w, size_t i, int off, const PtA &a_canonical) {
    uint32_t x[8], y[8];
    fe_to_words(x, a_canonical.x); fe_to_words(y, a_canonical.y);
#pragma unroll
    for (int k = 0; k < 8; k++) { ws_st(w, i, off + k, x[k]); ws_st(w, i, off + 8 + k, y[k]); }
}
BPPP_HD void ws_ld_merlin(Merlin &m, const WS &w, size_t i, int off) {
#pragma unroll
    for (int k = 0; k < 25; k++) m.st[k] = (uint64_t)ws_ld(w, i, off + 2 * k) | ((uint64_t)ws_ld(w, i, off + 2 * k + 1) << 32);
    uint32_t pp = ws_ld(w, i, off + 50);
    m.pos = pp & 0xFFu; m.pos_begin = (pp >> 8) & 0xFFu; m.cur_flags = (pp >> 16) & 0xFFu; m._pad = 0;
}
BPPP_HD void ws_st_merlin(const WS &w, size_t i, int off, const Merlin &m) {
#pragma unroll
    for (int k = 0; k < 25; k++) { ws_st(w, i, off + 2 * k, (uint32_t)m.st[k]); ws_st(w, i, off + 2 * k + 1, (uint32_t)(m.st[k] >> 32)); }
    ws_st(w, i, off + 50, m.pos | (m.pos_begin << 8) | (m.cur_flags << 16));
}

// absorb an affine point (canonical coords) as the reference's app_point does (transcript.rs:6-8)
BPPP_HD void merlin_append_point(Merlin &m, const char *label, uint32_t label_len, const PtA &a_canonical, bool is_identity) {
    uint8_t b[33];
    pta_compress(b, a_canonical, is_identity);
    merlin_append(m, label, label_len, b, 33);
}
// get_challenge (transcript.rs:10-14): 32 bytes big-endian -> scalar; false when >= n (reference panics)
BPPP_HD bool merlin_challenge_scalar(Merlin &m, const char *label, uint32_t label_len, Sc &out) {
    uint8_t b[32];
    merlin_challenge(m, label, label_len, b, 32);
    return sc_from_be32(out, b);
}

// Transcript endpoint of one proof within one phase.
//  * internal (ext == nullptr): the device-side Merlin of the batch entry points, which own a fresh Transcript::new(label);
//  * external: the CALLER owns the merlin::Transcript (the reference's single-instance signatures take `t: &mut Transcript`
//    with arbitrary prior state, and merlin keeps its STROBE state private).  The host was handed the bytes of the points
//    this phase appends by the previous step and did the appends itself, so appends are no-ops here and the phase's
//    challenges are read, in order, from `ext` (32 bytes big-endian each).
struct Tx {
    Merlin m;
    const uint8_t *ext;
    int k;
};
BPPP_HD void tx_init(Tx &t, const Merlin &init, const uint8_t *ext) { t.ext = ext; t.k = 0; if (!ext) t.m = init; }
BPPP_HD void tx_load(Tx &t, const WS &w, size_t i, int off, const uint8_t *ext) { t.ext = ext; t.k = 0; if (!ext) ws_ld_merlin(t.m, w, i, off); }
BPPP_HD void tx_store(const Tx &t, const WS &w, size_t i, int off) { if (!t.ext) ws_st_merlin(w, i, off, t.m); }
BPPP_HD void tx_point(Tx &t, const char *label, uint32_t label_len, const PtA &a_canonical, bool is_identity) {
    if (!t.ext) merlin_append_point(t.m, label, label_len, a_canonical, is_identity);
}
BPPP_HD void tx_u64(Tx &t, const char *label, uint32_t label_len, uint64_t v) { if (!t.ext) merlin_append_u64(t.m, label, label_len, v); }
BPPP_HD bool tx_challenge(Tx &t, const char *label, uint32_t label_len, Sc &out) {
    if (t.ext) return sc_from_be32(out, t.ext + 32 * t.k++);
    return merlin_challenge_scalar(t.m, label, label_len, out);
}

// status codes shared by the device code and the C ABI (include/bppp.h)
enum : int32_t {
    ST_FALSE = 0, ST_TRUE = 1,
    ST_PANIC_INVERT_ZERO = -1,     // reference: Scalar::invert().unwrap() on zero
    ST_PANIC_CHALLENGE_RANGE = -2, // reference: Scalar::from_repr(..).unwrap() on >= n (transcript.rs:13)
    ST_BAD_POINT = -3,             // encoding is not a curve point (reference: deserialisation error)
    ST_BAD_SCALAR = -4,            // scalar >= n (a k256::Scalar cannot hold it)
    ST_BAD_ARG = -5,
};

// ---- fixed-base tables ----
// tab[((g * nwin + w) * E + (d - 1))] = d * 2^(W w) * G_g, affine, 16 words (x[8], y[8]); zero = identity.
// Unsigned windows: E = 2^W - 1 entries, nwin = ceil(256 / W).  Signed windows (sgn): digits in (-2^(W-1), 2^(W-1)], so
// E = 2^(W-1) entries serve W bits -- one more bit per window for the same memory -- and nwin = ceil(257 / W) leaves
// room for the final carry; a negative digit adds the negated entry.
struct FixedTable {
    const uint4 *tab;
    int W;       // window bits
    int nwin;
    int ngens;
    uint32_t E;  // entries per (generator, window)
    int sgn;
};
BPPP_HD void fixed_table_shape(FixedTable &T, int W, bool sgn) {
    T.W = W; T.sgn = sgn ? 1 : 0;
    T.nwin = sgn ? (257 + W - 1) / W : (256 + W - 1) / W;
    T.E = sgn ? (1u << (W - 1)) : ((1u << W) - 1u);
}

BPPP_HD uint32_t scalar_window(const WS &w, size_t i, int sc_off, int win, int W) {
    int bit = win * W;
    int word = bit >> 5, sh = bit & 31;
    uint32_t lo = ws_ld(w, i, sc_off + word);
    uint64_t v = lo;
    if (sh + W > 32 && word + 1 < 8) v |= (uint64_t)ws_ld(w, i, sc_off + word + 1) << 32;
    return (uint32_t)(v >> sh) & ((1u << W) - 1u);
}

struct TableEntryRaw { uint4 a, b, c, e; };   // x words 0..7, y words 0..7
BPPP_HD TableEntryRaw table_fetch(const FixedTable &T, int g, int win, uint32_t d) {
    size_t idx = ((size_t)(g * T.nwin + win) * T.E + (d - 1)) * 4;
    TableEntryRaw r;
#if defined(__CUDA_ARCH__)
    r.a = __ldg(T.tab + idx); r.b = __ldg(T.tab + idx + 1); r.c = __ldg(T.tab + idx + 2); r.e = __ldg(T.tab + idx + 3);
#else
    r.a = T.tab[idx]; r.b = T.tab[idx + 1]; r.c = T.tab[idx + 2]; r.e = T.tab[idx + 3];
#endif
    return r;
}
BPPP_HD bool table_decode(PtA &q, const TableEntryRaw &r) {
    uint32_t x[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
    uint32_t y[8] = {r.c.x, r.c.y, r.c.z, r.c.w, r.e.x, r.e.y, r.e.z, r.e.w};
    q.x = fe_from_words(x); q.y = fe_from_words(y);
    uint32_t any = r.a.x | r.a.y | r.a.z | r.a.w | r.b.x | r.b.y | r.b.z | r.b.w | r.c.x | r.c.y | r.c.z | r.c.w | r.e.x | r.e.y | r.e.z | r.e.w;
    return any != 0;
}

// the two workspace words a W-bit window can straddle, and the window cut out of them
struct WindowWords { uint32_t lo, hi; };
BPPP_HD WindowWords scalar_window_words(const WS &w, size_t i, int sc_off, int win, int W) {
    int word = (win * W) >> 5;
    WindowWords r;
    r.lo = ws_ld(w, i, sc_off + word);
    r.hi = word + 1 < 8 ? ws_ld(w, i, sc_off + word + 1) : 0u;
    return r;
}
BPPP_HD uint32_t window_of_words(const WindowWords &ww, int win, int W) {
    int sh = (win * W) & 31;
    uint64_t v = (uint64_t)ww.lo | ((uint64_t)ww.hi << 32);
    return (uint32_t)(v >> sh) & ((1u << W) - 1u);
}

// One lane's share of sum_t scalar_t * G_{gen(t)}: items (t, win) are dealt round-robin to `nlanes` lanes.
// scalars: T consecutive Sc in the workspace starting at word sc_off; term_gen[t] = generator index.
// Software pipeline, two stages deep: while the mixed addition of item k runs, the 64-byte table entry of item k+1 is
// in flight from HBM and so are the scalar words of item k+2 (the window -> address -> entry chain is two dependent loads;
// ncu showed the first one exposed as long-scoreboard stalls).
// (term, window) of an item index, advanced by `step` items without dividing (step <= nwin)
struct ItemPos { int t, win; };
BPPP_HD ItemPos item_pos(int it, int nwin) { ItemPos p; p.t = it / nwin; p.win = it - p.t * nwin; return p; }
BPPP_HD void item_advance(ItemPos &p, int step, int nwin) { p.win += step; if (p.win >= nwin) { p.win -= nwin; p.t++; } }

BPPP_HD Pt msm_fixed_lane(const FixedTable &T, const WS &w, size_t i, int sc_off, const int *term_gen, int nterms, int lane, int nlanes) {
    PtX acc = ptx_identity();             // XYZZ accumulator: 8 M + 2 S per table point (ec.cuh)
    const int items = nterms * T.nwin;
    const bool small_step = nlanes <= T.nwin;
    TableEntryRaw cur, nxt;
    uint32_t dcur = 0, dnxt = 0;
    WindowWords ww_next; ww_next.lo = 0; ww_next.hi = 0;
    int it = lane;
    ItemPos p1 = item_pos(it, T.nwin);                     // position of the item whose entry is fetched next
    if (it < items) {
        dcur = window_of_words(scalar_window_words(w, i, sc_off + 8 * p1.t, p1.win, T.W), p1.win, T.W);
        if (dcur) cur = table_fetch(T, term_gen[p1.t], p1.win, dcur);
    }
    if (small_step) item_advance(p1, nlanes, T.nwin); else p1 = item_pos(it + nlanes, T.nwin);
    ItemPos p2 = p1;                                       // position of the item whose scalar words are loaded next
    if (it + nlanes < items) ww_next = scalar_window_words(w, i, sc_off + 8 * p1.t, p1.win, T.W);
    if (small_step) item_advance(p2, nlanes, T.nwin); else p2 = item_pos(it + 2 * nlanes, T.nwin);
#pragma unroll 1
    for (; it < items; it += nlanes) {
        const int itn = it + nlanes, itnn = it + 2 * nlanes;
        dnxt = 0;
        if (itn < items) {
            dnxt = window_of_words(ww_next, p1.win, T.W);
            if (dnxt) nxt = table_fetch(T, term_gen[p1.t], p1.win, dnxt);
        }
        if (itnn < items) ww_next = scalar_window_words(w, i, sc_off + 8 * p2.t, p2.win, T.W);
        p1 = p2;
        if (small_step) item_advance(p2, nlanes, T.nwin); else p2 = item_pos(itnn + nlanes, T.nwin);
        if (dcur != 0) {
            PtA q;
            if (table_decode(q, cur)) acc = ptx_add_mixed_hot(acc, q);
        }
        cur = nxt; dcur = dnxt;
    }
    return ptx_to_pt(acc);
}

// ---- signed windows ----
// Three workspace words starting at the word that holds the first bit of window win - 1 (of window 0 for win = 0): they
// cover window win - 1 (needed for the carry into win) and window win itself (2 W <= 46 bits from a bit offset < 32).
struct WindowWords3 { uint32_t a, b, c; };
BPPP_HD WindowWords3 scalar_window_words3(const WS &w, size_t i, int sc_off, int win, int W) {
    int word = (win ? (win - 1) * W : 0) >> 5;
    WindowWords3 r;
    r.a = word < 8 ? ws_ld(w, i, sc_off + word) : 0u;
    r.b = word + 1 < 8 ? ws_ld(w, i, sc_off + word + 1) : 0u;
    r.c = word + 2 < 8 ? ws_ld(w, i, sc_off + word + 2) : 0u;
    return r;
}
BPPP_HD uint32_t words3_bits(const WindowWords3 &ww, int rel, int W) {     // W bits from bit offset rel < 64
    uint64_t v = rel < 32 ? (((uint64_t)ww.a | ((uint64_t)ww.b << 32)) >> rel) : (((uint64_t)ww.b | ((uint64_t)ww.c << 32)) >> (rel - 32));
    return (uint32_t)v & ((1u << W) - 1u);
}
struct SignedDigit { uint32_t mag; bool neg; };
// digit of window win: value + carry-in, recentred.  The carry out of window j is 1 iff value_j + carry_j > 2^(W-1), which
// is decided by value_j alone unless value_j == 2^(W-1) exactly; only then does it look further down (probability 2^-W).
BPPP_HD SignedDigit signed_window_digit(const FixedTable &T, const WS &w, size_t i, int sc_off, int win, const WindowWords3 &ww) {
    const int W = T.W;
    const uint32_t H = 1u << (W - 1);
    const int base = ((win ? (win - 1) * W : 0) >> 5) * 32;
    uint32_t v = words3_bits(ww, win * W - base, W);
    uint32_t carry = 0;
    if (win) {
        uint32_t vp = words3_bits(ww, (win - 1) * W - base, W);
        carry = vp > H ? 1u : 0u;
        if (vp == H) {
#pragma unroll 1
            for (int j = win - 2; j >= 0; j--) {
                uint32_t vj = scalar_window(w, i, sc_off, j, W);
                if (vj != H) { carry = vj > H ? 1u : 0u; break; }
            }
        }
    }
    uint32_t d = v + carry;
    SignedDigit r;
    if (win < T.nwin - 1 && d > H) { r.mag = (1u << W) - d; r.neg = true; }
    else { r.mag = d; r.neg = false; }
    return r;
}
// msm_fixed_lane for signed tables: same two-stage software pipeline
BPPP_HD Pt msm_fixed_lane_signed(const FixedTable &T, const WS &w, size_t i, int sc_off, const int *term_gen, int nterms, int lane, int nlanes) {
    PtX acc = ptx_identity();
    const int items = nterms * T.nwin;
    const bool small_step = nlanes <= T.nwin;
    TableEntryRaw cur, nxt;
    SignedDigit dcur, dnxt;
    dcur.mag = 0; dcur.neg = false;
    WindowWords3 ww_next; ww_next.a = ww_next.b = ww_next.c = 0;
    int it = lane;
    ItemPos p1 = item_pos(it, T.nwin);
    if (it < items) {
        dcur = signed_window_digit(T, w, i, sc_off + 8 * p1.t, p1.win, scalar_window_words3(w, i, sc_off + 8 * p1.t, p1.win, T.W));
        if (dcur.mag) cur = table_fetch(T, term_gen[p1.t], p1.win, dcur.mag);
    }
    if (small_step) item_advance(p1, nlanes, T.nwin); else p1 = item_pos(it + nlanes, T.nwin);
    ItemPos p2 = p1;
    if (it + nlanes < items) ww_next = scalar_window_words3(w, i, sc_off + 8 * p1.t, p1.win, T.W);
    if (small_step) item_advance(p2, nlanes, T.nwin); else p2 = item_pos(it + 2 * nlanes, T.nwin);
#pragma unroll 1
    for (; it < items; it += nlanes) {
        const int itn = it + nlanes, itnn = it + 2 * nlanes;
        dnxt.mag = 0; dnxt.neg = false;
        if (itn < items) {
            dnxt = signed_window_digit(T, w, i, sc_off + 8 * p1.t, p1.win, ww_next);
            if (dnxt.mag) nxt = table_fetch(T, term_gen[p1.t], p1.win, dnxt.mag);
        }
        if (itnn < items) ww_next = scalar_window_words3(w, i, sc_off + 8 * p2.t, p2.win, T.W);
        p1 = p2;
        if (small_step) item_advance(p2, nlanes, T.nwin); else p2 = item_pos(itnn + nlanes, T.nwin);
        if (dcur.mag != 0) {
            PtA q;
            if (table_decode(q, cur)) {
                if (dcur.neg) q.y = fe_negate(q.y, 1);
                acc = ptx_add_mixed_hot(acc, q);
            }
        }
        cur = nxt; dcur = dnxt;
    }
    return ptx_to_pt(acc);
}

}  // namespace bppp

namespace bppp {
// Montgomery batch inversion of the Fe field at word offset in_off into out_off for the items
// t, t + T, t + 2T, ... < n (one thread's strided share, coalesced across threads): one fe_inv per
// share instead of one per item.  Zero inputs (identity points) produce zero outputs.
BPPP_HD void batch_inv_strided(const WS &w, int in_off, int out_off, size_t t, size_t T, size_t n) {
    Fe run = fe_one();
#pragma unroll 1
    for (size_t idx = t; idx < n; idx += T) {
        Fe z = ws_ld_fe(w, idx, in_off);
        ws_st_fe(w, idx, out_off, run);                 // prefix product before this item
        if (!fe_is_zero(z)) run = fe_mul(run, z);
    }
    Fe rinv = fe_inv(run);
    size_t cnt = n > t ? (n - t + T - 1) / T : 0;
#pragma unroll 1
    for (size_t k = cnt; k-- > 0;) {
        size_t idx = t + k * T;
        Fe z = ws_ld_fe(w, idx, in_off);
        Fe pre = ws_ld_fe(w, idx, out_off);
        if (fe_is_zero(z)) { Fe zero = fe_zero(); ws_st_fe(w, idx, out_off, zero); }
        else { ws_st_fe(w, idx, out_off, fe_mul(rinv, pre)); rinv = fe_mul(rinv, z); }
    }
}
// The same over a short list of (input, output) field offsets per proof: flat item index = f * n + proof, so consecutive
// threads still touch consecutive proofs.  One launch and one inversion chain serve several independent normalisations
// (the prover needs five after its first stage and three per WNLA round); each thread also gets more items per inversion.
struct InvList { int n; int in[8]; int out[8]; };
BPPP_HD void batch_inv_list_strided(const WS &w, const InvList &L, size_t t, size_t T) {
    const size_t total = (size_t)L.n * w.n;
    Fe run = fe_one();
#pragma unroll 1
    for (size_t idx = t; idx < total; idx += T) {
        size_t f = idx / w.n, i = idx - f * w.n;
        Fe z = ws_ld_fe(w, i, L.in[f]);
        ws_st_fe(w, i, L.out[f], run);
        if (!fe_is_zero(z)) run = fe_mul(run, z);
    }
    Fe rinv = fe_inv(run);
    size_t cnt = total > t ? (total - t + T - 1) / T : 0;
#pragma unroll 1
    for (size_t k = cnt; k-- > 0;) {
        size_t idx = t + k * T;
        size_t f = idx / w.n, i = idx - f * w.n;
        Fe z = ws_ld_fe(w, i, L.in[f]);
        Fe pre = ws_ld_fe(w, i, L.out[f]);
        if (fe_is_zero(z)) { Fe zero = fe_zero(); ws_st_fe(w, i, L.out[f], zero); }
        else { ws_st_fe(w, i, L.out[f], fe_mul(rinv, pre)); rinv = fe_mul(rinv, z); }
    }
}
// The same over `nfields` Fe fields per proof (field f: input at in_off + f * stride, output at out_off + f * stride):
// flat item index = f * n + proof, so consecutive threads still touch consecutive proofs.
BPPP_HD void batch_inv_multi_strided(const WS &w, int in_off, int out_off, int stride, int nfields, size_t t, size_t T) {
    const size_t total = (size_t)nfields * w.n;
    Fe run = fe_one();
#pragma unroll 1
    for (size_t idx = t; idx < total; idx += T) {
        size_t f = idx / w.n, i = idx - f * w.n;
        Fe z = ws_ld_fe(w, i, in_off + (int)f * stride);
        ws_st_fe(w, i, out_off + (int)f * stride, run);
        if (!fe_is_zero(z)) run = fe_mul(run, z);
    }
    Fe rinv = fe_inv(run);
    size_t cnt = total > t ? (total - t + T - 1) / T : 0;
#pragma unroll 1
    for (size_t k = cnt; k-- > 0;) {
        size_t idx = t + k * T;
        size_t f = idx / w.n, i = idx - f * w.n;
        Fe z = ws_ld_fe(w, i, in_off + (int)f * stride);
        Fe pre = ws_ld_fe(w, i, out_off + (int)f * stride);
        if (fe_is_zero(z)) { Fe zero = fe_zero(); ws_st_fe(w, i, out_off + (int)f * stride, zero); }
        else { ws_st_fe(w, i, out_off + (int)f * stride, fe_mul(rinv, pre)); rinv = fe_mul(rinv, z); }
    }
}
}  // namespace bppp
