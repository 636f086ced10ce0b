// libbppp.so, generic arithmetic-circuit / reciprocal translation unit: ArithmeticCircuit::{commit, prove, verify}
// (reference src/circuit.rs:146-653) for arbitrary dimensions, dense W_m / W_l and a tabulated partition function,
// and ReciprocalRangeProofProtocol::{commit_value, commit_poles, prove, verify, make_circuit}
// (src/range_proof/reciprocal.rs:88-214) for arbitrary (dim_nd, dim_np) on top of it.
//
// Division of labour (north_star): the host drives the Merlin transcript and the scalar-field (mod n) algebra of
// the coefficient vectors -- the same sc.cuh code the kernels run -- while every elliptic-curve operation
// (all commitments, the verifier's recombination, the whole WNLA) runs on the GPU through engine_msm.cu /
// engine_wnla.cu.  The batched u64 fast path (engine_prove.cu / engine_verify.cu) is the specialisation of this
// file with closed-form coefficients and on-device transcripts.
#define BPPP_FE_NOINLINE 1
#include <algorithm>
#include "engine_generic.cuh"

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

namespace {

typedef std::vector<Sc> SV;

// ---- util.rs on the host (scalars only; points go to the device) ----
Sc sv_get(const SV &a, size_t i) { return i < a.size() ? a[i] : sc_zero(); }                         // vector_extend, util.rs:24-26
Sc vmul(const SV &a, const SV &b) {                                                                   // vector_mul, util.rs:46-60
    Sc r = sc_zero(); size_t m = std::max(a.size(), b.size());
    for (size_t i = 0; i < m; i++) r = sc_add(r, sc_mul(sv_get(a, i), sv_get(b, i)));
    return r;
}
Sc wvmul(const SV &a, const SV &b, const Sc &w) {                                                     // weight_vector_mul, util.rs:28-44
    Sc r = sc_zero(), e = sc_one(); size_t m = std::max(a.size(), b.size());
    for (size_t i = 0; i < m; i++) { e = sc_mul(e, w); r = sc_add(r, sc_mul(sv_get(a, i), sc_mul(sv_get(b, i), e))); }
    return r;
}
SV vscale(const SV &a, const Sc &s) { SV r(a.size()); for (size_t i = 0; i < a.size(); i++) r[i] = sc_mul(a[i], s); return r; }   // util.rs:62-67
SV vadd(const SV &a, const SV &b) { size_t m = std::max(a.size(), b.size()); SV r(m); for (size_t i = 0; i < m; i++) r[i] = sc_add(sv_get(a, i), sv_get(b, i)); return r; }
SV vsub(const SV &a, const SV &b) { size_t m = std::max(a.size(), b.size()); SV r(m); for (size_t i = 0; i < m; i++) r[i] = sc_sub(sv_get(a, i), sv_get(b, i)); return r; }
SV e_pow(const Sc &v, size_t n) { SV r(n); Sc b = sc_one(); for (size_t i = 0; i < n; i++) { r[i] = b; b = sc_mul(b, v); } return r; }      // util.rs:87-95
Sc pow_u64(Sc base, uint64_t e) { Sc acc = sc_one(); while (e) { if (e & 1) acc = sc_mul(acc, base); base = sc_sqr(base); e >>= 1; } return acc; }
SV concat(const SV &a, const SV &b) { SV r(a); r.insert(r.end(), b.begin(), b.end()); return r; }
SV slice(const SV &a, size_t from, size_t to) { return SV(a.begin() + (long)std::min(from, a.size()), a.begin() + (long)std::min(to, a.size())); }
SV tensor(const SV &a, const SV &b) { SV r; for (auto &x : b) { SV t = vscale(a, x); r.insert(r.end(), t.begin(), t.end()); } return r; }     // util.rs:111-116
Sc sc_minus(const Sc &v) { return sc_neg(v); }                                                        // minus, util.rs:153-155

struct Mat { size_t rows = 0, cols = 0; SV v; Sc at(size_t i, size_t j) const { return v[i * cols + j]; } };
// vector_mul_on_matrix (util.rs:134-142): out[j] = sum_i a[i] m[i][j] with zero-extension; 0 / 1 entries short-cut
SV vmat(const SV &a, const Mat &m) {
    SV r(m.cols, sc_zero());
    const Sc one = sc_one();
    size_t rows = std::min(a.size(), m.rows);
    for (size_t i = 0; i < rows; i++) {
        if (sc_is_zero(a[i])) continue;
        for (size_t j = 0; j < m.cols; j++) {
            const Sc &e = m.v[i * m.cols + j];
            if (sc_is_zero(e)) continue;
            r[j] = sc_add(r[j], sc_eq(e, one) ? a[i] : sc_mul(a[i], e));
        }
    }
    return r;
}

// the same over the column block [from, to) of the first `rows` rows of m, without materialising the block
SV vmat_cols(const SV &a, const Mat &m, size_t rows_m, size_t from, size_t to) {
    SV r(to - from, sc_zero());
    const Sc one = sc_one();
    size_t rows = std::min(a.size(), rows_m);
    for (size_t i = 0; i < rows; i++) {
        if (sc_is_zero(a[i])) continue;
        const Sc *row = &m.v[i * m.cols];
        for (size_t j = from; j < to; j++) {
            const Sc &e = row[j];
            if (sc_is_zero(e)) continue;
            r[j - from] = sc_add(r[j - from], sc_eq(e, one) ? a[i] : sc_mul(a[i], e));
        }
    }
    return r;
}

struct Panic { int32_t code; };
Sc inv_or_panic(const Sc &a) { if (sc_is_zero(a)) throw Panic{ST_PANIC_INVERT_ZERO}; return sc_inv(a); }

struct Circuit {
    size_t dim_nm = 0, dim_no = 0, k = 0, dim_nl = 0, dim_nv = 0, dim_nw = 0;
    bool f_l = false, f_m = false;
    Mat W_m, W_l; SV a_m, a_l;
    std::vector<int32_t> part[4]; // LO, LL, LR, NO
    std::vector<uint8_t> g64, gvec64, hvec64, gvec2_64, hvec2_64;
    // device: [h_vec | g_vec | g]
    uint32_t *d_pts = nullptr; size_t hn = 0, gn = 0;
    int part_get(int typ, size_t j) const { return j < part[typ].size() ? part[typ][j] : -1; }
    void release() { cudaFree(d_pts); d_pts = nullptr; }
};
enum { P_LO = 0, P_LL = 1, P_LR = 2, P_NO = 3 };

int circuit_upload(Circuit &c) {
    c.hn = c.hvec64.size() / 64; c.gn = c.gvec64.size() / 64;
    std::vector<uint8_t> pb(c.hvec64);
    pb.insert(pb.end(), c.gvec64.begin(), c.gvec64.end());
    pb.insert(pb.end(), c.g64.begin(), c.g64.end());
    return decode_points_to_device(nullptr, pb.data(), FMT_AFFINE64, c.hn + c.gn + 1, &c.d_pts);
}
// <h_vec, hs> + <g_vec, gs> + gsc * g  on the device -> 33 bytes (+ projective copy when d_out30 != null)
int commit_hg(const Circuit &c, const SV &hs, const SV &gs, const Sc &gsc, uint8_t *out33, uint32_t *d_out30 = nullptr) {
    size_t n = c.hn + c.gn + 1;
    std::vector<uint32_t> sc(8 * n, 0);
    for (size_t i = 0; i < c.hn && i < hs.size(); i++) memcpy(&sc[8 * i], hs[i].v, 32);
    for (size_t i = 0; i < c.gn && i < gs.size(); i++) memcpy(&sc[8 * (c.hn + i)], gs[i].v, 32);
    memcpy(&sc[8 * (c.hn + c.gn)], gsc.v, 32);
    uint32_t *d_sc = nullptr, *d_o = d_out30;
    CUDA_OK(cudaMalloc(&d_sc, 32 * n));
    if (!d_o) CUDA_OK(cudaMalloc(&d_o, PT_BYTES));
    CUDA_OK(cudaMemcpy(d_sc, sc.data(), 32 * n, cudaMemcpyHostToDevice));
    int rc = msm_device(nullptr, c.d_pts, d_sc, n, nullptr, d_o);
    if (rc == BPPP_OK && out33) rc = encode_points_from_device(nullptr, d_o, 1, FMT_COMPRESSED, out33);
    cudaFree(d_sc); if (!d_out30) cudaFree(d_o);
    return rc;
}

// ---- circuit.rs:559-653 ----
Sc linear_comb_coef(const Circuit &c, size_t i, const Sc &lambda, const Sc &mu) {
    Sc coef = sc_zero();
    if (c.f_l) coef = sc_add(coef, pow_u64(lambda, c.dim_nv * i));
    if (c.f_m) coef = sc_add(coef, pow_u64(mu, c.dim_nv * i + 1));
    return coef;
}
SV collect_cl0(const Circuit &c, const Sc &lambda, const Sc &mu) {
    SV r(c.dim_nv - 1, sc_zero());
    if (c.f_l) r = slice(e_pow(lambda, c.dim_nv), 1, c.dim_nv);
    if (c.f_m) r = vsub(r, vscale(slice(e_pow(mu, c.dim_nv), 1, c.dim_nv), mu));
    return r;
}
SV collect_lambda(const Circuit &c, const Sc &lambda, const Sc &mu) {
    SV lv = e_pow(lambda, c.dim_nl);
    if (c.f_l && c.f_m)
        lv = vsub(lv, vadd(tensor(vscale(e_pow(lambda, c.dim_nv), mu), e_pow(pow_u64(mu, c.dim_nv), c.k)),
                           tensor(e_pow(mu, c.dim_nv), e_pow(pow_u64(lambda, c.dim_nv), c.k))));
    return lv;
}
Mat sub_cols(const Mat &W, size_t rows, size_t from, size_t to) {
    Mat m; m.rows = rows; m.cols = to - from; m.v.resize(m.rows * m.cols);
    for (size_t i = 0; i < rows; i++) for (size_t j = from; j < to; j++) m.v[i * m.cols + (j - from)] = W.v[i * W.cols + j];
    return m;
}
// vector_mul_on_matrix(a, map_f(..)) of circuit.rs:627-653 without building map_f's isz x jsz matrix (dim_nl x dim_nv scalars,
// almost all zero): out[j] = sum_i a[i] Wx[i][partition(typ, j)] for the j the partition maps, zero elsewhere
SV vmat_mapped(const SV &a, const Circuit &c, size_t isz, size_t jsz, int typ, const Mat &Wx) {
    SV r(jsz, sc_zero());
    const Sc one = sc_one();
    size_t rows = std::min(a.size(), isz);
    for (size_t j = 0; j < jsz; j++) {
        int j_ = c.part_get(typ, j);
        if (j_ < 0 || (size_t)j_ >= Wx.cols) continue;
        Sc acc = sc_zero();
        for (size_t i = 0; i < rows; i++) {
            const Sc &e = Wx.v[i * Wx.cols + (size_t)j_];
            if (sc_is_zero(e) || sc_is_zero(a[i])) continue;
            acc = sc_add(acc, sc_eq(e, one) ? a[i] : sc_mul(a[i], e));
        }
        r[j] = acc;
    }
    return r;
}
struct Coefs { SV nL, nR, nO, lL, lR, lO; };
Coefs collect_c(const Circuit &c, const SV &lambda_vec, const SV &mu_vec, const Sc &mu) {
    size_t nm = c.dim_nm;
    Mat W_lO = sub_cols(c.W_l, c.dim_nl, 2 * nm, c.W_l.cols), W_mO = sub_cols(c.W_m, c.dim_nm, 2 * nm, c.W_m.cols);
    // diag_inv(mu, nm) (util.rs:118-132) applied as a diagonal scaling
    Sc mu_inv = inv_or_panic(mu);
    SV dinv(nm); Sc val = sc_one();
    for (size_t i = 0; i < nm; i++) { val = sc_mul(val, mu_inv); dinv[i] = val; }
    auto scale_diag = [&](SV v) { v.resize(nm, sc_zero()); for (size_t j = 0; j < nm; j++) v[j] = sc_mul(v[j], dinv[j]); return v; };
    Coefs r;
    r.nL = scale_diag(vsub(vmat_cols(lambda_vec, c.W_l, c.dim_nl, 0, nm), vmat_cols(mu_vec, c.W_m, c.dim_nm, 0, nm)));
    r.nR = scale_diag(vsub(vmat_cols(lambda_vec, c.W_l, c.dim_nl, nm, 2 * nm), vmat_cols(mu_vec, c.W_m, c.dim_nm, nm, 2 * nm)));
    r.nO = scale_diag(vsub(vmat_mapped(lambda_vec, c, c.dim_nl, c.dim_nm, P_NO, W_lO), vmat_mapped(mu_vec, c, c.dim_nm, c.dim_nm, P_NO, W_mO)));
    r.lL = vsub(vmat_mapped(lambda_vec, c, c.dim_nl, c.dim_nv, P_LL, W_lO), vmat_mapped(mu_vec, c, c.dim_nm, c.dim_nv, P_LL, W_mO));
    r.lR = vsub(vmat_mapped(lambda_vec, c, c.dim_nl, c.dim_nv, P_LR, W_lO), vmat_mapped(mu_vec, c, c.dim_nm, c.dim_nv, P_LR, W_mO));
    r.lO = vsub(vmat_mapped(lambda_vec, c, c.dim_nl, c.dim_nv, P_LO, W_lO), vmat_mapped(mu_vec, c, c.dim_nm, c.dim_nv, P_LO, W_mO));
    return r;
}
SV make_cr_tau(const Sc &tau, const Sc &tau_inv, const Sc &tau2, const Sc &tau3, const Sc &beta) {
    return SV{sc_one(), sc_mul(tau_inv, beta), sc_mul(tau, beta), sc_mul(tau2, beta), sc_mul(tau3, beta), sc_mul(sc_mul(tau, tau3), beta),
              sc_mul(sc_mul(tau2, tau3), beta), sc_mul(sc_mul(tau3, tau3), beta), sc_mul(sc_mul(sc_mul(tau3, tau3), tau), beta)};
}
bool challenge(Merlin &t, const char *label, uint32_t ll, Sc &out) { if (!merlin_challenge_scalar(t, label, ll, out)) throw Panic{ST_PANIC_CHALLENGE_RANGE}; return true; }

struct ByteRng {
    const uint8_t *p; size_t len, pos = 0; bool exhausted = false;
    Sc draw() { if (pos + 64 > len) { exhausted = true; return sc_zero(); } Sc s = sc_from_wide_be64(p + pos); pos += 64; return s; }   // generate_biased
};

struct CircuitProofHost { uint8_t cl[33], cr[33], co[33], cs[33]; WnlaProofHost w; };

int wnla_from_circuit(const Circuit &c, const SV &cvec, const Sc &rho, const Sc &mu, size_t ln, size_t nn, WnlaDev &w) {
    std::vector<uint8_t> hv(c.hvec64); hv.insert(hv.end(), c.hvec2_64.begin(), c.hvec2_64.end());
    std::vector<uint8_t> gv(c.gvec64); gv.insert(gv.end(), c.gvec2_64.begin(), c.gvec2_64.end());
    std::vector<uint8_t> cb(32 * (cvec.size() ? cvec.size() : 1));
    for (size_t i = 0; i < cvec.size(); i++) sc_to_be32(&cb[32 * i], cvec[i]);
    uint8_t rb[32], mb[32]; sc_to_be32(rb, rho); sc_to_be32(mb, mu);
    return wnla_load(nullptr, w, c.g64.data(), gv.data(), gv.size() / 64, hv.data(), hv.size() / 64, cb.data(), cvec.size(), rb, mb, ln, nn);
}

// circuit.rs:260-556
int circuit_prove(const Circuit &c, const std::vector<std::vector<uint8_t>> &v33, const std::vector<SV> &wv, const SV &s_v, const SV &w_l, const SV &w_r, const SV &w_o,
                  Merlin &t, ByteRng &rng, CircuitProofHost &proof) {
    auto draw = [&]() { return rng.draw(); };
    SV ro{draw(), draw(), draw(), draw(), sc_zero(), draw(), draw(), draw(), sc_zero()};
    SV rl{draw(), draw(), draw(), sc_zero(), draw(), draw(), draw(), sc_zero(), sc_zero()};
    SV rr{draw(), draw(), sc_zero(), draw(), draw(), draw(), sc_zero(), sc_zero(), sc_zero()};
    const SV &nl = w_l, &nr = w_r;
    auto part_vec = [&](int typ, size_t size) { SV r(size, sc_zero()); for (size_t j = 0; j < size; j++) { int i = c.part_get(typ, j); if (i >= 0) r[j] = w_o[(size_t)i]; } return r; };
    SV no = part_vec(P_NO, c.dim_nm), lo = part_vec(P_LO, c.dim_nv), ll = part_vec(P_LL, c.dim_nv), lr = part_vec(P_LR, c.dim_nv);
    int rc;
    if ((rc = commit_hg(c, concat(ro, lo), no, sc_zero(), proof.co)) != BPPP_OK) return rc;
    if ((rc = commit_hg(c, concat(rl, ll), nl, sc_zero(), proof.cl)) != BPPP_OK) return rc;
    if ((rc = commit_hg(c, concat(rr, lr), nr, sc_zero(), proof.cr)) != BPPP_OK) return rc;
    merlin_append(t, BPPP_LBL("commitment_cl"), proof.cl, 33); merlin_append(t, BPPP_LBL("commitment_cr"), proof.cr, 33);
    merlin_append(t, BPPP_LBL("commitment_co"), proof.co, 33);
    for (auto &v : v33) merlin_append(t, BPPP_LBL("commitment_v"), v.data(), 33);
    Sc rho, lambda, beta, delta;
    challenge(t, BPPP_LBL("circuit_rho"), rho); challenge(t, BPPP_LBL("circuit_lambda"), lambda);
    challenge(t, BPPP_LBL("circuit_beta"), beta); challenge(t, BPPP_LBL("circuit_delta"), delta);
    Sc mu = sc_sqr(rho);
    SV lambda_vec = collect_lambda(c, lambda, mu), mu_vec = vscale(e_pow(mu, c.dim_nm), mu);
    Coefs cc = collect_c(c, lambda_vec, mu_vec, mu);
    SV ls(c.dim_nv), ns(c.dim_nm);
    for (auto &x : ls) x = draw();
    for (auto &x : ns) x = draw();
    Sc two = sc_from_u64(2), v_0 = sc_zero();
    SV rv(9, sc_zero()), v_1(c.dim_nv - 1, sc_zero());
    for (size_t i = 0; i < c.k; i++) {
        Sc cf = linear_comb_coef(c, i, lambda, mu);
        v_0 = sc_add(v_0, sc_mul(wv[i][0], cf));
        rv[0] = sc_add(rv[0], sc_mul(s_v[i], cf));
        v_1 = vadd(v_1, vscale(slice(wv[i], 1, wv[i].size()), cf));
    }
    v_0 = sc_mul(v_0, two); rv[0] = sc_mul(rv[0], two); v_1 = vscale(v_1, two);
    SV c_l0 = collect_cl0(c, lambda, mu);
    Sc delta2 = sc_sqr(delta), delta_inv = inv_or_panic(delta);
    auto W = [&](const SV &a, const SV &b) { return wvmul(a, b, mu); };
    auto m2 = [&](const Sc &x) { return sc_mul(x, two); };
    Sc f_[8];
    f_[0] = sc_minus(W(ns, ns));                                                                                            // circuit.rs:406
    f_[1] = sc_add(vmul(c_l0, ls), sc_mul(sc_mul(delta, two), W(ns, no)));                                                  // :409-410
    f_[2] = sc_sub(sc_sub(sc_sub(sc_minus(m2(vmul(cc.lR, ls))), sc_mul(vmul(c_l0, lo), delta)), m2(W(ns, vadd(nl, cc.nR)))),
                   sc_mul(W(no, no), delta2));                                                                              // :413-416
    f_[3] = sc_add(sc_add(sc_add(sc_add(m2(vmul(cc.lL, ls)), m2(sc_mul(vmul(cc.lR, lo), delta))), vmul(c_l0, ll)), m2(W(ns, vadd(nr, cc.nL)))),
                   sc_mul(m2(W(no, vadd(nl, cc.nR))), delta));                                                              // :419-423
    f_[4] = W(cc.nR, cc.nR);                                                                                                // :426-433
    f_[4] = sc_sub(f_[4], m2(sc_mul(vmul(cc.lO, ls), delta_inv)));
    f_[4] = sc_sub(f_[4], m2(sc_mul(vmul(cc.lL, lo), delta)));
    f_[4] = sc_sub(f_[4], m2(vmul(cc.lR, ll)));
    f_[4] = sc_sub(f_[4], vmul(c_l0, lr));
    f_[4] = sc_sub(f_[4], m2(sc_mul(W(ns, cc.nO), delta_inv)));
    f_[4] = sc_sub(f_[4], m2(sc_mul(W(no, vadd(nr, cc.nL)), delta)));
    f_[4] = sc_sub(f_[4], W(vadd(nl, cc.nR), vadd(nl, cc.nR)));
    f_[5] = sc_add(m2(sc_mul(W(cc.nO, cc.nR), delta_inv)), W(cc.nL, cc.nL));                                                // :438-444
    f_[5] = sc_sub(f_[5], m2(sc_mul(vmul(cc.lO, ll), delta_inv)));
    f_[5] = sc_sub(f_[5], m2(vmul(cc.lL, lr)));
    f_[5] = sc_sub(f_[5], m2(vmul(cc.lR, v_1)));
    f_[5] = sc_sub(f_[5], m2(sc_mul(W(vadd(nl, cc.nR), cc.nO), delta_inv)));
    f_[5] = sc_sub(f_[5], W(vadd(nr, cc.nL), vadd(nr, cc.nL)));
    f_[6] = sc_minus(m2(sc_mul(W(cc.nO, cc.nL), delta_inv)));                                                               // :447-450
    f_[6] = sc_add(f_[6], m2(sc_mul(vmul(cc.nO, lr), delta_inv)));
    f_[6] = sc_add(f_[6], m2(vmul(cc.lL, v_1)));
    f_[6] = sc_add(f_[6], m2(sc_mul(W(vadd(nr, cc.nL), cc.nO), delta_inv)));
    f_[7] = sc_minus(m2(sc_mul(vmul(cc.lO, v_1), delta_inv)));                                                              // :453
    Sc beta_inv = inv_or_panic(beta);
    SV rs(9);                                                                                                               // :457-467
    rs[0] = sc_add(f_[1], sc_mul(sc_mul(ro[1], delta), beta));
    rs[1] = sc_mul(f_[0], beta_inv);
    rs[2] = sc_sub(sc_mul(sc_add(sc_mul(ro[0], delta), f_[2]), beta_inv), rl[1]);
    rs[3] = sc_add(sc_mul(sc_sub(f_[3], rl[0]), beta_inv), sc_add(sc_mul(ro[2], delta), rr[1]));
    rs[4] = sc_add(sc_mul(sc_add(f_[4], rr[0]), beta_inv), sc_sub(sc_mul(ro[3], delta), rl[2]));
    rs[5] = sc_minus(sc_mul(rv[0], beta_inv));
    rs[6] = sc_sub(sc_add(sc_add(sc_mul(f_[5], beta_inv), sc_mul(ro[5], delta)), rr[3]), rl[4]);
    rs[7] = sc_sub(sc_add(sc_add(sc_mul(f_[6], beta_inv), rr[4]), sc_mul(ro[6], delta)), rl[5]);
    rs[8] = sc_add(sc_sub(sc_add(sc_mul(f_[7], beta_inv), sc_mul(ro[7], delta)), rl[6]), rr[5]);
    if ((rc = commit_hg(c, concat(rs, ls), ns, sc_zero(), proof.cs)) != BPPP_OK) return rc;                                 // :469-470
    merlin_append(t, BPPP_LBL("commitment_cs"), proof.cs, 33);
    Sc tau; challenge(t, BPPP_LBL("circuit_tau"), tau);
    Sc tau_inv = inv_or_panic(tau), tau2 = sc_sqr(tau), tau3 = sc_mul(tau2, tau), t3d = sc_mul(tau3, delta_inv);
    SV l = vscale(concat(rs, ls), tau_inv);                                                                                 // :479-483
    l = vsub(l, vscale(concat(ro, lo), delta));
    l = vadd(l, vscale(concat(rl, ll), tau));
    l = vsub(l, vscale(concat(rr, lr), tau2));
    l = vadd(l, vscale(concat(rv, v_1), tau3));
    SV pn_tau = vadd(vsub(vscale(cc.nO, t3d), vscale(cc.nL, tau2)), vscale(cc.nR, tau));
    Sc ps_tau = sc_sub(sc_add(W(pn_tau, pn_tau), m2(sc_mul(vmul(lambda_vec, c.a_l), tau3))), m2(sc_mul(vmul(mu_vec, c.a_m), tau3)));
    SV n_tau = vsub(vadd(vsub(vscale(ns, tau_inv), vscale(no, delta)), vscale(nl, tau)), vscale(nr, tau2));
    SV n = vadd(pn_tau, n_tau);
    SV cr_tau = make_cr_tau(tau, tau_inv, tau2, tau3, beta);
    SV cl_tau = vsub(vscale(vadd(vsub(vscale(cc.lO, t3d), vscale(cc.lL, tau2)), vscale(cc.lR, tau)), two), c_l0);
    SV cvec = concat(cr_tau, cl_tau);
    Sc vv = sc_add(ps_tau, sc_mul(tau3, v_0));
    uint32_t *d_com30 = nullptr;
    CUDA_OK(cudaMalloc(&d_com30, PT_BYTES));
    if ((rc = commit_hg(c, l, n, vv, nullptr, d_com30)) != BPPP_OK) { cudaFree(d_com30); return rc; }                       // :522-524
    size_t hn_all = (c.hvec64.size() + c.hvec2_64.size()) / 64, gn_all = (c.gvec64.size() + c.gvec2_64.size()) / 64;
    while (l.size() < hn_all) { l.push_back(sc_zero()); cvec.push_back(sc_zero()); }                                        // :526-529
    while (n.size() < gn_all) n.push_back(sc_zero());                                                                       // :531-533
    WnlaDev w;
    if ((rc = wnla_from_circuit(c, cvec, rho, mu, l.size(), n.size(), w)) != BPPP_OK) { cudaFree(d_com30); return rc; }
    std::vector<uint8_t> lb(32 * l.size()), nb(32 * n.size());
    for (size_t i = 0; i < l.size(); i++) sc_to_be32(&lb[32 * i], l[i]);
    for (size_t i = 0; i < n.size(); i++) sc_to_be32(&nb[32 * i], n[i]);
    uint32_t *d_l = nullptr, *d_n = nullptr;
    rc = upload_padded_scalars(nullptr, lb.data(), l.size(), w.Lh, &d_l);
    if (rc == BPPP_OK) rc = upload_padded_scalars(nullptr, nb.data(), n.size(), w.Lg, &d_n);
    int32_t st = ST_TRUE;
    if (rc == BPPP_OK) rc = wnla_prove_dev(nullptr, w, t, d_com30, d_l, d_n, l.size(), n.size(), proof.w, &st);
    cudaFree(d_com30); w.release();
    if (rc == BPPP_OK && st != ST_TRUE) throw Panic{st};
    if (rng.exhausted) return fail(BPPP_ERR_ARG, "rng buffer too short");
    return rc;
}

// circuit.rs:154-256
int circuit_verify(const Circuit &c, const std::vector<std::vector<uint8_t>> &v33, Merlin &t, const uint8_t *cl, const uint8_t *cr, const uint8_t *co, const uint8_t *cs,
                   const uint8_t *r33, size_t rn, const uint8_t *x33, size_t xn, const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn, int32_t *verdict) {
    merlin_append(t, BPPP_LBL("commitment_cl"), cl, 33); merlin_append(t, BPPP_LBL("commitment_cr"), cr, 33); merlin_append(t, BPPP_LBL("commitment_co"), co, 33);
    for (auto &v : v33) merlin_append(t, BPPP_LBL("commitment_v"), v.data(), 33);
    Sc rho, lambda, beta, delta;
    challenge(t, BPPP_LBL("circuit_rho"), rho); challenge(t, BPPP_LBL("circuit_lambda"), lambda);
    challenge(t, BPPP_LBL("circuit_beta"), beta); challenge(t, BPPP_LBL("circuit_delta"), delta);
    Sc mu = sc_sqr(rho);
    SV lambda_vec = collect_lambda(c, lambda, mu), mu_vec = vscale(e_pow(mu, c.dim_nm), mu);
    Coefs cc = collect_c(c, lambda_vec, mu_vec, mu);
    Sc two = sc_from_u64(2);
    merlin_append(t, BPPP_LBL("commitment_cs"), cs, 33);
    Sc tau; challenge(t, BPPP_LBL("circuit_tau"), tau);
    Sc tau_inv = inv_or_panic(tau), tau2 = sc_sqr(tau), tau3 = sc_mul(tau2, tau);
    Sc delta_inv = inv_or_panic(delta), t3d = sc_mul(tau3, delta_inv);
    SV pn_tau = vadd(vsub(vscale(cc.nO, t3d), vscale(cc.nL, tau2)), vscale(cc.nR, tau));
    Sc ps_tau = sc_sub(sc_add(wvmul(pn_tau, pn_tau, mu), sc_mul(sc_mul(vmul(lambda_vec, c.a_l), tau3), two)), sc_mul(sc_mul(vmul(mu_vec, c.a_m), tau3), two));
    uint32_t *d_pt30 = nullptr, *d_com30 = nullptr;
    CUDA_OK(cudaMalloc(&d_pt30, PT_BYTES)); CUDA_OK(cudaMalloc(&d_com30, PT_BYTES));
    int rc = commit_hg(c, SV(), pn_tau, ps_tau, nullptr, d_pt30);                                      // pt, circuit.rs:206
    SV cr_tau = make_cr_tau(tau, tau_inv, tau2, tau3, beta);
    SV c_l0 = collect_cl0(c, lambda, mu);
    SV cl_tau = vsub(vscale(vadd(vsub(vscale(cc.lO, t3d), vscale(cc.lL, tau2)), vscale(cc.lR, tau)), two), c_l0);
    SV cvec = concat(cr_tau, cl_tau);
    // commitment = pt + tau^-1 c_s - delta c_o + tau c_l - tau^2 c_r + tau^3 * 2 * sum_i coef_i v_i   (circuit.rs:182-187,230-235)
    if (rc == BPPP_OK) {
        size_t np = 4 + c.k;
        std::vector<uint8_t> pb(33 * np), sb(32 * np);
        memcpy(&pb[0], cs, 33); memcpy(&pb[33], co, 33); memcpy(&pb[66], cl, 33); memcpy(&pb[99], cr, 33);
        sc_to_be32(&sb[0], tau_inv); sc_to_be32(&sb[32], sc_neg(delta)); sc_to_be32(&sb[64], tau); sc_to_be32(&sb[96], sc_neg(tau2));
        for (size_t i = 0; i < c.k; i++) {
            memcpy(&pb[33 * (4 + i)], v33[i].data(), 33);
            sc_to_be32(&sb[32 * (4 + i)], sc_mul(sc_mul(linear_comb_coef(c, i, lambda, mu), two), tau3));
        }
        uint32_t *d_p = nullptr, *d_s = nullptr;
        rc = decode_points_to_device(nullptr, pb.data(), FMT_COMPRESSED, np, &d_p);
        if (rc != BPPP_OK) { *verdict = ST_BAD_POINT; cudaFree(d_pt30); cudaFree(d_com30); return BPPP_OK; }
        rc = decode_scalars_to_device(nullptr, sb.data(), np, &d_s);
        if (rc == BPPP_OK) rc = msm_device(nullptr, d_p, d_s, np, d_pt30, d_com30);
        cudaFree(d_p); cudaFree(d_s);
    }
    size_t hn_all = (c.hvec64.size() + c.hvec2_64.size()) / 64;
    while (cvec.size() < hn_all) cvec.push_back(sc_zero());                                               // circuit.rs:237-239
    WnlaDev w;
    if (rc == BPPP_OK) rc = wnla_from_circuit(c, cvec, rho, mu, 0, 0, w);
    if (rc == BPPP_OK) rc = wnla_verify_dev(nullptr, w, t, d_com30, r33, rn, x33, xn, l32, ln, n32, nn, verdict);
    cudaFree(d_pt30); cudaFree(d_com30); w.release();
    return rc;
}

int load_scalars(SV &out, const uint8_t *b, size_t n) { out.resize(n); for (size_t i = 0; i < n; i++) if (!sc_from_be32(out[i], b + 32 * i)) return fail(BPPP_ERR_ARG, "a scalar is not canonical (>= n)"); return BPPP_OK; }

int pick_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BPPP_ERR_ARG, "bad device index");
    CUDA_OK(cudaSetDevice(device));
    return BPPP_OK;
}

int circuit_from_desc(Circuit &c, const bppp_circuit_desc *d) {
    c.dim_nm = d->dim_nm; c.dim_no = d->dim_no; c.k = d->k; c.dim_nv = d->dim_nv; c.dim_nl = d->dim_nv * d->k; c.dim_nw = 2 * d->dim_nm + d->dim_no;
    c.f_l = d->f_l != 0; c.f_m = d->f_m != 0;
    c.g64.assign(d->g64, d->g64 + 64);
    c.gvec64.assign(d->gvec64, d->gvec64 + 64 * d->gn); c.hvec64.assign(d->hvec64, d->hvec64 + 64 * d->hn);
    c.gvec2_64.assign(d->gvec2_64, d->gvec2_64 + 64 * d->gn2); c.hvec2_64.assign(d->hvec2_64, d->hvec2_64 + 64 * d->hn2);
    c.W_m.rows = c.dim_nm; c.W_m.cols = c.dim_nw; c.W_l.rows = c.dim_nl; c.W_l.cols = c.dim_nw;
    int rc;
    if ((rc = load_scalars(c.W_m.v, d->W_m32, c.dim_nm * c.dim_nw)) != BPPP_OK) return rc;
    if ((rc = load_scalars(c.W_l.v, d->W_l32, c.dim_nl * c.dim_nw)) != BPPP_OK) return rc;
    if ((rc = load_scalars(c.a_m, d->a_m32, c.dim_nm)) != BPPP_OK) return rc;
    if ((rc = load_scalars(c.a_l, d->a_l32, c.dim_nl)) != BPPP_OK) return rc;
    const int32_t *parts[4] = {d->part_lo, d->part_ll, d->part_lr, d->part_no};
    for (int t = 0; t < 4; t++) c.part[t].assign(parts[t], parts[t] + d->part_n);
    return circuit_upload(c);
}

size_t write_circuit_record(uint8_t *out, const CircuitProofHost &p) {
    uint8_t *o = out;
    memcpy(o, p.cl, 33); memcpy(o + 33, p.cr, 33); memcpy(o + 66, p.co, 33); memcpy(o + 99, p.cs, 33); o += 132;
    memcpy(o, p.w.r33.data(), p.w.r33.size()); o += p.w.r33.size();
    memcpy(o, p.w.x33.data(), p.w.x33.size()); o += p.w.x33.size();
    memcpy(o, p.w.l32.data(), p.w.l32.size()); o += p.w.l32.size();
    memcpy(o, p.w.n32.data(), p.w.n32.size()); o += p.w.n32.size();
    return (size_t)(o - out);
}

// reciprocal.rs:150-214 with the np distinct inverses computed once (the reference recomputes them per row)
void make_reciprocal_circuit(Circuit &c, size_t nd, size_t np, const Sc &e) {
    c.dim_nm = nd; c.dim_no = np; c.k = 1; c.dim_nv = nd + 1; c.dim_nl = nd + 1; c.dim_nw = 2 * nd + np;
    c.f_l = true; c.f_m = false;
    c.a_m.assign(nd, sc_one()); c.a_l.assign(nd + 1, sc_zero());
    c.W_m.rows = nd; c.W_m.cols = c.dim_nw; c.W_m.v.assign(nd * c.dim_nw, sc_zero());
    Sc me = sc_neg(e);
    for (size_t i = 0; i < nd; i++) c.W_m.v[i * c.dim_nw + i + nd] = me;
    c.W_l.rows = nd + 1; c.W_l.cols = c.dim_nw; c.W_l.v.assign((nd + 1) * c.dim_nw, sc_zero());
    Sc base = sc_from_u64((uint64_t)(uint32_t)np), pw = sc_one();
    for (size_t i = 0; i < nd; i++) { c.W_l.v[i] = sc_neg(pw); pw = sc_mul(pw, base); }
    SV inv(np);
    for (size_t j = 0; j < np; j++) inv[j] = sc_neg(inv_or_panic(sc_add(e, sc_from_u64((uint64_t)(uint32_t)j))));
    for (size_t i = 0; i < nd; i++) {
        Sc *row = &c.W_l.v[(i + 1) * c.dim_nw];
        for (size_t j = 0; j < nd; j++) row[j + nd] = j == i ? sc_zero() : sc_one();
        for (size_t j = 0; j < np; j++) row[j + 2 * nd] = inv[j];
    }
    size_t pn = nd + 1;
    for (int t = 0; t < 4; t++) { c.part[t].assign(pn, -1); }
    for (size_t j = 0; j < pn && j < np; j++) c.part[P_LL][j] = (int32_t)j;
}

}  // namespace

// ArithmeticCircuit::commit (src/circuit.rs:146-151)
extern "C" int bppp_circuit_commit(int device, const bppp_circuit_desc *d, const uint8_t *v32, const uint8_t *s32, uint8_t *out33) {
    if (!d || !v32 || !s32 || !out33) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c; if ((rc = circuit_from_desc(c, d)) != BPPP_OK) return rc;
    SV v, s;
    if ((rc = load_scalars(v, v32, d->dim_nv)) == BPPP_OK && (rc = load_scalars(s, s32, 1)) == BPPP_OK) {
        SV hs(c.hn, sc_zero());
        hs[0] = s[0];
        for (size_t i = 1; i < v.size() && 8 + i < c.hn; i++) hs[8 + i] = v[i];          // <h_vec[9..], v[1..]>
        rc = commit_hg(c, hs, SV(), v[0], out33);
    }
    c.release();
    return rc;
}

// ArithmeticCircuit::prove (src/circuit.rs:260-556), fresh Transcript::new(label).  out: c_l c_r c_o c_s | r | x | l | n
extern "C" int bppp_circuit_prove(int device, const bppp_circuit_desc *d, const uint8_t *commits33, const uint8_t *v32, const uint8_t *sv32, const uint8_t *wl32,
                                  const uint8_t *wr32, const uint8_t *wo32, const uint8_t *rng_bytes, size_t rng_len, const uint8_t *label, size_t label_len,
                                  uint8_t *out, size_t out_cap, size_t *rounds_out, size_t *l_len_out, size_t *n_len_out, int32_t *status) {
    if (!d || !out || !rounds_out || !l_len_out || !n_len_out || !status) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c; if ((rc = circuit_from_desc(c, d)) != BPPP_OK) return rc;
    std::vector<std::vector<uint8_t>> v33(d->k);
    std::vector<SV> wv(d->k);
    SV s_v, w_l, w_r, w_o;
    for (size_t i = 0; i < d->k && rc == BPPP_OK; i++) { v33[i].assign(commits33 + 33 * i, commits33 + 33 * (i + 1)); rc = load_scalars(wv[i], v32 + 32 * d->dim_nv * i, d->dim_nv); }
    if (rc == BPPP_OK) rc = load_scalars(s_v, sv32, d->k);
    if (rc == BPPP_OK) rc = load_scalars(w_l, wl32, d->dim_nm);
    if (rc == BPPP_OK) rc = load_scalars(w_r, wr32, d->dim_nm);
    if (rc == BPPP_OK) rc = load_scalars(w_o, wo32, d->dim_no);
    *status = ST_TRUE;
    if (rc == BPPP_OK) {
        Merlin t; merlin_init(t, label, (uint32_t)label_len);
        ByteRng rng{rng_bytes, rng_len};
        CircuitProofHost proof;
        try { rc = circuit_prove(c, v33, wv, s_v, w_l, w_r, w_o, t, rng, proof); } catch (const Panic &p) { *status = p.code; }
        if (rc == BPPP_OK && *status == ST_TRUE) {
            size_t need = 132 + proof.w.r33.size() + proof.w.x33.size() + proof.w.l32.size() + proof.w.n32.size();
            if (need > out_cap) rc = fail(BPPP_ERR_ARG, "output buffer too small");
            else { write_circuit_record(out, proof); *rounds_out = proof.w.r33.size() / 33; *l_len_out = proof.w.l32.size() / 32; *n_len_out = proof.w.n32.size() / 32; }
        }
    }
    c.release();
    return rc;
}

// ArithmeticCircuit::verify (src/circuit.rs:154-256)
extern "C" int bppp_circuit_verify(int device, const bppp_circuit_desc *d, const uint8_t *commits33, const uint8_t *rec, size_t rounds_r, size_t rounds_x, size_t l_len,
                                   size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict) {
    if (!d || !rec || !verdict) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c; if ((rc = circuit_from_desc(c, d)) != BPPP_OK) return rc;
    std::vector<std::vector<uint8_t>> v33(d->k);
    for (size_t i = 0; i < d->k; i++) v33[i].assign(commits33 + 33 * i, commits33 + 33 * (i + 1));
    Merlin t; merlin_init(t, label, (uint32_t)label_len);
    const uint8_t *r33 = rec + 132, *x33 = r33 + 33 * rounds_r, *l32 = x33 + 33 * rounds_x, *n32 = l32 + 32 * l_len;
    try { rc = circuit_verify(c, v33, t, rec, rec + 33, rec + 66, rec + 99, r33, rounds_r, x33, rounds_x, l32, l_len, n32, n_len, verdict); }
    catch (const Panic &p) { *verdict = p.code; }
    c.release();
    return rc;
}

namespace {
int reciprocal_setup(Circuit &c, size_t nd, size_t np, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn, const uint8_t *gvec2_64, size_t gn2,
                     const uint8_t *hvec2_64, size_t hn2) {
    if (gn < nd || hn < nd + 10) return fail(BPPP_ERR_ARG, "g_vec needs dim_nd points and h_vec dim_nd + 10");   // the reference indexes out of bounds (panic)
    c.g64.assign(g64, g64 + 64); c.gvec64.assign(gvec64, gvec64 + 64 * gn); c.hvec64.assign(hvec64, hvec64 + 64 * hn);
    c.gvec2_64.assign(gvec2_64, gvec2_64 + 64 * gn2); c.hvec2_64.assign(hvec2_64, hvec2_64 + 64 * hn2);
    (void)np;
    return circuit_upload(c);
}
}  // namespace

// ReciprocalRangeProofProtocol::commit_value (src/range_proof/reciprocal.rs:88-90): x g + s h_vec[0]
extern "C" int bppp_reciprocal_commit_value(int device, const uint8_t *g64, const uint8_t *h0_64, const uint8_t *x32, const uint8_t *s32, uint8_t *out33) {
    if (!g64 || !h0_64 || !x32 || !s32 || !out33) return fail(BPPP_ERR_ARG, "null argument");
    uint8_t pts[128], sc[64];
    memcpy(pts, g64, 64); memcpy(pts + 64, h0_64, 64); memcpy(sc, x32, 32); memcpy(sc + 32, s32, 32);
    return bppp_msm(device, pts, FMT_AFFINE64, 2, sc, 2, FMT_COMPRESSED, out33);
}

// ReciprocalRangeProofProtocol::prove (src/range_proof/reciprocal.rs:110-146), fresh Transcript::new(label).
// digits: dim_nd values < dim_np (the witness m is their multiplicity vector).  rng: (19 + 2 dim_nd + 1) x 64 bytes.
// out: c_l c_r c_o c_s | r[rounds] | x[rounds] | l | n | r   ;  commit33_out = commit_value(x, s)
extern "C" int bppp_reciprocal_prove(int device, size_t dim_nd, size_t dim_np, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn,
                                     const uint8_t *gvec2_64, size_t gn2, const uint8_t *hvec2_64, size_t hn2, const uint8_t *x32, const uint8_t *s32,
                                     const uint32_t *digits, const uint8_t *rng_bytes, size_t rng_len, const uint8_t *label, size_t label_len, uint8_t *out,
                                     size_t out_cap, size_t *rounds_out, size_t *l_len_out, size_t *n_len_out, uint8_t *commit33_out, int32_t *status) {
    if (!out || !rounds_out || !l_len_out || !n_len_out || !status || !commit33_out) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c;
    if ((rc = reciprocal_setup(c, dim_nd, dim_np, g64, gvec64, gn, hvec64, hn, gvec2_64, gn2, hvec2_64, hn2)) != BPPP_OK) return rc;
    *status = ST_TRUE;
    try {
        SV xs, ss;
        if ((rc = load_scalars(xs, x32, 1)) != BPPP_OK || (rc = load_scalars(ss, s32, 1)) != BPPP_OK) { c.release(); return rc; }
        SV dg(dim_nd), m(dim_np, sc_zero());
        for (size_t i = 0; i < dim_nd; i++) {
            if (digits[i] >= dim_np) { c.release(); return fail(BPPP_ERR_ARG, "digit out of range"); }
            dg[i] = sc_from_u64(digits[i]); m[digits[i]] = sc_add(m[digits[i]], sc_one());
        }
        SV hs0(1, ss[0]);
        uint8_t com33[33];
        rc = commit_hg(c, hs0, SV(), xs[0], com33);                                   // commit_value
        Merlin t; merlin_init(t, label, (uint32_t)label_len);
        merlin_append(t, BPPP_LBL("reciprocal_commitment"), com33, 33);
        Sc e; challenge(t, BPPP_LBL("reciprocal_challenge"), e);
        SV r(dim_nd);
        for (size_t i = 0; i < dim_nd; i++) r[i] = inv_or_panic(sc_add(dg[i], e));     // reciprocal.rs:117-119
        ByteRng rng{rng_bytes, rng_len};
        Sc r_blind = rng.draw();
        SV hs(c.hn, sc_zero()); hs[0] = r_blind;
        for (size_t i = 0; i < dim_nd; i++) hs[9 + i] = r[i];
        uint8_t rcom33[33], ccom33[33];
        if (rc == BPPP_OK) rc = commit_hg(c, hs, SV(), sc_zero(), rcom33);             // commit_poles, reciprocal.rs:93-95
        make_reciprocal_circuit(c, dim_nd, dim_np, e);
        SV v = concat(SV{xs[0]}, r);
        Sc s_tot = sc_add(ss[0], r_blind);
        SV hs2(c.hn, sc_zero()); hs2[0] = s_tot;
        for (size_t i = 1; i < v.size(); i++) hs2[8 + i] = v[i];
        if (rc == BPPP_OK) rc = commit_hg(c, hs2, SV(), v[0], ccom33);                 // circuit.commit, reciprocal.rs:141
        CircuitProofHost proof;
        std::vector<std::vector<uint8_t>> v33{std::vector<uint8_t>(ccom33, ccom33 + 33)};
        if (rc == BPPP_OK) rc = circuit_prove(c, v33, std::vector<SV>{v}, SV{s_tot}, dg, r, m, t, rng, proof);
        if (rc == BPPP_OK) {
            size_t need = 165 + proof.w.r33.size() + proof.w.x33.size() + proof.w.l32.size() + proof.w.n32.size();
            if (need > out_cap) rc = fail(BPPP_ERR_ARG, "output buffer too small");
            else {
                size_t off = write_circuit_record(out, proof);
                memcpy(out + off, rcom33, 33);
                *rounds_out = proof.w.r33.size() / 33; *l_len_out = proof.w.l32.size() / 32; *n_len_out = proof.w.n32.size() / 32;
                memcpy(commit33_out, com33, 33);
            }
        }
    } catch (const Panic &p) { *status = p.code; }
    c.release();
    return rc;
}

// ReciprocalRangeProofProtocol::verify (src/range_proof/reciprocal.rs:98-107)
extern "C" int bppp_reciprocal_verify(int device, size_t dim_nd, size_t dim_np, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn,
                                      const uint8_t *gvec2_64, size_t gn2, const uint8_t *hvec2_64, size_t hn2, const uint8_t *commit33, const uint8_t *rec,
                                      size_t rounds_r, size_t rounds_x, size_t l_len, size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict) {
    if (!rec || !commit33 || !verdict) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c;
    if ((rc = reciprocal_setup(c, dim_nd, dim_np, g64, gvec64, gn, hvec64, hn, gvec2_64, gn2, hvec2_64, hn2)) != BPPP_OK) return rc;
    try {
        Merlin t; merlin_init(t, label, (uint32_t)label_len);
        merlin_append(t, BPPP_LBL("reciprocal_commitment"), commit33, 33);
        Sc e; challenge(t, BPPP_LBL("reciprocal_challenge"), e);
        make_reciprocal_circuit(c, dim_nd, dim_np, e);
        const uint8_t *r33 = rec + 132, *x33 = r33 + 33 * rounds_r, *l32 = x33 + 33 * rounds_x, *n32 = l32 + 32 * l_len, *pr = n32 + 32 * n_len;
        // circuit_commitment = commitment + proof.r  (reciprocal.rs:104)
        uint8_t two_pts[66], vp33[33];
        memcpy(two_pts, commit33, 33); memcpy(two_pts + 33, pr, 33);
        rc = bppp_points_sum(device, two_pts, FMT_COMPRESSED, 2, FMT_COMPRESSED, vp33);
        if (rc != BPPP_OK) { *verdict = ST_BAD_POINT; rc = BPPP_OK; }
        else {
            std::vector<std::vector<uint8_t>> v33{std::vector<uint8_t>(vp33, vp33 + 33)};
            rc = circuit_verify(c, v33, t, rec, rec + 33, rec + 66, rec + 99, r33, rounds_r, x33, rounds_x, l32, l_len, n32, n_len, verdict);
        }
    } catch (const Panic &p) { *verdict = p.code; }
    c.release();
    return rc;
}
