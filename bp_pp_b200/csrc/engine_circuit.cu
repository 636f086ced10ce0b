// libbppp.so, generic arithmetic-circuit / reciprocal translation unit: ArithmeticCircuit::{commit, prove, verify}
// (reference src/circuit.rs:146-653) for arbitrary dimensions, W_m / W_l dense or sparse (CSR) and a tabulated partition function,
// and ReciprocalRangeProofProtocol::{commit_value, commit_poles, prove, verify, make_circuit}
// (src/range_proof/reciprocal.rs:88-214) for arbitrary (dim_nd, dim_np) on top of it.
//
// Division of labour: W_m / W_l live on the device in compressed-sparse-column form with a dictionary of distinct values
// (circuits are mostly 0 / +-1 / a few constants), and the only super-linear step of the protocol -- the products
// lambda_vec x W_l and mu_vec x W_m behind collect_c (circuit.rs:584-653, util.rs:134-142) -- is a kernel (k_spmv_cols);
// every elliptic-curve operation (all commitments, the verifier's recombination, the whole WNLA) runs on the GPU through
// engine_msm.cu / engine_wnla.cu.  What remains on the host is the Merlin transcript and O(dim) vector algebra over F_n
// (the same sc.cuh code the kernels run).  The batched u64 fast path (engine_prove.cu / engine_verify.cu) is the specialisation of this
// file with closed-form coefficients and on-device transcripts.
#define BPPP_FE_NOINLINE 1
#define BPPP_GENERIC_ALLOC 1   // engine_generic.cuh: cudaMalloc / cudaFree of this file go through the caching allocator
#include <algorithm>
#include <map>
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

// W_m / W_l: compressed sparse columns, values through a dictionary (id 0 is reserved for the scalar one: no multiplication).
struct Sparse {
    size_t rows = 0, cols = 0;
    std::vector<uint32_t> colptr, rowidx, validx;
    SV table;                                   // distinct values, table[0] = 1
    uint32_t *d_colptr = nullptr, *d_rowidx = nullptr, *d_validx = nullptr, *d_table = nullptr;
    size_t nnz() const { return rowidx.size(); }
    void release() { cudaFree(d_colptr); cudaFree(d_rowidx); cudaFree(d_validx); cudaFree(d_table); d_colptr = d_rowidx = d_validx = d_table = nullptr; }
};
struct ScKey { uint32_t v[8]; bool operator<(const ScKey &o) const { return memcmp(v, o.v, 32) < 0; } };
struct ValueDict {
    Sparse &m; std::map<ScKey, uint32_t> ids;
    explicit ValueDict(Sparse &mm) : m(mm) { m.table.assign(1, sc_one()); ScKey k; memcpy(k.v, m.table[0].v, 32); ids[k] = 0; }
    uint32_t id(const Sc &s) {
        ScKey k; memcpy(k.v, s.v, 32);
        auto it = ids.find(k);
        if (it != ids.end()) return it->second;
        uint32_t n = (uint32_t)m.table.size(); m.table.push_back(s); ids[k] = n; return n;
    }
};
// entries as (row, col, value id) triples -> CSC by a counting sort over the columns
void sparse_from_triples(Sparse &m, size_t rows, size_t cols, const std::vector<uint32_t> &tr, const std::vector<uint32_t> &tc, const std::vector<uint32_t> &tv) {
    m.rows = rows; m.cols = cols;
    m.colptr.assign(cols + 1, 0);
    for (uint32_t c : tc) m.colptr[c + 1]++;
    for (size_t j = 0; j < cols; j++) m.colptr[j + 1] += m.colptr[j];
    m.rowidx.resize(tr.size()); m.validx.resize(tr.size());
    std::vector<uint32_t> cur(m.colptr.begin(), m.colptr.end() - 1);
    for (size_t k = 0; k < tr.size(); k++) { uint32_t p = cur[tc[k]]++; m.rowidx[p] = tr[k]; m.validx[p] = tv[k]; }
}
int sparse_upload(Sparse &m) {
    auto up = [](const void *h, size_t bytes, uint32_t **d) -> int {
        CUDA_OK(cudaMalloc(d, bytes ? bytes : 4));
        if (bytes) CUDA_OK(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
        return BPPP_OK;
    };
    int rc = up(m.colptr.data(), 4 * m.colptr.size(), &m.d_colptr);
    if (rc == BPPP_OK) rc = up(m.rowidx.data(), 4 * m.rowidx.size(), &m.d_rowidx);
    if (rc == BPPP_OK) rc = up(m.validx.data(), 4 * m.validx.size(), &m.d_validx);
    if (rc == BPPP_OK) rc = up(m.table.data(), 32 * m.table.size(), &m.d_table);
    return rc;
}
}  // namespace

// vector_mul_on_matrix (util.rs:134-142) for a sparse matrix: out[j] = sum_i a[i] m[i][j], rows beyond |a| contribute nothing
// (the reference zero-extends).  One warp per column: lanes stride over the column's entries, shuffle tree at the end.
__global__ void __launch_bounds__(128) k_spmv_cols(const uint32_t *colptr, const uint32_t *rowidx, const uint32_t *validx, const uint32_t *table, const uint32_t *a, uint32_t na,
                                                   uint32_t cols, uint32_t *out) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (j >= cols) return;
    Sc acc = sc_zero();
    for (uint32_t k = colptr[j] + lane; k < colptr[j + 1]; k += 32) {
        const uint32_t r = rowidx[k], vi = validx[k];
        if (r >= na) continue;
        Sc av, tv;
#pragma unroll
        for (int q = 0; q < 8; q++) av.v[q] = a[8 * (size_t)r + q];
        if (vi) {
#pragma unroll
            for (int q = 0; q < 8; q++) tv.v[q] = table[8 * (size_t)vi + q];
            av = sc_mul(av, tv);
        }
        acc = sc_add(acc, av);
    }
    for (int off = 16; off >= 1; off >>= 1) {
        Sc o;
#pragma unroll
        for (int q = 0; q < 8; q++) o.v[q] = __shfl_xor_sync(0xFFFFFFFFu, acc.v[q], off);
        acc = sc_add(acc, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) out[8 * (size_t)j + q] = acc.v[q];
    }
}

namespace {
// a x m on the device
int spmv(const SV &a, const Sparse &m, SV &out) {
    out.assign(m.cols, sc_zero());
    if (!m.cols) return BPPP_OK;
    uint32_t *d_a = nullptr, *d_o = nullptr;
    CUDA_OK(cudaMalloc(&d_a, 32 * std::max<size_t>(a.size(), 1))); CUDA_OK(cudaMalloc(&d_o, 32 * m.cols));
    if (!a.empty()) CUDA_OK(cudaMemcpy(d_a, a.data(), 32 * a.size(), cudaMemcpyHostToDevice));
    k_spmv_cols<<<(unsigned)((m.cols * 32 + 127) / 128), 128>>>(m.d_colptr, m.d_rowidx, m.d_validx, m.d_table, d_a, (uint32_t)a.size(), (uint32_t)m.cols, d_o);
    CUDA_OK(cudaMemcpy(out.data(), d_o, 32 * m.cols, cudaMemcpyDeviceToHost));
    cudaFree(d_a); cudaFree(d_o);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

struct Panic { int32_t code; };
Sc inv_or_panic(const Sc &a) { if (sc_is_zero(a)) throw Panic{ST_PANIC_INVERT_ZERO}; return sc_inv(a); }

struct Circuit {
    size_t dim_nm = 0, dim_no = 0, k = 0, dim_nl = 0, dim_nv = 0, dim_nw = 0;
    bool f_l = false, f_m = false;
    Sparse W_m, W_l; SV a_m, a_l;
    std::vector<int32_t> part[4]; // LO, LL, LR, NO
    std::vector<uint8_t> g64, gvec64, hvec64, gvec2_64, hvec2_64;
    // device: [h_vec | g_vec | g]
    uint32_t *d_pts = nullptr; size_t hn = 0, gn = 0;
    int part_get(int typ, size_t j) const { return j < part[typ].size() ? part[typ][j] : -1; }
    void release() { cudaFree(d_pts); d_pts = nullptr; W_m.release(); W_l.release(); }
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
struct Coefs { SV nL, nR, nO, lL, lR, lO; };
// circuit.rs:584-653.  u = lambda_vec x W_l and v = mu_vec x W_m over ALL dim_nw columns (two sparse products on the device);
// the twelve blocks the reference slices out of W (W_lL, W_lR, W_lO, W_mL, W_mR, W_mO and the partition-mapped copies of the
// O blocks, circuit.rs:627-653) are then index ranges / gathers of u and v.
Coefs collect_c(const Circuit &c, const SV &lambda_vec, const SV &mu_vec, const Sc &mu) {
    size_t nm = c.dim_nm;
    SV u, v;
    if (spmv(lambda_vec, c.W_l, u) != BPPP_OK || spmv(mu_vec, c.W_m, v) != BPPP_OK) throw Panic{ST_BAD_ARG};
    SV d = vsub(u, v);                                     // (lambda_vec W_l - mu_vec W_m), dim_nw entries
    // diag_inv(mu, nm) (util.rs:118-132) applied as a diagonal scaling
    Sc mu_inv = inv_or_panic(mu);
    SV dinv(nm); Sc val = sc_one();
    for (size_t i = 0; i < nm; i++) { val = sc_mul(val, mu_inv); dinv[i] = val; }
    auto mapped = [&](int typ, size_t jsz) {               // column j of the mapped O block is column partition(typ, j) of W_xO
        SV r(jsz, sc_zero());
        for (size_t j = 0; j < jsz; j++) { int j_ = c.part_get(typ, j); if (j_ >= 0) r[j] = d[2 * nm + (size_t)j_]; }
        return r;
    };
    Coefs r;
    r.nL.resize(nm); r.nR.resize(nm);
    for (size_t j = 0; j < nm; j++) { r.nL[j] = sc_mul(d[j], dinv[j]); r.nR[j] = sc_mul(d[nm + j], dinv[j]); }
    r.nO = mapped(P_NO, nm);
    for (size_t j = 0; j < nm; j++) r.nO[j] = sc_mul(r.nO[j], dinv[j]);
    r.lL = mapped(P_LL, c.dim_nv); r.lR = mapped(P_LR, c.dim_nv); r.lO = mapped(P_LO, c.dim_nv);
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
        if (rc != BPPP_OK) { cudaFree(d_pt30); cudaFree(d_com30); if (rc != BPPP_ERR_ENCODING) return rc; *verdict = ST_BAD_POINT; return BPPP_OK; }
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

int load_scalars(SV &out, const uint8_t *b, size_t n) { out.resize(n); for (size_t i = 0; i < n; i++) if (!sc_from_be32(out[i], b + 32 * i)) return fail(BPPP_ERR_ENCODING, "a scalar is not canonical (>= n)"); return BPPP_OK; }

int pick_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BPPP_ERR_ARG, "bad device index");
    CUDA_OK(cudaSetDevice(device));
    return BPPP_OK;
}

// the descriptor checks the reference enforces by panicking (index out of bounds): reject instead of reading past a buffer
int circuit_check_shape(size_t dim_nm, size_t dim_no, size_t k, size_t dim_nv, size_t gn, size_t hn, const int32_t *const parts[4], size_t part_n) {
    if (dim_nv < 1 || k < 1) return fail(BPPP_ERR_ARG, "circuit: dim_nv and k must be at least 1");
    if (gn < dim_nm) return fail(BPPP_ERR_ARG, "circuit: g_vec needs dim_nm points (circuit.rs:148-150 indexes it)");
    if (hn < dim_nv + 9) return fail(BPPP_ERR_ARG, "circuit: h_vec needs dim_nv + 9 points (circuit.rs:148-150)");
    (void)dim_nm;      // indices at or beyond part_n read as None (documented in bppp.h): a short table is a sparse partition, not an error
    for (int t = 0; t < 4; t++) {
        if (part_n && !parts[t]) return fail(BPPP_ERR_ARG, "circuit: null partition table");
        for (size_t j = 0; j < part_n; j++)
            if (parts[t][j] < -1 || (parts[t][j] >= 0 && (size_t)parts[t][j] >= dim_no)) return fail(BPPP_ERR_ARG, "circuit: partition entry outside [-1, dim_no) (the reference would index w_o out of bounds)");
    }
    return BPPP_OK;
}
int circuit_common(Circuit &c, size_t dim_nm, size_t dim_no, size_t k, size_t dim_nv, int f_l, int f_m, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64,
                   size_t hn, const uint8_t *gvec2_64, size_t gn2, const uint8_t *hvec2_64, size_t hn2, const uint8_t *a_m32, const uint8_t *a_l32,
                   const int32_t *const parts[4], size_t part_n) {
    if (!g64 || (gn && !gvec64) || (hn && !hvec64) || (gn2 && !gvec2_64) || (hn2 && !hvec2_64) || !a_m32 || !a_l32) return fail(BPPP_ERR_ARG, "circuit: null pointer with a non-zero size");
    int rc = circuit_check_shape(dim_nm, dim_no, k, dim_nv, gn, hn, parts, part_n);
    if (rc != BPPP_OK) return rc;
    c.dim_nm = dim_nm; c.dim_no = dim_no; c.k = k; c.dim_nv = dim_nv; c.dim_nl = dim_nv * k; c.dim_nw = 2 * dim_nm + dim_no;
    c.f_l = f_l != 0; c.f_m = f_m != 0;
    c.g64.assign(g64, g64 + 64);
    c.gvec64.assign(gvec64, gvec64 + 64 * gn); c.hvec64.assign(hvec64, hvec64 + 64 * hn);
    c.gvec2_64.assign(gvec2_64, gvec2_64 + 64 * gn2); c.hvec2_64.assign(hvec2_64, hvec2_64 + 64 * hn2);
    if ((rc = load_scalars(c.a_m, a_m32, c.dim_nm)) != BPPP_OK) return rc;
    if ((rc = load_scalars(c.a_l, a_l32, c.dim_nl)) != BPPP_OK) return rc;
    for (int t = 0; t < 4; t++) c.part[t].assign(parts[t], parts[t] + part_n);
    return BPPP_OK;
}
int sparse_from_dense(Sparse &m, const uint8_t *W32, size_t rows, size_t cols) {
    if (rows * cols && !W32) return fail(BPPP_ERR_ARG, "circuit: null matrix");
    ValueDict dict(m);
    std::vector<uint32_t> tr, tc, tv;
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < cols; j++) {
            Sc e;
            if (!sc_from_be32(e, W32 + 32 * (i * cols + j))) return fail(BPPP_ERR_ENCODING, "a scalar is not canonical (>= n)");
            if (sc_is_zero(e)) continue;
            tr.push_back((uint32_t)i); tc.push_back((uint32_t)j); tv.push_back(dict.id(e));
        }
    sparse_from_triples(m, rows, cols, tr, tc, tv);
    return sparse_upload(m);
}
int sparse_from_csr(Sparse &m, const bppp_sparse_matrix *d, size_t rows, size_t cols) {
    if (!d || d->rows != rows || d->cols != cols || !d->row_ptr || (d->nnz && (!d->col_idx || !d->values32))) return fail(BPPP_ERR_ARG, "circuit: sparse matrix shape / pointers");
    if (d->row_ptr[0] != 0 || d->row_ptr[rows] != d->nnz || d->nnz >= 0xFFFFFFFFull) return fail(BPPP_ERR_ARG, "circuit: sparse matrix row_ptr does not span nnz");
    const size_t nvals = d->value_idx ? d->n_values : d->nnz;
    SV vals(nvals);
    for (size_t k = 0; k < nvals; k++) if (!sc_from_be32(vals[k], d->values32 + 32 * k)) return fail(BPPP_ERR_ENCODING, "a scalar is not canonical (>= n)");
    ValueDict dict(m);
    std::vector<uint32_t> ids(nvals);
    for (size_t k = 0; k < nvals; k++) ids[k] = dict.id(vals[k]);
    std::vector<uint32_t> tr, tc, tv;
    tr.reserve(d->nnz); tc.reserve(d->nnz); tv.reserve(d->nnz);
    for (size_t i = 0; i < rows; i++) {
        if (d->row_ptr[i + 1] < d->row_ptr[i] || d->row_ptr[i + 1] > d->nnz) return fail(BPPP_ERR_ARG, "circuit: sparse matrix row_ptr not monotone");
        for (uint64_t k = d->row_ptr[i]; k < d->row_ptr[i + 1]; k++) {
            if (d->col_idx[k] >= cols) return fail(BPPP_ERR_ARG, "circuit: sparse matrix column index out of range");
            size_t vi = d->value_idx ? d->value_idx[k] : (size_t)k;
            if (vi >= nvals) return fail(BPPP_ERR_ARG, "circuit: sparse matrix value index out of range");
            if (sc_is_zero(vals[vi])) continue;
            tr.push_back((uint32_t)i); tc.push_back(d->col_idx[k]); tv.push_back(ids[vi]);
        }
    }
    sparse_from_triples(m, rows, cols, tr, tc, tv);
    return sparse_upload(m);
}
int circuit_from_desc(Circuit &c, const bppp_circuit_desc *d) {
    const int32_t *parts[4] = {d->part_lo, d->part_ll, d->part_lr, d->part_no};
    int rc = circuit_common(c, d->dim_nm, d->dim_no, d->k, d->dim_nv, d->f_l, d->f_m, d->g64, d->gvec64, d->gn, d->hvec64, d->hn, d->gvec2_64, d->gn2, d->hvec2_64, d->hn2,
                            d->a_m32, d->a_l32, parts, d->part_n);
    if (rc != BPPP_OK) return rc;
    if ((rc = sparse_from_dense(c.W_m, d->W_m32, c.dim_nm, c.dim_nw)) != BPPP_OK) return rc;
    if ((rc = sparse_from_dense(c.W_l, d->W_l32, c.dim_nl, c.dim_nw)) != BPPP_OK) return rc;
    return circuit_upload(c);
}
int circuit_from_sparse_desc(Circuit &c, const bppp_circuit_desc_sparse *d) {
    const int32_t *parts[4] = {d->part_lo, d->part_ll, d->part_lr, d->part_no};
    int rc = circuit_common(c, d->dim_nm, d->dim_no, d->k, d->dim_nv, d->f_l, d->f_m, d->g64, d->gvec64, d->gn, d->hvec64, d->hn, d->gvec2_64, d->gn2, d->hvec2_64, d->hn2,
                            d->a_m32, d->a_l32, parts, d->part_n);
    if (rc != BPPP_OK) return rc;
    if ((rc = sparse_from_csr(c.W_m, &d->W_m, c.dim_nm, c.dim_nw)) != BPPP_OK) return rc;
    if ((rc = sparse_from_csr(c.W_l, &d->W_l, c.dim_nl, c.dim_nw)) != BPPP_OK) return rc;
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

// reciprocal.rs:150-214 built directly in sparse form (the reference materialises dense (nd + 1) x (2 nd + np) matrices and
// recomputes the np distinct inverses per row): W_m has nd entries, W_l nd + nd (nd - 1) ones + nd np inverses
int make_reciprocal_circuit(Circuit &c, size_t nd, size_t np, const Sc &e) {
    c.dim_nm = nd; c.dim_no = np; c.k = 1; c.dim_nv = nd + 1; c.dim_nl = nd + 1; c.dim_nw = 2 * nd + np;
    c.f_l = true; c.f_m = false;
    c.a_m.assign(nd, sc_one()); c.a_l.assign(nd + 1, sc_zero());
    c.W_m.release(); c.W_l.release();
    c.W_m = Sparse(); c.W_l = Sparse();
    std::vector<uint32_t> tr, tc, tv;
    {
        ValueDict dict(c.W_m);
        uint32_t me = dict.id(sc_neg(e));
        for (size_t i = 0; i < nd; i++) { tr.push_back((uint32_t)i); tc.push_back((uint32_t)(i + nd)); tv.push_back(me); }       // -e on w_R (reciprocal.rs:166-169)
        sparse_from_triples(c.W_m, nd, c.dim_nw, tr, tc, tv);
    }
    tr.clear(); tc.clear(); tv.clear();
    {
        ValueDict dict(c.W_l);
        Sc base = sc_from_u64((uint64_t)(uint32_t)np), pw = sc_one();
        for (size_t i = 0; i < nd; i++) { tr.push_back(0); tc.push_back((uint32_t)i); tv.push_back(dict.id(sc_neg(pw))); pw = sc_mul(pw, base); }    // :172-176
        std::vector<uint32_t> inv(np);
        for (size_t j = 0; j < np; j++) inv[j] = dict.id(sc_neg(inv_or_panic(sc_add(e, sc_from_u64((uint64_t)(uint32_t)j)))));
        tr.reserve(nd * (nd + np)); tc.reserve(nd * (nd + np)); tv.reserve(nd * (nd + np));
        for (size_t i = 0; i < nd; i++) {
            for (size_t j = 0; j < nd; j++) if (j != i) { tr.push_back((uint32_t)(i + 1)); tc.push_back((uint32_t)(j + nd)); tv.push_back(0); }       // ones, :178-180
            for (size_t j = 0; j < np; j++) { tr.push_back((uint32_t)(i + 1)); tc.push_back((uint32_t)(j + 2 * nd)); tv.push_back(inv[j]); }           // :181-183
        }
        sparse_from_triples(c.W_l, nd + 1, c.dim_nw, tr, tc, tv);
    }
    size_t pn = nd + 1;
    for (int t = 0; t < 4; t++) { c.part[t].assign(pn, -1); }
    for (size_t j = 0; j < pn && j < np; j++) c.part[P_LL][j] = (int32_t)j;
    int rc = sparse_upload(c.W_m);
    if (rc == BPPP_OK) rc = sparse_upload(c.W_l);
    return rc;
}

}  // namespace

static int load_desc(Circuit &c, const bppp_circuit_desc *d) { return circuit_from_desc(c, d); }
static int load_desc(Circuit &c, const bppp_circuit_desc_sparse *d) { return circuit_from_sparse_desc(c, d); }

// ArithmeticCircuit::commit (src/circuit.rs:146-151)
template <class Desc>
static int circuit_commit_any(int device, const Desc *d, const uint8_t *v32, const uint8_t *s32, uint8_t *out33) {
    if (!d || !v32 || !s32 || !out33) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c; if ((rc = load_desc(c, d)) != BPPP_OK) { c.release(); return rc; }
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

extern "C" int bppp_circuit_commit(int device, const bppp_circuit_desc *d, const uint8_t *v32, const uint8_t *s32, uint8_t *out33) { return circuit_commit_any(device, d, v32, s32, out33); }
extern "C" int bppp_circuit_commit_sparse(int device, const bppp_circuit_desc_sparse *d, const uint8_t *v32, const uint8_t *s32, uint8_t *out33) { return circuit_commit_any(device, d, v32, s32, out33); }

// ArithmeticCircuit::prove (src/circuit.rs:260-556), fresh Transcript::new(label).  out: c_l c_r c_o c_s | r | x | l | n
template <class Desc>
static int circuit_prove_any(int device, const Desc *d, const uint8_t *commits33, const uint8_t *v32, const uint8_t *sv32, const uint8_t *wl32,
                                  const uint8_t *wr32, const uint8_t *wo32, const uint8_t *rng_bytes, size_t rng_len, const uint8_t *label, size_t label_len,
                                  uint8_t *out, size_t out_cap, size_t *rounds_out, size_t *l_len_out, size_t *n_len_out, int32_t *status) {
    if (!d || !out || !rounds_out || !l_len_out || !n_len_out || !status) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c; if ((rc = load_desc(c, d)) != BPPP_OK) { c.release(); return rc; }
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

#define BPPP_PROVE_ARGS const uint8_t *commits33, const uint8_t *v32, const uint8_t *sv32, const uint8_t *wl32, const uint8_t *wr32, const uint8_t *wo32, \
                        const uint8_t *rng_bytes, size_t rng_len, const uint8_t *label, size_t label_len, uint8_t *out, size_t out_cap, size_t *rounds_out,   \
                        size_t *l_len_out, size_t *n_len_out, int32_t *status
#define BPPP_PROVE_PASS commits33, v32, sv32, wl32, wr32, wo32, rng_bytes, rng_len, label, label_len, out, out_cap, rounds_out, l_len_out, n_len_out, status
extern "C" int bppp_circuit_prove(int device, const bppp_circuit_desc *d, BPPP_PROVE_ARGS) { return circuit_prove_any(device, d, BPPP_PROVE_PASS); }
extern "C" int bppp_circuit_prove_sparse(int device, const bppp_circuit_desc_sparse *d, BPPP_PROVE_ARGS) { return circuit_prove_any(device, d, BPPP_PROVE_PASS); }

// ArithmeticCircuit::verify (src/circuit.rs:154-256)
template <class Desc>
static int circuit_verify_any(int device, const Desc *d, const uint8_t *commits33, const uint8_t *rec, size_t rounds_r, size_t rounds_x, size_t l_len,
                              size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict) {
    if (!d || !rec || !verdict) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    Circuit c; if ((rc = load_desc(c, d)) != BPPP_OK) { c.release(); return rc; }
    std::vector<std::vector<uint8_t>> v33(d->k);
    for (size_t i = 0; i < d->k; i++) v33[i].assign(commits33 + 33 * i, commits33 + 33 * (i + 1));
    Merlin t; merlin_init(t, label, (uint32_t)label_len);
    const uint8_t *r33 = rec + 132, *x33 = r33 + 33 * rounds_r, *l32 = x33 + 33 * rounds_x, *n32 = l32 + 32 * l_len;
    try { rc = circuit_verify(c, v33, t, rec, rec + 33, rec + 66, rec + 99, r33, rounds_r, x33, rounds_x, l32, l_len, n32, n_len, verdict); }
    catch (const Panic &p) { *verdict = p.code; }
    c.release();
    return rc;
}
extern "C" int bppp_circuit_verify(int device, const bppp_circuit_desc *d, const uint8_t *commits33, const uint8_t *rec, size_t rounds_r, size_t rounds_x, size_t l_len,
                                   size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict) {
    return circuit_verify_any(device, d, commits33, rec, rounds_r, rounds_x, l_len, n_len, label, label_len, verdict);
}
extern "C" int bppp_circuit_verify_sparse(int device, const bppp_circuit_desc_sparse *d, const uint8_t *commits33, const uint8_t *rec, size_t rounds_r, size_t rounds_x,
                                          size_t l_len, size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict) {
    return circuit_verify_any(device, d, commits33, rec, rounds_r, rounds_x, l_len, n_len, label, label_len, verdict);
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
        if (rc == BPPP_OK) rc = make_reciprocal_circuit(c, dim_nd, dim_np, e);
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
        rc = make_reciprocal_circuit(c, dim_nd, dim_np, e);
        if (rc != BPPP_OK) { c.release(); return rc; }
        const uint8_t *r33 = rec + 132, *x33 = r33 + 33 * rounds_r, *l32 = x33 + 33 * rounds_x, *n32 = l32 + 32 * l_len, *pr = n32 + 32 * n_len;
        // circuit_commitment = commitment + proof.r  (reciprocal.rs:104)
        uint8_t two_pts[66], vp33[33];
        memcpy(two_pts, commit33, 33); memcpy(two_pts + 33, pr, 33);
        rc = bppp_points_sum(device, two_pts, FMT_COMPRESSED, 2, FMT_COMPRESSED, vp33);
        if (rc == BPPP_ERR_ENCODING) { *verdict = ST_BAD_POINT; rc = BPPP_OK; }
        else if (rc != BPPP_OK) { c.release(); return rc; }
        else {
            std::vector<std::vector<uint8_t>> v33{std::vector<uint8_t>(vp33, vp33 + 33)};
            rc = circuit_verify(c, v33, t, rec, rec + 33, rec + 66, rec + 99, r33, rounds_r, x33, rounds_x, l32, l_len, n32, n_len, verdict);
        }
    } catch (const Panic &p) { *verdict = p.code; }
    c.release();
    return rc;
}
