// libbppp.so, generic weight-norm-linear-argument translation unit: WeightNormLinearArgument::{commit, prove,
// verify} (reference src/wnla.rs:66-190) for arbitrary vector lengths on one GPU.
//
// All curve work runs on the device: X/R/commit are Pippenger MSMs (engine_msm.cu) over the current generators,
// generator folding h' = h0 + y h1, g' = rho g0 + y g1 (wnla.rs:170-171) is one GLV ladder per output point, scalar
// folding and the weighted inner products are element-wise / block-reduction kernels.  The host only drives the
// Merlin transcript (33-byte points in, 32-byte challenge out per round) -- SURVEY 7.1 step 8.
//
// Length semantics: the reference zero-extends mismatched vectors everywhere (util.rs:24-26) and halves every
// vector independently with ceil (util.rs:7-22).  Padding h, l, c to a common length with identity / zero and
// g, n likewise is exactly equivalent, provided the TRUE lengths are tracked for the transcript (l.sz, n.sz) and
// for the termination test (wnla.rs:126); that is what this file does.
//
// The verifier never folds generators: after R rounds the base-case commit(l, n) over the folded generators
// (wnla.rs:80-82) is ONE MSM over the original generators with scalars
//   H_i : l[i >> R] * prod_k y_k^bit_k(i)        G_i : n[i >> R] * prod_k (bit_k(i) ? y_k : rho_k)
// (rho_0 = rho, rho_{k+1} = mu_k, mu_{k+1} = mu_k^2), bit-identical to the round-by-round evaluation.
#define BPPP_FE_NOINLINE 1
#define BPPP_GENERIC_ALLOC 1   // engine_generic.cuh: cudaMalloc / cudaFree of this file go through the caching allocator
#include <algorithm>
#include <functional>
#include "engine_generic.cuh"

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

namespace bppp {

static std::atomic<uint64_t> g_wnla_launches{0};
#define WL(kern, grid, block, ...) do { kern<<<(grid), (block), 0, st>>>(__VA_ARGS__); g_wnla_launches++; } while (0)

__device__ __forceinline__ Sc ld_sc8(const uint32_t *p) { Sc r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = p[k]; return r; }
__device__ __forceinline__ void st_sc8(uint32_t *p, const Sc &a) {
#pragma unroll
    for (int k = 0; k < 8; k++) p[k] = a.v[k]; }
__device__ __forceinline__ Pt ld_pt30g(const uint32_t *p) { Pt r;
#pragma unroll
    for (int k = 0; k < FE_W; k++) { r.x.v[k] = p[k]; r.y.v[k] = p[FE_W + k]; r.z.v[k] = p[2 * FE_W + k]; }
    return r; }
__device__ __forceinline__ void st_pt30g(uint32_t *p, const Pt &a) {
#pragma unroll
    for (int k = 0; k < FE_W; k++) { p[k] = a.x.v[k]; p[FE_W + k] = a.y.v[k]; p[2 * FE_W + k] = a.z.v[k]; } }
__device__ __forceinline__ bool ld_pta16(PtA &q, const uint32_t *pts, size_t idx) {
    uint32_t x[8], y[8], any = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = pts[16 * idx + k]; y[k] = pts[16 * idx + 8 + k]; any |= x[k] | y[k]; }
    q.x = fe_from_words(x); q.y = fe_from_words(y);
    return any != 0;
}
__device__ __forceinline__ void st_pta16(uint32_t *pts, size_t idx, const Pt &p) {
    bool id = pt_is_identity(p);
    PtA a = pt_to_affine_with_zinv(p, fe_inv(p.z));
    uint32_t x[8], y[8];
    fe_to_words(x, a.x); fe_to_words(y, a.y);
#pragma unroll
    for (int k = 0; k < 8; k++) { pts[16 * idx + k] = id ? 0u : x[k]; pts[16 * idx + 8 + k] = id ? 0u : y[k]; }
}
__device__ __forceinline__ Sc sc_pow_u64_dev(Sc base, uint64_t e) {
    Sc acc = sc_one();
#pragma unroll 1
    while (e) { if (e & 1) acc = sc_mul(acc, base); base = sc_sqr(base); e >>= 1; }
    return acc;
}

struct ScParam { uint32_t v[8]; };
static ScParam to_param(const Sc &s) { ScParam p; for (int k = 0; k < 8; k++) p.v[k] = s.v[k]; return p; }
__device__ __forceinline__ Sc from_param(const ScParam &p) { Sc s;
#pragma unroll
    for (int k = 0; k < 8; k++) s.v[k] = p.v[k]; return s; }

// MSM scalars of X and R over [H (Lh) | G (Lg) | g] (wnla.rs:152-160): slots Lh + Lg hold vx / vr, filled by the host
// The G points are stored divided by sigma (WnlaDev): rho_p / rho_inv_p arrive multiplied by sigma, sigma_p scales R's G part.
__global__ void k_wnla_xr_scalars(const uint32_t *l, const uint32_t *n, size_t Lh, size_t Lg, ScParam rho_p, ScParam rho_inv_p, ScParam sigma_p, uint32_t *sx, uint32_t *sr) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Lh) {
        size_t o = t ^ 1;
        Sc lx = o < Lh ? ld_sc8(l + 8 * o) : sc_zero();                 // h0_i pairs with l1_i, h1_i with l0_i
        st_sc8(sx + 8 * t, lx);
        st_sc8(sr + 8 * t, (t & 1) ? ld_sc8(l + 8 * t) : sc_zero());    // <h1, l1>
    } else if (t < Lh + Lg) {
        size_t m = t - Lh, o = m ^ 1;
        Sc nx = o < Lg ? ld_sc8(n + 8 * o) : sc_zero();
        nx = sc_mul(nx, (m & 1) ? from_param(rho_inv_p) : from_param(rho_p));   // <g0, rho n1> + <g1, rho^-1 n0>
        st_sc8(sx + 8 * t, nx);
        st_sc8(sr + 8 * t, (m & 1) ? sc_mul(ld_sc8(n + 8 * m), from_param(sigma_p)) : sc_zero());    // <g1, n1>
    }
}
// per-block partial sums of
//   out[0]: sum_i n[2i] n[2i+1] mu2^(i+1)   out[1]: sum_i c[2i] l[2i+1] + c[2i+1] l[2i]
//   out[2]: sum_i n[2i+1]^2 mu2^(i+1)       out[3]: sum_i c[2i+1] l[2i+1]
// goff: global index of this block's first (halved) n pair -- 0 for a whole instance, the block offset for a shard
__global__ void __launch_bounds__(128) k_wnla_dots(const uint32_t *c, const uint32_t *l, const uint32_t *n, size_t Lh, size_t Lg, ScParam mu2_p, uint32_t *partials, size_t goff) {
    __shared__ uint32_t sh[4][128][8];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Sc a0 = sc_zero(), a1 = sc_zero(), a2 = sc_zero(), a3 = sc_zero();
    if (2 * i + 1 < Lg) {
        Sc n0 = ld_sc8(n + 8 * (2 * i)), n1 = ld_sc8(n + 8 * (2 * i + 1));
        Sc w = sc_pow_u64_dev(from_param(mu2_p), (uint64_t)(goff + i) + 1);
        Sc n1w = sc_mul(n1, w);
        a0 = sc_mul(n0, n1w); a2 = sc_mul(n1, n1w);
    }
    if (2 * i + 1 < Lh) {
        Sc c0 = ld_sc8(c + 8 * (2 * i)), c1 = ld_sc8(c + 8 * (2 * i + 1)), l0 = ld_sc8(l + 8 * (2 * i)), l1 = ld_sc8(l + 8 * (2 * i + 1));
        a1 = sc_add(sc_mul(c0, l1), sc_mul(c1, l0)); a3 = sc_mul(c1, l1);
    }
    st_sc8(sh[0][threadIdx.x], a0); st_sc8(sh[1][threadIdx.x], a1); st_sc8(sh[2][threadIdx.x], a2); st_sc8(sh[3][threadIdx.x], a3);
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll 1
            for (int q = 0; q < 4; q++) st_sc8(sh[q][threadIdx.x], sc_add(ld_sc8(sh[q][threadIdx.x]), ld_sc8(sh[q][threadIdx.x + s])));
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) st_sc8(partials + 8 * (4 * (size_t)blockIdx.x + threadIdx.x), ld_sc8(sh[threadIdx.x][0]));
}
// generic: per-block partial sums of  sum_i a[i] b[i]  and  sum_i n[i]^2 w^(i+1)   (wnla.commit's v)
__global__ void __launch_bounds__(128) k_commit_dots(const uint32_t *c, const uint32_t *l, size_t Lcl, const uint32_t *n, size_t Ln, ScParam mu_p, uint32_t *partials, size_t goff) {
    __shared__ uint32_t sh[2][128][8];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Sc a0 = sc_zero(), a1 = sc_zero();
    if (i < Lcl) a0 = sc_mul(ld_sc8(c + 8 * i), ld_sc8(l + 8 * i));
    if (i < Ln) { Sc v = ld_sc8(n + 8 * i); a1 = sc_mul(sc_sqr(v), sc_pow_u64_dev(from_param(mu_p), (uint64_t)(goff + i) + 1)); }
    st_sc8(sh[0][threadIdx.x], a0); st_sc8(sh[1][threadIdx.x], a1);
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) {
            st_sc8(sh[0][threadIdx.x], sc_add(ld_sc8(sh[0][threadIdx.x]), ld_sc8(sh[0][threadIdx.x + s])));
            st_sc8(sh[1][threadIdx.x], sc_add(ld_sc8(sh[1][threadIdx.x]), ld_sc8(sh[1][threadIdx.x + s])));
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) st_sc8(partials + 8 * (2 * (size_t)blockIdx.x + threadIdx.x), ld_sc8(sh[threadIdx.x][0]));
}
// generator folding (wnla.rs:170-171): out has ceil(L/2) points, out_i = in_2i + k in_2i+1.  h' = h0 + y h1 is k = y; the G half,
// g' = rho g0 + y g1, is stored as G0 + (y / rho) G1 with the factor rho moved into sigma (WnlaDev), so it is the same kernel with
// k = y / rho: one scalar multiplication per output for both halves.
__global__ void __launch_bounds__(64, 7) k_wnla_fold_points(const uint32_t *in, size_t L, ScParam k_p, uint32_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t Lo = (L + 1) / 2;
    if (i >= Lo) return;
    PtA p0, p1;
    bool ok0 = ld_pta16(p0, in, 2 * i);
    bool ok1 = 2 * i + 1 < L ? ld_pta16(p1, in, 2 * i + 1) : false;
    if (!ok1) { p1.x = fe_zero(); p1.y = fe_zero(); BPPP_SET_MAG(p1.x, 1); BPPP_SET_MAG(p1.y, 1); }
    PtA pts[1] = {p1}; bool ident[1] = {!ok1};
    Sc ks[1] = {from_param(k_p)};
    st_pta16(out, i, straus_var<1>(pts, ident, ks, pt_from_affine(p0, !ok0)));
}
// scalar folding (wnla.rs:172-175): c' = c0 + y c1, l' = l0 + y l1 over Lh; n' = rho^-1 n0 + y n1 over Lg
__global__ void k_wnla_fold_scalars(const uint32_t *c, const uint32_t *l, const uint32_t *n, size_t Lh, size_t Lg, ScParam y_p, ScParam rho_inv_p,
                                    uint32_t *c2, uint32_t *l2, uint32_t *n2, int have_ln) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Sc y = from_param(y_p);
    if (i < (Lh + 1) / 2) {
        bool odd = 2 * i + 1 < Lh;
        Sc c1 = odd ? ld_sc8(c + 8 * (2 * i + 1)) : sc_zero();
        st_sc8(c2 + 8 * i, sc_add(ld_sc8(c + 8 * (2 * i)), sc_mul(y, c1)));
        if (have_ln) {
            Sc l1 = odd ? ld_sc8(l + 8 * (2 * i + 1)) : sc_zero();
            st_sc8(l2 + 8 * i, sc_add(ld_sc8(l + 8 * (2 * i)), sc_mul(y, l1)));
        }
    }
    if (have_ln && i < (Lg + 1) / 2) {
        Sc n1 = 2 * i + 1 < Lg ? ld_sc8(n + 8 * (2 * i + 1)) : sc_zero();
        st_sc8(n2 + 8 * i, sc_add(sc_mul(ld_sc8(n + 8 * (2 * i)), from_param(rho_inv_p)), sc_mul(y, n1)));
    }
}
// s[i] *= k
__global__ void k_sc_scale(uint32_t *s, size_t n, ScParam k_p) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_sc8(s + 8 * i, sc_mul(ld_sc8(s + 8 * i), from_param(k_p)));
}
// pts[i] = k * pts[i] (affine 16-word points, in place): the true G points of a block whose stored points carry sigma
__global__ void __launch_bounds__(64) k_points_scale(uint32_t *pts, size_t n, ScParam k_p) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PtA p; bool ok = ld_pta16(p, pts, i);
    st_pta16(pts, i, pt_mul_glv(pt_from_affine(p, !ok), from_param(k_p)));
}
// com' = com + y X + (y^2 - 1) R   (wnla.rs:100-102; equals wnla'.commit(l', n') of wnla.rs:186)
__global__ void k_wnla_next_commitment(const uint32_t *com30, const uint32_t *x30, const uint32_t *r30, ScParam y_p, uint32_t *out30) {
    // two threads of one warp: the scalar multiplications y X and (y^2 - 1) R run side by side (a single thread is
    // latency-bound on ~130 dependent doublings each), thread 0 adds them to com
    __shared__ uint32_t sh[PT_W];
    const int t = threadIdx.x;
    if (blockIdx.x != 0 || t >= 2) return;
    Sc y = from_param(y_p);
    Sc k = t ? sc_sub(sc_sqr(y), sc_one()) : y;
    Pt part = pt_mul_glv(ld_pt30g(t ? r30 : x30), k);
    if (t == 1) st_pt30g(sh, part);
    __syncwarp(0x3);
    if (t == 0) st_pt30g(out30, pt_add(pt_add(ld_pt30g(com30), part), ld_pt30g(sh)));
}
// verifier: base-case scalars over the ORIGINAL generators (see file header).  ys / rhos: R scalars each.
__global__ void k_wnla_final_scalars(size_t Lh, size_t Lg, int R, const uint32_t *ys, const uint32_t *rhos, const uint32_t *l, size_t ln, const uint32_t *n, size_t nn,
                                     const uint32_t *c, size_t Lc, uint32_t *out_sc, uint32_t *cl_terms) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Lh) {
        Sc coef = sc_one();
#pragma unroll 1
        for (int k = 0; k < R; k++) if ((t >> k) & 1) coef = sc_mul(coef, ld_sc8(ys + 8 * k));
        size_t m = R >= 64 ? 0 : (t >> R);
        Sc lv = m < ln ? ld_sc8(l + 8 * m) : sc_zero();
        Sc s = sc_mul(coef, lv);
        st_sc8(out_sc + 8 * t, s);
        st_sc8(cl_terms + 8 * t, t < Lc ? sc_mul(s, ld_sc8(c + 8 * t)) : sc_zero());     // contribution to <c_final, l>
    } else if (t < Lh + Lg) {
        size_t i = t - Lh;
        Sc coef = sc_one();
#pragma unroll 1
        for (int k = 0; k < R; k++) coef = sc_mul(coef, ((i >> k) & 1) ? ld_sc8(ys + 8 * k) : ld_sc8(rhos + 8 * k));
        size_t m = R >= 64 ? 0 : (i >> R);
        Sc nv = m < nn ? ld_sc8(n + 8 * m) : sc_zero();
        st_sc8(out_sc + 8 * t, sc_mul(coef, nv));
    }
}
__global__ void __launch_bounds__(128) k_sc_sum_partials(const uint32_t *in, size_t L, uint32_t *partials) {
    __shared__ uint32_t sh[128][8];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    st_sc8(sh[threadIdx.x], i < L ? ld_sc8(in + 8 * i) : sc_zero());
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) st_sc8(sh[threadIdx.x], sc_add(ld_sc8(sh[threadIdx.x]), ld_sc8(sh[threadIdx.x + s])));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_sc8(partials + 8 * (size_t)blockIdx.x, ld_sc8(sh[0]));
}
__global__ void k_decode_one_point30(const uint32_t *pts16, uint32_t *out30) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    PtA a; bool ok = ld_pta16(a, pts16, 0);
    st_pt30g(out30, pt_from_affine(a, !ok));
}

static bool sc_is_one_host(const Sc &a) { Sc o = sc_one(); return memcmp(a.v, o.v, 32) == 0; }
// ---- host-side transcript helpers (same Merlin code the device runs) ----
static void host_append_point33(Merlin &m, const char *label, uint32_t ll, const uint8_t *b33) { merlin_append(m, label, ll, b33, 33); }
static bool host_challenge(Merlin &m, const char *label, uint32_t ll, Sc &out) { return merlin_challenge_scalar(m, label, ll, out); }

static int sum_partials_to_host(cudaStream_t st, const uint32_t *d_partials, size_t count, int stride, Sc *out /* stride sums */) {
    std::vector<uint32_t> h(8 * count * stride);
    CUDA_OK(cudaMemcpyAsync(h.data(), d_partials, 32 * count * stride, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int q = 0; q < stride; q++) out[q] = sc_zero();
    for (size_t b = 0; b < count; b++)
        for (int q = 0; q < stride; q++) { Sc v; memcpy(v.v, &h[8 * (b * stride + q)], 32); out[q] = sc_add(out[q], v); }
    return BPPP_OK;
}

int upload_padded_scalars(cudaStream_t st, const uint8_t *h32, size_t n, size_t L, uint32_t **d) {
    std::vector<uint8_t> buf(32 * (L ? L : 1), 0);
    memcpy(buf.data(), h32, 32 * n);
    return decode_scalars_to_device(st, buf.data(), L, d);
}

int wnla_load(cudaStream_t st, WnlaDev &w, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn, const uint8_t *c32, size_t cn,
              const uint8_t *rho32, const uint8_t *mu32, size_t ln, size_t nn) {
    w.len_h = hn; w.len_g = gn;
    w.Lh = std::max(std::max(hn, cn), ln); w.Lg = std::max(gn, nn);
    if (!sc_from_be32(w.rho, rho32) || !sc_from_be32(w.mu, mu32)) return fail(BPPP_ERR_ARG, "rho/mu not canonical");
    std::vector<uint8_t> pb(64 * (w.Lh + w.Lg + 1), 0);
    memcpy(pb.data(), hvec64, 64 * hn);
    memcpy(pb.data() + 64 * w.Lh, gvec64, 64 * gn);
    memcpy(pb.data() + 64 * (w.Lh + w.Lg), g64, 64);
    int rc = decode_points_to_device(st, pb.data(), FMT_AFFINE64, w.Lh + w.Lg + 1, &w.pts);
    if (rc != BPPP_OK) return rc;
    rc = upload_padded_scalars(st, c32, cn, w.Lh, &w.c);
    if (rc != BPPP_OK) { w.release(); return rc; }
    return BPPP_OK;
}

// C = v g + <h, l> + <g_vec, n>, v = <c, l> + |n|^2_mu  (wnla.rs:66-72); l, n device arrays padded to Lh, Lg
int wnla_commit_dev(cudaStream_t st, const WnlaDev &w, const uint32_t *d_l, const uint32_t *d_n, uint32_t *d_out30) {
    size_t L = std::max(w.Lh, w.Lg), nblk = (L + 127) / 128;
    uint32_t *d_part = nullptr, *d_sc = nullptr;
    DevScope scope; scope.own(&d_part); scope.own(&d_sc);
    CUDA_OK(cudaMalloc(&d_part, 64 * (nblk ? nblk : 1)));
    CUDA_OK(cudaMalloc(&d_sc, 32 * (w.Lh + w.Lg + 1)));
    if (nblk) WL(k_commit_dots, (unsigned)nblk, 128, w.c, d_l, w.Lh, d_n, w.Lg, to_param(w.mu), d_part, (size_t)0);
    Sc sums[2];
    int rc = sum_partials_to_host(st, d_part, nblk, 2, sums);
    if (rc != BPPP_OK) return rc;
    Sc v = sc_add(sums[0], sums[1]);
    CUDA_OK(cudaMemcpyAsync(d_sc, d_l, 32 * w.Lh, cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_sc + 8 * w.Lh, d_n, 32 * w.Lg, cudaMemcpyDeviceToDevice, st));
    if (w.scaled && w.Lg) WL(k_sc_scale, nblocks(w.Lg, 128), 128, d_sc + 8 * w.Lh, w.Lg, to_param(w.sigma));
    CUDA_OK(cudaMemcpyAsync(d_sc + 8 * (w.Lh + w.Lg), v.v, 32, cudaMemcpyHostToDevice, st));
    return msm_device(st, w.pts, d_sc, w.Lh + w.Lg + 1, nullptr, d_out30);
}


// wnla.rs:125-190.  d_com30: commitment (projective, device).  d_l / d_n: padded witness arrays (consumed).
// All scratch is allocated once: the scalar / partial-sum buffers at their first-round size and one half-size "B" set the
// folds ping-pong with (round k reads the set round k-1 wrote).  The next commitment C + y X + (y^2 - 1) R (two 128-doubling
// ladders, ~1.4 ms on two threads) runs on a side stream under the round's fold kernels and the next round's MSMs: only the
// next transcript append needs it.
int wnla_prove_dev(cudaStream_t st, WnlaDev &w, Merlin &t, uint32_t *d_com30, uint32_t *d_l, uint32_t *d_n, size_t len_l, size_t len_n, WnlaProofHost &proof,
                   int32_t *status) {
    std::vector<std::vector<uint8_t>> rs, xs;
    const size_t Lh0 = w.Lh, Lg0 = w.Lg, Lt0 = Lh0 + Lg0 + 1;
    const size_t LhB = (Lh0 + 1) / 2, LgB = (Lg0 + 1) / 2;
    uint32_t *d_xr30 = nullptr, *d_three = nullptr, *d_sx = nullptr, *d_sr = nullptr, *d_part = nullptr, *d_com_next = nullptr;
    uint32_t *ptsB = nullptr, *cB = nullptr, *lB = nullptr, *nB = nullptr;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_xr = nullptr, ev_com = nullptr;
    struct Cleanup {
        std::function<void()> f; ~Cleanup() { f(); }
    } cleanup{[&] {
        cudaFree(d_xr30); cudaFree(d_three); cudaFree(d_sx); cudaFree(d_sr); cudaFree(d_part); cudaFree(d_com_next);
        cudaFree(ptsB); cudaFree(cB); cudaFree(lB); cudaFree(nB);
        if (side) cudaStreamDestroy(side);
        if (ev_xr) cudaEventDestroy(ev_xr);
        if (ev_com) cudaEventDestroy(ev_com);
    }};
    // the caller's buffers are set A; d_l / d_n are consumed (freed on every path, like the earlier implementation)
    uint32_t *ptsA = w.pts, *cA = w.c, *lA = d_l, *nA = d_n;
    struct FreeLN { uint32_t *&l, *&n; ~FreeLN() { cudaFree(l); cudaFree(n); } } free_ln{lA, nA};
    CUDA_OK(cudaMalloc(&d_xr30, 2 * PT_BYTES)); CUDA_OK(cudaMalloc(&d_three, 3 * PT_BYTES)); CUDA_OK(cudaMalloc(&d_com_next, PT_BYTES));
    CUDA_OK(cudaMalloc(&d_sx, 32 * Lt0)); CUDA_OK(cudaMalloc(&d_sr, 32 * Lt0));
    {
        size_t half0 = (std::max(Lh0, Lg0) + 1) / 2, nblk0 = (half0 + 127) / 128;
        size_t Lc = std::max(LhB, LgB), nblkc = (Lc + 127) / 128;                 // wnla_commit_dev's partial sums after the first fold
        CUDA_OK(cudaMalloc(&d_part, 128 * std::max<size_t>(std::max(nblk0, nblkc), 1)));
    }
    CUDA_OK(cudaMalloc(&ptsB, 64 * (LhB + LgB + 1))); CUDA_OK(cudaMalloc(&cB, 32 * std::max<size_t>(LhB, 1)));
    CUDA_OK(cudaMalloc(&lB, 32 * std::max<size_t>(LhB, 1))); CUDA_OK(cudaMalloc(&nB, 32 * std::max<size_t>(LgB, 1)));
    CUDA_OK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&ev_xr, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&ev_com, cudaEventDisableTiming));
    uint32_t *pts = ptsA, *cc = cA, *ll = lA, *nn = nA;                  // current set
    uint32_t *pts2 = ptsB, *c2 = cB, *l2 = lB, *n2 = nB;                 // the set the next fold writes
    uint32_t *d_x30 = d_xr30, *d_r30 = d_xr30 + PT_W;
    int rc = BPPP_OK;
    bool first_round = true, com_pending = false;
    while (len_l + len_n >= 6) {     // wnla.rs:126
        size_t Lh = w.Lh, Lg = w.Lg, Lt = Lh + Lg + 1;
        if (sc_is_zero(w.rho)) { *status = ST_PANIC_INVERT_ZERO; break; }     // rho.invert_vartime().unwrap(), wnla.rs:135
        Sc rho_inv = sc_inv(w.rho), mu2 = sc_sqr(w.mu);
        size_t half = (std::max(Lh, Lg) + 1) / 2, nblk = (half + 127) / 128;
        const Sc sigma = w.scaled ? w.sigma : sc_one();
        WL(k_wnla_xr_scalars, nblocks(Lh + Lg, 128), 128, ll, nn, Lh, Lg, to_param(sc_mul(sigma, w.rho)), to_param(sc_mul(sigma, rho_inv)), to_param(sigma), d_sx, d_sr);
        if (nblk) WL(k_wnla_dots, (unsigned)nblk, 128, cc, ll, nn, Lh, Lg, to_param(mu2), d_part, (size_t)0);
        Sc sums[4];
        rc = sum_partials_to_host(st, d_part, nblk, 4, sums);
        if (rc != BPPP_OK) break;
        Sc vx = sc_add(sc_mul(sums[0], sc_dbl(rho_inv)), sums[1]);       // wnla.rs:145-148
        Sc vr = sc_add(sums[2], sums[3]);                                // wnla.rs:150
        CUDA_OK(cudaMemcpyAsync(d_sx + 8 * (Lh + Lg), vx.v, 32, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(d_sr + 8 * (Lh + Lg), vr.v, 32, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));                              // vx / vr live on this stack frame
        rc = msm_device(st, pts, d_sx, Lt, nullptr, d_x30); if (rc != BPPP_OK) break;
        rc = msm_device(st, pts, d_sr, Lt, nullptr, d_r30); if (rc != BPPP_OK) break;
        // transcript (wnla.rs:162-168)
        if (com_pending) { CUDA_OK(cudaStreamWaitEvent(st, ev_com, 0)); CUDA_OK(cudaMemcpyAsync(d_com30, d_com_next, PT_BYTES, cudaMemcpyDeviceToDevice, st)); com_pending = false; }
        CUDA_OK(cudaMemcpyAsync(d_three, d_com30, PT_BYTES, cudaMemcpyDeviceToDevice, st));
        CUDA_OK(cudaMemcpyAsync(d_three + PT_W, d_xr30, 2 * PT_BYTES, cudaMemcpyDeviceToDevice, st));
        uint8_t b[99];
        rc = encode_points_from_device(st, d_three, 3, FMT_COMPRESSED, b); if (rc != BPPP_OK) break;
        host_append_point33(t, BPPP_LBL("wnla_com"), b);
        host_append_point33(t, BPPP_LBL("wnla_x"), b + 33);
        host_append_point33(t, BPPP_LBL("wnla_r"), b + 66);
        merlin_append_u64(t, BPPP_LBL("l.sz"), (uint64_t)len_l);
        merlin_append_u64(t, BPPP_LBL("n.sz"), (uint64_t)len_n);
        Sc y;
        if (!host_challenge(t, BPPP_LBL("wnla_challenge"), y)) { *status = ST_PANIC_CHALLENGE_RANGE; break; }
        xs.emplace_back(b + 33, b + 66); rs.emplace_back(b + 66, b + 99);
        if (!first_round) {
            // com' = com + y X + (y^2 - 1) R on the side stream (reads d_three: this round's com, X, R stay untouched until the next encode)
            CUDA_OK(cudaEventRecord(ev_xr, st));
            CUDA_OK(cudaStreamWaitEvent(side, ev_xr, 0));
            k_wnla_next_commitment<<<1, 2, 0, side>>>(d_three, d_three + PT_W, d_three + 2 * PT_W, to_param(y), d_com_next); g_wnla_launches++;
            CUDA_OK(cudaEventRecord(ev_com, side));
            com_pending = true;
        }
        // fold generators and scalars (wnla.rs:170-175)
        size_t Lh2 = (Lh + 1) / 2, Lg2 = (Lg + 1) / 2;
        if (Lh2) WL(k_wnla_fold_points, nblocks(Lh2, 64), 64, pts, Lh, to_param(y), pts2);
        if (Lg2) WL(k_wnla_fold_points, nblocks(Lg2, 64), 64, pts + 16 * Lh, Lg, to_param(sc_mul(y, rho_inv)), pts2 + 16 * Lh2);
        w.sigma = sc_mul(sigma, w.rho); w.scaled = true;                 // g' = rho g0 + y g1 = (rho sigma) (G0 + (y / rho) G1)
        CUDA_OK(cudaMemcpyAsync(pts2 + 16 * (Lh2 + Lg2), pts + 16 * (Lh + Lg), 64, cudaMemcpyDeviceToDevice, st));
        WL(k_wnla_fold_scalars, nblocks(std::max(Lh2, Lg2), 128), 128, cc, ll, nn, Lh, Lg, to_param(y), to_param(rho_inv), c2, l2, n2, 1);
        std::swap(pts, pts2); std::swap(cc, c2); std::swap(ll, l2); std::swap(nn, n2);
        w.Lh = Lh2; w.Lg = Lg2;
        len_l = (len_l + 1) / 2; len_n = (len_n + 1) / 2;
        w.len_h = (w.len_h + 1) / 2; w.len_g = (w.len_g + 1) / 2;
        w.rho = w.mu; w.mu = mu2;
        if (first_round) {
            // the reference recomputes wnla'.commit(l', n') (wnla.rs:186); C + yX + (y^2-1)R equals it only when the caller's
            // commitment was consistent with (l, n), so the first re-commit is evaluated literally, later ones by the identity
            WnlaDev cur = w; cur.pts = pts; cur.c = cc;
            rc = wnla_commit_dev(st, cur, ll, nn, d_com30); if (rc != BPPP_OK) break;
            first_round = false;
        }
    }
    if (rc == BPPP_OK && com_pending) { CUDA_OK(cudaStreamSynchronize(side)); }
    if (rc == BPPP_OK) {
        CUDA_OK(cudaStreamSynchronize(st));
        // proof.r / proof.x are pushed after the recursion returns: innermost round first (wnla.rs:186-188)
        proof.r33.clear(); proof.x33.clear();
        for (size_t k = rs.size(); k-- > 0;) { proof.r33.insert(proof.r33.end(), rs[k].begin(), rs[k].end()); proof.x33.insert(proof.x33.end(), xs[k].begin(), xs[k].end()); }
        std::vector<uint32_t> hl(8 * (len_l ? len_l : 1)), hn(8 * (len_n ? len_n : 1));
        if (len_l) CUDA_OK(cudaMemcpy(hl.data(), ll, 32 * len_l, cudaMemcpyDeviceToHost));
        if (len_n) CUDA_OK(cudaMemcpy(hn.data(), nn, 32 * len_n, cudaMemcpyDeviceToHost));
        proof.l32.resize(32 * len_l); proof.n32.resize(32 * len_n);
        for (size_t i = 0; i < len_l; i++) { Sc s; memcpy(s.v, &hl[8 * i], 32); sc_to_be32(&proof.l32[32 * i], s); }
        for (size_t i = 0; i < len_n; i++) { Sc s; memcpy(s.v, &hn[8 * i], 32); sc_to_be32(&proof.n32[32 * i], s); }
    }
    cudaStreamSynchronize(side);
    cudaStreamSynchronize(st);
    // the caller releases w.pts / w.c (set A); they may currently be the "other" set, which is fine: both stay allocated until here
    w.pts = ptsA; w.c = cA;
    return rc;
}

// wnla.rs:75-121.  verdict: 1 / 0, or a negative status where the reference panics.
int wnla_verify_dev(cudaStream_t st, WnlaDev &w, Merlin &t, uint32_t *d_com30, const uint8_t *r33, size_t rn, const uint8_t *x33, size_t xn,
                    const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn, int32_t *verdict) {
    if (xn != rn) { *verdict = ST_FALSE; return BPPP_OK; }      // wnla.rs:76-78
    const int R = (int)xn;
    int rc = BPPP_OK;
    uint32_t *d_xr = nullptr, *d_x30 = nullptr, *d_r30 = nullptr, *d_l = nullptr, *d_n = nullptr, *d_ys = nullptr, *d_rhos = nullptr;
    uint32_t *d_sc = nullptr, *d_cl = nullptr, *d_part = nullptr, *d_f30 = nullptr;
    DevScope scope;                                             // every return path below releases what was allocated so far
    scope.own(&d_xr); scope.own(&d_x30); scope.own(&d_r30); scope.own(&d_l); scope.own(&d_n); scope.own(&d_ys); scope.own(&d_rhos);
    scope.own(&d_sc); scope.own(&d_cl); scope.own(&d_part); scope.own(&d_f30);
    // decode proof points (a malformed encoding is a deserialisation failure in the reference)
    std::vector<uint8_t> both(33 * 2 * (size_t)(R ? R : 1));
    memcpy(both.data(), x33, 33 * (size_t)R); memcpy(both.data() + 33 * (size_t)R, r33, 33 * (size_t)R);
    rc = decode_points_to_device(st, both.data(), FMT_COMPRESSED, 2 * (size_t)R, &d_xr);
    // only a bad ENCODING is a verdict (the reference fails to deserialise); out-of-memory / CUDA errors stay errors
    if (rc == BPPP_ERR_ENCODING) { *verdict = ST_BAD_POINT; return BPPP_OK; }
    if (rc != BPPP_OK) return rc;
    rc = decode_scalars_to_device(st, l32, ln, &d_l);
    if (rc == BPPP_OK) rc = decode_scalars_to_device(st, n32, nn, &d_n);
    if (rc != BPPP_OK) { if (rc != BPPP_ERR_ENCODING) return rc; *verdict = ST_BAD_SCALAR; return BPPP_OK; }
    CUDA_OK(cudaMalloc(&d_x30, PT_BYTES)); CUDA_OK(cudaMalloc(&d_r30, PT_BYTES));
    std::vector<Sc> ys(R), rhos(R);
    Sc rho = w.rho, mu = w.mu;
    size_t len_h = w.len_h, len_g = w.len_g;
    *verdict = ST_TRUE;
    for (int j = 0; j < R; j++) {
        int idx = R - 1 - j;                                   // proof.x.last() (wnla.rs:89-90)
        uint8_t cb[33];
        rc = encode_points_from_device(st, d_com30, 1, FMT_COMPRESSED, cb); if (rc != BPPP_OK) break;
        host_append_point33(t, BPPP_LBL("wnla_com"), cb);
        host_append_point33(t, BPPP_LBL("wnla_x"), x33 + 33 * (size_t)idx);
        host_append_point33(t, BPPP_LBL("wnla_r"), r33 + 33 * (size_t)idx);
        merlin_append_u64(t, BPPP_LBL("l.sz"), (uint64_t)len_h);
        merlin_append_u64(t, BPPP_LBL("n.sz"), (uint64_t)len_g);
        Sc y;
        if (!host_challenge(t, BPPP_LBL("wnla_challenge"), y)) { *verdict = ST_PANIC_CHALLENGE_RANGE; break; }
        ys[j] = y; rhos[j] = rho;
        WL(k_decode_one_point30, 1, 1, d_xr + 16 * (size_t)idx, d_x30);
        WL(k_decode_one_point30, 1, 1, d_xr + 16 * ((size_t)R + idx), d_r30);
        WL(k_wnla_next_commitment, 1, 2, d_com30, d_x30, d_r30, to_param(y), d_com30);
        len_h = (len_h + 1) / 2; len_g = (len_g + 1) / 2;
        rho = mu; mu = sc_sqr(mu);
    }
    if (rc == BPPP_OK && *verdict == ST_TRUE) {
        // base case: commitment == commit(l, n) over the folded generators == one MSM over the original ones
        size_t Lh = w.Lh, Lg = w.Lg, Lt = Lh + Lg + 1;
        CUDA_OK(cudaMalloc(&d_sc, 32 * Lt)); CUDA_OK(cudaMalloc(&d_cl, 32 * (Lh ? Lh : 1)));
        CUDA_OK(cudaMalloc(&d_ys, 32 * (size_t)(R ? R : 1))); CUDA_OK(cudaMalloc(&d_rhos, 32 * (size_t)(R ? R : 1)));
        CUDA_OK(cudaMalloc(&d_f30, PT_BYTES));
        if (R) { CUDA_OK(cudaMemcpyAsync(d_ys, ys.data(), 32 * (size_t)R, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(d_rhos, rhos.data(), 32 * (size_t)R, cudaMemcpyHostToDevice, st)); }
        WL(k_wnla_final_scalars, nblocks(Lh + Lg, 128), 128, Lh, Lg, R, d_ys, d_rhos, d_l, ln, d_n, nn, w.c, Lh, d_sc, d_cl);
        size_t nblk = (Lh + 127) / 128;
        CUDA_OK(cudaMalloc(&d_part, 32 * (nblk ? nblk : 1)));
        if (nblk) WL(k_sc_sum_partials, (unsigned)nblk, 128, d_cl, Lh, d_part);
        Sc cl;
        rc = sum_partials_to_host(st, d_part, nblk, 1, &cl);
        if (rc == BPPP_OK) {
            // |n|^2_{mu_R}: sum_m n[m]^2 mu_R^(m+1) over the proof's (short) n
            std::vector<uint8_t> nb(n32, n32 + 32 * nn);
            Sc wn = sc_zero(), pw = sc_one();
            for (size_t m = 0; m < nn; m++) { Sc v; sc_from_be32(v, &nb[32 * m]); pw = sc_mul(pw, mu); wn = sc_add(wn, sc_mul(sc_sqr(v), pw)); }
            Sc v = sc_add(cl, wn);
            CUDA_OK(cudaMemcpyAsync(d_sc + 8 * (Lh + Lg), v.v, 32, cudaMemcpyHostToDevice, st));
            rc = msm_device(st, w.pts, d_sc, Lt, nullptr, d_f30);
        }
        if (rc == BPPP_OK) {
            uint8_t a[33], b[33];
            rc = encode_points_from_device(st, d_com30, 1, FMT_COMPRESSED, a);
            if (rc == BPPP_OK) rc = encode_points_from_device(st, d_f30, 1, FMT_COMPRESSED, b);
            if (rc == BPPP_OK) *verdict = memcmp(a, b, 33) == 0 ? ST_TRUE : ST_FALSE;     // ProjectivePoint::eq (wnla.rs:81)
        }
    }
    return rc;
}

uint64_t wnla_launch_count() { return g_wnla_launches.load(); }

int point_bytes_to_pt30(cudaStream_t st, const uint8_t *p, int fmt, uint32_t *d_out30) {
    uint32_t *d16 = nullptr;
    int rc = decode_points_to_device(st, p, fmt, 1, &d16);
    if (rc != BPPP_OK) return rc;
    k_decode_one_point30<<<1, 1, 0, st>>>(d16, d_out30);
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(d16);
    return BPPP_OK;
}

}  // namespace bppp

static int pick_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BPPP_ERR_ARG, "bad device index");
    CUDA_OK(cudaSetDevice(device));
    return BPPP_OK;
}

// WeightNormLinearArgument::commit (src/wnla.rs:66-72)
extern "C" int bppp_wnla_commit(int device, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn, const uint8_t *c32, size_t cn,
                                const uint8_t *rho32, const uint8_t *mu32, const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn, uint8_t *out33) {
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    cudaStream_t st = nullptr;
    WnlaDev w;
    rc = wnla_load(st, w, g64, gvec64, gn, hvec64, hn, c32, cn, rho32, mu32, ln, nn); if (rc != BPPP_OK) return rc;
    uint32_t *d_l = nullptr, *d_n = nullptr, *d_out = nullptr;
    rc = upload_padded_scalars(st, l32, ln, w.Lh, &d_l);
    if (rc == BPPP_OK) rc = upload_padded_scalars(st, n32, nn, w.Lg, &d_n);
    if (rc == BPPP_OK) { CUDA_OK(cudaMalloc(&d_out, PT_BYTES)); rc = wnla_commit_dev(st, w, d_l, d_n, d_out); }
    if (rc == BPPP_OK) rc = encode_points_from_device(st, d_out, 1, FMT_COMPRESSED, out33);
    cudaFree(d_l); cudaFree(d_n); cudaFree(d_out); w.release();
    return rc;
}

// WeightNormLinearArgument::prove (src/wnla.rs:125-190) with a fresh Transcript::new(label).
// r_out / x_out: rounds x 33 bytes (innermost round first); l_out / n_out: final vectors.  *status: 1 or a panic code.
extern "C" int bppp_wnla_prove(int device, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn, const uint8_t *c32, size_t cn,
                               const uint8_t *rho32, const uint8_t *mu32, const uint8_t *commit33, const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn,
                               const uint8_t *label, size_t label_len, uint8_t *r_out, uint8_t *x_out, size_t *rounds_out, uint8_t *l_out, size_t *l_out_len,
                               uint8_t *n_out, size_t *n_out_len, int32_t *status) {
    if (!status || !rounds_out || !l_out_len || !n_out_len) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    cudaStream_t st = nullptr;
    WnlaDev w;
    rc = wnla_load(st, w, g64, gvec64, gn, hvec64, hn, c32, cn, rho32, mu32, ln, nn); if (rc != BPPP_OK) return rc;
    uint32_t *d_l = nullptr, *d_n = nullptr, *d_com16 = nullptr, *d_com30 = nullptr;
    rc = upload_padded_scalars(st, l32, ln, w.Lh, &d_l);
    if (rc == BPPP_OK) rc = upload_padded_scalars(st, n32, nn, w.Lg, &d_n);
    if (rc == BPPP_OK) rc = decode_points_to_device(st, commit33, FMT_COMPRESSED, 1, &d_com16);
    if (rc != BPPP_OK) { cudaFree(d_l); cudaFree(d_n); w.release(); return rc; }
    CUDA_OK(cudaMalloc(&d_com30, PT_BYTES));
    k_decode_one_point30<<<1, 1, 0, st>>>(d_com16, d_com30);
    Merlin t; merlin_init(t, label, (uint32_t)label_len);
    WnlaProofHost proof;
    *status = ST_TRUE;
    rc = wnla_prove_dev(st, w, t, d_com30, d_l, d_n, ln, nn, proof, status);    // consumes d_l, d_n
    if (rc == BPPP_OK && *status == ST_TRUE) {
        memcpy(r_out, proof.r33.data(), proof.r33.size()); memcpy(x_out, proof.x33.data(), proof.x33.size());
        *rounds_out = proof.r33.size() / 33;
        memcpy(l_out, proof.l32.data(), proof.l32.size()); *l_out_len = proof.l32.size() / 32;
        memcpy(n_out, proof.n32.data(), proof.n32.size()); *n_out_len = proof.n32.size() / 32;
    }
    cudaFree(d_com16); cudaFree(d_com30); w.release();
    return rc;
}

// WeightNormLinearArgument::verify (src/wnla.rs:75-121) with a fresh Transcript::new(label)
extern "C" int bppp_wnla_verify(int device, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn, const uint8_t *c32, size_t cn,
                                const uint8_t *rho32, const uint8_t *mu32, const uint8_t *commit33, const uint8_t *r33, size_t rn, const uint8_t *x33, size_t xn,
                                const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn, const uint8_t *label, size_t label_len, int32_t *verdict) {
    if (!verdict) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    cudaStream_t st = nullptr;
    WnlaDev w;
    rc = wnla_load(st, w, g64, gvec64, gn, hvec64, hn, c32, cn, rho32, mu32, 0, 0); if (rc != BPPP_OK) return rc;
    uint32_t *d_com16 = nullptr, *d_com30 = nullptr;
    rc = decode_points_to_device(st, commit33, FMT_COMPRESSED, 1, &d_com16);
    if (rc != BPPP_OK) { w.release(); if (rc != BPPP_ERR_ENCODING) return rc; *verdict = ST_BAD_POINT; return BPPP_OK; }
    CUDA_OK(cudaMalloc(&d_com30, PT_BYTES));
    k_decode_one_point30<<<1, 1, 0, st>>>(d_com16, d_com30);
    Merlin t; merlin_init(t, label, (uint32_t)label_len);
    rc = wnla_verify_dev(st, w, t, d_com30, r33, rn, x33, xn, l32, ln, n32, nn, verdict);
    cudaFree(d_com16); cudaFree(d_com30); w.release();
    return rc;
}


// ---- a block of a standalone WNLA instance resident on one GPU (BASELINE config 5) ----------------------------
// The generator index range is cut into contiguous blocks with even offsets, one per GPU.  Folding maps the pair
// (2i, 2i + 1) to i (util.rs:7-22), so a block with an even offset and an even length folds locally; each round a block
// contributes ONE partial point to X and one to R (its share of vx / vr rides on g inside the partial), the partials
// of all blocks are added (the only exchange: 2 x 64 bytes per block per round) and every holder runs the identical
// transcript.  A shard created with whole = 1 is the entire instance (odd lengths fold with the reference's zero
// extension): that is also the stepped single-GPU prover for a caller-owned transcript.
struct bppp_wnla_shard {
    int device = 0;
    cudaStream_t st = nullptr;
    size_t nh = 0, ng = 0, h_off = 0, g_off = 0;
    int whole = 0;
    uint32_t *pts[2] = {nullptr, nullptr};      // [H (nh) | G (ng) | g], 16 words each; ping-pong across folds
    uint32_t *c[2] = {nullptr, nullptr}, *l[2] = {nullptr, nullptr}, *n[2] = {nullptr, nullptr};
    uint32_t *sx = nullptr, *sr = nullptr, *part = nullptr, *out30 = nullptr;
    int cur = 0;
    Sc rho, mu;
    Sc sigma;                                   // the stored G points are the true ones divided by sigma (WnlaDev::sigma)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

extern "C" void bppp_wnla_shard_destroy(bppp_wnla_shard *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (int k = 0; k < 2; k++) { cudaFree(s->pts[k]); cudaFree(s->c[k]); cudaFree(s->l[k]); cudaFree(s->n[k]); }
    cudaFree(s->sx); cudaFree(s->sr); cudaFree(s->part); cudaFree(s->out30);
    if (s->e0) cudaEventDestroy(s->e0);
    if (s->e1) cudaEventDestroy(s->e1);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

extern "C" int bppp_wnla_shard_create(bppp_wnla_shard **out, int device, const uint8_t *g64, const uint8_t *hvec64, const uint8_t *c32, const uint8_t *l32,
                                      size_t nh, size_t h_off, const uint8_t *gvec64, const uint8_t *n32, size_t ng, size_t g_off, const uint8_t *rho32,
                                      const uint8_t *mu32, int whole) {
    if (!out || !g64 || !rho32 || !mu32 || (nh && (!hvec64 || !c32 || !l32)) || (ng && (!gvec64 || !n32))) return fail(BPPP_ERR_ARG, "null argument");
    *out = nullptr;
    if (!whole && ((h_off | g_off) & 1)) return fail(BPPP_ERR_ARG, "a block must start at an even index");
    int rc = pick_device(device); if (rc != BPPP_OK) return rc;
    bppp_wnla_shard *s = new bppp_wnla_shard();
    struct Guard { bppp_wnla_shard *s; ~Guard() { if (s) bppp_wnla_shard_destroy(s); } } guard{s};
    s->device = device; s->nh = nh; s->ng = ng; s->h_off = h_off; s->g_off = g_off; s->whole = whole;
    if (!sc_from_be32(s->rho, rho32) || !sc_from_be32(s->mu, mu32)) return fail(BPPP_ERR_ARG, "rho/mu not canonical");
    s->sigma = sc_one();
    CUDA_OK(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&s->e0)); CUDA_OK(cudaEventCreate(&s->e1));
    cudaStream_t st = s->st;
    std::vector<uint8_t> pb(64 * (nh + ng + 1));
    if (nh) memcpy(pb.data(), hvec64, 64 * nh);
    if (ng) memcpy(pb.data() + 64 * nh, gvec64, 64 * ng);
    memcpy(pb.data() + 64 * (nh + ng), g64, 64);
    rc = decode_points_to_device(st, pb.data(), FMT_AFFINE64, nh + ng + 1, &s->pts[0]); if (rc != BPPP_OK) return rc;
    rc = decode_scalars_to_device(st, c32, nh, &s->c[0]); if (rc != BPPP_OK) return rc;
    rc = decode_scalars_to_device(st, l32, nh, &s->l[0]); if (rc != BPPP_OK) return rc;
    rc = decode_scalars_to_device(st, n32, ng, &s->n[0]); if (rc != BPPP_OK) return rc;
    const size_t nh2 = (nh + 1) / 2, ng2 = (ng + 1) / 2, half = std::max(nh, ng) + 1;
    CUDA_OK(cudaMalloc(&s->pts[1], 64 * (nh2 + ng2 + 1)));
    CUDA_OK(cudaMalloc(&s->c[1], 32 * (nh2 ? nh2 : 1))); CUDA_OK(cudaMalloc(&s->l[1], 32 * (nh2 ? nh2 : 1))); CUDA_OK(cudaMalloc(&s->n[1], 32 * (ng2 ? ng2 : 1)));
    CUDA_OK(cudaMalloc(&s->sx, 32 * (nh + ng + 1))); CUDA_OK(cudaMalloc(&s->sr, 32 * (nh + ng + 1)));
    CUDA_OK(cudaMalloc(&s->part, 128 * ((half + 127) / 128)));
    CUDA_OK(cudaMalloc(&s->out30, 2 * PT_BYTES));
    guard.s = nullptr;
    *out = s;
    return BPPP_OK;
}

extern "C" int bppp_wnla_shard_state(const bppp_wnla_shard *s, size_t *nh, size_t *ng, size_t *h_off, size_t *g_off, uint8_t *rho32, uint8_t *mu32) {
    if (!s) return fail(BPPP_ERR_ARG, "null argument");
    if (nh) *nh = s->nh;
    if (ng) *ng = s->ng;
    if (h_off) *h_off = s->h_off;
    if (g_off) *g_off = s->g_off;
    if (rho32) sc_to_be32(rho32, s->rho);
    if (mu32) sc_to_be32(mu32, s->mu);
    return BPPP_OK;
}

// this block's share of wnla.commit(l, n) (wnla.rs:66-72): (<c, l> + |n|^2_mu restricted to the block) g + <h, l> + <g_vec, n>
extern "C" int bppp_wnla_shard_commit_partial(bppp_wnla_shard *s, uint8_t *out64) {
    if (!s || !out64) return fail(BPPP_ERR_ARG, "null argument");
    CUDA_OK(cudaSetDevice(s->device));
    cudaStream_t st = s->st;
    const int k = s->cur;
    const size_t L = std::max(s->nh, s->ng), nblk = (L + 127) / 128, Lt = s->nh + s->ng + 1;
    if (nblk) WL(k_commit_dots, (unsigned)nblk, 128, s->c[k], s->l[k], s->nh, s->n[k], s->ng, to_param(s->mu), s->part, s->g_off);
    Sc sums[2];
    int rc = sum_partials_to_host(st, s->part, nblk, 2, sums); if (rc != BPPP_OK) return rc;
    Sc v = sc_add(sums[0], sums[1]);
    if (s->nh) CUDA_OK(cudaMemcpyAsync(s->sx, s->l[k], 32 * s->nh, cudaMemcpyDeviceToDevice, st));
    if (s->ng) CUDA_OK(cudaMemcpyAsync(s->sx + 8 * s->nh, s->n[k], 32 * s->ng, cudaMemcpyDeviceToDevice, st));
    if (s->ng) WL(k_sc_scale, nblocks(s->ng, 128), 128, s->sx + 8 * s->nh, s->ng, to_param(s->sigma));
    CUDA_OK(cudaMemcpyAsync(s->sx + 8 * (s->nh + s->ng), v.v, 32, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));       // v is a stack variable
    rc = msm_device(st, s->pts[k], s->sx, Lt, nullptr, s->out30); if (rc != BPPP_OK) return rc;
    return encode_points_from_device(st, s->out30, 1, FMT_AFFINE64, out64);
}

// this block's shares of X and R (wnla.rs:143-160), 64-byte affine each; *device_ms: device time of the kernels
extern "C" int bppp_wnla_shard_xr_partial(bppp_wnla_shard *s, uint8_t *out128, float *device_ms) {
    if (!s || !out128) return fail(BPPP_ERR_ARG, "null argument");
    if (sc_is_zero(s->rho)) return fail(BPPP_ERR_ARG, "rho is zero: the reference panics (rho.invert_vartime().unwrap(), wnla.rs:135)");
    CUDA_OK(cudaSetDevice(s->device));
    cudaStream_t st = s->st;
    const int k = s->cur;
    const size_t Lh = s->nh, Lg = s->ng, Lt = Lh + Lg + 1;
    Sc rho_inv = sc_inv(s->rho), mu2 = sc_sqr(s->mu);
    const size_t half = (std::max(Lh, Lg) + 1) / 2, nblk = (half + 127) / 128;
    CUDA_OK(cudaEventRecord(s->e0, st));
    if (Lh + Lg) WL(k_wnla_xr_scalars, nblocks(Lh + Lg, 128), 128, s->l[k], s->n[k], Lh, Lg, to_param(sc_mul(s->sigma, s->rho)), to_param(sc_mul(s->sigma, rho_inv)),
                    to_param(s->sigma), s->sx, s->sr);
    if (nblk) WL(k_wnla_dots, (unsigned)nblk, 128, s->c[k], s->l[k], s->n[k], Lh, Lg, to_param(mu2), s->part, s->g_off / 2);
    Sc sums[4];
    int rc = sum_partials_to_host(st, s->part, nblk, 4, sums); if (rc != BPPP_OK) return rc;
    Sc vx = sc_add(sc_mul(sums[0], sc_dbl(rho_inv)), sums[1]);       // wnla.rs:145-148, this block's terms
    Sc vr = sc_add(sums[2], sums[3]);                                // wnla.rs:150
    CUDA_OK(cudaMemcpyAsync(s->sx + 8 * (Lh + Lg), vx.v, 32, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(s->sr + 8 * (Lh + Lg), vr.v, 32, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
    rc = msm_device(st, s->pts[k], s->sx, Lt, nullptr, s->out30); if (rc != BPPP_OK) return rc;
    rc = msm_device(st, s->pts[k], s->sr, Lt, nullptr, s->out30 + PT_W); if (rc != BPPP_OK) return rc;
    CUDA_OK(cudaEventRecord(s->e1, st));
    rc = encode_points_from_device(st, s->out30, 2, FMT_AFFINE64, out128); if (rc != BPPP_OK) return rc;
    if (device_ms) CUDA_OK(cudaEventElapsedTime(device_ms, s->e0, s->e1));
    return BPPP_OK;
}

// fold the block with the round's challenge (wnla.rs:170-175, 177-184): h' = h0 + y h1, g' = rho g0 + y g1, c' = c0 + y c1,
// l' = l0 + y l1, n' = rho^-1 n0 + y n1; rho <- mu, mu <- mu^2; offsets and lengths halve
extern "C" int bppp_wnla_shard_fold(bppp_wnla_shard *s, const uint8_t *y32, float *device_ms) {
    if (!s || !y32) return fail(BPPP_ERR_ARG, "null argument");
    if (!s->whole && ((s->nh | s->ng | s->h_off | s->g_off) & 1))
        return fail(BPPP_ERR_ARG, "this block no longer folds locally (odd length or offset): gather the blocks into a whole instance");
    Sc y;
    if (!sc_from_be32(y, y32)) return fail(BPPP_ERR_ARG, "challenge is not a canonical scalar");
    if (sc_is_zero(s->rho)) return fail(BPPP_ERR_ARG, "rho is zero: the reference panics (wnla.rs:135)");
    CUDA_OK(cudaSetDevice(s->device));
    cudaStream_t st = s->st;
    const int k = s->cur, o = k ^ 1;
    const size_t Lh = s->nh, Lg = s->ng, Lh2 = (Lh + 1) / 2, Lg2 = (Lg + 1) / 2;
    Sc rho_inv = sc_inv(s->rho);
    CUDA_OK(cudaEventRecord(s->e0, st));
    if (Lh2) WL(k_wnla_fold_points, nblocks(Lh2, 64), 64, s->pts[k], Lh, to_param(y), s->pts[o]);
    if (Lg2) WL(k_wnla_fold_points, nblocks(Lg2, 64), 64, s->pts[k] + 16 * Lh, Lg, to_param(sc_mul(y, rho_inv)), s->pts[o] + 16 * Lh2);
    s->sigma = sc_mul(s->sigma, s->rho);
    CUDA_OK(cudaMemcpyAsync(s->pts[o] + 16 * (Lh2 + Lg2), s->pts[k] + 16 * (Lh + Lg), 64, cudaMemcpyDeviceToDevice, st));
    if (Lh2 + Lg2) WL(k_wnla_fold_scalars, nblocks(std::max(Lh2, Lg2), 128), 128, s->c[k], s->l[k], s->n[k], Lh, Lg, to_param(y), to_param(rho_inv), s->c[o], s->l[o], s->n[o], 1);
    CUDA_OK(cudaEventRecord(s->e1, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    if (device_ms) CUDA_OK(cudaEventElapsedTime(device_ms, s->e0, s->e1));
    s->cur = o; s->nh = Lh2; s->ng = Lg2; s->h_off /= 2; s->g_off /= 2;
    Sc mu2 = sc_sqr(s->mu);
    s->rho = s->mu; s->mu = mu2;
    return BPPP_OK;
}

__global__ void k_words_to_be(const uint32_t *w, size_t n, int words, uint8_t *out) {      // 8 LE words per 32-byte big-endian field
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (size_t)(words / 8)) return;
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = w[8 * t + k];
    words_to_be32(out + 32 * t, v);
}
// the block's current contents (64-byte affine points, 32-byte scalars): what the holders exchange once the blocks are too
// short to fold locally, and how a proof's final l, n leave the device
extern "C" int bppp_wnla_shard_export(bppp_wnla_shard *s, uint8_t *hvec64, uint8_t *c32, uint8_t *l32, uint8_t *gvec64, uint8_t *n32) {
    if (!s) return fail(BPPP_ERR_ARG, "null argument");
    CUDA_OK(cudaSetDevice(s->device));
    cudaStream_t st = s->st;
    const int k = s->cur;
    uint8_t *d_tmp = nullptr;
    const size_t cap = 64 * std::max(s->nh, s->ng) + 64;
    CUDA_OK(cudaMalloc(&d_tmp, cap));
    auto dump = [&](const uint32_t *src, size_t n, int words, uint8_t *dst) -> int {
        if (!dst || !n) return BPPP_OK;
        WL(k_words_to_be, nblocks(n * (size_t)(words / 8), 128), 128, src, n, words, d_tmp);
        CUDA_OK(cudaMemcpyAsync(dst, d_tmp, 4 * (size_t)words * n, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        return BPPP_OK;
    };
    if (gvec64 && s->ng && !sc_is_one_host(s->sigma)) {          // hand out the true generators: sigma * stored (a few points by the time blocks are exported)
        WL(k_points_scale, nblocks(s->ng, 64), 64, s->pts[k] + 16 * s->nh, s->ng, to_param(s->sigma));
        s->sigma = sc_one();
    }
    int rc = dump(s->pts[k], s->nh, 16, hvec64);
    if (rc == BPPP_OK) rc = dump(s->c[k], s->nh, 8, c32);
    if (rc == BPPP_OK) rc = dump(s->l[k], s->nh, 8, l32);
    if (rc == BPPP_OK) rc = dump(s->pts[k] + 16 * s->nh, s->ng, 16, gvec64);
    if (rc == BPPP_OK) rc = dump(s->n[k], s->ng, 8, n32);
    cudaFree(d_tmp);
    return rc;
}
