// libbppp.so: CUDA kernels for sm_100a + the C ABI of include/bppp.h.
//
// A batch of N independent proofs is run in lockstep as a sequence of data-parallel kernels over a
// word-major workspace in HBM (ws.cuh).  Three kernel shapes carry the work:
//   * k_msm_fixed   -- fixed-base multi-scalar multiplication from per-generator window tables in HBM
//                      (LANES threads per proof, warp-shuffle reduction of the partial points);
//   * k_*_var*      -- joint variable-base Straus ladder for the per-proof proof points (1 thread/proof);
//   * k_batch_inv   -- Montgomery batch inversion across proofs for the affine normalisations that
//                      feed the Fiat-Shamir transcript;
// the transcript itself (Merlin/STROBE/Keccak) and all challenge-derived scalar algebra run on the
// device, one proof per thread, so a batch never returns to the host between phases.
// There is no CPU fallback: without a CUDA device every entry point fails with BPPP_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bppp.h"
#include "u64_prove.cuh"
#include "u64_verify.cuh"

using namespace bppp;

static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) { g_last_error = msg; return code; }
#define CUDA_OK(expr)                                                                                      \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) return fail(BPPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct TermMap { int gen[NUM_GENS]; };

struct bppp_ctx {
    int device = 0;
    FixedTable T{};
    uint4 *d_tab = nullptr;
    size_t table_bytes = 0;
    double table_build_ms = 0;
    size_t max_batch = 0;
    uint32_t *d_ws = nullptr;       // workspace, max(VL::WORDS, PL::WORDS) * max_batch words
    size_t ws_words_per_proof = 0;
    uint8_t *d_in_a = nullptr, *d_in_b = nullptr, *d_in_c = nullptr;   // staging for host-buffer entry points
    uint8_t *d_out = nullptr;
    int32_t *d_status = nullptr;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
};

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <int LANES>
__device__ __forceinline__ Pt lanes_reduce(Pt acc) {
#pragma unroll 1
    for (int off = LANES / 2; off >= 1; off >>= 1) {
        Pt o;
#pragma unroll
        for (int k = 0; k < 10; k++) {
            o.x.n[k] = __shfl_xor_sync(0xFFFFFFFFu, acc.x.n[k], off);
            o.y.n[k] = __shfl_xor_sync(0xFFFFFFFFu, acc.y.n[k], off);
            o.z.n[k] = __shfl_xor_sync(0xFFFFFFFFu, acc.z.n[k], off);
        }
        acc = pt_add(acc, o);
    }
    return acc;
}

// sum_t scalar[t] * G_{gen[t]} for every proof: LANES threads per proof
template <int LANES>
__global__ void __launch_bounds__(128) k_msm_fixed(FixedTable T, WS w, int sc_off, TermMap tm, int nterms, int out_off) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = tid / LANES;
    int lane = (int)(tid % LANES);
    bool live = i < w.n;
    if (!live) i = w.n - 1;
    Pt acc = msm_fixed_lane(T, w, i, sc_off, tm.gen, nterms, lane, LANES);
    acc = lanes_reduce<LANES>(acc);
    if (live && lane == 0) ws_st_pt(w, i, out_off, acc);
}

__global__ void __launch_bounds__(128) k_batch_inv(WS w, int in_off, int out_off, size_t nthreads) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) batch_inv_strided(w, in_off, out_off, t, nthreads, w.n);
}

__global__ void __launch_bounds__(64) k_v_load(WS w, const uint8_t *commits, const uint8_t *proofs, int fmt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    u64v_load_one(w, i, commits + csz * i, proofs + psz * i, fmt);
}
__global__ void __launch_bounds__(64) k_v_phase1(WS w, Merlin init) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_phase1_one(w, i, init);
}
__global__ void __launch_bounds__(64, 7) k_v_var5(WS w) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_var5_one(w, i);
}
__global__ void __launch_bounds__(64) k_v_round(WS w, int j) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_round_one(w, i, j);
}
__global__ void __launch_bounds__(64, 7) k_v_var2(WS w, int j) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_var2_one(w, i, j);
}
__global__ void __launch_bounds__(64) k_v_final_scalars(WS w) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_final_scalars_one(w, i);
}
__global__ void __launch_bounds__(64) k_v_verdict(WS w, int32_t *status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    u64v_verdict_one(w, i);
    status[i] = (int32_t)ws_ld(w, i, VL::STATUS);
}

// ---- prove kernels ----
__global__ void __launch_bounds__(64) k_p_load(WS w, const uint64_t *xs, const uint8_t *blinds) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_load_one(w, i, xs[i], blinds + 32 * i);
}
__global__ void __launch_bounds__(64) k_p_phase1(WS w, Merlin init, const uint8_t *rng) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_phase1_one(w, i, init, rng + (size_t)U64_RNG_BYTES * i);
}
__global__ void __launch_bounds__(64) k_p_phase2(WS w, const uint8_t *rng) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_phase2_one(w, i, rng + (size_t)U64_RNG_BYTES * i);
}
__global__ void __launch_bounds__(64) k_p_phase3(WS w) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_phase3_one(w, i);
}
__global__ void __launch_bounds__(64) k_p_round(WS w, int j) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_round_one(w, i, j);
}
__global__ void __launch_bounds__(64, 7) k_p_var2(WS w, int j) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_var2_one(w, i, j);
}
__global__ void __launch_bounds__(64) k_p_output(WS w, uint8_t *proofs, int32_t *status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    u64p_output_one(w, i, proofs + (size_t)U64_PROOF_BYTES_COMPRESSED * i);
    status[i] = (int32_t)ws_ld(w, i, PL::STATUS);
}
// V' = V + r_com (reciprocal.rs:141 via SURVEY App. C.2)
__global__ void __launch_bounds__(64) k_p_vprime(WS w) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_vprime_one(w, i);
}

// ---- commit ----
__global__ void __launch_bounds__(64) k_c_load(WS w, const uint64_t *xs, const uint8_t *blinds) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    Sc s; int32_t st = ST_TRUE;
    if (!sc_from_be32(s, blinds + 32 * i)) { st = ST_BAD_SCALAR; s = sc_zero(); }
    ws_st_sc(w, i, VL::FS, sc_from_u64(xs[i]));
    ws_st_sc(w, i, VL::FS + 8, s);
    ws_st(w, i, VL::STATUS, (uint32_t)st);
}
__global__ void __launch_bounds__(64) k_c_store(WS w, uint8_t *out, int fmt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    bool id;
    PtA a = ws_affine(w, i, VL::ACC, VL::ZINV, id);
    if (fmt == FMT_COMPRESSED) pta_compress(out + 33 * i, a, id);
    else pta_to_xy64(out + 64 * i, a, id);
}

// ---- fixed-base table construction (one generator per pass) ----
// scratch layout per entry (word-major over nent = nwin * E entries): Pt at 0..29, zinv at 30..39
__global__ void k_tab_bases(WS tmp, PtA gen, bool gen_id, int W, int nwin, uint32_t E) {
    // single thread: B_w = 2^(W w) G, stored at entry (w, d = 1)
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Pt b = pt_from_affine(gen, gen_id);
    for (int w = 0; w < nwin; w++) {
        ws_st_pt(tmp, (size_t)w * E + 0, 0, b);
        for (int k = 0; k < W; k++) b = pt_double(b);
    }
}
// level l >= 1: for m in [2^(l-1), 2^l): E[2m] = 2 E[m], E[2m+1] = E[2m] + B   (entry index = d - 1)
__global__ void __launch_bounds__(128) k_tab_level(WS tmp, int nwin, uint32_t E, int level) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t per = 1u << (level - 1);
    if (t >= (size_t)nwin * per) return;
    uint32_t w = (uint32_t)(t / per), m = per + (uint32_t)(t % per);
    size_t base = (size_t)w * E;
    Pt em = ws_ld_pt(tmp, base + (m - 1), 0);
    Pt b = ws_ld_pt(tmp, base + 0, 0);
    Pt e2 = pt_double(em);
    if (2 * m <= E) ws_st_pt(tmp, base + (2 * m - 1), 0, e2);
    if (2 * m + 1 <= E) ws_st_pt(tmp, base + (2 * m), 0, pt_add(e2, b));
}
__global__ void __launch_bounds__(128) k_tab_write(WS tmp, uint4 *dst) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tmp.n) return;
    bool id;
    PtA a = ws_affine(tmp, t, 0, 30, id);
    uint32_t x[8], y[8];
    fe_to_words(x, a.x); fe_to_words(y, a.y);
    if (id) {
#pragma unroll
        for (int k = 0; k < 8; k++) { x[k] = 0; y[k] = 0; }
    }
    dst[4 * t] = make_uint4(x[0], x[1], x[2], x[3]); dst[4 * t + 1] = make_uint4(x[4], x[5], x[6], x[7]);
    dst[4 * t + 2] = make_uint4(y[0], y[1], y[2], y[3]); dst[4 * t + 3] = make_uint4(y[4], y[5], y[6], y[7]);
}

// ---- microbenchmarks ----
__global__ void __launch_bounds__(256) k_mb_imad(uint64_t *out, uint32_t seed, int iters) {
    uint64_t acc[8];
    uint32_t y = seed | 1u;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = (uint64_t)(threadIdx.x + k) * 0x9E3779B97F4A7C15ULL;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = (uint64_t)(uint32_t)acc[k] * y + acc[k];
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
__global__ void __launch_bounds__(64) k_mb_op(uint32_t *out, uint32_t seed, int iters) {
    Fe a = fe_from_u32((seed + threadIdx.x) & FE_M26), b = fe_from_u32((seed * 3 + 1 + blockIdx.x) & FE_M26);
    a.n[3] = threadIdx.x + 5; b.n[7] = blockIdx.x + 9; b.n[9] = 77;
    Pt p; p.x = a; p.y = b; p.z = fe_from_u32(1);
    PtA q; q.x = b; q.y = a;
    Sc s, u;
#pragma unroll
    for (int k = 0; k < 8; k++) { s.v[k] = seed * (k + 1) + threadIdx.x; u.v[k] = seed + 7 * k + blockIdx.x; }
    s.v[7] &= 0x7FFFFFFFu; u.v[7] &= 0x7FFFFFFFu;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (OP == 1) a = fe_mul(a, b);
        if (OP == 2) a = fe_sqr(a);
        if (OP == 3) s = sc_mul(s, u);
        if (OP == 4) p = pt_add_mixed(p, q);
        if (OP == 5) p = pt_double(p);
        if (OP == 6) p = pt_add(p, p);
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 10; k++) r ^= a.n[k] ^ p.x.n[k] ^ p.y.n[k] ^ p.z.n[k];
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= s.v[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline unsigned nblocks(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }
#define LAUNCH(ctx, kern, grid, block, ...)                        \
    do {                                                           \
        kern<<<(grid), (block), 0, st>>>(__VA_ARGS__);             \
        (ctx)->launches++;                                         \
    } while (0)

static constexpr int MSM_LANES = 8;

static void launch_msm_fixed(bppp_ctx *c, cudaStream_t st, WS w, int sc_off, const TermMap &tm, int nterms, int out_off) {
    size_t threads = w.n * MSM_LANES;
    LAUNCH(c, k_msm_fixed<MSM_LANES>, nblocks(threads, 128), 128, c->T, w, sc_off, tm, nterms, out_off);
}
static void launch_batch_inv(bppp_ctx *c, cudaStream_t st, WS w, int in_off, int out_off) {
    // one inversion per thread, >= 8 items per thread when the batch is large enough to still fill the GPU
    size_t per = 8;
    size_t nthreads = (w.n + per - 1) / per;
    size_t min_threads = (size_t)c->sm_count * 128;
    if (nthreads < min_threads) nthreads = w.n < min_threads ? w.n : min_threads;
    LAUNCH(c, k_batch_inv, nblocks(nthreads, 128), 128, w, in_off, out_off, nthreads);
}

static int build_tables(bppp_ctx *c, const PtA *gens, const bool *gen_id) {
    const int W = c->T.W, nwin = c->T.nwin;
    const uint32_t E = (1u << W) - 1u;
    const size_t nent = (size_t)nwin * E;
    uint32_t *d_tmp = nullptr;
    CUDA_OK(cudaMalloc(&d_tmp, nent * 40 * sizeof(uint32_t)));
    cudaStream_t st = c->stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    WS tmp{d_tmp, nent};
    for (int g = 0; g < NUM_GENS; g++) {
        LAUNCH(c, k_tab_bases, 1, 1, tmp, gens[g], gen_id[g], W, nwin, E);
        for (int level = 1; level < W; level++) {
            size_t threads = (size_t)nwin << (level - 1);
            LAUNCH(c, k_tab_level, nblocks(threads, 128), 128, tmp, nwin, E, level);
        }
        launch_batch_inv(c, st, tmp, 20, 30);
        LAUNCH(c, k_tab_write, nblocks(nent, 128), 128, tmp, c->d_tab + (size_t)g * nent * 4);
    }
    cudaEventRecord(e1, st);
    CUDA_OK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    c->table_build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_OK(cudaFree(d_tmp));
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

extern "C" const char *bppp_last_error(void) { return g_last_error.c_str(); }

extern "C" int bppp_ctx_create(bppp_ctx **out, int device, const uint8_t *gens64, int window_bits, size_t max_batch) {
    if (!out || !gens64) return fail(BPPP_ERR_ARG, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BPPP_ERR_ARG, "bad device index");
    if (window_bits == 0) window_bits = 16;
    if (window_bits < 2 || window_bits > 16) return fail(BPPP_ERR_ARG, "window_bits must be in 2..16");
    if (max_batch == 0) max_batch = 65536;
    PtA gens[NUM_GENS]; bool gen_id[NUM_GENS];
    for (int g = 0; g < NUM_GENS; g++) {
        int s = pta_from_xy64(gens[g], gens64 + 64 * g);
        if (s < 0) return fail(BPPP_ERR_GENERATOR, "generator " + std::to_string(g) + " is not on the curve");
        gens[g].x = fe_normalize(gens[g].x); gens[g].y = fe_normalize(gens[g].y);
        gen_id[g] = s == 1;
    }
    CUDA_OK(cudaSetDevice(device));
    bppp_ctx *c = new bppp_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->T.W = window_bits; c->T.nwin = (256 + window_bits - 1) / window_bits; c->T.ngens = NUM_GENS;
    size_t nent = (size_t)c->T.nwin * ((1u << window_bits) - 1u);
    c->table_bytes = (size_t)NUM_GENS * nent * 64;
    CUDA_OK(cudaMalloc(&c->d_tab, c->table_bytes));
    c->T.tab = c->d_tab;
    c->max_batch = max_batch;
    c->ws_words_per_proof = VL::WORDS > PL::WORDS ? VL::WORDS : PL::WORDS;
    CUDA_OK(cudaMalloc(&c->d_ws, c->ws_words_per_proof * max_batch * sizeof(uint32_t)));
    CUDA_OK(cudaMalloc(&c->d_in_a, (size_t)64 * max_batch));
    CUDA_OK(cudaMalloc(&c->d_in_b, (size_t)U64_PROOF_BYTES_AFFINE * max_batch));
    CUDA_OK(cudaMalloc(&c->d_in_c, (size_t)U64_RNG_BYTES * max_batch));
    CUDA_OK(cudaMalloc(&c->d_out, (size_t)U64_PROOF_BYTES_COMPRESSED * max_batch));
    CUDA_OK(cudaMalloc(&c->d_status, sizeof(int32_t) * max_batch));
    int rc = build_tables(c, gens, gen_id);
    if (rc != BPPP_OK) { bppp_ctx_destroy(c); return rc; }
    *out = c;
    return BPPP_OK;
}

extern "C" void bppp_ctx_destroy(bppp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_tab); cudaFree(c->d_ws); cudaFree(c->d_in_a); cudaFree(c->d_in_b); cudaFree(c->d_in_c);
    cudaFree(c->d_out); cudaFree(c->d_status);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int bppp_ctx_info(const bppp_ctx *c, size_t *table_bytes, size_t *workspace_bytes, double *table_build_ms, int *window_bits) {
    if (!c) return BPPP_ERR_ARG;
    if (table_bytes) *table_bytes = c->table_bytes;
    if (workspace_bytes) *workspace_bytes = c->ws_words_per_proof * c->max_batch * sizeof(uint32_t);
    if (table_build_ms) *table_build_ms = c->table_build_ms;
    if (window_bits) *window_bits = c->T.W;
    return BPPP_OK;
}
extern "C" uint64_t bppp_launch_count(const bppp_ctx *c) { return c ? c->launches : 0; }

static TermMap identity_map() { TermMap tm; for (int t = 0; t < NUM_GENS; t++) tm.gen[t] = t; return tm; }

// ---- verify ----
static int verify_slice(bppp_ctx *c, cudaStream_t st, size_t n, const uint8_t *d_commits, const uint8_t *d_proofs, int fmt,
                        const Merlin &init, int32_t *d_status) {
    WS w{c->d_ws, n};
    const unsigned g64 = nblocks(n, 64);
    LAUNCH(c, k_v_load, g64, 64, w, d_commits, d_proofs, fmt);
    launch_batch_inv(c, st, w, VL::VP + 20, VL::ZINV);
    LAUNCH(c, k_v_phase1, g64, 64, w, init);
    TermMap tm = identity_map();
    launch_msm_fixed(c, st, w, VL::FS, tm, 17, VL::ACC);      // pt = ps_tau g + <g_vec, pn_tau>  (circuit.rs:206)
    LAUNCH(c, k_v_var5, g64, 64, w);
    for (int j = 0; j < 4; j++) {
        launch_batch_inv(c, st, w, VL::COM + 20, VL::ZINV);
        LAUNCH(c, k_v_round, g64, 64, w, j);
        LAUNCH(c, k_v_var2, g64, 64, w, j);
    }
    LAUNCH(c, k_v_final_scalars, g64, 64, w);
    launch_msm_fixed(c, st, w, VL::FS, tm, NUM_GENS, VL::ACC);  // commit(l, n) over the original generators
    LAUNCH(c, k_v_verdict, g64, 64, w, d_status);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

extern "C" int bppp_u64_verify_batch_dev(bppp_ctx *c, size_t n, const void *d_commits, const void *d_proofs, int fmt,
                                         const uint8_t *label, size_t label_len, void *d_status, void *stream) {
    if (!c || (n && (!d_commits || !d_proofs || !d_status))) return fail(BPPP_ERR_ARG, "null argument");
    if (fmt != FMT_COMPRESSED && fmt != FMT_AFFINE64) return fail(BPPP_ERR_ARG, "bad point format");
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    Merlin init; merlin_init(init, label, (uint32_t)label_len);
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        int rc = verify_slice(c, st, m, (const uint8_t *)d_commits + csz * off, (const uint8_t *)d_proofs + psz * off, fmt, init,
                              (int32_t *)d_status + off);
        if (rc != BPPP_OK) return rc;
    }
    return BPPP_OK;
}

extern "C" int bppp_u64_verify_batch(bppp_ctx *c, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt,
                                     const uint8_t *label, size_t label_len, int32_t *status) {
    if (!c || (n && (!commits || !proofs || !status))) return fail(BPPP_ERR_ARG, "null argument");
    if (fmt != FMT_COMPRESSED && fmt != FMT_AFFINE64) return fail(BPPP_ERR_ARG, "bad point format");
    CUDA_OK(cudaSetDevice(c->device));
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    Merlin init; merlin_init(init, label, (uint32_t)label_len);
    cudaStream_t st = c->stream;
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        CUDA_OK(cudaMemcpyAsync(c->d_in_a, commits + csz * off, csz * m, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_in_b, proofs + psz * off, psz * m, cudaMemcpyHostToDevice, st));
        int rc = verify_slice(c, st, m, c->d_in_a, c->d_in_b, fmt, init, c->d_status);
        if (rc != BPPP_OK) return rc;
        CUDA_OK(cudaMemcpyAsync(status + off, c->d_status, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    return BPPP_OK;
}

// ---- commit ----
extern "C" int bppp_u64_commit_batch(bppp_ctx *c, size_t n, const uint64_t *x, const uint8_t *blinds32, int fmt, uint8_t *out) {
    if (!c || (n && (!x || !blinds32 || !out))) return fail(BPPP_ERR_ARG, "null argument");
    if (fmt != FMT_COMPRESSED && fmt != FMT_AFFINE64) return fail(BPPP_ERR_ARG, "bad point format");
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    size_t osz = fmt == FMT_COMPRESSED ? 33 : 64;
    TermMap tm = identity_map(); tm.gen[0] = GEN_G; tm.gen[1] = GEN_HVEC;
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        WS w{c->d_ws, m};
        CUDA_OK(cudaMemcpyAsync(c->d_in_a, x + off, 8 * m, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_in_b, blinds32 + 32 * off, 32 * m, cudaMemcpyHostToDevice, st));
        LAUNCH(c, k_c_load, nblocks(m, 64), 64, w, (const uint64_t *)c->d_in_a, c->d_in_b);
        launch_msm_fixed(c, st, w, VL::FS, tm, 2, VL::ACC);
        launch_batch_inv(c, st, w, VL::ACC + 20, VL::ZINV);
        LAUNCH(c, k_c_store, nblocks(m, 64), 64, w, c->d_out, fmt);
        CUDA_OK(cudaMemcpyAsync(out + osz * off, c->d_out, osz * m, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

// ---- prove ----
static int prove_slice(bppp_ctx *c, cudaStream_t st, size_t n, const uint64_t *d_x, const uint8_t *d_blinds, const uint8_t *d_rng,
                       const Merlin &init, uint8_t *d_proofs, int32_t *d_status) {
    WS w{c->d_ws, n};
    const unsigned g64 = nblocks(n, 64);
    TermMap tm;
    LAUNCH(c, k_p_load, g64, 64, w, d_x, d_blinds);
    // V = x g + s h_0  (reciprocal.rs:88-90)
    u64p_termmap_commit(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 2, PL::PTS + 30 * PP_V);
    launch_batch_inv(c, st, w, PL::PTS + 30 * PP_V + 20, PL::ZINV + 10 * PP_V);
    LAUNCH(c, k_p_phase1, g64, 64, w, init, d_rng);
    // r_com, c_o, c_l, c_r
    for (int k = 0; k < 4; k++) {
        int nterms = u64p_termmap_stage1(tm.gen, k);
        launch_msm_fixed(c, st, w, PL::FS + 8 * u64p_stage1_scalar_base(k), tm, nterms, PL::PTS + 30 * u64p_stage1_point(k));
    }
    LAUNCH(c, k_p_vprime, g64, 64, w);
    for (int k = 0; k < 5; k++) {
        int p = u64p_stage1_norm_point(k);
        launch_batch_inv(c, st, w, PL::PTS + 30 * p + 20, PL::ZINV + 10 * p);
    }
    LAUNCH(c, k_p_phase2, g64, 64, w, d_rng);
    u64p_termmap_cs(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 42, PL::PTS + 30 * PP_CS);
    launch_batch_inv(c, st, w, PL::PTS + 30 * PP_CS + 20, PL::ZINV + 10 * PP_CS);
    LAUNCH(c, k_p_phase3, g64, 64, w);
    // C_0 = v g + <h, l> + <g_vec, n>  (circuit.rs:522-524): 43 terms
    u64p_termmap_c0(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 43, PL::COM);
    for (int j = 0; j < 4; j++) {
        // X_j (49 terms), R_j (25 terms) over the original generators
        TermMap all = identity_map();
        launch_msm_fixed(c, st, w, PL::XS, all, NUM_GENS, PL::PTS + 30 * (PP_X + j));
        u64p_termmap_r(tm.gen, j);
        launch_msm_fixed(c, st, w, PL::RS, tm, 25, PL::PTS + 30 * (PP_R + j));
        launch_batch_inv(c, st, w, PL::COM + 20, PL::ZINV + 10 * PP_COM);
        launch_batch_inv(c, st, w, PL::PTS + 30 * (PP_X + j) + 20, PL::ZINV + 10 * (PP_X + j));
        launch_batch_inv(c, st, w, PL::PTS + 30 * (PP_R + j) + 20, PL::ZINV + 10 * (PP_R + j));
        LAUNCH(c, k_p_round, g64, 64, w, j);
        if (j < 3) LAUNCH(c, k_p_var2, g64, 64, w, j);
    }
    LAUNCH(c, k_p_output, g64, 64, w, d_proofs, d_status);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

extern "C" int bppp_u64_prove_batch_dev(bppp_ctx *c, size_t n, const void *d_x, const void *d_blinds32, const void *d_rng,
                                        const uint8_t *label, size_t label_len, void *d_proofs_out, void *d_status, void *stream) {
    if (!c || (n && (!d_x || !d_blinds32 || !d_rng || !d_proofs_out || !d_status))) return fail(BPPP_ERR_ARG, "null argument");
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    Merlin init; merlin_init(init, label, (uint32_t)label_len);
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        int rc = prove_slice(c, st, m, (const uint64_t *)d_x + off, (const uint8_t *)d_blinds32 + 32 * off,
                             (const uint8_t *)d_rng + (size_t)U64_RNG_BYTES * off, init,
                             (uint8_t *)d_proofs_out + (size_t)U64_PROOF_BYTES_COMPRESSED * off, (int32_t *)d_status + off);
        if (rc != BPPP_OK) return rc;
    }
    return BPPP_OK;
}

extern "C" int bppp_u64_prove_batch(bppp_ctx *c, size_t n, const uint64_t *x, const uint8_t *blinds32, const uint8_t *rng,
                                    const uint8_t *label, size_t label_len, uint8_t *proofs_out, int32_t *status) {
    if (!c || (n && (!x || !blinds32 || !rng || !proofs_out || !status))) return fail(BPPP_ERR_ARG, "null argument");
    CUDA_OK(cudaSetDevice(c->device));
    Merlin init; merlin_init(init, label, (uint32_t)label_len);
    cudaStream_t st = c->stream;
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        CUDA_OK(cudaMemcpyAsync(c->d_in_a, x + off, 8 * m, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_in_b, blinds32 + 32 * off, 32 * m, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_in_c, rng + (size_t)U64_RNG_BYTES * off, (size_t)U64_RNG_BYTES * m, cudaMemcpyHostToDevice, st));
        int rc = prove_slice(c, st, m, (const uint64_t *)c->d_in_a, c->d_in_b, c->d_in_c, init, c->d_out, c->d_status);
        if (rc != BPPP_OK) return rc;
        CUDA_OK(cudaMemcpyAsync(proofs_out + (size_t)U64_PROOF_BYTES_COMPRESSED * off, c->d_out, (size_t)U64_PROOF_BYTES_COMPRESSED * m, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(status + off, c->d_status, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    return BPPP_OK;
}

// ---- microbench ----
extern "C" int bppp_microbench(int device, double *out, int n_out) {
    if (!out || n_out < 8) return fail(BPPP_ERR_ARG, "need 8 outputs");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    int sms = prop.multiProcessorCount;
    uint64_t *d64 = nullptr; uint32_t *d32 = nullptr;
    const int blocks = sms * 8;
    CUDA_OK(cudaMalloc(&d64, sizeof(uint64_t) * blocks * 256));
    CUDA_OK(cudaMalloc(&d32, sizeof(uint32_t) * blocks * 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time_ms = [&](auto launch) -> float {
        launch(); cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        return best;
    };
    {
        const int iters = 2000;
        float ms = time_ms([&] { k_mb_imad<<<blocks, 256>>>(d64, 12345u, iters); });
        out[0] = (double)blocks * 256 * iters * 64 / (ms * 1e-3);
    }
    const int opblocks = sms * 16;
    auto run_op = [&](int op, int iters) -> double {
        float ms = 0;
        switch (op) {
            case 1: ms = time_ms([&] { k_mb_op<1><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 2: ms = time_ms([&] { k_mb_op<2><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 3: ms = time_ms([&] { k_mb_op<3><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 4: ms = time_ms([&] { k_mb_op<4><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 5: ms = time_ms([&] { k_mb_op<5><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            default: ms = time_ms([&] { k_mb_op<6><<<opblocks, 64>>>(d32, 777u, iters); }); break;
        }
        return (double)opblocks * 64 * iters / (ms * 1e-3);
    };
    out[1] = run_op(1, 4000); out[2] = run_op(2, 4000); out[3] = run_op(3, 2000);
    out[4] = run_op(4, 400); out[5] = run_op(5, 400); out[6] = run_op(6, 400);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, device);
    out[7] = clk / 1000.0;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d64); cudaFree(d32);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}
