// libbppp.so, core translation unit: context, fixed-base window tables, the fixed-base MSM and
// batch-inversion kernels every pipeline shares, and commit_value.
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
#ifndef BPPP_CORE_INLINE
#define BPPP_FE_NOINLINE 1   // see fe.cuh: call-based fe_mul keeps the MSM loop inside the instruction cache
#endif
#ifndef BPPP_CORE_PT_INLINE
#define BPPP_PT_NOINLINE 1   // ec.cuh: the XYZZ mixed addition of k_msm_fixed is one real function with its 10 products
#endif                       // inlined (measured -6 % on the kernel; the Jacobian ladders are faster with per-product calls)
#include "engine_common.cuh"

using namespace bppp;

static thread_local std::string g_last_error;
namespace bppp { int engine_fail(int code, const std::string &msg) { g_last_error = msg; return code; } }
static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

// sum_t scalar[t] * G_{gen[t]} for every proof: LANES threads per proof
template <int LANES>
#ifndef BPPP_MSM_BLOCK
#define BPPP_MSM_BLOCK 64
#endif
#ifndef BPPP_MSM_MINBLOCKS
#define BPPP_MSM_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(BPPP_MSM_BLOCK, BPPP_MSM_MINBLOCKS) k_msm_fixed(FixedTable T, WS w, int sc_off, TermMap tm, int nterms, int out_off) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = tid / LANES;
    int lane = (int)(tid % LANES);
    bool live = i < w.n;
    if (!live) i = w.n - 1;
    Pt acc = T.sgn ? msm_fixed_lane_signed(T, w, i, sc_off, tm.gen, nterms, lane, LANES) : msm_fixed_lane(T, w, i, sc_off, tm.gen, nterms, lane, LANES);
    acc = lanes_reduce<LANES>(acc);
    if (live && lane == 0) ws_st_pt(w, i, out_off, acc);
}

__global__ void __launch_bounds__(128) k_batch_inv(WS w, int in_off, int out_off, size_t nthreads) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) batch_inv_strided(w, in_off, out_off, t, nthreads, w.n);
}

__global__ void __launch_bounds__(128) k_batch_inv_list(WS w, InvList L, size_t nthreads) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) batch_inv_list_strided(w, L, t, nthreads);
}

__global__ void __launch_bounds__(64) k_emit_points(WS w, EmitList L, uint8_t *out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= w.n * (size_t)L.n) return;
    size_t i = t / L.n; int k = (int)(t % L.n);
    bool id;
    PtA a = ws_affine(w, i, L.pt[k], L.zinv[k], id);
    pta_compress(out + 33 * t, a, id);
}

// ---- commit ----
__global__ void __launch_bounds__(64) k_c_load(WS w, const uint64_t *xs, const uint8_t *blinds, int32_t *bad_flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    Sc s; int32_t st = ST_TRUE;
    // a k256::Scalar cannot hold a value >= n: the call fails (BPPP_ERR_ARG) instead of committing with a zero blinding
    if (!sc_from_be32(s, blinds + 32 * i)) { st = ST_BAD_SCALAR; s = sc_zero(); atomicOr(bad_flag, 1); }
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
// scratch layout per entry (word-major over nent = nwin * E entries): Pt at 0..PT_W-1, zinv after it
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
    PtA a = ws_affine(tmp, t, 0, PT_W, id);
    uint32_t x[8], y[8];
    fe_to_words(x, a.x); fe_to_words(y, a.y);
    if (id) {
#pragma unroll
        for (int k = 0; k < 8; k++) { x[k] = 0; y[k] = 0; }
    }
    dst[4 * t] = make_uint4(x[0], x[1], x[2], x[3]); dst[4 * t + 1] = make_uint4(x[4], x[5], x[6], x[7]);
    dst[4 * t + 2] = make_uint4(y[0], y[1], y[2], y[3]); dst[4 * t + 3] = make_uint4(y[4], y[5], y[6], y[7]);
}

#ifndef BPPP_MSM_LANES
#define BPPP_MSM_LANES 4
#endif
static constexpr int MSM_LANES = BPPP_MSM_LANES;

namespace bppp {
SubPlan plan_sub(bppp_ctx *c, size_t n, int kind) {
    SubPlan sp;
    int parts = c->profiling ? 1 : (kind == SUB_HOST ? c->nsub_host : kind == SUB_HOST_PROVE ? c->nsub_host_prove : c->nsub);          // per-kernel timing wants kernels back to back on one stream
    const size_t min_part = 2048;                    // below this a sub-batch cannot fill the GPU anyway
    while (parts > 1 && n / parts < min_part) parts--;
    sp.parts = parts;
    c->active_parts = parts;
    for (int k = 0; k <= parts; k++) sp.lo[k] = n * k / parts;
    return sp;
}
int fork_streams(bppp_ctx *c, cudaStream_t caller, const SubPlan &sp) {
    if (sp.parts == 1) return BPPP_OK;
    CUDA_OK(cudaEventRecord(c->ev_fork, caller));
    for (int k = 0; k < sp.parts; k++) CUDA_OK(cudaStreamWaitEvent(c->sub_stream[k], c->ev_fork, 0));
    return BPPP_OK;
}
int join_streams(bppp_ctx *c, cudaStream_t caller, const SubPlan &sp) {
    if (sp.parts == 1) return BPPP_OK;
    for (int k = 0; k < sp.parts; k++) {
        CUDA_OK(cudaEventRecord(c->ev_join[k], c->sub_stream[k]));
        CUDA_OK(cudaStreamWaitEvent(caller, c->ev_join[k], 0));
    }
    return BPPP_OK;
}
// lanes per proof: MSM_LANES when the batch alone fills the GPU, more for small (sub-)batches so that a rank holding a
// 1/8 share of a batch (strong scaling) still has ~14 warps per SM; BPPP_MSM_LANES_RT overrides (experiments)
int msm_lanes_for(const bppp_ctx *c, size_t n) {
    if (c->msm_lanes_override) return c->msm_lanes_override;
    const size_t full = (size_t)c->sm_count * 448;     // threads of one full wave at 7 blocks x 64
    n *= (size_t)c->active_parts * (size_t)c->inflight_hint;
    // measured on a B200 (tools/batch_sweep.py --lane-sweep, profiles/r2_lane_sweep.json): 8 lanes pay below ~16k proofs
    // in flight, 16 below ~4k
    if (n * 16 <= full) return 16;
    if (n * 4 <= full) return 8;
    return MSM_LANES;
}
void launch_msm_fixed(bppp_ctx *c, cudaStream_t st, WS w, int sc_off, const TermMap &tm, int nterms, int out_off) {
    int lanes = msm_lanes_for(c, w.n);
    while (lanes > MSM_LANES && nterms * c->T.nwin < 32 * lanes) lanes /= 2;     // short sums: the lane reduction (log2 lanes full additions) must stay small
    size_t threads = w.n * lanes;
    if (lanes >= 16) LAUNCH(c, k_msm_fixed<16>, nblocks(threads, BPPP_MSM_BLOCK), BPPP_MSM_BLOCK, c->T, w, sc_off, tm, nterms, out_off);
    else if (lanes == 8) LAUNCH(c, k_msm_fixed<8>, nblocks(threads, BPPP_MSM_BLOCK), BPPP_MSM_BLOCK, c->T, w, sc_off, tm, nterms, out_off);
    else LAUNCH(c, k_msm_fixed<MSM_LANES>, nblocks(threads, BPPP_MSM_BLOCK), BPPP_MSM_BLOCK, c->T, w, sc_off, tm, nterms, out_off);
}
void launch_batch_inv_list(bppp_ctx *c, cudaStream_t st, WS w, const InvList &L) {
    size_t items = (size_t)L.n * w.n, per = 8;
    size_t nthreads = (items + per - 1) / per;
    size_t min_threads = (size_t)c->sm_count * 128;
    if (nthreads < min_threads) nthreads = items < min_threads ? items : min_threads;
    LAUNCH(c, k_batch_inv_list, nblocks(nthreads, 128), 128, w, L, nthreads);
}
void launch_emit_points(bppp_ctx *c, cudaStream_t st, WS w, const EmitList &L, uint8_t *d_out) {
    LAUNCH(c, k_emit_points, nblocks(w.n * L.n, 64), 64, w, L, d_out);
}
void launch_batch_inv(bppp_ctx *c, cudaStream_t st, WS w, int in_off, int out_off) {
    // one inversion per thread, >= 8 items per thread when the batch is large enough to still fill the GPU
    size_t per = 8;
    size_t nthreads = (w.n + per - 1) / per;
    size_t min_threads = (size_t)c->sm_count * 128;
    if (nthreads < min_threads) nthreads = w.n < min_threads ? w.n : min_threads;
    LAUNCH(c, k_batch_inv, nblocks(nthreads, 128), 128, w, in_off, out_off, nthreads);
}
}  // namespace bppp

static int build_tables(bppp_ctx *c, const PtA *gens, const bool *gen_id) {
    const int W = c->T.W, nwin = c->T.nwin;
    const uint32_t E = c->T.E;
    const size_t nent = (size_t)nwin * E;
    uint32_t *d_tmp = nullptr;
    CUDA_OK(cudaMalloc(&d_tmp, nent * (size_t)(4 * FE_W) * sizeof(uint32_t)));
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
        launch_batch_inv(c, st, tmp, 2 * FE_W, PT_W);
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

// streams, events, workspace and staging buffers of a context for `max_batch` proofs per launch sequence
static int alloc_work(bppp_ctx *c, size_t max_batch) {
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (const char *e = getenv("BPPP_NSUB")) { int v = atoi(e); if (v >= 1 && v <= bppp_ctx::MAX_SUB) c->nsub = v; }
    if (const char *e = getenv("BPPP_NSUB_HOST")) { int v = atoi(e); if (v >= 1 && v <= bppp_ctx::MAX_SUB) c->nsub_host = v; }
    if (const char *e = getenv("BPPP_NSUB_HOST_PROVE")) { int v = atoi(e); if (v >= 1 && v <= bppp_ctx::MAX_SUB) c->nsub_host_prove = v; }
    if (const char *e = getenv("BPPP_TAB_AFFINE")) c->tab_affine = atoi(e) != 0;
    if (const char *e = getenv("BPPP_TAB_K")) {     // "k1,k2,k3": items per thread (= per inversion) of the three levels
        int a = 0, b = 0, d = 0;
        if (sscanf(e, "%d,%d,%d", &a, &b, &d) == 3 && a >= 1 && b >= 1 && d >= 1) { c->tab_k[0] = a; c->tab_k[1] = b; c->tab_k[2] = d; }
    }
    if (const char *e = getenv("BPPP_VAR_SEG")) { int v = atoi(e); if (v >= 1 && v <= 33) c->var_seg = v; }
    if (const char *e = getenv("BPPP_VAR_SEG_ONE")) c->var_seg_one_item = atoi(e) != 0;
    if (const char *e = getenv("BPPP_VAR_SEG_WARPS")) { int v = atoi(e); if (v >= 1) c->var_seg_warps = v; }
    if (const char *e = getenv("BPPP_MSM_LANES_RT")) { int v = atoi(e); if (v == 4 || v == 8 || v == 16) c->msm_lanes_override = v; }
    if (const char *e = getenv("BPPP_VAR_LANES_RT")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) c->var_lanes_override = v; }
    for (int k = 0; k < bppp_ctx::MAX_SUB; k++) {
        CUDA_OK(cudaStreamCreateWithFlags(&c->sub_stream[k], cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2 * bppp_ctx::MAX_SUB; k++) CUDA_OK(cudaEventCreateWithFlags(&c->ev_up[k], cudaEventDisableTiming));
    c->max_batch = max_batch;
    c->ws_words_per_proof = VL::WORDS > PL::WORDS ? VL::WORDS : PL::WORDS;
    c->ws_words_per_proof = (c->ws_words_per_proof + 3) & ~(size_t)3;      // sub-batch bases stay 16-byte aligned (vtab_entry)
    CUDA_OK(cudaMalloc(&c->d_ws, c->ws_words_per_proof * max_batch * sizeof(uint32_t)));
    CUDA_OK(cudaMalloc(&c->d_in_a, (size_t)64 * max_batch));
    CUDA_OK(cudaMalloc(&c->d_in_b, (size_t)U64_PROOF_BYTES_AFFINE * max_batch));
    CUDA_OK(cudaMalloc(&c->d_in_c, (size_t)U64_RNG_BYTES * max_batch));
    CUDA_OK(cudaMalloc(&c->d_out, (size_t)U64_PROOF_BYTES_COMPRESSED * max_batch));
    CUDA_OK(cudaMalloc(&c->d_status, sizeof(int32_t) * max_batch));
    CUDA_OK(cudaMalloc(&c->d_flag, sizeof(int32_t)));
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
    // 2..20: unsigned windows; 21..23: signed windows (2^(W-1) entries each); a negative value asks for signed windows of
    // |window_bits| bits at any size (tests)
    const bool signed_windows = window_bits < 0 || window_bits > 20;
    if (window_bits < 0) window_bits = -window_bits;
    if (window_bits < 2 || window_bits > 23) return fail(BPPP_ERR_ARG, "window_bits must be in 2..23 (or negative for signed windows)");
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
    struct Guard { bppp_ctx *c; ~Guard() { if (c) bppp_ctx_destroy(c); } } guard{c};      // any early return below frees what was allocated
    c->device = device;
    fixed_table_shape(c->T, window_bits, signed_windows); c->T.ngens = NUM_GENS;
    size_t nent = (size_t)c->T.nwin * c->T.E;
    c->table_bytes = (size_t)NUM_GENS * nent * 64;
    CUDA_OK(cudaMalloc(&c->d_tab, c->table_bytes));
    c->owns_tab = true;
    c->T.tab = c->d_tab;
    int rc = alloc_work(c, max_batch);
    if (rc != BPPP_OK) return rc;
    rc = build_tables(c, gens, gen_id);
    if (rc != BPPP_OK) return rc;
    guard.c = nullptr;
    *out = c;
    return BPPP_OK;
}

// A second context on the same GPU that shares the parent's window tables (read-only after construction) and owns its own
// workspace, staging buffers and streams: independent batches can then be in flight side by side (one host thread per
// context), which is what keeps the GPU full when each batch is small.  The parent must outlive it.
extern "C" int bppp_ctx_create_shared(bppp_ctx **out, const bppp_ctx *parent, size_t max_batch) {
    if (!out || !parent) return fail(BPPP_ERR_ARG, "null argument");
    *out = nullptr;
    CUDA_OK(cudaSetDevice(parent->device));
    bppp_ctx *c = new bppp_ctx();
    struct Guard { bppp_ctx *c; ~Guard() { if (c) bppp_ctx_destroy(c); } } guard{c};
    c->device = parent->device;
    c->T = parent->T; c->d_tab = parent->d_tab; c->owns_tab = false;
    c->table_bytes = parent->table_bytes; c->table_build_ms = 0;
    int rc = alloc_work(c, max_batch ? max_batch : parent->max_batch);
    if (rc != BPPP_OK) return rc;
    guard.c = nullptr;
    *out = c;
    return BPPP_OK;
}

extern "C" void bppp_ctx_destroy(bppp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->owns_tab) cudaFree(c->d_tab);
    cudaFree(c->d_ws); cudaFree(c->d_in_a); cudaFree(c->d_in_b); cudaFree(c->d_in_c);
    cudaFree(c->d_out); cudaFree(c->d_status); cudaFree(c->d_flag);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int k = 0; k < bppp_ctx::MAX_SUB; k++) {
        if (c->sub_stream[k]) cudaStreamDestroy(c->sub_stream[k]);
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int k = 0; k < 2 * bppp_ctx::MAX_SUB; k++) if (c->ev_up[k]) cudaEventDestroy(c->ev_up[k]);
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
extern "C" int bppp_ctx_set_inflight(bppp_ctx *c, int batches) {
    if (!c || batches < 1 || batches > 64) return fail(BPPP_ERR_ARG, "batches in flight must be in 1..64");
    c->inflight_hint = batches;
    return BPPP_OK;
}

extern "C" int bppp_ctx_profile_begin(bppp_ctx *c) {
    if (!c) return BPPP_ERR_ARG;
    c->prof.clear();
    c->profiling = true;
    return BPPP_OK;
}
// Aggregates the launches recorded since profile_begin by kernel name.  names: n_max slots of 48 bytes.
extern "C" int bppp_ctx_profile_end(bppp_ctx *c, char *names, double *total_ms, uint32_t *counts, int n_max, int *n_out) {
    if (!c || !names || !total_ms || !counts || !n_out) return BPPP_ERR_ARG;
    c->profiling = false;
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaDeviceSynchronize());
    int n = 0;
    for (auto &r : c->prof) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        int k = 0;
        for (; k < n; k++) if (strncmp(names + 48 * k, r.name, 47) == 0) break;
        if (k == n) {
            if (n >= n_max) continue;
            strncpy(names + 48 * k, r.name, 47); names[48 * k + 47] = 0;
            total_ms[k] = 0; counts[k] = 0; n++;
        }
        total_ms[k] += ms; counts[k]++;
    }
    c->prof.clear();
    *n_out = n;
    return BPPP_OK;
}


// ---- commit ----
extern "C" int bppp_u64_commit_batch(bppp_ctx *c, size_t n, const uint64_t *x, const uint8_t *blinds32, int fmt, uint8_t *out) {
    if (!c || (n && (!x || !blinds32 || !out))) return fail(BPPP_ERR_ARG, "null argument");
    if (fmt != FMT_COMPRESSED && fmt != FMT_AFFINE64) return fail(BPPP_ERR_ARG, "bad point format");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    size_t osz = fmt == FMT_COMPRESSED ? 33 : 64;
    TermMap tm = identity_map(); tm.gen[0] = GEN_G; tm.gen[1] = GEN_HVEC;
    int32_t bad = 0;
    CUDA_OK(cudaMemsetAsync(c->d_flag, 0, sizeof(int32_t), st));
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        WS w{c->d_ws, m};
        CUDA_OK(cudaMemcpyAsync(c->d_in_a, x + off, 8 * m, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_in_b, blinds32 + 32 * off, 32 * m, cudaMemcpyHostToDevice, st));
        LAUNCH(c, k_c_load, nblocks(m, 64), 64, w, (const uint64_t *)c->d_in_a, c->d_in_b, c->d_flag);
        launch_msm_fixed(c, st, w, VL::FS, tm, 2, VL::ACC);
        launch_batch_inv(c, st, w, VL::ACC + 2 * FE_W, VL::ZINV);
        LAUNCH(c, k_c_store, nblocks(m, 64), 64, w, c->d_out, fmt);
        CUDA_OK(cudaMemcpyAsync(out + osz * off, c->d_out, osz * m, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(&bad, c->d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        if (bad) return fail(BPPP_ERR_ARG, "commit: a blinding factor is not a canonical scalar (>= n); no commitment is returned for such input");
    }
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

