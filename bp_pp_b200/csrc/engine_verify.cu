// libbppp.so, verify translation unit: U64RangeProofProtocol::verify over a batch (u64_proof.rs:42-54).
#define BPPP_FE_NOINLINE 1   // phase kernels are not hot: call-based fe_mul keeps them small and quick to compile
#include "engine_common.cuh"

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

// phase 0a: one thread per (point, proof), point-major.
// flags[i]: bits 0..13 identity mask, bit 31 = a point failed to decode (merged with atomicOr)
__global__ void __launch_bounds__(128) k_v_decode(WS w, const uint8_t *commits, const uint8_t *proofs, int fmt, uint32_t *flags, size_t lo, size_t cnt) {
    // proofs lo .. lo + cnt of the part (the host-buffer entry point decodes a part chunk by chunk while the next chunk uploads)
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = (int)(t / cnt); size_t i = lo + (t - (size_t)k * cnt);       // point-major: coalesced stores of the decoded words
    if (k >= VP_COUNT) return;
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    uint32_t bit; bool bad;
    u64v_decode_point_one(w, i, k, commits + csz * i, proofs + psz * i, fmt, &bit, &bad);
    uint32_t f = bit | (bad ? 0x80000000u : 0u);
    if (f) atomicOr(flags + i, f);
}
__global__ void __launch_bounds__(64) k_v_load_finish(WS w, const uint8_t *proofs, int fmt, const uint32_t *flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    size_t psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    uint32_t f = flags[i];
    u64v_load_finish_one(w, i, proofs + psz * i, fmt, f & 0x7FFFFFFFu, (f >> 31) != 0);
}
// ladder tables: build (one thread per (proof, point)), then batch-invert every Z and normalise in place
__global__ void __launch_bounds__(64) k_v_tables_build(WS w) {
    // point-major thread order: a warp handles one table point of 32 consecutive proofs, so its word-major loads and
    // stores are contiguous (proof-major order put 16 different points, i.e. 16 different rows, into every access)
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int p = (int)(t / w.n); size_t i = t - (size_t)p * w.n;
    if (p < VL::TAB_POINTS) u64v_table_build_one(w, i, p);
}
__global__ void __launch_bounds__(128) k_v_tables_normalize(WS w, size_t nthreads) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) u64v_tables_normalize_strided(w, t, nthreads);
}
// ext != nullptr: challenges of a caller-owned transcript, ext_stride bytes per proof (ws.cuh: Tx)
__global__ void __launch_bounds__(64) k_v_phase1(WS w, Merlin init, const uint8_t *ext, int ext_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_phase1_one(w, i, init, ext ? ext + (size_t)ext_stride * i : nullptr);
}
__global__ void __launch_bounds__(64) k_v_round(WS w, int j, const uint8_t *ext, int ext_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_round_one(w, i, j, ext ? ext + (size_t)ext_stride * i : nullptr);
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

// the same tables in affine coordinates, one launch per dependent level (u64_verify.cuh:tables_affine_level)
__global__ void __launch_bounds__(128) k_v_tables_affine(WS w, int level, size_t nthreads) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) u64v_tables_affine_level(w, level, t, nthreads);
}
// Ladder tables of the 13 per-proof points: three affine levels (~60 M per point against 148 M for the projective build +
// normalisation pass, kept behind BPPP_TAB_AFFINE=0), each level one cross-proof inversion per thread.
static int launch_v_tables(bppp_ctx *c, cudaStream_t st, WS w) {
    const size_t n = w.n;
    if (c->tab_affine) {
        for (int level = 1; level <= 3; level++) {
            const size_t nthreads = tab_level_threads(c, n * VL::TAB_POINTS * (size_t)aff_level_nops(level), level);
            LAUNCH(c, k_v_tables_affine, nblocks(nthreads, 128), 128, w, level, nthreads);
        }
    } else {
        LAUNCH(c, k_v_tables_build, nblocks(n * VL::TAB_POINTS, 64), 64, w);
        size_t items = n * VL::TAB_ENTRIES, nthreads = (items + 63) / 64;
        LAUNCH(c, k_v_tables_normalize, nblocks(nthreads, 128), 128, w, nthreads);
    }
    return BPPP_OK;
}

// ---- verify ----
// decoded: the caller has already run k_v_decode over the whole part (chunk by chunk, bppp_u64_verify_batch)
static int verify_part(bppp_ctx *c, cudaStream_t st, WS w, const uint8_t *d_commits, const uint8_t *d_proofs, int fmt,
                       const Merlin &init, int32_t *d_status, bool decoded = false) {
    const size_t n = w.n;
    const unsigned g64 = nblocks(n, 64);
    // decode flags live in the workspace's IDMASK row until k_v_load_finish rewrites it
    uint32_t *flags = w.p + (size_t)VL::IDMASK * w.n;
    if (!decoded) {
        CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * n, st));
        LAUNCH(c, k_v_decode, nblocks(n * VP_COUNT, 128), 128, w, d_commits, d_proofs, fmt, flags, (size_t)0, n);
    }
    LAUNCH(c, k_v_load_finish, g64, 64, w, d_proofs, fmt, flags);
    launch_batch_inv(c, st, w, VL::VP + 2 * FE_W, VL::ZINV);
    LAUNCH(c, k_v_phase1, g64, 64, w, init, (const uint8_t *)nullptr, 0);
    launch_v_tables(c, st, w);      // affine 1P..8P tables of the 13 per-proof points
    TermMap tm = identity_map();
    launch_msm_fixed(c, st, w, VL::FS, tm, 17, VL::ACC);      // pt = ps_tau g + <g_vec, pn_tau>  (circuit.rs:206)
    launch_v_var5(c, st, w);
    for (int j = 0; j < 4; j++) {
        launch_batch_inv(c, st, w, VL::COM + 2 * FE_W, VL::ZINV);
        LAUNCH(c, k_v_round, g64, 64, w, j, (const uint8_t *)nullptr, 0);
        launch_v_var2(c, st, w, j);
    }
    LAUNCH(c, k_v_final_scalars, g64, 64, w);
    launch_msm_fixed(c, st, w, VL::FS, tm, NUM_GENS, VL::ACC);  // commit(l, n) over the original generators
    LAUNCH(c, k_v_verdict, g64, 64, w, d_status);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

static int verify_slice(bppp_ctx *c, cudaStream_t st, size_t n, const uint8_t *d_commits, const uint8_t *d_proofs, int fmt,
                        const Merlin &init, int32_t *d_status) {
    const size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    SubPlan sp = plan_sub(c, n);
    if (verify_one_part(c, n)) { sp.parts = 1; sp.lo[0] = 0; sp.lo[1] = n; c->active_parts = 1; }
    int rc = fork_streams(c, st, sp);
    if (rc != BPPP_OK) return rc;
    for (int k = 0; k < sp.parts; k++) {
        cudaStream_t s = sp.parts == 1 ? st : c->sub_stream[k];
        rc = verify_part(c, s, sub_ws(c, sp, k), d_commits + csz * sp.lo[k], d_proofs + psz * sp.lo[k], fmt, init, d_status + sp.lo[k]);
        if (rc != BPPP_OK) return rc;
    }
    return join_streams(c, st, sp);
}

extern "C" int bppp_u64_verify_batch_dev(bppp_ctx *c, size_t n, const void *d_commits, const void *d_proofs, int fmt,
                                         const uint8_t *label, size_t label_len, void *d_status, void *stream) {
    if (!c || (n && (!d_commits || !d_proofs || !d_status))) return fail(BPPP_ERR_ARG, "null argument");
    if (fmt != FMT_COMPRESSED && fmt != FMT_AFFINE64) return fail(BPPP_ERR_ARG, "bad point format");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream, as in the CUDA runtime
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
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    Merlin init; merlin_init(init, label, (uint32_t)label_len);
    // each sub-batch copies in, runs its kernel sequence and copies out on its own stream, so the transfers of one
    // sub-batch overlap with the kernels of the others (pinned host buffers make the copies truly asynchronous)
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        SubPlan sp = plan_sub(c, m, SUB_HOST);
        if (verify_one_part(c, m)) {
            // One part (segmented ladders want the GPU to themselves).  The records upload in four chunks on the copy stream and
            // each chunk is decoded as it lands (the square roots of a chunk take longer than the next chunk's transfer), so only
            // the first chunk's upload is exposed; everything after the decode runs over the whole part.
            sp.parts = 1; sp.lo[0] = 0; sp.lo[1] = m; c->active_parts = 1;
            cudaStream_t st = c->stream, cs = c->copy_stream;
            WS w = sub_ws(c, sp, 0);
            uint32_t *flags = w.p + (size_t)VL::IDMASK * w.n;
            CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * m, st));
            const int chunks = 4;
            for (int q = 0; q < chunks; q++) {
                size_t lo = m * q / chunks, cnt = m * (q + 1) / chunks - lo;
                CUDA_OK(cudaMemcpyAsync(c->d_in_a + csz * lo, commits + csz * (off + lo), csz * cnt, cudaMemcpyHostToDevice, cs));
                CUDA_OK(cudaMemcpyAsync(c->d_in_b + psz * lo, proofs + psz * (off + lo), psz * cnt, cudaMemcpyHostToDevice, cs));
                CUDA_OK(cudaEventRecord(c->ev_up[q], cs));
                CUDA_OK(cudaStreamWaitEvent(st, c->ev_up[q], 0));
                LAUNCH(c, k_v_decode, nblocks(cnt * VP_COUNT, 128), 128, w, c->d_in_a, c->d_in_b, fmt, flags, lo, cnt);
            }
            int rc = verify_part(c, st, w, c->d_in_a, c->d_in_b, fmt, init, c->d_status, true);
            if (rc != BPPP_OK) { cudaStreamSynchronize(cs); return rc; }
            CUDA_OK(cudaMemcpyAsync(status + off, c->d_status, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
            CUDA_OK(cudaStreamSynchronize(cs));
            CUDA_OK(cudaStreamSynchronize(st));
            continue;
        }
        for (int k = 0; k < sp.parts; k++) {
            cudaStream_t st = sp.parts == 1 ? c->stream : c->sub_stream[k];
            size_t lo = sp.lo[k], cnt = sp.lo[k + 1] - sp.lo[k];
            CUDA_OK(cudaMemcpyAsync(c->d_in_a + csz * lo, commits + csz * (off + lo), csz * cnt, cudaMemcpyHostToDevice, st));
            CUDA_OK(cudaMemcpyAsync(c->d_in_b + psz * lo, proofs + psz * (off + lo), psz * cnt, cudaMemcpyHostToDevice, st));
            int rc = verify_part(c, st, sub_ws(c, sp, k), c->d_in_a + csz * lo, c->d_in_b + psz * lo, fmt, init, c->d_status + lo);
            if (rc != BPPP_OK) return rc;
            CUDA_OK(cudaMemcpyAsync(status + off + lo, c->d_status + lo, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, st));
        }
        for (int k = 0; k < sp.parts; k++) CUDA_OK(cudaStreamSynchronize(sp.parts == 1 ? c->stream : c->sub_stream[k]));
    }
    return BPPP_OK;
}


// ---- phase-stepped verify for a caller-owned transcript (include/bppp.h) -------------------------------------------
// The same kernels as verify_part, cut at the transcript's challenge points: every step returns the compressed points
// the host has to append (those it does not already hold) and takes the challenges it drew.
static int step_check(bppp_ctx *c, int kind, int stage, const char *what) {
    if (!c) return fail(BPPP_ERR_ARG, "null context");
    if (c->step.kind != kind || c->step.stage != stage) return fail(BPPP_ERR_ARG, std::string(what) + ": called out of order for this context's stepped session");
    return BPPP_OK;
}
extern "C" int bppp_u64_verify_begin(bppp_ctx *c, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt, uint8_t *vprime33_out) {
    if (!c || !n || !commits || !proofs || !vprime33_out) return fail(BPPP_ERR_ARG, "null argument");
    if (fmt != FMT_COMPRESSED && fmt != FMT_AFFINE64) return fail(BPPP_ERR_ARG, "bad point format");
    if (n > c->max_batch) return fail(BPPP_ERR_ARG, "a stepped session holds at most max_batch proofs");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->step = {}; c->active_parts = 1;
    const size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    WS w{c->d_ws, n};
    uint32_t *flags = w.p + (size_t)VL::IDMASK * w.n;
    CUDA_OK(cudaMemcpyAsync(c->d_in_a, commits, csz * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(c->d_in_b, proofs, psz * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * n, st));
    LAUNCH(c, k_v_decode, nblocks(n * VP_COUNT, 128), 128, w, c->d_in_a, c->d_in_b, fmt, flags, (size_t)0, n);
    LAUNCH(c, k_v_load_finish, nblocks(n, 64), 64, w, c->d_in_b, fmt, flags);
    launch_batch_inv(c, st, w, VL::VP + 2 * FE_W, VL::ZINV);
    EmitList L; L.n = 1; L.pt[0] = VL::VP; L.zinv[0] = VL::ZINV;     // V' = V + r: "commitment_v" (circuit.rs:159)
    launch_emit_points(c, st, w, L, c->d_out);
    CUDA_OK(cudaMemcpyAsync(vprime33_out, c->d_out, 33 * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    c->step.kind = 1; c->step.n = n; c->step.stage = 1; c->step.fmt = fmt;
    return BPPP_OK;
}
extern "C" int bppp_u64_verify_circuit(bppp_ctx *c, const uint8_t *chal, uint8_t *com33_out) {
    int rc = step_check(c, 1, 1, "bppp_u64_verify_circuit");
    if (rc != BPPP_OK) return rc;
    if (!chal || !com33_out) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    Merlin unused{};
    CUDA_OK(cudaMemcpyAsync(c->d_in_c, chal, 192 * n, cudaMemcpyHostToDevice, st));
    LAUNCH(c, k_v_phase1, nblocks(n, 64), 64, w, unused, (const uint8_t *)c->d_in_c, 192);
    launch_v_tables(c, st, w);
    TermMap tm = identity_map();
    launch_msm_fixed(c, st, w, VL::FS, tm, 17, VL::ACC);
    launch_v_var5(c, st, w);
    launch_batch_inv(c, st, w, VL::COM + 2 * FE_W, VL::ZINV);
    EmitList L; L.n = 1; L.pt[0] = VL::COM; L.zinv[0] = VL::ZINV;    // "wnla_com" of round 0 (wnla.rs:88)
    launch_emit_points(c, st, w, L, c->d_out);
    CUDA_OK(cudaMemcpyAsync(com33_out, c->d_out, 33 * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    c->step.stage = 2;
    return BPPP_OK;
}
extern "C" int bppp_u64_verify_round(bppp_ctx *c, int j, const uint8_t *y32, uint8_t *com33_out) {
    if (j < 0 || j > 3) return fail(BPPP_ERR_ARG, "round index out of range");
    int rc = step_check(c, 1, 2 + j, "bppp_u64_verify_round");
    if (rc != BPPP_OK) return rc;
    if (!y32 || (j < 3 && !com33_out)) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    CUDA_OK(cudaMemcpyAsync(c->d_in_c, y32, 32 * n, cudaMemcpyHostToDevice, st));
    LAUNCH(c, k_v_round, nblocks(n, 64), 64, w, j, (const uint8_t *)c->d_in_c, 32);
    launch_v_var2(c, st, w, j);
    if (j < 3) {
        launch_batch_inv(c, st, w, VL::COM + 2 * FE_W, VL::ZINV);
        EmitList L; L.n = 1; L.pt[0] = VL::COM; L.zinv[0] = VL::ZINV;
        launch_emit_points(c, st, w, L, c->d_out);
        CUDA_OK(cudaMemcpyAsync(com33_out, c->d_out, 33 * n, cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    c->step.stage = 3 + j;
    return BPPP_OK;
}
extern "C" int bppp_u64_verify_finish(bppp_ctx *c, int32_t *status) {
    int rc = step_check(c, 1, 6, "bppp_u64_verify_finish");
    if (rc != BPPP_OK) return rc;
    if (!status) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    TermMap tm = identity_map();
    LAUNCH(c, k_v_final_scalars, nblocks(n, 64), 64, w);
    launch_msm_fixed(c, st, w, VL::FS, tm, NUM_GENS, VL::ACC);
    LAUNCH(c, k_v_verdict, nblocks(n, 64), 64, w, c->d_status);
    CUDA_OK(cudaMemcpyAsync(status, c->d_status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    c->step = {};
    return BPPP_OK;
}
extern "C" void bppp_u64_step_abort(bppp_ctx *c) { if (c) { std::lock_guard<std::mutex> lock(c->mu); c->step = {}; } }
