// libbppp.so, prove translation unit: U64RangeProofProtocol::prove over a batch (u64_proof.rs:57-82).
#define BPPP_FE_NOINLINE 1   // phase kernels are not hot: call-based fe_mul keeps them small and quick to compile
#include "engine_common.cuh"

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

// ---- prove kernels ----
__global__ void __launch_bounds__(64) k_p_load(WS w, const uint64_t *xs, const uint8_t *blinds) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_load_one(w, i, xs[i], blinds + 32 * i);
}
// ext != nullptr: challenges of a caller-owned transcript, ext_stride bytes per proof (ws.cuh: Tx)
__global__ void __launch_bounds__(64) k_p_phase1(WS w, Merlin init, const uint8_t *rng, const uint8_t *ext, int ext_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_phase1_one(w, i, init, rng + (size_t)U64_RNG_BYTES * i, ext ? ext + (size_t)ext_stride * i : nullptr);
}
__global__ void __launch_bounds__(64) k_p_phase2(WS w, const uint8_t *rng, const uint8_t *ext, int ext_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_phase2_one(w, i, rng + (size_t)U64_RNG_BYTES * i, ext ? ext + (size_t)ext_stride * i : nullptr);
}
__global__ void __launch_bounds__(64) k_p_phase3(WS w, const uint8_t *ext, int ext_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_phase3_one(w, i, ext ? ext + (size_t)ext_stride * i : nullptr);
}
__global__ void __launch_bounds__(64) k_p_round(WS w, int j, const uint8_t *ext, int ext_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_round_one(w, i, j, ext ? ext + (size_t)ext_stride * i : nullptr);
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

// ---- prove ----
// rng_late: when set, the event after which the RNG bytes of the second transcript phase (scalars 19..51 of each record) are
// in place -- the host-buffer entry point uploads them while the first-stage sums run
static int prove_part(bppp_ctx *c, cudaStream_t st, WS w, const uint64_t *d_x, const uint8_t *d_blinds, const uint8_t *d_rng,
                      const Merlin &init, uint8_t *d_proofs, int32_t *d_status, cudaEvent_t rng_late = nullptr) {
    const size_t n = w.n;
    const unsigned g64 = nblocks(n, 64);
    TermMap tm;
    LAUNCH(c, k_p_load, g64, 64, w, d_x, d_blinds);
    // V = x g + s h_0  (reciprocal.rs:88-90)
    u64p_termmap_commit(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 2, PL::PTS + PT_W * PP_V);
    launch_batch_inv(c, st, w, PL::PTS + PT_W * PP_V + 2 * FE_W, PL::ZINV + FE_W * PP_V);
    LAUNCH(c, k_p_phase1, g64, 64, w, init, d_rng, (const uint8_t *)nullptr, 0);
    // r_com, c_o, c_l, c_r
    for (int k = 0; k < 4; k++) {
        int nterms = u64p_termmap_stage1(tm.gen, k);
        launch_msm_fixed(c, st, w, PL::FS + 8 * u64p_stage1_scalar_base(k), tm, nterms, PL::PTS + PT_W * u64p_stage1_point(k));
    }
    LAUNCH(c, k_p_vprime, g64, 64, w);
    {   // the five first-stage normalisations in one launch
        InvList L; L.n = 5;
        for (int k = 0; k < 5; k++) { int p = u64p_stage1_norm_point(k); L.in[k] = PL::PTS + PT_W * p + 2 * FE_W; L.out[k] = PL::ZINV + FE_W * p; }
        launch_batch_inv_list(c, st, w, L);
    }
    if (rng_late) CUDA_OK(cudaStreamWaitEvent(st, rng_late, 0));
    LAUNCH(c, k_p_phase2, g64, 64, w, d_rng, (const uint8_t *)nullptr, 0);
    u64p_termmap_cs(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 42, PL::PTS + PT_W * PP_CS);
    launch_batch_inv(c, st, w, PL::PTS + PT_W * PP_CS + 2 * FE_W, PL::ZINV + FE_W * PP_CS);
    LAUNCH(c, k_p_phase3, g64, 64, w, (const uint8_t *)nullptr, 0);
    // C_0 = v g + <h, l> + <g_vec, n>  (circuit.rs:522-524): 43 terms
    u64p_termmap_c0(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 43, PL::COM);
    for (int j = 0; j < 4; j++) {
        // X_j (49 terms), R_j (25 terms) over the original generators
        TermMap all = identity_map();
        launch_msm_fixed(c, st, w, PL::XS, all, NUM_GENS, PL::PTS + PT_W * (PP_X + j));
        u64p_termmap_r(tm.gen, j);
        launch_msm_fixed(c, st, w, PL::RS, tm, 25, PL::PTS + PT_W * (PP_R + j));
        {   // com_j, X_j, R_j normalised together for the transcript
            InvList L; L.n = 3;
            L.in[0] = PL::COM + 2 * FE_W; L.out[0] = PL::ZINV + FE_W * PP_COM;
            L.in[1] = PL::PTS + PT_W * (PP_X + j) + 2 * FE_W; L.out[1] = PL::ZINV + FE_W * (PP_X + j);
            L.in[2] = PL::PTS + PT_W * (PP_R + j) + 2 * FE_W; L.out[2] = PL::ZINV + FE_W * (PP_R + j);
            launch_batch_inv_list(c, st, w, L);
        }
        LAUNCH(c, k_p_round, g64, 64, w, j, (const uint8_t *)nullptr, 0);
        if (j < 3) launch_p_var2(c, st, w, j);
    }
    LAUNCH(c, k_p_output, g64, 64, w, d_proofs, d_status);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

static int prove_slice(bppp_ctx *c, cudaStream_t st, size_t n, const uint64_t *d_x, const uint8_t *d_blinds, const uint8_t *d_rng,
                       const Merlin &init, uint8_t *d_proofs, int32_t *d_status) {
    SubPlan sp = plan_sub(c, n);
    int rc = fork_streams(c, st, sp);
    if (rc != BPPP_OK) return rc;
    for (int k = 0; k < sp.parts; k++) {
        cudaStream_t s = sp.parts == 1 ? st : c->sub_stream[k];
        rc = prove_part(c, s, sub_ws(c, sp, k), d_x + sp.lo[k], d_blinds + 32 * sp.lo[k], d_rng + (size_t)U64_RNG_BYTES * sp.lo[k], init,
                        d_proofs + (size_t)U64_PROOF_BYTES_COMPRESSED * sp.lo[k], d_status + sp.lo[k]);
        if (rc != BPPP_OK) return rc;
    }
    return join_streams(c, st, sp);
}

extern "C" int bppp_u64_prove_batch_dev(bppp_ctx *c, size_t n, const void *d_x, const void *d_blinds32, const void *d_rng,
                                        const uint8_t *label, size_t label_len, void *d_proofs_out, void *d_status, void *stream) {
    if (!c || (n && (!d_x || !d_blinds32 || !d_rng || !d_proofs_out || !d_status))) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream, as in the CUDA runtime
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
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    Merlin init; merlin_init(init, label, (uint32_t)label_len);
    for (size_t off = 0; off < n; off += c->max_batch) {
        size_t m = n - off < c->max_batch ? n - off : c->max_batch;
        SubPlan sp = plan_sub(c, m, SUB_HOST_PROVE);
        // Uploads on their own stream, in the order the phases need them: x, blindings and the 19 scalars (1,216 bytes) of the first
        // transcript phase for every part, then the 33 scalars (2,112 bytes) of the second phase, which arrive while the
        // first-stage sums run.  A part starts after a third of its RNG bytes instead of all 3,328 per proof.
        const size_t early = (size_t)U64_RNG_EARLY * 64, late = (size_t)U64_RNG_BYTES - early;
        cudaStream_t cs = c->copy_stream;
        for (int k = 0; k < sp.parts; k++) {
            size_t lo = sp.lo[k], cnt = sp.lo[k + 1] - sp.lo[k];
            CUDA_OK(cudaMemcpyAsync(c->d_in_a + 8 * lo, x + off + lo, 8 * cnt, cudaMemcpyHostToDevice, cs));
            CUDA_OK(cudaMemcpyAsync(c->d_in_b + 32 * lo, blinds32 + 32 * (off + lo), 32 * cnt, cudaMemcpyHostToDevice, cs));
            CUDA_OK(cudaMemcpy2DAsync(c->d_in_c + (size_t)U64_RNG_BYTES * lo, U64_RNG_BYTES, rng + (size_t)U64_RNG_BYTES * (off + lo), U64_RNG_BYTES,
                                      early, cnt, cudaMemcpyHostToDevice, cs));
            CUDA_OK(cudaEventRecord(c->ev_up[2 * k], cs));
        }
        for (int k = 0; k < sp.parts; k++) {
            size_t lo = sp.lo[k], cnt = sp.lo[k + 1] - sp.lo[k];
            CUDA_OK(cudaMemcpy2DAsync(c->d_in_c + (size_t)U64_RNG_BYTES * lo + early, U64_RNG_BYTES, rng + (size_t)U64_RNG_BYTES * (off + lo) + early,
                                      U64_RNG_BYTES, late, cnt, cudaMemcpyHostToDevice, cs));
            CUDA_OK(cudaEventRecord(c->ev_up[2 * k + 1], cs));
        }
        for (int k = 0; k < sp.parts; k++) {
            cudaStream_t st = sp.parts == 1 ? c->stream : c->sub_stream[k];
            size_t lo = sp.lo[k], cnt = sp.lo[k + 1] - sp.lo[k];
            CUDA_OK(cudaStreamWaitEvent(st, c->ev_up[2 * k], 0));
            int rc = prove_part(c, st, sub_ws(c, sp, k), (const uint64_t *)c->d_in_a + lo, c->d_in_b + 32 * lo, c->d_in_c + (size_t)U64_RNG_BYTES * lo,
                                init, c->d_out + (size_t)U64_PROOF_BYTES_COMPRESSED * lo, c->d_status + lo, c->ev_up[2 * k + 1]);
            if (rc != BPPP_OK) { cudaStreamSynchronize(cs); return rc; }
            CUDA_OK(cudaMemcpyAsync(proofs_out + (size_t)U64_PROOF_BYTES_COMPRESSED * (off + lo), c->d_out + (size_t)U64_PROOF_BYTES_COMPRESSED * lo,
                                    (size_t)U64_PROOF_BYTES_COMPRESSED * cnt, cudaMemcpyDeviceToHost, st));
            CUDA_OK(cudaMemcpyAsync(status + off + lo, c->d_status + lo, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, st));
        }
        CUDA_OK(cudaStreamSynchronize(cs));
        for (int k = 0; k < sp.parts; k++) CUDA_OK(cudaStreamSynchronize(sp.parts == 1 ? c->stream : c->sub_stream[k]));
    }
    return BPPP_OK;
}


// ---- phase-stepped prove for a caller-owned transcript (include/bppp.h) ---------------------------------------------
// prove_part cut at the transcript's challenge points: each step returns the compressed points
// the host appends next and takes the challenges it drew.  The RNG bytes stay in the context's staging area.
static int pstep_check(bppp_ctx *c, int stage, const char *what) {
    if (!c) return fail(BPPP_ERR_ARG, "null context");
    if (c->step.kind != 2 || c->step.stage != stage) return fail(BPPP_ERR_ARG, std::string(what) + ": called out of order for this context's stepped session");
    return BPPP_OK;
}
static int pstep_emit(bppp_ctx *c, cudaStream_t st, WS w, const EmitList &L, uint8_t *host_out) {
    launch_emit_points(c, st, w, L, c->d_out);
    CUDA_OK(cudaMemcpyAsync(host_out, c->d_out, (size_t)33 * L.n * w.n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}
static void pstep_xr(bppp_ctx *c, cudaStream_t st, WS w, int j) {        // X_j, R_j and the three normalisations round j's appends need
    TermMap all = identity_map(), tm;
    launch_msm_fixed(c, st, w, PL::XS, all, NUM_GENS, PL::PTS + PT_W * (PP_X + j));
    u64p_termmap_r(tm.gen, j);
    launch_msm_fixed(c, st, w, PL::RS, tm, 25, PL::PTS + PT_W * (PP_R + j));
    InvList L; L.n = 3;
    L.in[0] = PL::COM + 2 * FE_W; L.out[0] = PL::ZINV + FE_W * PP_COM;
    L.in[1] = PL::PTS + PT_W * (PP_X + j) + 2 * FE_W; L.out[1] = PL::ZINV + FE_W * (PP_X + j);
    L.in[2] = PL::PTS + PT_W * (PP_R + j) + 2 * FE_W; L.out[2] = PL::ZINV + FE_W * (PP_R + j);
    launch_batch_inv_list(c, st, w, L);
}
static EmitList emit_round(int j) {
    EmitList E; E.n = 3;
    E.pt[0] = PL::COM; E.zinv[0] = PL::ZINV + FE_W * PP_COM;
    E.pt[1] = PL::PTS + PT_W * (PP_X + j); E.zinv[1] = PL::ZINV + FE_W * (PP_X + j);
    E.pt[2] = PL::PTS + PT_W * (PP_R + j); E.zinv[2] = PL::ZINV + FE_W * (PP_R + j);
    return E;
}
extern "C" int bppp_u64_prove_begin(bppp_ctx *c, size_t n, const uint64_t *x, const uint8_t *blinds32, const uint8_t *rng, uint8_t *v33_out) {
    if (!c || !n || !x || !blinds32 || !rng || !v33_out) return fail(BPPP_ERR_ARG, "null argument");
    if (n > c->max_batch) return fail(BPPP_ERR_ARG, "a stepped session holds at most max_batch proofs");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->step = {}; c->active_parts = 1;
    WS w{c->d_ws, n};
    CUDA_OK(cudaMemcpyAsync(c->d_in_a, x, 8 * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(c->d_in_b, blinds32, 32 * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(c->d_in_c, rng, (size_t)U64_RNG_BYTES * n, cudaMemcpyHostToDevice, st));
    TermMap tm;
    LAUNCH(c, k_p_load, nblocks(n, 64), 64, w, (const uint64_t *)c->d_in_a, c->d_in_b);
    u64p_termmap_commit(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 2, PL::PTS + PT_W * PP_V);
    launch_batch_inv(c, st, w, PL::PTS + PT_W * PP_V + 2 * FE_W, PL::ZINV + FE_W * PP_V);
    EmitList E; E.n = 1; E.pt[0] = PL::PTS + PT_W * PP_V; E.zinv[0] = PL::ZINV + FE_W * PP_V;     // "reciprocal_commitment" (reciprocal.rs:114)
    int rc = pstep_emit(c, st, w, E, v33_out);
    if (rc != BPPP_OK) return rc;
    c->step.kind = 2; c->step.n = n; c->step.stage = 1;
    return BPPP_OK;
}
extern "C" int bppp_u64_prove_reciprocal(bppp_ctx *c, const uint8_t *e32, uint8_t *pts33_out) {
    int rc = pstep_check(c, 1, "bppp_u64_prove_reciprocal");
    if (rc != BPPP_OK) return rc;
    if (!e32 || !pts33_out) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    Merlin unused{};
    TermMap tm;
    uint8_t *d_chal = c->d_in_b;                     // the blindings were consumed by k_p_load
    CUDA_OK(cudaMemcpyAsync(d_chal, e32, 32 * n, cudaMemcpyHostToDevice, st));
    LAUNCH(c, k_p_phase1, nblocks(n, 64), 64, w, unused, (const uint8_t *)c->d_in_c, (const uint8_t *)d_chal, 32);
    for (int k = 0; k < 4; k++) {
        int nterms = u64p_termmap_stage1(tm.gen, k);
        launch_msm_fixed(c, st, w, PL::FS + 8 * u64p_stage1_scalar_base(k), tm, nterms, PL::PTS + PT_W * u64p_stage1_point(k));
    }
    LAUNCH(c, k_p_vprime, nblocks(n, 64), 64, w);
    InvList L; L.n = 5;
    for (int k = 0; k < 5; k++) { int p = u64p_stage1_norm_point(k); L.in[k] = PL::PTS + PT_W * p + 2 * FE_W; L.out[k] = PL::ZINV + FE_W * p; }
    launch_batch_inv_list(c, st, w, L);
    EmitList E; E.n = 4;                              // commitment_cl, commitment_cr, commitment_co, commitment_v (circuit.rs:347-350)
    const int slots[4] = {PP_CL, PP_CR, PP_CO, PP_VP};
    for (int k = 0; k < 4; k++) { E.pt[k] = PL::PTS + PT_W * slots[k]; E.zinv[k] = PL::ZINV + FE_W * slots[k]; }
    rc = pstep_emit(c, st, w, E, pts33_out);
    if (rc != BPPP_OK) return rc;
    c->step.stage = 2;
    return BPPP_OK;
}
extern "C" int bppp_u64_prove_circuit(bppp_ctx *c, const uint8_t *chal, uint8_t *cs33_out) {
    int rc = pstep_check(c, 2, "bppp_u64_prove_circuit");
    if (rc != BPPP_OK) return rc;
    if (!chal || !cs33_out) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    TermMap tm;
    CUDA_OK(cudaMemcpyAsync(c->d_in_b, chal, 128 * n, cudaMemcpyHostToDevice, st));
    LAUNCH(c, k_p_phase2, nblocks(n, 64), 64, w, (const uint8_t *)c->d_in_c, (const uint8_t *)c->d_in_b, 128);
    u64p_termmap_cs(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 42, PL::PTS + PT_W * PP_CS);
    launch_batch_inv(c, st, w, PL::PTS + PT_W * PP_CS + 2 * FE_W, PL::ZINV + FE_W * PP_CS);
    EmitList E; E.n = 1; E.pt[0] = PL::PTS + PT_W * PP_CS; E.zinv[0] = PL::ZINV + FE_W * PP_CS;    // commitment_cs (circuit.rs:472)
    rc = pstep_emit(c, st, w, E, cs33_out);
    if (rc != BPPP_OK) return rc;
    c->step.stage = 3;
    return BPPP_OK;
}
extern "C" int bppp_u64_prove_tau(bppp_ctx *c, const uint8_t *tau32, uint8_t *pts33_out) {
    int rc = pstep_check(c, 3, "bppp_u64_prove_tau");
    if (rc != BPPP_OK) return rc;
    if (!tau32 || !pts33_out) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    TermMap tm;
    CUDA_OK(cudaMemcpyAsync(c->d_in_b, tau32, 32 * n, cudaMemcpyHostToDevice, st));
    LAUNCH(c, k_p_phase3, nblocks(n, 64), 64, w, (const uint8_t *)c->d_in_b, 32);
    u64p_termmap_c0(tm.gen);
    launch_msm_fixed(c, st, w, PL::FS, tm, 43, PL::COM);
    pstep_xr(c, st, w, 0);
    rc = pstep_emit(c, st, w, emit_round(0), pts33_out);       // wnla_com, wnla_x, wnla_r of round 0 (wnla.rs:162-164)
    if (rc != BPPP_OK) return rc;
    c->step.stage = 4;
    return BPPP_OK;
}
extern "C" int bppp_u64_prove_round(bppp_ctx *c, int j, const uint8_t *y32, uint8_t *pts33_out) {
    if (j < 0 || j > 3) return fail(BPPP_ERR_ARG, "round index out of range");
    int rc = pstep_check(c, 4 + j, "bppp_u64_prove_round");
    if (rc != BPPP_OK) return rc;
    if (!y32 || (j < 3 && !pts33_out)) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    CUDA_OK(cudaMemcpyAsync(c->d_in_b, y32, 32 * n, cudaMemcpyHostToDevice, st));
    LAUNCH(c, k_p_round, nblocks(n, 64), 64, w, j, (const uint8_t *)c->d_in_b, 32);
    if (j < 3) {
        launch_p_var2(c, st, w, j);
        pstep_xr(c, st, w, j + 1);
        rc = pstep_emit(c, st, w, emit_round(j + 1), pts33_out);
        if (rc != BPPP_OK) return rc;
    } else {
        CUDA_OK(cudaStreamSynchronize(st));
        CUDA_OK(cudaGetLastError());
    }
    c->step.stage = 5 + j;
    return BPPP_OK;
}
extern "C" int bppp_u64_prove_finish(bppp_ctx *c, uint8_t *proofs_out, int32_t *status) {
    int rc = pstep_check(c, 8, "bppp_u64_prove_finish");
    if (rc != BPPP_OK) return rc;
    if (!proofs_out || !status) return fail(BPPP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->step.n;
    WS w{c->d_ws, n};
    LAUNCH(c, k_p_output, nblocks(n, 64), 64, w, c->d_out, c->d_status);
    CUDA_OK(cudaMemcpyAsync(proofs_out, c->d_out, (size_t)U64_PROOF_BYTES_COMPRESSED * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(status, c->d_status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    c->step = {};
    return BPPP_OK;
}
