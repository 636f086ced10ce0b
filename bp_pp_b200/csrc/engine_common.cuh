// Shared host-side declarations of the engine's translation units (engine_*.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bppp.h"
#include "u64_prove.cuh"
#include "u64_verify.cuh"

namespace bppp {

int engine_fail(int code, const std::string &msg);
#define CUDA_OK(expr)                                                                                             \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess) {                                                                                  \
            (void)cudaGetLastError();   /* a recoverable error (e.g. out of memory) must not poison the next call */ \
            return engine_fail(_e == cudaErrorMemoryAllocation ? BPPP_ERR_NOMEM : BPPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
        }                                                                                                         \
    } while (0)

struct TermMap { int gen[NUM_GENS]; };

}  // namespace bppp

struct bppp_ctx {
    std::mutex mu;                  // entry points serialise per context (one workspace, one staging area); use one context per host thread for concurrency
    int device = 0;
    bppp::FixedTable T{};
    uint4 *d_tab = nullptr;
    bool owns_tab = false;          // false for a context created by bppp_ctx_create_shared
    size_t table_bytes = 0;
    double table_build_ms = 0;
    size_t max_batch = 0;
    uint32_t *d_ws = nullptr;       // workspace, max(VL::WORDS, PL::WORDS) * max_batch words
    size_t ws_words_per_proof = 0;
    uint8_t *d_in_a = nullptr, *d_in_b = nullptr, *d_in_c = nullptr;   // staging for host-buffer entry points
    uint8_t *d_out = nullptr;
    int32_t *d_status = nullptr;
    int32_t *d_flag = nullptr;      // one word: set by a kernel when an input the call must reject was seen
    cudaStream_t stream = nullptr;
    // a batch slice is cut into up to MAX_SUB sub-batches, each running its kernel sequence on its own stream,
    // so that tail waves and low-parallelism kernels of one sub-batch overlap with work of the others
    static constexpr int MAX_SUB = 8;
    int nsub = 2;
    int nsub_host = 4;              // host-buffer entry points: more, smaller sub-batches so that the first upload is short
    int nsub_host_prove = 2;        // host-buffer prove: its uploads are staged by phase (engine_prove.cu), two parts keep the kernels efficient
    cudaStream_t sub_stream[MAX_SUB] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_SUB] = {};
    // host-buffer prove: the uploads run on their own stream in the order the phases need them (engine_prove.cu), one event per part and stage
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_up[2 * MAX_SUB] = {};
    uint64_t launches = 0;
    int sm_count = 148;
    // phase-stepped session (bppp_u64_{verify,prove}_begin .. _finish): one at a time per context, lives in the workspace
    struct Step { int kind = 0; size_t n = 0; int stage = 0; int fmt = 0; } step;
    int active_parts = 1;
    int inflight_hint = 1;          // bppp_ctx_set_inflight: independent batches the caller keeps in flight on sibling contexts           // sub-batches of the slice in flight (they run concurrently: lane choices look at their sum)
    // ladder tables (u64_verify.cuh:tables_affine_level): affine levels; items per thread (= per inversion) by level when forced (BPPP_TAB_K)
    int tab_affine = 1, tab_k[3] = {0, 0, 0};
    int var_seg = 16;               // segments the verifier's one-thread ladders are cut into (engine_var.cu:k_v_var_seg); 1 = whole ladders
    int var_seg_one_item = 1;       // one block per work item (0: persistent warps -- measured slower, kept for experiments)
    int var_seg_warps = 0;          // > 0: warps per segmented launch (experiments); default = the launch's share of the warp slots
    int msm_lanes_override = 0, var_lanes_override = 0;   // BPPP_MSM_LANES_RT / BPPP_VAR_LANES_RT (experiments)
    // optional per-launch timing (bppp_ctx_profile_begin/end): CUDA events on the launching stream
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
};

namespace bppp {

static inline unsigned nblocks(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }
#define LAUNCH(ctx, kern, grid, block, ...)                                                    \
    do {                                                                                       \
        if ((ctx)->profiling) {                                                                \
            bppp_ctx::ProfRec _r; _r.name = #kern;                                             \
            cudaEventCreate(&_r.a); cudaEventCreate(&_r.b);                                    \
            cudaEventRecord(_r.a, st);                                                         \
            kern<<<(grid), (block), 0, st>>>(__VA_ARGS__);                                     \
            cudaEventRecord(_r.b, st);                                                         \
            (ctx)->prof.push_back(_r);                                                         \
        } else {                                                                               \
            kern<<<(grid), (block), 0, st>>>(__VA_ARGS__);                                     \
        }                                                                                      \
        (ctx)->launches++;                                                                     \
    } while (0)

#if defined(__CUDACC__)
// sum of the partial points held by the LANES adjacent threads of one proof (butterfly over warp shuffles, complete additions)
template <int LANES>
__device__ __forceinline__ Pt lanes_reduce(Pt acc) {
#pragma unroll 1
    for (int off = LANES / 2; off >= 1; off >>= 1) {
        Pt o;
#pragma unroll
        for (int k = 0; k < FE_W; k++) {
            o.x.v[k] = __shfl_xor_sync(0xFFFFFFFFu, acc.x.v[k], off);
            o.y.v[k] = __shfl_xor_sync(0xFFFFFFFFu, acc.y.v[k], off);
            o.z.v[k] = __shfl_xor_sync(0xFFFFFFFFu, acc.z.v[k], off);
        }
        acc = pt_add(acc, o);
    }
    return acc;
}
#endif

// threads of one affine table level over `items` (operation, point, proof) items: about 360 per SM over the concurrent sub-batches
// (B200 sweep, profiles/r2_kernel_experiments.txt #13: 16 / 32 / 64 items per inversion at 32,768 proofs per sub-batch,
// 4 / 8 / 16 at 8,192, 1 / 2 / 4 at 1,024), never more than 64 items per thread
static inline size_t tab_level_threads(const bppp_ctx *c, size_t items, int level) {
    if (c->tab_k[level - 1] > 0) return (items + c->tab_k[level - 1] - 1) / c->tab_k[level - 1];
    const int parts = c->active_parts > 1 ? c->active_parts : 1;
    size_t t = (size_t)c->sm_count * 360 / (size_t)parts, lo = (items + 63) / 64;
    if (t < lo) t = lo;
    return t < items ? t : (items ? items : 1);
}

// A verification batch this large runs as ONE part with segmented ladders (engine_var.cu:k_v_var_seg) instead of two
// sub-batches with whole ladders: 21.4 against 22.4 ms at 65,536 proofs (profiles/r2_kernel_experiments.txt #22).
static inline bool verify_one_part(const bppp_ctx *c, size_t n) {
    return c->var_seg > 1 && c->inflight_hint == 1 && !c->var_lanes_override && n >= 49152;     // measured at 65,536 only: smaller batches keep the two-part path
}

static inline TermMap identity_map() { TermMap tm; for (int t = 0; t < NUM_GENS; t++) tm.gen[t] = t; return tm; }

// engine_core.cu
// sub-batch plan for a slice of n proofs: part k covers [lo[k], lo[k+1]) and owns workspace words starting at
// d_ws + words_per_proof * lo[k] with row stride (lo[k+1] - lo[k])
struct SubPlan { int parts; size_t lo[bppp_ctx::MAX_SUB + 1]; };
enum { SUB_DEVICE = 0, SUB_HOST = 1, SUB_HOST_PROVE = 2 };
SubPlan plan_sub(bppp_ctx *c, size_t n, int kind = SUB_DEVICE);     // also records the number of concurrent parts in c->active_parts
int fork_streams(bppp_ctx *c, cudaStream_t caller, const SubPlan &sp);
int join_streams(bppp_ctx *c, cudaStream_t caller, const SubPlan &sp);
static inline WS sub_ws(const bppp_ctx *c, const SubPlan &sp, int k) {
    return WS{c->d_ws + c->ws_words_per_proof * sp.lo[k], sp.lo[k + 1] - sp.lo[k]};
}
void launch_msm_fixed(bppp_ctx *c, cudaStream_t st, WS w, int sc_off, const TermMap &tm, int nterms, int out_off);
void launch_batch_inv(bppp_ctx *c, cudaStream_t st, WS w, int in_off, int out_off);
void launch_batch_inv_list(bppp_ctx *c, cudaStream_t st, WS w, const InvList &L);
// compressed SEC1 bytes of up to 4 projective workspace points per proof (with their batch-inverted Z): out[(i * L.n + k) * 33]
struct EmitList { int n; int pt[4]; int zinv[4]; };
void launch_emit_points(bppp_ctx *c, cudaStream_t st, WS w, const EmitList &L, uint8_t *d_out);
// engine_var.cu: joint variable-base ladders (one thread per proof)
void launch_v_var5(bppp_ctx *c, cudaStream_t st, WS w);
void launch_v_var2(bppp_ctx *c, cudaStream_t st, WS w, int j);
void launch_p_var2(bppp_ctx *c, cudaStream_t st, WS w, int j);

}  // namespace bppp
