// libbppp.so, variable-base translation unit: the joint Straus ladders over per-proof points
// (one thread per proof, tables of 1P..8P per point in thread-local memory).
//
// engine_var_lat.cu compiles this file a second time with BPPP_VAR_LAT: only the 4-lane kernels, with the call boundary at the
// point operation (ec.cuh: BPPP_PTJ_*_NOINLINE -- a doubling / mixed addition is one function with its products inlined, so
// ptxas overlaps the independent products of one formula).  That shortens the dependent chain of a proof: 4-lane ladders
// 1.84 -> 1.53 ms at 2,048 proofs, 2.37 -> 2.17 ms at 8,192 (profiles/r2_kernel_experiments.txt); with the GPU full it loses
// (instruction-cache footprint), so the throughput kernels keep the field-level calls.
#ifndef BPPP_VAR_INLINE
#define BPPP_FE_NOINLINE 1   // see fe.cuh: keeps the ladder loop inside the instruction cache
#endif
#if defined(BPPP_VAR_LAT)
#define BPPP_PTJ_DBL_NOINLINE 1
#define BPPP_PTJ_ADD_NOINLINE 1
#define LATNAME(x) x##_lat
#else
#define LATNAME(x) x
#endif
#include "engine_common.cuh"

using namespace bppp;

#ifndef BPPP_VAR_BLOCK
#define BPPP_VAR_BLOCK 64
#endif
#ifndef BPPP_VAR_MINBLOCKS
#define BPPP_VAR_MINBLOCKS 7
#endif
#if !defined(BPPP_VAR_LAT)
__global__ void __launch_bounds__(BPPP_VAR_BLOCK, BPPP_VAR_MINBLOCKS) k_v_var5(WS w) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_var5_one(w, i);
}
__global__ void __launch_bounds__(BPPP_VAR_BLOCK, BPPP_VAR_MINBLOCKS) k_v_var2(WS w, int j) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64v_var2_one(w, i, j);
}
__global__ void __launch_bounds__(BPPP_VAR_BLOCK, BPPP_VAR_MINBLOCKS) k_p_var2(WS w, int j) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.n) u64p_var2_one(w, i, j);
}
#endif
// The verifier's ladders in segments (u64_verify.cuh:straus_tables_seg), persistent warps.  Work items are (segment, 32-proof
// chain) pairs in segment-major order; a warp draws a ticket, waits until the chain's previous segment has been published,
// continues the chain from the scratch rows, publishes its own segment and draws again.  The ticket order makes the wait
// safe: the previous segment's ticket was drawn earlier by a warp that is running (or done).  The grid is this launch's share
// of the GPU's warp slots, a little more than it has chains, so a chain continues on whichever warp frees first instead of
// staying on the scheduler it started on, and concurrent sub-batches do not crowd each other out with waiting blocks.
// sync: [0] ticket counter, [1] set if a wait ever timed out (seconds; never in a correct run), [32 + g] segments done of chain g.
template <int KIND>     // 0: five-point group (COM = ACC + ...), 1: two-point group of round j
__global__ void __launch_bounds__(32, 16) k_v_var_seg(WS w, int j, int nseg, uint32_t *sync, int one_item) {
    const unsigned lane = threadIdx.x;
    const unsigned nchains = (unsigned)((w.n + 31) / 32), total = nchains * (unsigned)nseg;
#pragma unroll 1
    for (;;) {
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(sync, 1u);
        ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
        if (ticket >= total) return;
        const unsigned seg = ticket / nchains, g = ticket - seg * nchains;
        volatile uint32_t *done = sync + 32 + g;
        const size_t i = (size_t)g * 32 + lane;
        if (seg) {
            unsigned spins = 0;
            while (*done < seg) {
                __nanosleep(100);
                if (++spins > (1u << 26)) {                       // several seconds: fail loudly instead of hanging the GPU
                    sync[1] = 1u;
                    if (i < w.n) ws_st(w, i, VL::STATUS, (uint32_t)ST_BAD_ARG);
                    return;
                }
            }
            __threadfence();
        }
        if (i < w.n) {
            if (KIND == 0) u64v_var5_seg(w, i, (int)seg, nseg);
            else u64v_var2_seg(w, i, j, (int)seg, nseg);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) *done = seg + 1;
        if (one_item) return;
    }
}
// The same ladders with LANES adjacent threads per proof (straus_tables_partial): the GLV halves are shared out, every lane
// repeats the doublings, the partial sums meet through warp shuffles.  More total work, a shorter dependent chain and
// LANES times the warps: selected when the (sub-)batch alone leaves most of the GPU idle.
template <int LANES>
__global__ void __launch_bounds__(BPPP_VAR_BLOCK, BPPP_VAR_MINBLOCKS) LATNAME(k_v_var5_lanes)(WS w) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = t / LANES; const int lane = (int)(t % LANES);
    const bool live = i < w.n;
    if (!live) i = w.n - 1;
    Pt part = lanes_reduce<LANES>(u64v_var5_partial(w, i, lane, LANES));
    if (live && lane == 0) ws_st_pt(w, i, VL::COM, pt_add(part, ws_ld_pt(w, i, VL::ACC)));
}
template <int LANES>
__global__ void __launch_bounds__(BPPP_VAR_BLOCK, BPPP_VAR_MINBLOCKS) LATNAME(k_v_var2_lanes)(WS w, int j) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = t / LANES; const int lane = (int)(t % LANES);
    const bool live = i < w.n;
    if (!live) i = w.n - 1;
    Pt part = lanes_reduce<LANES>(u64v_var2_partial(w, i, j, lane, LANES));
    if (live && lane == 0) ws_st_pt(w, i, VL::COM, pt_add(part, ws_ld_pt(w, i, VL::COM)));
}
template <int LANES>
__global__ void __launch_bounds__(BPPP_VAR_BLOCK, BPPP_VAR_MINBLOCKS) LATNAME(k_p_var2_lanes)(WS w, int j) {
    (void)j;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = t / LANES; const int lane = (int)(t % LANES);
    const bool live = i < w.n;
    if (!live) i = w.n - 1;
    Pt part = lanes_reduce<LANES>(u64p_var2_partial(w, i, lane, LANES));
    if (live && lane == 0) ws_st_pt(w, i, PL::COM, pt_add(part, ws_ld_pt(w, i, PL::COM)));
}
#if defined(BPPP_VAR_LAT)
namespace bppp {
void launch_v_var5_lat(bppp_ctx *c, cudaStream_t st, WS w) { LAUNCH(c, k_v_var5_lanes_lat<4>, nblocks(w.n * 4, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w); }
void launch_v_var2_lat(bppp_ctx *c, cudaStream_t st, WS w, int j) { LAUNCH(c, k_v_var2_lanes_lat<4>, nblocks(w.n * 4, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j); }
void launch_p_var2_lat(bppp_ctx *c, cudaStream_t st, WS w, int j) { LAUNCH(c, k_p_var2_lanes_lat<4>, nblocks(w.n * 4, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j); }
}  // namespace bppp
#else
__global__ void __launch_bounds__(64) k_p_tables_build(WS w, int j) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int p = (int)(t / w.n); size_t i = t - (size_t)p * w.n;
    if (p < 2) u64p_table_build_one(w, i, j, p);
}
__global__ void __launch_bounds__(128) k_p_tables_normalize(WS w, size_t nthreads) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) tables_normalize_strided(w, ptab_region(), t, nthreads);
}
namespace bppp {
// engine_var_lat.cu: the 4-lane kernels built for latency
void launch_v_var5_lat(bppp_ctx *c, cudaStream_t st, WS w);
void launch_v_var2_lat(bppp_ctx *c, cudaStream_t st, WS w, int j);
void launch_p_var2_lat(bppp_ctx *c, cudaStream_t st, WS w, int j);
// lanes per proof for the ladders, by the number of proofs in flight on the GPU
static int var_lanes_for(const bppp_ctx *c, size_t n) {
    if (c->var_lanes_override) return c->var_lanes_override;
    const size_t full = (size_t)c->sm_count * 448;
    n *= (size_t)c->active_parts * (size_t)c->inflight_hint;
    // measured (tools/batch_sweep.py --lane-sweep): 4 lanes win up to ~8k proofs in flight, 2 lanes up to ~16k
    if (n * 8 <= full) return 4;
    if (n * 4 <= full) return 2;
    return 1;
}
// segmented ladders: where one thread per proof is chosen, the launch has the GPU to itself (a single part, no sibling batches
// in flight: two such launches side by side crowd each other out with waiting blocks) and is large enough to matter
static bool var_segments(const bppp_ctx *c, size_t n) { return verify_one_part(c, n) && c->active_parts == 1; }
// warps of one segmented launch: its share of the 16 warp slots per SM (the sub-batches of a call and the caller's batches in
// flight run side by side), never more than it has work items
static unsigned var_seg_grid(const bppp_ctx *c, size_t n) {
    const size_t items = ((n + 31) / 32) * (size_t)c->var_seg;
    size_t share = (size_t)c->sm_count * 16 / ((size_t)c->active_parts * (size_t)c->inflight_hint);
    if (c->var_seg_warps > 0) share = (size_t)c->var_seg_warps;
    if (c->var_seg_one_item) return (unsigned)items;
    if (share < 64) share = 64;
    return (unsigned)(items < share ? items : share);
}
static uint32_t *var_seg_sync(cudaStream_t st, WS w) {
    uint32_t *sync = w.p + (size_t)(VL::TAB + LS_SYNC) * w.n;
    cudaMemsetAsync(sync, 0, sizeof(uint32_t) * (32 + (w.n + 31) / 32), st);
    return sync;
}
void launch_v_var5(bppp_ctx *c, cudaStream_t st, WS w) {
    const int lanes = var_lanes_for(c, w.n);
    if (lanes == 1 && var_segments(c, w.n)) LAUNCH(c, k_v_var_seg<0>, var_seg_grid(c, w.n), 32, w, 0, c->var_seg, var_seg_sync(st, w), c->var_seg_one_item);
    else if (lanes == 1) LAUNCH(c, k_v_var5, nblocks(w.n, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w);
    else if (lanes == 2) LAUNCH(c, k_v_var5_lanes<2>, nblocks(w.n * 2, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w);
    else launch_v_var5_lat(c, st, w);
}
void launch_v_var2(bppp_ctx *c, cudaStream_t st, WS w, int j) {
    const int lanes = var_lanes_for(c, w.n);
    if (lanes == 1 && var_segments(c, w.n)) LAUNCH(c, k_v_var_seg<1>, var_seg_grid(c, w.n), 32, w, j, c->var_seg, var_seg_sync(st, w), c->var_seg_one_item);
    else if (lanes == 1) LAUNCH(c, k_v_var2, nblocks(w.n, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j);
    else if (lanes == 2) LAUNCH(c, k_v_var2_lanes<2>, nblocks(w.n * 2, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j);
    else launch_v_var2_lat(c, st, w, j);
}
void launch_p_var2(bppp_ctx *c, cudaStream_t st, WS w, int j) {
    // tables of X_j, R_j (point-major), one cross-proof inversion for their 16 entries, then the ladder.  (Two points per
    // proof are too few items for the verifier's affine levels to pay: three latency-bound launches, 1.12 against 1.13 ms.)
    LAUNCH(c, k_p_tables_build, nblocks(w.n * 2, 64), 64, w, j);
    size_t items = w.n * PL::TAB_ENTRIES, nthreads = (items + 15) / 16;
    size_t min_threads = (size_t)c->sm_count * 128;
    if (nthreads < min_threads) nthreads = items < min_threads ? items : min_threads;
    LAUNCH(c, k_p_tables_normalize, nblocks(nthreads, 128), 128, w, nthreads);
    const int lanes = var_lanes_for(c, w.n);
    if (lanes == 1) LAUNCH(c, k_p_var2, nblocks(w.n, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j);
    else if (lanes == 2) LAUNCH(c, k_p_var2_lanes<2>, nblocks(w.n * 2, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j);
    else launch_p_var2_lat(c, st, w, j);
}
}  // namespace bppp
#endif
