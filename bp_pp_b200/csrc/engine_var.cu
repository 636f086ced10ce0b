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
void launch_v_var5(bppp_ctx *c, cudaStream_t st, WS w) {
    const int lanes = var_lanes_for(c, w.n);
    if (lanes == 1) LAUNCH(c, k_v_var5, nblocks(w.n, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w);
    else if (lanes == 2) LAUNCH(c, k_v_var5_lanes<2>, nblocks(w.n * 2, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w);
    else launch_v_var5_lat(c, st, w);
}
void launch_v_var2(bppp_ctx *c, cudaStream_t st, WS w, int j) {
    const int lanes = var_lanes_for(c, w.n);
    if (lanes == 1) LAUNCH(c, k_v_var2, nblocks(w.n, BPPP_VAR_BLOCK), BPPP_VAR_BLOCK, w, j);
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
