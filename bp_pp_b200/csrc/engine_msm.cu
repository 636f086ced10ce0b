// libbppp.so, variable-base MSM translation unit: sum_i k_i * P_i over arbitrary points, the device-side
// replacement of util::vector_mul<ProjectivePoint> (reference src/util.rs:46-60, which performs one full
// scalar multiplication per term) for large n -- WNLA X/R commitments (src/wnla.rs:152-160), wnla.commit
// (src/wnla.rs:66-72) and the circuit commitments (src/circuit.rs:335-345,469-470,522-524).
//
// Pippenger over the GLV halves of the scalars (128-bit magnitudes, the lambda half on (beta x, y)) with signed c-bit
// windows, every kernel in this file (no library call):
//   k_msm_digits      GLV split + signed 16-bit digits of both halves of every scalar, window-major;  k_msm_endo_x: beta x
//   k_msm_hist / k_scan_* / k_msm_scatter   counting sort of the point indices by (window, bucket): shared-memory histograms per
//                     (chunk, window) block, one exclusive scan, scatter through shared-memory cursors -- no global atomics
//   k_msm_slices      one thread per 64 consecutive sorted entries (XYZZ mixed additions, next point prefetched): perfectly
//                     balanced warps; whole buckets are written directly, runs cut by a slice boundary as head / tail pieces
//   k_msm_fixup(_wide) adds the pieces of cut buckets (a bucket of skewed scalars spanning > 32 slices gets a block)
//   k_msm_chunks      per window: chunked running-sum reduction  sum_b (b+1) B_b  (two adds per bucket)
//   k_pt_sum_groups   tree sums;  k_msm_horner: sum_w 2^(c w) W_w
// Small inputs (n <= 1024) use one GLV scalar multiplication per point and the same tree sum.
#define BPPP_FE_NOINLINE 1
#define BPPP_GENERIC_ALLOC 1   // engine_generic.cuh: cudaMalloc / cudaFree of this file go through the caching allocator
#define BPPP_PTX_ADD_NOINLINE 1   // ec.cuh: the bucket accumulation's XYZZ addition as one call with inlined products
#include "engine_generic.cuh"

#include <mutex>

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

namespace bppp {

// ---- point / scalar array helpers (AoS words in device memory) ----
__device__ __forceinline__ bool load_dev_point(PtA &q, const uint32_t *pts, size_t idx) {
    const uint4 *p = reinterpret_cast<const uint4 *>(pts + 16 * idx);
    TableEntryRaw r; r.a = __ldg(p); r.b = __ldg(p + 1); r.c = __ldg(p + 2); r.e = __ldg(p + 3);
    return table_decode(q, r);
}
__device__ __forceinline__ Pt ld_pt30(const uint32_t *p) { Pt r;
#pragma unroll
    for (int k = 0; k < FE_W; k++) { r.x.v[k] = p[k]; r.y.v[k] = p[FE_W + k]; r.z.v[k] = p[2 * FE_W + k]; }
    return r; }
__device__ __forceinline__ void st_pt30(uint32_t *p, const Pt &a) {
#pragma unroll
    for (int k = 0; k < FE_W; k++) { p[k] = a.x.v[k]; p[FE_W + k] = a.y.v[k]; p[2 * FE_W + k] = a.z.v[k]; } }

__global__ void k_decode_points(const uint8_t *in, int fmt, uint32_t *out, int32_t *bad, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PtA a;
    int s = fmt == FMT_COMPRESSED ? pta_decompress(a, in + 33 * i) : pta_from_xy64(a, in + 64 * i);
    uint32_t x[8], y[8];
    if (s == 0) { fe_to_words(x, fe_normalize(a.x)); fe_to_words(y, fe_normalize(a.y)); }
    else {
#pragma unroll
        for (int k = 0; k < 8; k++) { x[k] = 0; y[k] = 0; }
        if (s < 0) atomicExch(bad, (int32_t)ST_BAD_POINT);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { out[16 * i + k] = x[k]; out[16 * i + 8 + k] = y[k]; }
}
__global__ void k_decode_scalars(const uint8_t *in, uint32_t *out, int32_t *bad, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sc s;
    if (!sc_from_be32(s, in + 32 * i)) { atomicExch(bad, (int32_t)ST_BAD_SCALAR); s = sc_zero(); }
#pragma unroll
    for (int k = 0; k < 8; k++) out[8 * i + k] = s.v[k];
}
__global__ void k_encode_scalars(const uint32_t *in, uint8_t *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sc s;
#pragma unroll
    for (int k = 0; k < 8; k++) s.v[k] = in[8 * i + k];
    sc_to_be32(out + 32 * i, s);
}
// projective (PT_W words each) -> affine bytes; one inversion per point (used for a handful of outputs)
__global__ void k_encode_points(const uint32_t *pts30, int fmt, uint8_t *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pt p = ld_pt30(pts30 + PT_W * i);
    bool id = pt_is_identity(p);
    PtA a = pt_to_affine_with_zinv(p, fe_inv(p.z));
    if (fmt == FMT_COMPRESSED) pta_compress(out + 33 * i, a, id); else pta_to_xy64(out + 64 * i, a, id);
}

// ---- Pippenger ----
// GLV: k = k1 + k2 lambda with |k1|, |k2| < 2^128 (ec.cuh:glv_split), so scalar i becomes two ENTRIES -- 2i for (k1, P_i) and
// 2i + 1 for (k2, lambda P_i = (beta x_i, y_i)) -- of 128-bit magnitudes: half the windows, half the buckets to reduce and half
// the doublings of the Horner tail for the same number of bucket additions.
// Signed c-bit digits of every entry, window-major (row w holds the 2n entries of window w), 16 bits each: low 15 bits
// magnitude - 1 (the magnitude is at most 2^(c-1) <= 2^15), bit 15 = sign (the half-scalar's sign folded in).  Zero digits are
// stored as 0xFFFF.  nwin = floor(128 / c) + 1 windows absorb the last carry (the top window holds < c - 1 real bits).
__global__ void k_msm_digits(const uint32_t *sc, size_t n, int c, int nwin, uint16_t *digits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sc k;
#pragma unroll
    for (int j = 0; j < 8; j++) k.v[j] = sc[8 * i + j];
    const GlvSplit g = glv_split(k);
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1u;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        uint32_t m[6];
#pragma unroll
        for (int j = 0; j < 4; j++) m[j] = h ? g.k2[j] : g.k1[j];
        m[4] = 0; m[5] = 0;
        const uint32_t sneg = (h ? g.neg2 : g.neg1) ? 1u : 0u;
        uint32_t carry = 0;
        for (int w = 0; w < nwin; w++) {
            int bit = w * c, word = bit >> 5, sh = bit & 31;
            uint64_t v = word < 5 ? ((uint64_t)m[word] | ((uint64_t)m[word + 1] << 32)) : 0;
            uint32_t raw = ((uint32_t)(v >> sh) & mask) + carry;
            uint32_t neg = 0, mag = raw;
            carry = 0;
            // raw == half may be written +half or -half + carry: at c = 16 the code (magnitude 2^15, final sign 1) is the zero
            // marker 0xFFFF, so the tie takes whichever sign leaves the FINAL sign clear (never needed in the top window)
            if (raw > half || (raw == half && sneg && w + 1 < nwin)) { mag = (1u << c) - raw; neg = 1; carry = 1; }
            digits[(size_t)w * 2 * n + 2 * i + h] = mag == 0 ? (uint16_t)0xFFFFu : (uint16_t)((mag - 1) | ((neg ^ sneg) << 15));
        }
    }
}
// beta * x of every point (canonical words): the x coordinate of the lambda-half entries; (0, 0) stays the identity sentinel
__global__ void k_msm_endo_x(const uint32_t *pts, size_t n, uint32_t *bx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = pts[16 * i + j];
    fe_to_words(x, fe_normalize(fe_mul(fe_from_words(x), fe_beta())));
#pragma unroll
    for (int j = 0; j < 8; j++) bx[8 * i + j] = x[j];
}
// point of sorted entry e: P_(e >> 1), its x taken from the beta * x array for the lambda half
__device__ __forceinline__ bool load_entry_point(PtA &q, const uint32_t *pts, const uint32_t *bx, uint32_t e) {
    const size_t idx = e >> 1;
    const uint4 *px = (e & 1u) ? reinterpret_cast<const uint4 *>(bx + 8 * idx) : reinterpret_cast<const uint4 *>(pts + 16 * idx);
    const uint4 *py = reinterpret_cast<const uint4 *>(pts + 16 * idx + 8);
    TableEntryRaw r; r.a = __ldg(px); r.b = __ldg(px + 1); r.c = __ldg(py); r.e = __ldg(py + 1);
    return table_decode(q, r);
}
// Counting sort of the (window, bucket) keys without a library and without global atomics.  Pass 1: block (j, w) histograms
// chunk j of window w's digits in shared memory (2^(c-1) counters, 128 KB at c = 16) and writes the counts to
// counts[(w * half + b) * NCH + j]; an exclusive scan of that array gives every bucket's start (entry b * NCH).
// The order inside a bucket depends on thread timing; bucket SUMS do not (exact group arithmetic).
__global__ void __launch_bounds__(1024) k_msm_hist(const uint16_t *digits, size_t n, uint32_t half, uint32_t nch, uint32_t *counts) {
    extern __shared__ uint32_t sh[];
    const uint32_t j = blockIdx.x, w = blockIdx.y;
    for (uint32_t b = threadIdx.x; b < half; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    size_t per = (n + nch - 1) / nch; per = (per + 7) & ~(size_t)7;           // chunk starts stay 16-byte aligned
    const size_t lo = (size_t)j * per < n ? (size_t)j * per : n, hi = lo + per < n ? lo + per : n;
    const uint16_t *d16 = digits + (size_t)w * n + lo;
    const size_t m = hi - lo, m8 = (reinterpret_cast<uintptr_t>(d16) & 15) == 0 ? m / 8 : 0;
    for (size_t i = threadIdx.x; i < m8; i += blockDim.x) {
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(d16) + i);
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t a = wd[k] & 0xFFFFu, b = wd[k] >> 16;
            if (a != 0xFFFFu) atomicAdd(&sh[a & 0x7FFFu], 1u);
            if (b != 0xFFFFu) atomicAdd(&sh[b & 0x7FFFu], 1u);
        }
    }
    for (size_t i = m8 * 8 + threadIdx.x; i < m; i += blockDim.x) { uint16_t d = d16[i]; if (d != 0xFFFFu) atomicAdd(&sh[d & 0x7FFFu], 1u); }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < half; b += blockDim.x) counts[((size_t)w * half + b) * nch + j] = sh[b];
}
// Pass 2: block (r, w) owns the bucket RANGE [r * span, (r + 1) * span) of window w, streams all n digits of the window
// (coalesced 16-bit reads) and scatters the entries of its buckets through shared-memory cursors.  The block's writes land
// in one contiguous region of `vals` (its buckets are adjacent), so they combine in L2 instead of dirtying a 32-byte sector
// per 4-byte entry all over the array (a chunk-of-scalars scatter measured 0.75 ms at 2^21 points against 0.14 ms for pass 1).
__global__ void __launch_bounds__(1024) k_msm_scatter(const uint16_t *digits, size_t n, uint32_t half, uint32_t span, uint32_t nch, const uint32_t *scan, uint32_t *vals) {
    extern __shared__ uint32_t sh[];
    const uint32_t w = blockIdx.y, base = blockIdx.x * span, top = base + span < half ? base + span : half;
    for (uint32_t b = base + threadIdx.x; b < top; b += blockDim.x) sh[b - base] = scan[((size_t)w * half + b) * nch];
    __syncthreads();
    const uint16_t *d16 = digits + (size_t)w * n;
    auto place = [&](uint32_t d, size_t i) {
        uint32_t b = d & 0x7FFFu;
        if (d != 0xFFFFu && b >= base && b < top) vals[atomicAdd(&sh[b - base], 1u)] = (uint32_t)i | ((d >> 15) << 31);
    };
    // eight digits per 16-byte load: one load in flight per warp made this pass latency-bound (1.1 ms at 2^21 points)
    const size_t n8 = (reinterpret_cast<uintptr_t>(d16) & 15) == 0 ? n / 8 : 0;
    for (size_t i = threadIdx.x; i < n8; i += blockDim.x) {
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(d16) + i);
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { place(wd[k] & 0xFFFFu, 8 * i + 2 * k); place(wd[k] >> 16, 8 * i + 2 * k + 1); }
    }
    for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) place(d16[i], i);
}
// exclusive prefix sum of `count` 32-bit values (+ the total at out[count]) in three launches: 2048 values per block,
// one block over the block totals, offsets added back
__global__ void __launch_bounds__(256) k_scan_blocks(const uint32_t *in, size_t count, uint32_t *out, uint32_t *block_sums) {
    __shared__ uint32_t sh[256];
    const size_t base = (size_t)blockIdx.x * 2048 + (size_t)threadIdx.x * 8;
    uint32_t v[8], run = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = base + k < count ? in[base + k] : 0u; run += v[k]; }
    sh[threadIdx.x] = run;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        uint32_t t = (int)threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t excl = sh[threadIdx.x] - run;
#pragma unroll
    for (int k = 0; k < 8; k++) { if (base + k < count) out[base + k] = excl; excl += v[k]; }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = sh[255];
}
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t *block_sums, size_t nblk, uint32_t *total_out) {
    __shared__ uint32_t sh[1024];
    uint32_t carry = 0;
    for (size_t base = 0; base < nblk; base += 1024) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nblk ? block_sums[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            uint32_t t = (int)threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblk) block_sums[i] = carry + sh[threadIdx.x] - v;
        carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(256) k_scan_add(uint32_t *out, size_t count, const uint32_t *block_sums) {
    const size_t base = (size_t)blockIdx.x * 2048 + (size_t)threadIdx.x * 8;
    const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; k++) if (base + k < count) out[base + k] += add;
}

// ---- bucket accumulation over equal SLICES of the sorted index array ----
// Thread t adds the points of entries [t * SL, (t + 1) * SL) whatever buckets they fall in: every thread of a warp does the
// same number of additions (one thread per bucket wasted ~25 % of every warp on the largest of its 32 Poisson-sized buckets).
// A run of entries of one bucket that covers the whole bucket is written to the bucket; a run cut by a slice boundary goes
// to the slice's head slot (first run) or tail slot (last run) as an XYZZ point, and k_msm_fixup adds the pieces.
// Slice length: 64 when the input alone gives every SM enough slices; shorter for small inputs, which would otherwise be a few
// thousand threads each walking 64 dependent additions (2^16 points: 407 us of pure latency).
struct SortedView { const uint32_t *scan; uint32_t nch, nb, sl; };   // start of bucket b = scan[b * nch]; scan[nb * nch] = entries; sl = slice length
__device__ __forceinline__ uint32_t sv_start(const SortedView &v, uint32_t b) { return v.scan[(size_t)b * v.nch]; }
__device__ __forceinline__ void st_ptx32(uint32_t *p, const PtX &a) {
#pragma unroll
    for (int k = 0; k < FE_W; k++) { p[k] = a.inf ? 0u : a.x.v[k]; p[FE_W + k] = a.inf ? 0u : a.y.v[k]; p[2 * FE_W + k] = a.inf ? 0u : a.zz.v[k]; p[3 * FE_W + k] = a.inf ? 0u : a.zzz.v[k]; }
}
__device__ __forceinline__ Pt ld_ptx32_as_pt(const uint32_t *p) {
    PtX a;
#pragma unroll
    for (int k = 0; k < FE_W; k++) { a.x.v[k] = p[k]; a.y.v[k] = p[FE_W + k]; a.zz.v[k] = p[2 * FE_W + k]; a.zzz.v[k] = p[3 * FE_W + k]; }
    a.inf = fe_normalizes_to_zero(a.zz);
    return ptx_to_pt(a);
}
__global__ void __launch_bounds__(64, 7) k_msm_slices(const uint32_t *pts, const uint32_t *bx, const uint32_t *vals, SortedView sv, uint32_t *buckets, uint32_t *head, uint32_t *tail) {
    const uint32_t total = sv_start(sv, sv.nb);
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t MSM_SL = sv.sl;
    if (t * MSM_SL >= total) return;
    const uint32_t s = (uint32_t)(t * MSM_SL), e = s + MSM_SL < total ? s + MSM_SL : total;
    // bucket of entry s: the last b with start[b] <= s (empty buckets share a start with their successor: take the last)
    uint32_t lo = 0, hi = sv.nb;           // invariant: start[lo] <= s < start[hi]
    while (hi - lo > 1) { uint32_t mid = lo + (hi - lo) / 2; if (sv_start(sv, mid) <= s) lo = mid; else hi = mid; }
    uint32_t b = lo, bend = sv_start(sv, b + 1);
    uint32_t run_begin = s;
    bool first_run = true;
    PtX acc = ptx_identity();
    uint32_t v = vals[s];
    PtA q; bool ok = load_entry_point(q, pts, bx, v & 0x7FFFFFFFu);
#pragma unroll 1
    for (uint32_t p = s; p < e; p++) {
        if (p == bend) {                   // bucket b ends here: the run is complete unless it began at a slice boundary inside b
            if (run_begin == sv_start(sv, b)) st_pt30(buckets + PT_W * (size_t)b, ptx_to_pt(acc));
            else st_ptx32(head + 32 * t, acc);                      // only the first run can have begun inside its bucket
            acc = ptx_identity(); first_run = false; run_begin = p;
            do { b++; bend = sv_start(sv, b + 1); } while (bend <= p);
        }
        const uint32_t vc = v; const PtA qc = q; const bool okc = ok;
        if (p + 1 < e) { v = vals[p + 1]; ok = load_entry_point(q, pts, bx, v & 0x7FFFFFFFu); }       // next point in flight during this addition
        if (okc) {
            PtA qa = qc;
            if (vc >> 31) qa.y = fe_normalize_weak(fe_negate(qa.y, 1));
            acc = ptx_add_mixed_hot(acc, qa);
        }
    }
    if (run_begin == sv_start(sv, b) && e == bend) st_pt30(buckets + PT_W * (size_t)b, ptx_to_pt(acc));
    else if (first_run) st_ptx32(head + 32 * t, acc);
    else st_ptx32(tail + 32 * t, acc);
}
// buckets cut by slice boundaries: sum of their pieces (slot of bucket B in slice t: head when B began at or before the slice's
// first entry, else tail); empty buckets become the identity; buckets spanning more than 32 slices go to the block kernel
struct SpanQueue { uint32_t *count; uint32_t *bucket; uint32_t cap; };
__device__ __forceinline__ const uint32_t *slice_piece(const uint32_t *head, const uint32_t *tail, uint32_t S, uint32_t t, uint32_t MSM_SL) {
    return (S <= t * MSM_SL ? head : tail) + 32 * (size_t)t;
}
__global__ void __launch_bounds__(64) k_msm_fixup(SortedView sv, const uint32_t *head, const uint32_t *tail, uint32_t *buckets, SpanQueue sq) {
    const size_t B = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (B >= sv.nb) return;
    const uint32_t S = sv_start(sv, (uint32_t)B), E = sv_start(sv, (uint32_t)B + 1);
    if (S == E) { st_pt30(buckets + PT_W * B, pt_identity()); return; }
    const uint32_t MSM_SL = sv.sl;
    const uint32_t t0 = S / MSM_SL, t1 = (E - 1) / MSM_SL;
    if (t0 == t1) return;                                   // whole bucket inside one slice: k_msm_slices wrote it
    if (t1 - t0 >= 32) { uint32_t slot = atomicAdd(sq.count, 1u); if (slot < sq.cap) { sq.bucket[slot] = (uint32_t)B; return; } }
    Pt acc = ld_ptx32_as_pt(slice_piece(head, tail, S, t0, MSM_SL));
#pragma unroll 1
    for (uint32_t t = t0 + 1; t <= t1; t++) acc = pt_add(acc, ld_ptx32_as_pt(slice_piece(head, tail, S, t, MSM_SL)));
    st_pt30(buckets + PT_W * B, acc);
}
// A bucket spanning 32 slices or more (skewed scalars -- and, at 16-bit windows, the carry bucket of the top window, which
// holds half of all entries): WIDE_SPLIT blocks each sum an interleaved share of its pieces into `part`, k_msm_fixup_wide_sum
// adds the shares.  A fixed, small grid strides over the queue (it is short: one block per possible entry cost 0.23 ms of
// empty launches at 2^21 points).
static constexpr uint32_t WIDE_SPLIT = 8;
__global__ void __launch_bounds__(128) k_msm_fixup_wide(SortedView sv, const uint32_t *head, const uint32_t *tail, uint32_t *part, SpanQueue sq) {
    __shared__ uint32_t sh[128 * PT_W];
    uint32_t cnt = *sq.count; if (cnt > sq.cap) cnt = sq.cap;
    const uint32_t split = blockIdx.y;
    for (uint32_t q = blockIdx.x; q < cnt; q += gridDim.x) {
        const uint32_t B = sq.bucket[q];
        const uint32_t MSM_SL = sv.sl;
        const uint32_t S = sv_start(sv, B), E = sv_start(sv, B + 1), t0 = S / MSM_SL, t1 = (E - 1) / MSM_SL;
        Pt acc = pt_identity();
        for (uint32_t t = t0 + split * 128 + threadIdx.x; t <= t1; t += 128 * WIDE_SPLIT) acc = pt_add(acc, ld_ptx32_as_pt(slice_piece(head, tail, S, t, MSM_SL)));
        st_pt30(sh + PT_W * threadIdx.x, acc);
        __syncthreads();
        for (int s = 64; s >= 1; s >>= 1) {
            if ((int)threadIdx.x < s) st_pt30(sh + PT_W * threadIdx.x, pt_add(ld_pt30(sh + PT_W * threadIdx.x), ld_pt30(sh + PT_W * (threadIdx.x + s))));
            __syncthreads();
        }
        if (threadIdx.x == 0) st_pt30(part + PT_W * ((size_t)q * WIDE_SPLIT + split), ld_pt30(sh));
        __syncthreads();
    }
}
__global__ void k_msm_fixup_wide_sum(const uint32_t *part, uint32_t *buckets, SpanQueue sq) {
    uint32_t cnt = *sq.count; if (cnt > sq.cap) cnt = sq.cap;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += gridDim.x * blockDim.x) {
        Pt acc = ld_pt30(part + PT_W * (size_t)q * WIDE_SPLIT);
#pragma unroll 1
        for (uint32_t s = 1; s < WIDE_SPLIT; s++) acc = pt_add(acc, ld_pt30(part + PT_W * ((size_t)q * WIDE_SPLIT + s)));
        st_pt30(buckets + PT_W * (size_t)sq.bucket[q], acc);
    }
}
__global__ void k_pt_fill_identity(uint32_t *pts30, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_pt30(pts30 + PT_W * i, pt_identity());
}
// per window w, chunk j of CH buckets: out[w * nchunks + j] = sum_{b in chunk} (b + 1) B_b
__global__ void __launch_bounds__(64) k_msm_chunks(const uint32_t *buckets, int nwin, uint32_t half, uint32_t CH, uint32_t nchunks, uint32_t *out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nwin * nchunks) return;
    uint32_t w = (uint32_t)(t / nchunks), j = (uint32_t)(t % nchunks);
    uint32_t base = j * CH, top = base + CH < half ? base + CH : half;
    Pt S = pt_identity(), T = pt_identity();
#pragma unroll 1
    for (uint32_t b = top; b-- > base;) {
        S = pt_add(S, ld_pt30(buckets + PT_W * ((size_t)w * half + b)));
        T = pt_add(T, S);
    }
    // + base * S: double-and-add
    Pt BS = pt_identity();
    if (base) {
        BS = S;
#pragma unroll 1
        for (int bit = 30 - __clz(base); bit >= 0; bit--) {     // below the leading one of base (< 2^15)
            BS = pt_double(BS);
            if ((base >> bit) & 1u) BS = pt_add(BS, S);
        }
    }
    st_pt30(out + PT_W * t, pt_add(T, BS));
}
// out[g] = sum_{t < group} in[g * group + t]  (entries beyond count_in are skipped)
__global__ void __launch_bounds__(64) k_pt_sum_groups(const uint32_t *in, size_t count_in, uint32_t group, uint32_t *out, size_t count_out) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count_out) return;
    Pt acc = pt_identity();
#pragma unroll 1
    for (uint32_t t = 0; t < group; t++) {
        size_t idx = g * group + t;
        if (idx < count_in) acc = pt_add(acc, ld_pt30(in + PT_W * idx));
    }
    st_pt30(out + PT_W * g, acc);
}
// out[w] = sum of the per-chunk results of window w (in[w * per + j], j < per): one block per window, a strided partial sum per
// thread and a shared-memory tree -- one launch instead of log16(per) rounds of k_pt_sum_groups (three launches of latency at c = 16)
__global__ void __launch_bounds__(256) k_msm_window_sums(const uint32_t *in, uint32_t per, uint32_t *out) {
    __shared__ uint32_t sh[256 * PT_W];
    const uint32_t w = blockIdx.x;
    Pt acc = pt_identity();
    for (uint32_t j = threadIdx.x; j < per; j += 256) acc = pt_add(acc, ld_pt30(in + PT_W * ((size_t)w * per + j)));
    st_pt30(sh + PT_W * threadIdx.x, acc);
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) st_pt30(sh + PT_W * threadIdx.x, pt_add(ld_pt30(sh + PT_W * threadIdx.x), ld_pt30(sh + PT_W * (threadIdx.x + s))));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_pt30(out + PT_W * (size_t)w, ld_pt30(sh));
}
// result = sum_w 2^(c w) W_w, optionally + *addend
__global__ void k_msm_horner(const uint32_t *win, int c, int nwin, const uint32_t *addend, uint32_t *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    // One thread, 256 dependent doublings: latency is what counts, so the doublings run on the Jacobian formula with the
    // field arithmetic inlined (ptxas overlaps the independent squarings of one doubling).
    Pt acc = ld_pt30(win + PT_W * (nwin - 1));
    for (int w = nwin - 2; w >= 0; w--) {
        if (!fe_normalizes_to_zero(acc.z)) {
            PtJ j;
            j.x = fe_mul(acc.x, acc.z); j.y = fe_mul(acc.y, fe_sqr(acc.z)); j.z = acc.z; j.inf = false;      // (X/Z, Y/Z) = (Xj/Z^2, Yj/Z^3)
#pragma unroll 1
            for (int k = 0; k < c; k++) j = ptj_double_t<true>(j);
            acc = ptj_to_pt(j);
        }
        acc = pt_add(acc, ld_pt30(win + PT_W * w));
    }
    if (addend) acc = pt_add(acc, ld_pt30(addend));
    st_pt30(out, acc);
}
// small n: one GLV scalar multiplication per point
__global__ void __launch_bounds__(64) k_msm_small(const uint32_t *pts, const uint32_t *sc, size_t n, uint32_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PtA q; Sc k;
#pragma unroll
    for (int j = 0; j < 8; j++) k.v[j] = sc[8 * i + j];
    bool ok = load_dev_point(q, pts, i);
    Pt r = pt_mul_glv(pt_from_affine(q, !ok), k);
    st_pt30(out + PT_W * i, r);
}

static std::atomic<uint64_t> g_generic_launches{0};
uint64_t generic_launch_count() { return g_generic_launches.load(); }
#define GL(kern, grid, block, ...) do { kern<<<(grid), (block), 0, st>>>(__VA_ARGS__); g_generic_launches++; } while (0)

// per-device scratch slab that only grows (single host thread per device, like the contexts)
struct Carver { size_t total = 0; size_t take(size_t bytes) { size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; } };
static uint8_t *g_slab[16] = {};
static size_t g_slab_cap[16] = {};
static std::mutex g_slab_mu[16];     // one Pippenger run at a time per device: the slab is shared by every caller in the process
static int scratch_reserve(size_t bytes, uint8_t **out) {
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return fail(BPPP_ERR_ARG, "device index out of range");
    if (g_slab_cap[dev] < bytes) {
        if (g_slab[dev]) { CUDA_OK(cudaDeviceSynchronize()); cudaFree(g_slab[dev]); g_slab[dev] = nullptr; g_slab_cap[dev] = 0; }
        size_t want = bytes + bytes / 4;
        CUDA_OK(cudaMalloc(&g_slab[dev], want));
        g_slab_cap[dev] = want;
    }
    *out = g_slab[dev];
    return BPPP_OK;
}

static int tree_sum(cudaStream_t st, uint32_t *a, uint32_t *b, size_t count, uint32_t **result) {
    // repeatedly sums groups of 16 until one point remains; a holds the input, b is scratch of >= count/16 + 1 points
    uint32_t *in = a, *out = b;
    while (count > 1) {
        size_t nout = (count + 15) / 16;
        GL(k_pt_sum_groups, nblocks(nout, 64), 64, in, count, 16u, out, nout);
        std::swap(in, out); count = nout;
    }
    *result = in;
    return BPPP_OK;
}

int msm_choose_window(size_t n) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c = lg - 3;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    // Measured on a B200 (2^12 .. 2^21 points, widths 10..16): 16 wins from 2^19 points up, 13 below (down to 2^12 points).  The top window of the
    // 129-bit signed GLV halves decides: at c = 16 it holds only the carry, at c = 13 twelve real bits (thousands of buckets), while
    // c = 14 / 15 leave it 2 / 8 bits -- a whole window's entries in a handful of buckets, which serialises the shared-memory
    // cursors of the counting sort.
    if (c >= 10) c = n >= ((size_t)1 << 20) ? 16 : 13;      // n counts the GLV halves: 2^19 points and up take 16
    if (const char *e = getenv("BPPP_MSM_C")) { int v = atoi(e); if (v >= 4 && v <= 16) c = v; }       // experiments
    return c;
}

// d_out30: projective result (PT_W words, device).  d_addend30 may be null.  Synchronises before returning.
int msm_device(cudaStream_t st, const uint32_t *d_pts, const uint32_t *d_sc, size_t n, const uint32_t *d_addend30, uint32_t *d_out30) {
    if (n == 0) {
        if (d_addend30) CUDA_OK(cudaMemcpyAsync(d_out30, d_addend30, PT_BYTES, cudaMemcpyDeviceToDevice, st));
        else GL(k_pt_fill_identity, 1, 1, d_out30, (size_t)1);
        CUDA_OK(cudaStreamSynchronize(st));
        return BPPP_OK;
    }
    if (n <= 1024) {
        uint32_t *a = nullptr, *b = nullptr, *res = nullptr;
        DevScope scope; scope.own(&a); scope.own(&b);
        CUDA_OK(cudaMalloc(&a, PT_BYTES * (n + 1)));
        CUDA_OK(cudaMalloc(&b, PT_BYTES * (n / 16 + 2)));
        GL(k_msm_small, nblocks(n, 64), 64, d_pts, d_sc, n, a);
        size_t count = n;
        if (d_addend30) { CUDA_OK(cudaMemcpyAsync(a + PT_W * n, d_addend30, PT_BYTES, cudaMemcpyDeviceToDevice, st)); count = n + 1; }
        tree_sum(st, a, b, count, &res);
        CUDA_OK(cudaMemcpyAsync(d_out30, res, PT_BYTES, cudaMemcpyDeviceToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));
        CUDA_OK(cudaGetLastError());
        return BPPP_OK;
    }
    const size_t n2 = 2 * n;                   // sorted entries per window: the two GLV halves of every scalar
    const int c = msm_choose_window(n2);
    const int nwin = 128 / c + 1;              // 129 bits of signed digits per half
    const uint32_t half = 1u << (c - 1);
    const uint32_t nb = (uint32_t)nwin * half;
    const size_t total = (size_t)nwin * n2;    // upper bound of the sorted entries (zero digits drop out)
    if (total >= 0x7FFFFFFFull || n2 >= 0x80000000ull) return fail(BPPP_ERR_ARG, "MSM too large for 31-bit point indices");
    uint32_t CH = 16; if (CH > half) CH = half;        // 2 CH sequential additions per thread: short chains, many threads
    const uint32_t nchunks = (half + CH - 1) / CH;
    int dev_for_lock = 0, sms = 148;
    CUDA_OK(cudaGetDevice(&dev_for_lock));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_for_lock);
    uint32_t nch = (uint32_t)(sms / nwin); if (nch < 1) nch = 1;      // sort blocks: nch chunks x nwin windows ~ one per SM
    while (nch > 1 && n2 / nch < 4096) nch /= 2;
    const size_t ncount = (size_t)nb * nch, nscanblk = (ncount + 2047) / 2048;
    // measured on a B200 (2^12 .. 2^20 points): 16-entry slices win below ~0.5 M sorted entries, 32 up to ~8 M, 64 beyond
    const uint32_t MSM_SL = total < (512u << 10) ? 16u : (total < (8u << 20) ? 32u : 64u);
    const size_t nslices = (total + MSM_SL - 1) / MSM_SL;
    SpanQueue sq; sq.cap = (uint32_t)(total / (32 * MSM_SL) + 2);
    // one cached slab per device, carved into the working arrays (cudaMalloc per call cost more than the kernels)
    Carver cv;
    size_t o_digits = cv.take(2 * total), o_vals = cv.take(4 * total), o_counts = cv.take(4 * ncount), o_scan = cv.take(4 * (ncount + 1));
    size_t o_bsum = cv.take(4 * (nscanblk + 1)), o_buckets = cv.take((size_t)PT_BYTES * nb), o_head = cv.take(128 * nslices), o_tail = cv.take(128 * nslices);
    size_t o_sq = cv.take(256 + 4 * (size_t)sq.cap), o_chunks = cv.take((size_t)PT_BYTES * nwin * nchunks);
    size_t o_tmp = cv.take((size_t)PT_BYTES * ((size_t)nwin + 2)), o_bx = cv.take(32 * n);       // tmp: one point per window
    size_t o_wide = cv.take((size_t)PT_BYTES * WIDE_SPLIT * sq.cap);
    std::lock_guard<std::mutex> slab_lock(g_slab_mu[dev_for_lock & 15]);     // released after the final synchronise below
    uint8_t *slab = nullptr;
    int rc = scratch_reserve(cv.total, &slab);
    if (rc != BPPP_OK) return rc;
    uint16_t *digits = (uint16_t *)(slab + o_digits);
    uint32_t *vals = (uint32_t *)(slab + o_vals), *counts = (uint32_t *)(slab + o_counts), *scan = (uint32_t *)(slab + o_scan), *bsum = (uint32_t *)(slab + o_bsum);
    uint32_t *buckets = (uint32_t *)(slab + o_buckets), *head = (uint32_t *)(slab + o_head), *tail = (uint32_t *)(slab + o_tail);
    uint32_t *chunks = (uint32_t *)(slab + o_chunks), *tmp = (uint32_t *)(slab + o_tmp), *bx = (uint32_t *)(slab + o_bx), *wide = (uint32_t *)(slab + o_wide);
    sq.count = (uint32_t *)(slab + o_sq); sq.bucket = sq.count + 64;
    static bool smem_set[16] = {};
    if (!smem_set[dev_for_lock & 15]) {
        CUDA_OK(cudaFuncSetAttribute(k_msm_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << 15));
        CUDA_OK(cudaFuncSetAttribute(k_msm_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << 15));
        smem_set[dev_for_lock & 15] = true;
    }
    CUDA_OK(cudaMemsetAsync(sq.count, 0, 4, st));
    GL(k_msm_digits, nblocks(n, 128), 128, d_sc, n, c, nwin, digits);
    GL(k_msm_endo_x, nblocks(n, 128), 128, d_pts, n, bx);
    k_msm_hist<<<dim3(nch, (unsigned)nwin), 1024, 4 * (size_t)half, st>>>(digits, n2, half, nch, counts); g_generic_launches++;
    GL(k_scan_blocks, (unsigned)nscanblk, 256, counts, ncount, scan, bsum);
    GL(k_scan_sums, 1, 1024, bsum, nscanblk, scan + ncount);
    GL(k_scan_add, (unsigned)nscanblk, 256, scan, ncount, bsum);
    uint32_t nrange = (uint32_t)(sms / nwin); if (nrange < 1) nrange = 1;
    if (nrange > half) nrange = half;
    const uint32_t span = (half + nrange - 1) / nrange;
    k_msm_scatter<<<dim3((half + span - 1) / span, (unsigned)nwin), 1024, 4 * (size_t)span, st>>>(digits, n2, half, span, nch, scan, vals); g_generic_launches++;
    SortedView sv; sv.scan = scan; sv.nch = nch; sv.nb = nb; sv.sl = MSM_SL;
    GL(k_msm_slices, nblocks(nslices, 64), 64, d_pts, bx, vals, sv, buckets, head, tail);
    GL(k_msm_fixup, nblocks(nb, 64), 64, sv, head, tail, buckets, sq);
    GL(k_msm_fixup_wide, dim3(std::min<uint32_t>(sq.cap, 32u), WIDE_SPLIT), 128, sv, head, tail, wide, sq);
    GL(k_msm_fixup_wide_sum, 8, 32, wide, buckets, sq);
    GL(k_msm_chunks, nblocks((size_t)nwin * nchunks, 64), 64, buckets, nwin, half, CH, nchunks, chunks);
    // per-window sum of the chunk results
    uint32_t *in = chunks;
    if (nchunks > 1) { GL(k_msm_window_sums, (unsigned)nwin, 256, chunks, nchunks, tmp); in = tmp; }
    GL(k_msm_horner, 1, 1, in, c, nwin, d_addend30, d_out30);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

int decode_points_to_device(cudaStream_t st, const uint8_t *h_pts, int fmt, size_t n, uint32_t **d_words) {
    size_t psz = fmt == FMT_COMPRESSED ? 33 : 64;
    uint8_t *d_raw = nullptr; int32_t *d_bad = nullptr; uint32_t *d_w = nullptr;
    CUDA_OK(cudaMalloc(&d_raw, psz * (n ? n : 1))); CUDA_OK(cudaMalloc(&d_bad, 4)); CUDA_OK(cudaMalloc(&d_w, 64 * (n ? n : 1)));
    CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, st));
    CUDA_OK(cudaMemcpyAsync(d_raw, h_pts, psz * n, cudaMemcpyHostToDevice, st));
    if (n) GL(k_decode_points, nblocks(n, 64), 64, d_raw, fmt, d_w, d_bad, n);
    int32_t bad = 0;
    CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(d_raw); cudaFree(d_bad);
    if (bad) { cudaFree(d_w); return fail(BPPP_ERR_ENCODING, "a point is not on the curve"); }
    *d_words = d_w;
    return BPPP_OK;
}
int decode_scalars_to_device(cudaStream_t st, const uint8_t *h_sc, size_t n, uint32_t **d_words) {
    uint8_t *d_raw = nullptr; int32_t *d_bad = nullptr; uint32_t *d_w = nullptr;
    CUDA_OK(cudaMalloc(&d_raw, 32 * (n ? n : 1))); CUDA_OK(cudaMalloc(&d_bad, 4)); CUDA_OK(cudaMalloc(&d_w, 32 * (n ? n : 1)));
    CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, st));
    CUDA_OK(cudaMemcpyAsync(d_raw, h_sc, 32 * n, cudaMemcpyHostToDevice, st));
    if (n) GL(k_decode_scalars, nblocks(n, 128), 128, d_raw, d_w, d_bad, n);
    int32_t bad = 0;
    CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(d_raw); cudaFree(d_bad);
    if (bad) { cudaFree(d_w); return fail(BPPP_ERR_ENCODING, "a scalar is not canonical (>= n)"); }
    *d_words = d_w;
    return BPPP_OK;
}
int encode_points_from_device(cudaStream_t st, const uint32_t *d_pts30, size_t n, int fmt, uint8_t *h_out) {
    size_t psz = fmt == FMT_COMPRESSED ? 33 : 64;
    uint8_t *d_raw = nullptr;
    CUDA_OK(cudaMalloc(&d_raw, psz * (n ? n : 1)));
    if (n) GL(k_encode_points, nblocks(n, 64), 64, d_pts30, fmt, d_raw, n);
    CUDA_OK(cudaMemcpyAsync(h_out, d_raw, psz * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(d_raw);
    CUDA_OK(cudaGetLastError());
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

// util::vector_mul<ProjectivePoint> (src/util.rs:46-60) with zero-extension of the shorter operand
extern "C" int bppp_msm(int device, const uint8_t *points, int points_fmt, size_t n_points, const uint8_t *scalars32, size_t n_scalars,
                        int out_fmt, uint8_t *out) {
    if ((n_points && !points) || (n_scalars && !scalars32) || !out) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    size_t n = n_points < n_scalars ? n_points : n_scalars;   // missing terms multiply the identity or zero
    cudaStream_t st = nullptr;
    uint32_t *d_pts = nullptr, *d_sc = nullptr, *d_out = nullptr;
    // decode everything the caller passed so malformed trailing entries are still rejected, as deserialisation would
    rc = decode_points_to_device(st, points, points_fmt, n_points, &d_pts);
    if (rc != BPPP_OK) return rc;
    rc = decode_scalars_to_device(st, scalars32, n_scalars, &d_sc);
    if (rc != BPPP_OK) { cudaFree(d_pts); return rc; }
    CUDA_OK(cudaMalloc(&d_out, PT_BYTES));
    rc = msm_device(st, d_pts, d_sc, n, nullptr, d_out);
    if (rc == BPPP_OK) rc = encode_points_from_device(st, d_out, 1, out_fmt, out);
    cudaFree(d_pts); cudaFree(d_sc); cudaFree(d_out);
    return rc;
}

// Device-resident variant for throughput measurement: points already decoded by bppp_points_upload.
extern "C" int bppp_points_upload(int device, const uint8_t *points, int points_fmt, size_t n, void **handle) {
    if (!handle || (n && !points)) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    uint32_t *d = nullptr;
    rc = decode_points_to_device(nullptr, points, points_fmt, n, &d);
    if (rc == BPPP_OK) *handle = d;
    return rc;
}
extern "C" int bppp_scalars_upload(int device, const uint8_t *scalars32, size_t n, void **handle) {
    if (!handle || (n && !scalars32)) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    uint32_t *d = nullptr;
    rc = decode_scalars_to_device(nullptr, scalars32, n, &d);
    if (rc == BPPP_OK) *handle = d;
    return rc;
}
extern "C" void bppp_device_free(int device, void *handle) { if (handle) { cudaSetDevice(device); cudaFree(handle); } }
// MSM over uploaded arrays; *elapsed_ms (optional) = device time of the MSM alone (CUDA events)
extern "C" int bppp_msm_uploaded(int device, const void *points_handle, const void *scalars_handle, size_t n, int out_fmt, uint8_t *out, float *elapsed_ms) {
    if (!out || (n && (!points_handle || !scalars_handle))) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    uint32_t *d_out = nullptr;
    CUDA_OK(cudaMalloc(&d_out, PT_BYTES));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, nullptr);
    rc = msm_device(nullptr, (const uint32_t *)points_handle, (const uint32_t *)scalars_handle, n, nullptr, d_out);
    cudaEventRecord(e1, nullptr); cudaEventSynchronize(e1);
    if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc == BPPP_OK) rc = encode_points_from_device(nullptr, d_out, 1, out_fmt, out);
    cudaFree(d_out);
    return rc;
}
// SEC1 compressed <-> 64-byte affine conversion of a point array: the device-side counterpart of the
// SerializableProof <-> Proof conversions (src/wnla.rs:41-61, src/circuit.rs:48-76, src/range_proof/reciprocal.rs:43-59)
__global__ void k_words_to_bytes(const uint32_t *words, int fmt, uint8_t *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PtA a; bool ok = load_dev_point(a, words, i);
    if (fmt == FMT_COMPRESSED) pta_compress(out + 33 * i, a, !ok); else pta_to_xy64(out + 64 * i, a, !ok);
}
extern "C" int bppp_points_convert(int device, const uint8_t *in, int in_fmt, size_t n, int out_fmt, uint8_t *out) {
    if (n && (!in || !out)) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    uint32_t *d_w = nullptr; uint8_t *d_out = nullptr;
    rc = decode_points_to_device(nullptr, in, in_fmt, n, &d_w);
    if (rc != BPPP_OK) return rc;
    size_t osz = out_fmt == FMT_COMPRESSED ? 33 : 64;
    CUDA_OK(cudaMalloc(&d_out, osz * (n ? n : 1)));
    if (n) k_words_to_bytes<<<nblocks(n, 128), 128>>>(d_w, out_fmt, d_out, n);
    CUDA_OK(cudaMemcpy(out, d_out, osz * n, cudaMemcpyDeviceToHost));
    cudaFree(d_w); cudaFree(d_out);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

// synthetic generators for large-n measurements: out[i] = base + i * step (affine 64-byte points), computed on the device
__global__ void __launch_bounds__(64) k_points_generate(const uint32_t *two16, size_t n, uint8_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PtA base, step;
    bool okb = load_dev_point(base, two16, 0), oks = load_dev_point(step, two16, 1);
    Pt r = pt_add(pt_from_affine(base, !okb), pt_mul_glv(pt_from_affine(step, !oks), sc_from_u64((uint64_t)i)));
    bool id = pt_is_identity(r);
    PtA a = pt_to_affine_with_zinv(r, fe_inv(r.z));
    pta_to_xy64(out + 64 * i, a, id);
}
extern "C" int bppp_points_generate(int device, const uint8_t *base64, const uint8_t *step64, size_t n, uint8_t *out64) {
    if (!base64 || !step64 || (n && !out64)) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    uint8_t two[128]; memcpy(two, base64, 64); memcpy(two + 64, step64, 64);
    uint32_t *d_two = nullptr; uint8_t *d_out = nullptr;
    rc = decode_points_to_device(nullptr, two, FMT_AFFINE64, 2, &d_two);
    if (rc != BPPP_OK) return rc;
    CUDA_OK(cudaMalloc(&d_out, 64 * (n ? n : 1)));
    if (n) k_points_generate<<<nblocks(n, 64), 64>>>(d_two, n, d_out);
    CUDA_OK(cudaMemcpy(out64, d_out, 64 * n, cudaMemcpyDeviceToHost));
    cudaFree(d_two); cudaFree(d_out);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}

// sum of n points (e.g. the partial sums gathered from the ranks of a split MSM, the shares of X and R of a sharded WNLA round)
__global__ void k_points_sum_small(const uint32_t *pts, size_t n, uint32_t *out30) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Pt acc = pt_identity();
    for (size_t i = 0; i < n; i++) {
        PtA q;
        if (load_dev_point(q, pts, i)) acc = pt_add_mixed(acc, q);
    }
    st_pt30(out30, acc);
}
extern "C" int bppp_points_sum(int device, const uint8_t *points, int points_fmt, size_t n, int out_fmt, uint8_t *out) {
    if (!out || (n && !points)) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    if (n <= 256) {                        // a handful of points: n - 1 additions by one thread instead of n scalar multiplications by one
        uint32_t *d_pts = nullptr, *d_out = nullptr;
        rc = decode_points_to_device(nullptr, points, points_fmt, n, &d_pts);
        if (rc != BPPP_OK) return rc;
        CUDA_OK(cudaMalloc(&d_out, PT_BYTES));
        k_points_sum_small<<<1, 1>>>(d_pts, n, d_out);
        rc = encode_points_from_device(nullptr, d_out, 1, out_fmt, out);
        cudaFree(d_pts); cudaFree(d_out);
        return rc;
    }
    std::vector<uint8_t> ones(32 * n, 0);
    for (size_t i = 0; i < n; i++) ones[32 * i + 31] = 1;
    return bppp_msm(device, points, points_fmt, n, ones.data(), n, out_fmt, out);
}
