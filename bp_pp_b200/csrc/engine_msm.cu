// libbppp.so, variable-base MSM translation unit: sum_i k_i * P_i over arbitrary points, the device-side
// replacement of util::vector_mul<ProjectivePoint> (reference src/util.rs:46-60, which performs one full
// scalar multiplication per term) for large n -- WNLA X/R commitments (src/wnla.rs:152-160), wnla.commit
// (src/wnla.rs:66-72) and the circuit commitments (src/circuit.rs:335-345,469-470,522-524).
//
// Pippenger with signed c-bit windows:
//   k_msm_digits      signed digits of every scalar -> (bucket key, point index | sign) pairs, window-major
//   cub radix sort    pairs by bucket key (the only library call; it moves 8-byte pairs, no curve arithmetic)
//   k_msm_bounds      first / one-past-last sorted position of every bucket
//   k_msm_buckets     one thread per bucket: mixed-adds its points; buckets above the heavy threshold (skewed scalars, the
//                     partially filled top window) are deferred to k_msm_heavy: one 128-thread block per segment of at
//                     most MSM_SEG entries, shared-memory tree reduction; k_msm_heavy_sum adds the segments of a bucket
//   k_msm_chunks      per window: chunked running-sum reduction  sum_b (b+1) B_b  (two adds per bucket)
//   k_pt_sum_groups   tree sums;  k_msm_horner: sum_w 2^(c w) W_w
// Small inputs (n <= 1024) use one GLV scalar multiplication per point and the same tree sum.
#define BPPP_FE_NOINLINE 1
#define BPPP_PTX_ADD_NOINLINE 1   // ec.cuh: the bucket accumulation's XYZZ addition as one call with inlined products
#include "engine_generic.cuh"

#include <cub/device/device_radix_sort.cuh>
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
__global__ void k_msm_digits(const uint32_t *sc, size_t n, int c, int nwin, uint32_t *keys, uint32_t *vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[9];
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = sc[8 * i + j];
    k[8] = 0;
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1u;
    uint32_t carry = 0;
    for (int w = 0; w < nwin; w++) {
        int bit = w * c, word = bit >> 5, sh = bit & 31;
        uint64_t v = word < 8 ? k[word] : 0;
        if (word + 1 < 9) v |= (uint64_t)k[word + 1] << 32;
        uint32_t raw = ((uint32_t)(v >> sh) & mask) + carry;
        uint32_t neg = 0, mag = raw;
        carry = 0;
        if (raw > half) { mag = (1u << c) - raw; neg = 1; carry = 1; }
        keys[(size_t)w * n + i] = mag == 0 ? (uint32_t)nwin * half : (uint32_t)w * half + (mag - 1);     // zero digits: key nb, sorts last
        vals[(size_t)w * n + i] = (uint32_t)i | (neg << 31);
    }
}

// Buckets holding more than `heavy_thr` entries go to the block-per-bucket kernel.  The threshold follows the mean bucket
// size (max(32, 2 n / 2^(c-1))): skewed scalars, and the partially filled top window of every MSM (its few buckets share
// all n points), would otherwise serialise the launch behind a handful of threads.

__global__ void k_msm_bounds(const uint32_t *keys, size_t total, uint32_t nb, uint32_t *start, uint32_t *end) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    uint32_t k = keys[p];
    if (k >= nb) return;
    if (p == 0 || keys[p - 1] != k) start[k] = (uint32_t)p;
    if (p + 1 == total || keys[p + 1] != k) end[k] = (uint32_t)(p + 1);
}
__device__ __forceinline__ Pt msm_accumulate_range(const uint32_t *pts, const uint32_t *vals, uint32_t p0, uint32_t p1, uint32_t stride) {
    PtX acc = ptx_identity();             // XYZZ accumulator: 8 M + 2 S per point, exceptional cases handled exactly
#pragma unroll 1
    for (uint32_t p = p0; p < p1; p += stride) {
        uint32_t v = vals[p];
        PtA q;
        if (load_dev_point(q, pts, v & 0x7FFFFFFFu)) {
            if (v >> 31) q.y = fe_normalize_weak(fe_negate(q.y, 1));
            acc = ptx_add_mixed_hot(acc, q);
        }
    }
    return ptx_to_pt(acc);
}
// bucket sums are stored AoS, PT_W words per point
static constexpr uint32_t MSM_SEG = 2048;     // entries per heavy-bucket work item (16 per thread of a 128-thread block)
// Heavy work items: region M (slots [0, capM)) holds buckets of thr < size <= MSM_SEG, one slot each; region L (slots
// [capM, capM + capL)) holds the segments of larger buckets.  capL = 2 total / MSM_SEG can never overflow (every such bucket
// needs at most size / MSM_SEG + 1 <= 2 size / MSM_SEG slots); a bucket that finds region M full is summed in place.
struct HeavyQueue {
    uint32_t *count;      // [0] region M, [1] region L
    uint32_t *bucket;     // per slot: bucket id
    uint32_t *seg;        // per slot: segment index within its bucket
    uint32_t *seg_out;    // per slot of region L: partial sum (PT_W words)
    uint32_t capM, capL;
};
__global__ void __launch_bounds__(64, 7) k_msm_buckets(const uint32_t *pts, const uint32_t *vals, const uint32_t *start, const uint32_t *end, uint32_t nb,
                                                        uint32_t *buckets, HeavyQueue hq, uint32_t heavy_thr) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t p0 = start[b], p1 = end[b], size = p1 - p0;
    if (size > MSM_SEG) {
        uint32_t nseg = (size + MSM_SEG - 1) / MSM_SEG;
        uint32_t slot = hq.capM + atomicAdd(hq.count + 1, nseg);
        for (uint32_t k = 0; k < nseg; k++) { hq.bucket[slot + k] = (uint32_t)b; hq.seg[slot + k] = k; }
        return;                                            // k_msm_heavy_sum writes the bucket
    }
    if (size > heavy_thr) {
        uint32_t slot = atomicAdd(hq.count, 1u);
        if (slot < hq.capM) { hq.bucket[slot] = (uint32_t)b; hq.seg[slot] = 0; return; }      // k_msm_heavy writes the bucket
    }
    st_pt30(buckets + PT_W * b, msm_accumulate_range(pts, vals, p0, p1, 1));
}
__global__ void __launch_bounds__(128) k_msm_heavy(const uint32_t *pts, const uint32_t *vals, const uint32_t *start, const uint32_t *end, HeavyQueue hq,
                                                    uint32_t *buckets) {
    __shared__ uint32_t sh[128 * PT_W];
    const uint32_t slot = blockIdx.x;
    const bool large = slot >= hq.capM;
    uint32_t cnt = large ? hq.count[1] : hq.count[0];
    if (!large && cnt > hq.capM) cnt = hq.capM;
    if ((large ? slot - hq.capM : slot) >= cnt) return;
    uint32_t b = hq.bucket[slot], k = hq.seg[slot];
    uint32_t p0 = start[b] + k * MSM_SEG, p1 = end[b];
    if (p1 - p0 > MSM_SEG && large) p1 = p0 + MSM_SEG;
    Pt acc = msm_accumulate_range(pts, vals, p0 + threadIdx.x, p1, 128);
    st_pt30(sh + PT_W * threadIdx.x, acc);
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) st_pt30(sh + PT_W * threadIdx.x, pt_add(ld_pt30(sh + PT_W * threadIdx.x), ld_pt30(sh + PT_W * (threadIdx.x + s))));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_pt30(large ? hq.seg_out + PT_W * (size_t)(slot - hq.capM) : buckets + PT_W * (size_t)b, ld_pt30(sh));
}
// region L: the block of a bucket's first segment adds the partial sums of all its segments (they occupy adjacent slots)
__global__ void __launch_bounds__(128) k_msm_heavy_sum(const uint32_t *start, const uint32_t *end, HeavyQueue hq, uint32_t *buckets) {
    __shared__ uint32_t sh[128 * PT_W];
    const uint32_t i = blockIdx.x;                          // index within region L
    if (i >= hq.count[1] || hq.seg[hq.capM + i] != 0) return;
    uint32_t b = hq.bucket[hq.capM + i];
    uint32_t nseg = (end[b] - start[b] + MSM_SEG - 1) / MSM_SEG;
    Pt acc = pt_identity();
    for (uint32_t k = threadIdx.x; k < nseg; k += 128) acc = pt_add(acc, ld_pt30(hq.seg_out + PT_W * (size_t)(i + k)));
    st_pt30(sh + PT_W * threadIdx.x, acc);
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) st_pt30(sh + PT_W * threadIdx.x, pt_add(ld_pt30(sh + PT_W * threadIdx.x), ld_pt30(sh + PT_W * (threadIdx.x + s))));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_pt30(buckets + PT_W * (size_t)b, ld_pt30(sh));
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
        CUDA_OK(cudaMalloc(&a, PT_BYTES * (n + 1)));
        CUDA_OK(cudaMalloc(&b, PT_BYTES * (n / 16 + 2)));
        GL(k_msm_small, nblocks(n, 64), 64, d_pts, d_sc, n, a);
        size_t count = n;
        if (d_addend30) { CUDA_OK(cudaMemcpyAsync(a + PT_W * n, d_addend30, PT_BYTES, cudaMemcpyDeviceToDevice, st)); count = n + 1; }
        tree_sum(st, a, b, count, &res);
        CUDA_OK(cudaMemcpyAsync(d_out30, res, PT_BYTES, cudaMemcpyDeviceToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(a); cudaFree(b);
        CUDA_OK(cudaGetLastError());
        return BPPP_OK;
    }
    const int c = msm_choose_window(n);
    const int nwin = (256 + c) / c;            // 257 bits of signed digits
    const uint32_t half = 1u << (c - 1);
    const uint32_t nb = (uint32_t)nwin * half;
    const size_t total = (size_t)nwin * n;
    uint32_t CH = 16; if (CH > half) CH = half;        // 2 CH sequential additions per thread: short chains, many threads
    const uint32_t nchunks = (half + CH - 1) / CH;
    HeavyQueue hq;
    hq.capM = 8192; hq.capL = (uint32_t)(2 * total / MSM_SEG + 2);
    const size_t hslots = (size_t)hq.capM + hq.capL;
    uint32_t heavy_thr = (uint32_t)(2 * ((n + half - 1) / half)); if (heavy_thr < 32) heavy_thr = 32;
    int key_bits = 1; while (((uint64_t)1 << key_bits) <= (uint64_t)nb) key_bits++;      // keys are 0 .. nb
    size_t cub_bytes = 0;
    CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, total, 0, key_bits, st));
    // one cached slab per device, carved into the working arrays (cudaMalloc per call cost more than the kernels)
    Carver cv;
    size_t o_keys = cv.take(4 * total), o_vals = cv.take(4 * total), o_keys2 = cv.take(4 * total), o_vals2 = cv.take(4 * total);
    size_t o_start = cv.take(4 * (size_t)nb), o_end = cv.take(4 * (size_t)nb), o_buckets = cv.take((size_t)PT_BYTES * nb);
    size_t o_heavy = cv.take(256 + 8 * hslots), o_segout = cv.take((size_t)PT_BYTES * hq.capL), o_chunks = cv.take((size_t)PT_BYTES * nwin * nchunks);
    size_t o_tmp = cv.take((size_t)PT_BYTES * ((size_t)nwin * nchunks / 16 + 2)), o_cub = cv.take(cub_bytes);
    int dev_for_lock = 0;
    CUDA_OK(cudaGetDevice(&dev_for_lock));
    std::lock_guard<std::mutex> slab_lock(g_slab_mu[dev_for_lock & 15]);     // released after the final synchronise below
    uint8_t *slab = nullptr;
    int rc = scratch_reserve(cv.total, &slab);
    if (rc != BPPP_OK) return rc;
    uint32_t *keys = (uint32_t *)(slab + o_keys), *vals = (uint32_t *)(slab + o_vals), *keys2 = (uint32_t *)(slab + o_keys2), *vals2 = (uint32_t *)(slab + o_vals2);
    uint32_t *start = (uint32_t *)(slab + o_start), *end = (uint32_t *)(slab + o_end), *buckets = (uint32_t *)(slab + o_buckets);
    uint32_t *chunks = (uint32_t *)(slab + o_chunks), *tmp = (uint32_t *)(slab + o_tmp);
    hq.count = (uint32_t *)(slab + o_heavy); hq.bucket = hq.count + 64; hq.seg = hq.bucket + hslots; hq.seg_out = (uint32_t *)(slab + o_segout);
    void *cub_tmp = slab + o_cub;
    CUDA_OK(cudaMemsetAsync(start, 0, 4 * (size_t)nb, st));          // empty buckets keep start == end == 0
    CUDA_OK(cudaMemsetAsync(end, 0, 4 * (size_t)nb, st));
    CUDA_OK(cudaMemsetAsync(hq.count, 0, 8, st));
    GL(k_msm_digits, nblocks(n, 128), 128, d_sc, n, c, nwin, keys, vals);
    CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys2, vals, vals2, total, 0, key_bits, st));   // only the bits a key can have
    g_generic_launches += 4;
    GL(k_msm_bounds, nblocks(total, 256), 256, keys2, total, nb, start, end);
    GL(k_msm_buckets, nblocks(nb, 64), 64, d_pts, vals2, start, end, nb, buckets, hq, heavy_thr);
    GL(k_msm_heavy, (unsigned)hslots, 128, d_pts, vals2, start, end, hq, buckets);
    GL(k_msm_heavy_sum, hq.capL, 128, start, end, hq, buckets);
    GL(k_msm_chunks, nblocks((size_t)nwin * nchunks, 64), 64, buckets, nwin, half, CH, nchunks, chunks);
    // per-window sum of chunk results: groups of 16 until nwin points remain
    uint32_t *in = chunks, *out = tmp;
    size_t per = nchunks;
    while (per > 1) {
        uint32_t group = per >= 16 ? 16u : (uint32_t)per;
        size_t nout = (size_t)nwin * (per / group);
        GL(k_pt_sum_groups, nblocks(nout, 64), 64, in, (size_t)nwin * per, group, out, nout);
        std::swap(in, out); per /= group;
    }
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
    if (bad) { cudaFree(d_w); return fail(BPPP_ERR_ARG, "a point is not on the curve"); }
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
    if (bad) { cudaFree(d_w); return fail(BPPP_ERR_ARG, "a scalar is not canonical (>= n)"); }
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

// sum of n points (e.g. the partial sums gathered from the ranks of a split MSM)
extern "C" int bppp_points_sum(int device, const uint8_t *points, int points_fmt, size_t n, int out_fmt, uint8_t *out) {
    if (!out || (n && !points)) return fail(BPPP_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if (rc != BPPP_OK) return rc;
    std::vector<uint8_t> ones(32 * (n ? n : 1), 0);
    for (size_t i = 0; i < n; i++) ones[32 * i + 31] = 1;
    return bppp_msm(device, points, points_fmt, n, ones.data(), n, out_fmt, out);
}
