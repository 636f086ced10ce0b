// libbppp.so, multi-GPU translation unit: one PROCESS driving several GPUs (what a Rust or C caller of prove_batch /
// verify_batch gets on an 8 x B200 box without torchrun).  A bppp_multi_ctx owns one bppp_ctx per listed device --
// generators and window tables replicated, SURVEY 8(e) -- and every batch call cuts [0, n) into contiguous per-device
// ranges run by one host thread per device.  Proofs are independent (the reference has no shared mutable state,
// SURVEY 8b "Threading"), so there is no data-path collective and the output for proof i does not depend on the
// number of devices.
#include <algorithm>
#include <functional>
#include <thread>

#include <map>

#include "engine_generic.cuh"

using namespace bppp;

// ---- caching device allocator (declared in engine_generic.cuh; this file calls the real cudaMalloc / cudaFree) ----
namespace {
struct DevCache {
    std::multimap<size_t, void *> parked;          // size class -> block
    size_t parked_bytes = 0;
};
struct Live { int dev; size_t cls; };
std::mutex g_cache_mu;                             // one lock: the calls are rare next to the kernels they feed
DevCache g_cache[16];
std::map<void *, Live> g_live;                     // block -> (device it was allocated on, size class)
constexpr size_t CACHE_CAP = (size_t)24 << 30;     // parked bytes per device before blocks go back to CUDA
// size classes: eight per octave (at most 12.5 % slack), 512-byte granularity at the bottom
size_t size_class(size_t n) {
    if (n < 512) return 512;
    size_t p = 512;
    while (p * 2 <= n) p *= 2;                     // p <= n < 2p
    size_t step = p / 8;
    return (n + step - 1) / step * step;
}
}  // namespace
namespace bppp {
cudaError_t dev_malloc(void **out, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t cls = size_class(bytes);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        DevCache &c = g_cache[dev & 15];
        auto it = c.parked.find(cls);
        if (it != c.parked.end()) {
            *out = it->second; g_live[*out] = Live{dev, cls}; c.parked_bytes -= cls; c.parked.erase(it);
            return cudaSuccess;
        }
    }
    e = cudaMalloc(out, cls);
    if (e == cudaErrorMemoryAllocation) { (void)cudaGetLastError(); dev_trim(dev); e = cudaMalloc(out, cls); }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(g_cache_mu); g_live[*out] = Live{dev, cls}; }
    return e;
}
cudaError_t dev_free(void *p) {
    if (!p) return cudaSuccess;
    Live l{};
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_live.find(p);
        if (it == g_live.end()) return cudaFree(p);          // not ours (allocated by a unit that calls CUDA directly)
        l = it->second; g_live.erase(it);
    }
    // cudaFree's contract: nothing in flight still uses the block.  The block belongs to the device it was allocated on,
    // whatever device is current in the calling thread.
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != l.dev) cudaSetDevice(l.dev);
    cudaError_t e = cudaDeviceSynchronize();
    bool park;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        DevCache &c = g_cache[l.dev & 15];
        park = c.parked_bytes + l.cls <= CACHE_CAP;
        if (park) { c.parked.emplace(l.cls, p); c.parked_bytes += l.cls; }
    }
    if (!park) e = cudaFree(p);
    if (cur != l.dev) cudaSetDevice(cur);
    return e;
}
void dev_trim(int device) {
    std::multimap<size_t, void *> blocks;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        DevCache &c = g_cache[device & 15];
        blocks.swap(c.parked); c.parked_bytes = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    for (auto &kv : blocks) cudaFree(kv.second);
    if (cur != device) cudaSetDevice(cur);
}
}  // namespace bppp

struct bppp_multi_ctx {
    std::vector<bppp_ctx *> ctx;
};

static int mfail(int code, const std::string &msg) { return engine_fail(code, msg); }

// run fn(k) on one host thread per device; the first failure (lowest device) wins and its message becomes this thread's last error
static int for_each_device(size_t ndev, const std::function<int(size_t)> &fn) {
    std::vector<int> rc(ndev, BPPP_OK);
    std::vector<std::string> err(ndev);
    std::vector<std::thread> th;
    for (size_t k = 0; k < ndev; k++)
        th.emplace_back([&, k] { rc[k] = fn(k); if (rc[k] != BPPP_OK) err[k] = bppp_last_error(); });
    for (auto &t : th) t.join();
    for (size_t k = 0; k < ndev; k++) if (rc[k] != BPPP_OK) return mfail(rc[k], "device slot " + std::to_string(k) + ": " + err[k]);
    return BPPP_OK;
}
static inline size_t cut(size_t n, size_t k, size_t parts) { return n * k / parts; }

extern "C" int bppp_multi_ctx_create(bppp_multi_ctx **out, const int *devices, int ndev, const uint8_t *gens64, int window_bits, size_t max_batch_per_device) {
    if (!out || !devices || ndev < 1 || !gens64) return mfail(BPPP_ERR_ARG, "null argument or empty device list");
    *out = nullptr;
    bppp_multi_ctx *m = new bppp_multi_ctx();
    m->ctx.assign((size_t)ndev, nullptr);
    int rc = for_each_device((size_t)ndev, [&](size_t k) { return bppp_ctx_create(&m->ctx[k], devices[k], gens64, window_bits, max_batch_per_device); });
    if (rc != BPPP_OK) {
        for (auto *c : m->ctx) bppp_ctx_destroy(c);
        delete m;
        return rc;
    }
    *out = m;
    return BPPP_OK;
}
extern "C" void bppp_multi_ctx_destroy(bppp_multi_ctx *m) {
    if (!m) return;
    for (auto *c : m->ctx) bppp_ctx_destroy(c);
    delete m;
}
extern "C" int bppp_multi_device_count(const bppp_multi_ctx *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" bppp_ctx *bppp_multi_ctx_get(const bppp_multi_ctx *m, int k) { return (m && k >= 0 && (size_t)k < m->ctx.size()) ? m->ctx[(size_t)k] : nullptr; }

extern "C" int bppp_multi_u64_commit_batch(bppp_multi_ctx *m, size_t n, const uint64_t *x, const uint8_t *blinds32, int fmt, uint8_t *out) {
    if (!m || (n && (!x || !blinds32 || !out))) return mfail(BPPP_ERR_ARG, "null argument");
    const size_t D = m->ctx.size(), osz = fmt == FMT_COMPRESSED ? 33 : 64;
    return for_each_device(D, [&](size_t k) {
        size_t lo = cut(n, k, D), cnt = cut(n, k + 1, D) - lo;
        return cnt ? bppp_u64_commit_batch(m->ctx[k], cnt, x + lo, blinds32 + 32 * lo, fmt, out + osz * lo) : BPPP_OK;
    });
}
extern "C" int bppp_multi_u64_verify_batch(bppp_multi_ctx *m, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt,
                                           const uint8_t *label, size_t label_len, int32_t *status) {
    if (!m || (n && (!commits || !proofs || !status))) return mfail(BPPP_ERR_ARG, "null argument");
    const size_t D = m->ctx.size();
    const size_t csz = fmt == FMT_COMPRESSED ? 33 : 64, psz = fmt == FMT_COMPRESSED ? U64_PROOF_BYTES_COMPRESSED : U64_PROOF_BYTES_AFFINE;
    return for_each_device(D, [&](size_t k) {
        size_t lo = cut(n, k, D), cnt = cut(n, k + 1, D) - lo;
        return cnt ? bppp_u64_verify_batch(m->ctx[k], cnt, commits + csz * lo, proofs + psz * lo, fmt, label, label_len, status + lo) : BPPP_OK;
    });
}
extern "C" int bppp_multi_u64_prove_batch(bppp_multi_ctx *m, size_t n, const uint64_t *x, const uint8_t *blinds32, const uint8_t *rng,
                                          const uint8_t *label, size_t label_len, uint8_t *proofs_out, int32_t *status) {
    if (!m || (n && (!x || !blinds32 || !rng || !proofs_out || !status))) return mfail(BPPP_ERR_ARG, "null argument");
    const size_t D = m->ctx.size();
    return for_each_device(D, [&](size_t k) {
        size_t lo = cut(n, k, D), cnt = cut(n, k + 1, D) - lo;
        return cnt ? bppp_u64_prove_batch(m->ctx[k], cnt, x + lo, blinds32 + 32 * lo, rng + (size_t)U64_RNG_BYTES * lo, label, label_len,
                                          proofs_out + (size_t)U64_PROOF_BYTES_COMPRESSED * lo, status + lo) : BPPP_OK;
    });
}
