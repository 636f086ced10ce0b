// libbppp.so, peer-exchange translation unit: the exchange steps of the path done by our own kernels over NVLink / NVSwitch
// peer memory instead of a collective library.
//
// The path has exactly two exchanges: the partial sums of a point-range-split MSM (util::vector_mul over a
// generator vector too long for one GPU's share of the time, src/util.rs:46-60) and the per-round shares of X and R of a
// block-sharded WeightNormLinearArgument::prove (src/wnla.rs:152-160).  Both are a few hundred bytes per rank, so what
// matters is latency, not bandwidth: every rank owns a MAILBOX in its HBM (one 256-byte slot per rank and parity), the
// producer's tail kernel stores its payload straight into the slot it owns in EVERY peer's mailbox (remote stores over
// NVLink), fences at system scope and then publishes the epoch number in the slot's flag word; the consumer kernel on
// each GPU spins on its local flags and reduces / copies the slots.  No host round trip, no NCCL launch: the exchange is
// two tiny kernels stream-ordered behind the MSM.
//
// One process per GPU (torchrun): mailboxes are shared through CUDA IPC handles which the caller all-gathers once at
// set-up (bp_pp_b200/shard.py: PeerGroup).  Ranks call the collective entry points in the same order; the epoch counter
// (one per call) and two slot parities make a slot reusable as soon as its owner has finished the previous collective.
// A peer that never arrives ends the wait after BPPP_PEER_TIMEOUT_NS with BPPP_ERR_CUDA instead of hanging the GPU.
#define BPPP_FE_NOINLINE 1
#include "engine_generic.cuh"

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

static constexpr int PEER_MAX = 16;
static constexpr int SLOT_WORDS = 64;                  // 256 bytes: payload words 0..59, flag = word 63
static constexpr int SLOT_PAYLOAD_WORDS = 60;
static constexpr unsigned long long BPPP_PEER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

struct PeerPtrs { uint32_t *p[PEER_MAX]; };

struct bppp_peer {
    int device = 0, world = 1, rank = 0;
    uint32_t *mine = nullptr;                          // 2 parities x world slots
    PeerPtrs peers{};                                  // peers.p[r]: rank r's mailbox mapped into this process (p[rank] = mine)
    bool opened[PEER_MAX] = {};
    bool connected = false;
    uint32_t epoch = 0;
    uint32_t *d_part = nullptr, *d_out = nullptr;      // PT_W words each / world * SLOT_PAYLOAD_WORDS
    int32_t *d_err = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

__device__ __forceinline__ unsigned long long peer_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// store `words` payload words into this rank's slot of every mailbox, then publish the epoch (one warp)
__global__ void k_peer_post(PeerPtrs pp, int world, int rank, const uint32_t *payload, int words, uint32_t epoch) {
    const int lane = threadIdx.x;
    const size_t slot = ((size_t)(epoch & 1u) * world + rank) * SLOT_WORDS;
    for (int r = 0; r < world; r++) {
        for (int k = lane; k < words; k += 32) pp.p[r][slot + k] = payload[k];
    }
    __threadfence_system();
    __syncwarp();
    __threadfence_system();
    if (lane < world) {
        volatile uint32_t *flag = pp.p[lane] + slot + (SLOT_WORDS - 1);
        *flag = epoch;
    }
}
// wait for every rank's slot of this epoch; returns false on timeout
__device__ __forceinline__ bool peer_wait_slot(const uint32_t *mine, int world, int r, uint32_t epoch) {
    const volatile uint32_t *flag = mine + ((size_t)(epoch & 1u) * world + r) * SLOT_WORDS + (SLOT_WORDS - 1);
    const unsigned long long t0 = peer_now_ns();
    while (*flag != epoch) {
        if (peer_now_ns() - t0 > BPPP_PEER_TIMEOUT_NS) return false;
        __nanosleep(200);
    }
    __threadfence_system();
    return true;
}
// out30 = sum over ranks (in rank order) of the projective points in the slots
__global__ void k_peer_sum_points(const uint32_t *mine, int world, uint32_t epoch, uint32_t *out30, int32_t *err) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Pt acc = pt_identity();
    for (int r = 0; r < world; r++) {
        if (!peer_wait_slot(mine, world, r, epoch)) { *err = 1; return; }
        const volatile uint32_t *s = mine + ((size_t)(epoch & 1u) * world + r) * SLOT_WORDS;
        Pt p;
#pragma unroll
        for (int k = 0; k < FE_W; k++) { p.x.v[k] = s[k]; p.y.v[k] = s[FE_W + k]; p.z.v[k] = s[2 * FE_W + k]; }
        acc = pt_add(acc, p);
    }
#pragma unroll
    for (int k = 0; k < FE_W; k++) { out30[k] = acc.x.v[k]; out30[FE_W + k] = acc.y.v[k]; out30[2 * FE_W + k] = acc.z.v[k]; }
}
// out[r * words + k] = payload word k of rank r (one warp)
__global__ void k_peer_gather(const uint32_t *mine, int world, uint32_t epoch, int words, uint32_t *out, int32_t *err) {
    const int lane = threadIdx.x;
    for (int r = 0; r < world; r++) {
        bool ok = true;
        if (lane == 0) ok = peer_wait_slot(mine, world, r, epoch);
        ok = __shfl_sync(0xFFFFFFFFu, ok ? 1 : 0, 0) != 0;
        if (!ok) { if (lane == 0) *err = 1; return; }
        const volatile uint32_t *s = mine + ((size_t)(epoch & 1u) * world + r) * SLOT_WORDS;
        for (int k = lane; k < words; k += 32) out[(size_t)r * words + k] = s[k];
    }
}

extern "C" void bppp_peer_destroy(bppp_peer *p);
extern "C" int bppp_peer_create(bppp_peer **out, int device, int world, int rank, uint8_t *ipc_handle64_out) {
    if (!out || !ipc_handle64_out || world < 1 || world > PEER_MAX || rank < 0 || rank >= world) return fail(BPPP_ERR_ARG, "bad peer group arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BPPP_ERR_ARG, "bad device index");
    CUDA_OK(cudaSetDevice(device));
    bppp_peer *p = new bppp_peer();
    p->device = device; p->world = world; p->rank = rank;
    const size_t bytes = (size_t)2 * world * SLOT_WORDS * sizeof(uint32_t);
    auto bail = [&](int rc) { bppp_peer_destroy(p); return rc; };        // frees whatever was created so far
#define PEER_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { (void)cudaGetLastError(); \
        return bail(fail(_e == cudaErrorMemoryAllocation ? BPPP_ERR_NOMEM : BPPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); } } while (0)
    PEER_OK(cudaMalloc(&p->mine, bytes));
    PEER_OK(cudaMemset(p->mine, 0, bytes));
    PEER_OK(cudaMalloc(&p->d_part, PT_BYTES > SLOT_PAYLOAD_WORDS * 4 ? PT_BYTES : SLOT_PAYLOAD_WORDS * 4));
    PEER_OK(cudaMalloc(&p->d_out, (size_t)world * SLOT_PAYLOAD_WORDS * 4));
    PEER_OK(cudaMalloc(&p->d_err, 4));
    PEER_OK(cudaMemset(p->d_err, 0, 4));
    PEER_OK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PEER_OK(cudaEventCreate(&p->e0));
    PEER_OK(cudaEventCreate(&p->e1));
    cudaIpcMemHandle_t h;
    PEER_OK(cudaIpcGetMemHandle(&h, p->mine));
    PEER_OK(cudaDeviceSynchronize());
#undef PEER_OK
    memcpy(ipc_handle64_out, &h, 64);
    p->peers.p[rank] = p->mine;
    if (world == 1) p->connected = true;
    *out = p;
    return BPPP_OK;
}
// handles: world x 64 bytes in rank order (this rank's own entry is ignored)
extern "C" int bppp_peer_connect(bppp_peer *p, const uint8_t *handles) {
    if (!p || (!handles && p->world > 1)) return fail(BPPP_ERR_ARG, "null argument");
    CUDA_OK(cudaSetDevice(p->device));
    for (int r = 0; r < p->world; r++) {
        if (r == p->rank || p->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->peers.p[r] = (uint32_t *)ptr;
        p->opened[r] = true;
    }
    p->connected = true;
    return BPPP_OK;
}
extern "C" void bppp_peer_destroy(bppp_peer *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; r++) if (p->opened[r]) cudaIpcCloseMemHandle(p->peers.p[r]);
    cudaFree(p->mine); cudaFree(p->d_part); cudaFree(p->d_out); cudaFree(p->d_err);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->e0) cudaEventDestroy(p->e0);
    if (p->e1) cudaEventDestroy(p->e1);
    delete p;
}
static int peer_check_err(bppp_peer *p) {
    int32_t e = 0;
    CUDA_OK(cudaMemcpyAsync(&e, p->d_err, 4, cudaMemcpyDeviceToHost, p->stream));
    CUDA_OK(cudaStreamSynchronize(p->stream));
    if (e) {
        cudaMemsetAsync(p->d_err, 0, 4, p->stream);
        return fail(BPPP_ERR_CUDA, "peer exchange timed out: a rank of the group did not reach the collective");
    }
    return BPPP_OK;
}

// sum over ALL ranks of (sum_i scalars[i] * points[i]) over each rank's uploaded block (bppp_points_upload / bppp_scalars_upload):
// the block MSM, the remote stores of its partial sum and the reduction of the world's partial sums run back to back on one
// stream.  Every rank gets the same encoded point.  *elapsed_ms (optional): device time from the first MSM kernel to the reduced sum.
extern "C" int bppp_peer_msm_allsum(bppp_peer *p, const void *points_handle, const void *scalars_handle, size_t n, int out_fmt, uint8_t *out, float *elapsed_ms) {
    if (!p || !out || (n && (!points_handle || !scalars_handle))) return fail(BPPP_ERR_ARG, "null argument");
    if (!p->connected) return fail(BPPP_ERR_ARG, "bppp_peer_connect has not been called");
    CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t st = p->stream;
    const uint32_t epoch = ++p->epoch;
    CUDA_OK(cudaEventRecord(p->e0, st));
    int rc = msm_device(st, (const uint32_t *)points_handle, (const uint32_t *)scalars_handle, n, nullptr, p->d_part);
    if (rc != BPPP_OK) return rc;
    k_peer_post<<<1, 32, 0, st>>>(p->peers, p->world, p->rank, p->d_part, PT_W, epoch);
    k_peer_sum_points<<<1, 1, 0, st>>>(p->mine, p->world, epoch, p->d_out, p->d_err);
    CUDA_OK(cudaEventRecord(p->e1, st));
    rc = peer_check_err(p);
    if (rc != BPPP_OK) return rc;
    if (elapsed_ms) CUDA_OK(cudaEventElapsedTime(elapsed_ms, p->e0, p->e1));
    return encode_points_from_device(st, p->d_out, 1, out_fmt, out);
}

// all-gather of one short byte string per rank (<= 240 bytes, a multiple of 4): out = world x bytes in rank order
extern "C" int bppp_peer_allgather(bppp_peer *p, const uint8_t *in, size_t bytes, uint8_t *out) {
    if (!p || !in || !out) return fail(BPPP_ERR_ARG, "null argument");
    if (bytes == 0 || bytes > (size_t)SLOT_PAYLOAD_WORDS * 4 || (bytes & 3)) return fail(BPPP_ERR_ARG, "payload must be 4..240 bytes, a multiple of 4");
    if (!p->connected) return fail(BPPP_ERR_ARG, "bppp_peer_connect has not been called");
    CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t st = p->stream;
    const uint32_t epoch = ++p->epoch;
    const int words = (int)(bytes / 4);
    CUDA_OK(cudaMemcpyAsync(p->d_part, in, bytes, cudaMemcpyHostToDevice, st));
    k_peer_post<<<1, 32, 0, st>>>(p->peers, p->world, p->rank, p->d_part, words, epoch);
    k_peer_gather<<<1, 32, 0, st>>>(p->mine, p->world, epoch, words, p->d_out, p->d_err);
    CUDA_OK(cudaMemcpyAsync(out, p->d_out, bytes * (size_t)p->world, cudaMemcpyDeviceToHost, st));
    return peer_check_err(p);
}
extern "C" int bppp_peer_world(const bppp_peer *p) { return p ? p->world : 0; }
extern "C" int bppp_peer_rank(const bppp_peer *p) { return p ? p->rank : -1; }
