// Shared declarations of the generic (arbitrary-size, single-instance) translation units:
// engine_msm.cu (variable-base Pippenger MSM) and engine_wnla.cu (WNLA / circuit / reciprocal protocols).
#pragma once
#include <atomic>
#include "engine_common.cuh"

namespace bppp {

// Device arrays used by the generic paths:
//   points  : AoS, 16 canonical words per point (x[8], y[8] little-endian), all-zero = identity
//   scalars : AoS, 8 canonical little-endian words per scalar
//   pt30    : projective point, PT_W words (x, y, z as 8 x 32-bit limbs; the name is historical)
int msm_choose_window(size_t n);
static constexpr size_t PT_BYTES = PT_W * sizeof(uint32_t);
int msm_device(cudaStream_t st, const uint32_t *d_pts, const uint32_t *d_sc, size_t n, const uint32_t *d_addend30, uint32_t *d_out30);
int decode_points_to_device(cudaStream_t st, const uint8_t *h_pts, int fmt, size_t n, uint32_t **d_words);
int decode_scalars_to_device(cudaStream_t st, const uint8_t *h_sc, size_t n, uint32_t **d_words);
int encode_points_from_device(cudaStream_t st, const uint32_t *d_pts30, size_t n, int fmt, uint8_t *h_out);
uint64_t generic_launch_count();

// engine_wnla.cu -- device-resident WNLA instance.  pts = [H (Lh) | G (Lg) | g] affine words; c scalar words.
struct WnlaDev {
    size_t Lh = 0, Lg = 0;            // padded lengths
    size_t len_h = 0, len_g = 0;      // true generator lengths (verify absorbs these, wnla.rs:91-92)
    uint32_t *pts = nullptr, *c = nullptr;
    Sc rho, mu;
    // The stored G points are the true generators divided by sigma (wnla_prove_dev: folding g' = rho g0 + y g1 as
    // sigma' = rho sigma, G' = G0 + (y / rho) G1 costs one scalar multiplication per output instead of two; the factor rides on the
    // MSM scalars of the G part).  scaled = false means sigma = 1.
    bool scaled = false;
    Sc sigma;
    void release() { cudaFree(pts); cudaFree(c); pts = c = nullptr; }
};
struct WnlaProofHost { std::vector<uint8_t> r33, x33, l32, n32; };   // r/x in push order (innermost round first)
int upload_padded_scalars(cudaStream_t st, const uint8_t *h32, size_t n, size_t L, uint32_t **d);
int wnla_load(cudaStream_t st, WnlaDev &w, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn, const uint8_t *c32, size_t cn,
              const uint8_t *rho32, const uint8_t *mu32, size_t ln, size_t nn);
int wnla_commit_dev(cudaStream_t st, const WnlaDev &w, const uint32_t *d_l, const uint32_t *d_n, uint32_t *d_out30);
int wnla_prove_dev(cudaStream_t st, WnlaDev &w, Merlin &t, uint32_t *d_com30, uint32_t *d_l, uint32_t *d_n, size_t len_l, size_t len_n, WnlaProofHost &proof,
                   int32_t *status);
int wnla_verify_dev(cudaStream_t st, WnlaDev &w, Merlin &t, uint32_t *d_com30, const uint8_t *r33, size_t rn, const uint8_t *x33, size_t xn,
                    const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn, int32_t *verdict);
int point_bytes_to_pt30(cudaStream_t st, const uint8_t *p, int fmt, uint32_t *d_out30);
uint64_t wnla_launch_count();

// Caching device allocator of the generic paths (implemented in engine_multi.cu, which keeps the real calls).  One WNLA prove makes dozens of short-lived device buffers;
// cudaMalloc / cudaFree each cost a device-wide map / unmap (far worse once peer access is enabled: every allocation is mapped
// into the peers), which dominated the wall time next to ~0.15 s of kernels.  dev_free keeps cudaFree's contract -- it waits
// for the device before the block can be handed out again -- but parks the block in a per-device size-class list.
cudaError_t dev_malloc(void **p, size_t bytes);
cudaError_t dev_free(void *p);
void dev_trim(int device);            // return every parked block of the device to CUDA
// frees the registered device pointers when the scope ends, whatever path leaves it (CUDA_OK returns early on errors)
struct DevScope {
    std::vector<void **> slots;
    template <typename T> void own(T **p) { slots.push_back(reinterpret_cast<void **>(p)); }
    ~DevScope() { for (void **p : slots) { if (*p) dev_free(*p); *p = nullptr; } }
};

}  // namespace bppp

// the generic translation units allocate through the cache
#if defined(BPPP_GENERIC_ALLOC)
#define cudaMalloc(p, n) bppp::dev_malloc((void **)(p), (n))
#define cudaFree(p) bppp::dev_free((void *)(p))
#endif
