// Shared declarations of the generic (arbitrary-size, single-instance) translation units:
// engine_msm.cu (variable-base Pippenger MSM) and engine_wnla.cu (WNLA / circuit / reciprocal protocols).
#pragma once
#include <atomic>
#include "engine_common.cuh"

namespace bppp {

// Device arrays used by the generic paths:
//   points  : AoS, 16 canonical words per point (x[8], y[8] little-endian), all-zero = identity
//   scalars : AoS, 8 canonical little-endian words per scalar
//   pt30    : projective point, 30 words (x, y, z as 10x26 limbs)
int msm_choose_window(size_t n);
int msm_device(cudaStream_t st, const uint32_t *d_pts, const uint32_t *d_sc, size_t n, const uint32_t *d_addend30, uint32_t *d_out30);
int decode_points_to_device(cudaStream_t st, const uint8_t *h_pts, int fmt, size_t n, uint32_t **d_words);
int decode_scalars_to_device(cudaStream_t st, const uint8_t *h_sc, size_t n, uint32_t **d_words);
int encode_points_from_device(cudaStream_t st, const uint32_t *d_pts30, size_t n, int fmt, uint8_t *h_out);
uint64_t generic_launch_count();

}  // namespace bppp
