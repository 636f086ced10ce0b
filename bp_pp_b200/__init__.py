"""bp_pp_b200 -- B200-native Bulletproofs++ engine behind the reference's API surface.

(The distribution is called `bp-pp_b200`; a Python package name cannot contain a hyphen.)

Host-side mirror of distributed-lab/bp-pp's public items for the u64 range-proof hot path:
`U64RangeProofProtocol::{commit_value, prove, verify}` (src/range_proof/u64_proof.rs:37-82) plus the
batch entry points `commit_batch` / `prove_batch` / `verify_batch` over N independent proofs.  All
curve, field, transcript and protocol arithmetic runs on the GPU inside libbppp.so (CUDA, sm_100a)
through the C ABI declared in include/bppp.h; this module only marshals bytes.
"""
from .api import (BpppError, Context, MultiContext, U64RangeProofProtocol, FMT_AFFINE64, FMT_COMPRESSED, G_VEC_FULL_SZ,
                  H_VEC_CIRCUIT_SZ, H_VEC_FULL_SZ, ST_BAD_POINT, ST_BAD_SCALAR, ST_FALSE, ST_PANIC_CHALLENGE_RANGE,
                  ST_PANIC_INVERT_ZERO, ST_TRUE, U64_PROOF_BYTES, U64_RNG_BYTES, ArithmeticCircuit, ReciprocalRangeProofProtocol, UploadedMsm, WeightNormLinearArgument, microbench, msm, points_convert, points_generate, u64_proofs_to_affine,
                  points_sum)

__all__ = ["BpppError", "Context", "MultiContext", "U64RangeProofProtocol", "FMT_AFFINE64", "FMT_COMPRESSED", "G_VEC_FULL_SZ",
           "H_VEC_CIRCUIT_SZ", "H_VEC_FULL_SZ", "ST_BAD_POINT", "ST_BAD_SCALAR", "ST_FALSE",
           "ST_PANIC_CHALLENGE_RANGE", "ST_PANIC_INVERT_ZERO", "ST_TRUE", "U64_PROOF_BYTES", "U64_RNG_BYTES",
           "microbench", "msm", "points_sum", "UploadedMsm", "WeightNormLinearArgument", "ReciprocalRangeProofProtocol", "ArithmeticCircuit", "points_generate", "points_convert", "u64_proofs_to_affine"]
