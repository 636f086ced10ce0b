"""ctypes loader for libbppp.so (the CUDA engine + C ABI of include/bppp.h).

There is deliberately no CPU fallback: if the shared library is missing, or no CUDA device is
present when a context is created, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BPPP_LIB") or os.path.join(_HERE, "libbppp.so")   # BPPP_LIB: kernel-variant experiments

EXPORTS = [
    "bppp_ctx_create", "bppp_ctx_create_shared", "bppp_ctx_set_inflight", "bppp_ctx_destroy", "bppp_last_error", "bppp_ctx_info", "bppp_u64_commit_batch",
    "bppp_u64_verify_batch", "bppp_u64_verify_batch_dev", "bppp_u64_prove_batch", "bppp_u64_prove_batch_dev",
    "bppp_u64_verify_begin", "bppp_u64_verify_circuit", "bppp_u64_verify_round", "bppp_u64_verify_finish",
    "bppp_u64_prove_begin", "bppp_u64_prove_reciprocal", "bppp_u64_prove_circuit", "bppp_u64_prove_tau", "bppp_u64_prove_round",
    "bppp_u64_prove_finish", "bppp_u64_step_abort",
    "bppp_multi_ctx_create", "bppp_multi_ctx_destroy", "bppp_multi_device_count", "bppp_multi_ctx_get", "bppp_multi_u64_commit_batch",
    "bppp_multi_u64_verify_batch", "bppp_multi_u64_prove_batch",
    "bppp_launch_count", "bppp_microbench", "bppp_ctx_profile_begin", "bppp_ctx_profile_end",
    "bppp_msm", "bppp_points_upload", "bppp_scalars_upload", "bppp_device_free", "bppp_msm_uploaded", "bppp_points_sum", "bppp_points_generate", "bppp_points_convert",
    "bppp_wnla_commit", "bppp_wnla_prove", "bppp_wnla_verify",
    "bppp_wnla_shard_create", "bppp_wnla_shard_destroy", "bppp_wnla_shard_state", "bppp_wnla_shard_commit_partial", "bppp_wnla_shard_xr_partial",
    "bppp_wnla_shard_fold", "bppp_wnla_shard_export",
    "bppp_peer_create", "bppp_peer_connect", "bppp_peer_destroy", "bppp_peer_world", "bppp_peer_rank", "bppp_peer_msm_allsum", "bppp_peer_allgather",
    "bppp_circuit_commit", "bppp_circuit_prove", "bppp_circuit_verify", "bppp_circuit_commit_sparse", "bppp_circuit_prove_sparse", "bppp_circuit_verify_sparse",
    "bppp_reciprocal_commit_value", "bppp_reciprocal_prove", "bppp_reciprocal_verify",
]

_lib = None


class BpppError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BpppError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        L.bppp_last_error.restype = C.c_char_p
        L.bppp_launch_count.restype = C.c_uint64
        L.bppp_launch_count.argtypes = [C.c_void_p]
        L.bppp_ctx_destroy.argtypes = [C.c_void_p]
        L.bppp_ctx_destroy.restype = None
        L.bppp_multi_ctx_destroy.argtypes = [C.c_void_p]
        L.bppp_multi_ctx_destroy.restype = None
        L.bppp_multi_ctx_get.restype = C.c_void_p
        L.bppp_multi_ctx_get.argtypes = [C.c_void_p, C.c_int]
        L.bppp_wnla_shard_destroy.argtypes = [C.c_void_p]
        L.bppp_wnla_shard_destroy.restype = None
        L.bppp_peer_destroy.argtypes = [C.c_void_p]
        L.bppp_peer_destroy.restype = None
        L.bppp_device_free.restype = None
        L.bppp_device_free.argtypes = [C.c_int, C.c_void_p]
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise BpppError(f"{what} failed with {rc}: {lib().bppp_last_error().decode()}")
