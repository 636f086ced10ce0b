"""Seeded synthetic inputs of SURVEY 8(d), computed by the product itself:
S(tag, i) = SHAKE256("bppp-bench" || tag || LE64(i)).  Generators are hash-to-scalar(S("gen", j)) * G with the scalar
multiplication done by the engine's own MSM on the GPU, so bench.py and the tools need only this package."""
from __future__ import annotations

import hashlib
import struct

# secp256k1 domain parameters (SEC 2): group order and base point
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
G64 = GX.to_bytes(32, "big") + GY.to_bytes(32, "big")


def S(tag: str, i: int, nbytes: int) -> bytes:
    return hashlib.shake_256(b"bppp-bench" + tag.encode() + struct.pack("<Q", i)).digest(nbytes)


def synth_scalar(tag: str, i: int) -> bytes:
    """64 bytes of S(tag, i) reduced mod n, 32-byte big-endian (the wide reduction Scalar::generate_biased uses)."""
    return (int.from_bytes(S(tag, i, 64), "big") % N).to_bytes(32, "big")


def synth_generators64(device: int = 0, count: int = 49) -> bytes:
    """g || g_vec[16] || h_vec[32] as 64-byte affine points: hash-to-scalar(S("gen", j)) * G, on the GPU."""
    from .api import FMT_AFFINE64, msm
    return b"".join(msm(G64, synth_scalar("gen", j), FMT_AFFINE64, FMT_AFFINE64, device) for j in range(count))


def synth_x(i: int) -> int:
    if i < 3:
        return (0, 1, 2**64 - 1)[i]
    return int.from_bytes(S("x", i, 8), "little")


def synth_blind(i: int) -> bytes:
    return synth_scalar("blind", i)


def synth_rng_bytes(i: int) -> bytes:
    return S("rng", i, 52 * 64)


def synth_batch(n: int, start: int = 0):
    """(xs uint64[n], blinds uint8[n,32], rng uint8[n,3328]) of proofs start..start+n of the SURVEY 8(d) batch."""
    import numpy as np
    xs = np.array([synth_x(start + i) for i in range(n)], dtype=np.uint64)
    blinds = np.frombuffer(b"".join(synth_blind(start + i) for i in range(n)), dtype=np.uint8).reshape(n, 32).copy()
    rng = np.frombuffer(b"".join(synth_rng_bytes(start + i) for i in range(n)), dtype=np.uint8).reshape(n, 3328).copy()
    return xs, blinds, rng
