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


# ---- the tamper set of SURVEY 8(d) config 2 -------------------------------------------------------------------------
TAMPER_EVERY = 16
_PT_OFF = [33 * k for k in range(12)] + [492]       # the 13 points of a 525-byte record


def tamper_batch(proofs: bytes, commits: bytes, add_g, first_index: int = 0, every: int = TAMPER_EVERY):
    """Every `every`-th record (by GLOBAL index first_index + i) mutated by rule (index / every) mod 8, positions drawn from
    S("tamper", index).  Rules (the shape-preserving forms of SURVEY 8d config 2, as in tests/test_gpu_parity.py):
      0 point += G at one of the 13 positions     1 scalar += 1 at one of 3           2 point := identity
      3 commitment += G                           4 swap X and R of one round         5 swap two rounds
      6 one random bit flipped anywhere           7 non-canonical scalar (even) / x >= p (odd)
    add_g(list of 33-byte points) -> list of the same points + G (the caller supplies the group operation: the engine in
    bench.py, the checker in the golden generator).  Returns (proofs, commits, tampered local indices)."""
    n = len(commits) // 33
    recs = bytearray(proofs)
    coms = bytearray(commits)
    idx = [i for i in range(n) if (first_index + i) % every == 0]
    want = []                                          # (local index, where) of points that need + G
    for i in idx:
        gi = first_index + i
        rule = (gi // every) % 8
        r = S("tamper", gi, 8)
        o = 525 * i
        if rule == 0:
            want.append((i, _PT_OFF[r[0] % 13]))
        elif rule == 1:
            p = o + 396 + 32 * (r[0] % 3)
            recs[p:p + 32] = ((int.from_bytes(recs[p:p + 32], "big") + 1) % N).to_bytes(32, "big")
        elif rule == 2:
            p = o + _PT_OFF[r[0] % 13]
            recs[p:p + 33] = bytes(33)
        elif rule == 3:
            want.append((i, -1))
        elif rule == 4:
            j = r[0] % 4
            a, b = o + 132 + 33 * j, o + 264 + 33 * j
            recs[a:a + 33], recs[b:b + 33] = recs[b:b + 33], recs[a:a + 33]
        elif rule == 5:
            a, b = o + 132, o + 132 + 33 * 3
            recs[a:a + 33], recs[b:b + 33] = recs[b:b + 33], recs[a:a + 33]
        elif rule == 6:
            recs[o + int.from_bytes(r[:4], "little") % 525] ^= 1 << (r[4] % 8)
        else:
            if r[0] & 1:
                recs[o + 1:o + 33] = b"\xff" * 32
            else:
                recs[o + 396:o + 428] = b"\xff" * 32
    if want:
        src = [bytes(coms[33 * i:33 * i + 33]) if w < 0 else bytes(recs[525 * i + w:525 * i + w + 33]) for i, w in want]
        for (i, w), q in zip(want, add_g(src)):
            if w < 0:
                coms[33 * i:33 * i + 33] = q
            else:
                recs[525 * i + w:525 * i + w + 33] = q
    return bytes(recs), bytes(coms), idx


def engine_add_g(device: int = 0):
    """add_g for tamper_batch computed by the engine: P + G as the two-term sum 1*P + 1*G on the GPU."""
    from .api import FMT_AFFINE64, FMT_COMPRESSED, msm, points_convert
    one2 = (1).to_bytes(32, "big") * 2

    def add_g(pts):
        aff = points_convert(b"".join(pts), FMT_COMPRESSED, FMT_AFFINE64, device)
        return [msm(aff[64 * k:64 * k + 64] + G64, one2, FMT_AFFINE64, FMT_COMPRESSED, device) for k in range(len(pts))]
    return add_g


def block_hashes(data: bytes, item: int, block: int = 4096):
    """sha256 of every `block`-item slice (hex): how full-size outputs are compared without shipping them."""
    return [hashlib.sha256(data[o:o + item * block]).hexdigest() for o in range(0, len(data), item * block)]
