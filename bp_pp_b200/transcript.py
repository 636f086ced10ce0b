"""Host-side `merlin::Transcript` (merlin 3.0.0: STROBE-128 over Keccak-f[1600]) and the reference's two helpers
`transcript::app_point` / `transcript::get_challenge` (src/transcript.rs:6-14).

The batch entry points run this construction on the GPU (csrc/merlin.cuh).  This module is the HOST counterpart the
single-instance API needs: the reference's `prove` / `verify` take a caller-owned `&mut Transcript` in arbitrary prior
state, so the host drives its own transcript and exchanges points / challenges with the engine through the
phase-stepped C ABI (include/bppp.h).  Any object with `append_message(label, msg)` and `challenge_bytes(label, n)`
can stand in for it (in Rust: `merlin::Transcript` itself).
"""
from __future__ import annotations

import struct

_M64 = (1 << 64) - 1
_ROUND_CONSTANTS = []
_ROTATIONS = [[0] * 5 for _ in range(5)]


def _setup():
    # round constants from the degree-8 LFSR, rotation offsets from the (x, y) -> (y, 2x + 3y) walk (FIPS 202, 3.2)
    r = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if r & 1:
                rc |= 1 << ((1 << j) - 1)
            r = ((r << 1) ^ (0x71 if r & 0x80 else 0)) & 0xFF
        _ROUND_CONSTANTS.append(rc)
    x, y = 1, 0
    for t in range(24):
        _ROTATIONS[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5


_setup()


def _rotl(v, n):
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(a):
    """In-place Keccak-f[1600] on 25 lanes, a[x + 5 y]."""
    for rc in _ROUND_CONSTANTS:
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        for x in range(5):
            d = c[(x + 4) % 5] ^ _rotl(c[(x + 1) % 5], 1)
            for y in range(0, 25, 5):
                a[x + y] ^= d
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rotl(a[x + 5 * y], _ROTATIONS[x][y])
        for y in range(0, 25, 5):
            for x in range(5):
                a[x + y] = b[x + y] ^ (~b[(x + 1) % 5 + y] & _M64 & b[(x + 2) % 5 + y])
        a[0] ^= rc


_RATE = 166
_I, _A, _C, _T, _M, _K = 1, 2, 4, 8, 16, 32


class _Strobe128:
    def __init__(self, protocol: bytes):
        self.state = bytearray(200)
        self.state[0:6] = bytes([1, _RATE + 2, 1, 0, 1, 96])
        self.state[6:18] = b"STROBEv1.0.2"
        self._permute()
        self.pos = self.pos_begin = self.cur_flags = 0
        self.meta_ad(protocol, False)

    def _permute(self):
        lanes = list(struct.unpack("<25Q", self.state))
        keccak_f1600(lanes)
        self.state[:] = struct.pack("<25Q", *lanes)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[_RATE + 1] ^= 0x80
        self._permute()
        self.pos = self.pos_begin = 0

    def _absorb(self, data: bytes):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == _RATE:
                self._run_f()

    def _begin(self, flags: int, more: bool):
        if more:
            if flags != self.cur_flags:
                raise ValueError("STROBE: continued operation with different flags")
            return
        old = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old, flags]))
        if flags & (_C | _K) and self.pos:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin(_M | _A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin(_A, more)
        self._absorb(data)

    def prf(self, n: int) -> bytes:
        self._begin(_I | _A | _C, False)
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == _RATE:
                self._run_f()
        return bytes(out)


class Transcript:
    """`merlin::Transcript`: new(label), append_message, append_u64, challenge_bytes."""

    def __init__(self, label: bytes):
        self._strobe = _Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self._strobe.meta_ad(label, False)
        self._strobe.meta_ad(struct.pack("<I", len(message)), True)
        self._strobe.ad(message, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, struct.pack("<Q", x))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self._strobe.meta_ad(label, False)
        self._strobe.meta_ad(struct.pack("<I", n), True)
        return self._strobe.prf(n)


SCALAR_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def app_point(label: bytes, p33: bytes, t) -> None:
    """transcript.rs:6-8: the point's 33-byte `to_bytes()` form (identity = 33 zero bytes) as a message."""
    if len(p33) != 33:
        raise ValueError("app_point takes the 33-byte compressed form")
    t.append_message(label, p33)


def get_challenge(label: bytes, t) -> bytes:
    """transcript.rs:10-14: 32 challenge bytes read big-endian; the reference unwraps `from_repr`, i.e. panics when the
    value is not below the group order (probability ~2^-128)."""
    b = t.challenge_bytes(label, 32)
    if int.from_bytes(b, "big") >= SCALAR_ORDER:
        raise ValueError("get_challenge: challenge is not a canonical scalar (the reference panics here)")
    return b
