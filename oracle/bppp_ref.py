"""CPU oracle #1 (pure Python ints) -- TEST INFRASTRUCTURE ONLY, never the product path.

A literal restatement of the reference's Bulletproofs++ algorithm
(distributed-lab/bp-pp 0.1.1) over secp256k1 using Python integers.  Every
function cites the reference file:line it follows (paths relative to
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.

PARITY STATUS: **parity unpinned** against the real k256/merlin crates.  The
reference holds no golden vectors (all of its tests draw from OsRng and only
assert verify(prove(..)) == true, src/tests.rs:41,135,170) and no Rust toolchain
exists in the build container, so the oracle is pinned only against
  * the published Merlin conformance vector (tests/golden/merlin_kat.json),
  * OpenSSL secp256k1 (via `cryptography`) for the group law and encodings,
  * Keccak-f[1600] via hashlib.sha3_256,
and against itself (frozen goldens under tests/golden/, see make_golden.py).
Conventions that are *recalled* from k256 0.13.3 / merlin 3.0.0 rather than
checked are tagged [recalled].
"""
from __future__ import annotations

import hashlib
import struct

# ----------------------------------------------------------------------------
# secp256k1 (k256 0.13.3, Cargo.lock:411-414)
# ----------------------------------------------------------------------------
P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
B = 7

# Points are `None` (identity) or an affine (x, y) tuple of ints.  The
# reference's ProjectivePoint representation is unobservable:
# only equality, to_affine and to_bytes are visible, all canonical.
IDENTITY = None
G = (GX, GY)


def on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return 0 <= x < P and 0 <= y < P and (y * y - x * x * x - B) % P == 0


def _jac_double(X, Y, Z):
    if Z == 0 or Y == 0:
        return (0, 1, 0)
    S = 4 * X * Y * Y % P
    M = 3 * X * X % P
    X3 = (M * M - 2 * S) % P
    Y3 = (M * (S - X3) - 8 * pow(Y, 4, P)) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jac_add_affine(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    R = (S2 - Y1) % P
    if H == 0:
        if R == 0:
            return _jac_double(X1, Y1, Z1)
        return (0, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (R * R - HHH - 2 * V) % P
    Y3 = (R * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def _jac_to_affine(X, Y, Z):
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def pt_add(a, b):
    """ProjectivePoint + ProjectivePoint (exact group law)."""
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def pt_neg(a):
    if a is None:
        return None
    return (a[0], (-a[1]) % P)


def pt_sub(a, b):
    return pt_add(a, pt_neg(b))


def pt_mul(a, k: int):
    """ProjectivePoint * Scalar.  k is taken mod n (a Scalar is always canonical)."""
    k %= N
    if a is None or k == 0:
        return None
    x, y = a
    acc = (0, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(*acc)
        if bit == "1":
            acc = _jac_add_affine(*acc, x, y)
    return _jac_to_affine(*acc)


def pt_eq(a, b) -> bool:
    return a == b


def pt_to_bytes(pt) -> bytes:
    """GroupEncoding::to_bytes -- SEC1 compressed, 33 B; identity = 33 zero bytes [recalled]
    (transcript.rs:7)."""
    if pt is None:
        return b"\x00" * 33
    x, y = pt
    return bytes([2 + (y & 1)]) + x.to_bytes(32, "big")


def pt_from_bytes(b: bytes):
    """SEC1 compressed decode; raises ValueError if x is not on the curve."""
    if len(b) != 33:
        raise ValueError("bad length")
    if b == b"\x00" * 33:
        return None
    if b[0] not in (2, 3):
        raise ValueError("bad tag")
    x = int.from_bytes(b[1:], "big")
    if x >= P:
        raise ValueError("x >= p")
    y2 = (x * x * x + B) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise ValueError("not on curve")
    if (y & 1) != (b[0] & 1):
        y = P - y
    return (x, y)


# ----------------------------------------------------------------------------
# Scalars (k256::Scalar): ints in [0, N)
# ----------------------------------------------------------------------------
def sc(x: int) -> int:
    return x % N


def sc_inv(x: int) -> int:
    """Scalar::invert / invert_vartime .unwrap(): panics (raises) on zero."""
    if x % N == 0:
        raise ZeroDivisionError("Scalar::invert().unwrap() on zero")
    return pow(x, -1, N)


def sc_from_repr(b: bytes) -> int:
    """Scalar::from_repr: 32 B big-endian, None if >= n (transcript.rs:13 unwraps)."""
    v = int.from_bytes(b, "big")
    if v >= N:
        raise ValueError("Scalar::from_repr(..).unwrap() on value >= n")
    return v


def sc_to_bytes(x: int) -> bytes:
    return (x % N).to_bytes(32, "big")


def scalar_generate_biased(rng) -> int:
    """Scalar::generate_biased: one fill_bytes of 64 B, big-endian 512-bit integer
    reduced mod n [recalled] (reciprocal.rs:121; circuit.rs:265-294,371-372)."""
    return int.from_bytes(rng.fill_bytes(64), "big") % N


class ByteRng:
    """RngCore stand-in: hands out a pre-drawn byte string in order (the ABI's RNG contract)."""

    def __init__(self, data: bytes):
        self.data = data
        self.pos = 0

    def fill_bytes(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise ValueError("rng buffer exhausted")
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out


class ShakeRng(ByteRng):
    """Harness stream: SHAKE256(seed)."""

    def __init__(self, seed: bytes, nbytes: int = 52 * 64):
        super().__init__(hashlib.shake_256(seed).digest(nbytes))


# ----------------------------------------------------------------------------
# Keccak-f[1600], STROBE-128, Merlin 3.0.0
# ----------------------------------------------------------------------------
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(lanes):
    """lanes: list of 25 u64, index x + 5*y.  Returns new list."""
    a = list(lanes)
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _ROT[x][y])
        a = [b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & b[(x + 2) % 5 + 5 * y]) for y in range(5) for x in range(5)]
        a[0] ^= _RC[rnd]
    return a


def keccak_f1600_bytes(state: bytearray) -> None:
    lanes = list(struct.unpack("<25Q", bytes(state)))
    lanes = keccak_f1600(lanes)
    state[:] = struct.pack("<25Q", *lanes)


class Strobe128:
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, proto: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600_bytes(st)
        self.st = st
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(proto, False)

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[self.R + 1] ^= 0x80
        keccak_f1600_bytes(self.st)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data: bytes):
        for b in data:
            self.st[self.pos] ^= b
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray()
        for _ in range(n):
            out.append(self.st[self.pos])
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert flags == self.cur_flags
            return
        assert not (flags & self.FLAG_T)
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        force_f = (flags & (self.FLAG_C | self.FLAG_K)) != 0
        if force_f and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin::Transcript (3.0.0)."""

    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", len(message)), True)
        self.strobe.ad(message, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, struct.pack("<Q", x))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", n), True)
        return self.strobe.prf(n, False)


# transcript.rs:6-8
def app_point(label: bytes, p, t: Transcript):
    t.append_message(label, pt_to_bytes(p))


# transcript.rs:10-14
def get_challenge(label: bytes, t: Transcript) -> int:
    return sc_from_repr(t.challenge_bytes(label, 32))


# ----------------------------------------------------------------------------
# util.rs -- generic over T in {Scalar, ProjectivePoint}
# ----------------------------------------------------------------------------
class _ScalarOps:
    zero = 0

    @staticmethod
    def add(a, b):
        return (a + b) % N

    @staticmethod
    def sub(a, b):
        return (a - b) % N

    @staticmethod
    def mul(a, s):
        return a * s % N


class _PointOps:
    zero = None

    add = staticmethod(pt_add)
    sub = staticmethod(pt_sub)
    mul = staticmethod(pt_mul)


def _ops(v):
    for x in v:
        if x is None or isinstance(x, tuple):
            return _PointOps
        return _ScalarOps
    return None


def _ops2(a, b=None, kind=None):
    if kind is not None:
        return kind
    o = _ops(a)
    if o is None and b is not None:
        o = _ops(b)
    return o or _ScalarOps


def reduce(v):
    """util.rs:7-22 -- even/odd split."""
    return list(v[0::2]), list(v[1::2])


def vector_extend(v, n, zero):
    """util.rs:24-26"""
    return [v[i] if i < len(v) else zero for i in range(n)]


def weight_vector_mul(a, b, weight, kind=None):
    """util.rs:28-44 -- sum a_i * (b_i * w^(i+1)); a may be points or scalars, b scalars."""
    o = _ops2(a, kind=kind)
    exp = 1
    result = o.zero
    m = max(len(a), len(b))
    a_ext = vector_extend(a, m, o.zero)
    b_ext = vector_extend(b, m, 0)
    for a_val, b_val in zip(a_ext, b_ext):
        exp = exp * weight % N
        result = o.add(result, o.mul(a_val, b_val * exp % N))
    return result


def vector_mul(a, b, kind=None):
    """util.rs:46-60 -- the naive MSM / inner product."""
    o = _ops2(a, kind=kind)
    result = o.zero
    m = max(len(a), len(b))
    a_ext = vector_extend(a, m, o.zero)
    b_ext = vector_extend(b, m, 0)
    for a_val, b_val in zip(a_ext, b_ext):
        result = o.add(result, o.mul(a_val, b_val))
    return result


def vector_mul_on_scalar(a, s, kind=None):
    """util.rs:62-67"""
    o = _ops2(a, kind=kind)
    return [o.mul(x, s) for x in a]


def vector_add(a, b, kind=None):
    """util.rs:69-76"""
    o = _ops2(a, b, kind=kind)
    m = max(len(a), len(b))
    return [o.add(x, y) for x, y in zip(vector_extend(a, m, o.zero), vector_extend(b, m, o.zero))]


def vector_sub(a, b, kind=None):
    """util.rs:78-85"""
    o = _ops2(a, b, kind=kind)
    m = max(len(a), len(b))
    return [o.sub(x, y) for x, y in zip(vector_extend(a, m, o.zero), vector_extend(b, m, o.zero))]


def e(v, n):
    """util.rs:87-95 -- [1, v, v^2, ...]"""
    out, buf = [], 1
    for _ in range(n):
        out.append(buf)
        buf = buf * v % N
    return out


def pow_(s, n):
    """util.rs:97-99"""
    return pow(s, n, N)


def vector_tensor_mul(a, b):
    """util.rs:111-116"""
    out = []
    for x in b:
        out += vector_mul_on_scalar(a, x)
    return out


def diag_inv(x, n):
    """util.rs:118-132"""
    x_inv = sc_inv(x)
    val = 1
    rows = []
    for i in range(n):
        row = []
        for j in range(n):
            if i == j:
                val = val * x_inv % N
                row.append(val)
            else:
                row.append(0)
        rows.append(row)
    return rows


def vector_mul_on_matrix(a, m):
    """util.rs:134-142 (scalars only on the path)"""
    return [vector_mul(a, [row[j] for row in m], kind=_ScalarOps) for j in range(len(m[0]))]


def minus(v):
    """util.rs:153-155"""
    if v is None or isinstance(v, tuple):
        return pt_mul(v, N - 1)
    return (-v) % N


# ----------------------------------------------------------------------------
# wnla.rs
# ----------------------------------------------------------------------------
class WnlaProof:
    """wnla.rs:25-30"""

    def __init__(self, r, x, l, n):
        self.r, self.x, self.l, self.n = list(r), list(x), list(l), list(n)


class WeightNormLinearArgument:
    """wnla.rs:12-19"""

    def __init__(self, g, g_vec, h_vec, c, rho, mu):
        self.g, self.g_vec, self.h_vec, self.c, self.rho, self.mu = g, list(g_vec), list(h_vec), list(c), rho, mu

    def commit(self, l, n):
        """wnla.rs:66-72"""
        v = (vector_mul(self.c, l, kind=_ScalarOps) + weight_vector_mul(n, n, self.mu, kind=_ScalarOps)) % N
        return pt_add(pt_add(pt_mul(self.g, v), vector_mul(self.h_vec, l, kind=_PointOps)),
                      vector_mul(self.g_vec, n, kind=_PointOps))

    def verify(self, commitment, t: Transcript, proof: WnlaProof) -> bool:
        """wnla.rs:75-121"""
        if len(proof.x) != len(proof.r):
            return False
        if len(proof.x) == 0:
            return pt_eq(commitment, self.commit(proof.l, proof.n))
        c0, c1 = reduce(self.c)
        g0, g1 = reduce(self.g_vec)
        h0, h1 = reduce(self.h_vec)
        app_point(b"wnla_com", commitment, t)
        app_point(b"wnla_x", proof.x[-1], t)
        app_point(b"wnla_r", proof.r[-1], t)
        t.append_u64(b"l.sz", len(self.h_vec))
        t.append_u64(b"n.sz", len(self.g_vec))
        y = get_challenge(b"wnla_challenge", t)
        h_ = vector_add(h0, vector_mul_on_scalar(h1, y, kind=_PointOps), kind=_PointOps)
        g_ = vector_add(vector_mul_on_scalar(g0, self.rho, kind=_PointOps),
                        vector_mul_on_scalar(g1, y, kind=_PointOps), kind=_PointOps)
        c_ = vector_add(c0, vector_mul_on_scalar(c1, y, kind=_ScalarOps), kind=_ScalarOps)
        com_ = pt_add(pt_add(commitment, pt_mul(proof.x[-1], y)), pt_mul(proof.r[-1], (y * y - 1) % N))
        wnla = WeightNormLinearArgument(self.g, g_, h_, c_, self.mu, self.mu * self.mu % N)
        proof_ = WnlaProof(proof.r[:-1], proof.x[:-1], proof.l, proof.n)
        return wnla.verify(com_, t, proof_)

    def prove(self, commitment, t: Transcript, l, n) -> WnlaProof:
        """wnla.rs:125-190"""
        l, n = list(l), list(n)
        if len(l) + len(n) < 6:
            return WnlaProof([], [], l, n)
        rho_inv = sc_inv(self.rho)
        c0, c1 = reduce(self.c)
        l0, l1 = reduce(l)
        n0, n1 = reduce(n)
        g0, g1 = reduce(self.g_vec)
        h0, h1 = reduce(self.h_vec)
        mu2 = self.mu * self.mu % N
        S, Pt = _ScalarOps, _PointOps
        vx = (weight_vector_mul(n0, n1, mu2, kind=S) * (rho_inv * 2 % N)
              + vector_mul(c0, l1, kind=S) + vector_mul(c1, l0, kind=S)) % N
        vr = (weight_vector_mul(n1, n1, mu2, kind=S) + vector_mul(c1, l1, kind=S)) % N
        x = pt_mul(self.g, vx)
        x = pt_add(x, vector_mul(h0, l1, kind=Pt))
        x = pt_add(x, vector_mul(h1, l0, kind=Pt))
        x = pt_add(x, vector_mul(g0, vector_mul_on_scalar(n1, self.rho, kind=S), kind=Pt))
        x = pt_add(x, vector_mul(g1, vector_mul_on_scalar(n0, rho_inv, kind=S), kind=Pt))
        r = pt_mul(self.g, vr)
        r = pt_add(r, vector_mul(h1, l1, kind=Pt))
        r = pt_add(r, vector_mul(g1, n1, kind=Pt))
        app_point(b"wnla_com", commitment, t)
        app_point(b"wnla_x", x, t)
        app_point(b"wnla_r", r, t)
        t.append_u64(b"l.sz", len(l))
        t.append_u64(b"n.sz", len(n))
        y = get_challenge(b"wnla_challenge", t)
        h_ = vector_add(h0, vector_mul_on_scalar(h1, y, kind=Pt), kind=Pt)
        g_ = vector_add(vector_mul_on_scalar(g0, self.rho, kind=Pt), vector_mul_on_scalar(g1, y, kind=Pt), kind=Pt)
        c_ = vector_add(c0, vector_mul_on_scalar(c1, y, kind=S), kind=S)
        l_ = vector_add(l0, vector_mul_on_scalar(l1, y, kind=S), kind=S)
        n_ = vector_add(vector_mul_on_scalar(n0, rho_inv, kind=S), vector_mul_on_scalar(n1, y, kind=S), kind=S)
        wnla = WeightNormLinearArgument(self.g, g_, h_, c_, self.mu, mu2)
        proof = wnla.prove(wnla.commit(l_, n_), t, l_, n_)
        proof.r.append(r)
        proof.x.append(x)
        return proof


# ----------------------------------------------------------------------------
# circuit.rs
# ----------------------------------------------------------------------------
LO, LL, LR, NO = "LO", "LL", "LR", "NO"  # circuit.rs:15-20 PartitionType


class CircuitProof:
    """circuit.rs:24-33"""

    def __init__(self, c_l, c_r, c_o, c_s, r, x, l, n):
        self.c_l, self.c_r, self.c_o, self.c_s = c_l, c_r, c_o, c_s
        self.r, self.x, self.l, self.n = list(r), list(x), list(l), list(n)


class CircuitWitness:
    """circuit.rs:80-91"""

    def __init__(self, v, s_v, w_l, w_r, w_o):
        self.v, self.s_v, self.w_l, self.w_r, self.w_o = v, s_v, w_l, w_r, w_o


class ArithmeticCircuit:
    """circuit.rs:95-139"""

    def __init__(self, dim_nm, dim_no, k, dim_nl, dim_nv, dim_nw, g, g_vec, h_vec, W_m, W_l, a_m, a_l,
                 f_l, f_m, g_vec_, h_vec_, partition):
        self.dim_nm, self.dim_no, self.k = dim_nm, dim_no, k
        self.dim_nl, self.dim_nv, self.dim_nw = dim_nl, dim_nv, dim_nw
        self.g, self.g_vec, self.h_vec = g, list(g_vec), list(h_vec)
        self.W_m, self.W_l, self.a_m, self.a_l = W_m, W_l, a_m, a_l
        self.f_l, self.f_m = f_l, f_m
        self.g_vec_, self.h_vec_ = list(g_vec_), list(h_vec_)
        self.partition = partition

    def commit(self, v, s):
        """circuit.rs:146-151"""
        return pt_add(pt_add(pt_mul(self.g, v[0]), pt_mul(self.h_vec[0], s)),
                      vector_mul(self.h_vec[9:], v[1:], kind=_PointOps))

    # -- helper collectors, circuit.rs:559-653 --
    def linear_comb_coef(self, i, lam, mu):
        coef = 0
        if self.f_l:
            coef = (coef + pow_(lam, self.dim_nv * i)) % N
        if self.f_m:
            coef = (coef + pow_(mu, self.dim_nv * i + 1)) % N
        return coef

    def collect_cl0(self, lam, mu):
        c_l0 = [0] * (self.dim_nv - 1)
        if self.f_l:
            c_l0 = e(lam, self.dim_nv)[1:]
        if self.f_m:
            c_l0 = vector_sub(c_l0, vector_mul_on_scalar(e(mu, self.dim_nv)[1:], mu, kind=_ScalarOps), kind=_ScalarOps)
        return c_l0

    def collect_lambda(self, lam, mu):
        lambda_vec = e(lam, self.dim_nl)
        if self.f_l and self.f_m:
            lambda_vec = vector_sub(
                lambda_vec,
                vector_add(
                    vector_tensor_mul(vector_mul_on_scalar(e(lam, self.dim_nv), mu, kind=_ScalarOps),
                                      e(pow_(mu, self.dim_nv), self.k)),
                    vector_tensor_mul(e(mu, self.dim_nv), e(pow_(lam, self.dim_nv), self.k)),
                    kind=_ScalarOps),
                kind=_ScalarOps)
        return lambda_vec

    def collect_m_rl(self):
        nm = self.dim_nm
        M_lnL = [list(self.W_l[i][:nm]) for i in range(self.dim_nl)]
        M_mnL = [list(self.W_m[i][:nm]) for i in range(self.dim_nm)]
        M_lnR = [list(self.W_l[i][nm:2 * nm]) for i in range(self.dim_nl)]
        M_mnR = [list(self.W_m[i][nm:2 * nm]) for i in range(self.dim_nm)]
        return M_lnL, M_mnL, M_lnR, M_mnR

    def collect_m_o(self):
        nm = self.dim_nm
        W_lO = [list(self.W_l[i][2 * nm:]) for i in range(self.dim_nl)]
        W_mO = [list(self.W_m[i][2 * nm:]) for i in range(self.dim_nm)]

        def map_f(isz, jsz, typ, W_x):
            out = []
            for i in range(isz):
                row = []
                for j in range(jsz):
                    j_ = self.partition(typ, j)
                    row.append(W_x[i][j_] if j_ is not None else 0)
                out.append(row)
            return out

        M_lnO = map_f(self.dim_nl, self.dim_nm, NO, W_lO)
        M_llL = map_f(self.dim_nl, self.dim_nv, LL, W_lO)
        M_llR = map_f(self.dim_nl, self.dim_nv, LR, W_lO)
        M_llO = map_f(self.dim_nl, self.dim_nv, LO, W_lO)
        M_mnO = map_f(self.dim_nm, self.dim_nm, NO, W_mO)
        M_mlL = map_f(self.dim_nm, self.dim_nv, LL, W_mO)
        M_mlR = map_f(self.dim_nm, self.dim_nv, LR, W_mO)
        M_mlO = map_f(self.dim_nm, self.dim_nv, LO, W_mO)
        return M_lnO, M_mnO, M_llL, M_mlL, M_llR, M_mlR, M_llO, M_mlO

    def collect_c(self, lambda_vec, mu_vec, mu):
        S = _ScalarOps
        M_lnL, M_mnL, M_lnR, M_mnR = self.collect_m_rl()
        M_lnO, M_mnO, M_llL, M_mlL, M_llR, M_mlR, M_llO, M_mlO = self.collect_m_o()
        mu_diag_inv = diag_inv(mu, self.dim_nm)
        vm = vector_mul_on_matrix

        def sub(a, b):
            return vector_sub(a, b, kind=S)

        c_nL = vm(sub(vm(lambda_vec, M_lnL), vm(mu_vec, M_mnL)), mu_diag_inv)
        c_nR = vm(sub(vm(lambda_vec, M_lnR), vm(mu_vec, M_mnR)), mu_diag_inv)
        c_nO = vm(sub(vm(lambda_vec, M_lnO), vm(mu_vec, M_mnO)), mu_diag_inv)
        c_lL = sub(vm(lambda_vec, M_llL), vm(mu_vec, M_mlL))
        c_lR = sub(vm(lambda_vec, M_llR), vm(mu_vec, M_mlR))
        c_lO = sub(vm(lambda_vec, M_llO), vm(mu_vec, M_mlO))
        return c_nL, c_nR, c_nO, c_lL, c_lR, c_lO

    def verify(self, v, t: Transcript, proof: CircuitProof) -> bool:
        """circuit.rs:154-256"""
        S, Pt = _ScalarOps, _PointOps
        app_point(b"commitment_cl", proof.c_l, t)
        app_point(b"commitment_cr", proof.c_r, t)
        app_point(b"commitment_co", proof.c_o, t)
        for v_val in v:
            app_point(b"commitment_v", v_val, t)
        rho = get_challenge(b"circuit_rho", t)
        lam = get_challenge(b"circuit_lambda", t)
        beta = get_challenge(b"circuit_beta", t)
        delta = get_challenge(b"circuit_delta", t)
        mu = rho * rho % N
        lambda_vec = self.collect_lambda(lam, mu)
        mu_vec = vector_mul_on_scalar(e(mu, self.dim_nm), mu, kind=S)
        c_nL, c_nR, c_nO, c_lL, c_lR, c_lO = self.collect_c(lambda_vec, mu_vec, mu)
        two = 2
        v_ = None
        for i in range(self.k):
            v_ = pt_add(v_, pt_mul(v[i], self.linear_comb_coef(i, lam, mu)))
        v_ = pt_mul(v_, two)
        app_point(b"commitment_cs", proof.c_s, t)
        tau = get_challenge(b"circuit_tau", t)
        tau_inv = sc_inv(tau)
        tau2 = tau * tau % N
        tau3 = tau2 * tau % N
        delta_inv = sc_inv(delta)
        pn_tau = vector_mul_on_scalar(c_nO, tau3 * delta_inv % N, kind=S)
        pn_tau = vector_sub(pn_tau, vector_mul_on_scalar(c_nL, tau2, kind=S), kind=S)
        pn_tau = vector_add(pn_tau, vector_mul_on_scalar(c_nR, tau, kind=S), kind=S)
        ps_tau = (weight_vector_mul(pn_tau, pn_tau, mu, kind=S)
                  + vector_mul(lambda_vec, self.a_l, kind=S) * tau3 * two
                  - vector_mul(mu_vec, self.a_m, kind=S) * tau3 * two) % N
        pt = pt_add(pt_mul(self.g, ps_tau), vector_mul(self.g_vec, pn_tau, kind=Pt))
        cr_tau = [1, tau_inv * beta % N, tau * beta % N, tau2 * beta % N, tau3 * beta % N,
                  tau * tau3 * beta % N, tau2 * tau3 * beta % N, tau3 * tau3 * beta % N,
                  tau3 * tau3 * tau * beta % N]
        c_l0 = self.collect_cl0(lam, mu)
        cl_tau = vector_mul_on_scalar(c_lO, tau3 * delta_inv % N, kind=S)
        cl_tau = vector_sub(cl_tau, vector_mul_on_scalar(c_lL, tau2, kind=S), kind=S)
        cl_tau = vector_add(cl_tau, vector_mul_on_scalar(c_lR, tau, kind=S), kind=S)
        cl_tau = vector_mul_on_scalar(cl_tau, two, kind=S)
        cl_tau = vector_sub(cl_tau, c_l0, kind=S)
        c = cr_tau + cl_tau
        commitment = pt
        commitment = pt_add(commitment, pt_mul(proof.c_s, tau_inv))
        commitment = pt_sub(commitment, pt_mul(proof.c_o, delta))
        commitment = pt_add(commitment, pt_mul(proof.c_l, tau))
        commitment = pt_sub(commitment, pt_mul(proof.c_r, tau2))
        commitment = pt_add(commitment, pt_mul(v_, tau3))
        while len(c) < len(self.h_vec) + len(self.h_vec_):
            c.append(0)
        wnla = WeightNormLinearArgument(self.g, self.g_vec + self.g_vec_, self.h_vec + self.h_vec_, c, rho, mu)
        return wnla.verify(commitment, t, WnlaProof(proof.r, proof.x, proof.l, proof.n))

    def prove(self, v, witness: CircuitWitness, t: Transcript, rng) -> CircuitProof:
        """circuit.rs:260-556"""
        S, Pt = _ScalarOps, _PointOps
        gb = lambda: scalar_generate_biased(rng)  # noqa: E731
        ro = [gb(), gb(), gb(), gb(), 0, gb(), gb(), gb(), 0]
        rl = [gb(), gb(), gb(), 0, gb(), gb(), gb(), 0, 0]
        rr = [gb(), gb(), 0, gb(), gb(), gb(), 0, 0, 0]
        nl = list(witness.w_l)
        nr = list(witness.w_r)

        def part(typ, size):
            out = []
            for j in range(size):
                i = self.partition(typ, j)
                out.append(witness.w_o[i] if i is not None else 0)
            return out

        no = part(NO, self.dim_nm)
        lo = part(LO, self.dim_nv)
        ll = part(LL, self.dim_nv)
        lr = part(LR, self.dim_nv)
        co = pt_add(vector_mul(self.h_vec, ro + lo, kind=Pt), vector_mul(self.g_vec, no, kind=Pt))
        cl = pt_add(vector_mul(self.h_vec, rl + ll, kind=Pt), vector_mul(self.g_vec, nl, kind=Pt))
        cr = pt_add(vector_mul(self.h_vec, rr + lr, kind=Pt), vector_mul(self.g_vec, nr, kind=Pt))
        app_point(b"commitment_cl", cl, t)
        app_point(b"commitment_cr", cr, t)
        app_point(b"commitment_co", co, t)
        for v_val in v:
            app_point(b"commitment_v", v_val, t)
        rho = get_challenge(b"circuit_rho", t)
        lam = get_challenge(b"circuit_lambda", t)
        beta = get_challenge(b"circuit_beta", t)
        delta = get_challenge(b"circuit_delta", t)
        mu = rho * rho % N
        lambda_vec = self.collect_lambda(lam, mu)
        mu_vec = vector_mul_on_scalar(e(mu, self.dim_nm), mu, kind=S)
        c_nL, c_nR, c_nO, c_lL, c_lR, c_lO = self.collect_c(lambda_vec, mu_vec, mu)
        ls = [gb() for _ in range(self.dim_nv)]
        ns = [gb() for _ in range(self.dim_nm)]
        two = 2
        v_0 = 0
        for i in range(self.k):
            v_0 = (v_0 + witness.v[i][0] * self.linear_comb_coef(i, lam, mu)) % N
        v_0 = v_0 * two % N
        rv = [0] * 9
        for i in range(self.k):
            rv[0] = (rv[0] + witness.s_v[i] * self.linear_comb_coef(i, lam, mu)) % N
        rv[0] = rv[0] * two % N
        v_1 = [0] * (self.dim_nv - 1)
        for i in range(self.k):
            v_1 = vector_add(v_1, vector_mul_on_scalar(witness.v[i][1:], self.linear_comb_coef(i, lam, mu), kind=S),
                             kind=S)
        v_1 = vector_mul_on_scalar(v_1, two, kind=S)
        c_l0 = self.collect_cl0(lam, mu)
        f_ = [0] * 8
        delta2 = delta * delta % N
        delta_inv = sc_inv(delta)
        vm = lambda a, b: vector_mul(a, b, kind=S)  # noqa: E731
        wvm = lambda a, b: weight_vector_mul(a, b, mu, kind=S)  # noqa: E731
        va = lambda a, b: vector_add(a, b, kind=S)  # noqa: E731
        f_[0] = minus(wvm(ns, ns))
        f_[1] = (vm(c_l0, ls) + delta * two * wvm(ns, no)) % N
        f_[2] = (minus(vm(c_lR, ls) * two % N) - vm(c_l0, lo) * delta - wvm(ns, va(nl, c_nR)) * two
                 - wvm(no, no) * delta2) % N
        f_[3] = (vm(c_lL, ls) * two + vm(c_lR, lo) * delta * two + vm(c_l0, ll)
                 + wvm(ns, va(nr, c_nL)) * two + wvm(no, va(nl, c_nR)) * two * delta) % N
        f_[4] = (wvm(c_nR, c_nR) - vm(c_lO, ls) * delta_inv * two - vm(c_lL, lo) * delta * two
                 - vm(c_lR, ll) * two - vm(c_l0, lr) - wvm(ns, c_nO) * delta_inv * two
                 - wvm(no, va(nr, c_nL)) * delta * two - wvm(va(nl, c_nR), va(nl, c_nR))) % N
        f_[5] = (wvm(c_nO, c_nR) * delta_inv * two + wvm(c_nL, c_nL) - vm(c_lO, ll) * delta_inv * two
                 - vm(c_lL, lr) * two - vm(c_lR, v_1) * two - wvm(va(nl, c_nR), c_nO) * delta_inv * two
                 - wvm(va(nr, c_nL), va(nr, c_nL))) % N
        f_[6] = (minus(wvm(c_nO, c_nL) * delta_inv * two % N) + vm(c_nO, lr) * delta_inv * two
                 + vm(c_lL, v_1) * two + wvm(va(nr, c_nL), c_nO) * delta_inv * two) % N
        f_[7] = minus(vm(c_lO, v_1) * delta_inv * two % N)
        beta_inv = sc_inv(beta)
        rs = [
            (f_[1] + ro[1] * delta * beta) % N,
            f_[0] * beta_inv % N,
            ((ro[0] * delta + f_[2]) * beta_inv - rl[1]) % N,
            ((f_[3] - rl[0]) * beta_inv + (ro[2] * delta + rr[1])) % N,
            ((f_[4] + rr[0]) * beta_inv + (ro[3] * delta - rl[2])) % N,
            minus(rv[0] * beta_inv % N),
            (f_[5] * beta_inv + ro[5] * delta + rr[3] - rl[4]) % N,
            (f_[6] * beta_inv + rr[4] + ro[6] * delta - rl[5]) % N,
            (f_[7] * beta_inv + ro[7] * delta - rl[6] + rr[5]) % N,
        ]
        cs = pt_add(vector_mul(self.h_vec, rs + ls, kind=Pt), vector_mul(self.g_vec, ns, kind=Pt))
        app_point(b"commitment_cs", cs, t)
        tau = get_challenge(b"circuit_tau", t)
        tau_inv = sc_inv(tau)
        tau2 = tau * tau % N
        tau3 = tau2 * tau % N
        vms = lambda a, s: vector_mul_on_scalar(a, s, kind=S)  # noqa: E731
        vs = lambda a, b: vector_sub(a, b, kind=S)  # noqa: E731
        l = vms(rs + ls, tau_inv)
        l = vs(l, vms(ro + lo, delta))
        l = va(l, vms(rl + ll, tau))
        l = vs(l, vms(rr + lr, tau2))
        l = va(l, vms(rv + v_1, tau3))
        pn_tau = vms(c_nO, tau3 * delta_inv % N)
        pn_tau = vs(pn_tau, vms(c_nL, tau2))
        pn_tau = va(pn_tau, vms(c_nR, tau))
        ps_tau = (wvm(pn_tau, pn_tau) + vm(lambda_vec, self.a_l) * tau3 * two
                  - vm(mu_vec, self.a_m) * tau3 * two) % N
        n_tau = vms(ns, tau_inv)
        n_tau = vs(n_tau, vms(no, delta))
        n_tau = va(n_tau, vms(nl, tau))
        n_tau = vs(n_tau, vms(nr, tau2))
        n = va(pn_tau, n_tau)
        cr_tau = [1, tau_inv * beta % N, tau * beta % N, tau2 * beta % N, tau3 * beta % N,
                  tau * tau3 * beta % N, tau2 * tau3 * beta % N, tau3 * tau3 * beta % N,
                  tau3 * tau3 * tau * beta % N]
        cl_tau = vms(c_lO, tau3 * delta_inv % N)
        cl_tau = vs(cl_tau, vms(c_lL, tau2))
        cl_tau = va(cl_tau, vms(c_lR, tau))
        cl_tau = vms(cl_tau, two)
        cl_tau = vs(cl_tau, c_l0)
        c = cr_tau + cl_tau
        vv = (ps_tau + tau3 * v_0) % N
        commitment = pt_add(pt_add(pt_mul(self.g, vv), vector_mul(self.h_vec, l, kind=Pt)),
                            vector_mul(self.g_vec, n, kind=Pt))
        while len(l) < len(self.h_vec) + len(self.h_vec_):
            l.append(0)
            c.append(0)
        while len(n) < len(self.g_vec) + len(self.g_vec_):
            n.append(0)
        wnla = WeightNormLinearArgument(self.g, self.g_vec + self.g_vec_, self.h_vec + self.h_vec_, c, rho, mu)
        pw = wnla.prove(commitment, t, l, n)
        return CircuitProof(cl, cr, co, cs, pw.r, pw.x, pw.l, pw.n)


# ----------------------------------------------------------------------------
# range_proof/reciprocal.rs
# ----------------------------------------------------------------------------
class ReciprocalWitness:
    """reciprocal.rs:17-26"""

    def __init__(self, x, s, m, digits):
        self.x, self.s, self.m, self.digits = x, s, list(m), list(digits)


class ReciprocalProof:
    """reciprocal.rs:30-33"""

    def __init__(self, circuit_proof: CircuitProof, r):
        self.circuit_proof, self.r = circuit_proof, r


class ReciprocalRangeProofProtocol:
    """reciprocal.rs:64-84"""

    def __init__(self, dim_nd, dim_np, g, g_vec, h_vec, g_vec_, h_vec_):
        self.dim_nd, self.dim_np, self.g = dim_nd, dim_np, g
        self.g_vec, self.h_vec, self.g_vec_, self.h_vec_ = list(g_vec), list(h_vec), list(g_vec_), list(h_vec_)

    def commit_value(self, x, s):
        """reciprocal.rs:88-90"""
        return pt_add(pt_mul(self.g, x), pt_mul(self.h_vec[0], s))

    def commit_poles(self, r, s):
        """reciprocal.rs:93-95"""
        return pt_add(pt_mul(self.h_vec[0], s), vector_mul(self.h_vec[9:], r, kind=_PointOps))

    def verify(self, commitment, proof: ReciprocalProof, t: Transcript) -> bool:
        """reciprocal.rs:98-107"""
        app_point(b"reciprocal_commitment", commitment, t)
        e_ = get_challenge(b"reciprocal_challenge", t)
        circuit = self.make_circuit(e_)
        circuit_commitment = pt_add(commitment, proof.r)
        return circuit.verify([circuit_commitment], t, proof.circuit_proof)

    def prove(self, commitment, witness: ReciprocalWitness, t: Transcript, rng) -> ReciprocalProof:
        """reciprocal.rs:110-146"""
        app_point(b"reciprocal_commitment", commitment, t)
        e_ = get_challenge(b"reciprocal_challenge", t)
        r = [sc_inv((witness.digits[i] + e_) % N) for i in range(self.dim_nd)]
        r_blind = scalar_generate_biased(rng)
        r_com = self.commit_poles(r, r_blind)
        v = [witness.x] + r
        circuit = self.make_circuit(e_)
        cw = CircuitWitness([v], [(witness.s + r_blind) % N], witness.digits, r, witness.m)
        circuit_commitment = circuit.commit(cw.v[0], cw.s_v[0])
        return ReciprocalProof(circuit.prove([circuit_commitment], cw, t, rng), r_com)

    def make_circuit(self, e_) -> ArithmeticCircuit:
        """reciprocal.rs:150-214"""
        dim_nm = self.dim_nd
        dim_no = self.dim_np
        dim_nv = self.dim_nd + 1
        dim_nl = dim_nv
        dim_nw = self.dim_nd * 2 + self.dim_np
        a_m = [1] * dim_nm
        W_m = [[0] * dim_nw for _ in range(dim_nm)]
        for i in range(dim_nm):
            W_m[i][i + dim_nm] = minus(e_)
        a_l = [0] * dim_nl
        base = self.dim_np % N
        W_l = [[0] * dim_nw for _ in range(dim_nl)]
        for i in range(dim_nm):
            W_l[0][i] = minus(pow_(base, i))
        for i in range(dim_nm):
            for j in range(dim_nm):
                W_l[i + 1][j + dim_nm] = 1
        for i in range(dim_nm):
            W_l[i + 1][i + dim_nm] = 0
        inv = [minus(sc_inv((e_ + j) % N)) for j in range(dim_no)]  # same values the reference recomputes per row
        for i in range(dim_nm):
            for j in range(dim_no):
                W_l[i + 1][j + 2 * dim_nm] = inv[j]
        dim_np = self.dim_np

        def partition(typ, index):
            if typ == LL and index < dim_np:
                return index
            return None

        return ArithmeticCircuit(dim_nm, dim_no, 1, dim_nl, dim_nv, dim_nw, self.g, self.g_vec, self.h_vec,
                                 W_m, W_l, a_m, a_l, True, False, self.g_vec_, self.h_vec_, partition)


# ----------------------------------------------------------------------------
# range_proof/u64_proof.rs
# ----------------------------------------------------------------------------
G_VEC_FULL_SZ = 16      # u64_proof.rs:12
H_VEC_CIRCUIT_SZ = 26   # u64_proof.rs:13
H_VEC_FULL_SZ = 32      # u64_proof.rs:14


class U64RangeProofProtocol:
    """u64_proof.rs:19-28"""
    DIM_ND = 16
    DIM_NP = 16

    def __init__(self, g, g_vec, h_vec):
        self.g, self.g_vec, self.h_vec = g, list(g_vec), list(h_vec)

    def _reciprocal(self):
        return ReciprocalRangeProofProtocol(self.DIM_ND, self.DIM_NP, self.g, self.g_vec,
                                            self.h_vec[:H_VEC_CIRCUIT_SZ], [], self.h_vec[H_VEC_CIRCUIT_SZ:])

    def commit_value(self, x: int, s: int):
        """u64_proof.rs:37-39"""
        return pt_add(pt_mul(self.g, x % N), pt_mul(self.h_vec[0], s))

    def verify(self, v, proof: ReciprocalProof, t: Transcript) -> bool:
        """u64_proof.rs:42-54"""
        return self._reciprocal().verify(v, proof, t)

    def prove(self, x: int, s: int, t: Transcript, rng) -> ReciprocalProof:
        """u64_proof.rs:57-82"""
        digits = self.u64_to_hex(x)
        poles = self.u64_to_hex_mapped(x)
        reciprocal = self._reciprocal()
        witness = ReciprocalWitness(x % N, s, poles, digits)
        return reciprocal.prove(reciprocal.commit_value(witness.x, witness.s), witness, t, rng)

    @staticmethod
    def u64_to_hex(x: int):
        """u64_proof.rs:84-90"""
        out = []
        for _ in range(16):
            out.append(x % 16)
            x //= 16
        return out

    @staticmethod
    def u64_to_hex_mapped(x: int):
        """u64_proof.rs:92-102"""
        result = [0] * 16
        for _ in range(16):
            result[x % 16] += 1
            x //= 16
        return result


# ----------------------------------------------------------------------------
# Wire format (SerializableProof, wnla.rs:33-39 / circuit.rs:36-46 / reciprocal.rs:37-41)
# The canonical parity artefact is the raw record: points SEC1-compressed (33 B),
# scalars 32 B big-endian, in struct field order; r[]/x[] in the reference's push
# order (innermost round first, wnla.rs:186-188).
# ----------------------------------------------------------------------------
def serialize_reciprocal_proof(proof: ReciprocalProof) -> bytes:
    cp = proof.circuit_proof
    out = b"".join(pt_to_bytes(p) for p in (cp.c_l, cp.c_r, cp.c_o, cp.c_s))
    out += b"".join(pt_to_bytes(p) for p in cp.r)
    out += b"".join(pt_to_bytes(p) for p in cp.x)
    out += b"".join(sc_to_bytes(s) for s in cp.l)
    out += b"".join(sc_to_bytes(s) for s in cp.n)
    out += pt_to_bytes(proof.r)
    return out


def deserialize_u64_proof(rec: bytes) -> ReciprocalProof:
    """Inverse of serialize_reciprocal_proof for the canonical u64 shape (525 B)."""
    if len(rec) != 525:
        raise ValueError("u64 proof record must be 525 bytes")
    pts = [pt_from_bytes(rec[33 * i:33 * i + 33]) for i in range(12)]
    off = 12 * 33
    sc_ = [sc_from_repr(rec[off + 32 * i:off + 32 * i + 32]) for i in range(3)]
    r = pt_from_bytes(rec[off + 96:off + 96 + 33])
    return ReciprocalProof(CircuitProof(pts[0], pts[1], pts[2], pts[3], pts[4:8], pts[8:12], sc_[0:2], sc_[2:3]), r)


def serialize_wnla_proof(proof: WnlaProof) -> bytes:
    out = b"".join(pt_to_bytes(p) for p in proof.r)
    out += b"".join(pt_to_bytes(p) for p in proof.x)
    out += b"".join(sc_to_bytes(s) for s in proof.l)
    out += b"".join(sc_to_bytes(s) for s in proof.n)
    return out


def serialize_circuit_proof(cp: CircuitProof) -> bytes:
    out = b"".join(pt_to_bytes(p) for p in (cp.c_l, cp.c_r, cp.c_o, cp.c_s))
    return out + serialize_wnla_proof(WnlaProof(cp.r, cp.x, cp.l, cp.n))


def _hex_pt(p) -> str:
    # k256's `AffinePoint: Serialize` goes through sec1::EncodedPoint, whose identity encoding is the single byte 00
    # (SEC1 2.3.3), not the 33 zero bytes GroupEncoding::to_bytes yields for the transcript; serdect upper-hex [recalled]
    if p is None:
        return "00"
    return pt_to_bytes(p).hex().upper()


def _hex_sc(s) -> str:
    return sc_to_bytes(s).hex().upper()


def reciprocal_proof_to_json_obj(proof: ReciprocalProof) -> dict:
    """serde_json form of reciprocal::SerializableProof [recalled hex conventions]."""
    cp = proof.circuit_proof
    return {
        "circuit_proof": {
            "c_l": _hex_pt(cp.c_l), "c_r": _hex_pt(cp.c_r), "c_o": _hex_pt(cp.c_o), "c_s": _hex_pt(cp.c_s),
            "r": [_hex_pt(p) for p in cp.r], "x": [_hex_pt(p) for p in cp.x],
            "l": [_hex_sc(s) for s in cp.l], "n": [_hex_sc(s) for s in cp.n],
        },
        "r": _hex_pt(proof.r),
    }


# ----------------------------------------------------------------------------
# Seeded synthetic inputs: S(tag, i) = SHAKE256("bppp-bench" || tag || LE64(i))
# ----------------------------------------------------------------------------
def S(tag: str, i: int, nbytes: int) -> bytes:
    return hashlib.shake_256(b"bppp-bench" + tag.encode() + struct.pack("<Q", i)).digest(nbytes)


def synth_generators():
    """g, g_vec[16], h_vec[32] = hash-to-scalar(S("gen", j)) * G."""
    pts = [pt_mul(G, int.from_bytes(S("gen", j, 64), "big") % N) for j in range(1 + 16 + 32)]
    return pts[0], pts[1:17], pts[17:49]


def synth_x(i: int) -> int:
    if i == 0:
        return 0
    if i == 1:
        return 1
    if i == 2:
        return 2**64 - 1
    return int.from_bytes(S("x", i, 8), "little")


def synth_blind(i: int) -> int:
    return int.from_bytes(S("blind", i, 64), "big") % N


def synth_rng_bytes(i: int) -> bytes:
    return S("rng", i, 52 * 64)
