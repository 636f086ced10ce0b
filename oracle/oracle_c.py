"""ctypes loader for the C oracle (oracle/oracle.c).  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  `build()` compiles oracle/_build/liboracle.so with gcc when it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def use_native():
    """Rebuild with -march=native on THIS machine and use that build (CPU-baseline timing)."""
    global _lib, _SO
    native = os.path.join(_HERE, "_build", "liboracle_native.so")
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"])
    _SO = native
    _lib = None
    return lib()


def lib():
    global _lib
    if _lib is None:
        if _SO.endswith("liboracle.so"):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_wnla_rounds.restype = C.c_size_t
        _lib.oracle_wnla_rounds.argtypes = [C.c_size_t, C.c_size_t]
    return _lib


def _buf(b: bytes):
    return (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b if len(b) else b"\0")


def _out(n: int):
    return (C.c_uint8 * max(n, 1))()


def u64_commit(gens64: bytes, x: int, s32: bytes) -> bytes:
    out = _out(33)
    st = lib().oracle_u64_commit(_buf(gens64), C.c_uint64(x), _buf(s32), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def u64_prove_batch(gens64: bytes, xs, blinds: bytes, rngs: bytes, label: bytes, threads: int = 1):
    n = len(xs)
    assert len(blinds) == 32 * n and len(rngs) == 3328 * n
    xs_a = (C.c_uint64 * max(n, 1))(*xs)
    out = _out(525 * n)
    status = (C.c_int32 * max(n, 1))()
    st = lib().oracle_u64_prove_batch(_buf(gens64), C.c_size_t(n), xs_a, _buf(blinds), _buf(rngs), _buf(label),
                                      C.c_size_t(len(label)), out, status, C.c_int(threads))
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)[:525 * n], list(status)[:n]


def u64_verify_batch(gens64: bytes, commits: bytes, proofs: bytes, label: bytes, threads: int = 1):
    n = len(commits) // 33
    assert len(commits) == 33 * n and len(proofs) == 525 * n
    status = (C.c_int32 * max(n, 1))()
    st = lib().oracle_u64_verify_batch(_buf(gens64), C.c_size_t(n), _buf(commits), _buf(proofs), _buf(label),
                                       C.c_size_t(len(label)), status, C.c_int(threads))
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return list(status)[:n]


def wnla_commit(g64, gvec64, hvec64, c32, rho32, mu32, l32, n32) -> bytes:
    out = _out(33)
    st = lib().oracle_wnla_commit(_buf(g64), _buf(gvec64), C.c_size_t(len(gvec64) // 64), _buf(hvec64),
                                  C.c_size_t(len(hvec64) // 64), _buf(c32), C.c_size_t(len(c32) // 32), _buf(rho32),
                                  _buf(mu32), _buf(l32), C.c_size_t(len(l32) // 32), _buf(n32),
                                  C.c_size_t(len(n32) // 32), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def wnla_prove(g64, gvec64, hvec64, c32, rho32, mu32, commit33, l32, n32, label):
    """-> (r33s, x33s, l32s, n32s) as byte strings (innermost round first)."""
    ln, nn = len(l32) // 32, len(n32) // 32
    rounds = lib().oracle_wnla_rounds(ln, nn)
    r_out, x_out = _out(33 * rounds), _out(33 * rounds)
    l_out, n_out = _out(32 * max(ln, 1)), _out(32 * max(nn, 1))
    ro, lo, no = C.c_size_t(), C.c_size_t(), C.c_size_t()
    st = lib().oracle_wnla_prove(_buf(g64), _buf(gvec64), C.c_size_t(len(gvec64) // 64), _buf(hvec64),
                                 C.c_size_t(len(hvec64) // 64), _buf(c32), C.c_size_t(len(c32) // 32), _buf(rho32),
                                 _buf(mu32), _buf(commit33), _buf(l32), C.c_size_t(ln), _buf(n32), C.c_size_t(nn),
                                 _buf(label), C.c_size_t(len(label)), r_out, x_out, C.byref(ro), l_out, C.byref(lo),
                                 n_out, C.byref(no))
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return (bytes(r_out)[:33 * ro.value], bytes(x_out)[:33 * ro.value], bytes(l_out)[:32 * lo.value],
            bytes(n_out)[:32 * no.value])


def wnla_verify(g64, gvec64, hvec64, c32, rho32, mu32, commit33, r33, x33, l32, n32, label) -> int:
    return lib().oracle_wnla_verify(_buf(g64), _buf(gvec64), C.c_size_t(len(gvec64) // 64), _buf(hvec64),
                                    C.c_size_t(len(hvec64) // 64), _buf(c32), C.c_size_t(len(c32) // 32), _buf(rho32),
                                    _buf(mu32), _buf(commit33), _buf(r33), C.c_size_t(len(r33) // 33), _buf(x33),
                                    C.c_size_t(len(x33) // 33), _buf(l32), C.c_size_t(len(l32) // 32), _buf(n32),
                                    C.c_size_t(len(n32) // 32), _buf(label), C.c_size_t(len(label)))


def reciprocal_prove(dim_nd, dim_np, g64, gvec64, hvec64, gvec2_64, hvec2_64, x32, s32, digits, rng, label):
    """-> (record, rounds, l_len, n_len, commit33)"""
    cap = 33 * (5 + 2 * 72) + 32 * 16
    out = _out(cap)
    ro, lo, no = C.c_size_t(), C.c_size_t(), C.c_size_t()
    com = _out(33)
    dg = (C.c_uint32 * dim_nd)(*digits)
    st = lib().oracle_reciprocal_prove(C.c_size_t(dim_nd), C.c_size_t(dim_np), _buf(g64), _buf(gvec64),
                                       C.c_size_t(len(gvec64) // 64), _buf(hvec64), C.c_size_t(len(hvec64) // 64),
                                       _buf(gvec2_64), C.c_size_t(len(gvec2_64) // 64), _buf(hvec2_64),
                                       C.c_size_t(len(hvec2_64) // 64), _buf(x32), _buf(s32), dg, _buf(rng),
                                       C.c_size_t(len(rng)), _buf(label), C.c_size_t(len(label)), out, C.c_size_t(cap),
                                       C.byref(ro), C.byref(lo), C.byref(no), com)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    n = 33 * (5 + 2 * ro.value) + 32 * (lo.value + no.value)
    return bytes(out)[:n], ro.value, lo.value, no.value, bytes(com)


def reciprocal_verify(dim_nd, dim_np, g64, gvec64, hvec64, gvec2_64, hvec2_64, commit33, rec, rounds_r, rounds_x, l_len,
                      n_len, label) -> int:
    return lib().oracle_reciprocal_verify(C.c_size_t(dim_nd), C.c_size_t(dim_np), _buf(g64), _buf(gvec64),
                                          C.c_size_t(len(gvec64) // 64), _buf(hvec64), C.c_size_t(len(hvec64) // 64),
                                          _buf(gvec2_64), C.c_size_t(len(gvec2_64) // 64), _buf(hvec2_64),
                                          C.c_size_t(len(hvec2_64) // 64), _buf(commit33), _buf(rec),
                                          C.c_size_t(rounds_r), C.c_size_t(rounds_x), C.c_size_t(l_len),
                                          C.c_size_t(n_len), _buf(label), C.c_size_t(len(label)))


class CircuitDesc(C.Structure):
    _fields_ = [
        ("dim_nm", C.c_size_t), ("dim_no", C.c_size_t), ("k", C.c_size_t), ("dim_nv", C.c_size_t),
        ("f_l", C.c_int), ("f_m", C.c_int),
        ("g64", u8p), ("gvec64", u8p), ("hvec64", u8p), ("gvec2_64", u8p), ("hvec2_64", u8p),
        ("gn", C.c_size_t), ("hn", C.c_size_t), ("gn2", C.c_size_t), ("hn2", C.c_size_t),
        ("W_m32", u8p), ("W_l32", u8p), ("a_m32", u8p), ("a_l32", u8p),
        ("part_lo", C.POINTER(C.c_int32)), ("part_ll", C.POINTER(C.c_int32)),
        ("part_lr", C.POINTER(C.c_int32)), ("part_no", C.POINTER(C.c_int32)), ("part_n", C.c_size_t),
    ]


def make_circuit_desc(dim_nm, dim_no, k, dim_nv, f_l, f_m, g64, gvec64, hvec64, gvec2_64, hvec2_64, W_m32, W_l32, a_m32,
                      a_l32, part_lo, part_ll, part_lr, part_no):
    """Keeps the ctypes buffers alive on the returned object (`._keep`)."""
    d = CircuitDesc()
    keep = []

    def pb(b):
        a = _buf(b)
        keep.append(a)
        return C.cast(a, u8p)

    def pi(v):
        a = (C.c_int32 * max(len(v), 1))(*v)
        keep.append(a)
        return C.cast(a, C.POINTER(C.c_int32))

    d.dim_nm, d.dim_no, d.k, d.dim_nv, d.f_l, d.f_m = dim_nm, dim_no, k, dim_nv, int(f_l), int(f_m)
    d.g64, d.gvec64, d.hvec64, d.gvec2_64, d.hvec2_64 = pb(g64), pb(gvec64), pb(hvec64), pb(gvec2_64), pb(hvec2_64)
    d.gn, d.hn, d.gn2, d.hn2 = len(gvec64) // 64, len(hvec64) // 64, len(gvec2_64) // 64, len(hvec2_64) // 64
    d.W_m32, d.W_l32, d.a_m32, d.a_l32 = pb(W_m32), pb(W_l32), pb(a_m32), pb(a_l32)
    d.part_lo, d.part_ll, d.part_lr, d.part_no = pi(part_lo), pi(part_ll), pi(part_lr), pi(part_no)
    d.part_n = len(part_lo)
    d._keep = keep
    return d


def circuit_commit(desc, v32, s32) -> bytes:
    out = _out(33)
    st = lib().oracle_circuit_commit(C.byref(desc), _buf(v32), _buf(s32), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def circuit_prove(desc, commits33, v32, sv32, wl32, wr32, wo32, rng, label):
    cap = 33 * (4 + 2 * 72) + 32 * 16
    out = _out(cap)
    ro, lo, no = C.c_size_t(), C.c_size_t(), C.c_size_t()
    st = lib().oracle_circuit_prove(C.byref(desc), _buf(commits33), _buf(v32), _buf(sv32), _buf(wl32), _buf(wr32),
                                    _buf(wo32), _buf(rng), C.c_size_t(len(rng)), _buf(label), C.c_size_t(len(label)),
                                    out, C.c_size_t(cap), C.byref(ro), C.byref(lo), C.byref(no))
    if st != 0:
        raise ValueError(f"oracle status {st}")
    n = 33 * (4 + 2 * ro.value) + 32 * (lo.value + no.value)
    return bytes(out)[:n], ro.value, lo.value, no.value


def circuit_verify(desc, commits33, rec, rounds_r, rounds_x, l_len, n_len, label) -> int:
    return lib().oracle_circuit_verify(C.byref(desc), _buf(commits33), _buf(rec), C.c_size_t(rounds_r),
                                       C.c_size_t(rounds_x), C.c_size_t(l_len), C.c_size_t(n_len), _buf(label),
                                       C.c_size_t(len(label)))


def msm(pts64: bytes, sc32: bytes) -> bytes:
    out = _out(33)
    st = lib().oracle_msm(_buf(pts64), _buf(sc32), C.c_size_t(len(sc32) // 32), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def point_mul(p64: bytes, k32: bytes) -> bytes:
    out = _out(64)
    st = lib().oracle_point_mul(_buf(p64), _buf(k32), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def point_add(p64: bytes, q64: bytes) -> bytes:
    out = _out(64)
    st = lib().oracle_point_add(_buf(p64), _buf(q64), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def point_decompress(b33: bytes) -> bytes:
    out = _out(64)
    st = lib().oracle_point_decompress(_buf(b33), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def point_compress(b64: bytes) -> bytes:
    out = _out(33)
    st = lib().oracle_point_compress(_buf(b64), out)
    if st != 0:
        raise ValueError(f"oracle status {st}")
    return bytes(out)


def fe_mul(a, b):
    out = _out(32)
    lib().oracle_fe_mul(_buf(a), _buf(b), out)
    return bytes(out)


def fe_inv(a):
    out = _out(32)
    lib().oracle_fe_inv(_buf(a), out)
    return bytes(out)


def sc_mul(a, b):
    out = _out(32)
    lib().oracle_sc_mul(_buf(a), _buf(b), out)
    return bytes(out)


def sc_inv(a):
    out = _out(32)
    st = lib().oracle_sc_inv(_buf(a), out)
    if st != 0:
        raise ZeroDivisionError(f"oracle status {st}")
    return bytes(out)


def sc_from_wide(b64):
    out = _out(32)
    lib().oracle_sc_from_wide(_buf(b64), out)
    return bytes(out)


def merlin_simple(label, mlabel, msg, clabel, n):
    out = _out(n)
    lib().oracle_merlin_simple(_buf(label), C.c_size_t(len(label)), C.c_char_p(mlabel), _buf(msg),
                               C.c_size_t(len(msg)), C.c_char_p(clabel), out, C.c_size_t(n))
    return bytes(out)


def bench_point_mul(p64, k32, iters):
    out = _out(64)
    lib().oracle_bench_point_mul(_buf(p64), _buf(k32), C.c_int(iters), out)
    return bytes(out)


def bench_sc_inv(a32, iters):
    out = _out(32)
    lib().oracle_bench_sc_inv(_buf(a32), C.c_int(iters), out)
    return bytes(out)
