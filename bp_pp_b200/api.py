"""Python mirror of the reference's u64 range-proof interface over the C ABI (include/bppp.h)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence, Tuple

from ._lib import BpppError, check, lib

# src/range_proof/u64_proof.rs:12-14
G_VEC_FULL_SZ = 16
H_VEC_CIRCUIT_SZ = 26
H_VEC_FULL_SZ = 32

FMT_COMPRESSED = 0
FMT_AFFINE64 = 1
U64_PROOF_BYTES = 525
U64_PROOF_BYTES_AFFINE = 928
U64_RNG_BYTES = 52 * 64

ST_FALSE, ST_TRUE = 0, 1
ST_PANIC_INVERT_ZERO, ST_PANIC_CHALLENGE_RANGE, ST_BAD_POINT, ST_BAD_SCALAR = -1, -2, -3, -4


def _in(b):
    """Read-only buffer argument.  `bytes` go through as a pointer to their own storage (no copy: these are 100 MB+ for the
    large generic calls); other buffer types are copied into a ctypes array."""
    if isinstance(b, bytes):
        return b if len(b) else b"\0"
    return (C.c_uint8 * max(len(b), 1)).from_buffer_copy(bytes(b) if len(b) else b"\0")


class Context:
    """Owns the device-side state of one U64RangeProofProtocol on one GPU (tables + workspace)."""

    def __init__(self, gens64: bytes, device: int = 0, window_bits: int = 0, max_batch: int = 65536, _shared_from=None):
        self._h = C.c_void_p()
        self._parent = _shared_from          # keeps the owner of the tables alive
        if _shared_from is not None:
            check(lib().bppp_ctx_create_shared(C.byref(self._h), _shared_from._h, C.c_size_t(max_batch)), "bppp_ctx_create_shared")
            self.device = _shared_from.device
            return
        if len(gens64) != 64 * 49:
            raise ValueError("gens64 must be 49 x 64 bytes: g || g_vec[16] || h_vec[32]")
        check(lib().bppp_ctx_create(C.byref(self._h), C.c_int(device), _in(gens64), C.c_int(window_bits),
                                    C.c_size_t(max_batch)), "bppp_ctx_create")
        self.device = device

    def shared(self, max_batch: int = 0) -> "Context":
        """Another context on the same GPU sharing this one's window tables, with its own workspace and streams: batches
        submitted through different contexts run side by side."""
        return Context(b"", self.device, 0, max_batch, _shared_from=self)

    def set_inflight(self, batches: int):
        """Hint that `batches` independent batches are kept in flight on this GPU (affects lane choices only)."""
        check(lib().bppp_ctx_set_inflight(self._h, C.c_int(batches)), "bppp_ctx_set_inflight")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().bppp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        tb, wb, ms, wbits = C.c_size_t(), C.c_size_t(), C.c_double(), C.c_int()
        check(lib().bppp_ctx_info(self._h, C.byref(tb), C.byref(wb), C.byref(ms), C.byref(wbits)), "bppp_ctx_info")
        return {"table_bytes": tb.value, "workspace_bytes": wb.value, "table_build_ms": ms.value,
                "window_bits": wbits.value}

    def launch_count(self) -> int:
        return int(lib().bppp_launch_count(self._h))

    def profile_begin(self):
        check(lib().bppp_ctx_profile_begin(self._h), "bppp_ctx_profile_begin")

    def profile_end(self) -> dict:
        """{kernel name: (total ms, launches)} for everything launched since profile_begin()."""
        nmax = 64
        names = C.create_string_buffer(48 * nmax)
        ms = (C.c_double * nmax)()
        cnt = (C.c_uint32 * nmax)()
        n = C.c_int()
        check(lib().bppp_ctx_profile_end(self._h, names, ms, cnt, C.c_int(nmax), C.byref(n)), "bppp_ctx_profile_end")
        out = {}
        for k in range(n.value):
            nm = names.raw[48 * k:48 * k + 48].split(b"\0")[0].decode()
            out[nm] = (ms[k], int(cnt[k]))
        return out

    # ---- host-buffer entry points ----
    def commit_batch(self, xs: Sequence[int], blinds32: bytes, fmt: int = FMT_COMPRESSED) -> bytes:
        n = len(xs)
        if len(blinds32) != 32 * n:
            raise ValueError("blinds32 must be n x 32 bytes")
        osz = 33 if fmt == FMT_COMPRESSED else 64
        out = (C.c_uint8 * max(osz * n, 1))()
        xa = (C.c_uint64 * max(n, 1))(*xs)
        check(lib().bppp_u64_commit_batch(self._h, C.c_size_t(n), xa, _in(blinds32), C.c_int(fmt), out),
              "bppp_u64_commit_batch")
        return bytes(out)[:osz * n]

    def verify_batch(self, commits: bytes, proofs: bytes, label: bytes, fmt: int = FMT_COMPRESSED) -> List[int]:
        csz = 33 if fmt == FMT_COMPRESSED else 64
        psz = U64_PROOF_BYTES if fmt == FMT_COMPRESSED else U64_PROOF_BYTES_AFFINE
        n = len(commits) // csz
        if len(commits) != csz * n or len(proofs) != psz * n:
            raise ValueError("commits/proofs length mismatch")
        status = (C.c_int32 * max(n, 1))()
        check(lib().bppp_u64_verify_batch(self._h, C.c_size_t(n), _in(commits), _in(proofs), C.c_int(fmt), _in(label),
                                          C.c_size_t(len(label)), status), "bppp_u64_verify_batch")
        return list(status)[:n]

    def prove_batch(self, xs: Sequence[int], blinds32: bytes, rng: bytes, label: bytes) -> Tuple[bytes, List[int]]:
        n = len(xs)
        if len(blinds32) != 32 * n or len(rng) != U64_RNG_BYTES * n:
            raise ValueError("blinds32 must be n x 32 bytes and rng n x 3328 bytes")
        xa = (C.c_uint64 * max(n, 1))(*xs)
        out = (C.c_uint8 * max(U64_PROOF_BYTES * n, 1))()
        status = (C.c_int32 * max(n, 1))()
        check(lib().bppp_u64_prove_batch(self._h, C.c_size_t(n), xa, _in(blinds32), _in(rng), _in(label),
                                         C.c_size_t(len(label)), out, status), "bppp_u64_prove_batch")
        return bytes(out)[:U64_PROOF_BYTES * n], list(status)[:n]

    # ---- phase-stepped entry points for caller-owned transcripts (include/bppp.h) ----
    def verify_with_transcripts(self, commits: bytes, proofs: bytes, transcripts: Sequence) -> List[int]:
        """U64RangeProofProtocol::verify (u64_proof.rs:42-54 -> reciprocal.rs:98-107 -> circuit.rs:154-256 ->
        wnla.rs:75-121) for n proofs, proof i continuing the caller-owned transcripts[i] (any prior state).  The host does
        every app_point / get_challenge of SURVEY App. B on the caller's objects; the GPU does all group arithmetic.
        The transcripts are left exactly where the reference leaves them."""
        from .transcript import app_point, get_challenge
        n = len(transcripts)
        if len(commits) != 33 * n or len(proofs) != U64_PROOF_BYTES * n:
            raise ValueError("commits / proofs / transcripts length mismatch")
        L = lib()
        rec = lambda i, k: proofs[U64_PROOF_BYTES * i + 33 * k:U64_PROOF_BYTES * i + 33 * k + 33]      # noqa: E731  record point k
        vp = (C.c_uint8 * (33 * n))()
        check(L.bppp_u64_verify_begin(self._h, C.c_size_t(n), _in(commits), _in(proofs), C.c_int(FMT_COMPRESSED), vp), "bppp_u64_verify_begin")
        try:
            vp = bytes(vp)
            chal = bytearray()
            for i, t in enumerate(transcripts):
                app_point(b"reciprocal_commitment", commits[33 * i:33 * i + 33], t)                  # reciprocal.rs:99
                chal += get_challenge(b"reciprocal_challenge", t)                                    # :100
                app_point(b"commitment_cl", rec(i, 0), t); app_point(b"commitment_cr", rec(i, 1), t)  # circuit.rs:155-159
                app_point(b"commitment_co", rec(i, 2), t); app_point(b"commitment_v", vp[33 * i:33 * i + 33], t)
                for lbl in (b"circuit_rho", b"circuit_lambda", b"circuit_beta", b"circuit_delta"):   # :161-164
                    chal += get_challenge(lbl, t)
                app_point(b"commitment_cs", rec(i, 3), t)                                            # :189
                chal += get_challenge(b"circuit_tau", t)                                             # :191
            com = (C.c_uint8 * (33 * n))()
            check(L.bppp_u64_verify_circuit(self._h, _in(bytes(chal)), com), "bppp_u64_verify_circuit")
            for j in range(4):                                                                       # wnla.rs:88-94
                ys = bytearray()
                cb = bytes(com)
                for i, t in enumerate(transcripts):
                    app_point(b"wnla_com", cb[33 * i:33 * i + 33], t)
                    app_point(b"wnla_x", rec(i, 8 + (3 - j)), t)
                    app_point(b"wnla_r", rec(i, 4 + (3 - j)), t)
                    t.append_u64(b"l.sz", 32 >> j); t.append_u64(b"n.sz", 16 >> j)
                    ys += get_challenge(b"wnla_challenge", t)
                check(L.bppp_u64_verify_round(self._h, C.c_int(j), _in(bytes(ys)), com), "bppp_u64_verify_round")
            status = (C.c_int32 * n)()
            check(L.bppp_u64_verify_finish(self._h, status), "bppp_u64_verify_finish")
            return list(status)
        except Exception:
            L.bppp_u64_step_abort(self._h)
            raise

    def prove_with_transcripts(self, xs: Sequence[int], blinds32: bytes, rng: bytes, transcripts: Sequence) -> Tuple[bytes, List[int]]:
        """U64RangeProofProtocol::prove (u64_proof.rs:57-82 -> reciprocal.rs:110-146 -> circuit.rs:260-556 ->
        wnla.rs:125-190) for n witnesses, witness i continuing the caller-owned transcripts[i]."""
        from .transcript import app_point, get_challenge
        n = len(transcripts)
        if len(xs) != n or len(blinds32) != 32 * n or len(rng) != U64_RNG_BYTES * n:
            raise ValueError("xs / blinds32 / rng / transcripts length mismatch")
        L = lib()
        xa = (C.c_uint64 * n)(*xs)
        v = (C.c_uint8 * (33 * n))()
        check(L.bppp_u64_prove_begin(self._h, C.c_size_t(n), xa, _in(blinds32), _in(rng), v), "bppp_u64_prove_begin")
        try:
            v = bytes(v)
            es = bytearray()
            for i, t in enumerate(transcripts):
                app_point(b"reciprocal_commitment", v[33 * i:33 * i + 33], t)                        # reciprocal.rs:114
                es += get_challenge(b"reciprocal_challenge", t)                                      # :115
            p4 = (C.c_uint8 * (132 * n))()
            check(L.bppp_u64_prove_reciprocal(self._h, _in(bytes(es)), p4), "bppp_u64_prove_reciprocal")
            p4 = bytes(p4)
            chal = bytearray()
            for i, t in enumerate(transcripts):                                                      # circuit.rs:347-355
                for k, lbl in enumerate((b"commitment_cl", b"commitment_cr", b"commitment_co", b"commitment_v")):
                    app_point(lbl, p4[132 * i + 33 * k:132 * i + 33 * k + 33], t)
                for lbl in (b"circuit_rho", b"circuit_lambda", b"circuit_beta", b"circuit_delta"):
                    chal += get_challenge(lbl, t)
            cs = (C.c_uint8 * (33 * n))()
            check(L.bppp_u64_prove_circuit(self._h, _in(bytes(chal)), cs), "bppp_u64_prove_circuit")
            cs = bytes(cs)
            taus = bytearray()
            for i, t in enumerate(transcripts):                                                      # circuit.rs:472-474
                app_point(b"commitment_cs", cs[33 * i:33 * i + 33], t)
                taus += get_challenge(b"circuit_tau", t)
            p3 = (C.c_uint8 * (99 * n))()
            check(L.bppp_u64_prove_tau(self._h, _in(bytes(taus)), p3), "bppp_u64_prove_tau")
            for j in range(4):                                                                       # wnla.rs:162-168
                pb = bytes(p3)
                ys = bytearray()
                for i, t in enumerate(transcripts):
                    for k, lbl in enumerate((b"wnla_com", b"wnla_x", b"wnla_r")):
                        app_point(lbl, pb[99 * i + 33 * k:99 * i + 33 * k + 33], t)
                    t.append_u64(b"l.sz", 32 >> j); t.append_u64(b"n.sz", 16 >> j)
                    ys += get_challenge(b"wnla_challenge", t)
                check(L.bppp_u64_prove_round(self._h, C.c_int(j), _in(bytes(ys)), p3), "bppp_u64_prove_round")
            out = (C.c_uint8 * (U64_PROOF_BYTES * n))()
            status = (C.c_int32 * n)()
            check(L.bppp_u64_prove_finish(self._h, out, status), "bppp_u64_prove_finish")
            return bytes(out), list(status)
        except Exception:
            L.bppp_u64_step_abort(self._h)
            raise

    # ---- raw-pointer entry points (host numpy/pinned buffers or device pointers) ----
    def verify_batch_ptr(self, n: int, commits_ptr: int, proofs_ptr: int, label: bytes, status_ptr: int,
                         fmt: int = FMT_COMPRESSED):
        check(lib().bppp_u64_verify_batch(self._h, C.c_size_t(n), C.c_void_p(commits_ptr), C.c_void_p(proofs_ptr),
                                          C.c_int(fmt), _in(label), C.c_size_t(len(label)), C.c_void_p(status_ptr)),
              "bppp_u64_verify_batch")

    def prove_batch_ptr(self, n: int, xs_ptr: int, blinds_ptr: int, rng_ptr: int, label: bytes, proofs_ptr: int,
                        status_ptr: int):
        check(lib().bppp_u64_prove_batch(self._h, C.c_size_t(n), C.c_void_p(xs_ptr), C.c_void_p(blinds_ptr),
                                         C.c_void_p(rng_ptr), _in(label), C.c_size_t(len(label)),
                                         C.c_void_p(proofs_ptr), C.c_void_p(status_ptr)), "bppp_u64_prove_batch")

    def verify_batch_dev(self, n: int, d_commits: int, d_proofs: int, label: bytes, d_status: int,
                         fmt: int = FMT_COMPRESSED, stream: int = 0):
        check(lib().bppp_u64_verify_batch_dev(self._h, C.c_size_t(n), C.c_void_p(d_commits), C.c_void_p(d_proofs),
                                              C.c_int(fmt), _in(label), C.c_size_t(len(label)), C.c_void_p(d_status),
                                              C.c_void_p(stream)), "bppp_u64_verify_batch_dev")

    def prove_batch_dev(self, n: int, d_x: int, d_blinds: int, d_rng: int, label: bytes, d_proofs: int, d_status: int,
                        stream: int = 0):
        check(lib().bppp_u64_prove_batch_dev(self._h, C.c_size_t(n), C.c_void_p(d_x), C.c_void_p(d_blinds),
                                             C.c_void_p(d_rng), _in(label), C.c_size_t(len(label)),
                                             C.c_void_p(d_proofs), C.c_void_p(d_status), C.c_void_p(stream)),
              "bppp_u64_prove_batch_dev")


class MultiContext:
    """One process driving several GPUs (bppp_multi_ctx): the batch is cut into contiguous per-device ranges, one host
    thread per device inside the library, generators and tables replicated, no data-path collective."""

    def __init__(self, gens64: bytes, devices: Sequence[int], window_bits: int = 0, max_batch_per_device: int = 65536):
        if len(gens64) != 64 * 49:
            raise ValueError("gens64 must be 49 x 64 bytes: g || g_vec[16] || h_vec[32]")
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        check(lib().bppp_multi_ctx_create(C.byref(self._h), devs, C.c_int(len(devices)), _in(gens64), C.c_int(window_bits),
                                          C.c_size_t(max_batch_per_device)), "bppp_multi_ctx_create")
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().bppp_multi_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def commit_batch(self, xs: Sequence[int], blinds32: bytes, fmt: int = FMT_COMPRESSED) -> bytes:
        n = len(xs)
        osz = 33 if fmt == FMT_COMPRESSED else 64
        out = (C.c_uint8 * max(osz * n, 1))()
        xa = (C.c_uint64 * max(n, 1))(*xs)
        check(lib().bppp_multi_u64_commit_batch(self._h, C.c_size_t(n), xa, _in(blinds32), C.c_int(fmt), out), "bppp_multi_u64_commit_batch")
        return bytes(out)[:osz * n]

    def verify_batch(self, commits: bytes, proofs: bytes, label: bytes, fmt: int = FMT_COMPRESSED) -> List[int]:
        csz = 33 if fmt == FMT_COMPRESSED else 64
        n = len(commits) // csz
        status = (C.c_int32 * max(n, 1))()
        check(lib().bppp_multi_u64_verify_batch(self._h, C.c_size_t(n), _in(commits), _in(proofs), C.c_int(fmt), _in(label),
                                                C.c_size_t(len(label)), status), "bppp_multi_u64_verify_batch")
        return list(status)[:n]

    def prove_batch(self, xs: Sequence[int], blinds32: bytes, rng: bytes, label: bytes) -> Tuple[bytes, List[int]]:
        n = len(xs)
        xa = (C.c_uint64 * max(n, 1))(*xs)
        out = (C.c_uint8 * max(U64_PROOF_BYTES * n, 1))()
        status = (C.c_int32 * max(n, 1))()
        check(lib().bppp_multi_u64_prove_batch(self._h, C.c_size_t(n), xa, _in(blinds32), _in(rng), _in(label), C.c_size_t(len(label)),
                                               out, status), "bppp_multi_u64_prove_batch")
        return bytes(out)[:U64_PROOF_BYTES * n], list(status)[:n]


class U64RangeProofProtocol:
    """Mirror of `bp_pp::range_proof::u64_proof::U64RangeProofProtocol` (u64_proof.rs:19-102).

    Points are 64-byte affine `x || y` (what a shim holding k256 points passes), scalars 32-byte
    big-endian.  `prove` takes the RNG as the byte string the reference's `RngCore` would have produced
    (52 x 64 bytes).  A panic in the reference surfaces as `BpppError` here.
    """
    DIM_ND = 16
    DIM_NP = 16

    def __init__(self, g: bytes, g_vec: Iterable[bytes], h_vec: Iterable[bytes], device: int = 0, window_bits: int = 0,
                 max_batch: int = 65536):
        g_vec, h_vec = list(g_vec), list(h_vec)
        if len(g_vec) != G_VEC_FULL_SZ or len(h_vec) != H_VEC_FULL_SZ:
            raise ValueError("g_vec must hold 16 points and h_vec 32")   # the reference indexes out of bounds (panic)
        self.g, self.g_vec, self.h_vec = g, g_vec, h_vec
        self.ctx = Context(g + b"".join(g_vec) + b"".join(h_vec), device, window_bits, max_batch)

    # u64_proof.rs:37-39
    def commit_value(self, x: int, s: bytes) -> bytes:
        return self.ctx.commit_batch([x], s)

    def commit_batch(self, xs: Sequence[int], blinds32: bytes, fmt: int = FMT_COMPRESSED) -> bytes:
        return self.ctx.commit_batch(xs, blinds32, fmt)

    # u64_proof.rs:57-82 with the reference's own signature: prove(&self, x, s, t: &mut Transcript, rng)
    def prove_t(self, x: int, s: bytes, t, rng_bytes: bytes) -> bytes:
        """`t` is the caller's transcript (bp_pp_b200.transcript.Transcript or anything with append_message /
        append_u64 / challenge_bytes) in any prior state; it is advanced exactly as the reference advances it."""
        proofs, status = self.ctx.prove_with_transcripts([x], s, rng_bytes, [t])
        if status[0] != ST_TRUE:
            raise BpppError(f"prove: the reference would panic here (status {status[0]})")
        return proofs

    # u64_proof.rs:42-54 with the reference's own signature: verify(&self, v, proof, t: &mut Transcript) -> bool
    def verify_t(self, v: bytes, proof: bytes, t) -> bool:
        status = self.ctx.verify_with_transcripts(v, proof, [t])
        if status[0] < 0:
            raise BpppError(f"verify: malformed input or reference panic (status {status[0]})")
        return status[0] == ST_TRUE

    # u64_proof.rs:57-82 with a fresh Transcript::new(transcript_label) (the batch entry points' convention)
    def prove(self, x: int, s: bytes, transcript_label: bytes, rng_bytes: bytes) -> bytes:
        proofs, status = self.ctx.prove_batch([x], s, rng_bytes, transcript_label)
        if status[0] != ST_TRUE:
            raise BpppError(f"prove: the reference would panic here (status {status[0]})")
        return proofs

    def prove_batch(self, xs: Sequence[int], blinds32: bytes, rng: bytes, transcript_label: bytes):
        return self.ctx.prove_batch(xs, blinds32, rng, transcript_label)

    # u64_proof.rs:42-54
    def verify(self, v: bytes, proof: bytes, transcript_label: bytes) -> bool:
        status = self.ctx.verify_batch(v, proof, transcript_label, FMT_COMPRESSED if len(v) == 33 else FMT_AFFINE64)
        if status[0] < 0:
            raise BpppError(f"verify: malformed input or reference panic (status {status[0]})")
        return status[0] == ST_TRUE

    def verify_batch(self, commits: bytes, proofs: bytes, transcript_label: bytes, fmt: int = FMT_COMPRESSED):
        return self.ctx.verify_batch(commits, proofs, transcript_label, fmt)

    # u64_proof.rs:84-102
    @staticmethod
    def u64_to_hex(x: int) -> List[int]:
        return [(x >> (4 * i)) & 15 for i in range(16)]

    @staticmethod
    def u64_to_hex_mapped(x: int) -> List[int]:
        out = [0] * 16
        for i in range(16):
            out[(x >> (4 * i)) & 15] += 1
        return out


def microbench(device: int = 0) -> dict:
    out = (C.c_double * 14)()
    check(lib().bppp_microbench(C.c_int(device), out, C.c_int(14)), "bppp_microbench")
    keys = ["imad_wide_per_s", "fe_mul_per_s", "fe_sqr_per_s", "sc_mul_per_s", "pt_add_mixed_per_s",
            "pt_double_per_s", "pt_add_per_s", "sm_clock_mhz", "imad_per_s", "iadd_per_s",
            "fe_inv_per_s", "fe_inv_fermat_per_s", "sc_inv_per_s", "sc_inv_fermat_per_s"]
    return dict(zip(keys, list(out)))


# ---- generic entry points (arbitrary sizes) ----
def _psz(fmt: int) -> int:
    return 33 if fmt == FMT_COMPRESSED else 64


def msm(points: bytes, scalars32: bytes, points_fmt: int = FMT_AFFINE64, out_fmt: int = FMT_COMPRESSED, device: int = 0) -> bytes:
    """`util::vector_mul` over points (src/util.rs:46-60): sum scalars[i] * points[i], zero-extending the shorter side."""
    npts, nsc = len(points) // _psz(points_fmt), len(scalars32) // 32
    out = (C.c_uint8 * _psz(out_fmt))()
    check(lib().bppp_msm(C.c_int(device), _in(points), C.c_int(points_fmt), C.c_size_t(npts), _in(scalars32), C.c_size_t(nsc),
                         C.c_int(out_fmt), out), "bppp_msm")
    return bytes(out)


def points_sum(points: bytes, points_fmt: int = FMT_COMPRESSED, out_fmt: int = FMT_COMPRESSED, device: int = 0) -> bytes:
    n = len(points) // _psz(points_fmt)
    out = (C.c_uint8 * _psz(out_fmt))()
    check(lib().bppp_points_sum(C.c_int(device), _in(points), C.c_int(points_fmt), C.c_size_t(n), C.c_int(out_fmt), out), "bppp_points_sum")
    return bytes(out)


class UploadedMsm:
    """Points and scalars decoded once and kept in HBM; `run()` times the MSM alone on the device."""

    def __init__(self, points: bytes, scalars32: bytes, points_fmt: int = FMT_AFFINE64, device: int = 0):
        self.device, self.n = device, min(len(points) // _psz(points_fmt), len(scalars32) // 32)
        self._p, self._s = C.c_void_p(), C.c_void_p()
        check(lib().bppp_points_upload(C.c_int(device), _in(points), C.c_int(points_fmt), C.c_size_t(self.n), C.byref(self._p)), "bppp_points_upload")
        check(lib().bppp_scalars_upload(C.c_int(device), _in(scalars32), C.c_size_t(self.n), C.byref(self._s)), "bppp_scalars_upload")

    def run(self, out_fmt: int = FMT_COMPRESSED):
        out = (C.c_uint8 * _psz(out_fmt))()
        ms = C.c_float()
        check(lib().bppp_msm_uploaded(C.c_int(self.device), self._p, self._s, C.c_size_t(self.n), C.c_int(out_fmt), out, C.byref(ms)), "bppp_msm_uploaded")
        return bytes(out), ms.value

    def close(self):
        for h in (self._p, self._s):
            if h.value:
                lib().bppp_device_free(C.c_int(self.device), h)
        self._p, self._s = C.c_void_p(), C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WeightNormLinearArgument:
    """Mirror of `bp_pp::wnla::WeightNormLinearArgument` (src/wnla.rs:12-19) for arbitrary lengths.
    Points are 64-byte affine, scalars 32-byte big-endian; `prove`/`verify` use a fresh Transcript::new(label)."""

    def __init__(self, g: bytes, g_vec: bytes, h_vec: bytes, c: bytes, rho: bytes, mu: bytes, device: int = 0):
        self.g, self.g_vec, self.h_vec, self.c, self.rho, self.mu, self.device = g, g_vec, h_vec, c, rho, mu, device

    def _pub(self):
        return (C.c_int(self.device), _in(self.g), _in(self.g_vec), C.c_size_t(len(self.g_vec) // 64), _in(self.h_vec),
                C.c_size_t(len(self.h_vec) // 64), _in(self.c), C.c_size_t(len(self.c) // 32), _in(self.rho), _in(self.mu))

    # src/wnla.rs:66-72
    def commit(self, l: bytes, n: bytes) -> bytes:
        out = (C.c_uint8 * 33)()
        check(lib().bppp_wnla_commit(*self._pub(), _in(l), C.c_size_t(len(l) // 32), _in(n), C.c_size_t(len(n) // 32), out), "bppp_wnla_commit")
        return bytes(out)

    # src/wnla.rs:125-190 -> (r, x, l, n) byte strings, r/x innermost round first
    def prove(self, commitment: bytes, label: bytes, l: bytes, n: bytes):
        ln, nn = len(l) // 32, len(n) // 32
        r_out, x_out = (C.c_uint8 * (33 * 64))(), (C.c_uint8 * (33 * 64))()
        l_out, n_out = (C.c_uint8 * max(32 * ln, 1))(), (C.c_uint8 * max(32 * nn, 1))()
        rounds, lo, no, st = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int32()
        check(lib().bppp_wnla_prove(*self._pub(), _in(commitment), _in(l), C.c_size_t(ln), _in(n), C.c_size_t(nn), _in(label), C.c_size_t(len(label)),
                                    r_out, x_out, C.byref(rounds), l_out, C.byref(lo), n_out, C.byref(no), C.byref(st)), "bppp_wnla_prove")
        if st.value != ST_TRUE:
            raise BpppError(f"wnla prove: the reference would panic here (status {st.value})")
        return bytes(r_out)[:33 * rounds.value], bytes(x_out)[:33 * rounds.value], bytes(l_out)[:32 * lo.value], bytes(n_out)[:32 * no.value]

    # src/wnla.rs:75-121 -> 1 / 0 / negative status
    def verify(self, commitment: bytes, label: bytes, r: bytes, x: bytes, l: bytes, n: bytes) -> int:
        verdict = C.c_int32()
        check(lib().bppp_wnla_verify(*self._pub(), _in(commitment), _in(r), C.c_size_t(len(r) // 33), _in(x), C.c_size_t(len(x) // 33), _in(l),
                                     C.c_size_t(len(l) // 32), _in(n), C.c_size_t(len(n) // 32), _in(label), C.c_size_t(len(label)), C.byref(verdict)),
              "bppp_wnla_verify")
        return verdict.value


class ReciprocalRangeProofProtocol:
    """Mirror of `bp_pp::range_proof::reciprocal::ReciprocalRangeProofProtocol` (reciprocal.rs:64-84), any (dim_nd, dim_np)."""

    def __init__(self, dim_nd: int, dim_np: int, g: bytes, g_vec: bytes, h_vec: bytes, g_vec_: bytes, h_vec_: bytes, device: int = 0):
        self.dim_nd, self.dim_np, self.g, self.g_vec, self.h_vec, self.g_vec_, self.h_vec_, self.device = dim_nd, dim_np, g, g_vec, h_vec, g_vec_, h_vec_, device

    def _pub(self):
        return (C.c_int(self.device), C.c_size_t(self.dim_nd), C.c_size_t(self.dim_np), _in(self.g), _in(self.g_vec), C.c_size_t(len(self.g_vec) // 64),
                _in(self.h_vec), C.c_size_t(len(self.h_vec) // 64), _in(self.g_vec_), C.c_size_t(len(self.g_vec_) // 64), _in(self.h_vec_),
                C.c_size_t(len(self.h_vec_) // 64))

    # reciprocal.rs:88-90
    def commit_value(self, x32: bytes, s32: bytes) -> bytes:
        out = (C.c_uint8 * 33)()
        check(lib().bppp_reciprocal_commit_value(C.c_int(self.device), _in(self.g), _in(self.h_vec[:64]), _in(x32), _in(s32), out), "bppp_reciprocal_commit_value")
        return bytes(out)

    # reciprocal.rs:93-95: s * h_vec[0] + <h_vec[9..], r>
    def commit_poles(self, r32: bytes, s32: bytes) -> bytes:
        nr = len(r32) // 32
        return msm(self.h_vec[:64] + self.h_vec[64 * 9:64 * (9 + nr)], s32 + r32, FMT_AFFINE64, FMT_COMPRESSED, self.device)

    # reciprocal.rs:110-146 -> (record, rounds, l_len, n_len, commitment33)
    def prove(self, x32: bytes, s32: bytes, digits, rng: bytes, label: bytes):
        cap = 33 * (5 + 2 * 64) + 32 * 16
        out = (C.c_uint8 * cap)()
        ro, lo, no, st = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int32()
        com = (C.c_uint8 * 33)()
        dg = (C.c_uint32 * max(self.dim_nd, 1))(*digits)
        check(lib().bppp_reciprocal_prove(*self._pub(), _in(x32), _in(s32), dg, _in(rng), C.c_size_t(len(rng)), _in(label), C.c_size_t(len(label)), out,
                                          C.c_size_t(cap), C.byref(ro), C.byref(lo), C.byref(no), com, C.byref(st)), "bppp_reciprocal_prove")
        if st.value != ST_TRUE:
            raise BpppError(f"reciprocal prove: the reference would panic here (status {st.value})")
        n = 33 * (5 + 2 * ro.value) + 32 * (lo.value + no.value)
        return bytes(out)[:n], ro.value, lo.value, no.value, bytes(com)

    # reciprocal.rs:98-107 -> 1 / 0 / negative status
    def verify(self, commitment33: bytes, rec: bytes, rounds_r: int, rounds_x: int, l_len: int, n_len: int, label: bytes) -> int:
        verdict = C.c_int32()
        check(lib().bppp_reciprocal_verify(*self._pub(), _in(commitment33), _in(rec), C.c_size_t(rounds_r), C.c_size_t(rounds_x), C.c_size_t(l_len),
                                           C.c_size_t(n_len), _in(label), C.c_size_t(len(label)), C.byref(verdict)), "bppp_reciprocal_verify")
        return verdict.value


class CircuitDesc(C.Structure):
    """`bppp_circuit_desc` of include/bppp.h."""
    _u8p = C.POINTER(C.c_uint8)
    _fields_ = [
        ("dim_nm", C.c_size_t), ("dim_no", C.c_size_t), ("k", C.c_size_t), ("dim_nv", C.c_size_t), ("f_l", C.c_int), ("f_m", C.c_int),
        ("g64", _u8p), ("gvec64", _u8p), ("hvec64", _u8p), ("gvec2_64", _u8p), ("hvec2_64", _u8p),
        ("gn", C.c_size_t), ("hn", C.c_size_t), ("gn2", C.c_size_t), ("hn2", C.c_size_t),
        ("W_m32", _u8p), ("W_l32", _u8p), ("a_m32", _u8p), ("a_l32", _u8p),
        ("part_lo", C.POINTER(C.c_int32)), ("part_ll", C.POINTER(C.c_int32)), ("part_lr", C.POINTER(C.c_int32)), ("part_no", C.POINTER(C.c_int32)),
        ("part_n", C.c_size_t),
    ]


class SparseMatrix(C.Structure):
    """`bppp_sparse_matrix` of include/bppp.h (CSR, values direct or through a dictionary)."""
    _fields_ = [("rows", C.c_size_t), ("cols", C.c_size_t), ("nnz", C.c_size_t), ("row_ptr", C.POINTER(C.c_uint64)), ("col_idx", C.POINTER(C.c_uint32)),
                ("value_idx", C.POINTER(C.c_uint32)), ("values32", C.POINTER(C.c_uint8)), ("n_values", C.c_size_t)]


class CircuitDescSparse(C.Structure):
    """`bppp_circuit_desc_sparse` of include/bppp.h."""
    _u8p = C.POINTER(C.c_uint8)
    _fields_ = [
        ("dim_nm", C.c_size_t), ("dim_no", C.c_size_t), ("k", C.c_size_t), ("dim_nv", C.c_size_t), ("f_l", C.c_int), ("f_m", C.c_int),
        ("g64", _u8p), ("gvec64", _u8p), ("hvec64", _u8p), ("gvec2_64", _u8p), ("hvec2_64", _u8p),
        ("gn", C.c_size_t), ("hn", C.c_size_t), ("gn2", C.c_size_t), ("hn2", C.c_size_t),
        ("W_m", SparseMatrix), ("W_l", SparseMatrix), ("a_m32", _u8p), ("a_l32", _u8p),
        ("part_lo", C.POINTER(C.c_int32)), ("part_ll", C.POINTER(C.c_int32)), ("part_lr", C.POINTER(C.c_int32)), ("part_no", C.POINTER(C.c_int32)),
        ("part_n", C.c_size_t),
    ]


def dense_to_csr(W32: bytes, rows: int, cols: int, dictionary: bool = True):
    """Row-major dense matrix of 32-byte scalars -> (row_ptr, col_idx, value_idx or None, values32): the CSR form of
    bppp_sparse_matrix, with the distinct values collected into a dictionary when `dictionary`."""
    zero = bytes(32)
    row_ptr, col_idx, val_idx, values, table = [0], [], [], [], {}
    for i in range(rows):
        for j in range(cols):
            e = W32[32 * (i * cols + j):32 * (i * cols + j) + 32]
            if e == zero:
                continue
            col_idx.append(j)
            if dictionary:
                val_idx.append(table.setdefault(e, len(table)))
            else:
                values.append(e)
        row_ptr.append(len(col_idx))
    if dictionary:
        values = sorted(table, key=table.get)
    return row_ptr, col_idx, (val_idx if dictionary else None), b"".join(values)


class ArithmeticCircuit:
    """Mirror of `bp_pp::circuit::ArithmeticCircuit` (circuit.rs:95-139): row-major W_m / W_l (32-byte scalars) and the
    partition function tabulated as four index lists (-1 = None).  sparse = None passes the dense descriptor; "csr" /
    "csr-dict" pass the same matrices through bppp_circuit_desc_sparse (values per non-zero / through a dictionary)."""

    def __init__(self, dim_nm, dim_no, k, dim_nv, g, g_vec, h_vec, W_m, W_l, a_m, a_l, f_l, f_m, g_vec_, h_vec_, part_lo, part_ll, part_lr, part_no, device=0,
                 sparse=None):
        self.device = device
        self._keep = []
        self._sfx = "_sparse" if sparse else ""
        d = CircuitDescSparse() if sparse else CircuitDesc()

        def pb(b):
            a = _in(b); self._keep.append(a); return C.cast(a, CircuitDesc._u8p)

        def pi(v):
            a = (C.c_int32 * max(len(v), 1))(*v); self._keep.append(a); return C.cast(a, C.POINTER(C.c_int32))

        d.dim_nm, d.dim_no, d.k, d.dim_nv, d.f_l, d.f_m = dim_nm, dim_no, k, dim_nv, int(f_l), int(f_m)
        d.g64, d.gvec64, d.hvec64, d.gvec2_64, d.hvec2_64 = pb(g), pb(g_vec), pb(h_vec), pb(g_vec_), pb(h_vec_)
        d.gn, d.hn, d.gn2, d.hn2 = len(g_vec) // 64, len(h_vec) // 64, len(g_vec_) // 64, len(h_vec_) // 64
        if sparse:
            dim_nw = 2 * dim_nm + dim_no
            for name, W, rows in (("W_m", W_m, dim_nm), ("W_l", W_l, dim_nv * k)):
                rp, ci, vi, vals = dense_to_csr(W, rows, dim_nw, dictionary=(sparse == "csr-dict"))
                m = SparseMatrix()
                m.rows, m.cols, m.nnz = rows, dim_nw, len(ci)
                a_rp, a_ci = (C.c_uint64 * len(rp))(*rp), (C.c_uint32 * max(len(ci), 1))(*ci)
                self._keep += [a_rp, a_ci]
                m.row_ptr, m.col_idx = C.cast(a_rp, C.POINTER(C.c_uint64)), C.cast(a_ci, C.POINTER(C.c_uint32))
                if vi is not None:
                    a_vi = (C.c_uint32 * max(len(vi), 1))(*vi)
                    self._keep.append(a_vi)
                    m.value_idx = C.cast(a_vi, C.POINTER(C.c_uint32))
                m.values32, m.n_values = pb(vals), len(vals) // 32
                setattr(d, name, m)
            d.a_m32, d.a_l32 = pb(a_m), pb(a_l)
        else:
            d.W_m32, d.W_l32, d.a_m32, d.a_l32 = pb(W_m), pb(W_l), pb(a_m), pb(a_l)
        d.part_lo, d.part_ll, d.part_lr, d.part_no = pi(part_lo), pi(part_ll), pi(part_lr), pi(part_no)
        d.part_n = len(part_lo)
        self.desc = d

    # circuit.rs:146-151
    def commit(self, v32: bytes, s32: bytes) -> bytes:
        out = (C.c_uint8 * 33)()
        check(getattr(lib(), "bppp_circuit_commit" + self._sfx)(C.c_int(self.device), C.byref(self.desc), _in(v32), _in(s32), out), "bppp_circuit_commit")
        return bytes(out)

    # circuit.rs:260-556 -> (record, rounds, l_len, n_len)
    def prove(self, commits33: bytes, v32: bytes, sv32: bytes, wl32: bytes, wr32: bytes, wo32: bytes, rng: bytes, label: bytes):
        cap = 33 * (4 + 2 * 64) + 32 * 16
        out = (C.c_uint8 * cap)()
        ro, lo, no, st = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int32()
        check(getattr(lib(), "bppp_circuit_prove" + self._sfx)(C.c_int(self.device), C.byref(self.desc), _in(commits33), _in(v32), _in(sv32), _in(wl32), _in(wr32), _in(wo32), _in(rng),
                                       C.c_size_t(len(rng)), _in(label), C.c_size_t(len(label)), out, C.c_size_t(cap), C.byref(ro), C.byref(lo), C.byref(no),
                                       C.byref(st)), "bppp_circuit_prove")
        if st.value != ST_TRUE:
            raise BpppError(f"circuit prove: the reference would panic here (status {st.value})")
        n = 33 * (4 + 2 * ro.value) + 32 * (lo.value + no.value)
        return bytes(out)[:n], ro.value, lo.value, no.value

    # circuit.rs:154-256
    def verify(self, commits33: bytes, rec: bytes, rounds_r: int, rounds_x: int, l_len: int, n_len: int, label: bytes) -> int:
        verdict = C.c_int32()
        check(getattr(lib(), "bppp_circuit_verify" + self._sfx)(C.c_int(self.device), C.byref(self.desc), _in(commits33), _in(rec), C.c_size_t(rounds_r), C.c_size_t(rounds_x),
                                        C.c_size_t(l_len), C.c_size_t(n_len), _in(label), C.c_size_t(len(label)), C.byref(verdict)), "bppp_circuit_verify")
        return verdict.value


def points_generate(base64: bytes, step64: bytes, n: int, device: int = 0) -> bytes:
    """n synthetic generators base + i*step as 64-byte affine points (computed on the GPU)."""
    out = (C.c_uint8 * max(64 * n, 1))()
    check(lib().bppp_points_generate(C.c_int(device), _in(base64), _in(step64), C.c_size_t(n), out), "bppp_points_generate")
    return bytes(out)[:64 * n]


def points_convert(points: bytes, in_fmt: int, out_fmt: int, device: int = 0) -> bytes:
    """SEC1 compressed <-> 64-byte affine for an array of points (SerializableProof <-> Proof conversions)."""
    n = len(points) // _psz(in_fmt)
    out = (C.c_uint8 * max(_psz(out_fmt) * n, 1))()
    check(lib().bppp_points_convert(C.c_int(device), _in(points), C.c_int(in_fmt), C.c_size_t(n), C.c_int(out_fmt), out), "bppp_points_convert")
    return bytes(out)[:_psz(out_fmt) * n]


def u64_proofs_to_affine(proofs: bytes, device: int = 0) -> bytes:
    """525-byte compressed u64 records -> 928-byte records with 64-byte points (BPPP_FMT_AFFINE64)."""
    n = len(proofs) // U64_PROOF_BYTES
    import numpy as np
    rec = np.frombuffer(proofs, dtype=np.uint8).reshape(n, U64_PROOF_BYTES)
    pts = np.concatenate([rec[:, :396].reshape(n, 12, 33), rec[:, 492:525].reshape(n, 1, 33)], axis=1)
    aff = np.frombuffer(points_convert(pts.tobytes(), FMT_COMPRESSED, FMT_AFFINE64, device), dtype=np.uint8).reshape(n, 13, 64)
    out = np.concatenate([aff[:, :12].reshape(n, 768), rec[:, 396:492], aff[:, 12].reshape(n, 64)], axis=1)
    return out.tobytes()
