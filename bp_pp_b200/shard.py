"""Sharding of a proof batch across the GPUs of one box: one process per GPU, contiguous slices, no
data-path collective (proofs are independent -- the reference has no shared state between calls,
src/range_proof/u64_proof.rs:42-82).  Generators/tables are replicated per GPU by each rank's Context.
`gather_status` is a convenience for callers that want the whole verdict vector on every rank."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`; sizes differ by at most one; concatenation over ranks is 0..n."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bytes(buf: bytes, item: int, n: int, world: int, rank: int) -> bytes:
    lo, hi = shard_bounds(n, world, rank)
    return buf[item * lo:item * hi]


def gather_status(local: Sequence[int], n: int) -> List[int]:
    """All ranks contribute their slice of verdicts; returns the full vector (needs torch.distributed)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    assert len(local) == sizes[rank]
    mx = max(sizes) if sizes else 0
    t = torch.full((mx,), -99, dtype=torch.int32)
    t[:len(local)] = torch.tensor(list(local), dtype=torch.int32)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full: List[int] = []
    for r in range(world):
        full += outs[r][:sizes[r]].tolist()
    return full


class PeerGroup:
    """The ranks' mailboxes for the library's own exchange kernels (bppp_peer, include/bppp.h): remote stores over NVLink /
    NVSwitch into every peer's HBM plus an epoch flag, instead of a collective-library call per exchange.  One per process;
    the 64-byte CUDA IPC handles are all-gathered once here through torch.distributed (set-up, not data path)."""

    def __init__(self, device: int):
        import ctypes as C
        import torch.distributed as dist
        from ._lib import check, lib
        from .api import _in
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world, self.rank = (dist.get_world_size(), dist.get_rank()) if multi else (1, 0)
        self.device = device
        self._h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(lib().bppp_peer_create(C.byref(self._h), C.c_int(device), C.c_int(self.world), C.c_int(self.rank), handle), "bppp_peer_create")
        if multi:
            import torch
            from ._lib import BpppError
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
            err = None
            try:
                check(lib().bppp_peer_connect(self._h, _in(b"".join(handles))), "bppp_peer_connect")
            except BpppError as e:              # e.g. CUDA IPC not permitted between the ranks' containers
                err = e
            # every mailbox is mapped everywhere before the first remote store -- or every rank gives up together
            ok = torch.tensor([0 if err else 1], device=torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu"))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.close()
                raise err or BpppError("bppp_peer_connect failed on another rank of the group")

    def msm_allsum(self, up, out_fmt: int = 0):
        """Sum over all ranks of each rank's resident block MSM (api.UploadedMsm): (encoded point, device ms from the first
        MSM kernel to the reduced sum).  Every rank must call it."""
        import ctypes as C
        from ._lib import check, lib
        from .api import _psz
        out, ms = (C.c_uint8 * _psz(out_fmt))(), C.c_float()
        check(lib().bppp_peer_msm_allsum(self._h, up._p, up._s, C.c_size_t(up.n), C.c_int(out_fmt), out, C.byref(ms)), "bppp_peer_msm_allsum")
        return bytes(out), ms.value

    def allgather(self, mine: bytes) -> List[bytes]:
        """One short byte string per rank (4..240 bytes, a multiple of 4), returned in rank order on every rank."""
        import ctypes as C
        from ._lib import check, lib
        from .api import _in
        out = (C.c_uint8 * (len(mine) * self.world))()
        check(lib().bppp_peer_allgather(self._h, _in(mine), C.c_size_t(len(mine)), out), "bppp_peer_allgather")
        raw = bytes(out)
        return [raw[len(mine) * r:len(mine) * (r + 1)] for r in range(self.world)]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            import ctypes as C
            from ._lib import lib
            lib().bppp_peer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def msm_sharded_resident(up, peer: "PeerGroup", out_fmt: int = 0):
    """`util::vector_mul` over a point vector cut into one resident block per rank (src/util.rs:46-60): block Pippenger, remote
    stores of the partial sums and their reduction fused on each GPU's stream (PeerGroup.msm_allsum) -> (point, device ms)."""
    return peer.msm_allsum(up, out_fmt)


def msm_sharded(points: bytes, scalars32: bytes, device: int, points_fmt: int = 1) -> bytes:
    """One large MSM split by point range over the ranks of the process group (one GPU each): every rank computes the
    partial sum of its contiguous block on its own GPU, the 33-byte partial points are all-gathered (NCCL when the group
    is NCCL: the only data-path collective in this package, a few hundred bytes) and every rank adds them.
    `points`/`scalars32` hold the FULL vectors on every rank (generators are replicated); returns the 33-byte sum."""
    import torch
    import torch.distributed as dist
    from . import api
    world, rank = dist.get_world_size(), dist.get_rank()
    psz = 33 if points_fmt == api.FMT_COMPRESSED else 64
    n = min(len(points) // psz, len(scalars32) // 32)
    lo, hi = shard_bounds(n, world, rank)
    part = api.msm(points[psz * lo:psz * hi], scalars32[32 * lo:32 * hi], points_fmt, api.FMT_COMPRESSED, device)
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
    gathered = [torch.empty(33, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine)
    allparts = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
    return api.points_sum(allparts, api.FMT_COMPRESSED, api.FMT_COMPRESSED, device)


# ---- a standalone WNLA instance over several GPUs (BASELINE config 5) -------------------------------------
class WnlaShard:
    """One contiguous block of a WeightNormLinearArgument instance resident on one GPU (bppp_wnla_shard, include/bppp.h)."""

    def __init__(self, device: int, g64: bytes, hvec64: bytes, c32: bytes, l32: bytes, h_off: int, gvec64: bytes, n32: bytes, g_off: int,
                 rho32: bytes, mu32: bytes, whole: bool = False):
        import ctypes as C
        from ._lib import check, lib
        from .api import _in
        nh, ng = len(hvec64) // 64, len(gvec64) // 64
        if len(c32) != 32 * nh or len(l32) != 32 * nh or len(n32) != 32 * ng:
            raise ValueError("a block holds equally long h_vec / c / l and equally long g_vec / n")
        self._h, self.device = C.c_void_p(), device
        check(lib().bppp_wnla_shard_create(C.byref(self._h), C.c_int(device), _in(g64), _in(hvec64), _in(c32), _in(l32), C.c_size_t(nh), C.c_size_t(h_off),
                                           _in(gvec64), _in(n32), C.c_size_t(ng), C.c_size_t(g_off), _in(rho32), _in(mu32), C.c_int(int(whole))),
              "bppp_wnla_shard_create")

    def state(self):
        import ctypes as C
        from ._lib import check, lib
        nh, ng, ho, go = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        rho, mu = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
        check(lib().bppp_wnla_shard_state(self._h, C.byref(nh), C.byref(ng), C.byref(ho), C.byref(go), rho, mu), "bppp_wnla_shard_state")
        return {"nh": nh.value, "ng": ng.value, "h_off": ho.value, "g_off": go.value, "rho": bytes(rho), "mu": bytes(mu)}

    def commit_partial(self) -> bytes:
        import ctypes as C
        from ._lib import check, lib
        out = (C.c_uint8 * 64)()
        check(lib().bppp_wnla_shard_commit_partial(self._h, out), "bppp_wnla_shard_commit_partial")
        return bytes(out)

    def xr_partial(self):
        import ctypes as C
        from ._lib import check, lib
        out, ms = (C.c_uint8 * 128)(), C.c_float()
        check(lib().bppp_wnla_shard_xr_partial(self._h, out, C.byref(ms)), "bppp_wnla_shard_xr_partial")
        return bytes(out), ms.value

    def fold(self, y32: bytes) -> float:
        import ctypes as C
        from ._lib import check, lib
        from .api import _in
        ms = C.c_float()
        check(lib().bppp_wnla_shard_fold(self._h, _in(y32), C.byref(ms)), "bppp_wnla_shard_fold")
        return ms.value

    def export(self):
        import ctypes as C
        from ._lib import check, lib
        st = self.state()
        nh, ng = st["nh"], st["ng"]
        h, c, l = (C.c_uint8 * max(64 * nh, 1))(), (C.c_uint8 * max(32 * nh, 1))(), (C.c_uint8 * max(32 * nh, 1))()
        g, n = (C.c_uint8 * max(64 * ng, 1))(), (C.c_uint8 * max(32 * ng, 1))()
        check(lib().bppp_wnla_shard_export(self._h, h, c, l, g, n), "bppp_wnla_shard_export")
        return {"hvec64": bytes(h)[:64 * nh], "c32": bytes(c)[:32 * nh], "l32": bytes(l)[:32 * nh], "gvec64": bytes(g)[:64 * ng], "n32": bytes(n)[:32 * ng]}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            from ._lib import lib
            import ctypes as C
            lib().bppp_wnla_shard_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def compress64(p64: bytes) -> bytes:
    """64-byte affine (x || y, all-zero = identity) -> the engine's 33-byte form (SEC1 tag from y's parity; identity = 33 zero
    bytes): a re-encoding of canonical coordinates, no curve arithmetic."""
    if p64 == b"\0" * 64:
        return b"\0" * 33
    return bytes([2 + (p64[63] & 1)]) + p64[:32]


def _gather_bytes(mine: bytes, device: int, peer: "PeerGroup" = None) -> List[bytes]:
    """All-gather of one equal-length byte string per process: through the peer mailboxes when a PeerGroup is given and the
    string fits a slot, else torch.distributed (NCCL on the GPU under an nccl group, gloo on the CPU); a single-process
    call returns [mine]."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [mine]
    if peer is not None and 4 <= len(mine) <= 240 and len(mine) % 4 == 0:
        return peer.allgather(mine)
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def wnla_prove_sharded(g64: bytes, blocks: Sequence[dict], rho32: bytes, mu32: bytes, commitment33: bytes, t, devices: Sequence[int], stats: dict = None,
                       peer: "PeerGroup" = None):
    """`WeightNormLinearArgument::prove(&self, commitment, t, l, n)` (src/wnla.rs:125-190) for an instance cut into equal,
    contiguous blocks: this process holds `blocks` (dicts with hvec64, c32, l32, gvec64, n32), block j resident on
    devices[j]; under torch.distributed the processes' blocks follow each other in rank order.  `t` is the caller's
    transcript (bp_pp_b200.transcript.Transcript or compatible); commitment33 = None computes `commit(l, n)` first (returned
    in stats["commitment33"]).  Every process returns the same (r, x, l, n) byte
    strings, r / x innermost round first (wnla.rs:186-188), byte-identical to the single-GPU bppp_wnla_prove.

    Per round: each block's shares of X and R on its own GPU (host threads within a process), ONE all-gather of
    128 bytes per block (through `peer`'s mailboxes when given: remote stores by the library's own kernels; else
    torch.distributed), identical transcript everywhere, local fold.  Once the blocks are too short to fold locally
    (2^17 -> 1 element after 17 rounds for 2^20 generators on 8 GPUs) their contents are all-gathered and every process
    finishes the last rounds on one GPU."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    import torch.distributed as dist
    from . import api
    from .transcript import app_point, get_challenge
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world, rank = (dist.get_world_size(), dist.get_rank()) if multi else (1, 0)
    nb = len(blocks)
    if nb < 1 or len(devices) < nb:
        raise ValueError("one device per local block")
    nh, ng = len(blocks[0]["hvec64"]) // 64, len(blocks[0]["gvec64"]) // 64
    if any(len(b["hvec64"]) != 64 * nh or len(b["gvec64"]) != 64 * ng for b in blocks):
        raise ValueError("blocks must be equally long")
    total_blocks = world * nb
    whole = total_blocks == 1
    dev0 = devices[0]
    t0 = time.perf_counter()
    shards = [WnlaShard(devices[j], g64, b["hvec64"], b["c32"], b["l32"], (rank * nb + j) * nh, b["gvec64"], b["n32"], (rank * nb + j) * ng, rho32, mu32, whole)
              for j, b in enumerate(blocks)]
    len_l, len_n = total_blocks * nh, total_blocks * ng
    st = {"rounds_sharded": 0, "rounds_whole": 0, "device_ms": 0.0, "upload_s": time.perf_counter() - t0, "exchange_bytes": 0}
    pool = ThreadPoolExecutor(max_workers=nb + 1)        # one worker per local block + one for the running commitment
    pmap = (lambda f, xs: list(pool.map(f, xs))) if nb > 1 else (lambda f, xs: [f(x) for x in xs])
    com_future = None

    def total(parts64):          # sum of 64-byte affine points -> (affine64, compressed33)
        p64 = api.points_sum(b"".join(parts64), api.FMT_AFFINE64, api.FMT_AFFINE64, dev0)
        return p64, compress64(p64)

    if commitment33 is None:
        # WeightNormLinearArgument::commit(l, n) (wnla.rs:66-72) block by block: what a caller computes before proving
        cp = b"".join(pmap(lambda sh: sh.commit_partial(), shards))
        cp = cp if whole else b"".join(_gather_bytes(cp, dev0, peer))
        com64, com33 = total([cp[o:o + 64] for o in range(0, len(cp), 64)])
        st["commitment33"] = com33
    else:
        com64, com33 = api.points_convert(commitment33, api.FMT_COMPRESSED, api.FMT_AFFINE64, dev0), commitment33
    rs, xs = [], []
    first = True
    t1 = time.perf_counter()
    try:
        while len_l + len_n >= 6:                                    # wnla.rs:126
            s0 = shards[0].state()
            if not whole and ((s0["nh"] | s0["ng"]) & 1 or s0["nh"] < 2 or s0["ng"] < 2 or (s0["h_off"] | s0["g_off"]) & 1):
                # the blocks no longer fold locally: gather their contents, finish as one instance on one GPU per process
                ex = [sh.export() for sh in shards]
                keys = ("hvec64", "c32", "l32", "gvec64", "n32")
                mine = {k: b"".join(e[k] for e in ex) for k in keys}
                full = {k: b"".join(_gather_bytes(mine[k], dev0)) for k in keys}
                st["exchange_bytes"] += sum(len(v) for v in full.values())
                for sh in shards:
                    sh.close()
                shards = [WnlaShard(dev0, g64, full["hvec64"], full["c32"], full["l32"], 0, full["gvec64"], full["n32"], 0, s0["rho"], s0["mu"], True)]
                whole, pmap = True, (lambda f, xs: [f(x) for x in xs])
                nb = 1
            parts = pmap(lambda sh: sh.xr_partial(), shards)
            st["device_ms"] += max(ms for _, ms in parts)
            if com_future is not None:                               # C + y X + (y^2 - 1) R of the previous round, computed under this round's MSMs
                com64, com33 = com_future.result()
                com_future = None
            mine = b"".join(p for p, _ in parts)
            allparts = mine if whole else b"".join(_gather_bytes(mine, dev0, peer))
            if not whole:
                st["exchange_bytes"] += len(allparts)
            X64, X33 = total([allparts[o:o + 64] for o in range(0, len(allparts), 128)])
            R64, R33 = total([allparts[o + 64:o + 128] for o in range(0, len(allparts), 128)])
            app_point(b"wnla_com", com33, t); app_point(b"wnla_x", X33, t); app_point(b"wnla_r", R33, t)        # wnla.rs:162-164
            t.append_u64(b"l.sz", len_l); t.append_u64(b"n.sz", len_n)                                        # :165-166
            y32 = get_challenge(b"wnla_challenge", t)
            st["device_ms"] += max(pmap(lambda sh: sh.fold(y32), shards))
            if first:
                # the reference recomputes wnla'.commit(l', n') (wnla.rs:186); it equals C + y X + (y^2 - 1) R only when the
                # caller's commitment matched (l, n), so the first re-commit is evaluated literally, block by block
                cp = b"".join(pmap(lambda sh: sh.commit_partial(), shards))
                cp = cp if whole else b"".join(_gather_bytes(cp, dev0, peer))
                com64, com33 = total([cp[o:o + 64] for o in range(0, len(cp), 64)])
                first = False
            else:
                def next_commitment(c64=com64, x64=X64, r64=R64, yb=y32):                                         # wnla.rs:100-102
                    y = int.from_bytes(yb, "big")
                    sc = (1).to_bytes(32, "big") + yb + ((y * y - 1) % _N).to_bytes(32, "big")
                    n64 = api.msm(c64 + x64 + r64, sc, api.FMT_AFFINE64, api.FMT_AFFINE64, dev0)
                    return n64, compress64(n64)
                com_future = pool.submit(next_commitment)
            rs.append(R33); xs.append(X33)
            len_l, len_n = (len_l + 1) // 2, (len_n + 1) // 2
            st["rounds_whole" if whole else "rounds_sharded"] += 1
        if com_future is not None:
            com_future.result()
        ex = [sh.export() for sh in shards]
        l_out, n_out = b"".join(e["l32"] for e in ex), b"".join(e["n32"] for e in ex)
        if not whole:
            l_out, n_out = b"".join(_gather_bytes(l_out, dev0)), b"".join(_gather_bytes(n_out, dev0))
    finally:
        for sh in shards:
            sh.close()
        pool.shutdown()
    st["prove_s"] = time.perf_counter() - t1
    if stats is not None:
        stats.update(st)
    return b"".join(reversed(rs)), b"".join(reversed(xs)), l_out, n_out
