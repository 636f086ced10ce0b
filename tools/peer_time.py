"""Latency of the exchange step on N ranks (torchrun): the peer-mailbox all-gather against torch.distributed's, and the sharded WNLA
prover with either.  python -m torch.distributed.run --nproc-per-node N tools/peer_time.py [log2n]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import bp_pp_b200 as B
    from bp_pp_b200 import synth
    from bp_pp_b200.shard import PeerGroup, _gather_bytes, wnla_prove_sharded
    from bp_pp_b200.transcript import Transcript
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    peer = PeerGroup(lr)
    mine = bytes([rank]) * 128
    res = {}
    for name, fn in (("peer_allgather_us", lambda: peer.allgather(mine)), ("nccl_allgather_us", lambda: _gather_bytes(mine, lr))):
        for _ in range(5):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        res[name] = round((time.perf_counter() - t0) / 200 * 1e6, 1)
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << log2n
    be = lambda v: (v % synth.N).to_bytes(32, "big")  # noqa: E731
    step64 = B.msm(synth.G64, be(29), B.FMT_AFFINE64, B.FMT_AFFINE64, lr)
    per = n // world
    lo = rank * per
    g64 = B.msm(synth.G64, be(11), B.FMT_AFFINE64, B.FMT_AFFINE64, lr)
    gvec = B.points_generate(B.msm(synth.G64, be(11 + 29 * (1 + lo)), B.FMT_AFFINE64, B.FMT_AFFINE64, lr), step64, per, lr)
    hvec = B.points_generate(B.msm(synth.G64, be(11 + 29 * (1 + n + lo)), B.FMT_AFFINE64, B.FMT_AFFINE64, lr), step64, per, lr)
    rnd = np.random.default_rng(77)

    def scalars():
        a = np.frombuffer(rnd.bytes(32 * n), dtype=np.uint8).reshape(n, 32)[lo:lo + per].copy()
        a[:, 0] &= 0x7F
        return a.tobytes()
    c, l, nn = scalars(), scalars(), scalars()
    rho = 0x1234567890ABCDEF1234567890ABCDEF
    blk = [dict(hvec64=hvec, c32=c, l32=l, gvec64=gvec, n32=nn)]
    for name, pg in (("wnla_nccl", None), ("wnla_peer", peer), ("wnla_nccl2", None), ("wnla_peer2", peer)):
        dist.barrier()
        st = {}
        proof = wnla_prove_sharded(g64, blk, be(rho), be(rho * rho), None, Transcript(b"wnla config 5"), [lr], st, pg)
        res[name] = {"prove_s": round(st["prove_s"], 4), "upload_s": round(st["upload_s"], 4), "device_ms": round(st["device_ms"], 1)}
    if rank == 0:
        print(json.dumps(res))
    peer.close()
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
