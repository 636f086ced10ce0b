"""CPU suite: batch sharding across ranks (world_size 2, gloo).  The verify inside each rank is the oracle here
(no GPU in CI); on the GPU box the same slices go through Context.verify_batch."""
import os
import socket
import sys

import pytest

from conftest import ROOT, synth_batch

LABEL = b"u64 range proof"


def test_shard_bounds_partition():
    from bp_pp_b200.shard import shard_bounds
    for n in [0, 1, 2, 7, 8, 65536, 65537]:
        for world in [1, 2, 3, 4, 8]:
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n, gens64, commits, proofs, expect, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle_c
    from bp_pp_b200.shard import gather_status, shard_bytes
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    c = shard_bytes(commits, 33, n, world, rank)
    p = shard_bytes(proofs, 525, n, world, rank)
    local = oracle_c.u64_verify_batch(gens64, c, p, LABEL, 2)
    full = gather_status(local, n)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, full == expect))


def test_sharded_verify_two_ranks_gloo(ref, oracle, gens64):
    import torch.multiprocessing as mp
    n = 7   # ragged split: 4 + 3
    xs, blinds, rngs = synth_batch(ref, n, start=40)
    proofs, _ = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, 4)
    commits = b"".join(oracle.u64_commit(gens64, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n))
    bad = bytearray(proofs)
    bad[525 * 2 + 400] ^= 1
    bad[525 * 5 + 470] ^= 1
    expect = oracle.u64_verify_batch(gens64, commits, bytes(bad), LABEL, 4)
    assert expect == [1, 1, 0, 1, 1, 0, 1]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, gens64, commits, bytes(bad), expect, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def _msm_worker(rank, world, port, pts, sc, expect, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle_c
    import bp_pp_b200.api as api
    from bp_pp_b200.shard import msm_sharded
    # no GPU in CI: stand the oracle in for the two device calls; the split / all_gather / combine logic is what is under test
    api.msm = lambda P, S, points_fmt=1, out_fmt=0, device=0: oracle_c.msm(P, S)
    api.points_sum = lambda P, points_fmt=0, out_fmt=0, device=0: oracle_c.msm(b"".join(oracle_c.point_decompress(P[33 * i:33 * i + 33]) for i in range(len(P) // 33)),
                                                                               (1).to_bytes(32, "big") * (len(P) // 33))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    res = msm_sharded(pts, sc, 0)
    dist.barrier(); dist.destroy_process_group()
    q.put((rank, res == expect))


def test_msm_split_by_point_range_two_ranks_gloo(ref, oracle):
    import random
    import torch.multiprocessing as mp
    from conftest import xy
    rnd = random.Random(3)
    n = 9   # ragged: 5 + 4
    p, qq = xy(ref.pt_mul(ref.G, 5)), xy(ref.pt_mul(ref.G, 9))
    pts = []
    for _ in range(n):
        pts.append(p); p = oracle.point_add(p, qq)
    pts = b"".join(pts)
    sc = b"".join(rnd.randrange(ref.N).to_bytes(32, "big") for _ in range(n))
    expect = oracle.msm(pts, sc)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_msm_worker, args=(r, 2, port, pts, sc, expect, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bp_pp_b200.shard import _gather_bytes
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = bytes([rank + 1]) * 128                       # one block's shares of X and R
    got = _gather_bytes(mine, 0)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, got == [bytes([r + 1]) * 128 for r in range(world)]))


def test_wnla_exchange_step_two_ranks_gloo():
    """The sharded WNLA's only data-path exchange: an all-gather of one equal-length byte string per rank, in rank order
    (world_size 2, gloo on the CPU; NCCL on the GPU box -- tests/test_gpu_multirank.py runs the whole protocol)."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
