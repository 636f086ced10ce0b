"""GPU suite, two processes (one per GPU when the box has two, else both on GPU 0; gloo rendezvous so that a single-GPU
box can run it): what BASELINE configs 3 and 5 promise across GPU counts -- each rank's slice of THE seeded batch gives
the C oracle's bytes (so the concatenation is identical for every N), and a WNLA instance cut into one block per rank
gives the single-GPU proof."""
import hashlib
import json
import os
import socket
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
LABEL = b"u64 range proof"


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import numpy as np
        import torch
        import torch.distributed as dist
        import bp_pp_b200 as B
        from bp_pp_b200 import synth
        from bp_pp_b200.shard import PeerGroup, wnla_prove_sharded
        from bp_pp_b200.transcript import Transcript
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "u64_batch_golden.json")))
        gens = synth.synth_generators64(dev)
        assert hashlib.sha256(gens).hexdigest() == gold["generators_sha256"]
        # ---- config 3: this rank's block of the seeded batch (blocks of 4,096: the golden file's granularity) ----
        m = gold["block"]
        ctx = B.Context(gens, dev, 12, m)
        xs, blinds, rng = synth.synth_batch(m, rank * m)
        commits = ctx.commit_batch(xs.tolist(), blinds.tobytes())
        proofs, st = ctx.prove_batch(xs.tolist(), blinds.tobytes(), rng.tobytes(), LABEL)
        ok = st == [1] * m and synth.block_hashes(proofs, 525) == [gold["proof_block_sha256"][rank]] and synth.block_hashes(commits, 33) == [gold["commit_block_sha256"][rank]]
        bad, bcom, idx = synth.tamper_batch(proofs, commits, synth.engine_add_g(dev), rank * m)
        expect = [1] * m
        for i in idx:
            expect[i] = gold["tampered_verdicts"][(rank * m + i) // gold["tamper_every"]]
        ok = ok and ctx.verify_batch(bcom, bad, LABEL) == expect
        ctx.close()
        # ---- config 5: one WNLA block per rank against the single-GPU prover ----
        n = 1 << 12
        be = lambda v: (v % synth.N).to_bytes(32, "big")  # noqa: E731
        base, step = B.msm(synth.G64, be(11), B.FMT_AFFINE64, B.FMT_AFFINE64, dev), B.msm(synth.G64, be(29), B.FMT_AFFINE64, B.FMT_AFFINE64, dev)
        pts = B.points_generate(base, step, 2 * n + 1, dev)
        g, gvec, hvec = pts[:64], pts[64:64 * (n + 1)], pts[64 * (n + 1):]
        rnd = np.random.default_rng(12)

        def scalars():
            a = np.frombuffer(rnd.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
            a[:, 0] &= 0x7F
            return a.tobytes()
        c, l, nn = scalars(), scalars(), scalars()
        rho, mu = be(0xABCDEF123456789), be(0xABCDEF123456789 ** 2)
        per = n // world
        sl = lambda b, item: b[item * per * rank:item * per * (rank + 1)]      # noqa: E731
        stats = {}
        peer = PeerGroup(dev)            # mailboxes in both ranks' HBM (CUDA IPC between the two processes): the library's own exchange kernels
        got = wnla_prove_sharded(g, [dict(hvec64=sl(hvec, 64), c32=sl(c, 32), l32=sl(l, 32), gvec64=sl(gvec, 64), n32=sl(nn, 32))], rho, mu, None,
                                 Transcript(b"two ranks"), [dev], stats, peer)
        w = B.WeightNormLinearArgument(g, gvec, hvec, c, rho, mu, device=dev)
        com = w.commit(l, nn)
        ok = ok and stats["commitment33"] == com and got == w.prove(com, b"two ranks", l, nn) and stats["rounds_sharded"] >= 10
        # ---- the split MSM: block Pippenger + remote stores of the partial sums + their reduction, fused on each GPU's stream ----
        up = B.UploadedMsm(sl(hvec, 64), sl(l, 32), device=dev)
        for _ in range(3):               # the epoch counter and both slot parities
            total, ms = peer.msm_allsum(up)
            ok = ok and total == B.msm(hvec, l, device=dev) and ms > 0
        ok = ok and peer.allgather(bytes([rank + 1]) * 240) == [bytes([r + 1]) * 240 for r in range(world)]
        up.close(); peer.close()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:          # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


def test_two_ranks_reproduce_the_oracle_and_the_single_gpu_proof():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
    assert [(r, ok) for r, ok, _ in res] == [(0, True), (1, True)], [msg for _, _, msg in res]
