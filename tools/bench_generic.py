"""Measurements for the generic paths (BASELINE configs 4 and 5): MSM points/s, standalone WNLA prove/verify, wide reciprocal.
Single GPU:   python tools/bench_generic.py
Split MSM:    torchrun --nproc-per-node N tools/bench_generic.py --split     (point range per rank, partial sums all-gathered over NCCL)
Prints one JSON object."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bp_pp_b200 as B
import bppp_ref as R


def xy(p): return p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def rand_scalars(rnd, n):
    raw = bytearray(rnd.randbytes(32 * n))
    raw[0::32] = bytes(b & 0x7F for b in raw[0::32])
    return bytes(raw)


def main():
    split = "--split" in sys.argv
    rnd = random.Random(2026)
    base, step = xy(R.pt_mul(R.G, 11)), xy(R.pt_mul(R.G, 29))
    out = {}
    if split:
        import torch, torch.distributed as dist
        from bp_pp_b200.shard import msm_sharded
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        n = 1 << 21
        pts = B.points_generate(base, step, n, device=local)
        sc = rand_scalars(rnd, n)
        msm_sharded(pts, sc, local)     # warm-up
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); res = msm_sharded(pts, sc, local); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            single = B.msm(pts, sc, device=local)
            print(json.dumps({"msm_split": {"n_points": n, "n_gpus": world, "wall_ms_max_over_ranks_incl_h2d": round(float(t.item()) * 1e3, 2),
                                            "points_per_s": round(n / float(t.item())), "equals_single_gpu_result": res == single}}), flush=True)
        dist.barrier(); dist.destroy_process_group()
        return
    # ---- MSM points/s, operands resident in HBM, device time ----
    msm = {}
    nmax = 1 << 21
    pts = B.points_generate(base, step, nmax)
    sc = rand_scalars(rnd, nmax)
    for logn in (16, 20, 21):
        n = 1 << logn
        up = B.UploadedMsm(pts[:64 * n], sc[:32 * n])
        up.run()
        best = min(up.run()[1] for _ in range(3))
        msm[f"2^{logn}"] = {"ms": round(best, 3), "points_per_s": round(n / best * 1e3)}
        up.close()
    out["msm_points_per_s"] = msm
    # ---- standalone WNLA (config 5) ----
    wn = {}
    for logn in (16, 18, 20):
        n = 1 << logn
        g, gvec, hvec = pts[:64], pts[64:64 * (n + 1)], pts[64 * (n + 1):64 * (2 * n + 1)] if 2 * n + 1 <= nmax else B.points_generate(step, base, n)
        c, l, nn = rand_scalars(rnd, n), rand_scalars(rnd, n), rand_scalars(rnd, n)
        rho = rnd.randrange(1, R.N)
        w = B.WeightNormLinearArgument(g, gvec, hvec, c, rho.to_bytes(32, "big"), (rho * rho % R.N).to_bytes(32, "big"))
        def best_of(fn, reps=2):      # the first call of a size also grows the cached device scratch slab
            ts, out = [], None
            for _ in range(reps):
                t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
            return out, ts
        com, t_c = best_of(lambda: w.commit(l, nn))
        (r, x, lo, no), t_p = best_of(lambda: w.prove(com, b"wnla big", l, nn))
        ok, t_v = best_of(lambda: w.verify(com, b"wnla big", r, x, lo, no))
        wn[f"2^{logn}"] = {"commit_s": round(min(t_c), 3), "prove_s": round(min(t_p), 3), "verify_s": round(min(t_v), 3),
                           "first_call_s": {"commit": round(t_c[0], 3), "prove": round(t_p[0], 3), "verify": round(t_v[0], 3)},
                           "rounds": len(r) // 33, "verified": ok == 1,
                           "note": "wall clock through the host-buffer C ABI, uploads included; best of 2 calls"}
    out["wnla_standalone"] = wn
    # ---- wide reciprocal (config 4) ----
    nd, np_ = 1024, 16
    allp = pts[:64 * (1 + nd + nd + 10 + 1014)]
    g, gvec = allp[:64], allp[64:64 * (1 + nd)]
    hvec, hvec2 = allp[64 * (1 + nd):64 * (1 + nd + nd + 10)], allp[64 * (1 + nd + nd + 10):]
    digits = [rnd.randrange(np_) for _ in range(nd)]
    x = sum(d * pow(np_, i, R.N) for i, d in enumerate(digits)) % R.N
    proto = B.ReciprocalRangeProofProtocol(nd, np_, g, gvec, hvec, b"", hvec2)
    rng = rnd.randbytes((1 + 18 + nd + 1 + nd) * 64)
    t0 = time.perf_counter(); rec, rounds, ll, nl, com = proto.prove(x.to_bytes(32, "big"), (5).to_bytes(32, "big"), digits, rng, b"wide"); t_p = time.perf_counter() - t0
    t0 = time.perf_counter(); ok = proto.verify(com, rec, rounds, rounds, ll, nl, b"wide"); t_v = time.perf_counter() - t0
    out["reciprocal_dim1024"] = {"prove_s": round(t_p, 3), "verify_s": round(t_v, 3), "rounds": rounds, "verified": ok == 1}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
