"""Device time of one verify / prove step against the batch size on one GPU (what each rank sees under strong scaling).

  python tools/batch_sweep.py [--sizes 2048,8192,...] [--window-bits 20] [--profile]

Prints one JSON object; --profile adds the per-kernel CUDA-event times of one step at every size.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LABEL = b"u64 range proof"


def main():
    import numpy as np
    import torch
    import bp_pp_b200 as B
    from bp_pp_b200.synth import synth_generators64

    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2048,4096,8192,16384,32768,65536")
    ap.add_argument("--window-bits", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--lane-sweep", action="store_true", help="time every (ladder lanes, fixed-base lanes) override at every size")
    ap.add_argument("--inflight", default="1,2,4,8", help="independent batches in flight (one shared-table context and stream each)")
    args = ap.parse_args()
    sizes = [int(s) for s in args.sizes.split(",")]
    nmax = max(sizes)
    gens = synth_generators64(0)
    ctx = B.Context(gens, 0, args.window_bits, nmax)
    rnd = np.random.default_rng(7)
    xs = rnd.integers(0, 2**64, size=nmax, dtype=np.uint64)
    blinds = np.frombuffer(rnd.bytes(32 * nmax), dtype=np.uint8).reshape(nmax, 32).copy()
    blinds[:, 0] &= 0x7F
    rng = np.frombuffer(rnd.bytes(3328 * nmax), dtype=np.uint8).copy()
    commits = np.frombuffer(ctx.commit_batch(xs.tolist(), blinds.tobytes()), dtype=np.uint8).copy()
    proofs, st = ctx.prove_batch(xs.tolist(), blinds.tobytes(), rng.tobytes(), LABEL)
    assert all(s == 1 for s in st)
    proofs = np.frombuffer(proofs, dtype=np.uint8).copy()
    dev = torch.device("cuda", 0)
    d_commits, d_proofs = torch.from_numpy(commits).to(dev), torch.from_numpy(proofs).to(dev)
    d_status = torch.empty(nmax, dtype=torch.int32, device=dev)
    d_x = torch.from_numpy(xs.view(np.int64)).to(dev)
    d_blinds, d_rng = torch.from_numpy(blinds).to(dev), torch.from_numpy(rng).to(dev)
    d_out = torch.empty(nmax * 525, dtype=torch.uint8, device=dev)
    d_pst = torch.empty(nmax, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    out = {"window_bits": args.window_bits, "sizes": {}}
    for n in sizes:
        def vstep():
            ctx.verify_batch_dev(n, d_commits.data_ptr(), d_proofs.data_ptr(), LABEL, d_status.data_ptr(), stream=stream.cuda_stream)

        def pstep():
            ctx.prove_batch_dev(n, d_x.data_ptr(), d_blinds.data_ptr(), d_rng.data_ptr(), LABEL, d_out.data_ptr(), d_pst.data_ptr(),
                                stream=stream.cuda_stream)

        def timed(step):
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(args.steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); step(); e1.record(stream); e1.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / args.steps

        v, p = timed(vstep), timed(pstep)
        assert bool((d_status[:n] == 1).all()) and bool((d_pst[:n] == 1).all())
        assert bytes(d_out[:525 * n].cpu().numpy()) == proofs[:525 * n].tobytes()
        rec = {"verify_ms": round(v, 3), "prove_ms": round(p, 3), "verify_per_s": round(n / v * 1e3), "prove_per_s": round(n / p * 1e3)}
        if args.profile:
            ctx.profile_begin(); vstep(); pv = ctx.profile_end()
            ctx.profile_begin(); pstep(); pp = ctx.profile_end()
            rec["kernels_verify_ms"] = {k: [round(ms, 3), c] for k, (ms, c) in sorted(pv.items(), key=lambda kv: -kv[1][0])}
            rec["kernels_prove_ms"] = {k: [round(ms, 3), c] for k, (ms, c) in sorted(pp.items(), key=lambda kv: -kv[1][0])}
        out["sizes"][str(n)] = rec
    if args.lane_sweep:
        out["lane_sweep"] = {}
        for vl, ml in [(1, 4), (2, 4), (4, 4), (1, 8), (1, 16), (2, 8), (4, 8), (4, 16)]:
            os.environ["BPPP_VAR_LANES_RT"], os.environ["BPPP_MSM_LANES_RT"] = str(vl), str(ml)
            c2 = ctx.shared(nmax)
            for n in sizes:
                def t(step):
                    for _ in range(2):
                        step()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(args.steps):
                        step()
                    e1.record(stream); e1.synchronize()
                    return e0.elapsed_time(e1) / args.steps
                v = t(lambda: c2.verify_batch_dev(n, d_commits.data_ptr(), d_proofs.data_ptr(), LABEL, d_status.data_ptr(), stream=stream.cuda_stream))
                p = t(lambda: c2.prove_batch_dev(n, d_x.data_ptr(), d_blinds.data_ptr(), d_rng.data_ptr(), LABEL, d_out.data_ptr(), d_pst.data_ptr(),
                                                 stream=stream.cuda_stream))
                out["lane_sweep"][f"var{vl}_msm{ml}_n{n}"] = [round(v, 3), round(p, 3)]
            c2.close()
        del os.environ["BPPP_VAR_LANES_RT"], os.environ["BPPP_MSM_LANES_RT"]
    # ---- throughput with S independent batches in flight: S contexts sharing the tables, one stream each ----
    S_list = [int(v) for v in args.inflight.split(",") if int(v) > 1]
    smax = max(S_list) if S_list else 1
    ctxs = [ctx] + [ctx.shared(nmax) for _ in range(smax - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(smax)]
    d_sts = [torch.empty(nmax, dtype=torch.int32, device=dev) for _ in range(smax)]
    d_outs = [torch.empty(nmax * 525, dtype=torch.uint8, device=dev) for _ in range(smax)]
    for n in sizes:
        for S in S_list:
            def run(kind, steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                for s_ in streams[:S]:
                    s_.wait_event(e0)
                for k in range(steps):
                    j = k % S
                    if kind == "v":
                        ctxs[j].verify_batch_dev(n, d_commits.data_ptr(), d_proofs.data_ptr(), LABEL, d_sts[j].data_ptr(), stream=streams[j].cuda_stream)
                    else:
                        ctxs[j].prove_batch_dev(n, d_x.data_ptr(), d_blinds.data_ptr(), d_rng.data_ptr(), LABEL, d_outs[j].data_ptr(), d_sts[j].data_ptr(),
                                                stream=streams[j].cuda_stream)
                for s_ in streams[:S]:
                    stream.wait_stream(s_)
                e1.record(stream); e1.synchronize()
                return e0.elapsed_time(e1) / steps
            steps = max(args.steps, 2 * S)
            run("v", S); v = run("v", steps)
            run("p", S); p = run("p", steps)
            ok = all(bytes(d_outs[j][:525 * n].cpu().numpy()) == proofs[:525 * n].tobytes() for j in range(S))
            out["sizes"][str(n)][f"inflight{S}"] = {"verify_ms_per_batch": round(v, 3), "prove_ms_per_batch": round(p, 3),
                                                     "verify_per_s": round(n / v * 1e3), "prove_per_s": round(n / p * 1e3), "bytes_ok": ok}
    big = out["sizes"][str(nmax)]
    for n in sizes:
        r = out["sizes"][str(n)]
        r["verify_eff_vs_largest"] = round(r["verify_per_s"] / big["verify_per_s"], 3)
        r["prove_eff_vs_largest"] = round(r["prove_per_s"] / big["prove_per_s"], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
