"""Times verify/prove (device-resident, CUDA events, L2 flushed between steps) for the library named by BPPP_LIB.
BPPP_PROFILE=1 adds per-kernel CUDA-event times, BPPP_MICRO=1 the field/point microbenchmarks.  The proof bytes are
hashed so that kernel variants can be checked against each other (they must not differ)."""
import hashlib
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    ctx = B.Context(synth_generators64(0), 0, int(os.environ.get("BPPP_W", "20")), n)
    rnd = np.random.default_rng(7)
    xs = rnd.integers(0, 2**64, size=n, dtype=np.uint64)
    blinds = np.frombuffer(rnd.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    blinds[:, 0] &= 0x7F
    rng = np.frombuffer(rnd.bytes(3328 * n), dtype=np.uint8).copy()
    commits = np.frombuffer(ctx.commit_batch(xs.tolist(), blinds.tobytes()), dtype=np.uint8).copy()
    proofs, st = ctx.prove_batch(xs.tolist(), blinds.tobytes(), rng.tobytes(), LABEL)
    assert all(s == 1 for s in st)
    proofs = np.frombuffer(proofs, dtype=np.uint8).copy()
    bad = proofs.copy()
    bad[525 * 5 + 400] ^= 1                      # one tampered scalar: the verdict vector must show exactly this one
    dev = torch.device("cuda", 0)
    d_commits, d_proofs = torch.from_numpy(commits).to(dev), torch.from_numpy(bad).to(dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    d_x = torch.from_numpy(xs.view(np.int64)).to(dev)
    d_blinds, d_rng = torch.from_numpy(blinds).to(dev), torch.from_numpy(rng).to(dev)
    d_out = torch.empty(n * 525, dtype=torch.uint8, device=dev)
    d_pst = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def v():
        ctx.verify_batch_dev(n, d_commits.data_ptr(), d_proofs.data_ptr(), LABEL, d_status.data_ptr(), stream=st.cuda_stream)

    def p():
        ctx.prove_batch_dev(n, d_x.data_ptr(), d_blinds.data_ptr(), d_rng.data_ptr(), LABEL, d_out.data_ptr(), d_pst.data_ptr(), stream=st.cuda_stream)

    def t(fn, reps):
        fn(); fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    vm = t(v, 5)
    verdicts = d_status.cpu().numpy()
    ok = bool(verdicts[5] == 0 and (np.delete(verdicts, 5) == 1).all())
    pm = t(p, 3)
    pok = bool((d_pst.cpu().numpy() == 1).all()) and bytes(d_out.cpu().numpy()) == proofs.tobytes()
    res = {"lib": os.path.basename(os.environ.get("BPPP_LIB", "libbppp.so")), "nsub": os.environ.get("BPPP_NSUB", "default"), "n": n,
           "verify_ms": round(vm, 3), "verify_per_s": round(n / vm * 1e3), "prove_ms": round(pm, 3), "prove_per_s": round(n / pm * 1e3),
           "ok": ok and pok, "proofs_sha256_16": hashlib.sha256(proofs.tobytes()).hexdigest()[:16]}
    if os.environ.get("BPPP_PROFILE"):
        ctx.profile_begin(); v(); pv = ctx.profile_end()
        res["kernels_verify"] = {k: round(ms, 3) for k, (ms, c) in sorted(pv.items(), key=lambda kv: -kv[1][0])[:12]}
        ctx.profile_begin(); p(); pp = ctx.profile_end()
        res["kernels_prove"] = {k: round(ms, 3) for k, (ms, c) in sorted(pp.items(), key=lambda kv: -kv[1][0])[:6]}
    if os.environ.get("BPPP_MICRO"):
        res["microbench"] = {k: float(f"{v:.4g}") for k, v in B.microbench(0).items()}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
