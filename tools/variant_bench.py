"""Times verify/prove (device-resident, CUDA events) for the library named by BPPP_LIB and the BPPP_NSUB setting."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bp_pp_b200 as B, bppp_ref as R
import bench

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    g, gv, hv = R.synth_generators()
    gens = b"".join(bench.xy(p) for p in [g] + gv + hv)
    ctx = B.Context(gens, 0, int(os.environ.get("BPPP_W", "16")), n)
    xs, blinds, rng, commits, proofs, expect = bench.make_workload(ctx, n, R)
    dev = torch.device("cuda", 0)
    d_commits = torch.from_numpy(commits).to(dev); d_proofs = torch.from_numpy(proofs).to(dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    d_x = torch.from_numpy(xs.view(np.int64)).to(dev); d_blinds = torch.from_numpy(blinds).to(dev); d_rng = torch.from_numpy(rng).to(dev)
    d_out = torch.empty(n * 525, dtype=torch.uint8, device=dev); d_pst = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    def v(): ctx.verify_batch_dev(n, d_commits.data_ptr(), d_proofs.data_ptr(), bench.LABEL, d_status.data_ptr(), stream=st.cuda_stream)
    def p(): ctx.prove_batch_dev(n, d_x.data_ptr(), d_blinds.data_ptr(), d_rng.data_ptr(), bench.LABEL, d_out.data_ptr(), d_pst.data_ptr(), stream=st.cuda_stream)
    def t(fn, reps):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st); e1.synchronize()
        return e0.elapsed_time(e1) / reps
    vm = t(v, 4); ok = bool((d_status.cpu().numpy() == expect).all())
    pm = t(p, 2); pok = bool((d_pst.cpu().numpy() == 1).all())
    res = {"lib": os.path.basename(os.environ.get("BPPP_LIB", "libbppp.so")), "nsub": os.environ.get("BPPP_NSUB", "4"), "verify_ms": round(vm, 2),
           "verify_per_s": round(n / vm * 1e3), "prove_ms": round(pm, 2), "prove_per_s": round(n / pm * 1e3), "ok": ok and pok}
    if os.environ.get("BPPP_PROFILE"):
        ctx.profile_begin(); v(); pv = ctx.profile_end()
        res["kernels_verify"] = {k: round(ms, 2) for k, (ms, c) in sorted(pv.items(), key=lambda kv: -kv[1][0])[:12]}
        ctx.profile_begin(); p(); pp = ctx.profile_end()
        res["kernels_prove"] = {k: round(ms, 2) for k, (ms, c) in sorted(pp.items(), key=lambda kv: -kv[1][0])[:4]}
    print(json.dumps(res), flush=True)

if __name__ == "__main__":
    main()
