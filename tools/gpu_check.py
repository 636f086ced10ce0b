"""First-light check on a GPU box: microbenchmarks, table build, commit/verify/prove parity vs the C oracle."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bp_pp_b200 as B
import bppp_ref as R, oracle_c as OC

def xy(p): return b"\0" * 64 if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    big = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    print("microbench", json.dumps(B.microbench(0)), flush=True)
    g, gv, hv = R.synth_generators()
    gens = b"".join(xy(p) for p in [g] + gv + hv)
    t0 = time.time()
    ctx = B.Context(gens, 0, W, max(big, n))
    print("ctx", time.time() - t0, ctx.info(), flush=True)
    xs = [R.synth_x(i) for i in range(n)]
    blinds = b"".join(R.sc_to_bytes(R.synth_blind(i)) for i in range(n))
    rngs = b"".join(R.synth_rng_bytes(i) for i in range(n))
    label = b"u64 range proof"
    OC.use_native()
    thr = os.cpu_count()
    t0 = time.time(); oproofs, ost = OC.u64_prove_batch(gens, xs, blinds, rngs, label, thr); t_op = time.time() - t0
    ocommits = b"".join(OC.u64_commit(gens, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n))
    print(f"oracle prove {n} proofs on {thr} threads: {t_op:.2f}s", flush=True)
    commits = ctx.commit_batch(xs, blinds)
    print("commit parity:", commits == ocommits, flush=True)
    st = ctx.verify_batch(ocommits, oproofs, label)
    print("verify honest:", sum(1 for s in st if s == 1), "of", n, flush=True)
    bad = bytearray(oproofs)
    for i in range(n):
        pos = (i * 37) % 525
        if i % 2 == 0: bad[525 * i + pos] ^= 1 << (i % 8)
    t0 = time.time(); ost2 = OC.u64_verify_batch(gens, ocommits, bytes(bad), label, thr); t_ov = time.time() - t0
    st2 = ctx.verify_batch(ocommits, bytes(bad), label)
    print(f"oracle verify {n} on {thr} threads: {t_ov:.2f}s; tamper verdict parity:", st2 == ost2, "hist", {k: st2.count(k) for k in set(st2)}, flush=True)
    if st2 != ost2:
        print([(i, a, b) for i, (a, b) in enumerate(zip(st2, ost2)) if a != b][:10])
    proofs, pst = ctx.prove_batch(xs, blinds, rngs, label)
    print("prove parity:", proofs == oproofs, "status ok:", all(s == 1 for s in pst), flush=True)
    if proofs != oproofs:
        for i in range(n):
            a, b = proofs[525 * i:525 * i + 525], oproofs[525 * i:525 * i + 525]
            if a != b:
                print("first mismatching proof", i, "byte", next(k for k in range(525) if a[k] != b[k])); break
    # throughput (host buffers, wall clock) at the bench size
    reps = (big + n - 1) // n
    bc, bp = (ocommits * reps)[:33 * big], (oproofs * reps)[:525 * big]
    for _ in range(2):
        t0 = time.time(); st = ctx.verify_batch(bc, bp, label); dt = time.time() - t0
        print(f"verify_batch {big}: {dt*1e3:.1f} ms -> {big/dt:.0f} proofs/s, all true: {all(s == 1 for s in st)}", flush=True)
    bx, bb, br = (xs * reps)[:big], (blinds * reps)[:32 * big], (rngs * reps)[:3328 * big]
    for _ in range(2):
        t0 = time.time(); pr, pst = ctx.prove_batch(bx, bb, br, label); dt = time.time() - t0
        print(f"prove_batch {big}: {dt*1e3:.1f} ms -> {big/dt:.0f} proofs/s, matches: {pr[:525*n] == oproofs}", flush=True)
    print("launches", ctx.launch_count())

if __name__ == "__main__":
    main()
