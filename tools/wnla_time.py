"""Repeated wall-clock timing of the generic WNLA entry points at n = 2^20 (first call vs warm calls) and of an MSM upload."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bp_pp_b200 as B, bppp_ref as R
from tools.bench_generic import xy, rand_scalars
rnd = random.Random(1)
base, step = xy(R.pt_mul(R.G, 11)), xy(R.pt_mul(R.G, 29))
n = 1 << 20
pts = B.points_generate(base, step, 2 * n + 1)
g, gvec, hvec = pts[:64], pts[64:64 * (n + 1)], pts[64 * (n + 1):]
c, l, nn = rand_scalars(rnd, n), rand_scalars(rnd, n), rand_scalars(rnd, n)
rho = rnd.randrange(1, R.N)
w = B.WeightNormLinearArgument(g, gvec, hvec, c, rho.to_bytes(32, "big"), (rho * rho % R.N).to_bytes(32, "big"))
for i in range(3):
    t0 = time.perf_counter(); com = w.commit(l, nn); print("commit", round(time.perf_counter() - t0, 3))
for i in range(2):
    t0 = time.perf_counter(); out = w.prove(com, b"x", l, nn); print("prove", round(time.perf_counter() - t0, 3))
for i in range(2):
    t0 = time.perf_counter(); ok = w.verify(com, b"x", *out); print("verify", round(time.perf_counter() - t0, 3), ok)
t0 = time.perf_counter(); up = B.UploadedMsm(pts, rand_scalars(rnd, 2 * n + 1)); print("upload msm operands", round(time.perf_counter() - t0, 3))
t0 = time.perf_counter(); up.run(); print("msm run", round(time.perf_counter() - t0, 3))
t0 = time.perf_counter(); up.run(); print("msm run", round(time.perf_counter() - t0, 3))
