"""Repeated wall-clock timing of the generic WNLA entry points at n = 2^20 (first call vs warm calls) and of an MSM upload."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bp_pp_b200 as B
from bp_pp_b200 import synth


class R:
    N = synth.N


def rand_scalars(rnd, n):
    a = np.frombuffer(np.random.default_rng(rnd.randrange(1 << 30)).bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    a[:, 0] &= 0x7F
    return a.tobytes()


rnd = random.Random(1)
be = lambda v: (v % synth.N).to_bytes(32, "big")  # noqa: E731
base, step = B.msm(synth.G64, be(11), B.FMT_AFFINE64, B.FMT_AFFINE64), B.msm(synth.G64, be(29), B.FMT_AFFINE64, B.FMT_AFFINE64)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
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
