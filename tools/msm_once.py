"""One resident 2^k-point MSM (for `ncu` launch lists / captures):  python tools/msm_once.py [log2n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bp_pp_b200 as B  # noqa: E402
from bp_pp_b200 import synth  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 21)
be = lambda v: (v % synth.N).to_bytes(32, "big")  # noqa: E731
pts = B.points_generate(B.msm(synth.G64, be(11), B.FMT_AFFINE64, B.FMT_AFFINE64), B.msm(synth.G64, be(29), B.FMT_AFFINE64, B.FMT_AFFINE64), n)
sc = np.frombuffer(np.random.default_rng(1).bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
sc[:, 0] &= 0x7F
up = B.UploadedMsm(pts, sc.tobytes())
print(up.run()[1]); print(up.run()[1])
