import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bp_pp_b200 as B, bppp_ref as R
from tools.bench_generic import xy, rand_scalars
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
rnd = random.Random(1)
pts = B.points_generate(xy(R.pt_mul(R.G, 11)), xy(R.pt_mul(R.G, 29)), n)
up = B.UploadedMsm(pts, rand_scalars(rnd, n))
print(up.run()[1]); print(up.run()[1])
