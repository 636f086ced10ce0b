"""Small end-to-end run for compute-sanitizer (memcheck): u64 commit/prove/verify, MSM (both paths), WNLA, reciprocal."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bp_pp_b200 as B, bppp_ref as R
def xy(p): return p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")
g, gv, hv = R.synth_generators()
proto = B.U64RangeProofProtocol(xy(g), [xy(p) for p in gv], [xy(p) for p in hv], window_bits=4, max_batch=40)
n = 33
xs = [R.synth_x(i) for i in range(n)]
blinds = b"".join(R.sc_to_bytes(R.synth_blind(i)) for i in range(n))
rng = b"".join(R.synth_rng_bytes(i) for i in range(n))
commits = proto.commit_batch(xs, blinds)
proofs, st = proto.prove_batch(xs, blinds, rng, b"u64 range proof")
bad = bytearray(proofs); bad[7] ^= 1; bad[525 + 400] ^= 1
print("u64:", proto.verify_batch(commits, bytes(bad), b"u64 range proof")[:4], all(s == 1 for s in st))
rnd = random.Random(1)
pts = B.points_generate(xy(g), xy(gv[0]), 1300)
sc = b"".join(rnd.randrange(R.N).to_bytes(32, "big") for _ in range(1300))
print("msm:", B.msm(pts, sc).hex()[:16], B.msm(pts[:64 * 50], sc[:32 * 50]).hex()[:16])
# skewed scalars: medium-heavy buckets and multi-segment buckets (k_msm_heavy / k_msm_heavy_sum)
pts7 = B.points_generate(xy(g), xy(gv[1]), 7000)
sc7 = b"".join((rnd.choice((1, 2, 2**40 + 3)) if i % 50 else rnd.randrange(R.N)).to_bytes(32, "big") for i in range(7000))
print("msm skewed:", B.msm(pts7, sc7).hex()[:16])
# signed fixed-base windows
proto_s = B.U64RangeProofProtocol(xy(g), [xy(p) for p in gv], [xy(p) for p in hv], window_bits=-5, max_batch=40)
ps, sts = proto_s.prove_batch(xs, blinds, rng, b"u64 range proof")
print("signed windows:", ps == proofs, proto_s.verify_batch(commits, ps, b"u64 range proof") == [1] * n)
w = B.WeightNormLinearArgument(pts[:64], pts[64:64 * 9], pts[64 * 9:64 * 17], sc[:32 * 8], sc[32 * 8:32 * 9], sc[32 * 9:32 * 10])
l, nn = sc[32 * 10:32 * 18], sc[32 * 18:32 * 26]
com = w.commit(l, nn)
r, x, lo, no = w.prove(com, b"t", l, nn)
print("wnla:", w.verify(com, b"t", r, x, lo, no))
rp = B.ReciprocalRangeProofProtocol(4, 4, pts[:64], pts[64:64 * 5], pts[64 * 5:64 * 19], b"", pts[64 * 19:64 * 21])
rec, ro, ll, nl, c33 = rp.prove((27).to_bytes(32, "big"), (9).to_bytes(32, "big"), [3, 2, 1, 0], rnd.randbytes(28 * 64), b"r")
print("reciprocal:", rp.verify(c33, rec, ro, ro, ll, nl, b"r"))
