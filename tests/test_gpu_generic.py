"""GPU suite (`-m gpu`): generic (arbitrary-size) entry points against the C oracle: variable-base MSM, WNLA,
reciprocal range proofs of other dimensions, the generic arithmetic circuit."""
import os
import random

import pytest

from conftest import xy

pytestmark = pytest.mark.gpu


def _points(oracle, ref, n, seed=1):
    """n distinct points by repeated addition (fast through the C oracle)."""
    rnd = random.Random(seed)
    p = xy(ref.pt_mul(ref.G, rnd.randrange(1, ref.N)))
    q = xy(ref.pt_mul(ref.G, rnd.randrange(1, ref.N)))
    out = []
    for _ in range(n):
        out.append(p)
        p = oracle.point_add(p, q)
    return out


def _be(v):
    return v.to_bytes(32, "big")


@pytest.mark.parametrize("n", [0, 1, 2, 7, 64, 1024, 1025, 1500, 5000, 20000])
def test_msm_matches_naive_vector_mul(oracle, ref, n):
    import bp_pp_b200 as B
    rnd = random.Random(100 + n)
    pts = _points(oracle, ref, n, seed=n + 3)
    sc = [rnd.randrange(ref.N) for _ in range(n)]
    for i in range(0, n, 7):
        sc[i] = [0, 1, 2, ref.N - 1, ref.N - 2, 2**128, 15][i % 7]         # edge scalars
    if n >= 16:
        pts[5] = pts[9]                                                      # repeated point
        pts[11] = b"\0" * 64                                                 # identity among the inputs
    P, S = b"".join(pts), b"".join(_be(s) for s in sc)
    assert B.msm(P, S) == oracle.msm(P, S) if n else B.msm(P, S) == b"\0" * 33


def test_msm_zero_extension_and_degenerate_buckets(oracle, ref):
    import bp_pp_b200 as B
    n = 3000
    pts = _points(oracle, ref, n, seed=77)
    P = b"".join(pts)
    # every scalar equal and tiny: one giant bucket in window 0 (exercises the segment merge path)
    S = _be(3) * n
    assert B.msm(P, S) == oracle.msm(P, S)
    # all points identical with opposite scalars: sum is the identity
    P2 = pts[0] * 2000
    S2 = (_be(5) + _be(ref.N - 5)) * 1000
    assert B.msm(P2, S2) == b"\0" * 33
    # lengths differ: the shorter operand is zero-extended (util.rs:24-26,52-53)
    S3 = b"".join(_be(random.Random(5).randrange(ref.N)) for _ in range(1200))
    assert B.msm(P, S3) == oracle.msm(P[:64 * 1200], S3)
    assert B.msm(P[:64 * 1100], S3) == oracle.msm(P[:64 * 1100], S3[:32 * 1100])
    # compressed input format
    comp = b"".join(oracle.point_compress(p) for p in pts[:1300])
    assert B.msm(comp, S3, points_fmt=B.FMT_COMPRESSED) == oracle.msm(P[:64 * 1200], S3)
    with pytest.raises(B.BpppError):
        B.msm(b"\x01" * 64, _be(1))                                          # off-curve point
    with pytest.raises(B.BpppError):
        B.msm(pts[0], b"\xff" * 32)                                          # scalar >= n


def test_points_sum_combines_partial_sums(oracle, ref):
    import bp_pp_b200 as B
    n = 4000
    rnd = random.Random(9)
    pts = _points(oracle, ref, n, seed=9)
    sc = [_be(rnd.randrange(ref.N)) for _ in range(n)]
    full = B.msm(b"".join(pts), b"".join(sc))
    parts = [B.msm(b"".join(pts[a:a + 500]), b"".join(sc[a:a + 500])) for a in range(0, n, 500)]   # 8 "ranks"
    assert B.points_sum(b"".join(parts)) == full


def _wnla_instance(oracle, ref, gn, hn, ln, nn, seed):
    rnd = random.Random(seed)
    pts = _points(oracle, ref, 1 + gn + hn, seed=seed)
    g, gvec, hvec = pts[0], b"".join(pts[1:1 + gn]), b"".join(pts[1 + gn:])
    c = b"".join(_be(rnd.randrange(ref.N)) for _ in range(hn))
    rho = rnd.randrange(1, ref.N)
    mu = rho * rho % ref.N
    l = b"".join(_be(rnd.randrange(ref.N)) for _ in range(ln))
    n = b"".join(_be(rnd.randrange(ref.N)) for _ in range(nn))
    return g, gvec, hvec, c, _be(rho), _be(mu), l, n


@pytest.mark.parametrize("gn,hn,ln,nn", [(4, 4, 4, 4), (8, 8, 8, 8), (16, 32, 32, 16), (5, 7, 7, 5), (3, 9, 6, 2), (64, 64, 64, 64), (1200, 1200, 1200, 1200)])
def test_wnla_commit_prove_verify_match_oracle(oracle, ref, gn, hn, ln, nn):
    import bp_pp_b200 as B
    g, gvec, hvec, c, rho, mu, l, n = _wnla_instance(oracle, ref, gn, hn, ln, nn, seed=1000 + gn + hn)
    label = b"wnla test"
    w = B.WeightNormLinearArgument(g, gvec, hvec, c, rho, mu)
    com = oracle.wnla_commit(g, gvec, hvec, c, rho, mu, l, n)
    assert w.commit(l, n) == com
    r, x, lo, no = w.prove(com, label, l, n)
    assert (r, x, lo, no) == oracle.wnla_prove(g, gvec, hvec, c, rho, mu, com, l, n, label)
    # when |l| != |h_vec| the prover absorbs l.len() (wnla.rs:165) but the verifier |h_vec| (wnla.rs:91): the reference
    # itself then rejects its own proof; parity with the oracle is what is asserted
    expect = oracle.wnla_verify(g, gvec, hvec, c, rho, mu, com, r, x, lo, no, label)
    assert expect == (1 if (ln == hn and nn == gn) or len(r) == 0 else 0)
    assert w.verify(com, label, r, x, lo, no) == expect
    if len(lo):
        bad = bytearray(lo); bad[31] ^= 1
        assert w.verify(com, label, r, x, bytes(bad), no) == 0
    if len(r) >= 33:
        swapped = x[:33] + r[33:]
        assert w.verify(com, label, swapped, x, lo, no) == oracle.wnla_verify(g, gvec, hvec, c, rho, mu, com, swapped, x, lo, no, label)
        assert w.verify(com, label, r + r[:33], x, lo, no) == 0          # x.len() != r.len()
    assert w.verify(com, b"other label", r, x, lo, no) == (expect if len(r) == 0 else 0)


def test_wnla_golden_fixture(oracle):
    import json
    import bp_pp_b200 as B
    from conftest import ROOT
    wn = json.load(open(os.path.join(ROOT, "tests", "golden", "wnla_golden.json")))
    b = bytes.fromhex
    rho = b(wn["rho"]); rho_i = int.from_bytes(rho, "big")
    N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    w = B.WeightNormLinearArgument(b(wn["g"]), b"".join(b(p) for p in wn["g_vec"]), b"".join(b(p) for p in wn["h_vec"]),
                                   b"".join(b(v) for v in wn["c"]), rho, _be(rho_i * rho_i % N))
    l = b"".join(_be(v) for v in wn["l"]); n = b"".join(_be(v) for v in wn["n"])
    assert w.commit(l, n).hex() == wn["commitment"]
    r, x, lo, no = w.prove(b(wn["commitment"]), b"wnla test", l, n)
    assert (r + x + lo + no).hex() == wn["proof"]
    assert w.verify(b(wn["commitment"]), b"wnla test", r, x, lo, no) == 1
