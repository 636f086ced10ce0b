"""CPU suite: the DEVICE arithmetic and per-proof phase logic (bp_pp_b200/csrc/*.cuh) compiled for the host
with the bound assertions of fe.cuh enabled, against the oracles.  This is how the kernels' code is exercised without a GPU."""
import ctypes as C
import random

import pytest

from conftest import synth_batch, xy

LABEL = b"u64 range proof"


def B(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


def O(n):
    return (C.c_uint8 * n)()


def be(x):
    return x.to_bytes(32, "big")


def test_field_ops(emu_prims, ref):
    L, P = emu_prims, ref.P
    rnd = random.Random(7)
    edge = [0, 1, 2, P - 1, P - 2, P, P + 1, 2**256 - 1, 2**255, 977, 2**32 + 977, (1 << 26) - 1, (1 << 52) - 1]
    vals = edge + [rnd.randrange(2**256) for _ in range(150)]
    for a in vals:
        o = O(32); L.emu_fe_norm(B(be(a)), o); assert int.from_bytes(o, "big") == a % P
        o = O(32); L.emu_fe_sqr(B(be(a)), o); assert int.from_bytes(o, "big") == a * a % P
        b, c, d = rnd.choice(vals), rnd.choice(vals), rnd.choice(vals)
        o = O(32); L.emu_fe_mul(B(be(a)), B(be(b)), o); assert int.from_bytes(o, "big") == a * b % P
        k1, k2 = rnd.randrange(1, 8), rnd.randrange(1, 8)
        o = O(32); L.emu_fe_expr(B(be(a)), B(be(b)), B(be(c)), B(be(d)), k1, k2, o)
        assert int.from_bytes(o, "big") == (a * k1 + b * k2 - c) * d % P
    for a in vals[:40]:
        o = O(32); L.emu_fe_inv(B(be(a)), o); assert int.from_bytes(o, "big") == pow(a % P, P - 2, P)
        sq = a * a % P
        o = O(32); ok = L.emu_fe_sqrt(B(be(sq)), o); r = int.from_bytes(o, "big")
        assert ok == 1 and r * r % P == sq


def test_safegcd_inversion_both_moduli(emu_prims, ref):
    """modinv.cuh (600 division steps in 20 batches of 30) against Python's modular inverse and against the Fermat powers it
    replaced: 0 -> 0, 1, -1, +-2, powers of two around the limb boundaries of both the 32-bit words and the 30-bit limbs,
    non-canonical field representatives in [p, 2^256), and 4,000 random values per modulus."""
    L = emu_prims
    rnd = random.Random(11)
    for field, M in ((0, ref.P), (1, ref.N)):
        edge = [0, 1, 2, 3, M - 1, M - 2, M - 3, (M + 1) // 2, (M - 1) // 2, 2**255, 2**255 - 1, 2**128, 2**128 - 1, M >> 1, M >> 128]
        edge += [2**k for k in (29, 30, 31, 32, 59, 60, 61, 64, 90, 120, 150, 180, 210, 239, 240, 241, 254)]
        edge += [(2**k - 1) for k in (30, 60, 90, 240, 250)]
        if field == 0:
            edge += [M, M + 1, M + 5, 2**256 - 1, 2**256 - 2, M + 2**32]          # representatives fe_inv must normalise first
        vals = edge + [rnd.randrange(M) for _ in range(4000)] + [rnd.randrange(1, 2**k) for k in range(1, 256, 3)]
        vals = [v % 2**256 if field == 0 else v % M for v in vals]
        buf = b"".join(be(v) for v in vals)
        out = O(32 * len(vals))
        L.emu_inv_many(field, C.c_size_t(len(vals)), B(buf), out)
        out = bytes(out)
        for k, v in enumerate(vals):
            got = int.from_bytes(out[32 * k:32 * k + 32], "big")
            want = pow(v % M, -1, M) if v % M else 0
            assert got == want, (field, hex(v))
        for v in edge[:12] + vals[-20:]:
            o = O(32)
            (L.emu_fe_inv_fermat if field == 0 else L.emu_sc_inv_fermat)(B(be(v)), o)
            assert int.from_bytes(o, "big") == (pow(v % M, -1, M) if v % M else 0)


def test_field_ops_on_non_canonical_representatives(emu_prims, ref):
    """every pair of edge values, including the representatives in [p, 2^256) that trigger the rare double wrap
    of fe_add / double borrow of fe_sub / second fold of the product reduction"""
    L, P = emu_prims, ref.P
    top = 2**256
    edge = [0, 1, 976, 977, 978, 2**32 + 976, 2**32 + 977, 2**32 + 978, P - 1, P, P + 1, P + 977, top - 2**32, top - 978, top - 977, top - 2, top - 1,
            2**255, 2**224 - 1, (2**256 - 1) // 3]
    for a in edge:
        for b in edge:
            o = O(32); L.emu_fe_mul(B(be(a)), B(be(b)), o); assert int.from_bytes(o, "big") == a * b % P
            for c in (0, 1, P, top - 1, 2**32 + 977):
                o = O(32); L.emu_fe_expr(B(be(a)), B(be(b)), B(be(c)), B(be(1)), 1, 1, o)
                assert int.from_bytes(o, "big") == (a + b - c) % P
                o = O(32); L.emu_fe_expr(B(be(a)), B(be(b)), B(be(c)), B(be(top - 1)), 7, 8, o)
                assert int.from_bytes(o, "big") == (7 * a + 8 * b - c) * (top - 1) % P


def test_field_and_scalar_ops_property(emu_prims, ref):
    """hypothesis: the device field / scalar code (host build) against Python integers on operands biased towards limb
    boundaries (all-ones / all-zero limbs, values around p, n and 2^256)."""
    from hypothesis import given, settings, strategies as st
    L, P, N = emu_prims, ref.P, ref.N
    limb = st.sampled_from([0, 1, 2, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 977, 0xFFFFFC2F]) | st.integers(0, 2**32 - 1)
    words = st.lists(limb, min_size=8, max_size=8).map(lambda ws: sum(w << (32 * k) for k, w in enumerate(ws)))
    near = st.sampled_from([0, P, N, 2**256 - 1, 2**255]).flatmap(lambda c: st.integers(-2**33, 2**33).map(lambda d: (c + d) % 2**256))
    val = words | near | st.integers(0, 2**256 - 1)

    @settings(max_examples=400, deadline=None)
    @given(val, val, val, st.integers(1, 8), st.integers(1, 8))
    def check(a, b, c, k1, k2):
        o = O(32); L.emu_fe_mul(B(be(a)), B(be(b)), o); assert int.from_bytes(o, "big") == a * b % P
        o = O(32); L.emu_fe_sqr(B(be(a)), o); assert int.from_bytes(o, "big") == a * a % P
        o = O(32); L.emu_fe_expr(B(be(a)), B(be(b)), B(be(c)), B(be(b)), k1, k2, o)
        assert int.from_bytes(o, "big") == (a * k1 + b * k2 - c) * b % P
        sa, sb = a % N, b % N
        o = O(32); L.emu_sc_mul(B(be(sa)), B(be(sb)), o); assert int.from_bytes(o, "big") == sa * sb % N
        o = O(32); L.emu_sc_wide(B(be(a) + be(b)), o); assert int.from_bytes(o, "big") == ((a << 256) | b) % N
        o = O(32); L.emu_sc_sub(B(be(sa)), B(be(sb)), o); assert int.from_bytes(o, "big") == (sa - sb) % N

    check()


def test_scalar_ops(emu_prims, ref):
    L, N = emu_prims, ref.N
    rnd = random.Random(8)
    vals = [0, 1, 2, N - 1, N - 2, 2**255, 2**128, 2**129 - 1, (1 << 32) - 1] + [rnd.randrange(N) for _ in range(150)]
    for a in vals:
        b = rnd.choice(vals)
        o = O(32); L.emu_sc_mul(B(be(a)), B(be(b)), o); assert int.from_bytes(o, "big") == a * b % N
        o = O(32); L.emu_sc_add(B(be(a)), B(be(b)), o); assert int.from_bytes(o, "big") == (a + b) % N
        o = O(32); L.emu_sc_sub(B(be(a)), B(be(b)), o); assert int.from_bytes(o, "big") == (a - b) % N
        o = O(32); L.emu_sc_neg(B(be(a)), o); assert int.from_bytes(o, "big") == (-a) % N
    for a in vals[1:25]:
        o = O(32); L.emu_sc_inv(B(be(a)), o); assert int.from_bytes(o, "big") == pow(a, -1, N)
    for w in [b"\xff" * 64, b"\0" * 64, N.to_bytes(64, "big"), (N * N - 1).to_bytes(64, "big")] + [rnd.randbytes(64) for _ in range(100)]:
        o = O(32); L.emu_sc_wide(B(w), o); assert int.from_bytes(o, "big") == int.from_bytes(w, "big") % N
    assert L.emu_sc_from_repr(B(be(N))) == 0 and L.emu_sc_from_repr(B(be(N - 1))) == 1


def test_group_law_complete_formulas(emu_prims, ref):
    L, N = emu_prims, ref.N
    rnd = random.Random(9)
    unxy = lambda b: None if b == b"\0" * 64 else (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))  # noqa: E731
    pts = [ref.pt_mul(ref.G, rnd.randrange(1, N)) for _ in range(6)]
    cases = []
    for p in pts:
        cases += [(p, rnd.choice(pts)), (p, p), (p, ref.pt_neg(p)), (p, None), (None, p), (None, None)]
    for p, q in cases:   # identity, P+P and P+(-P) all go through the same branch-free code
        o = O(64); assert L.emu_pt_add(B(xy(p)), B(xy(q)), o) == 0; assert unxy(bytes(o)) == ref.pt_add(p, q)
        if q is not None:
            o = O(64); assert L.emu_pt_add_mixed(B(xy(p)), B(xy(q)), o) == 0; assert unxy(bytes(o)) == ref.pt_add(p, q)
            o = O(64); assert L.emu_pt_add_mixed_proj(B(xy(p)), B(xy(q)), o) == 0
            assert unxy(bytes(o)) == ref.pt_add(ref.pt_mul(p, 3), q)
        o = O(64); assert L.emu_pt_double(B(xy(p)), o) == 0; assert unxy(bytes(o)) == ref.pt_add(p, p)
    for p in pts[:3] + [None]:
        for k in [0, 1, 2, 8, 15, 16, N - 1, N - 2, 2**255, int("8" * 64, 16) % N, rnd.randrange(N)]:
            o = O(64); assert L.emu_pt_mul(B(xy(p)), B(be(k)), o) == 0; assert unxy(bytes(o)) == ref.pt_mul(p, k)
            assert L.emu_pt_equal(B(xy(p)), B(xy(ref.pt_mul(p, k))), B(be(k))) == 1
            assert L.emu_pt_equal(B(xy(p)), B(xy(ref.pt_add(ref.pt_mul(p, k), ref.G))), B(be(k))) == 0


def test_xyzz_accumulator_including_exceptional_cases(emu_prims, ref):
    L = emu_prims
    rnd = random.Random(21)
    unxy = lambda b: None if b == b"\0" * 64 else (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))  # noqa: E731
    pts = [ref.pt_mul(ref.G, rnd.randrange(1, ref.N)) for _ in range(8)]
    p, q = pts[0], pts[1]
    seqs = [pts, [p], [p, p], [p, p, p], [p, ref.pt_neg(p)], [p, ref.pt_neg(p), q], [p, q, ref.pt_neg(ref.pt_add(p, q))],
            [p, q, ref.pt_add(p, q)], [p, p, ref.pt_neg(p), ref.pt_neg(p), q], pts + [ref.pt_neg(x) for x in pts]]
    for seq in seqs:
        exp = None
        for x in seq:
            exp = ref.pt_add(exp, x)
        o = O(64)
        assert L.emu_ptx_sum(B(b"".join(xy(x) for x in seq)), len(seq), o) == 0
        assert unxy(bytes(o)) == exp


def test_jacobian_accumulator_including_exceptional_cases(emu_prims, ref):
    L, N = emu_prims, ref.N
    rnd = random.Random(22)
    unxy = lambda b: None if b == b"\0" * 64 else (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))  # noqa: E731
    p = ref.pt_mul(ref.G, rnd.randrange(1, N))
    q = ref.pt_mul(ref.G, rnd.randrange(1, N))
    for k in [0, 1, 2, 3, 5, N - 1, N - 2, 2**255, rnd.randrange(N), rnd.randrange(N)]:
        kp = ref.pt_mul(p, k)
        extras = [[], [q], [p], [ref.pt_neg(p)], [q, ref.pt_neg(q)], [q, q, q]]
        if kp is not None:
            extras += [[kp], [ref.pt_neg(kp)], [ref.pt_neg(kp), kp, kp]]     # acc == Q (doubling) and acc == -Q (identity)
        for ex in extras:
            exp = kp
            for e in ex:
                exp = ref.pt_add(exp, e)
            o = O(64)
            assert L.emu_ptj_ladder(B(xy(p)), B(be(k)), B(b"".join(xy(e) for e in ex)), len(ex), o) == 0
            assert unxy(bytes(o)) == exp, (hex(k), len(ex))


def test_glv_split_and_multiplication(emu_prims, ref):
    L, N = emu_prims, ref.N
    lam = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
    rnd = random.Random(13)
    unxy = lambda b: None if b == b"\0" * 64 else (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))  # noqa: E731
    ks = [0, 1, 2, N - 1, N - 2, N // 2, N // 2 + 1, lam, N - lam, 2**255 % N, 2**128, 2**128 - 1] + [rnd.randrange(N) for _ in range(300)]
    for k in ks:
        o = O(34); L.emu_glv_split(B(be(k)), o)
        k1, k2 = int.from_bytes(bytes(o[:16]), "big"), int.from_bytes(bytes(o[16:32]), "big")
        if o[32]: k1 = -k1
        if o[33]: k2 = -k2
        assert (k1 + k2 * lam - k) % N == 0 and abs(k1) < 2**128 and abs(k2) < 2**128
    p = ref.pt_mul(ref.G, rnd.randrange(1, N))
    for q in [p, None]:
        for k in ks[:40]:
            o = O(64); assert L.emu_pt_mul_glv(B(xy(q)), B(be(k)), o) == 0
            assert unxy(bytes(o)) == ref.pt_mul(q, k), hex(k)


def test_point_encodings(emu_prims, ref):
    L = emu_prims
    rnd = random.Random(10)
    for _ in range(8):
        p = ref.pt_mul(ref.G, rnd.randrange(1, ref.N))
        c = ref.pt_to_bytes(p)
        o = O(64); assert L.emu_pt_decompress(B(c), o) == 0 and bytes(o) == xy(p)
        o = O(33); assert L.emu_pt_compress(B(xy(p)), o) == 0 and bytes(o) == c
    o = O(64); assert L.emu_pt_decompress(B(b"\0" * 33), o) == 1          # identity
    o = O(33); assert L.emu_pt_compress(B(b"\0" * 64), o) == 1 and bytes(o) == b"\0" * 33
    x = 5
    while True:
        try:
            ref.pt_from_bytes(b"\x02" + be(x)); x += 1
        except ValueError:
            break
    assert L.emu_pt_decompress(B(b"\x02" + be(x)), O(64)) == -1           # x not on the curve
    assert L.emu_pt_decompress(B(b"\x04" + be(ref.GX)), O(64)) == -1        # bad tag
    assert L.emu_pt_decompress(B(b"\x02" + be(ref.P)), O(64)) == -1         # x >= p
    bad = bytearray(xy(ref.G)); bad[63] ^= 1
    assert L.emu_pt_compress(B(bytes(bad)), O(33)) == -1                    # off-curve affine input


def test_keccak_and_merlin(emu_prims, ref):
    L = emu_prims
    rnd = random.Random(12)
    lanes = [rnd.randrange(2**64) for _ in range(25)]
    arr = (C.c_uint64 * 25)(*lanes); L.emu_keccak(arr); assert list(arr) == ref.keccak_f1600(lanes)
    o = O(32); L.emu_merlin_simple(B(b"test protocol"), 13, b"some label", 10, B(b"some data"), 9, b"challenge", 9, o, 32)
    assert bytes(o).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    msg = rnd.randbytes(33)
    o = O(32 * 20); L.emu_merlin_long(B(LABEL), len(LABEL), B(msg), 33, 20, o)
    t = ref.Transcript(LABEL); exp = b""
    for i in range(20):   # crosses the 166-byte STROBE rate several times
        t.append_message(b"wnla_com", msg); t.append_u64(b"l.sz", 32 >> (i & 3)); exp += t.challenge_bytes(b"wnla_challenge", 32)
    assert bytes(o) == exp


def _emu_ctx(emu_u64, gens64, W=4):
    return C.c_void_p(emu_u64.emu_ctx_create(B(gens64), W))


def test_signed_window_digits_on_adversarial_scalars(emu_u64, ref, gens64):
    """Signed fixed-base windows: scalars made of windows equal to 2^(W-1) exactly (the carry has to look further down),
    all-ones, n - 1 and small values, through the commit path x*g + s*h_0 against plain integer arithmetic."""
    for W in (-4, -6):
        w = -W
        H = 1 << (w - 1)
        chain = sum(H << (w * k) for k in range(256 // w - 1))
        blinds = [chain, chain + 1, chain - 1, (chain << w) % ref.N, ref.N - 1, ref.N - 2, 1, 0, H, H + 1, (1 << 255) + chain % (1 << 200), 2**256 % ref.N]
        blinds = [b % ref.N for b in blinds]
        xs = [0, 1, 2**64 - 1, H, H - 1, H + 1, 5, 6, 7, 8, 9, 10]
        n = len(xs)
        ctx = _emu_ctx(emu_u64, gens64, W=W)
        out = (C.c_uint8 * (33 * n))()
        xa = (C.c_uint64 * n)(*xs)
        emu_u64.emu_u64_commit_batch(ctx, C.c_size_t(n), xa, B(b"".join(be(b) for b in blinds)), out)
        g = (int.from_bytes(gens64[:32], "big"), int.from_bytes(gens64[32:64], "big"))
        h0 = (int.from_bytes(gens64[64 * 17:64 * 17 + 32], "big"), int.from_bytes(gens64[64 * 17 + 32:64 * 18], "big"))
        for i in range(n):
            expect = ref.pt_add(ref.pt_mul(g, xs[i]), ref.pt_mul(h0, blinds[i]))
            assert bytes(out)[33 * i:33 * i + 33] == ref.pt_to_bytes(expect), (W, i)
        emu_u64.emu_ctx_destroy(ctx)


def test_device_verify_logic_matches_oracle(emu_u64, ref, oracle, gens64):
    n = 6
    xs, blinds, rngs = synth_batch(ref, n)
    proofs, st = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, 4)
    commits = b"".join(oracle.u64_commit(gens64, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n))
    ctx = _emu_ctx(emu_u64, gens64)
    status = (C.c_int32 * n)()
    emu_u64.emu_u64_verify_batch(ctx, C.c_size_t(n), B(commits), B(proofs), 0, B(LABEL), len(LABEL), status)
    assert list(status) == [1] * n
    bad = bytearray(proofs)
    for i, pos in enumerate([400, 5, 470, 500, 140, 0]):
        bad[525 * i + pos] ^= 1
    emu_u64.emu_u64_verify_batch(ctx, C.c_size_t(n), B(commits), B(bytes(bad)), 0, B(LABEL), len(LABEL), status)
    assert list(status) == oracle.u64_verify_batch(gens64, commits, bytes(bad), LABEL, 4)
    # 64-byte affine input format gives the same verdicts
    def to_affine_rec(rec):
        pts = [oracle.point_decompress(rec[33 * k:33 * k + 33]) for k in range(12)]
        return b"".join(pts) + rec[396:492] + oracle.point_decompress(rec[492:525])
    aff = b"".join(to_affine_rec(proofs[525 * i:525 * i + 525]) for i in range(n))
    acom = b"".join(oracle.point_decompress(commits[33 * i:33 * i + 33]) for i in range(n))
    emu_u64.emu_u64_verify_batch(ctx, C.c_size_t(n), B(acom), B(aff), 1, B(LABEL), len(LABEL), status)
    assert list(status) == [1] * n
    emu_u64.emu_ctx_destroy(ctx)


def test_affine_ladder_tables_equal_the_projective_construction(emu_u64, ref, oracle, gens64):
    """u64_verify.cuh:tables_affine_level (affine chords / tangents with cross-proof Montgomery inversions) against
    u64v_table_build_one + tables_normalize_strided (complete projective formulas, then normalisation): the 104 finished
    entries (x, y, beta x) of every proof must agree word for word -- honest proofs, identity points (33 zero bytes) in
    several slots, undecodable points, and X_j == R_j."""
    n = 7
    xs, blinds, rngs = synth_batch(ref, n)
    proofs, st = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, 4)
    commits = bytearray(b"".join(oracle.u64_commit(gens64, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n)))
    rec = bytearray(proofs)
    rec[525 * 1 + 33 * 2:525 * 1 + 33 * 3] = bytes(33)                    # c_o = identity
    rec[525 * 2 + 33 * 4:525 * 2 + 33 * 5] = bytes(33)                    # r[0] = identity
    rec[525 * 2 + 33 * 11:525 * 2 + 33 * 12] = bytes(33)                  # x[3] = identity
    rec[525 * 3 + 33 * 8:525 * 3 + 33 * 9] = rec[525 * 3 + 33 * 4:525 * 3 + 33 * 5]      # x[0] = r[0]
    rec[525 * 4 + 1:525 * 4 + 33] = (5).to_bytes(32, "big")               # c_l: x = 5 is not on the curve
    commits[33 * 5:33 * 6] = bytes(33)                                    # V = identity
    rec[525 * 6 + 492:525 * 6 + 525] = commits[33 * 6:33 * 6 + 1].replace(b"\x02", b"\x13").replace(b"\x03", b"\x02").replace(b"\x13", b"\x03") + commits[33 * 6 + 1:33 * 7]   # r = -V: V' = identity
    words = n * 104 * 24
    out = [(C.c_uint32 * words)(), (C.c_uint32 * words)()]
    for mode in (0, 1):
        emu_u64.emu_u64_verify_tables(C.c_size_t(n), B(bytes(commits)), B(bytes(rec)), 0, mode, out[mode])
    a, b = list(out[0]), list(out[1])
    for i in range(n):
        for e in range(104):
            lo = (i * 104 + e) * 24
            assert a[lo:lo + 24] == b[lo:lo + 24], (i, e)
    assert any(a), "tables are empty"
    # V' = identity really is the sentinel in both (slot 12 of proof 6), and the honest proof 0 has none
    assert a[(6 * 104 + 12 * 8) * 24:(6 * 104 + 13 * 8) * 24] == [0] * (8 * 24)
    assert all(any(a[(0 * 104 + e) * 24:(0 * 104 + e) * 24 + 16]) for e in range(104))


@pytest.mark.parametrize("nseg", [1, 3, 4, 7, 33])
def test_segmented_ladders_match_the_oracle(emu_u64, ref, oracle, gens64, nseg):
    """u64v_var5_seg / u64v_var2_seg: the ladders cut into nseg runs of windows with the accumulator and the GLV halves parked
    in the scratch rows between runs (what k_v_var_seg does on the GPU) -- verdicts on honest and tampered records as the oracle's."""
    n = 5
    xs, blinds, rngs = synth_batch(ref, n, start=30)
    proofs, st = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, 4)
    commits = b"".join(oracle.u64_commit(gens64, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n))
    bad = bytearray(proofs)
    bad[525 * 1 + 33 * 6:525 * 1 + 33 * 7] = bytes(33)                                   # r[2] = identity
    bad[525 * 2 + 33 * 8:525 * 2 + 33 * 9] = bad[525 * 2 + 33 * 4:525 * 2 + 33 * 5]      # x[0] = r[0]
    bad[525 * 3 + 430] ^= 2                                                               # l[1] tampered
    ctx = _emu_ctx(emu_u64, gens64)
    status = (C.c_int32 * n)()
    emu_u64.emu_set_var_segments(nseg)
    try:
        emu_u64.emu_u64_verify_batch(ctx, C.c_size_t(n), B(commits), B(bytes(bad)), 0, B(LABEL), len(LABEL), status)
    finally:
        emu_u64.emu_set_var_segments(0)
    assert list(status) == oracle.u64_verify_batch(gens64, commits, bytes(bad), LABEL, 4)
    assert status[0] == 1 and status[4] == 1 and status[3] == 0
    emu_u64.emu_ctx_destroy(ctx)


@pytest.mark.parametrize("tab_affine", [0, 1])
def test_device_verify_logic_with_either_table_construction(emu_u64, ref, oracle, gens64, tab_affine):
    n = 4
    xs, blinds, rngs = synth_batch(ref, n, start=9)
    proofs, st = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, 4)
    commits = b"".join(oracle.u64_commit(gens64, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n))
    bad = bytearray(proofs)
    bad[525 * 1 + 33 * 5:525 * 1 + 33 * 6] = bytes(33)      # r[1] = identity
    bad[525 * 2 + 33 * 8 + 7] ^= 4                            # x[0] tampered (most likely undecodable or another point)
    ctx = _emu_ctx(emu_u64, gens64)
    status = (C.c_int32 * n)()
    emu_u64.emu_set_tab_affine(tab_affine)
    try:
        emu_u64.emu_u64_verify_batch(ctx, C.c_size_t(n), B(commits), B(bytes(bad)), 0, B(LABEL), len(LABEL), status)
    finally:
        emu_u64.emu_set_tab_affine(1)
    assert list(status) == oracle.u64_verify_batch(gens64, commits, bytes(bad), LABEL, 4)
    assert status[0] == 1 and status[3] == 1 and status[1] <= 0
    emu_u64.emu_ctx_destroy(ctx)


@pytest.mark.parametrize("W", [5, -5, -7])     # 5: a width that does not divide 32; negative: signed windows (ws.cuh:FixedTable)
def test_device_prove_logic_matches_oracle(emu_u64, ref, oracle, gens64, W):
    n = 5
    xs, blinds, rngs = synth_batch(ref, n)   # includes x = 0, 1, 2^64 - 1
    proofs, st = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, 4)
    ctx = _emu_ctx(emu_u64, gens64, W=W)
    out = (C.c_uint8 * (525 * n))(); status = (C.c_int32 * n)()
    xa = (C.c_uint64 * n)(*xs)
    emu_u64.emu_u64_prove_batch(ctx, C.c_size_t(n), xa, B(blinds), B(rngs), B(LABEL), len(LABEL), out, status)
    assert list(status) == [1] * n
    assert bytes(out) == proofs
    emu_u64.emu_ctx_destroy(ctx)


def test_device_prove_matches_golden(emu_u64, ref, golden, gens64):
    c = golden["cases"][0]
    ctx = _emu_ctx(emu_u64, gens64)
    out = (C.c_uint8 * 525)(); status = (C.c_int32 * 1)()
    xa = (C.c_uint64 * 1)(c["x"])
    emu_u64.emu_u64_prove_batch(ctx, C.c_size_t(1), xa, B(bytes.fromhex(c["blind"])), B(ref.synth_rng_bytes(c["rng_index"])), B(LABEL), len(LABEL), out, status)
    assert bytes(out).hex() == c["proof"]
    emu_u64.emu_ctx_destroy(ctx)
