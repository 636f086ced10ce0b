"""GPU suite (`-m gpu`): the CUDA path through the C ABI against the oracle and the golden fixtures.
Bit-exact is the bar: identical 525-byte proofs, identical commitments, identical verdicts (tampered too)."""
import hashlib
import os
import random

import pytest

from conftest import synth_batch, xy

pytestmark = pytest.mark.gpu
LABEL = b"u64 range proof"
THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def ctx(gens64):
    import bp_pp_b200 as B
    c = B.Context(gens64, 0, 16, 65536)
    yield c
    c.close()


@pytest.fixture(scope="module")
def batch(ref, oracle, gens64):
    n = 384
    xs, blinds, rngs = synth_batch(ref, n)      # x_0..2 = 0, 1, 2^64-1
    proofs, st = oracle.u64_prove_batch(gens64, xs, blinds, rngs, LABEL, THREADS)
    assert st == [0] * n
    commits = b"".join(oracle.u64_commit(gens64, xs[i], blinds[32 * i:32 * i + 32]) for i in range(n))
    return {"n": n, "xs": xs, "blinds": blinds, "rngs": rngs, "proofs": proofs, "commits": commits}


def test_native_library_is_the_one_loaded(ctx):
    from bp_pp_b200._lib import SO_PATH
    maps = open("/proc/self/maps").read()
    assert os.path.basename(SO_PATH) in maps
    assert ctx.info()["table_bytes"] == 49 * 16 * 65535 * 64


def test_commit_parity(ctx, batch):
    import bp_pp_b200 as B
    assert ctx.commit_batch(batch["xs"], batch["blinds"]) == batch["commits"]
    aff = ctx.commit_batch(batch["xs"][:8], batch["blinds"][:256], B.FMT_AFFINE64)
    import oracle_c
    assert aff == b"".join(oracle_c.point_decompress(batch["commits"][33 * i:33 * i + 33]) for i in range(8))


def test_prove_is_byte_identical_to_the_oracle(ctx, batch):
    proofs, st = ctx.prove_batch(batch["xs"], batch["blinds"], batch["rngs"], LABEL)
    assert st == [1] * batch["n"]
    assert proofs == batch["proofs"]


def test_prove_matches_golden_fixtures(ctx, ref, golden):
    cases = golden["cases"]
    xs = [c["x"] for c in cases]
    blinds = b"".join(bytes.fromhex(c["blind"]) for c in cases)
    rngs = b"".join(ref.synth_rng_bytes(c["rng_index"]) for c in cases)
    proofs, st = ctx.prove_batch(xs, blinds, rngs, LABEL)
    assert st == [1] * len(cases)
    for i, c in enumerate(cases):
        assert proofs[525 * i:525 * i + 525].hex() == c["proof"]
    commits = ctx.commit_batch(xs, blinds)
    assert commits.hex() == "".join(c["commitment"] for c in cases)
    assert ctx.verify_batch(commits, proofs, LABEL) == [1] * len(cases)


def test_verify_honest_and_golden_tampers(ctx, golden):
    c = golden["cases"][0]
    rec, com = bytes.fromhex(c["proof"]), bytes.fromhex(c["commitment"])
    recs, exp = [rec], [1]
    for t in golden["tampers_case0"]:
        bad = bytearray(rec); bad[t["pos"]] ^= 1 << t["bit"]
        recs.append(bytes(bad)); exp.append(t["verdict"])
    assert ctx.verify_batch(com * len(recs), b"".join(recs), LABEL) == exp
    assert ctx.verify_batch(com, rec, b"u64 range prooF") == [0]           # wrong transcript label


def _tamper(rec: bytes, com: bytes, rule: int, rnd, oracle):
    """The tamper rules of SURVEY 8d config 2 that keep the canonical record shape."""
    rec = bytearray(rec)
    G = xy((0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
            0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8))
    pt_off = [33 * k for k in range(12)] + [492]
    if rule == 0:      # point += G at one of 13 positions
        o = rnd.choice(pt_off)
        rec[o:o + 33] = oracle.point_compress(oracle.point_add(oracle.point_decompress(bytes(rec[o:o + 33])), G))
    elif rule == 1:    # scalar += 1 at one of 3 positions
        o = 396 + 32 * rnd.randrange(3)
        v = (int.from_bytes(rec[o:o + 32], "big") + 1) % 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
        rec[o:o + 32] = v.to_bytes(32, "big")
    elif rule == 2:    # point := identity
        o = rnd.choice(pt_off); rec[o:o + 33] = b"\0" * 33
    elif rule == 3:    # commitment += G
        com = oracle.point_compress(oracle.point_add(oracle.point_decompress(com), G))
    elif rule == 4:    # swap X and R of one round
        j = rnd.randrange(4); a, b = 132 + 33 * j, 264 + 33 * j
        rec[a:a + 33], rec[b:b + 33] = rec[b:b + 33], rec[a:a + 33]
    elif rule == 5:    # swap two rounds
        a, b = 132, 132 + 33 * 3
        rec[a:a + 33], rec[b:b + 33] = rec[b:b + 33], rec[a:a + 33]
    elif rule == 6:    # random bit flip anywhere (often an invalid encoding)
        o = rnd.randrange(525); rec[o] ^= 1 << rnd.randrange(8)
    elif rule == 7:    # non-canonical scalar / x >= p
        if rnd.random() < 0.5: rec[396:428] = b"\xff" * 32
        else: rec[1:33] = b"\xff" * 32
    return bytes(rec), com


def test_verify_verdicts_match_oracle_on_tampered_batch(ctx, batch, oracle, gens64):
    rnd = random.Random(2024)
    n = batch["n"]
    recs, coms = [], []
    for i in range(n):
        rec, com = batch["proofs"][525 * i:525 * i + 525], batch["commits"][33 * i:33 * i + 33]
        if i % 2 == 1:
            rec, com = _tamper(rec, com, (i // 2) % 8, rnd, oracle)
        recs.append(rec); coms.append(com)
    recs, coms = b"".join(recs), b"".join(coms)
    expect = oracle.u64_verify_batch(gens64, coms, recs, LABEL, THREADS)
    got = ctx.verify_batch(coms, recs, LABEL)
    assert got == expect
    assert expect.count(1) >= n // 2 and expect.count(0) > 0 and any(v < 0 for v in expect)


def test_projective_table_construction_gives_the_same_verdicts(gens64, batch, oracle, monkeypatch):
    """BPPP_TAB_AFFINE=0 (read at context creation) switches the verifier's ladder tables back to the projective build +
    normalisation pass; the 8-rule tampered batch must come out exactly as with the affine levels and as the oracle says.
    Identity points, swapped X / R and undecodable bytes all pass through the tables."""
    import bp_pp_b200 as B
    rnd = random.Random(7)
    n = 192
    recs, coms = [], []
    for i in range(n):
        rec, com = batch["proofs"][525 * i:525 * i + 525], batch["commits"][33 * i:33 * i + 33]
        if i % 3:
            rec, com = _tamper(rec, com, i % 8, rnd, oracle)
        recs.append(rec); coms.append(com)
    recs, coms = b"".join(recs), b"".join(coms)
    expect = oracle.u64_verify_batch(gens64, coms, recs, LABEL, THREADS)
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("BPPP_TAB_AFFINE", mode)
        c = B.Context(gens64, 0, 8, 256)
        try:
            got[mode] = c.verify_batch(coms, recs, LABEL)
        finally:
            c.close()
    assert got["0"] == expect and got["1"] == expect


def test_affine64_input_format(ctx, batch, oracle):
    import bp_pp_b200 as B
    n = 64
    def to_affine_rec(rec):
        return b"".join(oracle.point_decompress(rec[33 * k:33 * k + 33]) for k in range(12)) + rec[396:492] + oracle.point_decompress(rec[492:525])
    aff = b"".join(to_affine_rec(batch["proofs"][525 * i:525 * i + 525]) for i in range(n))
    acom = b"".join(oracle.point_decompress(batch["commits"][33 * i:33 * i + 33]) for i in range(n))
    assert ctx.verify_batch(acom, aff, LABEL, B.FMT_AFFINE64) == [1] * n
    bad = bytearray(aff); bad[928 * 3 + 63] ^= 1          # off-curve y
    bad[928 * 5 + 12 * 64 + 5] ^= 1                       # l scalar
    exp = [1] * n; exp[3] = -3; exp[5] = 0
    assert ctx.verify_batch(acom, bytes(bad), LABEL, B.FMT_AFFINE64) == exp


def test_empty_and_ragged_batches(ctx, batch):
    assert ctx.verify_batch(b"", b"", LABEL) == []
    assert ctx.prove_batch([], b"", b"", LABEL) == (b"", [])
    for n in [1, 31, 33, 65]:
        assert ctx.verify_batch(batch["commits"][:33 * n], batch["proofs"][:525 * n], LABEL) == [1] * n


def test_slicing_when_batch_exceeds_workspace(gens64, batch):
    import bp_pp_b200 as B
    small = B.Context(gens64, 0, 8, 50)      # max_batch 50 < 128: three slices; w = 8 tables
    n = 128
    assert small.verify_batch(batch["commits"][:33 * n], batch["proofs"][:525 * n], LABEL) == [1] * n
    proofs, st = small.prove_batch(batch["xs"][:n], batch["blinds"][:32 * n], batch["rngs"][:3328 * n], LABEL)
    assert proofs == batch["proofs"][:525 * n]
    small.close()


def test_window_width_does_not_change_results(gens64, batch):
    import bp_pp_b200 as B
    for W in (5, 11):
        c = B.Context(gens64, 0, W, 64)
        proofs, _ = c.prove_batch(batch["xs"][:16], batch["blinds"][:512], batch["rngs"][:3328 * 16], LABEL)
        assert proofs == batch["proofs"][:525 * 16]
        assert c.verify_batch(batch["commits"][:33 * 16], proofs, LABEL) == [1] * 16
        c.close()


def test_non_canonical_blinding_is_rejected(ctx):
    proofs, st = ctx.prove_batch([5, 6], b"\xff" * 32 + (7).to_bytes(32, "big"), bytes(2 * 3328), LABEL)
    assert st == [-4, 1]


def test_full_size_round_trip_properties(ctx, ref):
    """BASELINE configs 2/3 at full size (65,536) through size-independent properties: prove -> verify all true,
    determinism, every tampered record rejected, untouched neighbours unaffected."""
    import numpy as np
    n = 65536
    rnd = np.random.default_rng(7)
    xs = rnd.integers(0, 2**64, size=n, dtype=np.uint64).tolist()
    xs[0], xs[1], xs[2] = 0, 1, 2**64 - 1
    blinds = bytearray(rnd.bytes(32 * n))
    for i in range(n):
        blinds[32 * i] &= 0x7F          # < n
    blinds = bytes(blinds)
    rngs = rnd.bytes(3328 * n)
    commits = ctx.commit_batch(xs, blinds)
    proofs, st = ctx.prove_batch(xs, blinds, rngs, LABEL)
    assert st == [1] * n
    proofs2, _ = ctx.prove_batch(xs, blinds, rngs, LABEL)
    assert hashlib.sha256(proofs).digest() == hashlib.sha256(proofs2).digest()
    assert ctx.verify_batch(commits, proofs, LABEL) == [1] * n
    bad = bytearray(proofs)
    idx = list(range(0, n, 16))
    for i in idx:
        bad[525 * i + 396 + (i % 96)] ^= 1          # one bit in l/n scalars
    verdicts = ctx.verify_batch(commits, bytes(bad), LABEL)
    assert all(verdicts[i] <= 0 for i in idx)
    assert sum(1 for v in verdicts if v == 1) == n - len(idx)
    # proofs are not transferable between commitments
    shifted = commits[33:] + commits[:33]
    assert ctx.verify_batch(shifted[:33 * 1024], proofs[:525 * 1024], LABEL).count(1) == 0


def test_device_field_arithmetic_matches_the_host_branches():
    """tests/cuda/fe_selftest.cu: every PTX carry-chain path (field mul / sqr / add / sub / small multiples / normalise,
    scalar mul / sqr / wide reduction / inversion, the three point-formula families) against the portable host branches
    of the same headers on 20,400 operand pairs including non-canonical representatives.  The host branches are
    themselves checked against Python integers by tests/test_hostemu.py."""
    import subprocess
    import __graft_entry__ as G
    exe = G.build_selftest()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fe_selftest: ok" in r.stdout and r.stdout.count(" 0 / ") >= 22


@pytest.mark.parametrize("W", [8, -8, -5])
def test_signed_and_unsigned_window_tables_agree_with_the_oracle(gens64, batch, W):
    """bppp_ctx_create window_bits: positive = unsigned windows, negative = signed windows of |W| bits (the layout a W = 22
    table uses); commit, prove and verify must stay byte-identical."""
    import bp_pp_b200 as B
    n = 96
    c = B.Context(gens64, 0, W, n)
    xs, blinds, rngs = batch["xs"][:n], batch["blinds"][:32 * n], batch["rngs"][:3328 * n]
    assert c.commit_batch(xs, blinds) == batch["commits"][:33 * n]
    proofs, st = c.prove_batch(xs, blinds, rngs, LABEL)
    assert st == [1] * n and proofs == batch["proofs"][:525 * n]
    bad = bytearray(proofs)
    for i in range(0, n, 3):
        bad[525 * i + 396 + (i % 96)] ^= 1 << (i % 8)          # scalars l / n: stays decodable
    import oracle_c
    assert c.verify_batch(batch["commits"][:33 * n], bytes(bad), LABEL) == oracle_c.u64_verify_batch(gens64, batch["commits"][:33 * n], bytes(bad), LABEL, THREADS)
    c.close()


def test_context_creation_failure_is_clean(gens64):
    """A workspace that cannot be allocated fails with an error (no crash, nothing leaked that blocks the next context)."""
    import bp_pp_b200 as B
    with pytest.raises(B.BpppError):
        B.Context(gens64, 0, 8, 1 << 27)            # ~3.7 TB of workspace
    c = B.Context(gens64, 0, 8, 16)
    assert c.info()["window_bits"] == 8
    c.close()


def test_device_resident_entry_points_on_a_side_stream(ctx, batch):
    """bppp_u64_{prove,verify}_batch_dev: operands already in HBM, work enqueued on the caller's (non-default) stream,
    results identical to the host-buffer entry points."""
    import numpy as np
    import torch
    n = batch["n"]
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        d_x = torch.from_numpy(np.array(batch["xs"], dtype=np.uint64).view(np.int64)).to(dev)
        d_bl = torch.frombuffer(bytearray(batch["blinds"]), dtype=torch.uint8).to(dev)
        d_rng = torch.frombuffer(bytearray(batch["rngs"]), dtype=torch.uint8).to(dev)
        d_out = torch.zeros(525 * n, dtype=torch.uint8, device=dev)
        d_pst = torch.zeros(n, dtype=torch.int32, device=dev)
        ctx.prove_batch_dev(n, d_x.data_ptr(), d_bl.data_ptr(), d_rng.data_ptr(), LABEL, d_out.data_ptr(), d_pst.data_ptr(), stream=side.cuda_stream)
        d_com = torch.frombuffer(bytearray(batch["commits"]), dtype=torch.uint8).to(dev)
        d_vst = torch.zeros(n, dtype=torch.int32, device=dev)
        ctx.verify_batch_dev(n, d_com.data_ptr(), d_out.data_ptr(), LABEL, d_vst.data_ptr(), stream=side.cuda_stream)
    side.synchronize()
    assert bytes(d_out.cpu().numpy()) == batch["proofs"]
    assert d_pst.cpu().tolist() == [1] * n and d_vst.cpu().tolist() == [1] * n


def test_commit_rejects_a_non_canonical_blinding(ctx):
    """A k256::Scalar cannot hold a value >= n: commit_batch must fail rather than return an unblinded x*g."""
    import bp_pp_b200 as B
    n_be = (0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141).to_bytes(32, "big")
    with pytest.raises(B.BpppError):
        ctx.commit_batch([5, 6], (7).to_bytes(32, "big") + n_be)
    assert len(ctx.commit_batch([5, 6], (7).to_bytes(32, "big") * 2)) == 66       # the context stays usable


@pytest.mark.parametrize("var_lanes,msm_lanes", [(1, 4), (2, 8), (4, 16)])
def test_lane_variants_are_byte_identical(gens64, batch, oracle, var_lanes, msm_lanes, monkeypatch):
    """Intra-proof parallelism for small (sub-)batches: the ladders with 2 / 4 lanes per proof (GLV halves shared out,
    partial sums met through warp shuffles) and the fixed-base sums with 8 / 16 lanes must give the single-lane bytes."""
    import bp_pp_b200 as B
    monkeypatch.setenv("BPPP_VAR_LANES_RT", str(var_lanes))
    monkeypatch.setenv("BPPP_MSM_LANES_RT", str(msm_lanes))
    n = 200
    c = B.Context(gens64, 0, 8, n)
    xs, blinds, rngs = batch["xs"][:n], batch["blinds"][:32 * n], batch["rngs"][:3328 * n]
    proofs, st = c.prove_batch(xs, blinds, rngs, LABEL)
    assert st == [1] * n and proofs == batch["proofs"][:525 * n]
    rnd = random.Random(99)
    recs, coms = [], []
    for i in range(n):
        rec, com = batch["proofs"][525 * i:525 * i + 525], batch["commits"][33 * i:33 * i + 33]
        if i % 3 == 1:
            rec, com = _tamper(rec, com, (i // 3) % 8, rnd, oracle)
        recs.append(rec); coms.append(com)
    recs, coms = b"".join(recs), b"".join(coms)
    assert c.verify_batch(coms, recs, LABEL) == oracle.u64_verify_batch(gens64, coms, recs, LABEL, THREADS)
    c.close()


def test_small_batches_pick_lanes_automatically(ctx, batch):
    """The default lane choice (by batch size) at sizes on both sides of every threshold the 16-bit context can reach."""
    for n in (1, 7, 64, 384):
        proofs, st = ctx.prove_batch(batch["xs"][:n], batch["blinds"][:32 * n], batch["rngs"][:3328 * n], LABEL)
        assert proofs == batch["proofs"][:525 * n] and st == [1] * n
        assert ctx.verify_batch(batch["commits"][:33 * n], proofs, LABEL) == [1] * n


def test_shared_table_contexts_run_side_by_side(ctx, batch):
    """bppp_ctx_create_shared: contexts sharing one set of window tables, each with its own workspace, driven from
    concurrent host threads: every thread must get the oracle's bytes."""
    import threading
    n = batch["n"]
    others = [ctx.shared(n) for _ in range(3)]
    results = {}

    def work(k, c):
        for _ in range(3):
            proofs, st = c.prove_batch(batch["xs"], batch["blinds"], batch["rngs"], LABEL)
            verdicts = c.verify_batch(batch["commits"], proofs, LABEL)
        results[k] = (proofs, st, verdicts)

    threads = [threading.Thread(target=work, args=(k, c)) for k, c in enumerate([ctx] + others)]
    for t in threads: t.start()
    for t in threads: t.join()
    for k in range(4):
        proofs, st, verdicts = results[k]
        assert proofs == batch["proofs"] and st == [1] * n and verdicts == [1] * n
    for c in others:
        c.close()


def test_full_batch_matches_the_oracle_block_hashes(gens64):
    """BASELINE configs 2/3 at FULL size against the C oracle: all 65,536 commitments and proofs (sha256 per 4,096-proof
    block) and the complete verdict vector of the 8-rule tampered batch, from tests/golden/u64_batch_golden.json
    (generated by tests/golden/make_batch_golden.py with the oracle on all host cores)."""
    import json
    import hashlib
    import bp_pp_b200 as B
    from bp_pp_b200 import synth
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "u64_batch_golden.json")))
    n = gold["n"]
    assert hashlib.sha256(gens64).hexdigest() == gold["generators_sha256"]
    assert synth.synth_generators64(0) == gens64             # the product-side generator synthesis is the oracle's
    ctx = B.Context(gens64, 0, 16, n)
    xs, blinds, rng = synth.synth_batch(n)
    commits = ctx.commit_batch(xs.tolist(), blinds.tobytes())
    assert synth.block_hashes(commits, 33) == gold["commit_block_sha256"]
    proofs, st = ctx.prove_batch(xs.tolist(), blinds.tobytes(), rng.tobytes(), LABEL)
    assert st == [1] * n
    assert synth.block_hashes(proofs, 525) == gold["proof_block_sha256"]
    bad, bcom, idx = synth.tamper_batch(proofs, commits, synth.engine_add_g(0))
    assert synth.block_hashes(bad, 525) == gold["tampered_proof_block_sha256"]
    assert synth.block_hashes(bcom, 33) == gold["tampered_commit_block_sha256"]
    verdicts = ctx.verify_batch(bcom, bad, LABEL)
    expect = [1] * n
    for i, v in zip(idx, gold["tampered_verdicts"]):
        expect[i] = v
    assert verdicts == expect
    ctx.close()


def test_multi_context_in_one_process(gens64, batch):
    """bppp_multi_ctx: one process, a list of devices (here the visible GPUs, or GPU 0 three times when there is only
    one -- three contexts, three host threads, same splitting logic): results must not depend on the device list."""
    import torch
    import bp_pp_b200 as B
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    m = B.MultiContext(gens64, devices, 8, 64)          # max_batch 64 per device: 384 proofs also exercise slicing
    n = batch["n"]
    assert m.commit_batch(batch["xs"], batch["blinds"]) == batch["commits"]
    proofs, st = m.prove_batch(batch["xs"], batch["blinds"], batch["rngs"], LABEL)
    assert proofs == batch["proofs"] and st == [1] * n
    bad = bytearray(proofs)
    for i in range(0, n, 5):
        bad[525 * i + 430] ^= 2
    exp = [0 if i % 5 == 0 else 1 for i in range(n)]
    assert m.verify_batch(batch["commits"], bytes(bad), LABEL) == exp
    assert m.verify_batch(batch["commits"][:33 * 2], proofs[:525 * 2], LABEL) == [1, 1]      # fewer proofs than devices
    m.close()
