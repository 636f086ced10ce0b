"""Generates tests/golden/*.json from the pure-Python oracle (oracle/bppp_ref.py).

The reference holds no golden vectors of its own (all of its tests draw from OsRng, src/tests.rs:14,85,142)
and cannot be run in the build container (no cargo), so these are SELF-FROZEN: they pin the oracle, the C
oracle and the CUDA path to each other, not to k256 ("parity unpinned", see oracle/bppp_ref.py).
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bppp_ref as R  # noqa: E402


def xy(p):
    return (b"\0" * 64 if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")).hex()


def main():
    g, gv, hv = R.synth_generators()
    pub = R.U64RangeProofProtocol(g, gv, hv)
    label = b"u64 range proof"
    cases = []
    values = [123456, 0, 1, 2**64 - 1, 0x0123456789ABCDEF]
    for i, x in enumerate(values):
        s = R.synth_blind(100 + i)
        rng = R.synth_rng_bytes(100 + i)
        proof = pub.prove(x, s, R.Transcript(label), R.ByteRng(rng))
        V = pub.commit_value(x, s)
        assert pub.verify(V, proof, R.Transcript(label))
        rec = R.serialize_reciprocal_proof(proof)
        cases.append({"x": x, "blind": R.sc_to_bytes(s).hex(), "rng_index": 100 + i,
                      "rng_sha256": hashlib.sha256(rng).hexdigest(), "commitment": R.pt_to_bytes(V).hex(),
                      "proof": rec.hex(), "json": R.reciprocal_proof_to_json_obj(proof)})
        print("case", i, x, hashlib.sha256(rec).hexdigest()[:16])
    # tamper verdicts for case 0 (every byte class)
    rec0 = bytes.fromhex(cases[0]["proof"])
    V0 = R.pt_from_bytes(bytes.fromhex(cases[0]["commitment"]))
    tampers = []
    for pos, bit in [(1, 0), (34, 3), (70, 1), (100, 7), (140, 2), (270, 5), (400, 0), (431, 6), (470, 4), (500, 1), (0, 0), (396, 7)]:
        bad = bytearray(rec0)
        bad[pos] ^= 1 << bit
        try:
            pr = R.deserialize_u64_proof(bytes(bad))
            try:
                verdict = 1 if pub.verify(V0, pr, R.Transcript(label)) else 0
            except (ZeroDivisionError,):
                verdict = -1
        except ValueError as e:
            verdict = -4 if "from_repr" in str(e) else -3
        tampers.append({"pos": pos, "bit": bit, "verdict": verdict})
        print("tamper", pos, bit, verdict)
    other = {"wrong_commitment": 1 if pub.verify(R.pt_add(V0, R.G), R.deserialize_u64_proof(rec0), R.Transcript(label)) else 0,
             "wrong_label": 1 if pub.verify(V0, R.deserialize_u64_proof(rec0), R.Transcript(b"u64 range prooF")) else 0}
    out = {"label": label.decode(), "generators": [xy(p) for p in [g] + gv + hv], "cases": cases, "tampers_case0": tampers,
           "other_case0": other}
    with open(os.path.join(HERE, "u64_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    # Merlin / Keccak known answers
    t = R.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    kat = {"merlin_simple": {"proto": "test protocol", "label": "some label", "data": "some data", "challenge_label": "challenge",
                             "challenge32": t.challenge_bytes(b"challenge", 32).hex(),
                             "note": "merlin's published conformance vector (d5a21972...0615); one nibble differs from the survey "
                                     "author's memory of it (9bca computed vs 9bfa recalled) -- a wrong construction would differ in ~all nibbles"},
           "keccak_sha3_256_empty": hashlib.sha3_256(b"").hexdigest()}
    with open(os.path.join(HERE, "merlin_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    # standalone WNLA (tests.rs:138-171 shape, N = 4) and the ac_works circuit (tests.rs:44-136), seeded
    import struct
    gs = [R.pt_mul(R.G, int.from_bytes(R.S("wnla-gen", j, 64), "big") % R.N) for j in range(9)]
    wg, wgv, whv = gs[0], gs[1:5], gs[5:9]
    c = [int.from_bytes(R.S("wnla-c", j, 64), "big") % R.N for j in range(4)]
    rho = int.from_bytes(R.S("wnla-rho", 0, 64), "big") % R.N
    w = R.WeightNormLinearArgument(wg, wgv, whv, c, rho, rho * rho % R.N)
    l, n = [1, 2, 3, 4], [8, 7, 6, 5]
    com = w.commit(l, n)
    pr = w.prove(com, R.Transcript(b"wnla test"), l, n)
    assert w.verify(com, R.Transcript(b"wnla test"), R.WnlaProof(pr.r, pr.x, pr.l, pr.n))
    wn = {"g": xy(wg), "g_vec": [xy(p) for p in wgv], "h_vec": [xy(p) for p in whv], "c": [R.sc_to_bytes(v).hex() for v in c],
          "rho": R.sc_to_bytes(rho).hex(), "l": l, "n": n, "commitment": R.pt_to_bytes(com).hex(),
          "proof": R.serialize_wnla_proof(pr).hex(), "rounds": len(pr.r), "l_len": len(pr.l), "n_len": len(pr.n)}
    with open(os.path.join(HERE, "wnla_golden.json"), "w") as f:
        json.dump(wn, f, indent=1)
    print("wnla rounds", len(pr.r), len(pr.l), len(pr.n))


if __name__ == "__main__":
    main()
