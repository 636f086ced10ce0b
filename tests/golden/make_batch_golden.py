"""Generates tests/golden/u64_batch_golden.json: the FULL BASELINE config 2/3 batch (65,536 proofs, SURVEY 8d inputs
S(tag, i)) run through the C oracle (oracle/oracle.c, all host threads), stored as sha256 per 4,096-proof block of
commitments, honest proofs and tampered records, plus the oracle's verdict for every tampered record.

The -m gpu suite and bench.py compare the CUDA path against these at full size and at every GPU count without needing
the oracle (or minutes of CPU time) at run time.  SELF-FROZEN like the other goldens: it pins the CUDA path to the
oracle, not to k256 ("parity unpinned").
Run:  python tests/golden/make_batch_golden.py          (~2-3 minutes on 8 cores)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bppp_ref as R  # noqa: E402
import oracle_c as OC  # noqa: E402
from bp_pp_b200 import synth  # noqa: E402   (seeded inputs and the tamper rules only: no engine call is made here)

N_BATCH = 65536
LABEL = b"u64 range proof"


def xy(p):
    return p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def main():
    OC.build(); OC.use_native()
    threads = os.cpu_count() or 1
    g, gv, hv = R.synth_generators()
    gens = b"".join(xy(p) for p in [g] + gv + hv)
    xs, blinds, rng = synth.synth_batch(N_BATCH)
    assert int(xs[5]) == R.synth_x(5) and blinds[7].tobytes() == R.sc_to_bytes(R.synth_blind(7)) and rng[9].tobytes() == R.synth_rng_bytes(9)
    t0 = time.time()
    proofs, st = OC.u64_prove_batch(gens, xs.tolist(), blinds.tobytes(), rng.tobytes(), LABEL, threads)
    assert st == [0] * N_BATCH
    print(f"proved in {time.time() - t0:.1f} s")
    commits = b"".join(OC.u64_commit(gens, int(xs[i]), blinds[i].tobytes()) for i in range(N_BATCH))
    G64 = xy(R.G)

    def add_g(pts):
        return [OC.point_compress(OC.point_add(OC.point_decompress(p), G64)) for p in pts]

    bad, bcom, idx = synth.tamper_batch(proofs, commits, add_g)
    sub_p = b"".join(bad[525 * i:525 * i + 525] for i in idx)
    sub_c = b"".join(bcom[33 * i:33 * i + 33] for i in idx)
    t0 = time.time()
    verdicts = OC.u64_verify_batch(gens, sub_c, sub_p, LABEL, threads)
    print(f"verified {len(idx)} tampered records in {time.time() - t0:.1f} s:", {v: verdicts.count(v) for v in sorted(set(verdicts))})
    # the honest records verify (spot check; the full set is what the GPU arm proves and verifies)
    assert OC.u64_verify_batch(gens, commits[:33 * 64], proofs[:525 * 64], LABEL, threads) == [1] * 64
    out = {
        "n": N_BATCH, "label": LABEL.decode(), "block": 4096, "tamper_every": synth.TAMPER_EVERY,
        "generators_sha256": __import__("hashlib").sha256(gens).hexdigest(),
        "commit_block_sha256": synth.block_hashes(commits, 33),
        "proof_block_sha256": synth.block_hashes(proofs, 525),
        "tampered_proof_block_sha256": synth.block_hashes(bad, 525),
        "tampered_commit_block_sha256": synth.block_hashes(bcom, 33),
        "tampered_verdicts": verdicts,
    }
    with open(os.path.join(HERE, "u64_batch_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote u64_batch_golden.json")


if __name__ == "__main__":
    main()
