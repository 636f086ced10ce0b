"""include/bppp.hpp (the C++ host layer that mirrors the reference's Rust items) compiles against libbppp.so, refuses to
run without a GPU, and on a GPU produces the oracle's bytes."""
import os
import struct
import subprocess

import pytest

from conftest import ROOT, has_cuda, synth_batch

LABEL = b"u64 range proof"


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from bp_pp_b200._lib import SO_PATH
    out = str(tmp_path_factory.mktemp("cpp") / "hpp_roundtrip")
    libdir = os.path.dirname(SO_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "hpp_roundtrip.cpp"),
           "-o", out, "-L", libdir, "-l:" + os.path.basename(SO_PATH), "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def _stdin(ref, gens64):
    xs, blinds, rngs = synth_batch(ref, 4)
    i = 3
    return xs[i], blinds[32 * i:32 * i + 32], rngs[3328 * i:3328 * i + 3328]


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_cpp_layer_fails_loudly_without_a_gpu(exe, ref, gens64):
    x, blind, rng = _stdin(ref, gens64)
    r = subprocess.run([exe], input=gens64 + struct.pack("<Q", x) + blind + rng, capture_output=True, timeout=120)
    assert r.returncode == 4
    assert b"no CPU fallback" in r.stdout and b"commit=" not in r.stdout


@pytest.mark.gpu
def test_cpp_layer_matches_the_oracle(exe, ref, oracle, gens64):
    x, blind, rng = _stdin(ref, gens64)
    r = subprocess.run([exe], input=gens64 + struct.pack("<Q", x) + blind + rng, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = dict(line.split("=", 1) for line in r.stdout.decode().splitlines())
    proofs, st = oracle.u64_prove_batch(gens64, [x], blind, rng, LABEL, 1)
    assert st == [0]
    assert kv["commit"] == oracle.u64_commit(gens64, x, blind).hex()
    assert kv["proof"] == proofs.hex()
    assert kv["verify"] == "1" and kv["verify_tampered"] == "0" and kv["malformed"] == "-3"
    assert kv["wnla_rounds"] == "1" and kv["wnla_verify"] == "1"      # 4 + 4 -> 2 + 2 < 6 stops the recursion (wnla.rs:128)
    sc = lambda v: v.to_bytes(32, "big")
    com = oracle.wnla_commit(gens64[:64], gens64[64:64 + 4 * 64], gens64[17 * 64:21 * 64], b"".join(map(sc, [1, 2, 4, 2])), sc(2), sc(4),
                             b"".join(map(sc, [2, 1, 4, 1])), b"".join(map(sc, [1, 4, 2, 2])))
    assert kv["wnla_commit"] == com.hex()
    # ReciprocalRangeProofProtocol mirror at the u64 dimensions: the fast path's record (and so the oracle's) plus the blinded pole commitment r
    assert kv["reciprocal_proof"] == proofs.hex() and kv["reciprocal_commit"] == kv["commit"] == kv["reciprocal_commit_value"]
    assert kv["reciprocal_shape"] == "4,2,1" and kv["reciprocal_verify"] == "1"
    # ArithmeticCircuit mirror: the reference's ac_works instance against the C oracle
    N = ref.N
    be = lambda v: (v % N).to_bytes(32, "big")  # noqa: E731
    P = lambda k: gens64[64 * k:64 * (k + 1)]  # noqa: E731
    g, g_vec, h_vec = P(0), [P(1 + k) for k in range(16)], [P(17 + k) for k in range(32)]
    desc = oracle.make_circuit_desc(1, 2, 1, 2, True, False, g, g_vec[0], b"".join(h_vec[:11]), b"", b"".join(h_vec[11:16]),
                                    b"".join(map(be, [0, 0, 1, 0])), b"".join(map(be, [0, 1, 0, 0, 0, N - 1, 1, 0])), be(0), be(-8) + be(-15),
                                    [-1, -1], [0, 1], [-1, -1], [-1, -1])
    ccom = oracle.circuit_commit(desc, be(3) + be(5), blind)
    rec, rounds, ll, nl = oracle.circuit_prove(desc, ccom, be(3) + be(5), blind, be(3), be(5), be(15) + be(8), rng[:(18 + 2 + 1) * 64], b"circuit test")
    assert kv["circuit_commit"] == ccom.hex() and kv["circuit_proof"] == rec.hex() and kv["circuit_shape"] == f"{rounds},{ll},{nl}"
    assert kv["circuit_verify"] == "1" and kv["circuit_verify_tampered"] == "0"
