import ctypes as C
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def xy(p):
    return b"\0" * 64 if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "u64_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gens64(golden):
    return b"".join(bytes.fromhex(h) for h in golden["generators"])


@pytest.fixture(scope="session")
def oracle():
    import oracle_c
    oracle_c.build()
    return oracle_c


@pytest.fixture(scope="session")
def ref():
    import bppp_ref
    return bppp_ref


def _build_emu(name, extra=()):
    src = os.path.join(ROOT, "tests", "hostemu", name + ".cpp")
    outdir = os.path.join(ROOT, "tests", "_hostemu")
    os.makedirs(outdir, exist_ok=True)
    so = os.path.join(outdir, "lib" + name + ".so")
    deps = [src] + [os.path.join(ROOT, "bp_pp_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "bp_pp_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DBPPP_VERIFY_MAG", *extra, "-x", "c++", src, "-o", so])
    return C.CDLL(so)


@pytest.fixture(scope="session")
def emu_prims():
    """bp_pp_b200/csrc/{fe,sc,ec,merlin}.cuh compiled for the host with the bound assertions of fe.cuh (BPPP_VERIFY_MAG)."""
    return _build_emu("emu_prims")


@pytest.fixture(scope="session")
def emu_u64():
    L = _build_emu("emu_u64", ("-DBPPP_EMU_PROVE",))
    L.emu_ctx_create.restype = C.c_void_p
    return L


def synth_batch(ref, n, start=0):
    xs = [ref.synth_x(start + i) for i in range(n)]
    blinds = b"".join(ref.sc_to_bytes(ref.synth_blind(start + i)) for i in range(n))
    rngs = b"".join(ref.synth_rng_bytes(start + i) for i in range(n))
    return xs, blinds, rngs


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
