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


def synth_circuit(ref, k, nv, nm, no, f_l, f_m, seed):
    """A random dense ArithmeticCircuit with a witness built to satisfy W_l w + f_l v + a_l = 0 and
    w_L o w_R = W_m w + f_m v + a_m (circuit.rs:95-139).  k > 1 and f_m = true are the branches the reference's own tests
    never reach (circuit.rs:559-570,603-611).  Returns python-int structures; callers flatten them for the C ABIs."""
    import random
    N = ref.N
    rnd = random.Random(seed)
    nl, nw = k * nv, 2 * nm + no
    pts = [ref.pt_mul(ref.G, int.from_bytes(ref.S("ac2-gen", j, 64), "big") % N) for j in range(1 + nm + 16)]
    g, g_vec, h_all = pts[0], pts[1:1 + nm], pts[1 + nm:]
    W_m = [[rnd.randrange(N) if rnd.random() < 0.5 else 0 for _ in range(nw)] for _ in range(nm)]
    W_l = [[rnd.randrange(N) if rnd.random() < 0.5 else 0 for _ in range(nw)] for _ in range(nl)]
    wl = [rnd.randrange(N) for _ in range(nm)]
    wr = [rnd.randrange(N) for _ in range(nm)]
    wo = [rnd.randrange(N) for _ in range(no)]
    w = wl + wr + wo
    v = [[rnd.randrange(N) for _ in range(nv)] for _ in range(k)]
    vflat = [x for row in v for x in row]
    dot = lambda row: sum(a * b for a, b in zip(row, w)) % N  # noqa: E731
    a_l = [(-dot(W_l[i]) - (vflat[i] if f_l else 0)) % N for i in range(nl)]
    a_m = [(wl[i] * wr[i] - dot(W_m[i]) - (vflat[i] if (f_m and i < len(vflat)) else 0)) % N for i in range(nm)]
    s_v = [rnd.randrange(N) for _ in range(k)]
    return dict(k=k, nv=nv, nm=nm, no=no, nl=nl, nw=nw, f_l=f_l, f_m=f_m, g=g, g_vec=g_vec, h_vec=h_all[:9 + nv], h_vec_=h_all[9 + nv:],
                W_m=W_m, W_l=W_l, a_m=a_m, a_l=a_l, wl=wl, wr=wr, wo=wo, v=v, s_v=s_v, part_ll=list(range(no)))


def circuit_bytes(ref, c):
    """Flattened byte arguments of a synth_circuit for oracle_c.make_circuit_desc / bp_pp_b200.ArithmeticCircuit."""
    be = lambda val: (val % ref.N).to_bytes(32, "big")  # noqa: E731
    flat = lambda m: b"".join(be(e) for row in m for e in row)  # noqa: E731
    vec = lambda a: b"".join(be(e) for e in a)  # noqa: E731
    pv = lambda ps: b"".join(xy(p) for p in ps)  # noqa: E731
    none = [-1] * c["no"]
    return dict(g=xy(c["g"]), g_vec=pv(c["g_vec"]), h_vec=pv(c["h_vec"]), h_vec_=pv(c["h_vec_"]), W_m=flat(c["W_m"]), W_l=flat(c["W_l"]), a_m=vec(c["a_m"]),
                a_l=vec(c["a_l"]), wl=vec(c["wl"]), wr=vec(c["wr"]), wo=vec(c["wo"]), v=flat(c["v"]), s_v=vec(c["s_v"]), part_lo=none, part_ll=c["part_ll"],
                part_lr=none, part_no=none)
