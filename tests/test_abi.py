"""CPU suite: the C-ABI library loads and exports every symbol include/bppp.h declares; no GPU => loud failure."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_cuda


def _header_functions():
    text = open(os.path.join(ROOT, "include", "bppp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bppp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bp_pp_b200._lib import EXPORTS, lib
    L = lib()
    declared = _header_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), f"libbppp.so does not export {name}"
    assert sorted(EXPORTS) == declared


def test_signatures_use_plain_c_types_only():
    text = open(os.path.join(ROOT, "include", "bppp.h")).read()
    assert "torch" not in text and "at::" not in text and "std::" not in text
    assert 'extern "C"' in text


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    import bp_pp_b200 as B
    with pytest.raises(B.BpppError) as ei:
        B.Context(b"\0" * 64 * 49)
    assert "-11" in str(ei.value) and "no CPU fallback" in str(ei.value)
    with pytest.raises(B.BpppError):
        B.microbench(0)
    from bp_pp_b200.shard import PeerGroup
    with pytest.raises(B.BpppError) as ei:
        PeerGroup(0)                                  # the exchange kernels need a device too: no host-side stand-in
    assert "-11" in str(ei.value)


def test_argument_validation_in_the_binding():
    import bp_pp_b200 as B
    with pytest.raises(ValueError):
        B.Context(b"\0" * 10)
    assert B.U64RangeProofProtocol.u64_to_hex(0x1234) == [4, 3, 2, 1] + [0] * 12
    assert B.U64RangeProofProtocol.u64_to_hex_mapped(0x1123)[:4] == [12, 2, 1, 1]


def test_compress64_is_a_pure_re_encoding(golden):
    """shard.compress64 turns 64-byte affine coordinates into the 33-byte SEC1 form without curve arithmetic: checked against
    the golden commitments (compressed by the oracle) and the identity convention."""
    from bp_pp_b200.shard import compress64
    assert compress64(b"\0" * 64) == b"\0" * 33
    x = bytes(range(32)); y_even = bytes(31) + b"\x02"; y_odd = bytes(31) + b"\x03"
    assert compress64(x + y_even) == b"\x02" + x and compress64(x + y_odd) == b"\x03" + x


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bp_pp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_serializable_proof_json_framing(golden):
    """bp_pp_b200.serde reproduces the serde_json object the oracle emits for reciprocal::SerializableProof and round-trips."""
    from bp_pp_b200 import serde
    c = golden["cases"][0]
    rec = bytes.fromhex(c["proof"])
    obj = serde.reciprocal_record_to_obj(rec)
    assert obj == c["json"]
    assert serde.reciprocal_obj_to_record(obj) == rec
    assert serde.loads_reciprocal(serde.dumps_reciprocal(rec)) == rec


def test_serde_identity_point_is_the_one_byte_sec1_encoding(golden):
    """k256 serialises AffinePoint::IDENTITY as the single SEC1 byte 00 (not 33 zero bytes): the JSON framing must emit
    and accept that form while the engine's record keeps the 33-byte GroupEncoding form."""
    from bp_pp_b200 import serde
    rec = bytearray(bytes.fromhex(golden["cases"][0]["proof"]))
    rec[33:66] = bytes(33)            # c_r := identity
    rec[492:525] = bytes(33)          # r := identity
    obj = serde.reciprocal_record_to_obj(bytes(rec))
    assert obj["circuit_proof"]["c_r"] == "00" and obj["r"] == "00"
    assert serde.reciprocal_obj_to_record(obj) == bytes(rec)
    import pytest
    with pytest.raises(ValueError):
        serde.reciprocal_obj_to_record({**obj, "r": "0000"})


def test_rust_ffi_declarations_match_the_header():
    """rust/bp-pp-gpu/src/ffi.rs (authored without a toolchain) must bind only symbols the library exports, with the
    argument COUNT the C header declares -- the cheapest check that the crate would at least link."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bppp.h")).read(), flags=re.S)
    c_args = {}
    for m in re.finditer(r"\b(bppp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        c_args[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    rs = open(os.path.join(ROOT, "rust", "bp-pp-gpu", "src", "ffi.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    found = re.findall(r"pub fn (bppp_[a-z0-9_]+)\s*\(([^)]*)\)", rs, flags=re.S)
    assert len(found) >= 35
    for name, args in found:
        assert name in c_args, f"ffi.rs declares {name}, which include/bppp.h does not"
        n = 0 if not args.strip() else args.count(":")
        assert n == c_args[name], f"{name}: {n} arguments in ffi.rs, {c_args[name]} in bppp.h"
