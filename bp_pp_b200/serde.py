"""Wire form of the reference's `SerializableProof` models (src/wnla.rs:33-39, src/circuit.rs:36-46,
src/range_proof/reciprocal.rs:37-41): serde_json objects whose points are upper-case hex of the 33-byte SEC1 compressed
form and whose scalars are upper-case hex of 32 big-endian bytes [recalled k256/serdect conventions, SURVEY App. D].
The engine itself consumes and produces the raw records; this module is the thin JSON framing around them."""
from __future__ import annotations

import json
from typing import Dict, List


_ID33 = bytes(33)


def _pt(b: bytes) -> str:
    """One point.  The identity is the 1-byte SEC1 encoding "00" on the wire (k256 `AffinePoint: Serialize` goes through
    sec1::EncodedPoint) although it is 33 zero bytes inside the engine's records (GroupEncoding::to_bytes)."""
    return "00" if b == _ID33 else b.hex().upper()


def _unpt(h: str) -> bytes:
    b = bytes.fromhex(h)
    if b == b"\0":
        return _ID33
    if len(b) != 33:
        raise ValueError("a compressed point is 33 bytes (or the single byte 00 for the identity)")
    return b


def _pts(b: bytes) -> List[str]:
    assert len(b) % 33 == 0
    return [_pt(b[i:i + 33]) for i in range(0, len(b), 33)]


def _scs(b: bytes) -> List[str]:
    assert len(b) % 32 == 0
    return [b[i:i + 32].hex().upper() for i in range(0, len(b), 32)]


def wnla_proof_to_obj(r: bytes, x: bytes, l: bytes, n: bytes) -> Dict:
    """wnla::SerializableProof { r, x, l, n }"""
    return {"r": _pts(r), "x": _pts(x), "l": _scs(l), "n": _scs(n)}


def circuit_record_to_obj(rec: bytes, rounds: int, l_len: int, n_len: int) -> Dict:
    """circuit::SerializableProof { c_l, c_r, c_o, c_s, r, x, l, n } from the record c_l c_r c_o c_s | r | x | l | n"""
    o = 132
    r, x = rec[o:o + 33 * rounds], rec[o + 33 * rounds:o + 66 * rounds]
    o += 66 * rounds
    l, n = rec[o:o + 32 * l_len], rec[o + 32 * l_len:o + 32 * (l_len + n_len)]
    d = {"c_l": _pt(rec[0:33]), "c_r": _pt(rec[33:66]), "c_o": _pt(rec[66:99]), "c_s": _pt(rec[99:132])}
    d.update(wnla_proof_to_obj(r, x, l, n))
    return d


def reciprocal_record_to_obj(rec: bytes, rounds: int = 4, l_len: int = 2, n_len: int = 1) -> Dict:
    """reciprocal::SerializableProof { circuit_proof, r }; defaults are the u64 shape (525-byte record)."""
    body = 132 + 66 * rounds + 32 * (l_len + n_len)
    assert len(rec) == body + 33
    return {"circuit_proof": circuit_record_to_obj(rec[:body], rounds, l_len, n_len), "r": _pt(rec[body:])}


def reciprocal_obj_to_record(obj: Dict) -> bytes:
    cp = obj["circuit_proof"]
    h, p = bytes.fromhex, _unpt
    out = p(cp["c_l"]) + p(cp["c_r"]) + p(cp["c_o"]) + p(cp["c_s"])
    out += b"".join(p(q) for q in cp["r"]) + b"".join(p(q) for q in cp["x"])
    out += b"".join(h(s) for s in cp["l"]) + b"".join(h(s) for s in cp["n"])
    return out + p(obj["r"])


def dumps_reciprocal(rec: bytes, **kw) -> str:
    return json.dumps(reciprocal_record_to_obj(rec, **kw), indent=2)


def loads_reciprocal(text: str) -> bytes:
    return reciprocal_obj_to_record(json.loads(text))
