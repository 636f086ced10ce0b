// Exercises include/bppp.hpp the way a reference user would write it:
//   let proto = U64RangeProofProtocol{g, g_vec, h_vec}; let v = proto.commit_value(x, &s);
//   let proof = proto.prove(x, &s, &mut pt, &mut rng);  assert!(proto.verify(&v, proof, &mut vt));
// (src/range_proof/u64_proof.rs tests, src/tests.rs u64_range_proof_works).  Generators arrive on stdin as 49 x 64 bytes,
// then x (8 bytes LE), blind (32), rng (3328).  Prints the commitment, the proof record and the verdicts as hex so the
// Python test can compare them with the oracle.  Without a CUDA device construction must throw (no CPU fallback).
#include <cstdio>
#include <cstring>
#include <iostream>

#include "bppp.hpp"

using namespace bp_pp;
using bp_pp::range_proof::u64_proof::U64RangeProofProtocol;

template <class T> static void hex(const char *k, const T &v) {
    std::printf("%s=", k);
    for (uint8_t b : v) std::printf("%02x", b);
    std::printf("\n");
}

int main() {
    std::vector<uint8_t> in((49 * 64) + 8 + 32 + BPPP_U64_RNG_BYTES);
    if (std::fread(in.data(), 1, in.size(), stdin) != in.size()) { std::fprintf(stderr, "short input\n"); return 2; }
    const uint8_t *p = in.data();
    Point g; std::memcpy(g.data(), p, 64); p += 64;
    std::vector<Point> g_vec(16), h_vec(32);
    for (auto &q : g_vec) { std::memcpy(q.data(), p, 64); p += 64; }
    for (auto &q : h_vec) { std::memcpy(q.data(), p, 64); p += 64; }
    uint64_t x; std::memcpy(&x, p, 8); p += 8;
    Scalar s; std::memcpy(s.data(), p, 32); p += 32;
    std::vector<uint8_t> rng(p, p + BPPP_U64_RNG_BYTES);
    try {
        U64RangeProofProtocol proto(g, g_vec, h_vec, 0, 8, 16);
        auto v = proto.commit_value(x, s);
        auto proof = proto.prove(x, s, "u64 range proof", rng);
        hex("commit", v);
        hex("proof", proof.record);
        std::printf("verify=%d\n", (int)proto.verify(v, proof, "u64 range proof"));
        auto bad = proof; bad.record[491] ^= 1;   // last byte of the scalar n
        std::printf("verify_tampered=%d\n", (int)proto.verify(v, bad, "u64 range proof"));
        auto mal = proof; std::memset(mal.record.data(), 0xff, 33);
        try { proto.verify(v, mal, "u64 range proof"); std::printf("malformed=accepted\n"); }
        catch (const Malformed &e) { std::printf("malformed=%d\n", e.status); }

        wnla::WeightNormLinearArgument w{g, {g_vec.begin(), g_vec.begin() + 4}, {h_vec.begin(), h_vec.begin() + 4}, {}, {}, {}};
        Scalar one{}; one[31] = 1; Scalar two{}; two[31] = 2; Scalar four{}; four[31] = 4;
        w.c = {one, two, four, two}; w.rho = two; w.mu = four;
        std::vector<Scalar> l = {two, one, four, one}, n = {one, four, two, two};
        auto com = w.commit(l, n);
        auto wp = w.prove(com, "wnla", l, n);
        hex("wnla_commit", com);
        std::printf("wnla_rounds=%zu\nwnla_verify=%d\n", wp.r.size(), (int)w.verify(com, "wnla", wp));
    } catch (const Panic &e) { std::printf("panic=%d %s\n", e.status, e.what()); return 3;
    } catch (const Error &e) { std::printf("error=%s\n", e.what()); return 4; }
    return 0;
}
