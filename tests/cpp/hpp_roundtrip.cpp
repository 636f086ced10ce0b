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

        // ReciprocalRangeProofProtocol with the u64 dimensions (u64_proof.rs:42-54 builds exactly this): must give the fast path's record
        namespace rr = bp_pp::range_proof::reciprocal;
        rr::ReciprocalRangeProofProtocol rp{16, 16, g, g_vec, {h_vec.begin(), h_vec.begin() + 26}, {}, {h_vec.begin() + 26, h_vec.end()}};
        rr::Witness rw; rw.x = Scalar{}; rw.s = s;
        for (int k = 0; k < 8; k++) rw.x[31 - k] = (uint8_t)(x >> (8 * k));
        for (uint64_t d : U64RangeProofProtocol::u64_to_hex(x)) rw.digits.push_back((uint32_t)d);
        CompressedPoint rcom;
        auto rproof = rp.prove(rw, "u64 range proof", rng, &rcom);
        hex("reciprocal_commit", rcom);
        hex("reciprocal_proof", rproof.record);
        std::printf("reciprocal_shape=%zu,%zu,%zu\nreciprocal_verify=%d\n", rproof.rounds, rproof.l_len, rproof.n_len, (int)rp.verify(rcom, rproof, "u64 range proof"));
        hex("reciprocal_commit_value", rp.commit_value(rw.x, rw.s));

        // ArithmeticCircuit: the reference's ac_works instance (src/tests.rs:44-136): x y = z, x + y = r with x = 3, y = 5
        auto sc = [](uint64_t v) { Scalar o{}; for (int k = 0; k < 8; k++) o[31 - k] = (uint8_t)(v >> (8 * k)); return o; };
        auto neg = [&](uint64_t v) {                      // n - v, n = group order
            static const uint8_t N[32] = {0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFE,
                                          0xBA,0xAE,0xDC,0xE6,0xAF,0x48,0xA0,0x3B,0xBF,0xD2,0x5E,0x8C,0xD0,0x36,0x41,0x41};
            Scalar o; int borrow = 0; Scalar sv = sc(v);
            for (int k = 31; k >= 0; k--) { int t = (int)N[k] - (int)sv[k] - borrow; borrow = t < 0; o[k] = (uint8_t)(t + (borrow ? 256 : 0)); }
            return o;
        };
        circuit::ArithmeticCircuit ac;
        ac.dim_nm = 1; ac.dim_no = 2; ac.k = 1; ac.dim_nv = 2; ac.f_l = true; ac.f_m = false;
        ac.g = g; ac.g_vec = {g_vec[0]}; ac.h_vec = {h_vec.begin(), h_vec.begin() + 11}; ac.h_vec_ = {h_vec.begin() + 11, h_vec.begin() + 16};
        ac.W_m = {sc(0), sc(0), sc(1), sc(0)};
        ac.W_l = {sc(0), sc(1), sc(0), sc(0), sc(0), neg(1), sc(1), sc(0)};
        ac.a_m = {sc(0)}; ac.a_l = {neg(8), neg(15)};
        ac.partition = {{-1, -1}, {0, 1}, {-1, -1}, {-1, -1}};
        circuit::Witness cw{{sc(3), sc(5)}, {s}, {sc(3)}, {sc(5)}, {sc(15), sc(8)}};
        auto ccom = ac.commit(cw.v, s);
        std::vector<uint8_t> crng(rng.begin(), rng.begin() + (18 + 2 + 1) * 64);
        auto cproof = ac.prove({ccom}, cw, "circuit test", crng);
        hex("circuit_commit", ccom);
        hex("circuit_proof", cproof.record);
        std::printf("circuit_shape=%zu,%zu,%zu\ncircuit_verify=%d\n", cproof.rounds, cproof.l_len, cproof.n_len, (int)ac.verify({ccom}, cproof, "circuit test"));
        auto cbad = cproof; cbad.record.back() ^= 1;
        std::printf("circuit_verify_tampered=%d\n", (int)ac.verify({ccom}, cbad, "circuit test"));
    } catch (const Panic &e) { std::printf("panic=%d %s\n", e.status, e.what()); return 3;
    } catch (const Error &e) { std::printf("error=%s\n", e.what()); return 4; }
    return 0;
}
