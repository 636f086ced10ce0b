// bppp.hpp -- header-only C++17 host layer over the C ABI (bppp.h), mirroring the reference's Rust items for the
// hot path: same type and method names, argument meaning and error behaviour (a reference `panic!` surfaces as
// bp_pp::Panic, a malformed encoding as bp_pp::Malformed).  Everything below only marshals bytes; all arithmetic
// runs on the GPU inside libbppp.so.
//
//   bp_pp::range_proof::u64_proof::U64RangeProofProtocol   src/range_proof/u64_proof.rs:19-102
//   bp_pp::wnla::WeightNormLinearArgument                   src/wnla.rs:12-190
//   bp_pp::range_proof::reciprocal::Proof (525-byte record) src/range_proof/reciprocal.rs:30-41
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "bppp.h"

namespace bp_pp {

using Scalar = std::array<uint8_t, 32>;        // 32-byte big-endian, canonical (k256::Scalar::to_bytes)
using Point = std::array<uint8_t, 64>;         // affine x || y, all-zero = identity (what a k256::AffinePoint yields uncompressed)
using CompressedPoint = std::array<uint8_t, 33>;

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
struct Panic : Error { int32_t status; Panic(int32_t s, const std::string &w) : Error(w), status(s) {} };       // the reference would panic
struct Malformed : Error { int32_t status; Malformed(int32_t s, const std::string &w) : Error(w), status(s) {} };  // deserialisation failure

inline void check(int rc, const char *what) {
    if (rc != BPPP_OK) throw Error(std::string(what) + " failed (" + std::to_string(rc) + "): " + bppp_last_error());
}

namespace range_proof {
namespace reciprocal {
// reciprocal::SerializableProof as the 525-byte record c_l c_r c_o c_s | r[4] | x[4] | l[2] | n[1] | r
struct Proof { std::array<uint8_t, BPPP_U64_PROOF_BYTES> record; };
}  // namespace reciprocal

namespace u64_proof {
constexpr size_t G_VEC_FULL_SZ = 16, H_VEC_CIRCUIT_SZ = 26, H_VEC_FULL_SZ = 32;   // u64_proof.rs:12-14

class U64RangeProofProtocol {
public:
    static constexpr size_t DIM_ND = 16, DIM_NP = 16;
    Point g; std::vector<Point> g_vec, h_vec;

    U64RangeProofProtocol(const Point &g_, std::vector<Point> g_vec_, std::vector<Point> h_vec_, int device = 0, int window_bits = 0,
                          size_t max_batch = 65536)
        : g(g_), g_vec(std::move(g_vec_)), h_vec(std::move(h_vec_)) {
        if (g_vec.size() != G_VEC_FULL_SZ || h_vec.size() != H_VEC_FULL_SZ) throw Panic(BPPP_ST_BAD_ARG, "index out of bounds: g_vec needs 16 points, h_vec 32");
        std::vector<uint8_t> gens;
        gens.insert(gens.end(), g.begin(), g.end());
        for (auto &p : g_vec) gens.insert(gens.end(), p.begin(), p.end());
        for (auto &p : h_vec) gens.insert(gens.end(), p.begin(), p.end());
        check(bppp_ctx_create(&ctx_, device, gens.data(), window_bits, max_batch), "bppp_ctx_create");
    }
    ~U64RangeProofProtocol() { bppp_ctx_destroy(ctx_); }
    U64RangeProofProtocol(const U64RangeProofProtocol &) = delete;
    U64RangeProofProtocol &operator=(const U64RangeProofProtocol &) = delete;

    // u64_proof.rs:37-39
    CompressedPoint commit_value(uint64_t x, const Scalar &s) const { return commit_batch({x}, {s}).at(0); }
    std::vector<CompressedPoint> commit_batch(const std::vector<uint64_t> &xs, const std::vector<Scalar> &blinds) const {
        if (blinds.size() != xs.size()) throw Error("commit_batch: xs / blinds length mismatch");
        std::vector<CompressedPoint> out(xs.size());
        check(bppp_u64_commit_batch(ctx_, xs.size(), xs.data(), flat(blinds), BPPP_FMT_COMPRESSED, out.empty() ? nullptr : out[0].data()), "bppp_u64_commit_batch");
        return out;
    }
    // u64_proof.rs:57-82 over N witnesses, each with a fresh Transcript::new(label).  rng: the 52 x 64 bytes per proof an
    // RngCore would have produced, in draw order.  Throws Panic where the reference panics.
    std::vector<reciprocal::Proof> prove_batch(const std::vector<uint64_t> &xs, const std::vector<Scalar> &blinds, const std::vector<uint8_t> &rng,
                                               const std::string &label) const {
        const size_t n = xs.size();
        if (blinds.size() != n || rng.size() != n * BPPP_U64_RNG_BYTES) throw Error("prove_batch: blinds / rng length mismatch");
        std::vector<reciprocal::Proof> out(n);
        std::vector<int32_t> st(n);
        check(bppp_u64_prove_batch(ctx_, n, xs.data(), flat(blinds), rng.data(), (const uint8_t *)label.data(), label.size(),
                                   n ? out[0].record.data() : nullptr, st.data()), "bppp_u64_prove_batch");
        for (size_t i = 0; i < n; i++) if (st[i] != BPPP_ST_TRUE) throw Panic(st[i], "prove: the reference would panic on proof " + std::to_string(i));
        return out;
    }
    reciprocal::Proof prove(uint64_t x, const Scalar &s, const std::string &transcript_label, const std::vector<uint8_t> &rng_bytes) const {
        return prove_batch({x}, {s}, rng_bytes, transcript_label).at(0);
    }
    // u64_proof.rs:42-54 over N proofs, one status per proof (BPPP_ST_*): 1 / 0 are the reference's true / false; a negative
    // status marks a record the reference could not have deserialised (BAD_POINT / BAD_SCALAR) or on which it would have
    // panicked.  One bad record never affects the verdicts of the others.
    std::vector<int32_t> verify_batch_status(const std::vector<CompressedPoint> &vs, const std::vector<reciprocal::Proof> &proofs, const std::string &label) const {
        const size_t n = vs.size();
        if (proofs.size() != n) throw Error("verify_batch: length mismatch");
        std::vector<int32_t> st(n);
        static_assert(sizeof(reciprocal::Proof) == BPPP_U64_PROOF_BYTES, "Proof must be the packed 525-byte record");
        check(bppp_u64_verify_batch(ctx_, n, n ? vs[0].data() : nullptr, n ? proofs[0].record.data() : nullptr, BPPP_FMT_COMPRESSED,
                                    (const uint8_t *)label.data(), label.size(), st.data()), "bppp_u64_verify_batch");
        return st;
    }
    // the same as booleans: only status 1 is true (a malformed or panicking record is "not verified"; use verify_batch_status to tell them apart)
    std::vector<bool> verify_batch(const std::vector<CompressedPoint> &vs, const std::vector<reciprocal::Proof> &proofs, const std::string &label) const {
        std::vector<int32_t> st = verify_batch_status(vs, proofs, label);
        std::vector<bool> out(st.size());
        for (size_t i = 0; i < st.size(); i++) out[i] = st[i] == BPPP_ST_TRUE;
        return out;
    }
    // single proof: true / false exactly as the reference; throws Panic where it panics, Malformed where it could not deserialise
    bool verify(const CompressedPoint &v, const reciprocal::Proof &proof, const std::string &transcript_label) const {
        int32_t st = verify_batch_status({v}, {proof}, transcript_label).at(0);
        if (st == BPPP_ST_PANIC_INVERT_ZERO || st == BPPP_ST_PANIC_CHALLENGE_RANGE) throw Panic(st, "verify: the reference would panic on this proof");
        if (st < 0) throw Malformed(st, "verify: the proof does not deserialise");
        return st == BPPP_ST_TRUE;
    }

    // u64_proof.rs:84-102
    static std::vector<uint64_t> u64_to_hex(uint64_t x) { std::vector<uint64_t> d(16); for (auto &v : d) { v = x % 16; x /= 16; } return d; }
    static std::vector<uint64_t> u64_to_hex_mapped(uint64_t x) { std::vector<uint64_t> m(16, 0); for (int i = 0; i < 16; i++) { m[x % 16]++; x /= 16; } return m; }

    bppp_ctx *raw() const { return ctx_; }

private:
    static const uint8_t *flat(const std::vector<Scalar> &v) { return v.empty() ? nullptr : v[0].data(); }
    bppp_ctx *ctx_ = nullptr;
};
}  // namespace u64_proof
}  // namespace range_proof

namespace wnla {
// wnla::Proof { r, x, l, n } with r / x in push order (innermost round first, wnla.rs:186-188)
struct Proof { std::vector<CompressedPoint> r, x; std::vector<Scalar> l, n; };

// wnla.rs:12-19
struct WeightNormLinearArgument {
    Point g; std::vector<Point> g_vec, h_vec; std::vector<Scalar> c; Scalar rho, mu;
    int device = 0;

    // wnla.rs:66-72
    CompressedPoint commit(const std::vector<Scalar> &l, const std::vector<Scalar> &n) const {
        CompressedPoint out;
        check(bppp_wnla_commit(device, g.data(), p(g_vec), g_vec.size(), p(h_vec), h_vec.size(), s(c), c.size(), rho.data(), mu.data(), s(l), l.size(), s(n), n.size(),
                               out.data()), "bppp_wnla_commit");
        return out;
    }
    // wnla.rs:125-190, fresh Transcript::new(label)
    Proof prove(const CompressedPoint &commitment, const std::string &label, const std::vector<Scalar> &l, const std::vector<Scalar> &n) const {
        std::vector<CompressedPoint> r(64), x(64);
        std::vector<Scalar> lo(l.size() ? l.size() : 1), no(n.size() ? n.size() : 1);
        size_t rounds = 0, ll = 0, nl = 0; int32_t st = 0;
        check(bppp_wnla_prove(device, g.data(), p(g_vec), g_vec.size(), p(h_vec), h_vec.size(), s(c), c.size(), rho.data(), mu.data(), commitment.data(), s(l), l.size(),
                              s(n), n.size(), (const uint8_t *)label.data(), label.size(), r[0].data(), x[0].data(), &rounds, lo[0].data(), &ll, no[0].data(), &nl, &st),
              "bppp_wnla_prove");
        if (st != BPPP_ST_TRUE) throw Panic(st, "wnla prove: the reference would panic");
        r.resize(rounds); x.resize(rounds); lo.resize(ll); no.resize(nl);
        return Proof{r, x, lo, no};
    }
    // wnla.rs:75-121
    bool verify(const CompressedPoint &commitment, const std::string &label, const Proof &proof) const {
        int32_t verdict = 0;
        check(bppp_wnla_verify(device, g.data(), p(g_vec), g_vec.size(), p(h_vec), h_vec.size(), s(c), c.size(), rho.data(), mu.data(), commitment.data(),
                               proof.r.empty() ? nullptr : proof.r[0].data(), proof.r.size(), proof.x.empty() ? nullptr : proof.x[0].data(), proof.x.size(), s(proof.l),
                               proof.l.size(), s(proof.n), proof.n.size(), (const uint8_t *)label.data(), label.size(), &verdict), "bppp_wnla_verify");
        if (verdict == BPPP_ST_PANIC_INVERT_ZERO || verdict == BPPP_ST_PANIC_CHALLENGE_RANGE) throw Panic(verdict, "wnla verify: the reference would panic");
        if (verdict < 0) throw Malformed(verdict, "wnla verify: proof does not deserialise");
        return verdict == BPPP_ST_TRUE;
    }

private:
    static const uint8_t *p(const std::vector<Point> &v) { return v.empty() ? nullptr : v[0].data(); }
    static const uint8_t *s(const std::vector<Scalar> &v) { return v.empty() ? nullptr : v[0].data(); }
};
}  // namespace wnla

}  // namespace bp_pp
