// bppp.hpp -- header-only C++17 host layer over the C ABI (bppp.h), mirroring the reference's Rust items for the
// hot path: same type and method names, argument meaning and error behaviour (a reference `panic!` surfaces as
// bp_pp::Panic, a malformed encoding as bp_pp::Malformed).  Everything below only marshals bytes; all arithmetic
// runs on the GPU inside libbppp.so.
//
//   bp_pp::range_proof::u64_proof::U64RangeProofProtocol   src/range_proof/u64_proof.rs:19-102
//   bp_pp::wnla::WeightNormLinearArgument                   src/wnla.rs:12-190
//   bp_pp::range_proof::reciprocal::Proof (525-byte record) src/range_proof/reciprocal.rs:30-41
//   bp_pp::range_proof::reciprocal::ReciprocalRangeProofProtocol   src/range_proof/reciprocal.rs:64-214 (any dim_nd, dim_np)
//   bp_pp::circuit::ArithmeticCircuit                              src/circuit.rs:95-556 (dense W_m / W_l, tabulated partition)
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

// ---- generic records: c_l c_r c_o c_s | r[rounds] | x[rounds] | l | n (| r for the reciprocal protocol) ----
struct GenericProof {
    std::vector<uint8_t> record;
    size_t rounds = 0, l_len = 0, n_len = 0;
};
inline const uint8_t *pts_ptr(const std::vector<Point> &v) { return v.empty() ? nullptr : v[0].data(); }
inline const uint8_t *sc_ptr(const std::vector<Scalar> &v) { return v.empty() ? nullptr : v[0].data(); }
inline bool verdict_to_bool(int32_t verdict, const char *what) {
    if (verdict == BPPP_ST_PANIC_INVERT_ZERO || verdict == BPPP_ST_PANIC_CHALLENGE_RANGE) throw Panic(verdict, std::string(what) + ": the reference would panic");
    if (verdict < 0) throw Malformed(verdict, std::string(what) + ": the proof does not deserialise");
    return verdict == BPPP_ST_TRUE;
}

namespace range_proof {
namespace reciprocal {
// reciprocal.rs:22-28: the committed value, its blinding and the digits (base dim_np, dim_nd of them)
struct Witness { Scalar x, s; std::vector<uint32_t> digits; };

// reciprocal.rs:64-84, for arbitrary (dim_nd, dim_np); make_circuit (:150-214) is built inside the engine from the challenge
struct ReciprocalRangeProofProtocol {
    size_t dim_nd = 0, dim_np = 0;
    Point g; std::vector<Point> g_vec, h_vec, g_vec_, h_vec_;
    int device = 0;

    // reciprocal.rs:88-90
    CompressedPoint commit_value(const Scalar &x, const Scalar &s) const {
        if (h_vec.empty()) throw Panic(BPPP_ST_BAD_ARG, "index out of bounds: h_vec is empty");
        CompressedPoint out;
        check(bppp_reciprocal_commit_value(device, g.data(), h_vec[0].data(), x.data(), s.data(), out.data()), "bppp_reciprocal_commit_value");
        return out;
    }
    // reciprocal.rs:110-146 with a fresh Transcript::new(label); rng = (1 + 18 + (dim_nd + 1) + dim_nd) x 64 bytes in draw order
    GenericProof prove(const Witness &w, const std::string &label, const std::vector<uint8_t> &rng, CompressedPoint *commitment_out = nullptr) const {
        GenericProof pr;
        pr.record.resize(33 * (4 + 2 * 64) + 32 * (2 * (dim_nd + dim_np) + 64));
        CompressedPoint com; int32_t st = 0;
        check(bppp_reciprocal_prove(device, dim_nd, dim_np, g.data(), pts_ptr(g_vec), g_vec.size(), pts_ptr(h_vec), h_vec.size(), pts_ptr(g_vec_), g_vec_.size(),
                                    pts_ptr(h_vec_), h_vec_.size(), w.x.data(), w.s.data(), w.digits.data(), rng.data(), rng.size(),
                                    (const uint8_t *)label.data(), label.size(), pr.record.data(), pr.record.size(), &pr.rounds, &pr.l_len, &pr.n_len, com.data(), &st),
              "bppp_reciprocal_prove");
        if (st != BPPP_ST_TRUE) throw Panic(st, "reciprocal prove: the reference would panic");
        pr.record.resize(33 * (4 + 2 * pr.rounds) + 32 * (pr.l_len + pr.n_len) + 33);      // ... | l | n | r (the pole commitment, a point)
        if (commitment_out) *commitment_out = com;
        return pr;
    }
    // reciprocal.rs:98-107
    bool verify(const CompressedPoint &commitment, const GenericProof &pr, const std::string &label) const {
        int32_t verdict = 0;
        check(bppp_reciprocal_verify(device, dim_nd, dim_np, g.data(), pts_ptr(g_vec), g_vec.size(), pts_ptr(h_vec), h_vec.size(), pts_ptr(g_vec_), g_vec_.size(),
                                     pts_ptr(h_vec_), h_vec_.size(), commitment.data(), pr.record.data(), pr.rounds, pr.rounds, pr.l_len, pr.n_len,
                                     (const uint8_t *)label.data(), label.size(), &verdict), "bppp_reciprocal_verify");
        return verdict_to_bool(verdict, "reciprocal verify");
    }
};
}  // namespace reciprocal
}  // namespace range_proof

namespace circuit {
// circuit.rs:15-20: the partition closure tabulated (index into w_o, or -1 for None)
struct Partition { std::vector<int32_t> lo, ll, lr, no; };
// circuit.rs:78-93
struct Witness { std::vector<Scalar> v, s_v, w_l, w_r, w_o; };     // v: k x dim_nv row-major

// circuit.rs:95-139 with dense row-major W_m (dim_nm x dim_nw) and W_l (dim_nl x dim_nw), dim_nl = dim_nv k, dim_nw = 2 dim_nm + dim_no
struct ArithmeticCircuit {
    size_t dim_nm = 0, dim_no = 0, k = 0, dim_nv = 0;
    bool f_l = false, f_m = false;
    Point g; std::vector<Point> g_vec, h_vec, g_vec_, h_vec_;
    std::vector<Scalar> W_m, W_l, a_m, a_l;
    Partition partition;
    int device = 0;

    bppp_circuit_desc desc() const {
        bppp_circuit_desc d{};
        d.dim_nm = dim_nm; d.dim_no = dim_no; d.k = k; d.dim_nv = dim_nv; d.f_l = f_l; d.f_m = f_m;
        d.g64 = g.data(); d.gvec64 = pts_ptr(g_vec); d.hvec64 = pts_ptr(h_vec); d.gvec2_64 = pts_ptr(g_vec_); d.hvec2_64 = pts_ptr(h_vec_);
        d.gn = g_vec.size(); d.hn = h_vec.size(); d.gn2 = g_vec_.size(); d.hn2 = h_vec_.size();
        d.W_m32 = sc_ptr(W_m); d.W_l32 = sc_ptr(W_l); d.a_m32 = sc_ptr(a_m); d.a_l32 = sc_ptr(a_l);
        d.part_lo = partition.lo.data(); d.part_ll = partition.ll.data(); d.part_lr = partition.lr.data(); d.part_no = partition.no.data();
        d.part_n = partition.lo.size();
        return d;
    }
    // circuit.rs:146-151
    CompressedPoint commit(const std::vector<Scalar> &v, const Scalar &s) const {
        bppp_circuit_desc d = desc(); CompressedPoint out;
        check(bppp_circuit_commit(device, &d, sc_ptr(v), s.data(), out.data()), "bppp_circuit_commit");
        return out;
    }
    // circuit.rs:260-556 with a fresh Transcript::new(label); rng = (18 + dim_nv + dim_nm) x 64 bytes in draw order
    GenericProof prove(const std::vector<CompressedPoint> &v_commitments, const Witness &w, const std::string &label, const std::vector<uint8_t> &rng) const {
        bppp_circuit_desc d = desc();
        GenericProof pr;
        pr.record.resize(33 * (4 + 2 * 64) + 32 * (2 * (dim_nm + dim_nv + 9) + 64));
        int32_t st = 0;
        check(bppp_circuit_prove(device, &d, v_commitments.empty() ? nullptr : v_commitments[0].data(), sc_ptr(w.v), sc_ptr(w.s_v), sc_ptr(w.w_l), sc_ptr(w.w_r),
                                 sc_ptr(w.w_o), rng.data(), rng.size(), (const uint8_t *)label.data(), label.size(), pr.record.data(), pr.record.size(), &pr.rounds,
                                 &pr.l_len, &pr.n_len, &st), "bppp_circuit_prove");
        if (st != BPPP_ST_TRUE) throw Panic(st, "circuit prove: the reference would panic");
        pr.record.resize(33 * (4 + 2 * pr.rounds) + 32 * (pr.l_len + pr.n_len));
        return pr;
    }
    // circuit.rs:154-256
    bool verify(const std::vector<CompressedPoint> &v_commitments, const GenericProof &pr, const std::string &label) const {
        bppp_circuit_desc d = desc(); int32_t verdict = 0;
        check(bppp_circuit_verify(device, &d, v_commitments.empty() ? nullptr : v_commitments[0].data(), pr.record.data(), pr.rounds, pr.rounds, pr.l_len, pr.n_len,
                                  (const uint8_t *)label.data(), label.size(), &verdict), "bppp_circuit_verify");
        return verdict_to_bool(verdict, "circuit verify");
    }
};
}  // namespace circuit

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
