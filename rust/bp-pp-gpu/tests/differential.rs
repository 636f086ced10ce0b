//! Differential tests against the REAL reference (`bp-pp` 0.1.1 on k256 / merlin) -- what finally pins the conventions the
//! build image could only recall (identity encoding, generate_biased, serde hex, field order).  Needs a CUDA
//! device and a Rust toolchain; never run in the build image.
//!
//!   cargo test --release -- --test-threads=1
use k256::elliptic_curve::group::GroupEncoding;
use k256::elliptic_curve::rand_core::{CryptoRng, RngCore, SeedableRng};
use k256::{ProjectivePoint, Scalar};
use rand_chacha::ChaCha20Rng;

use bp_pp_gpu::range_proof::reciprocal::{Proof, SerializableProof};
use bp_pp_gpu::range_proof::u64_proof::{U64RangeProofProtocol, G_VEC_FULL_SZ, H_VEC_FULL_SZ};

/// An RNG that replays a byte string: 64 bytes per `fill_bytes(64)` call, exactly what `Scalar::generate_biased` pulls.
struct Replay { bytes: Vec<u8>, pos: usize }
impl RngCore for Replay {
    fn next_u32(&mut self) -> u32 { let mut b = [0u8; 4]; self.fill_bytes(&mut b); u32::from_le_bytes(b) }
    fn next_u64(&mut self) -> u64 { let mut b = [0u8; 8]; self.fill_bytes(&mut b); u64::from_le_bytes(b) }
    fn fill_bytes(&mut self, dst: &mut [u8]) { dst.copy_from_slice(&self.bytes[self.pos..self.pos + dst.len()]); self.pos += dst.len(); }
    fn try_fill_bytes(&mut self, dst: &mut [u8]) -> Result<(), k256::elliptic_curve::rand_core::Error> { self.fill_bytes(dst); Ok(()) }
}
impl CryptoRng for Replay {}

fn gens(rng: &mut ChaCha20Rng) -> (ProjectivePoint, Vec<ProjectivePoint>, Vec<ProjectivePoint>) {
    use k256::elliptic_curve::Group;
    (ProjectivePoint::random(&mut *rng), (0..G_VEC_FULL_SZ).map(|_| ProjectivePoint::random(&mut *rng)).collect(), (0..H_VEC_FULL_SZ).map(|_| ProjectivePoint::random(&mut *rng)).collect())
}

/// Same witnesses, same RNG bytes, same (pre-loaded) transcript: the engine's proof must be the reference's, byte for byte,
/// and both must leave the transcript in the same state.
#[test]
fn proofs_are_byte_identical_to_bp_pp() {
    let mut seed = ChaCha20Rng::seed_from_u64(7);
    let (g, g_vec, h_vec) = gens(&mut seed);
    let ours = U64RangeProofProtocol::new(g, g_vec.clone(), h_vec.clone());
    let theirs = bp_pp::range_proof::u64_proof::U64RangeProofProtocol { g, g_vec, h_vec };
    for (i, x) in [0u64, 1, 123456, 0x0123456789ABCDEF, u64::MAX].into_iter().enumerate() {
        let s = Scalar::generate_biased(&mut seed);
        let mut draws = vec![0u8; 52 * 64];
        seed.fill_bytes(&mut draws);
        let mut t_a = merlin::Transcript::new(b"outer protocol");
        t_a.append_message(b"prior", &[i as u8; 9]);
        let mut t_b = t_a.clone();
        let p_ref = theirs.prove(x, &s, &mut t_a, &mut Replay { bytes: draws.clone(), pos: 0 });
        let p_gpu = ours.prove(x, &s, &mut t_b, &mut Replay { bytes: draws, pos: 0 });
        let j_ref = serde_json::to_string(&bp_pp::range_proof::reciprocal::SerializableProof::from(&p_ref)).unwrap();
        let j_gpu = serde_json::to_string(&SerializableProof::from(&p_gpu)).unwrap();
        assert_eq!(j_ref, j_gpu, "serialized proofs differ for x = {x}");
        let (mut a, mut b) = ([0u8; 32], [0u8; 32]);
        t_a.challenge_bytes(b"after", &mut a); t_b.challenge_bytes(b"after", &mut b);
        assert_eq!(a, b, "transcript state differs after prove");
        assert_eq!(theirs.commit_value(x, &s).to_bytes(), ours.commit_value(x, &s).to_bytes());
        // cross verification, and identical verdicts on a tampered proof
        let v = ours.commit_value(x, &s);
        assert!(theirs.verify(&v, (&serde_json::from_str::<bp_pp::range_proof::reciprocal::SerializableProof>(&j_gpu).unwrap()).into(), &mut merlin::Transcript::new(b"outer protocol").tap(i)));
        assert!(ours.verify(&v, (&serde_json::from_str::<SerializableProof>(&j_ref).unwrap()).into(), &mut merlin::Transcript::new(b"outer protocol").tap(i)));
        let mut bad: Proof = p_gpu.clone();
        bad.circuit_proof.n[0] += Scalar::ONE;
        assert!(!ours.verify(&v, bad, &mut merlin::Transcript::new(b"outer protocol").tap(i)));
    }
}

/// The golden vectors the CUDA path is pinned to in the build image (tests/golden/u64_golden.json, frozen from the Python
/// oracle) replayed through real k256: commitment, 525-byte record and serde_json form must match.
#[test]
fn golden_vectors_match_real_k256() {
    let gold: serde_json::Value = serde_json::from_str(include_str!("../../../tests/golden/u64_golden.json")).unwrap();
    let pt = |h: &str| -> ProjectivePoint {
        use k256::elliptic_curve::sec1::FromEncodedPoint;
        let b = hex::decode(h).unwrap();
        let enc = k256::EncodedPoint::from_affine_coordinates(k256::FieldBytes::from_slice(&b[..32]), k256::FieldBytes::from_slice(&b[32..]), false);
        ProjectivePoint::from(Option::<k256::AffinePoint>::from(k256::AffinePoint::from_encoded_point(&enc)).unwrap())
    };
    let g: Vec<ProjectivePoint> = gold["generators"].as_array().unwrap().iter().map(|v| pt(v.as_str().unwrap())).collect();
    let theirs = bp_pp::range_proof::u64_proof::U64RangeProofProtocol { g: g[0], g_vec: g[1..17].to_vec(), h_vec: g[17..49].to_vec() };
    for case in gold["cases"].as_array().unwrap() {
        use k256::elliptic_curve::PrimeField;
        let x = case["x"].as_u64().unwrap();
        let s = Option::<Scalar>::from(Scalar::from_repr(*k256::FieldBytes::from_slice(&hex::decode(case["blind"].as_str().unwrap()).unwrap()))).unwrap();
        // rng bytes: SHAKE256("bppp-bench" || "rng" || LE64(rng_index)), 3328 bytes -- regenerate with tests/golden/make_golden.py conventions
        let draws = bp_pp_gpu_test_support::shake_rng(case["rng_index"].as_u64().unwrap());
        let mut t = merlin::Transcript::new(b"u64 range proof");
        let p = theirs.prove(x, &s, &mut t, &mut Replay { bytes: draws, pos: 0 });
        assert_eq!(hex::encode(theirs.commit_value(x, &s).to_bytes()), case["commitment"].as_str().unwrap());
        let ours: Proof = (&serde_json::from_value::<SerializableProof>(serde_json::to_value(bp_pp::range_proof::reciprocal::SerializableProof::from(&p)).unwrap()).unwrap()).into();
        assert_eq!(hex::encode(ours.to_record()), case["proof"].as_str().unwrap(), "k256 disagrees with the frozen golden record for x = {x}");
        assert_eq!(serde_json::to_value(bp_pp::range_proof::reciprocal::SerializableProof::from(&p)).unwrap(), case["json"]);
    }
}

trait Tap { fn tap(self, i: usize) -> Self; }
impl Tap for merlin::Transcript { fn tap(mut self, i: usize) -> Self { self.append_message(b"prior", &[i as u8; 9]); self } }

mod bp_pp_gpu_test_support {
    /// SHAKE256("bppp-bench" || "rng" || LE64(i)) squeezed to 52 x 64 bytes (bp_pp_b200/synth.py: synth_rng_bytes); uses the Keccak
    /// inside merlin's own dependency tree to avoid another crate.
    pub fn shake_rng(i: u64) -> Vec<u8> {
        use keccak::f1600;
        let mut msg = b"bppp-benchrng".to_vec();
        msg.extend_from_slice(&i.to_le_bytes());
        let rate = 136usize;
        let mut st = [0u64; 25];
        let mut block = msg.clone();
        block.push(0x1F);
        while block.len() % rate != 0 { block.push(0); }
        let last = block.len() - 1;
        block[last] |= 0x80;
        for chunk in block.chunks(rate) {
            for (k, w) in chunk.chunks(8).enumerate() { st[k] ^= u64::from_le_bytes(w.try_into().unwrap()); }
            f1600(&mut st);
        }
        let mut out = Vec::with_capacity(52 * 64);
        while out.len() < 52 * 64 {
            for k in 0..rate / 8 { out.extend_from_slice(&st[k].to_le_bytes()); }
            f1600(&mut st);
        }
        out.truncate(52 * 64);
        out
    }
}
