//! `transcript::{app_point, get_challenge}` with the reference's signatures (src/transcript.rs:6-14).  These run on the HOST:
//! merlin keeps its STROBE state private, so a caller-owned transcript can only be advanced here; the engine receives
//! the challenges and returns the points to append (phase-stepped C ABI, include/bppp.h).
use k256::elliptic_curve::group::GroupEncoding;
use k256::elliptic_curve::PrimeField;
use k256::{FieldBytes, ProjectivePoint, Scalar};
use merlin::Transcript;

pub fn app_point(label: &'static [u8], p: &ProjectivePoint, t: &mut Transcript) {
    let encoded = p.to_bytes();
    t.append_message(label, encoded.as_slice());
}

pub fn get_challenge(label: &'static [u8], t: &mut Transcript) -> Scalar {
    let mut wide = [0u8; 32];
    t.challenge_bytes(label, &mut wide);
    // panics for a value >= n exactly as the reference does (probability ~2^-128)
    Scalar::from_repr(*FieldBytes::from_slice(&wide)).unwrap()
}

/// The same two operations on raw bytes (what the engine hands back / takes).
pub(crate) fn app_point33(label: &'static [u8], p33: &[u8], t: &mut Transcript) { t.append_message(label, p33); }
pub(crate) fn challenge32(label: &'static [u8], t: &mut Transcript) -> [u8; 32] { get_challenge(label, t).to_repr().into() }
