//! k256 values <-> the byte conventions of include/bppp.h.
use k256::elliptic_curve::group::GroupEncoding;
use k256::elliptic_curve::sec1::{FromEncodedPoint, ToEncodedPoint};
use k256::elliptic_curve::PrimeField;
use k256::{AffinePoint, EncodedPoint, FieldBytes, ProjectivePoint, Scalar};

/// 64 bytes x || y big-endian, all-zero for the identity (`BPPP_FMT_AFFINE64`): no square root on the device.
pub fn point64(p: &ProjectivePoint) -> [u8; 64] {
    let mut out = [0u8; 64];
    let enc = p.to_affine().to_encoded_point(false);
    if let (Some(x), Some(y)) = (enc.x(), enc.y()) {
        out[..32].copy_from_slice(x);
        out[32..].copy_from_slice(y);
    }
    out
}
pub fn points64(ps: &[ProjectivePoint]) -> Vec<u8> { ps.iter().flat_map(|p| point64(p)).collect() }
/// 33 bytes, what `GroupEncoding::to_bytes` yields (identity = 33 zero bytes): the form the transcript absorbs.
pub fn point33(p: &ProjectivePoint) -> [u8; 33] {
    let mut out = [0u8; 33];
    out.copy_from_slice(p.to_bytes().as_slice());
    out
}
pub fn point_from33(b: &[u8]) -> ProjectivePoint {
    if b.iter().all(|v| *v == 0) { return ProjectivePoint::IDENTITY; }
    let enc = EncodedPoint::from_bytes(b).expect("engine returned a malformed point encoding");
    ProjectivePoint::from(Option::<AffinePoint>::from(AffinePoint::from_encoded_point(&enc)).expect("engine returned an off-curve point"))
}
pub fn scalar32(s: &Scalar) -> [u8; 32] { s.to_repr().into() }
pub fn scalars32(ss: &[Scalar]) -> Vec<u8> { ss.iter().flat_map(|s| scalar32(s)).collect() }
pub fn scalar_from32(b: &[u8]) -> Scalar {
    Option::<Scalar>::from(Scalar::from_repr(*FieldBytes::from_slice(b))).expect("engine returned a non-canonical scalar")
}
