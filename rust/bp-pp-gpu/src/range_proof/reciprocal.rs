//! `reciprocal::{Witness, Proof, SerializableProof, ReciprocalRangeProofProtocol}` (src/range_proof/reciprocal.rs:17-214).
//! The u64 shape (dim 16/16) is served by the batched kernels; other dimensions by `bppp_reciprocal_{prove,verify}`, which
//! own a fresh `Transcript::new(label)` -- the generic path has no phase-stepped form yet, so `prove` / `verify` here take
//! the transcript LABEL for (dim_nd, dim_np) != (16, 16); see INTEGRATION.md section 2.
use k256::{AffinePoint, ProjectivePoint, Scalar};
use serde::{Deserialize, Serialize};

use crate::circuit;
use crate::convert::*;
use crate::{check, check_status, ffi};

/// Reciprocal range-proof witness (reciprocal.rs:17-27).
#[derive(Clone, Debug)]
pub struct Witness {
    pub x: Scalar,
    pub s: Scalar,
    pub m: Vec<Scalar>,
    pub digits: Vec<Scalar>,
}

/// zk-proof that the committed value lies in [0, dim_np^dim_nd) (reciprocal.rs:30-34).
#[derive(Clone, Debug)]
pub struct Proof {
    pub circuit_proof: circuit::Proof,
    pub r: ProjectivePoint,
}

#[derive(Serialize, Deserialize, Clone, Debug)]
pub struct SerializableProof {
    pub circuit_proof: circuit::SerializableProof,
    pub r: AffinePoint,
}

impl From<&SerializableProof> for Proof {
    fn from(value: &SerializableProof) -> Self {
        Proof { circuit_proof: circuit::Proof::from(&value.circuit_proof), r: ProjectivePoint::from(value.r) }
    }
}
impl From<&Proof> for SerializableProof {
    fn from(value: &Proof) -> Self {
        SerializableProof { circuit_proof: circuit::SerializableProof::from(&value.circuit_proof), r: value.r.to_affine() }
    }
}

impl Proof {
    /// The engine's record: c_l c_r c_o c_s | r[..] | x[..] | l[..] | n[..] | r  (include/bppp.h; 525 bytes for the u64 shape).
    pub fn to_record(&self) -> Vec<u8> {
        let cp = &self.circuit_proof;
        let mut out = Vec::new();
        for p in [&cp.c_l, &cp.c_r, &cp.c_o, &cp.c_s] { out.extend_from_slice(&point33(p)); }
        for p in cp.r.iter().chain(cp.x.iter()) { out.extend_from_slice(&point33(p)); }
        for s in cp.l.iter().chain(cp.n.iter()) { out.extend_from_slice(&scalar32(s)); }
        out.extend_from_slice(&point33(&self.r));
        out
    }
    pub fn from_record_shaped(rec: &[u8], rounds: usize, l_len: usize, n_len: usize) -> Proof {
        let p = |k: usize| point_from33(&rec[33 * k..33 * k + 33]);
        let so = 33 * (4 + 2 * rounds);
        let s = |k: usize| scalar_from32(&rec[so + 32 * k..so + 32 * k + 32]);
        Proof {
            circuit_proof: circuit::Proof {
                c_l: p(0), c_r: p(1), c_o: p(2), c_s: p(3),
                r: (0..rounds).map(|k| p(4 + k)).collect(), x: (0..rounds).map(|k| p(4 + rounds + k)).collect(),
                l: (0..l_len).map(s).collect(), n: (0..n_len).map(|k| s(l_len + k)).collect(),
            },
            r: point_from33(&rec[so + 32 * (l_len + n_len)..so + 32 * (l_len + n_len) + 33]),
        }
    }
    pub fn from_record(rec: &[u8]) -> Proof { Proof::from_record_shaped(rec, 4, 2, 1) }
}

/// Public information of the reciprocal protocol (reciprocal.rs:64-84).
#[derive(Clone, Debug)]
pub struct ReciprocalRangeProofProtocol {
    pub dim_nd: usize,
    pub dim_np: usize,
    pub g: ProjectivePoint,
    pub g_vec: Vec<ProjectivePoint>,
    pub h_vec: Vec<ProjectivePoint>,
    pub g_vec_: Vec<ProjectivePoint>,
    pub h_vec_: Vec<ProjectivePoint>,
}

impl ReciprocalRangeProofProtocol {
    /// `commitment = x*g + s*h_vec[0]` (reciprocal.rs:88-90)
    pub fn commit_value(&self, x: &Scalar, s: &Scalar) -> ProjectivePoint {
        let pts = [point64(&self.g), point64(&self.h_vec[0])].concat();
        let scs = [scalar32(x), scalar32(s)].concat();
        let mut out = [0u8; 33];
        check(unsafe { ffi::bppp_msm(0, pts.as_ptr(), ffi::BPPP_FMT_AFFINE64, 2, scs.as_ptr(), 2, ffi::BPPP_FMT_COMPRESSED, out.as_mut_ptr()) }, "bppp_msm");
        point_from33(&out)
    }
    /// `commitment = s*h_vec[0] + <r, h_vec[9:]>` (reciprocal.rs:93-95)
    pub fn commit_poles(&self, r: &[Scalar], s: &Scalar) -> ProjectivePoint {
        let mut pts = point64(&self.h_vec[0]).to_vec();
        pts.extend_from_slice(&points64(&self.h_vec[9..]));
        let mut scs = scalar32(s).to_vec();
        scs.extend_from_slice(&scalars32(r));
        let mut out = [0u8; 33];
        check(unsafe { ffi::bppp_msm(0, pts.as_ptr(), ffi::BPPP_FMT_AFFINE64, 1 + self.h_vec.len() - 9, scs.as_ptr(), 1 + r.len(), ffi::BPPP_FMT_COMPRESSED, out.as_mut_ptr()) }, "bppp_msm");
        point_from33(&out)
    }
    /// reciprocal.rs:110-146 with a fresh `Transcript::new(label)`; `rng_bytes` = (1 + 18 + (dim_nd + 1) + dim_nd) x 64 bytes in draw order.
    pub fn prove_with_label(&self, witness: Witness, digits: &[u32], label: &[u8], rng_bytes: &[u8]) -> (ProjectivePoint, Proof) {
        let cap = 33 * (5 + 2 * 64) + 32 * 16;
        let mut out = vec![0u8; cap];
        let (mut ro, mut lo, mut no, mut st) = (0usize, 0usize, 0usize, 0i32);
        let mut com = [0u8; 33];
        let (g, gv, hv, gv2, hv2) = (point64(&self.g), points64(&self.g_vec), points64(&self.h_vec), points64(&self.g_vec_), points64(&self.h_vec_));
        check(unsafe {
            ffi::bppp_reciprocal_prove(0, self.dim_nd, self.dim_np, g.as_ptr(), gv.as_ptr(), self.g_vec.len(), hv.as_ptr(), self.h_vec.len(), gv2.as_ptr(), self.g_vec_.len(),
                                       hv2.as_ptr(), self.h_vec_.len(), scalar32(&witness.x).as_ptr(), scalar32(&witness.s).as_ptr(), digits.as_ptr(), rng_bytes.as_ptr(),
                                       rng_bytes.len(), label.as_ptr(), label.len(), out.as_mut_ptr(), cap, &mut ro, &mut lo, &mut no, com.as_mut_ptr(), &mut st)
        }, "bppp_reciprocal_prove");
        check_status(st, "ReciprocalRangeProofProtocol::prove");
        (point_from33(&com), Proof::from_record_shaped(&out, ro, lo, no))
    }
    /// reciprocal.rs:98-107 with a fresh `Transcript::new(label)`.
    pub fn verify_with_label(&self, commitment: &ProjectivePoint, proof: Proof, label: &[u8]) -> bool {
        let rec = proof.to_record();
        let cp = &proof.circuit_proof;
        let mut verdict = 0i32;
        let (g, gv, hv, gv2, hv2) = (point64(&self.g), points64(&self.g_vec), points64(&self.h_vec), points64(&self.g_vec_), points64(&self.h_vec_));
        check(unsafe {
            ffi::bppp_reciprocal_verify(0, self.dim_nd, self.dim_np, g.as_ptr(), gv.as_ptr(), self.g_vec.len(), hv.as_ptr(), self.h_vec.len(), gv2.as_ptr(), self.g_vec_.len(),
                                        hv2.as_ptr(), self.h_vec_.len(), point33(commitment).as_ptr(), rec.as_ptr(), cp.r.len(), cp.x.len(), cp.l.len(), cp.n.len(),
                                        label.as_ptr(), label.len(), &mut verdict)
        }, "bppp_reciprocal_verify");
        check_status(verdict, "ReciprocalRangeProofProtocol::verify");
        verdict == ffi::BPPP_ST_TRUE
    }
}
