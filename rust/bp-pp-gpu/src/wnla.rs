//! `wnla::{WeightNormLinearArgument, Proof, SerializableProof}` (src/wnla.rs:12-190) on the GPU.  `prove` continues the
//! CALLER's transcript: the instance lives on the device as one whole block (bppp_wnla_shard, whole = 1) that is stepped
//! round by round -- shares of X and R out, challenge in.  `verify` with a caller-owned transcript replays the rounds on
//! the host transcript and checks the base case with one MSM (`verify_with_label` uses the engine's own transcript).
use std::ops::{Mul, Sub};

use k256::{AffinePoint, ProjectivePoint, Scalar};
use merlin::Transcript;
use serde::{Deserialize, Serialize};

use crate::convert::*;
use crate::transcript::{app_point, get_challenge};
use crate::{check, check_status, ffi};

#[derive(Clone, Debug)]
pub struct WeightNormLinearArgument {
    pub g: ProjectivePoint,
    pub g_vec: Vec<ProjectivePoint>,
    pub h_vec: Vec<ProjectivePoint>,
    pub c: Vec<Scalar>,
    pub rho: Scalar,
    pub mu: Scalar,
}

#[derive(Clone, Debug)]
pub struct Proof {
    pub r: Vec<ProjectivePoint>,
    pub x: Vec<ProjectivePoint>,
    pub l: Vec<Scalar>,
    pub n: Vec<Scalar>,
}

#[derive(Serialize, Deserialize, Clone, Debug)]
pub struct SerializableProof {
    pub r: Vec<AffinePoint>,
    pub x: Vec<AffinePoint>,
    pub l: Vec<Scalar>,
    pub n: Vec<Scalar>,
}

impl From<&SerializableProof> for Proof {
    fn from(s: &SerializableProof) -> Self {
        Proof { r: s.r.iter().map(ProjectivePoint::from).collect(), x: s.x.iter().map(ProjectivePoint::from).collect(), l: s.l.clone(), n: s.n.clone() }
    }
}
impl From<&Proof> for SerializableProof {
    fn from(p: &Proof) -> Self {
        SerializableProof { r: p.r.iter().map(|v| v.to_affine()).collect(), x: p.x.iter().map(|v| v.to_affine()).collect(), l: p.l.clone(), n: p.n.clone() }
    }
}

struct Shard(*mut ffi::bppp_wnla_shard);
impl Drop for Shard { fn drop(&mut self) { unsafe { ffi::bppp_wnla_shard_destroy(self.0) } } }

impl WeightNormLinearArgument {
    /// `C = v*g + <h_vec, l> + <g_vec, n>`, `v = |n|_mu^2 + <c, l>` (wnla.rs:66-72)
    pub fn commit(&self, l: &[Scalar], n: &[Scalar]) -> ProjectivePoint {
        let mut out = [0u8; 33];
        check(unsafe {
            ffi::bppp_wnla_commit(0, point64(&self.g).as_ptr(), points64(&self.g_vec).as_ptr(), self.g_vec.len(), points64(&self.h_vec).as_ptr(), self.h_vec.len(),
                                  scalars32(&self.c).as_ptr(), self.c.len(), scalar32(&self.rho).as_ptr(), scalar32(&self.mu).as_ptr(), scalars32(l).as_ptr(), l.len(),
                                  scalars32(n).as_ptr(), n.len(), out.as_mut_ptr())
        }, "bppp_wnla_commit");
        point_from33(&out)
    }

    /// wnla.rs:125-190, continuing the caller's transcript.  The reference zero-extends mismatched vectors (util.rs:24-26);
    /// the block form needs |h_vec| = |c| = |l| and |g_vec| = |n|, so shorter operands are padded here the same way.
    pub fn prove(&self, commitment: &ProjectivePoint, t: &mut Transcript, l: Vec<Scalar>, n: Vec<Scalar>) -> Proof {
        let (mut len_l, mut len_n) = (l.len(), n.len());
        let nh = self.h_vec.len().max(self.c.len()).max(len_l);
        let ng = self.g_vec.len().max(len_n);
        let pad_s = |v: &[Scalar], to: usize| { let mut b = scalars32(v); b.resize(32 * to, 0); b };
        let pad_p = |v: &[ProjectivePoint], to: usize| { let mut b = points64(v); b.resize(64 * to, 0); b };
        let mut raw = std::ptr::null_mut();
        check(unsafe {
            ffi::bppp_wnla_shard_create(&mut raw, 0, point64(&self.g).as_ptr(), pad_p(&self.h_vec, nh).as_ptr(), pad_s(&self.c, nh).as_ptr(), pad_s(&l, nh).as_ptr(), nh, 0,
                                        pad_p(&self.g_vec, ng).as_ptr(), pad_s(&n, ng).as_ptr(), ng, 0, scalar32(&self.rho).as_ptr(), scalar32(&self.mu).as_ptr(), 1)
        }, "bppp_wnla_shard_create");
        let shard = Shard(raw);
        let (mut rs, mut xs) = (Vec::new(), Vec::new());
        let mut com = *commitment;
        let mut first = true;
        while len_l + len_n >= 6 {                                            // wnla.rs:126
            let mut xr = [0u8; 128];
            check(unsafe { ffi::bppp_wnla_shard_xr_partial(shard.0, xr.as_mut_ptr(), std::ptr::null_mut()) }, "bppp_wnla_shard_xr_partial");
            let (x, r) = (point_from64(&xr[..64]), point_from64(&xr[64..]));
            app_point(b"wnla_com", &com, t); app_point(b"wnla_x", &x, t); app_point(b"wnla_r", &r, t);      // wnla.rs:162-164
            t.append_u64(b"l.sz", len_l as u64);
            t.append_u64(b"n.sz", len_n as u64);
            let y = get_challenge(b"wnla_challenge", t);
            check(unsafe { ffi::bppp_wnla_shard_fold(shard.0, scalar32(&y).as_ptr(), std::ptr::null_mut()) }, "bppp_wnla_shard_fold");
            com = if first {
                // wnla'.commit(l', n') evaluated literally the first time (wnla.rs:186): it only equals C + yX + (y^2-1)R for a consistent C
                let mut c64 = [0u8; 64];
                check(unsafe { ffi::bppp_wnla_shard_commit_partial(shard.0, c64.as_mut_ptr()) }, "bppp_wnla_shard_commit_partial");
                point_from64(&c64)
            } else {
                com + x.mul(y) + r.mul(y.mul(&y).sub(&Scalar::ONE))           // two scalar multiplications on the host: wnla.rs:100-102
            };
            first = false;
            rs.push(r); xs.push(x);
            len_l = (len_l + 1) / 2; len_n = (len_n + 1) / 2;
        }
        let (mut lb, mut nb) = (vec![0u8; 32 * nh.max(1)], vec![0u8; 32 * ng.max(1)]);
        let (mut cur_h, mut cur_g) = (0usize, 0usize);
        check(unsafe { ffi::bppp_wnla_shard_state(shard.0, &mut cur_h, &mut cur_g, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) },
              "bppp_wnla_shard_state");
        check(unsafe { ffi::bppp_wnla_shard_export(shard.0, std::ptr::null_mut(), std::ptr::null_mut(), lb.as_mut_ptr(), std::ptr::null_mut(), nb.as_mut_ptr()) },
              "bppp_wnla_shard_export");
        rs.reverse(); xs.reverse();                                           // pushed after the recursion returns: innermost first (wnla.rs:186-188)
        Proof { r: rs, x: xs, l: (0..len_l).map(|k| scalar_from32(&lb[32 * k..32 * k + 32])).collect(), n: (0..len_n).map(|k| scalar_from32(&nb[32 * k..32 * k + 32])).collect() }
    }

    /// wnla.rs:75-121 with a fresh `Transcript::new(label)` inside the engine.
    pub fn verify_with_label(&self, commitment: &ProjectivePoint, label: &[u8], proof: Proof) -> bool {
        let mut verdict = 0i32;
        let (r, x) = (proof.r.iter().flat_map(|p| point33(p)).collect::<Vec<u8>>(), proof.x.iter().flat_map(|p| point33(p)).collect::<Vec<u8>>());
        check(unsafe {
            ffi::bppp_wnla_verify(0, point64(&self.g).as_ptr(), points64(&self.g_vec).as_ptr(), self.g_vec.len(), points64(&self.h_vec).as_ptr(), self.h_vec.len(),
                                  scalars32(&self.c).as_ptr(), self.c.len(), scalar32(&self.rho).as_ptr(), scalar32(&self.mu).as_ptr(), point33(commitment).as_ptr(),
                                  r.as_ptr(), proof.r.len(), x.as_ptr(), proof.x.len(), scalars32(&proof.l).as_ptr(), proof.l.len(), scalars32(&proof.n).as_ptr(), proof.n.len(),
                                  label.as_ptr(), label.len(), &mut verdict)
        }, "bppp_wnla_verify");
        check_status(verdict, "WeightNormLinearArgument::verify");
        verdict == ffi::BPPP_ST_TRUE
    }
}

fn point_from64(b: &[u8]) -> ProjectivePoint {
    use k256::elliptic_curve::sec1::FromEncodedPoint;
    if b.iter().all(|v| *v == 0) { return ProjectivePoint::IDENTITY; }
    let enc = k256::EncodedPoint::from_affine_coordinates(k256::FieldBytes::from_slice(&b[..32]), k256::FieldBytes::from_slice(&b[32..]), false);
    ProjectivePoint::from(Option::<AffinePoint>::from(AffinePoint::from_encoded_point(&enc)).expect("engine returned an off-curve point"))
}
