//! The north star's new entry points: `prove_batch` / `verify_batch` / `commit_batch` over N independent proofs, each with
//! its own fresh `Transcript::new(label)` run on the device, on one GPU (`U64RangeProofProtocol`) or on a list of GPUs
//! driven from this one process (`MultiGpu`, bppp_multi_ctx: contiguous per-device ranges, no data-path collective).
use std::os::raw::c_int;

use k256::elliptic_curve::rand_core::{CryptoRng, RngCore};
use k256::{ProjectivePoint, Scalar};

use crate::convert::*;
use crate::range_proof::reciprocal::Proof;
use crate::range_proof::u64_proof::U64RangeProofProtocol;
use crate::{check, ffi};

/// Per-proof outcome of `verify_batch`: the reference's boolean, or why it could not have produced one.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Verdict { True, False, Malformed(i32), WouldPanic(i32) }
impl From<i32> for Verdict {
    fn from(st: i32) -> Self {
        match st { 1 => Verdict::True, 0 => Verdict::False, -1 | -2 => Verdict::WouldPanic(st), s => Verdict::Malformed(s) }
    }
}

fn draw<R: RngCore + CryptoRng>(n: usize, rng: &mut R) -> Vec<u8> {
    let mut b = vec![0u8; ffi::BPPP_U64_RNG_BYTES * n];
    for k in 0..52 * n { rng.fill_bytes(&mut b[64 * k..64 * k + 64]); }       // proof i consumes draws 52 i .. 52 i + 51, in the reference's order
    b
}

impl U64RangeProofProtocol {
    pub fn commit_batch(&self, xs: &[u64], blinds: &[Scalar]) -> Vec<ProjectivePoint> {
        assert_eq!(xs.len(), blinds.len());
        let mut out = vec![0u8; 33 * xs.len()];
        check(unsafe { ffi::bppp_u64_commit_batch(self.ctx(), xs.len(), xs.as_ptr(), scalars32(blinds).as_ptr(), ffi::BPPP_FMT_COMPRESSED, out.as_mut_ptr()) }, "bppp_u64_commit_batch");
        out.chunks(33).map(point_from33).collect()
    }
    pub fn prove_batch<R: RngCore + CryptoRng>(&self, xs: &[u64], blinds: &[Scalar], label: &'static [u8], rng: &mut R) -> Vec<Proof> {
        assert_eq!(xs.len(), blinds.len());
        let n = xs.len();
        let (mut out, mut st) = (vec![0u8; ffi::BPPP_U64_PROOF_BYTES * n], vec![0i32; n]);
        check(unsafe { ffi::bppp_u64_prove_batch(self.ctx(), n, xs.as_ptr(), scalars32(blinds).as_ptr(), draw(n, rng).as_ptr(), label.as_ptr(), label.len(), out.as_mut_ptr(), st.as_mut_ptr()) },
              "bppp_u64_prove_batch");
        st.iter().for_each(|s| crate::check_status(*s, "prove_batch"));
        out.chunks(ffi::BPPP_U64_PROOF_BYTES).map(Proof::from_record).collect()
    }
    pub fn verify_batch(&self, vs: &[ProjectivePoint], proofs: &[Proof], label: &'static [u8]) -> Vec<Verdict> {
        assert_eq!(vs.len(), proofs.len());
        let n = vs.len();
        let commits: Vec<u8> = vs.iter().flat_map(|p| point33(p)).collect();
        let recs: Vec<u8> = proofs.iter().flat_map(|p| p.to_record()).collect();
        let mut st = vec![0i32; n];
        check(unsafe { ffi::bppp_u64_verify_batch(self.ctx(), n, commits.as_ptr(), recs.as_ptr(), ffi::BPPP_FMT_COMPRESSED, label.as_ptr(), label.len(), st.as_mut_ptr()) },
              "bppp_u64_verify_batch");
        st.into_iter().map(Verdict::from).collect()
    }
}

/// One process, several GPUs.
pub struct MultiGpu(*mut ffi::bppp_multi_ctx);
impl Drop for MultiGpu { fn drop(&mut self) { unsafe { ffi::bppp_multi_ctx_destroy(self.0) } } }
impl MultiGpu {
    pub fn new(p: &U64RangeProofProtocol, devices: &[i32], window_bits: i32, max_batch_per_device: usize) -> MultiGpu {
        let mut gens = point64(&p.g).to_vec();
        gens.extend_from_slice(&points64(&p.g_vec));
        gens.extend_from_slice(&points64(&p.h_vec));
        let devs: Vec<c_int> = devices.iter().map(|d| *d as c_int).collect();
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::bppp_multi_ctx_create(&mut raw, devs.as_ptr(), devs.len() as c_int, gens.as_ptr(), window_bits as c_int, max_batch_per_device) }, "bppp_multi_ctx_create");
        MultiGpu(raw)
    }
    pub fn verify_batch(&self, vs: &[ProjectivePoint], proofs: &[Proof], label: &'static [u8]) -> Vec<Verdict> {
        let n = vs.len();
        let commits: Vec<u8> = vs.iter().flat_map(|p| point33(p)).collect();
        let recs: Vec<u8> = proofs.iter().flat_map(|p| p.to_record()).collect();
        let mut st = vec![0i32; n];
        check(unsafe { ffi::bppp_multi_u64_verify_batch(self.0, n, commits.as_ptr(), recs.as_ptr(), ffi::BPPP_FMT_COMPRESSED, label.as_ptr(), label.len(), st.as_mut_ptr()) },
              "bppp_multi_u64_verify_batch");
        st.into_iter().map(Verdict::from).collect()
    }
    pub fn prove_batch<R: RngCore + CryptoRng>(&self, xs: &[u64], blinds: &[Scalar], label: &'static [u8], rng: &mut R) -> Vec<Proof> {
        let n = xs.len();
        let (mut out, mut st) = (vec![0u8; ffi::BPPP_U64_PROOF_BYTES * n], vec![0i32; n]);
        check(unsafe { ffi::bppp_multi_u64_prove_batch(self.0, n, xs.as_ptr(), scalars32(blinds).as_ptr(), draw(n, rng).as_ptr(), label.as_ptr(), label.len(), out.as_mut_ptr(), st.as_mut_ptr()) },
              "bppp_multi_u64_prove_batch");
        st.iter().for_each(|s| crate::check_status(*s, "prove_batch"));
        out.chunks(ffi::BPPP_U64_PROOF_BYTES).map(Proof::from_record).collect()
    }
}
