//! `U64RangeProofProtocol` with the reference's fields, constants and signatures (src/range_proof/u64_proof.rs:12-102).
//! `prove` / `verify` continue the CALLER's transcript through the phase-stepped C ABI; the GPU context (window tables of
//! the 49 generators) is built on first use and cached inside the value.
use std::cell::OnceCell;
use std::os::raw::c_int;

use k256::elliptic_curve::rand_core::{CryptoRng, RngCore};
use k256::{ProjectivePoint, Scalar};
use merlin::Transcript;

use crate::convert::*;
use crate::range_proof::reciprocal::Proof;
use crate::transcript::{app_point33, challenge32};
use crate::{check, check_status, ffi};

#[allow(dead_code)]
const G_VEC_CIRCUIT_SZ: usize = 16;
pub const G_VEC_FULL_SZ: usize = 16;
pub const H_VEC_CIRCUIT_SZ: usize = 26;
pub const H_VEC_FULL_SZ: usize = 32;

pub(crate) struct Ctx(pub *mut ffi::bppp_ctx);
impl Drop for Ctx { fn drop(&mut self) { unsafe { ffi::bppp_ctx_destroy(self.0) } } }

/// Public information for the reciprocal range proof over [0, 2^64) -- field for field the reference's struct.
pub struct U64RangeProofProtocol {
    pub g: ProjectivePoint,
    /// Dimension: `16`
    pub g_vec: Vec<ProjectivePoint>,
    /// Dimension: `26+6=32`
    pub h_vec: Vec<ProjectivePoint>,
    /// CUDA device and table width used when the context is first needed (0 = the library default of 16 bits).
    pub device: i32,
    pub window_bits: i32,
    ctx: OnceCell<Ctx>,
}

impl Clone for U64RangeProofProtocol {
    fn clone(&self) -> Self { Self::new(self.g, self.g_vec.clone(), self.h_vec.clone()).on_device(self.device, self.window_bits) }
}
impl std::fmt::Debug for U64RangeProofProtocol {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        f.debug_struct("U64RangeProofProtocol").field("g", &self.g).field("g_vec", &self.g_vec).field("h_vec", &self.h_vec).finish()
    }
}

impl U64RangeProofProtocol {
    /// Count of digits of u64 in hex representation.
    pub const DIM_ND: usize = 16;
    /// Base (hex)
    pub const DIM_NP: usize = 16;

    pub fn new(g: ProjectivePoint, g_vec: Vec<ProjectivePoint>, h_vec: Vec<ProjectivePoint>) -> Self {
        U64RangeProofProtocol { g, g_vec, h_vec, device: 0, window_bits: 0, ctx: OnceCell::new() }
    }
    pub fn on_device(mut self, device: i32, window_bits: i32) -> Self { self.device = device; self.window_bits = window_bits; self }

    pub(crate) fn ctx(&self) -> *mut ffi::bppp_ctx {
        self.ctx.get_or_init(|| {
            // the reference indexes h_vec[..26], h_vec[26..] and g_vec by position: wrong lengths panic there too
            assert!(self.g_vec.len() == G_VEC_FULL_SZ && self.h_vec.len() == H_VEC_FULL_SZ, "index out of bounds: g_vec needs 16 points, h_vec 32");
            let mut gens = Vec::with_capacity(64 * 49);
            gens.extend_from_slice(&point64(&self.g));
            gens.extend_from_slice(&points64(&self.g_vec));
            gens.extend_from_slice(&points64(&self.h_vec));
            let mut raw = std::ptr::null_mut();
            check(unsafe { ffi::bppp_ctx_create(&mut raw, self.device as c_int, gens.as_ptr(), self.window_bits as c_int, 65536) }, "bppp_ctx_create");
            Ctx(raw)
        }).0
    }

    /// `commitment = x*g + s*h_vec[0]` (u64_proof.rs:37-39)
    pub fn commit_value(&self, x: u64, s: &Scalar) -> ProjectivePoint {
        let mut out = [0u8; 33];
        check(unsafe { ffi::bppp_u64_commit_batch(self.ctx(), 1, &x, scalar32(s).as_ptr(), ffi::BPPP_FMT_COMPRESSED, out.as_mut_ptr()) }, "bppp_u64_commit_batch");
        point_from33(&out)
    }

    /// Verifies that the value committed in `v` lies in [0, 2^64) (u64_proof.rs:42-54), continuing the caller's transcript.
    pub fn verify(&self, v: &ProjectivePoint, proof: Proof, t: &mut Transcript) -> bool {
        let c = self.ctx();
        let rec = proof.to_record();
        let pt = |k: usize| &rec[33 * k..33 * k + 33];
        let v33 = point33(v);
        let mut vp = [0u8; 33];
        check(unsafe { ffi::bppp_u64_verify_begin(c, 1, v33.as_ptr(), rec.as_ptr(), ffi::BPPP_FMT_COMPRESSED, vp.as_mut_ptr()) }, "bppp_u64_verify_begin");
        let mut chal = Vec::with_capacity(192);
        app_point33(b"reciprocal_commitment", &v33, t);                       // reciprocal.rs:99
        chal.extend_from_slice(&challenge32(b"reciprocal_challenge", t));     // :100
        app_point33(b"commitment_cl", pt(0), t); app_point33(b"commitment_cr", pt(1), t);   // circuit.rs:155-159
        app_point33(b"commitment_co", pt(2), t); app_point33(b"commitment_v", &vp, t);
        for l in [&b"circuit_rho"[..], b"circuit_lambda", b"circuit_beta", b"circuit_delta"] { chal.extend_from_slice(&challenge32(leak(l), t)); }
        app_point33(b"commitment_cs", pt(3), t);                              // :189
        chal.extend_from_slice(&challenge32(b"circuit_tau", t));
        let mut com = [0u8; 33];
        check(unsafe { ffi::bppp_u64_verify_circuit(c, chal.as_ptr(), com.as_mut_ptr()) }, "bppp_u64_verify_circuit");
        for j in 0..4usize {                                                  // wnla.rs:88-94
            app_point33(b"wnla_com", &com, t);
            app_point33(b"wnla_x", pt(8 + (3 - j)), t);
            app_point33(b"wnla_r", pt(4 + (3 - j)), t);
            t.append_u64(b"l.sz", (32 >> j) as u64);
            t.append_u64(b"n.sz", (16 >> j) as u64);
            let y = challenge32(b"wnla_challenge", t);
            check(unsafe { ffi::bppp_u64_verify_round(c, j as c_int, y.as_ptr(), com.as_mut_ptr()) }, "bppp_u64_verify_round");
        }
        let mut st = 0i32;
        check(unsafe { ffi::bppp_u64_verify_finish(c, &mut st) }, "bppp_u64_verify_finish");
        check_status(st, "U64RangeProofProtocol::verify");
        st == ffi::BPPP_ST_TRUE
    }

    /// Creates the proof that `x` with blinding `s` lies in [0, 2^64) (u64_proof.rs:57-82).  The 52 `generate_biased` draws
    /// are taken from `rng` up front: their count and order do not depend on the data.
    pub fn prove<R>(&self, x: u64, s: &Scalar, t: &mut Transcript, rng: &mut R) -> Proof
        where R: RngCore + CryptoRng
    {
        let c = self.ctx();
        let mut draws = vec![0u8; ffi::BPPP_U64_RNG_BYTES];
        for k in 0..52 { rng.fill_bytes(&mut draws[64 * k..64 * k + 64]); }
        let mut v = [0u8; 33];
        check(unsafe { ffi::bppp_u64_prove_begin(c, 1, &x, scalar32(s).as_ptr(), draws.as_ptr(), v.as_mut_ptr()) }, "bppp_u64_prove_begin");
        app_point33(b"reciprocal_commitment", &v, t);                         // reciprocal.rs:114
        let e = challenge32(b"reciprocal_challenge", t);
        let mut p4 = [0u8; 132];
        check(unsafe { ffi::bppp_u64_prove_reciprocal(c, e.as_ptr(), p4.as_mut_ptr()) }, "bppp_u64_prove_reciprocal");
        for (k, l) in [&b"commitment_cl"[..], b"commitment_cr", b"commitment_co", b"commitment_v"].iter().enumerate() { app_point33(leak(l), &p4[33 * k..33 * k + 33], t); }
        let mut chal = Vec::with_capacity(128);
        for l in [&b"circuit_rho"[..], b"circuit_lambda", b"circuit_beta", b"circuit_delta"] { chal.extend_from_slice(&challenge32(leak(l), t)); }
        let mut cs = [0u8; 33];
        check(unsafe { ffi::bppp_u64_prove_circuit(c, chal.as_ptr(), cs.as_mut_ptr()) }, "bppp_u64_prove_circuit");
        app_point33(b"commitment_cs", &cs, t);                                // circuit.rs:472
        let tau = challenge32(b"circuit_tau", t);
        let mut p3 = [0u8; 99];
        check(unsafe { ffi::bppp_u64_prove_tau(c, tau.as_ptr(), p3.as_mut_ptr()) }, "bppp_u64_prove_tau");
        for j in 0..4usize {                                                  // wnla.rs:162-168
            app_point33(b"wnla_com", &p3[0..33], t); app_point33(b"wnla_x", &p3[33..66], t); app_point33(b"wnla_r", &p3[66..99], t);
            t.append_u64(b"l.sz", (32 >> j) as u64);
            t.append_u64(b"n.sz", (16 >> j) as u64);
            let y = challenge32(b"wnla_challenge", t);
            check(unsafe { ffi::bppp_u64_prove_round(c, j as c_int, y.as_ptr(), p3.as_mut_ptr()) }, "bppp_u64_prove_round");
        }
        let mut rec = [0u8; ffi::BPPP_U64_PROOF_BYTES];
        let mut st = 0i32;
        check(unsafe { ffi::bppp_u64_prove_finish(c, rec.as_mut_ptr(), &mut st) }, "bppp_u64_prove_finish");
        check_status(st, "U64RangeProofProtocol::prove");
        Proof::from_record(&rec)
    }

    pub fn u64_to_hex(mut x: u64) -> Vec<Scalar> {
        (0..16).map(|_| { let d = x % 16; x /= 16; Scalar::from(d) }).collect()
    }
    pub fn u64_to_hex_mapped(mut x: u64) -> Vec<Scalar> {
        let mut m = [0u64; 16];
        for _ in 0..16 { m[(x % 16) as usize] += 1; x /= 16; }
        m.iter().map(|v| Scalar::from(*v)).collect()
    }
}

// merlin wants `&'static [u8]` labels; the label sets above are literals, this only launders the slice type
fn leak(l: &[u8]) -> &'static [u8] {
    match l {
        b"circuit_rho" => b"circuit_rho", b"circuit_lambda" => b"circuit_lambda", b"circuit_beta" => b"circuit_beta", b"circuit_delta" => b"circuit_delta",
        b"commitment_cl" => b"commitment_cl", b"commitment_cr" => b"commitment_cr", b"commitment_co" => b"commitment_co", b"commitment_v" => b"commitment_v",
        _ => unreachable!(),
    }
}
