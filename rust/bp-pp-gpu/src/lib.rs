//! `bp-pp-gpu`: the public items of distributed-lab/bp-pp (`src/lib.rs:3-6` of the reference: `wnla`, `circuit`,
//! `transcript`, `range_proof`) with identical names, fields, argument order and panics, computed on B200 GPUs by
//! libbppp.so (include/bppp.h).  Plus the batch entry points the north star adds (`prove_batch` / `verify_batch`).
//!
//! Host code only marshals bytes and drives the caller's `merlin::Transcript`; all field, curve and protocol arithmetic
//! runs on the device.  Never compiled in the build image (no Rust toolchain there) -- see Cargo.toml.
#![allow(non_snake_case)]

pub mod circuit;
pub mod range_proof;
pub mod transcript;
pub mod wnla;

pub mod batch;
mod convert;
pub mod ffi;

/// Re-panic where the reference panics; `what` names the call, `code` is a `BPPP_ST_*` / `BPPP_ERR_*` value.
pub(crate) fn check(rc: i32, what: &str) {
    if rc != ffi::BPPP_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::bppp_last_error()) }.to_string_lossy().into_owned();
        panic!("{what} failed with {rc}: {msg}");
    }
}
pub(crate) fn check_status(st: i32, what: &str) {
    match st {
        ffi::BPPP_ST_PANIC_INVERT_ZERO => panic!("{what}: called `Option::unwrap()` on a `None` value (inverse of zero)"),
        ffi::BPPP_ST_PANIC_CHALLENGE_RANGE => panic!("{what}: challenge is not a canonical scalar (from_repr().unwrap())"),
        s if s < 0 => panic!("{what}: malformed input (status {s})"),
        _ => {}
    }
}
