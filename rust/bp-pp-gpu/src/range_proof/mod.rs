//! `range_proof::{reciprocal, u64_proof}` (reference src/range_proof/).
pub mod reciprocal;
pub mod u64_proof;
