//! `extern "C"` declarations of include/bppp.h (the drop-in boundary).  Plain pointers and sizes only.
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct bppp_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bppp_multi_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bppp_wnla_shard { _p: [u8; 0] }
#[repr(C)] pub struct bppp_peer { _p: [u8; 0] }

pub const BPPP_OK: c_int = 0;
pub const BPPP_FMT_COMPRESSED: c_int = 0;
pub const BPPP_FMT_AFFINE64: c_int = 1;
pub const BPPP_ST_TRUE: i32 = 1;
pub const BPPP_ST_PANIC_INVERT_ZERO: i32 = -1;
pub const BPPP_ST_PANIC_CHALLENGE_RANGE: i32 = -2;
pub const BPPP_U64_PROOF_BYTES: usize = 525;
pub const BPPP_U64_RNG_BYTES: usize = 3328;

extern "C" {
    pub fn bppp_last_error() -> *const c_char;
    pub fn bppp_ctx_create(out: *mut *mut bppp_ctx, device: c_int, gens64: *const u8, window_bits: c_int, max_batch: usize) -> c_int;
    pub fn bppp_ctx_create_shared(out: *mut *mut bppp_ctx, parent: *const bppp_ctx, max_batch: usize) -> c_int;
    pub fn bppp_ctx_destroy(ctx: *mut bppp_ctx);
    pub fn bppp_u64_commit_batch(ctx: *mut bppp_ctx, n: usize, x: *const u64, blinds32: *const u8, fmt: c_int, out: *mut u8) -> c_int;
    pub fn bppp_u64_verify_batch(ctx: *mut bppp_ctx, n: usize, commits: *const u8, proofs: *const u8, fmt: c_int, label: *const u8, label_len: usize, status: *mut i32) -> c_int;
    pub fn bppp_u64_prove_batch(ctx: *mut bppp_ctx, n: usize, x: *const u64, blinds32: *const u8, rng: *const u8, label: *const u8, label_len: usize,
                                proofs_out: *mut u8, status: *mut i32) -> c_int;
    // phase-stepped variants for a caller-owned transcript
    pub fn bppp_u64_verify_begin(ctx: *mut bppp_ctx, n: usize, commits: *const u8, proofs: *const u8, fmt: c_int, vprime33_out: *mut u8) -> c_int;
    pub fn bppp_u64_verify_circuit(ctx: *mut bppp_ctx, chal: *const u8, com33_out: *mut u8) -> c_int;
    pub fn bppp_u64_verify_round(ctx: *mut bppp_ctx, j: c_int, y32: *const u8, com33_out: *mut u8) -> c_int;
    pub fn bppp_u64_verify_finish(ctx: *mut bppp_ctx, status: *mut i32) -> c_int;
    pub fn bppp_u64_prove_begin(ctx: *mut bppp_ctx, n: usize, x: *const u64, blinds32: *const u8, rng: *const u8, v33_out: *mut u8) -> c_int;
    pub fn bppp_u64_prove_reciprocal(ctx: *mut bppp_ctx, e32: *const u8, pts33_out: *mut u8) -> c_int;
    pub fn bppp_u64_prove_circuit(ctx: *mut bppp_ctx, chal: *const u8, cs33_out: *mut u8) -> c_int;
    pub fn bppp_u64_prove_tau(ctx: *mut bppp_ctx, tau32: *const u8, pts33_out: *mut u8) -> c_int;
    pub fn bppp_u64_prove_round(ctx: *mut bppp_ctx, j: c_int, y32: *const u8, pts33_out: *mut u8) -> c_int;
    pub fn bppp_u64_prove_finish(ctx: *mut bppp_ctx, proofs_out: *mut u8, status: *mut i32) -> c_int;
    pub fn bppp_u64_step_abort(ctx: *mut bppp_ctx);
    // one process, several GPUs
    pub fn bppp_multi_ctx_create(out: *mut *mut bppp_multi_ctx, devices: *const c_int, ndev: c_int, gens64: *const u8, window_bits: c_int, max_batch_per_device: usize) -> c_int;
    pub fn bppp_multi_ctx_destroy(m: *mut bppp_multi_ctx);
    pub fn bppp_multi_u64_commit_batch(m: *mut bppp_multi_ctx, n: usize, x: *const u64, blinds32: *const u8, fmt: c_int, out: *mut u8) -> c_int;
    pub fn bppp_multi_u64_verify_batch(m: *mut bppp_multi_ctx, n: usize, commits: *const u8, proofs: *const u8, fmt: c_int, label: *const u8, label_len: usize, status: *mut i32) -> c_int;
    pub fn bppp_multi_u64_prove_batch(m: *mut bppp_multi_ctx, n: usize, x: *const u64, blinds32: *const u8, rng: *const u8, label: *const u8, label_len: usize,
                                      proofs_out: *mut u8, status: *mut i32) -> c_int;
    // generic single-instance entry points
    pub fn bppp_msm(device: c_int, points: *const u8, points_fmt: c_int, n_points: usize, scalars32: *const u8, n_scalars: usize, out_fmt: c_int, out: *mut u8) -> c_int;
    pub fn bppp_points_sum(device: c_int, points: *const u8, points_fmt: c_int, n: usize, out_fmt: c_int, out: *mut u8) -> c_int;
    pub fn bppp_wnla_commit(device: c_int, g64: *const u8, gvec64: *const u8, gn: usize, hvec64: *const u8, hn: usize, c32: *const u8, cn: usize, rho32: *const u8, mu32: *const u8,
                            l32: *const u8, ln: usize, n32: *const u8, nn: usize, out33: *mut u8) -> c_int;
    pub fn bppp_wnla_verify(device: c_int, g64: *const u8, gvec64: *const u8, gn: usize, hvec64: *const u8, hn: usize, c32: *const u8, cn: usize, rho32: *const u8, mu32: *const u8,
                            commit33: *const u8, r33: *const u8, rn: usize, x33: *const u8, xn: usize, l32: *const u8, ln: usize, n32: *const u8, nn: usize,
                            label: *const u8, label_len: usize, verdict: *mut i32) -> c_int;
    pub fn bppp_wnla_shard_create(out: *mut *mut bppp_wnla_shard, device: c_int, g64: *const u8, hvec64: *const u8, c32: *const u8, l32: *const u8, nh: usize, h_off: usize,
                                  gvec64: *const u8, n32: *const u8, ng: usize, g_off: usize, rho32: *const u8, mu32: *const u8, whole: c_int) -> c_int;
    pub fn bppp_wnla_shard_destroy(s: *mut bppp_wnla_shard);
    pub fn bppp_wnla_shard_state(s: *const bppp_wnla_shard, nh: *mut usize, ng: *mut usize, h_off: *mut usize, g_off: *mut usize, rho32: *mut u8, mu32: *mut u8) -> c_int;
    pub fn bppp_wnla_shard_commit_partial(s: *mut bppp_wnla_shard, out64: *mut u8) -> c_int;
    pub fn bppp_wnla_shard_xr_partial(s: *mut bppp_wnla_shard, out128: *mut u8, device_ms: *mut f32) -> c_int;
    pub fn bppp_wnla_shard_fold(s: *mut bppp_wnla_shard, y32: *const u8, device_ms: *mut f32) -> c_int;
    pub fn bppp_wnla_shard_export(s: *mut bppp_wnla_shard, hvec64: *mut u8, c32: *mut u8, l32: *mut u8, gvec64: *mut u8, n32: *mut u8) -> c_int;
    pub fn bppp_peer_create(out: *mut *mut bppp_peer, device: c_int, world: c_int, rank: c_int, ipc_handle64_out: *mut u8) -> c_int;
    pub fn bppp_peer_connect(p: *mut bppp_peer, handles: *const u8) -> c_int;
    pub fn bppp_peer_destroy(p: *mut bppp_peer);
    pub fn bppp_peer_msm_allsum(p: *mut bppp_peer, points_handle: *const c_void, scalars_handle: *const c_void, n: usize, out_fmt: c_int, out: *mut u8, elapsed_ms: *mut f32) -> c_int;
    pub fn bppp_peer_allgather(p: *mut bppp_peer, input: *const u8, bytes: usize, out: *mut u8) -> c_int;
    pub fn bppp_reciprocal_prove(device: c_int, dim_nd: usize, dim_np: usize, g64: *const u8, gvec64: *const u8, gn: usize, hvec64: *const u8, hn: usize,
                                 gvec2_64: *const u8, gn2: usize, hvec2_64: *const u8, hn2: usize, x32: *const u8, s32: *const u8, digits: *const u32,
                                 rng: *const u8, rng_len: usize, label: *const u8, label_len: usize, out: *mut u8, out_cap: usize, rounds_out: *mut usize,
                                 l_len_out: *mut usize, n_len_out: *mut usize, commit33_out: *mut u8, status: *mut i32) -> c_int;
    pub fn bppp_reciprocal_verify(device: c_int, dim_nd: usize, dim_np: usize, g64: *const u8, gvec64: *const u8, gn: usize, hvec64: *const u8, hn: usize,
                                  gvec2_64: *const u8, gn2: usize, hvec2_64: *const u8, hn2: usize, commit33: *const u8, rec: *const u8, rounds_r: usize,
                                  rounds_x: usize, l_len: usize, n_len: usize, label: *const u8, label_len: usize, verdict: *mut i32) -> c_int;
}
#[allow(dead_code)]
pub(crate) type Opaque = c_void;
