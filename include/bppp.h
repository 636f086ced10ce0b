/*
 * bppp.h -- C ABI of the B200-native Bulletproofs++ engine (libbppp.so).
 *
 * This is the drop-in boundary for the hot path of distributed-lab/bp-pp: every entry point is what
 * a Rust shim for the reference's public API would bind over `extern "C"`.
 * The reference has no FFI of its own; each function cites the reference item it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   scalars   32 bytes big-endian, canonical (< n)  -- k256::Scalar::to_bytes / from_repr
 *   points    BPPP_FMT_COMPRESSED: 33-byte SEC1 compressed (identity = 33 zero bytes), the form
 *             GroupEncoding::to_bytes yields (src/transcript.rs:7) and serde uses for AffinePoint;
 *             BPPP_FMT_AFFINE64: 64 bytes x||y big-endian (identity = 64 zero bytes), what a shim holding
 *             k256::AffinePoint passes without a square root on the device.
 *   u64 proof one record in reciprocal::SerializableProof field order
 *             (src/range_proof/reciprocal.rs:37-41, src/circuit.rs:37-46):
 *               c_l c_r c_o c_s | r[0..4) | x[0..4) | l[0..2) | n[0..1) | r
 *             r[]/x[] in the reference's push order, innermost WNLA round first (src/wnla.rs:186-188).
 *             525 bytes compressed, 928 bytes with 64-byte points.
 *   rng       the reference is generic over RngCore (src/circuit.rs:260-263); the ABI takes the bytes the
 *             RNG would have produced: 52 draws x 64 bytes per u64 proof, in draw order,
 *             each draw reduced as Scalar::generate_biased does (big-endian 512 bit mod n).
 *   status    per proof, int32: 1 = verify true / proved, 0 = verify false, < 0 = BPPP_ST_* (the
 *             reference would have panicked or failed to deserialise).  Functions return 0 or BPPP_ERR_*
 *             and never unwind.
 *   buffers   caller-owned.  `_dev` variants take device pointers valid on the context's GPU and a
 *             CUDA stream handle (cudaStream_t as void*; NULL = the legacy default stream), enqueue work and return
 *             without synchronising; internally the batch fans out over sub-streams that fork from / join into it.
 *   threading entry points taking a context serialise on it (internal mutex): one workspace, one staging area.  For
 *             concurrency use one context per host thread (bppp_ctx_create_shared shares the tables).  `_dev` calls only
 *             enqueue: two of them on the SAME context must be ordered on one stream (they share the workspace).
 */
#ifndef BPPP_H
#define BPPP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bppp_ctx bppp_ctx;

enum { BPPP_FMT_COMPRESSED = 0, BPPP_FMT_AFFINE64 = 1 };

enum {
    BPPP_ST_FALSE = 0,
    BPPP_ST_TRUE = 1,
    BPPP_ST_PANIC_INVERT_ZERO = -1,     /* reference: Scalar::invert().unwrap() on zero */
    BPPP_ST_PANIC_CHALLENGE_RANGE = -2, /* reference: from_repr().unwrap(), src/transcript.rs:13 */
    BPPP_ST_BAD_POINT = -3,             /* not a curve point: the reference fails at deserialisation */
    BPPP_ST_BAD_SCALAR = -4,            /* >= n: not representable as k256::Scalar */
    BPPP_ST_BAD_ARG = -5
};

enum {
    BPPP_OK = 0,
    BPPP_ERR_ARG = -10,
    BPPP_ERR_NO_DEVICE = -11,  /* no CUDA device: there is deliberately no CPU fallback */
    BPPP_ERR_CUDA = -12,
    BPPP_ERR_GENERATOR = -13,  /* a generator is not a curve point */
    BPPP_ERR_NOMEM = -14,
    BPPP_ERR_ENCODING = -15    /* an input point is not on the curve / a scalar is >= n (verify entry points turn this into a
                                  BPPP_ST_BAD_* verdict; infrastructure failures -- NOMEM, CUDA -- are returned as such) */
};

#define BPPP_U64_NUM_GENS 49        /* g, g_vec[16], h_vec[32]: src/range_proof/u64_proof.rs:12-14,19-28 */
#define BPPP_U64_PROOF_BYTES 525    /* 13 points + 3 scalars, README.md:30-34 */
#define BPPP_U64_PROOF_BYTES_AFFINE 928
#define BPPP_U64_RNG_BYTES 3328     /* 52 draws x 64 bytes */

/* Context for U64RangeProofProtocol { g, g_vec[16], h_vec[32] } (src/range_proof/u64_proof.rs:19-28) on
 * CUDA device `device`.  gens64 = g || g_vec || h_vec as 49 x 64-byte affine points.  Builds the
 * fixed-base window tables on the device and a workspace for `max_batch` proofs per launch sequence (larger
 * batches are processed in slices).  window_bits: 0 = default 16 (3.3 GB of tables); 2..20 = unsigned windows
 * (20: 42.7 GB); 21..23 = signed windows with 2^(W-1) entries each (22: 78.9 GB); a negative value requests
 * signed windows of |window_bits| bits at any width. */
int bppp_ctx_create(bppp_ctx **out, int device, const uint8_t *gens64, int window_bits, size_t max_batch);
/* A further context on the same GPU sharing `parent`'s window tables (read-only after construction) with its own workspace,
 * staging buffers and streams (max_batch 0 = the parent's).  Independent batches submitted through different contexts run
 * side by side, which keeps the GPU full when each batch is small (a rank's 1/8 share of a 65,536 batch).  The parent
 * must outlive every context that shares its tables. */
int bppp_ctx_create_shared(bppp_ctx **out, const bppp_ctx *parent, size_t max_batch);
/* Hint: the caller keeps `batches` independent batches in flight on this GPU (sibling contexts).  Only affects how many
 * lanes per proof the kernels use for small batches (results are identical for every choice). */
int bppp_ctx_set_inflight(bppp_ctx *ctx, int batches);
void bppp_ctx_destroy(bppp_ctx *ctx);
const char *bppp_last_error(void);
/* bytes of device memory held by the context (tables + workspace), build time of the tables in ms */
int bppp_ctx_info(const bppp_ctx *ctx, size_t *table_bytes, size_t *workspace_bytes, double *table_build_ms, int *window_bits);

/* U64RangeProofProtocol::commit_value (src/range_proof/u64_proof.rs:37-39): out[i] = x[i]*g + s[i]*h_vec[0] */
int bppp_u64_commit_batch(bppp_ctx *ctx, size_t n, const uint64_t *x, const uint8_t *blinds32, int fmt, uint8_t *out);

/* U64RangeProofProtocol::verify (src/range_proof/u64_proof.rs:42-54) over n independent proofs, each with
 * a fresh merlin::Transcript::new(label).  commits: n points, proofs: n records (format `fmt`). */
int bppp_u64_verify_batch(bppp_ctx *ctx, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt,
                          const uint8_t *label, size_t label_len, int32_t *status);
int bppp_u64_verify_batch_dev(bppp_ctx *ctx, size_t n, const void *d_commits, const void *d_proofs, int fmt,
                              const uint8_t *label, size_t label_len, void *d_status, void *stream);

/* U64RangeProofProtocol::prove (src/range_proof/u64_proof.rs:57-82) over n independent witnesses.
 * rng: n x 3328 bytes.  proofs_out: n x 525-byte records (always compressed). */
int bppp_u64_prove_batch(bppp_ctx *ctx, size_t n, const uint64_t *x, const uint8_t *blinds32, const uint8_t *rng,
                         const uint8_t *label, size_t label_len, uint8_t *proofs_out, int32_t *status);
int bppp_u64_prove_batch_dev(bppp_ctx *ctx, size_t n, const void *d_x, const void *d_blinds32, const void *d_rng,
                             const uint8_t *label, size_t label_len, void *d_proofs_out, void *d_status, void *stream);

/* ---- phase-stepped variants for a CALLER-OWNED transcript -------------------------------------------------------------
 * The reference's single-instance signatures take `t: &mut merlin::Transcript` in whatever state the caller left it
 * (src/range_proof/u64_proof.rs:42,57; src/range_proof/reciprocal.rs:98,110; src/circuit.rs:154,260; src/wnla.rs:75,125)
 * and merlin keeps its STROBE state private, so a shim cannot hand the transcript to the device.  These entry points cut
 * the same kernel sequences at the transcript's challenge points: every step RETURNS the
 * 33-byte compressed points the host must `app_point` next (those it does not already hold) and TAKES the challenges
 * it then drew with `get_challenge` (32 bytes big-endian each).  n independent instances advance together (n = 1 for
 * the reference's signatures); all per-step arrays are proof-major.  One stepped session per context at a time, calls
 * in the order listed, synchronous; bppp_u64_step_abort drops a half-finished session.
 *
 * verify: host appends reciprocal_commitment V -> e; commitment_cl/cr/co from the proof and commitment_v = V' (returned by
 *   _begin) -> rho, lambda, beta, delta; commitment_cs -> tau; then per WNLA round j: wnla_com (returned), wnla_x, wnla_r
 *   from the proof (x[3-j], r[3-j]), l.sz = 32 >> j, n.sz = 16 >> j -> y_j. */
int bppp_u64_verify_begin(bppp_ctx *ctx, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt, uint8_t *vprime33_out);
int bppp_u64_verify_circuit(bppp_ctx *ctx, const uint8_t *chal /* n x 6 x 32: e rho lambda beta delta tau */, uint8_t *com33_out /* n x 33: wnla_com of round 0 */);
int bppp_u64_verify_round(bppp_ctx *ctx, int j, const uint8_t *y32 /* n x 32 */, uint8_t *com33_out /* j < 3: wnla_com of round j + 1; j == 3: unused */);
int bppp_u64_verify_finish(bppp_ctx *ctx, int32_t *status);
/* prove: _begin returns V (reciprocal_commitment) -> e; _reciprocal returns commitment_cl, cr, co, v -> rho, lambda, beta,
 *   delta; _circuit returns commitment_cs -> tau; _tau returns wnla_com, wnla_x, wnla_r of round 0 -> y_0; _round(j) returns
 *   those of round j + 1 (nothing for j == 3); _finish writes the 525-byte records.  rng as in bppp_u64_prove_batch. */
int bppp_u64_prove_begin(bppp_ctx *ctx, size_t n, const uint64_t *x, const uint8_t *blinds32, const uint8_t *rng, uint8_t *v33_out);
int bppp_u64_prove_reciprocal(bppp_ctx *ctx, const uint8_t *e32 /* n x 32 */, uint8_t *pts33_out /* n x 4 x 33 */);
int bppp_u64_prove_circuit(bppp_ctx *ctx, const uint8_t *chal /* n x 4 x 32: rho lambda beta delta */, uint8_t *cs33_out /* n x 33 */);
int bppp_u64_prove_tau(bppp_ctx *ctx, const uint8_t *tau32 /* n x 32 */, uint8_t *pts33_out /* n x 3 x 33: com X_0 R_0 */);
int bppp_u64_prove_round(bppp_ctx *ctx, int j, const uint8_t *y32 /* n x 32 */, uint8_t *pts33_out /* j < 3: n x 3 x 33: com X_{j+1} R_{j+1} */);
int bppp_u64_prove_finish(bppp_ctx *ctx, uint8_t *proofs_out, int32_t *status);
void bppp_u64_step_abort(bppp_ctx *ctx);

/* number of kernels launched by this context since creation (for the bench's gpu_launches claim) */
uint64_t bppp_launch_count(const bppp_ctx *ctx);

/* ---- one process, several GPUs ---------------------------------------------------------------------------------------
 * A bppp_multi_ctx owns one context per listed CUDA device (generators and tables replicated); each batch call
 * cuts [0, n) into contiguous per-device ranges, one host thread per device, no data-path collective.  The result for
 * proof i is independent of the device list.  Arguments as for the single-device entry points (host buffers). */
typedef struct bppp_multi_ctx bppp_multi_ctx;
int bppp_multi_ctx_create(bppp_multi_ctx **out, const int *devices, int ndev, const uint8_t *gens64, int window_bits, size_t max_batch_per_device);
void bppp_multi_ctx_destroy(bppp_multi_ctx *m);
int bppp_multi_device_count(const bppp_multi_ctx *m);
bppp_ctx *bppp_multi_ctx_get(const bppp_multi_ctx *m, int k);      /* the k-th device's context (owned by m) */
int bppp_multi_u64_commit_batch(bppp_multi_ctx *m, size_t n, const uint64_t *x, const uint8_t *blinds32, int fmt, uint8_t *out);
int bppp_multi_u64_verify_batch(bppp_multi_ctx *m, size_t n, const uint8_t *commits, const uint8_t *proofs, int fmt,
                                const uint8_t *label, size_t label_len, int32_t *status);
int bppp_multi_u64_prove_batch(bppp_multi_ctx *m, size_t n, const uint64_t *x, const uint8_t *blinds32, const uint8_t *rng,
                               const uint8_t *label, size_t label_len, uint8_t *proofs_out, int32_t *status);

/* ---- generic (arbitrary-size, single-instance) entry points ------------------------------------------------ */

/* util::vector_mul<ProjectivePoint> (src/util.rs:46-60): out = sum_i scalars[i] * points[i], the shorter operand
 * zero-extended as the reference does.  Pippenger on the device for n > 1024.  points_fmt / out_fmt: BPPP_FMT_*. */
int bppp_msm(int device, const uint8_t *points, int points_fmt, size_t n_points, const uint8_t *scalars32, size_t n_scalars,
             int out_fmt, uint8_t *out);
/* The same with operands decoded and resident in HBM (handles from *_upload); *elapsed_ms = device time of the MSM. */
int bppp_points_upload(int device, const uint8_t *points, int points_fmt, size_t n, void **handle);
int bppp_scalars_upload(int device, const uint8_t *scalars32, size_t n, void **handle);
void bppp_device_free(int device, void *handle);
int bppp_msm_uploaded(int device, const void *points_handle, const void *scalars_handle, size_t n, int out_fmt, uint8_t *out,
                      float *elapsed_ms);
/* SEC1 compressed <-> 64-byte affine for an array of points: the SerializableProof <-> Proof conversions
 * (src/wnla.rs:41-61, src/circuit.rs:48-76, src/range_proof/reciprocal.rs:43-59); fails if a point is not on the curve */
int bppp_points_convert(int device, const uint8_t *in, int in_fmt, size_t n, int out_fmt, uint8_t *out);
/* synthetic generator vectors for large-n measurements: out[i] = base + i * step, 64-byte affine, computed on the device */
int bppp_points_generate(int device, const uint8_t *base64, const uint8_t *step64, size_t n, uint8_t *out64);
/* sum of n points: combines the per-rank partial sums of an MSM split by point range across GPUs */
int bppp_points_sum(int device, const uint8_t *points, int points_fmt, size_t n, int out_fmt, uint8_t *out);

/* WeightNormLinearArgument { g, g_vec, h_vec, c, rho, mu } (src/wnla.rs:12-19) for arbitrary lengths on one GPU.
 * Points 64-byte affine, scalars 32-byte big-endian, commitments / proof points 33-byte compressed.  Mismatched
 * lengths are zero-extended exactly as the reference does (src/util.rs:24-26).  Fresh Transcript::new(label).
 *   commit  src/wnla.rs:66-72      prove  src/wnla.rs:125-190      verify  src/wnla.rs:75-121
 * prove: r_out / x_out need 33 * rounds bytes (rounds <= 64), innermost round first (src/wnla.rs:186-188); l_out / n_out
 * need 32 * ln / 32 * nn bytes.  *status / *verdict: 1, 0 (verify false) or a BPPP_ST_* panic / malformed code. */
int bppp_wnla_commit(int device, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn,
                     const uint8_t *c32, size_t cn, const uint8_t *rho32, const uint8_t *mu32, const uint8_t *l32, size_t ln,
                     const uint8_t *n32, size_t nn, uint8_t *out33);
int bppp_wnla_prove(int device, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn,
                    const uint8_t *c32, size_t cn, const uint8_t *rho32, const uint8_t *mu32, const uint8_t *commit33,
                    const uint8_t *l32, size_t ln, const uint8_t *n32, size_t nn, const uint8_t *label, size_t label_len,
                    uint8_t *r_out, uint8_t *x_out, size_t *rounds_out, uint8_t *l_out, size_t *l_out_len, uint8_t *n_out,
                    size_t *n_out_len, int32_t *status);
int bppp_wnla_verify(int device, const uint8_t *g64, const uint8_t *gvec64, size_t gn, const uint8_t *hvec64, size_t hn,
                     const uint8_t *c32, size_t cn, const uint8_t *rho32, const uint8_t *mu32, const uint8_t *commit33,
                     const uint8_t *r33, size_t rn, const uint8_t *x33, size_t xn, const uint8_t *l32, size_t ln,
                     const uint8_t *n32, size_t nn, const uint8_t *label, size_t label_len, int32_t *verdict);

/* ---- a standalone WNLA instance cut into blocks, one per GPU (BASELINE config 5) ---------------------------
 * A shard holds one contiguous block of the instance resident on one GPU: h_vec / c / l indices [h_off, h_off + nh) and
 * g_vec / n indices [g_off, g_off + ng), offsets even.  Folding maps the pair (2i, 2i+1) to i (src/util.rs:7-22), so an
 * even-offset, even-length block folds locally.  Per round every block yields its shares of X and R (src/wnla.rs:143-160;
 * the block's part of vx / vr rides on g), the holders add the shares -- the only exchange, 2 x 64 bytes per block --
 * run the identical transcript and fold with the challenge.  When the blocks become too short to fold locally their
 * contents are gathered (_export) into one shard created with whole = 1, which finishes the recursion; a whole shard is
 * also the stepped single-GPU prover for a caller-owned transcript.  Drivers: bp_pp_b200/shard.py (an
 * NCCL all-gather between one-GPU processes, host threads between the GPUs of one process). */
typedef struct bppp_wnla_shard bppp_wnla_shard;
int bppp_wnla_shard_create(bppp_wnla_shard **out, int device, const uint8_t *g64, const uint8_t *hvec64, const uint8_t *c32, const uint8_t *l32,
                           size_t nh, size_t h_off, const uint8_t *gvec64, const uint8_t *n32, size_t ng, size_t g_off, const uint8_t *rho32,
                           const uint8_t *mu32, int whole);
void bppp_wnla_shard_destroy(bppp_wnla_shard *s);
int bppp_wnla_shard_state(const bppp_wnla_shard *s, size_t *nh, size_t *ng, size_t *h_off, size_t *g_off, uint8_t *rho32, uint8_t *mu32);
int bppp_wnla_shard_commit_partial(bppp_wnla_shard *s, uint8_t *out64);                 /* share of wnla.commit(l, n), src/wnla.rs:66-72 */
int bppp_wnla_shard_xr_partial(bppp_wnla_shard *s, uint8_t *out128, float *device_ms);  /* shares of X and R, 64-byte affine each */
int bppp_wnla_shard_fold(bppp_wnla_shard *s, const uint8_t *y32, float *device_ms);     /* src/wnla.rs:170-184 on the block */
int bppp_wnla_shard_export(bppp_wnla_shard *s, uint8_t *hvec64, uint8_t *c32, uint8_t *l32, uint8_t *gvec64, uint8_t *n32);

/* ---- peer exchange over NVLink / NVSwitch (one process per GPU) ------------------------------------------------------
 * The two exchange steps of the path -- the partial sums of a point-range-split util::vector_mul (src/util.rs:46-60) and
 * the per-round shares of X and R of a block-sharded WeightNormLinearArgument::prove (src/wnla.rs:152-160) -- done by the
 * library's own kernels: every rank owns a mailbox in its HBM, a producer stores its payload straight into its slot of every
 * peer's mailbox (remote stores), fences and publishes an epoch flag; a consumer kernel spins on its local flags and
 * reduces / copies the slots.  _create returns the mailbox's 64-byte CUDA IPC handle; the caller gathers the handles of
 * all ranks once (any transport) and passes them, in rank order, to _connect.  The collective calls (_msm_allsum,
 * _allgather) must be made by every rank in the same order.  world <= 16.  A rank that never arrives ends the wait with
 * BPPP_ERR_CUDA after 20 s. */
typedef struct bppp_peer bppp_peer;
int bppp_peer_create(bppp_peer **out, int device, int world, int rank, uint8_t *ipc_handle64_out);
int bppp_peer_connect(bppp_peer *p, const uint8_t *handles /* world x 64 bytes */);
void bppp_peer_destroy(bppp_peer *p);
int bppp_peer_world(const bppp_peer *p);
int bppp_peer_rank(const bppp_peer *p);
/* sum over all ranks of each rank's block MSM (handles from bppp_points_upload / bppp_scalars_upload): Pippenger, the remote
 * stores of the partial sum and the reduction of the world's partial sums back to back on one stream; every rank receives
 * the same point.  *elapsed_ms (optional): device time from the first MSM kernel to the reduced sum. */
int bppp_peer_msm_allsum(bppp_peer *p, const void *points_handle, const void *scalars_handle, size_t n, int out_fmt, uint8_t *out, float *elapsed_ms);
/* all-gather of one short byte string per rank (4..240 bytes, a multiple of 4): out = world x bytes in rank order */
int bppp_peer_allgather(bppp_peer *p, const uint8_t *in, size_t bytes, uint8_t *out);

/* ArithmeticCircuit<P> (src/circuit.rs:95-139) with dense row-major W_m (dim_nm x dim_nw) and W_l (dim_nl x dim_nw),
 * dim_nl = dim_nv * k, dim_nw = 2 dim_nm + dim_no, and the partition closure tabulated: part_xx[j] = index or -1 for
 * j < part_n, None beyond (PartitionType LO / LL / LR / NO, src/circuit.rs:15-20); entries must lie in [-1, dim_no).  Generators 64-byte affine. */
typedef struct bppp_circuit_desc {
    size_t dim_nm, dim_no, k, dim_nv;
    int f_l, f_m;
    const uint8_t *g64, *gvec64, *hvec64, *gvec2_64, *hvec2_64;   /* g, g_vec, h_vec, g_vec_, h_vec_ */
    size_t gn, hn, gn2, hn2;
    const uint8_t *W_m32, *W_l32, *a_m32, *a_l32;
    const int32_t *part_lo, *part_ll, *part_lr, *part_no;
    size_t part_n;
} bppp_circuit_desc;
/* ArithmeticCircuit::commit / prove / verify (src/circuit.rs:146-151, 260-556, 154-256); fresh Transcript::new(label).
 * prove: witness v (k x dim_nv), s_v (k), w_l, w_r (dim_nm), w_o (dim_no); rng = (18 + dim_nv + dim_nm) x 64 bytes in draw
 * order; out record = c_l c_r c_o c_s | r[rounds] | x[rounds] | l | n.  A descriptor the reference would panic on (index out
 * of bounds: too few generators, partition entries outside [-1, dim_no)) is rejected with BPPP_ERR_ARG. */
int bppp_circuit_commit(int device, const bppp_circuit_desc *desc, const uint8_t *v32, const uint8_t *s32, uint8_t *out33);
int bppp_circuit_prove(int device, const bppp_circuit_desc *desc, const uint8_t *commits33, const uint8_t *v32, const uint8_t *sv32,
                       const uint8_t *wl32, const uint8_t *wr32, const uint8_t *wo32, const uint8_t *rng, size_t rng_len,
                       const uint8_t *label, size_t label_len, uint8_t *out, size_t out_cap, size_t *rounds_out,
                       size_t *l_len_out, size_t *n_len_out, int32_t *status);
int bppp_circuit_verify(int device, const bppp_circuit_desc *desc, const uint8_t *commits33, const uint8_t *rec, size_t rounds_r,
                        size_t rounds_x, size_t l_len, size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict);

/* The same circuit with W_m and W_l in compressed-sparse-row form (what a circuit compiler emits; the dense descriptor needs
 * O(dim^2) memory).  Values either one per non-zero (value_idx NULL, values32 = nnz x 32 bytes) or through a dictionary of
 * n_values distinct scalars (values32 = n_values x 32 bytes, value_idx[k] selects the k-th non-zero's value).  The engine
 * keeps the matrices on the device (compressed sparse columns) and evaluates lambda_vec x W_l, mu_vec x W_m -- the only
 * super-linear step of circuit.rs:584-653 -- in a kernel. */
typedef struct bppp_sparse_matrix {
    size_t rows, cols, nnz;
    const uint64_t *row_ptr;      /* rows + 1 entries, row_ptr[rows] = nnz */
    const uint32_t *col_idx;      /* nnz */
    const uint32_t *value_idx;    /* nnz, or NULL */
    const uint8_t *values32;
    size_t n_values;
} bppp_sparse_matrix;
typedef struct bppp_circuit_desc_sparse {
    size_t dim_nm, dim_no, k, dim_nv;
    int f_l, f_m;
    const uint8_t *g64, *gvec64, *hvec64, *gvec2_64, *hvec2_64;
    size_t gn, hn, gn2, hn2;
    bppp_sparse_matrix W_m, W_l;   /* dim_nm x dim_nw, dim_nl x dim_nw */
    const uint8_t *a_m32, *a_l32;
    const int32_t *part_lo, *part_ll, *part_lr, *part_no;
    size_t part_n;
} bppp_circuit_desc_sparse;
int bppp_circuit_commit_sparse(int device, const bppp_circuit_desc_sparse *desc, const uint8_t *v32, const uint8_t *s32, uint8_t *out33);
int bppp_circuit_prove_sparse(int device, const bppp_circuit_desc_sparse *desc, const uint8_t *commits33, const uint8_t *v32, const uint8_t *sv32,
                              const uint8_t *wl32, const uint8_t *wr32, const uint8_t *wo32, const uint8_t *rng, size_t rng_len,
                              const uint8_t *label, size_t label_len, uint8_t *out, size_t out_cap, size_t *rounds_out,
                              size_t *l_len_out, size_t *n_len_out, int32_t *status);
int bppp_circuit_verify_sparse(int device, const bppp_circuit_desc_sparse *desc, const uint8_t *commits33, const uint8_t *rec, size_t rounds_r,
                               size_t rounds_x, size_t l_len, size_t n_len, const uint8_t *label, size_t label_len, int32_t *verdict);

/* ReciprocalRangeProofProtocol { dim_nd, dim_np, g, g_vec, h_vec, g_vec_, h_vec_ } (src/range_proof/reciprocal.rs:64-84)
 * for arbitrary dimensions: commit_value (:88-90), prove (:110-146), verify (:98-107); make_circuit (:150-214) is built
 * internally from the challenge.  digits: dim_nd values < dim_np.  rng = (1 + 18 + (dim_nd + 1) + dim_nd) x 64 bytes.
 * out record = c_l c_r c_o c_s | r[rounds] | x[rounds] | l | n | r. */
int bppp_reciprocal_commit_value(int device, const uint8_t *g64, const uint8_t *h0_64, const uint8_t *x32, const uint8_t *s32, uint8_t *out33);
int bppp_reciprocal_prove(int device, size_t dim_nd, size_t dim_np, const uint8_t *g64, const uint8_t *gvec64, size_t gn,
                          const uint8_t *hvec64, size_t hn, const uint8_t *gvec2_64, size_t gn2, const uint8_t *hvec2_64, size_t hn2,
                          const uint8_t *x32, const uint8_t *s32, const uint32_t *digits, const uint8_t *rng, size_t rng_len,
                          const uint8_t *label, size_t label_len, uint8_t *out, size_t out_cap, size_t *rounds_out,
                          size_t *l_len_out, size_t *n_len_out, uint8_t *commit33_out, int32_t *status);
int bppp_reciprocal_verify(int device, size_t dim_nd, size_t dim_np, const uint8_t *g64, const uint8_t *gvec64, size_t gn,
                           const uint8_t *hvec64, size_t hn, const uint8_t *gvec2_64, size_t gn2, const uint8_t *hvec2_64, size_t hn2,
                           const uint8_t *commit33, const uint8_t *rec, size_t rounds_r, size_t rounds_x, size_t l_len, size_t n_len,
                           const uint8_t *label, size_t label_len, int32_t *verdict);

/* Per-kernel device timing of everything launched between begin and end (CUDA events on the launching
 * stream).  end() synchronises and fills up to n_max (name[48], total ms, launch count) triples. */
int bppp_ctx_profile_begin(bppp_ctx *ctx);
int bppp_ctx_profile_end(bppp_ctx *ctx, char *names48, double *total_ms, uint32_t *counts, int n_max, int *n_out);

/* Integer-pipe microbenchmarks on `device`: fills out[0..8) with
 *   [0] IMAD.WIDE.U32 multiply-accumulates/s   [1] fe_mul/s   [2] fe_sqr/s   [3] sc_mul/s
 *   [4] mixed point additions/s   [5] point doublings/s   [6] full point additions/s   [7] SM clock MHz seen
 * Used to state the integer roofline the path is bound by. */
int bppp_microbench(int device, double *out, int n_out);   /* n_out >= 10 also fills [8] IMAD/s, [9] IADD/s; n_out >= 14: [10] fe_inv/s (safegcd), [11] fe_inv/s by a^(p-2), [12] sc_inv/s (safegcd), [13] sc_inv/s by a^(n-2) */

#ifdef __cplusplus
}
#endif
#endif
