/* xhe.h -- C ABI of libxhe_cuda.so: the B200 (sm_100a) batch verifier for XELIS-HE confidential transactions.
 *
 * This is the drop-in boundary for the reference's batch-verification hot path.  The reference (pure Rust) has no
 * FFI; each entry point below names the reference call site it replaces (paths relative to the reference repo).
 * Conventions: int32 return, 0 = XHE_OK; >0 = verification verdicts mirroring ProofVerificationError /
 * VerificationError (src/lib.rs:71-89, src/tx/verify.rs:16-21); <0 = infrastructure failure (bad argument, CUDA
 * error; text via xhe_last_error).  No exceptions cross the ABI.  Points cross as 32-byte canonical ristretto255
 * encodings, scalars as 32-byte canonical little-endian.  A ctx is bound to one device and one host thread at a time.
 * There is NO CPU fallback: every compute entry point launches sm_100a kernels or fails with XHE_E_CUDA. */
#ifndef XHE_H
#define XHE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  XHE_OK = 0,
  XHE_ERR_SIGNATURE = 1,            /* ProofVerificationError::Signature               src/lib.rs:73 */
  XHE_ERR_DECOMPRESSION = 2,        /* ::Decompression                                 src/lib.rs:75 */
  XHE_ERR_COMMITMENT_EQ_PROOF = 3,  /* ::CommitmentEqProof                             src/lib.rs:77 */
  XHE_ERR_CT_VALIDITY_PROOF = 4,    /* ::CiphertextValidityProof                       src/lib.rs:79 */
  XHE_ERR_GENERIC_PROOF = 5,        /* ::GenericProof (sigma MSM != identity)          src/lib.rs:81, src/tx/verify.rs:500-502 */
  XHE_ERR_RANGE_PROOF = 6,          /* ::RangeProof(..)                                src/lib.rs:83, src/tx/verify.rs:504-514 */
  XHE_ERR_TRANSCRIPT = 7,           /* ::Transcript(IdentityPoint)                     src/lib.rs:85, src/transcript.rs:73-84 */
  XHE_ERR_FORMAT = 8,               /* ::Format                                        src/lib.rs:87 */
  XHE_ERR_INVALID_NONCE = 9,        /* VerificationError::InvalidNonce                 src/tx/verify.rs:19 */
  XHE_ERR_STATE = 10,               /* VerificationError::State(..)                    src/tx/verify.rs:18 */
  XHE_ERR_PARSE = 11,               /* wire-format / serde-level rejection (before verify is reachable) */
  XHE_E_ARG = -1, XHE_E_CUDA = -2, XHE_E_NOMEM = -3, XHE_E_NCCL = -4,
  XHE_E_CAPACITY = -5               /* a transaction's range proof needs more parties than the context was created for
                                       (party_capacity < m <= 512); nothing was verified or applied, *fail_index names the tx */
};

typedef struct xhe_ctx xhe_ctx;
struct xhe_ledger;

/* Replaces the lazy_statics H, BP_GENS = BulletproofGens::new(64, 512), PC_GENS (src/elgamal.rs:16-24,
 * src/proofs.rs:19-22): builds G/H tables and the 2*64*party_capacity generator table on `device`. */
int32_t xhe_ctx_create(int device, uint32_t party_capacity, xhe_ctx** out);
void xhe_ctx_destroy(xhe_ctx* ctx);
/* the party capacity the context was created with: aggregated range proofs with m <= this many parties verify
 * (the reference's BP_GENS holds 512, src/proofs.rs:20; a transaction with a assets and k transfers has m = next_pow2(a + k)) */
uint32_t xhe_ctx_party_capacity(const xhe_ctx* ctx);
const char* xhe_last_error(const xhe_ctx* ctx);
/* stream all subsequent *_dev calls are launched on (a cudaStream_t passed as void*; NULL = default stream) */
int32_t xhe_ctx_set_stream(xhe_ctx* ctx, void* cuda_stream);
int32_t xhe_ctx_sync(xhe_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py "gpu_launches") */
uint64_t xhe_ctx_launch_count(const xhe_ctx* ctx);

/* ---- host-buffer entry points (copies inside) ------------------------------------------------------------ */
/* CompressedRistretto::decompress, batched (src/compressed.rs:28-34,57-62,78-84,100-106; src/tx/verify.rs:85-92;
 * src/proofs.rs:168-179,306-317).  ok[i] = 1/0 per point, never aborts.  xy (optional) = canonical affine x||y, 64 B. */
int32_t xhe_ristretto_decompress(xhe_ctx* ctx, const uint8_t* enc, size_t n, uint8_t* xy, uint8_t* ok);
/* RistrettoPoint::compress, batched (src/compressed.rs:17-21,43-50,71-75,93-97; src/elgamal.rs:40,61) from affine x||y */
int32_t xhe_ristretto_compress(xhe_ctx* ctx, const uint8_t* xy, size_t n, uint8_t* enc);
/* ristretto255 one-way map RistrettoPoint::from_uniform_bytes (src/elgamal.rs:22; bulletproofs generator chains) */
int32_t xhe_ristretto_from_uniform(xhe_ctx* ctx, const uint8_t* uniform64, size_t n, uint8_t* enc);
/* RistrettoPoint::vartime_multiscalar_mul + is_identity (src/proofs.rs:49-67).  Accepts n = 0, identity and
 * duplicate points; a non-canonical scalar or invalid point is XHE_E_ARG. */
int32_t xhe_msm_vartime(xhe_ctx* ctx, const uint8_t* scalars, const uint8_t* enc_points, size_t n, uint8_t out_enc[32], int32_t* is_identity);
/* ElGamalCiphertext Add/Sub on compressed balances (src/elgamal.rs:322-342; src/tx/verify.rs:561-609): out = bal +/- delta */
int32_t xhe_ct_update(xhe_ctx* ctx, const uint8_t* bal, const uint8_t* delta, const uint8_t* sub, size_t n, uint8_t* out, uint8_t* ok);
/* Signature::verify group part (src/elgamal.rs:38-42): r = s*H - e*P, compressed; the SHA3-512 stays with the caller */
int32_t xhe_sig_r(xhe_ctx* ctx, const uint8_t* s, const uint8_t* e, const uint8_t* pk_enc, size_t n, uint8_t* r_enc, uint8_t* ok);

/* ---- device-pointer entry points (asynchronous on the ctx stream; buffers are caller-owned device memory) -- */
int32_t xhe_decompress_dev(xhe_ctx* ctx, const void* d_enc, size_t n, void* d_affine /* n*64 B, 8x32-bit limbs x,y */, void* d_niels /* n*96 B or NULL */, void* d_ok /* n bytes */);
int32_t xhe_compress_dev(xhe_ctx* ctx, const void* d_ext /* n*128 B X,Y,Z,T */, size_t n, void* d_enc);
int32_t xhe_from_uniform_dev(xhe_ctx* ctx, const void* d_uniform64, size_t n, void* d_enc);
/* resident balance update (config 4): bal (extended, 2 points per account, in place) +/- delta (affine Niels, 2 per account) */
int32_t xhe_ct_update_resident_dev(xhe_ctx* ctx, void* d_bal_ext, const void* d_delta_niels, const void* d_sub, size_t n);
int32_t xhe_ct_update_dev(xhe_ctx* ctx, const void* d_bal, const void* d_delta, const void* d_sub, size_t n, void* d_out, void* d_ok);
/* MSM over resident points: scalars n*32 B (canonical), points as affine Niels n*96 B; out = 32-byte encoding + flag word */
size_t  xhe_msm_workspace_bytes(const xhe_ctx* ctx, size_t n);
int32_t xhe_msm_dev(xhe_ctx* ctx, const void* d_scalars, const void* d_niels, size_t n, void* d_workspace, size_t workspace_bytes, void* d_out_enc32, void* d_is_identity_u32);

/* ---- Transaction::verify_batch, device part (src/tx/verify.rs:487-517 and everything it calls in src/proofs.rs,
 * src/elgamal.rs, src/compressed.rs and the bulletproofs/dalek crates).  The caller (the Rust host in north_star; the
 * C++ host layer xelis_he_b200/host here) parses transactions, resolves state lookups, replays the Merlin transcripts
 * and hands over a struct-of-arrays batch with the Fiat-Shamir challenges and the per-proof random batch factors.
 * All pointers are HOST memory; indices refer to the point table.  Point index 0 MUST be the identity encoding (used
 * for the dud commitments of src/tx/verify.rs:466-475).  Indices >= n_points address balance-chain outputs
 * (n_points + j = output of op j). ----------------------------------------------------------------------------------- */
#define XHE_OP_PLUS_AMOUNT (1LL << 50)
#define XHE_OP_FROM_LEDGER (1LL << 49)
typedef struct xhe_batch {
  uint32_t struct_size;            /* = sizeof(xhe_batch): a caller compiled against another layout is refused with XHE_E_ARG */
  uint32_t n_tx;
  uint32_t n_points; const uint8_t* points;              /* n_points x 32 compressed ristretto255 */
  /* Signature::verify (src/elgamal.rs:38-42): r_i = s_i*H - e_i*P[sig_pk_i]; the SHA3-512 compare stays with the caller */
  uint32_t n_sigs; const uint8_t* sig_s; const uint8_t* sig_e; const uint32_t* sig_pk;
  /* balance chains (src/tx/verify.rs:301-336,354-374; src/elgamal.rs:322-377): op j yields point n_points+j =
   * prev_j + sum(+-P[term]) - amount_j*G, where prev_j >= 0 is an earlier op of the same (account, asset, half) chain
   * and prev_j < 0 encodes the initial balance point -(1+index).  One op per ciphertext half (commitment / handle).
   * prev_j = -(1+index) - XHE_OP_PLUS_AMOUNT makes the op ADD amount_j*G instead: with index 0 (the identity) and
   * unsigned terms that is one half of get_sender_output_ct (src/tx/verify.rs:107-144), the ciphertext the reference
   * hands to BlockchainVerificationState::set_output_ciphertext (src/tx/verify.rs:339-340, 582).
   * prev_j = -(1+p) - XHE_OP_FROM_LEDGER takes the initial balance from point p of the device-resident ledger `ledger` below
   * (p = 2 * slot for the commitment, 2 * slot + 1 for the handle): nothing is uploaded or decompressed for it. */
  uint32_t n_ops; const int64_t* op_prev; const uint32_t* op_term_off /* n_ops+1 */; const uint32_t* op_terms /* bit 31 = subtract */;
  const uint64_t* op_amount; uint32_t max_chain /* longest chain length (>= 1) */;
  /* CommitmentEqProof::pre_verify (src/proofs.rs:134-211): points P_src,Y0,D_src,C_src,Y1,C_dst,Y2; scalars z_s,z_x,z_r,c,w,bf */
  uint32_t n_eq; const uint32_t* eq_points; const uint8_t* eq_scalars;
  /* CiphertextValidityProof::pre_verify (src/proofs.rs:281-361): points C,Y0,P_dest,D_dest,Y1,P_src,D_src,Y2; scalars z_r,z_x,c,w,bf */
  uint32_t n_val; const uint32_t* val_points; const uint8_t* val_scalars;
  /* RangeProof::verify_batch views (src/tx/verify.rs:504-514): per proof m (parties, power of two), points
   * A,S,T1,T2,L[lg],R[lg],V[m] (lg = 6 + log2 m), scalars t_x,t_x_blinding,e_blinding,a,b,c,rho (c = intra-proof random
   * weight, rho = cross-proof batch factor), challenges y,z,x,w,u[lg] */
  uint32_t n_rp; const uint32_t* rp_m; const uint32_t* rp_point_off /* n_rp+1 */; const uint32_t* rp_points;
  const uint8_t* rp_scalars /* n_rp x 7 x 32 */; const uint32_t* rp_chal_off /* n_rp+1, in scalars */; const uint8_t* rp_challenges;
  /* OPTIONAL device-side Fiat-Shamir.  When fs_blobs != NULL the library replays the Merlin transcripts,
   * derives c / w / y / z / x / w / u_j and the random batch factors, and checks the main signatures' SHA3-512 on the
   * device from the transactions' xtx1 wire bytes; the challenge / factor slots of eq_scalars, val_scalars, rp_scalars and
   * rp_challenges are then ignored (may be zero).  fs_plan: 6 words per tx = eq_begin, val_begin, rp slot (0xffffffff none),
   * rp challenge offset, main-signature slot (0xffffffff none), flags (bit 0: sigma / range stage reached). */
  const uint8_t* fs_blobs; const uint64_t* fs_blob_off /* n_tx+1 */; const uint32_t* fs_plan; uint8_t fs_seed[32];
  /* index of this batch's first transaction inside the whole (possibly sharded) batch: transaction i draws its factors from
   * SHAKE256("xhe-batch-factors" || fs_seed || fs_index_base + i), so two shards never share a factor stream.  fs_seed must
   * be unpredictable to whoever produced the proofs (the host layer folds OS entropy into it). */
  uint64_t fs_index_base;
  /* OPTIONAL device-side layout (fast path; requires fs_blobs; SURVEY.md 8 f.2 in spirit).  When layout_on_device != 0 the
   * library also builds the point table and every per-proof / signature array from the blobs (kernel k_layout), so
   * points / sig_* / eq_* / val_* / rp_points / rp_scalars may be NULL.  The host supplies counts (n_points, n_sigs, n_eq, n_val,
   * n_rp), rp_m, rp_point_off, rp_chal_off, the balance-chain ops, fs_plan with EIGHT words per tx (the six above + point base
   * + first balance-op index; the per-tx point layout is documented at k_layout) and the state-derived encodings
   * (initial balance halves) region_b, placed at point indices [n_points - n_region_b, n_points). */
  uint32_t layout_on_device; uint32_t n_region_b; const uint8_t* region_b;
  /* OPTIONAL device-resident ledger (xhe_ledger_* below): the table that XHE_OP_FROM_LEDGER ops read */
  const struct xhe_ledger* ledger;
} xhe_batch;

typedef struct xhe_verdict {
  uint32_t struct_size;          /* = sizeof(xhe_verdict) */
  int32_t sigma_is_identity;     /* BatchCollector::verify (src/proofs.rs:49-67) */
  int32_t range_is_identity;     /* RangeProof::verify_batch mega-check */
  uint8_t sigma_enc[32], range_enc[32];
  uint8_t sigma_ext[128], range_ext[128];   /* un-normalised partial sums (X,Y,Z,T packed) for multi-GPU combination */
  uint8_t* point_ok;             /* n_points: decompression flags (caller-allocated) */
  uint8_t* sig_r;                /* n_sigs x 32: compressed r_i (caller-allocated) */
  uint8_t* op_out;               /* n_ops x 32: compressed chain outputs = updated balance halves (caller-allocated) */
  uint8_t* sig_ok;               /* n_sigs: filled only in device Fiat-Shamir mode, for the main-signature slots (caller-allocated, optional) */
  uint32_t device_flags;         /* device-layout mode: bit 0 identity-encoded sigma-proof Y, bit 1 some point failed to decompress, bit 2 some signature
                                    mismatched, bit 3 a state-derived point (region_b) failed to decompress, bit 4 identity-encoded A/S/T/L/R of a range proof */
  uint8_t* tx_flags;             /* n_tx (caller-allocated, optional, device-layout mode): per transaction bits 0-2 as above and bit 3 = its range proof has
                                    an identity-encoded point; copied back only when device_flags != 0 -- lets the host re-decide ONE transaction, not the batch */
} xhe_verdict;

/* returns XHE_OK when the device work completed (verdict fields filled) -- the accept/reject decision and its error
 * precedence belong to the caller, which knows the transaction structure. */
int32_t xhe_verify_batch(xhe_ctx* ctx, const xhe_batch* batch, xhe_verdict* verdict);
/* The same call in three stages, so a batch can stay resident in HBM and be re-run (bench.py times xhe_batch_run alone
 * for `value`): prepare = arena allocation + H2D, run = kernels only (asynchronous), fetch = D2H + stream sync. */
int32_t xhe_batch_prepare(xhe_ctx* ctx, const xhe_batch* batch);
int32_t xhe_batch_run(xhe_ctx* ctx);
int32_t xhe_batch_fetch(xhe_ctx* ctx, xhe_verdict* verdict);
size_t  xhe_batch_h2d_bytes(const xhe_ctx* ctx);
size_t  xhe_batch_d2h_bytes(const xhe_ctx* ctx);
/* K7: add n partial sums (n x 128 B as produced in *_ext) and test the Ristretto identity; out_enc optional */
int32_t xhe_combine_partials(xhe_ctx* ctx, const uint8_t* ext, size_t n, uint8_t out_enc[32], int32_t* is_identity);
/* Cross-rank decision of a sharded batch: sum of n <= 224 canonical encodings -- the per-rank partial
 * results of the sigma or range MSM (src/proofs.rs:49-67 decides on the identity of the total) -- as one 32-thread
 * kernel with no allocation.  *all_valid = 0 if an encoding does not decode (such inputs are left out of the sum). */
int32_t xhe_sum_encodings(xhe_ctx* ctx, const uint8_t* enc, size_t n, uint8_t out_enc[32], int32_t* is_identity, int32_t* all_valid);
/* The same decision with everything on the device (asynchronous on the ctx stream), so that a sharded step -- kernels, exchange,
 * decision -- can be timed with CUDA events: xhe_batch_record_dev writes this rank's 80-byte record (int32 code, int64 first
 * failing tx at offset 4, sigma partial at 12, range partial at 44) of the batch that xhe_batch_run just processed; after an
 * all-gather of the records (ncclAllGather on the same stream), xhe_shard_decide_dev writes int32 code and, at offset 8, int64
 * index: first failing transaction of the whole batch, else GenericProof if the summed sigma partials are not the identity, else
 * RangeProof (structural failure in a shard, or the summed range partials), else 0.  Code 0xFF in a record = that shard saw a
 * per-transaction anomaly its host has to name. */
int32_t xhe_batch_record_dev(xhe_ctx* ctx, void* d_record80);
int32_t xhe_shard_decide_dev(xhe_ctx* ctx, const void* d_records, uint32_t world, void* d_out16);
/* copy nbytes (<= 1 MiB) between device-accessible addresses (device memory or pinned host memory) with a kernel on the
 * ctx stream -- for small control messages that must not queue behind bulk transfers on the copy engines */
int32_t xhe_copy_small(xhe_ctx* ctx, void* dst, const void* src, size_t nbytes);

/* ---- device-resident ledger: a state backend for BlockchainVerificationState (src/tx/verify.rs:25-77) whose
 * balances live on the device as decompressed points (coordinate-planar extended, 256 B per (account, asset)); the
 * (account || asset) -> slot index stays with the host.  Balances are compressed only on export.  A ledger belongs to one ctx. */
typedef struct xhe_ledger xhe_ledger;
int32_t xhe_ledger_create(xhe_ctx* ctx, size_t capacity /* (account, asset) slots */, xhe_ledger** out);
void    xhe_ledger_destroy(xhe_ledger* ledger);
size_t  xhe_ledger_size(const xhe_ledger* ledger);
/* insert / overwrite n balances: keys n x 64 (account || asset), cts n x 64 CompressedCiphertext (src/compressed.rs:37-41), decoded on
 * the device.  ok[i] (optional) = 0 when ciphertext i does not decode (it is then reported as not found). */
int32_t xhe_ledger_load(xhe_ledger* ledger, const uint8_t* keys, const uint8_t* cts, size_t n, uint8_t* ok);
/* ElGamalCiphertext Add / Sub in place (src/elgamal.rs:322-342; apply_without_verify's algebra, src/tx/verify.rs:574,602):
 * bal[key_i] +/- delta_i with 64-byte compressed deltas; repeated keys are applied in order.  status[i] (optional): 0 applied,
 * 1 unknown key, 2 ill-formed delta (balance untouched). */
int32_t xhe_ledger_update(xhe_ledger* ledger, const uint8_t* keys, const uint8_t* deltas, const uint8_t* sub, size_t n, uint8_t* status);
/* the same for EVERY slot [0, size) in slot (= insertion) order, deltas resident on the device as planar affine Niels
 * [y+x | y-x | 2dxy][2 size][8 words]: config 4's HBM-bound form, asynchronous on the ctx stream */
int32_t xhe_ledger_update_dense_dev(xhe_ledger* ledger, const void* d_delta_niels_planar, const void* d_sub /* size bytes */);
/* compressed export on demand (what get_account_balance returns): found[i] (optional) = 0 for an unknown key (64 zero bytes out) */
int32_t xhe_ledger_export(xhe_ledger* ledger, const uint8_t* keys, size_t n, uint8_t* out_cts, uint8_t* found);
/* After an ACCEPTED batch that is still resident on its ctx (xhe_verify_batch / xhe_batch_run): write the outputs of balance-chain
 * ops ops[i] -- commitment op, its handle op is ops[i] + 1 -- into slots slots[i], in order, on the device (asynchronous on the
 * ctx stream; no balance crosses the bus in either direction).  This is update_account_balance (src/tx/verify.rs:329-336,367-374)
 * for a state whose balances live in the ledger. */
int32_t xhe_ledger_commit_batch(xhe_ledger* ledger, xhe_ctx* ctx, const uint32_t* slots, const uint32_t* ops, size_t n);
/* slot of a key (0xFFFFFFFF: unknown, or stored ciphertext undecodable) */
uint32_t xhe_ledger_slot(const xhe_ledger* ledger, const uint8_t key64[64]);
/* device-side snapshot / restore of the whole table (tests and benchmarks that re-verify the same batch) */
int32_t xhe_ledger_snapshot(xhe_ledger* ledger);
int32_t xhe_ledger_restore(xhe_ledger* ledger);
/* the table itself for callers that launch their own kernels: 4 planes X, Y, Z, T of *plane_stride_points points x 32 bytes;
 * slot s owns points 2s (commitment) and 2s + 1 (handle) */
void*   xhe_ledger_device_table(const xhe_ledger* ledger, size_t* plane_stride_points);

/* ---- measurement helpers ---------------------------------------------------------------------------------- */
/* integer-multiply pipe microbenchmarks: which = 0 IMAD.lo, 1 IMAD.HI, 2 IMAD.WIDE.U32; returns
 * achieved instructions/s summed over the device in *rate. */
int32_t xhe_measure_int_peak(xhe_ctx* ctx, int which, double* rate);
/* CUDA-event timing of the main kernels of xhe_batch_run / the MSM (roofline evidence): enable, run, then read.
 * units[i] = algorithmic limb products (DESIGN.md work model) accumulated for kernel i. */
int32_t xhe_ctx_timing(xhe_ctx* ctx, int enable);
/* ---- decoding of decrypted amounts ---------------------------------------------------------------
 * ECDLPInstance::decode / par_decode (src/elgamal.rs:67-92; the curve25519-dalek fork's ecdlp module) and
 * ElGamalSecretKey::decrypt (src/elgamal.rs:140-145, M = C - s * D): find v in [0, 2^range_bits) with v * G == M by a
 * baby-step / giant-step search on the device.  A table holds 2^l1_bits baby steps (8 bytes per slot, 2^(l1_bits + 1) slots,
 * built on the device at creation); one decode then takes 2^(range_bits - l1_bits - 1) giant steps, spread over a warp.
 * out_value[i] = v, or -1; status[i] = 1 found, 0 no such v in range, 2 the input does not decode.  range_bits <= 62 and
 * at most l1_bits + 33. */
typedef struct xhe_ecdlp xhe_ecdlp;
int32_t xhe_ecdlp_create(xhe_ctx* ctx, uint32_t l1_bits, xhe_ecdlp** out);
void    xhe_ecdlp_destroy(xhe_ecdlp* table);
size_t  xhe_ecdlp_table_bytes(const xhe_ecdlp* table);
int32_t xhe_ecdlp_decode(xhe_ecdlp* table, const uint8_t* points /* n x 32: encodings of M */, size_t n, uint32_t range_bits, int64_t* out_value, uint8_t* status);
int32_t xhe_decrypt_decode(xhe_ecdlp* table, const uint8_t secret_key[32], const uint8_t* ciphertexts /* n x 64: commitment || handle */, size_t n, uint32_t range_bits,
                           int64_t* out_value, uint8_t* status);

/* run the independent pipelines of xhe_batch_run back to back on one stream (isolated per-kernel timing; slower) */
int32_t xhe_ctx_set_serial(xhe_ctx* ctx, int serial);
int32_t xhe_ctx_timing_read(xhe_ctx* ctx, const char** names, double* ms, uint64_t* launches, double* units, int cap);
/* start/end (ms since the start of the run) of every timed kernel of the last timed xhe_batch_run, as gathered by the
 * last xhe_ctx_timing_read; returns the number of spans written (diagnostics for the stream pipelines) */
int32_t xhe_ctx_timeline(xhe_ctx* ctx, const char** names, float* t0, float* t1, int cap);
/* self-test of the arithmetic layer: runs op (tests/hostemu op codes) on n operand pairs on the device */
int32_t xhe_selftest_fe(xhe_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out);
/* self-test of the warp-cooperative arithmetic of the MSM's Horner chain (csrc/oct.cuh): op 0 mul, 1 add, 2 sub over n field
 * elements (8 words each); op 3 doubling, 4 complete addition over n extended points (32 words each: X, Y, Z, T) */
int32_t xhe_selftest_oct(xhe_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
