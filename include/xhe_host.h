/* xhe_host.h -- C ABI of the HOST layer of libxhe_cuda.so (xelis_he_b200/host/verifier.*): the part of the reference's
 * verification API that north_star leaves on the host -- transaction parsing, state lookups through
 * BlockchainVerificationState, Merlin transcripts, message hashes, verdict precedence -- mirrored in C++ and exported for
 * hosts that are not Rust (the Python tests, bench.py).  A Rust integration keeps its own host code and binds include/xhe.h
 * directly; these entry points show the same call sequence end to end.
 *
 *   reference (Rust)                                            here
 *   mock::Ledger                       src/lib.rs:106-201         xheh_ledger_*
 *   Transaction::verify_batch          src/tx/verify.rs:487-517   xheh_verify_batch_ex / xheh_verify_batch_shard
 *   Transaction::apply_without_verify  src/tx/verify.rs:545-619   xheh_apply_without_verify
 *   Transaction::to_bytes              src/tx/verify.rs:623-688   xheh_tx_to_bytes
 *
 * Return values: 0 = XHE_OK, > 0 verdicts, < 0 infrastructure errors (include/xhe.h).  *fail_index = first failing
 * transaction (index into the batch that was passed), -1 for the two batch-level MSM checks.
 *
 * ---- wire format "xtx1" ---------------------------------------------------------------------------------------------
 * The reference has serde derives only (src/tx/mod.rs:19-119) and no canonical encoding; the host layer and the device
 * kernels that read transactions (k_layout, k_fiat_shamir, k_sig_hash_*) share this framing.  One blob per transaction,
 * all integers little-endian, every 32-byte field a compressed ristretto255 point, a canonical scalar or a hash:
 *
 *   offset  field
 *   0       u8  version
 *   1       u8  type: 0 Transfers, 1 Burn, 2 CallContract, 3 DeployContract, 4 MultiSig        (src/tx/mod.rs:83-93)
 *   2       u8  a = number of new source commitments                                          (src/tx/mod.rs:95-100)
 *   3       u8  number of multisig co-signatures, 0xFF = None
 *   4       u32 count: Transfers k | MultiSig signers | CallContract assets | else 0
 *   8       u32 aux:   CallContract params | DeployContract code bytes | MultiSig threshold | else 0
 *   12      u32 range-proof byte length = 32 * (9 + 2 lg), lg = 6 + log2(next_pow2(a + k))
 *   16      source public key [32]
 *   48      u64 fee
 *   56      u64 nonce
 *   64      body:
 *             Transfers:      k x { asset[32], dest[32], commitment[32], sender_handle[32], receiver_handle[32],
 *                                   ct_validity_proof[160] = Y_0,Y_1,Y_2,z_r,z_x, u32 extra_len (0xFFFFFFFF = None), extra[extra_len] }
 *             Burn:           asset[32], u64 amount
 *             CallContract:   contract[32], count x { asset[32], u64 amount }, aux x { u32 klen, key, u32 vlen, value }
 *             DeployContract: code[aux]
 *             MultiSig:       count x signer public key [32]
 *   then    range proof: A, S, T_1, T_2, t_x, t_x_blinding, e_blinding, lg x (L, R), a, b     (32 bytes each)
 *   then    a x { asset[32], new source commitment[32], eq_proof[192] = Y_0,Y_1,Y_2,z_s,z_x,z_r }
 *   then    multisig co-signatures x { u8 signer index, signature[64] = s, e }
 *   then    signature[64] = s, e                                                               (src/elgamal.rs:26-36)
 *
 * A blob whose framing is inconsistent, or that carries a non-canonical scalar where serde would reject one, is
 * XHE_ERR_PARSE.  Blobs handed to the device are padded to 16 bytes (fs_blob_off in xhe_batch). */
#ifndef XHE_HOST_H
#define XHE_HOST_H
#include "xhe.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- mock::Ledger (src/lib.rs:106-201): balances are 64-byte CompressedCiphertexts keyed by (account, asset) ---- */
void*   xheh_ledger_new(void);
void*   xheh_ledger_clone(const void* ledger);
void    xheh_ledger_free(void* ledger);
void    xheh_ledger_set_balance(void* ledger, const uint8_t pk[32], const uint8_t asset[32], const uint8_t ct[64]);
int     xheh_ledger_get_balance(void* ledger, const uint8_t pk[32], const uint8_t asset[32], uint8_t ct[64]);   /* 1 = found */
void    xheh_ledger_set_nonce(void* ledger, const uint8_t pk[32], uint64_t nonce);
void    xheh_ledger_set_multisig(void* ledger, const uint8_t pk[32], const uint8_t* signers, size_t n, uint8_t threshold);
int     xheh_ledger_has_multisig(void* ledger, const uint8_t pk[32]);
size_t  xheh_ledger_size(void* ledger);
void    xheh_ledger_import(void* ledger, const uint8_t* records /* n x (pk[32] asset[32] ct[64]) */, size_t n);   /* nonce 0 for new accounts */
size_t  xheh_ledger_export(void* ledger, uint8_t* out, size_t cap);                                              /* same record layout */
int32_t xheh_ledger_apply_records(void* ledger, const uint8_t* records, size_t n);   /* update_account_balance per record, in order */
/* set_output_ciphertext (src/tx/verify.rs:60-66): recorded only on request, like the reference mock drops them */
void    xheh_ledger_record_outputs(void* ledger, int on);
size_t  xheh_ledger_outputs_size(void* ledger);
size_t  xheh_ledger_export_outputs(void* ledger, uint8_t* out, size_t cap);

/* ---- device-resident state: the same interface over an xhe_ledger on ctx's device.  A handle returned here is
 * accepted wherever a `ledger` is (verify, apply, commit): the fast path then reads balances from the device table and commits the
 * accepted updates there -- no balance crosses the bus; other paths see the compressed view (export on demand). ---- */
void*   xheh_dledger_new(xhe_ctx* ctx, size_t capacity);
void    xheh_dledger_free(void* dledger);
int32_t xheh_dledger_import(void* dledger, const uint8_t* records, size_t n);        /* pk[32] asset[32] ct[64]; nonce 0 for new accounts */
size_t  xheh_dledger_export(void* dledger, uint8_t* out, size_t cap);                /* every balance, same record layout (compressed on demand) */
void    xheh_dledger_set_multisig(void* dledger, const uint8_t pk[32], const uint8_t* signers, size_t n, uint8_t threshold);
int32_t xheh_dledger_snapshot(void* dledger);                                        /* device-side copy of the table ... */
int32_t xheh_dledger_restore(void* dledger);                                         /* ... and back (re-verifying the same batch) */

/* ---- Transaction::verify_batch (src/tx/verify.rs:487-517) ---------------------------------------------------------
 * flags: bit 0 device-side Fiat-Shamir; bit 1 shard mode (partial64 != NULL: no identity decision here, the
 * partial sigma / range encodings are returned and the state updates are held back); bit 2 fast path (device-side layout,
 * optimistic; a failing transaction is re-decided alone by the exact path); bit 3 replayable batch factors (tests only).
 * seed = personalisation of the batch factors; OS entropy is always folded in unless bit 3 is set.
 * timings7 (optional): parse, resolve, transcript, device, finish, total (ms), Keccak permutations on the host (-1: fast path). */
int32_t xheh_verify_batch_ex(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, const uint8_t* seed, size_t seed_len,
                             int threads, uint32_t flags, long* fail_index, double* timings7, uint8_t* partial64);
int32_t xheh_verify_batch(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, const uint8_t* seed, size_t seed_len,
                          int threads, long* fail_index, double* timings7);
int32_t xheh_verify_batch_partial(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, const uint8_t* seed, size_t seed_len,
                                  int threads, long* fail_index, double* timings7, uint8_t* partial64);
/* one rank's share [lo, hi) of a batch sharded over several GPUs; blobs = the WHOLE batch.  Balance chains
 * and multisig settings that start in [0, lo) are followed, so verdicts equal the sequential walk of the reference
 * (src/tx/verify.rs:301-374) however the batch is cut.  Always shard mode. */
int32_t xheh_verify_batch_shard(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, size_t lo, size_t hi,
                                const uint8_t* seed, size_t seed_len, int threads, uint32_t flags, long* fail_index, double* timings7, uint8_t* partial64);
/* key-digest index of a batch: per transaction, 64-bit digests of the (account, asset) balances it moves and of a multisig
 * setting it makes (a few 8-byte words per transaction).  Built once, where the transactions are received and framed -- the
 * counterpart of the reference's deserialisation into `Transaction` values, which also happens before verify_batch -- and
 * handed to the shard-mode call, which then finds the earlier transactions its shard depends on (src/tx/verify.rs:301-374)
 * without reading the other shards' bytes.  xheh_shard_dependencies answers that question on its own (index may be NULL:
 * the blobs are scanned); it returns the number of such transactions, in batch order, and writes at most cap indices. */
void*   xheh_batch_index_build(const uint8_t* const* blobs, const size_t* lens, size_t n, int threads);
void    xheh_batch_index_free(void* index);
size_t  xheh_batch_index_bytes(const void* index);
long    xheh_shard_dependencies(const uint8_t* const* blobs, const size_t* lens, size_t n, size_t lo, size_t hi, const void* index, int threads, size_t* out, size_t cap);
int32_t xheh_verify_batch_shard_ix(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, size_t lo, size_t hi,
                                   const uint8_t* seed, size_t seed_len, int threads, uint32_t flags, long* fail_index, double* timings7, uint8_t* partial64, const void* index);
/* the state updates a shard-mode call held back: apply them (after the cross-rank decision), or detach them so the context can
 * take its next batch, then commit / export (128-byte records account, asset, ciphertext) / drop */
int32_t xheh_commit_pending(xhe_ctx* ctx, void* ledger);
void*   xheh_take_pending(xhe_ctx* ctx);
int32_t xheh_commit_taken(void* pending, void* ledger);
size_t  xheh_export_taken(void* pending, uint8_t* out, size_t cap);
void    xheh_drop_taken(void* pending);

/* zero-copy input: a page-locked buffer in which the caller lays the batch out the way the device reads it -- the
 * xtx1 blobs back to back, each padded to a multiple of 16 bytes.  When blobs[i] point into such a buffer in that order, the fast
 * path uploads them where they are instead of gathering them into its staging buffer first. */
void*   xheh_blob_arena_alloc(size_t bytes);
void    xheh_blob_arena_free(void* arena);

/* ---- Transaction::apply_without_verify over a list of transactions, in order (src/tx/verify.rs:545-619) ---- */
int32_t xheh_apply_without_verify(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n);

/* ---- host-only pieces, exported for tests against public vectors ---- */
int32_t xheh_tx_to_bytes(const uint8_t* blob, size_t len, uint8_t* out, size_t cap, size_t* out_len, size_t* multisig_index);   /* src/tx/verify.rs:623-688 */
void    xheh_merlin_test(const char* proto, const char* label, const uint8_t* msg, size_t n, const char* chal_label, uint8_t* out, size_t outlen);
void    xheh_sha3_512(const uint8_t* m, size_t n, uint8_t* out64);
void    xheh_shake256(const uint8_t* m, size_t n, uint8_t* out, size_t outlen);
void    xheh_blake3(const uint8_t* m, size_t n, uint8_t* out32);
void    xheh_reduce_wide(const uint8_t in64[64], uint8_t out32[32]);     /* Scalar::from_bytes_mod_order_wide */
void    xheh_const_g_2_64(uint8_t out32[32]);                            /* encoding of 2^64 * G */

#ifdef __cplusplus
}
#endif
#endif
