/* oracle/tx.h -- TEST INFRASTRUCTURE.  Transaction model, builder (test-vector minting), mock ledger and the three
 * reference entry points restated: Transaction::verify_batch (src/tx/verify.rs:487-517), Transaction::verify (520-542),
 * Transaction::apply_without_verify (545-619), with pre_verify (203-485), to_bytes (623-688), the builder
 * (src/tx/builder.rs:320-554) and mock::Ledger (src/lib.rs:97-242).
 *
 * Wire format "xtx1" (the reference has serde derives only, src/tx/mod.rs:19-119; this framing is ours and is shared
 * with the product's host parser as a SPEC, not as code).  All integers little-endian:
 *   0  u8  version | 1 u8 type (0 Transfers,1 Burn,2 CallContract,3 DeployContract,4 MultiSig) | 2 u8 n_source_commitments
 *   3  u8  multisig signature count (0xFF = None)
 *   4  u32 count  (Transfers: k; MultiSig: signers; CallContract: assets; else 0)
 *   8  u32 aux    (CallContract: params; DeployContract: code bytes; MultiSig: threshold; else 0)
 *   12 u32 range-proof byte length | 16 source[32] | 48 u64 fee | 56 u64 nonce | 64 body | range proof |
 *   a x {asset[32], commitment[32], eq_proof[192]} | multisig x {u8 index, sig[64]} | signature[64]
 *   Transfers body: k x {asset, dest, commitment, sender_handle, receiver_handle (32 each), proof[160], u32 extra_len (0xFFFFFFFF none), extra}
 *   Burn: asset[32], u64 amount.  CallContract: contract[32], assets x {asset[32], u64}, params x {u32 klen, key, u32 vlen, value}.
 *   DeployContract: code[aux].  MultiSig: signers x pubkey[32]. */
#ifndef XO_TX_H
#define XO_TX_H
#include "bp.h"
enum { XO_TX_TRANSFERS = 0, XO_TX_BURN = 1, XO_TX_CALL = 2, XO_TX_DEPLOY = 3, XO_TX_MULTISIG = 4 };
typedef struct { const uint8_t *asset, *dest, *commitment, *sender_handle, *receiver_handle, *proof, *extra; uint32_t extra_len; int has_extra; } xo_transfer;
typedef struct {
  const uint8_t *blob; size_t len;
  uint8_t version, type, n_sc; int n_ms; uint32_t count, aux, rp_len; const uint8_t *source; uint64_t fee, nonce;
  xo_transfer *transfers; const uint8_t *body; size_t body_len; const uint8_t *rp; const uint8_t *sc; const uint8_t *ms; const uint8_t *sig;
} xo_tx;
int  xo_tx_parse(xo_tx *tx, const uint8_t *blob, size_t len);   /* XO_OK / XO_ERR_PARSE; allocates tx->transfers */
void xo_tx_free(xo_tx *tx);
size_t xo_tx_to_bytes(const xo_tx *tx, uint8_t **out, size_t *multisig_index); /* src/tx/verify.rs:623-688 */
/* mock ledger (src/lib.rs:106-201) */
typedef struct xo_ledger xo_ledger;
xo_ledger *xo_ledger_new(void); xo_ledger *xo_ledger_clone(const xo_ledger *l); void xo_ledger_free(xo_ledger *l);
void xo_ledger_set_balance(xo_ledger *l, const uint8_t pk[32], const uint8_t asset[32], const uint8_t ct[64]);
int  xo_ledger_get_balance(const xo_ledger *l, const uint8_t pk[32], const uint8_t asset[32], uint8_t ct[64]);
void xo_ledger_set_nonce(xo_ledger *l, const uint8_t pk[32], uint64_t nonce);
int  xo_ledger_get_nonce(const xo_ledger *l, const uint8_t pk[32], uint64_t *nonce);
void xo_ledger_set_multisig(xo_ledger *l, const uint8_t pk[32], const uint8_t *signers, int n, uint8_t threshold);
int  xo_ledger_get_multisig(const xo_ledger *l, const uint8_t pk[32], const uint8_t **signers, int *n, uint8_t *threshold);
size_t xo_ledger_dump(const xo_ledger *l, uint8_t *out, size_t cap); /* sorted (pk,asset,ct) records, 128 B each */
void xo_ledger_record_outputs(xo_ledger *l, int on);   /* default off */
void xo_ledger_set_output(xo_ledger *l, const uint8_t pk[32], const uint8_t asset[32], const uint8_t ct[64]);   /* set_output_ciphertext, compressed */
size_t xo_ledger_dump_outputs(const xo_ledger *l, uint8_t *out, size_t cap); /* sorted (pk,asset,output ct) records */
/* entry points; *fail_index = index of the first failing tx (or -1 for the batch-level MSM checks) */
int xo_verify_batch(const uint8_t *const *blobs, const size_t *lens, size_t n, xo_ledger *state, xo_rng *rng, long *fail_index);
int xo_verify_batch_ex(const uint8_t *const *blobs, const size_t *lens, size_t n, xo_ledger *state, xo_rng *rng, long *fail_index, uint8_t *partial64);
int xo_verify(const uint8_t *blob, size_t len, xo_ledger *state, xo_rng *rng);
int xo_apply_without_verify(const uint8_t *blob, size_t len, xo_ledger *state);
/* builder (src/tx/builder.rs).  Spec for one transfer / the tx data; plaintext balances come from `balances` */
typedef struct { uint8_t asset[32], dest[32]; uint64_t amount; const uint8_t *extra; uint32_t extra_len; int has_extra; } xo_transfer_spec;
typedef struct {
  uint8_t version, type; uint8_t source[32]; uint64_t fee, nonce;
  const xo_transfer_spec *transfers; uint32_t n_transfers;
  uint8_t burn_asset[32]; uint64_t burn_amount;
  uint8_t contract[32]; const uint8_t *call_assets; const uint64_t *call_amounts; uint32_t n_call_assets;
  const uint8_t *raw_tail; uint32_t raw_tail_len;   /* CallContract params blob (pre-framed) or DeployContract code */
  uint32_t n_params;
  const uint8_t *signers; uint32_t n_signers; uint8_t threshold;
  const uint8_t *assets; const uint64_t *balances; uint32_t n_assets; /* plaintext source balances per used asset, in commitment order */
} xo_tx_spec;
/* builds an unsigned tx then signs; multisig = optional (index, sk) co-signers */
size_t xo_tx_build(uint8_t **out, const xo_tx_spec *spec, const sc *sk, const xo_ledger *state, xo_rng *rng,
                   const uint8_t *ms_index, const sc *ms_sk, int n_ms);
void xo_keygen(xo_rng *rng, sc *sk, uint8_t pk[32]);
void xo_pubkey_from_secret(const sc *sk, uint8_t pk[32], ge *P);
void xo_encrypt(uint8_t ct[64], const ge *P, uint64_t amount, const sc *opening);
void xo_blake3(const uint8_t *in, size_t n, uint8_t out[32]);
#endif
