/* oracle/bp.h -- TEST INFRASTRUCTURE.  Aggregated Bulletproofs range proofs (n = 64 bits, m parties) restated from the
 * public dalek-cryptography protocol that the un-vendored `bulletproofs 5.2.0 @288386bb` (xelis fork) implements:
 * generators (BulletproofGens::new(64, 512), PedersenGens::default(); reference src/proofs.rs:19-22), prover
 * (RangeProof::prove_multiple; src/tx/builder.rs:525-533), single verification (verify_multiple; src/tx/verify.rs:531-539)
 * and cross-proof batch verification (verify_batch / verification_view; src/tx/verify.rs:504-514).
 * PARITY UNPINNED at this boundary: the fork's source and fixtures are not in the container; transcript
 * labels follow upstream dalek bulletproofs; B_blinding is pinned by its well-known encoding. */
#ifndef XO_BP_H
#define XO_BP_H
#include "proofs.h"
#define XO_BP_N 64
#define XO_BP_PARTY_CAP 512
const ge *xo_bp_G(int party, int i); /* lazily derived; thread-safe only after xo_bp_ensure(m) */
const ge *xo_bp_H(int party, int i);
void xo_bp_ensure(int m);
size_t xo_rp_size(int m);  /* bytes: 32 * (9 + 2 lg(64 m)) */
/* prove: values/blindings of length m (m a power of two <= 512); appends to the transcript exactly like the verifier */
int xo_rp_prove(uint8_t *out, const uint64_t *values, const sc *blindings, int m, xo_transcript *t, xo_rng *rng);
/* one proof to verify in a batch: transcript already advanced to the range-proof position, commitments compressed (+ decoded) */
typedef struct { const uint8_t *proof; size_t len; xo_transcript *t; const uint8_t *commit_enc; const ge *commit_pts; int m; } xo_rp_item;
int xo_rp_verify_batch(const xo_rp_item *items, size_t n_items, xo_rng *rng, uint8_t out_enc[32]); /* XO_OK / XO_ERR_RANGE_PROOF */
int xo_rp_verify_batch_ex(const xo_rp_item *items, size_t n_items, xo_rng *rng, uint8_t out_enc[32], int no_decision);
#endif
