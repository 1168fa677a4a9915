/* oracle/proofs.h -- TEST INFRASTRUCTURE.  Restates src/proofs.rs (CommitmentEqProof 73-222, CiphertextValidityProof
 * 237-371, BatchCollector 40-68), src/transcript.rs:37-111 (ProtocolTranscript) and src/elgamal.rs:16-65,194-200
 * (H, Signature) of the reference over the oracle's own arithmetic. */
#ifndef XO_PROOFS_H
#define XO_PROOFS_H
#include "ge.h"
#include "merlin.h"
#include "keccak.h"
enum { XO_OK = 0, XO_ERR_SIGNATURE = 1, XO_ERR_DECOMPRESSION = 2, XO_ERR_COMMITMENT_EQ_PROOF = 3, XO_ERR_CT_VALIDITY_PROOF = 4,
       XO_ERR_GENERIC_PROOF = 5, XO_ERR_RANGE_PROOF = 6, XO_ERR_TRANSCRIPT = 7, XO_ERR_FORMAT = 8, XO_ERR_INVALID_NONCE = 9,
       XO_ERR_STATE = 10, XO_ERR_PARSE = 11 };
/* deterministic test RNG: SHAKE256 stream */
typedef struct { xo_sponge sp; } xo_rng;
void xo_rng_init(xo_rng *r, const void *seed, size_t n);
void xo_rng_bytes(xo_rng *r, void *out, size_t n);
void xo_rng_scalar(xo_rng *r, sc *s);
const ge *xo_G(void); const ge *xo_H(void);
/* transcript helpers (src/transcript.rs) */
void xo_challenge_scalar(xo_transcript *t, const char *label, sc *out);
int  xo_validate_and_append_point(xo_transcript *t, const char *label, const uint8_t p[32]);
/* batch collector (src/proofs.rs:40-68) */
typedef struct { sc *scalars; ge *points; size_t n, cap; sc g_scalar, h_scalar; } xo_collector;
void xo_collector_init(xo_collector *c); void xo_collector_free(xo_collector *c);
void xo_collector_push(xo_collector *c, const sc *s, const ge *p);
int  xo_collector_verify(const xo_collector *c, uint8_t out_enc[32]); /* 1 iff identity */
/* proofs: 192-byte eq proof (Y0,Y1,Y2,z_s,z_x,z_r), 160-byte validity proof (Y0,Y1,Y2,z_r,z_x) */
void xo_eq_proof_new(uint8_t out[192], const sc *sk, const ge *P_src, const ge *D_src, const sc *opening, uint64_t amount, xo_transcript *t, xo_rng *rng);
int  xo_eq_proof_pre_verify(const uint8_t proof[192], const ge *P_src, const ge *C_src, const ge *D_src, const ge *C_dst, xo_transcript *t, xo_collector *c, xo_rng *rng);
void xo_validity_proof_new(uint8_t out[160], const ge *P_dest, const ge *P_src, uint64_t amount, const sc *opening, xo_transcript *t, xo_rng *rng);
int  xo_validity_proof_pre_verify(const uint8_t proof[160], const ge *C, const ge *P_dest, const ge *P_src, const ge *D_dest, const ge *D_src, xo_transcript *t, xo_collector *c, xo_rng *rng);
/* signature (src/elgamal.rs:38-65,194-200): 64 bytes s || e */
void xo_sign(uint8_t sig[64], const sc *sk, const uint8_t pk_enc[32], const uint8_t *msg, size_t n, xo_rng *rng);
int  xo_sig_verify(const uint8_t sig[64], const uint8_t *msg, size_t n, const ge *pk, uint8_t r_enc_out[32]);
#endif
