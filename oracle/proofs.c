/* oracle/proofs.c -- TEST INFRASTRUCTURE.  See proofs.h. */
#include "proofs.h"
#include <stdlib.h>
#include <string.h>
void xo_rng_init(xo_rng *r, const void *seed, size_t n) { xo_sponge_init(&r->sp, 136); xo_sponge_absorb(&r->sp, "xhe-oracle-rng", 14); xo_sponge_absorb(&r->sp, seed, n); xo_sponge_finish(&r->sp, 0x1f); }
void xo_rng_bytes(xo_rng *r, void *out, size_t n) { xo_sponge_squeeze(&r->sp, out, n); }
void xo_rng_scalar(xo_rng *r, sc *s) { uint8_t b[64]; xo_rng_bytes(r, b, 64); sc_frombytes_wide(s, b); }
static ge G_pt, H_pt; static int gens_ready = 0;
static void init_gens(void) { /* src/elgamal.rs:16-24 */
  if (gens_ready) return; ge_basepoint(&G_pt); uint8_t e[32], h[64]; ristretto_encode(e, &G_pt); xo_sha3_512(e, 32, h); ristretto_from_uniform(&H_pt, h); gens_ready = 1;
}
const ge *xo_G(void) { init_gens(); return &G_pt; }
const ge *xo_H(void) { init_gens(); return &H_pt; }
void xo_challenge_scalar(xo_transcript *t, const char *label, sc *out) { uint8_t b[64]; xo_transcript_challenge(t, label, b, 64); sc_frombytes_wide(out, b); }
int xo_validate_and_append_point(xo_transcript *t, const char *label, const uint8_t p[32]) { /* src/transcript.rs:73-84 */
  uint8_t z = 0; for (int i = 0; i < 32; i++) z |= p[i]; if (!z) return 0; xo_transcript_append(t, label, p, 32); return 1;
}
void xo_collector_init(xo_collector *c) { memset(c, 0, sizeof *c); }
void xo_collector_free(xo_collector *c) { free(c->scalars); free(c->points); memset(c, 0, sizeof *c); }
void xo_collector_push(xo_collector *c, const sc *s, const ge *p) {
  if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 64; c->scalars = realloc(c->scalars, c->cap * sizeof(sc)); c->points = realloc(c->points, c->cap * sizeof(ge)); }
  c->scalars[c->n] = *s; c->points[c->n] = *p; c->n++;
}
int xo_collector_verify(const xo_collector *c, uint8_t out_enc[32]) { /* src/proofs.rs:49-67 */
  size_t n = c->n + 2; sc *s = malloc(n * sizeof(sc)); ge *p = malloc(n * sizeof(ge));
  memcpy(s, c->scalars, c->n * sizeof(sc)); memcpy(p, c->points, c->n * sizeof(ge));
  s[c->n] = c->g_scalar; p[c->n] = *xo_G(); s[c->n + 1] = c->h_scalar; p[c->n + 1] = *xo_H();
  ge r; ge_msm_vartime(&r, s, p, n); free(s); free(p);
  if (out_enc) ristretto_encode(out_enc, &r);
  return ge_ristretto_is_identity(&r);
}
static void msm2(ge *r, const sc *a, const ge *A, const sc *b, const ge *B) { sc s[2] = { *a, *b }; ge p[2] = { *A, *B }; ge_msm_vartime(r, s, p, 2); }
void xo_eq_proof_new(uint8_t out[192], const sc *sk, const ge *P_src, const ge *D_src, const sc *opening, uint64_t amount, xo_transcript *t, xo_rng *rng) {
  xo_transcript_append(t, "dom-sep", "equality-proof", 14);
  sc x, y_s, y_x, y_r, c, z; sc_from_u64(&x, amount); xo_rng_scalar(rng, &y_s); xo_rng_scalar(rng, &y_x); xo_rng_scalar(rng, &y_r);
  ge Y; ge_scalarmult(&Y, &y_s, P_src); ristretto_encode(out, &Y);
  msm2(&Y, &y_x, xo_G(), &y_s, D_src); ristretto_encode(out + 32, &Y);
  msm2(&Y, &y_x, xo_G(), &y_r, xo_H()); ristretto_encode(out + 64, &Y);
  xo_transcript_append(t, "Y_0", out, 32); xo_transcript_append(t, "Y_1", out + 32, 32); xo_transcript_append(t, "Y_2", out + 64, 32);
  xo_challenge_scalar(t, "c", &c);
  sc_muladd(&z, &c, sk, &y_s); sc_tobytes(out + 96, &z); sc_muladd(&z, &c, &x, &y_x); sc_tobytes(out + 128, &z); sc_muladd(&z, &c, opening, &y_r); sc_tobytes(out + 160, &z);
  xo_transcript_append(t, "z_s", out + 96, 32); xo_transcript_append(t, "z_x", out + 128, 32); xo_transcript_append(t, "z_r", out + 160, 32);
  sc w; xo_challenge_scalar(t, "w", &w);
}
int xo_eq_proof_pre_verify(const uint8_t pr[192], const ge *P_src, const ge *C_src, const ge *D_src, const ge *C_dst, xo_transcript *t, xo_collector *col, xo_rng *rng) {
  xo_transcript_append(t, "dom-sep", "equality-proof", 14);
  if (!xo_validate_and_append_point(t, "Y_0", pr) || !xo_validate_and_append_point(t, "Y_1", pr + 32) || !xo_validate_and_append_point(t, "Y_2", pr + 64)) return XO_ERR_TRANSCRIPT;
  sc c, w, ww, z_s, z_x, z_r; xo_challenge_scalar(t, "c", &c);
  xo_transcript_append(t, "z_s", pr + 96, 32); xo_transcript_append(t, "z_x", pr + 128, 32); xo_transcript_append(t, "z_r", pr + 160, 32);
  xo_challenge_scalar(t, "w", &w); sc_mul(&ww, &w, &w);
  /* the reference holds canonical `Scalar`s (serde rejects others); treat non-canonical as a parse error upstream */
  sc_frombytes_mod_order(&z_s, pr + 96); sc_frombytes_mod_order(&z_x, pr + 128); sc_frombytes_mod_order(&z_r, pr + 160);
  ge Y0, Y1, Y2;
  if (!ristretto_decode(&Y0, pr) || !ristretto_decode(&Y1, pr + 32) || !ristretto_decode(&Y2, pr + 64)) return XO_ERR_COMMITMENT_EQ_PROOF;
  sc bf, t1, t2, nw, nww, one, none; xo_rng_scalar(rng, &bf);
  sc_neg(&nw, &w); sc_neg(&nww, &ww); sc_1(&one); sc_neg(&none, &one);
  sc_mul(&t1, &w, &z_x); sc_mul(&t2, &ww, &z_x); sc_add(&t1, &t1, &t2); sc_mul(&t1, &t1, &bf); sc_add(&col->g_scalar, &col->g_scalar, &t1);
  sc_mul(&t1, &ww, &z_r); sc_sub(&t1, &t1, &c); sc_mul(&t1, &t1, &bf); sc_add(&col->h_scalar, &col->h_scalar, &t1);
  sc s[7]; s[0] = z_s; s[1] = none; sc_mul(&s[2], &w, &z_s); sc_mul(&s[3], &nw, &c); s[4] = nw; sc_mul(&s[5], &nww, &c); s[6] = nww;
  const ge *p[7] = { P_src, &Y0, D_src, C_src, &Y1, C_dst, &Y2 };
  for (int i = 0; i < 7; i++) { sc_mul(&s[i], &s[i], &bf); xo_collector_push(col, &s[i], p[i]); }
  return XO_OK;
}
void xo_validity_proof_new(uint8_t out[160], const ge *P_dest, const ge *P_src, uint64_t amount, const sc *opening, xo_transcript *t, xo_rng *rng) {
  xo_transcript_append(t, "dom-sep", "validity-proof", 14);
  sc x, y_r, y_x, c, z; sc_from_u64(&x, amount); xo_rng_scalar(rng, &y_r); xo_rng_scalar(rng, &y_x);
  ge Y; msm2(&Y, &y_r, xo_H(), &y_x, xo_G()); ristretto_encode(out, &Y);
  ge_scalarmult(&Y, &y_r, P_dest); ristretto_encode(out + 32, &Y); ge_scalarmult(&Y, &y_r, P_src); ristretto_encode(out + 64, &Y);
  xo_transcript_append(t, "Y_0", out, 32); xo_transcript_append(t, "Y_1", out + 32, 32); xo_transcript_append(t, "Y_2", out + 64, 32);
  xo_challenge_scalar(t, "c", &c);
  sc_muladd(&z, &c, opening, &y_r); sc_tobytes(out + 96, &z); sc_muladd(&z, &c, &x, &y_x); sc_tobytes(out + 128, &z);
  xo_transcript_append(t, "z_r", out + 96, 32); xo_transcript_append(t, "z_x", out + 128, 32);
  sc w; xo_challenge_scalar(t, "w", &w);
}
int xo_validity_proof_pre_verify(const uint8_t pr[160], const ge *C, const ge *P_dest, const ge *P_src, const ge *D_dest, const ge *D_src, xo_transcript *t, xo_collector *col, xo_rng *rng) {
  xo_transcript_append(t, "dom-sep", "validity-proof", 14);
  if (!xo_validate_and_append_point(t, "Y_0", pr) || !xo_validate_and_append_point(t, "Y_1", pr + 32) || !xo_validate_and_append_point(t, "Y_2", pr + 64)) return XO_ERR_TRANSCRIPT;
  sc c, w, z_r, z_x; xo_challenge_scalar(t, "c", &c);
  xo_transcript_append(t, "z_r", pr + 96, 32); xo_transcript_append(t, "z_x", pr + 128, 32);
  xo_challenge_scalar(t, "w", &w);
  sc_frombytes_mod_order(&z_r, pr + 96); sc_frombytes_mod_order(&z_x, pr + 128);
  ge Y0, Y1, Y2;
  if (!ristretto_decode(&Y0, pr) || !ristretto_decode(&Y1, pr + 32) || !ristretto_decode(&Y2, pr + 64)) return XO_ERR_CT_VALIDITY_PROOF;
  sc bf, t1, nw, one, none, nc, wzr, nwc; xo_rng_scalar(rng, &bf);
  sc_neg(&nw, &w); sc_1(&one); sc_neg(&none, &one); sc_neg(&nc, &c);
  sc_mul(&t1, &z_x, &bf); sc_add(&col->g_scalar, &col->g_scalar, &t1);
  sc_mul(&t1, &z_r, &bf); sc_add(&col->h_scalar, &col->h_scalar, &t1);
  sc_mul(&wzr, &w, &z_r); sc_mul(&nwc, &nw, &c);
  sc s[8]; s[0] = nc; s[1] = none; s[2] = wzr; s[3] = nwc; s[4] = nw; sc_mul(&s[5], &w, &wzr); sc_mul(&s[6], &w, &nwc); sc_mul(&s[7], &w, &nw);
  const ge *p[8] = { C, &Y0, P_dest, D_dest, &Y1, P_src, D_src, &Y2 };
  for (int i = 0; i < 8; i++) { sc_mul(&s[i], &s[i], &bf); xo_collector_push(col, &s[i], p[i]); }
  return XO_OK;
}
static void hash_and_point_to_scalar(sc *out, const uint8_t pk_enc[32], const uint8_t *msg, size_t n, const uint8_t r_enc[32]) { /* src/elgamal.rs:53-65 */
  xo_sponge sp; uint8_t h[64]; xo_sponge_init(&sp, 72); xo_sponge_absorb(&sp, pk_enc, 32); xo_sponge_absorb(&sp, msg, n); xo_sponge_absorb(&sp, r_enc, 32);
  xo_sponge_finish(&sp, 0x06); xo_sponge_squeeze(&sp, h, 64); sc_frombytes_wide(out, h);
}
void xo_sign(uint8_t sig[64], const sc *sk, const uint8_t pk_enc[32], const uint8_t *msg, size_t n, xo_rng *rng) { /* src/elgamal.rs:194-200 */
  sc k, e, s, inv; xo_rng_scalar(rng, &k); ge r; ge_scalarmult(&r, &k, xo_H()); uint8_t r_enc[32]; ristretto_encode(r_enc, &r);
  hash_and_point_to_scalar(&e, pk_enc, msg, n, r_enc); sc_invert(&inv, sk); sc_muladd(&s, &inv, &e, &k); sc_tobytes(sig, &s); sc_tobytes(sig + 32, &e);
}
int xo_sig_verify(const uint8_t sig[64], const uint8_t *msg, size_t n, const ge *pk, uint8_t r_enc_out[32]) { /* src/elgamal.rs:38-42 */
  sc s, e, ne, e2; sc_frombytes_mod_order(&s, sig); sc_frombytes_mod_order(&e, sig + 32); sc_neg(&ne, &e);
  ge a, b, r; ge_scalarmult(&a, &s, xo_H()); ge_scalarmult(&b, &ne, pk); ge_add(&r, &a, &b);
  uint8_t pk_enc[32], r_enc[32]; ristretto_encode(pk_enc, pk); ristretto_encode(r_enc, &r); if (r_enc_out) memcpy(r_enc_out, r_enc, 32);
  hash_and_point_to_scalar(&e2, pk_enc, msg, n, r_enc); return sc_eq(&e, &e2);
}
