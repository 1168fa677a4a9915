/* oracle/tx.c -- TEST INFRASTRUCTURE.  See tx.h. */
#include "tx.h"
#include <stdlib.h>
#include <string.h>
void xo_blake3(const uint8_t *in, size_t n, uint8_t out[32]);
static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
static void wr32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static void wr64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static void be64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * (7 - i))); }
static const uint8_t ZERO32[32] = {0};
/* ---------------------------------------------------------------- growable byte buffer */
typedef struct { uint8_t *p; size_t n, cap; } buf;
static void buf_put(buf *b, const void *d, size_t n) { if (b->n + n > b->cap) { b->cap = (b->n + n) * 2 + 64; b->p = realloc(b->p, b->cap); } memcpy(b->p + b->n, d, n); b->n += n; }
static void buf_u8(buf *b, uint8_t v) { buf_put(b, &v, 1); }
static void buf_be64(buf *b, uint64_t v) { uint8_t t[8]; be64(t, v); buf_put(b, t, 8); }
static void buf_le32(buf *b, uint32_t v) { uint8_t t[4]; wr32(t, v); buf_put(b, t, 4); }
static void buf_le64(buf *b, uint64_t v) { uint8_t t[8]; wr64(t, v); buf_put(b, t, 8); }
/* ---------------------------------------------------------------- parse */
int xo_tx_parse(xo_tx *tx, const uint8_t *blob, size_t len) {
  memset(tx, 0, sizeof *tx); if (len < 64 + 64) return XO_ERR_PARSE;
  tx->blob = blob; tx->len = len; tx->version = blob[0]; tx->type = blob[1]; tx->n_sc = blob[2]; tx->n_ms = blob[3] == 0xFF ? -1 : blob[3];
  tx->count = rd32(blob + 4); tx->aux = rd32(blob + 8); tx->rp_len = rd32(blob + 12); tx->source = blob + 16; tx->fee = rd64(blob + 48); tx->nonce = rd64(blob + 56);
  if (tx->type > XO_TX_MULTISIG) return XO_ERR_PARSE;
  size_t off = 64; const uint8_t *end = blob + len; tx->body = blob + off;
#define NEED(k) do { if ((size_t)(end - (blob + off)) < (size_t)(k)) { xo_tx_free(tx); return XO_ERR_PARSE; } } while (0)
  switch (tx->type) {
  case XO_TX_TRANSFERS:
    if (tx->count > 65535) return XO_ERR_PARSE;
    tx->transfers = calloc(tx->count ? tx->count : 1, sizeof(xo_transfer));
    for (uint32_t i = 0; i < tx->count; i++) {
      NEED(5 * 32 + 160 + 4); xo_transfer *t = &tx->transfers[i]; const uint8_t *p = blob + off;
      t->asset = p; t->dest = p + 32; t->commitment = p + 64; t->sender_handle = p + 96; t->receiver_handle = p + 128; t->proof = p + 160; off += 320;
      uint32_t el = rd32(blob + off); off += 4; t->has_extra = el != 0xFFFFFFFFu; t->extra_len = t->has_extra ? el : 0; NEED(t->extra_len); t->extra = blob + off; off += t->extra_len;
      for (int k = 0; k < 2; k++) { sc s; if (!sc_frombytes_canonical(&s, t->proof + 96 + 32 * k)) { xo_tx_free(tx); return XO_ERR_PARSE; } }
    } break;
  case XO_TX_BURN: NEED(40); off += 40; break;
  case XO_TX_CALL: NEED(32); off += 32; if (tx->count > 65535 || tx->aux > 65535) return XO_ERR_PARSE; NEED((size_t)tx->count * 40); off += (size_t)tx->count * 40;
    for (uint32_t i = 0; i < tx->aux * 2; i++) { NEED(4); uint32_t l = rd32(blob + off); off += 4; NEED(l); off += l; } break;
  case XO_TX_DEPLOY: NEED(tx->aux); off += tx->aux; break;
  case XO_TX_MULTISIG: if (tx->count > 255 || tx->aux > 255) return XO_ERR_PARSE; NEED((size_t)tx->count * 32); off += (size_t)tx->count * 32; break;
  }
  tx->body_len = off - 64;
  NEED(tx->rp_len); tx->rp = blob + off; off += tx->rp_len;
  NEED((size_t)tx->n_sc * 256); tx->sc = blob + off; off += (size_t)tx->n_sc * 256;
  for (int i = 0; i < tx->n_sc; i++) for (int k = 0; k < 3; k++) { sc s; if (!sc_frombytes_canonical(&s, tx->sc + 256 * i + 64 + 96 + 32 * k)) { xo_tx_free(tx); return XO_ERR_PARSE; } }
  if (tx->n_ms > 0) { NEED((size_t)tx->n_ms * 65); tx->ms = blob + off; off += (size_t)tx->n_ms * 65; }
  NEED(64); tx->sig = blob + off; off += 64;
  if (off != len) { xo_tx_free(tx); return XO_ERR_PARSE; }
  /* RangeProof::from_bytes structural rules (FormatError in serde, i.e. before verify is reachable) */
  if (tx->rp_len % 32 || tx->rp_len < 9 * 32 || ((tx->rp_len / 32 - 9) & 1) || (tx->rp_len / 32 - 9) / 2 >= 32) { xo_tx_free(tx); return XO_ERR_PARSE; }
  { sc s; const uint8_t *r = tx->rp; if (!sc_frombytes_canonical(&s, r + 128) || !sc_frombytes_canonical(&s, r + 160) || !sc_frombytes_canonical(&s, r + 192) ||
      !sc_frombytes_canonical(&s, r + tx->rp_len - 64) || !sc_frombytes_canonical(&s, r + tx->rp_len - 32)) { xo_tx_free(tx); return XO_ERR_PARSE; } }
  { sc s; if (!sc_frombytes_canonical(&s, tx->sig) || !sc_frombytes_canonical(&s, tx->sig + 32)) { xo_tx_free(tx); return XO_ERR_PARSE; }
    for (int i = 0; i < tx->n_ms; i++) if (!sc_frombytes_canonical(&s, tx->ms + 65 * i + 1) || !sc_frombytes_canonical(&s, tx->ms + 65 * i + 33)) { xo_tx_free(tx); return XO_ERR_PARSE; } }
  return XO_OK;
}
void xo_tx_free(xo_tx *tx) { free(tx->transfers); tx->transfers = NULL; }
size_t xo_tx_to_bytes(const xo_tx *tx, uint8_t **out, size_t *multisig_index) {
  buf b = {0}; buf_u8(&b, tx->version); buf_put(&b, tx->source, 32); buf_be64(&b, tx->fee); buf_be64(&b, tx->nonce);
  const uint8_t *p = tx->body;
  switch (tx->type) {
  case XO_TX_TRANSFERS: for (uint32_t i = 0; i < tx->count; i++) { const xo_transfer *t = &tx->transfers[i]; buf_put(&b, t->asset, 160); if (t->has_extra) buf_put(&b, t->extra, t->extra_len); buf_put(&b, t->proof, 160); } break;
  case XO_TX_BURN: buf_put(&b, p, 32); buf_be64(&b, rd64(p + 32)); break;
  case XO_TX_CALL: buf_put(&b, p, 32); p += 32; for (uint32_t i = 0; i < tx->count; i++) { buf_put(&b, p, 32); buf_be64(&b, rd64(p + 32)); p += 40; }
    for (uint32_t i = 0; i < tx->aux * 2; i++) { uint32_t l = rd32(p); buf_put(&b, p + 4, l); p += 4 + l; } break;
  case XO_TX_DEPLOY: buf_put(&b, p, tx->aux); break;
  case XO_TX_MULTISIG: buf_u8(&b, (uint8_t)tx->aux); buf_put(&b, p, (size_t)tx->count * 32); break;
  }
  buf_put(&b, tx->rp, tx->rp_len);
  buf_put(&b, tx->sc, (size_t)tx->n_sc * 256);
  if (multisig_index) *multisig_index = b.n;
  if (tx->n_ms >= 0) buf_put(&b, tx->ms, (size_t)tx->n_ms * 65);
  *out = b.p; return b.n;
}
/* ---------------------------------------------------------------- mock ledger: open-addressing tables */
typedef struct { uint8_t key[64]; uint8_t val[64]; uint8_t *ext; int ext_n; uint8_t used; } slot;
typedef struct { slot *s; size_t cap, n; } tbl;
struct xo_ledger { tbl bal, nonce, ms, out; int record_out; /* out: last set_output_ciphertext per (account, asset), src/tx/verify.rs:339-340 -- kept only on request: the reference hands the ciphertext over uncompressed and its mock drops it, so the CPU baseline must not pay two encodings for it */ };
static uint64_t khash(const uint8_t k[64]) { uint64_t h = 1469598103934665603ULL; for (int i = 0; i < 64; i++) { h ^= k[i]; h *= 1099511628211ULL; } return h; }
static slot *tbl_find(const tbl *t, const uint8_t k[64]) { if (!t->cap) return NULL; size_t i = khash(k) & (t->cap - 1); while (t->s[i].used) { if (!memcmp(t->s[i].key, k, 64)) return &t->s[i]; i = (i + 1) & (t->cap - 1); } return NULL; }
static slot *tbl_put(tbl *t, const uint8_t k[64]) {
  if ((t->n + 1) * 2 > t->cap) { tbl nt = { calloc(t->cap ? t->cap * 2 : 64, sizeof(slot)), t->cap ? t->cap * 2 : 64, 0 };
    for (size_t i = 0; i < t->cap; i++) if (t->s[i].used) { size_t j = khash(t->s[i].key) & (nt.cap - 1); while (nt.s[j].used) j = (j + 1) & (nt.cap - 1); nt.s[j] = t->s[i]; nt.n++; } free(t->s); *t = nt; }
  slot *s = tbl_find(t, k); if (s) return s; size_t i = khash(k) & (t->cap - 1); while (t->s[i].used) i = (i + 1) & (t->cap - 1);
  memset(&t->s[i], 0, sizeof(slot)); memcpy(t->s[i].key, k, 64); t->s[i].used = 1; t->n++; return &t->s[i];
}
static void mk(uint8_t k[64], const uint8_t a[32], const uint8_t b[32]) { memcpy(k, a, 32); if (b) memcpy(k + 32, b, 32); else memset(k + 32, 0, 32); }
xo_ledger *xo_ledger_new(void) { return calloc(1, sizeof(xo_ledger)); }
static void tbl_clone(tbl *d, const tbl *s) { *d = *s; if (s->cap) { d->s = malloc(s->cap * sizeof(slot)); memcpy(d->s, s->s, s->cap * sizeof(slot)); for (size_t i = 0; i < s->cap; i++) if (s->s[i].used && s->s[i].ext) { d->s[i].ext = malloc(s->s[i].ext_n * 32); memcpy(d->s[i].ext, s->s[i].ext, s->s[i].ext_n * 32); } } }
xo_ledger *xo_ledger_clone(const xo_ledger *l) { xo_ledger *c = calloc(1, sizeof *c); tbl_clone(&c->bal, &l->bal); tbl_clone(&c->nonce, &l->nonce); tbl_clone(&c->ms, &l->ms); tbl_clone(&c->out, &l->out); c->record_out = l->record_out; return c; }
void xo_ledger_free(xo_ledger *l) { if (!l) return; for (size_t i = 0; i < l->ms.cap; i++) if (l->ms.s[i].used) free(l->ms.s[i].ext); free(l->bal.s); free(l->nonce.s); free(l->ms.s); free(l->out.s); free(l); }
void xo_ledger_set_balance(xo_ledger *l, const uint8_t pk[32], const uint8_t asset[32], const uint8_t ct[64]) { uint8_t k[64]; mk(k, pk, asset); memcpy(tbl_put(&l->bal, k)->val, ct, 64); }
int xo_ledger_get_balance(const xo_ledger *l, const uint8_t pk[32], const uint8_t asset[32], uint8_t ct[64]) { uint8_t k[64]; mk(k, pk, asset); slot *s = tbl_find(&l->bal, k); if (!s) return 0; memcpy(ct, s->val, 64); return 1; }
void xo_ledger_set_nonce(xo_ledger *l, const uint8_t pk[32], uint64_t nonce) { uint8_t k[64]; mk(k, pk, NULL); wr64(tbl_put(&l->nonce, k)->val, nonce); }
int xo_ledger_get_nonce(const xo_ledger *l, const uint8_t pk[32], uint64_t *nonce) { uint8_t k[64]; mk(k, pk, NULL); slot *s = tbl_find(&l->nonce, k); if (!s) return 0; *nonce = rd64(s->val); return 1; }
void xo_ledger_set_multisig(xo_ledger *l, const uint8_t pk[32], const uint8_t *signers, int n, uint8_t threshold) { /* src/lib.rs:177-189: empty signer list removes */
  uint8_t k[64]; mk(k, pk, NULL); slot *s = tbl_put(&l->ms, k); free(s->ext); s->ext = NULL; s->ext_n = n; s->val[0] = threshold; s->val[1] = n > 0;
  if (n > 0) { s->ext = malloc(n * 32); memcpy(s->ext, signers, n * 32); }
}
int xo_ledger_get_multisig(const xo_ledger *l, const uint8_t pk[32], const uint8_t **signers, int *n, uint8_t *threshold) {
  uint8_t k[64]; mk(k, pk, NULL); slot *s = tbl_find(&l->ms, k); if (!s || !s->val[1]) return 0; *signers = s->ext; *n = s->ext_n; *threshold = s->val[0]; return 1;
}
/* multisig settings as records pk[32] n[1] threshold[1] signers[n x 32] (test plumbing: copy a minted batch's settings into another ledger) */
size_t xo_ledger_dump_multisig(const xo_ledger *l, uint8_t *out, size_t cap) {
  size_t o = 0; for (size_t i = 0; i < l->ms.cap; i++) if (l->ms.s[i].used && l->ms.s[i].val[1]) { size_t need = 34 + 32 * (size_t)l->ms.s[i].ext_n;
    if (o + need <= cap) { memcpy(out + o, l->ms.s[i].key, 32); out[o + 32] = (uint8_t)l->ms.s[i].ext_n; out[o + 33] = l->ms.s[i].val[0]; memcpy(out + o + 34, l->ms.s[i].ext, 32 * (size_t)l->ms.s[i].ext_n); } o += need; }
  return o;
}
static int cmp128(const void *a, const void *b) { return memcmp(a, b, 64); }
/* set_output_ciphertext (src/tx/verify.rs:60-66): the mock of the reference ignores it (src/lib.rs:166-175); recorded here
 * (compressed) so that the CUDA path's output ciphertexts can be checked byte for byte */
void xo_ledger_record_outputs(xo_ledger *l, int on) { l->record_out = on; }
void xo_ledger_set_output(xo_ledger *l, const uint8_t pk[32], const uint8_t asset[32], const uint8_t ct[64]) { uint8_t k[64]; mk(k, pk, asset); memcpy(tbl_put(&l->out, k)->val, ct, 64); }
size_t xo_ledger_dump_outputs(const xo_ledger *l, uint8_t *out, size_t cap) {
  size_t n = 0; for (size_t i = 0; i < l->out.cap; i++) if (l->out.s[i].used) { if ((n + 1) * 128 <= cap) { memcpy(out + n * 128, l->out.s[i].key, 64); memcpy(out + n * 128 + 64, l->out.s[i].val, 64); } n++; }
  if (n * 128 <= cap) qsort(out, n, 128, cmp128); return n;
}
size_t xo_ledger_dump(const xo_ledger *l, uint8_t *out, size_t cap) {
  size_t n = 0; for (size_t i = 0; i < l->bal.cap; i++) if (l->bal.s[i].used) { if ((n + 1) * 128 <= cap) { memcpy(out + n * 128, l->bal.s[i].key, 64); memcpy(out + n * 128 + 64, l->bal.s[i].val, 64); } n++; }
  if (n * 128 <= cap) qsort(out, n, 128, cmp128); return n;
}
/* ---------------------------------------------------------------- ciphertext algebra (src/elgamal.rs:322-377) */
typedef struct { ge C, D; } ct_t;
static int ct_decode(ct_t *c, const uint8_t e[64]) { return ristretto_decode(&c->C, e) && ristretto_decode(&c->D, e + 32); }
static void ct_encode(uint8_t e[64], const ct_t *c) { ristretto_encode(e, &c->C); ristretto_encode(e + 32, &c->D); }
static void ct_zero(ct_t *c) { ge_identity(&c->C); ge_identity(&c->D); }
static void ct_add_amount(ct_t *c, uint64_t amount) { sc s; sc_from_u64(&s, amount); ge t; ge_scalarmult(&t, &s, xo_G()); ge_add(&c->C, &c->C, &t); } /* `&G * o`: variable-base */
typedef struct { ge commitment, sender_handle, receiver_handle; } dtransfer;
/* src/tx/verify.rs:107-144 */
static void sender_output_ct(ct_t *out, const xo_tx *tx, const uint8_t asset[32], const dtransfer *dt) {
  ct_zero(out);
  if (!memcmp(asset, ZERO32, 32)) ct_add_amount(out, tx->fee);
  if (tx->type == XO_TX_TRANSFERS) { for (uint32_t i = 0; i < tx->count; i++) if (!memcmp(asset, tx->transfers[i].asset, 32)) { ge_add(&out->C, &out->C, &dt[i].commitment); ge_add(&out->D, &out->D, &dt[i].sender_handle); } }
  else if (tx->type == XO_TX_BURN) { if (!memcmp(asset, tx->body, 32)) ct_add_amount(out, rd64(tx->body + 32)); }
  else if (tx->type == XO_TX_CALL) { /* HashMap::get: one entry per key; with duplicate wire keys the last insert wins */
    const uint8_t *hit = NULL; for (uint32_t i = 0; i < tx->count; i++) if (!memcmp(asset, tx->body + 32 + 40 * i, 32)) hit = tx->body + 32 + 40 * i; if (hit) ct_add_amount(out, rd64(hit + 32)); }
}
static int has_commitment_for(const xo_tx *tx, const uint8_t asset[32]) { for (int i = 0; i < tx->n_sc; i++) if (!memcmp(tx->sc + 256 * i, asset, 32)) return 1; return 0; }
static int verify_commitment_assets(const xo_tx *tx) { /* src/tx/verify.rs:161-199 */
  if (!has_commitment_for(tx, ZERO32)) return 0;
  for (int i = 0; i < tx->n_sc; i++) for (int j = 0; j < tx->n_sc; j++) if (i != j && !memcmp(tx->sc + 256 * i, tx->sc + 256 * j, 32)) return 0;
  if (tx->type == XO_TX_TRANSFERS) { for (uint32_t i = 0; i < tx->count; i++) if (!has_commitment_for(tx, tx->transfers[i].asset)) return 0; }
  else if (tx->type == XO_TX_BURN) return has_commitment_for(tx, tx->body);
  else if (tx->type == XO_TX_CALL) { for (uint32_t i = 0; i < tx->count; i++) if (!has_commitment_for(tx, tx->body + 32 + 40 * i)) return 0; }
  return 1;
}
static void prepare_transcript(xo_transcript *t, const xo_tx *tx) { /* src/tx/verify.rs:146-158 */
  xo_transcript_init(t, "transaction-proof"); xo_transcript_append_u64(t, "version", tx->version); xo_transcript_append(t, "source_pubkey", tx->source, 32);
  xo_transcript_append_u64(t, "fee", tx->fee); xo_transcript_append_u64(t, "nonce", tx->nonce);
}
typedef struct { xo_transcript t; uint8_t *commit_enc; ge *commit_pts; int m; } prepared;
/* src/tx/verify.rs:203-485 */
static int pre_verify(const xo_tx *tx, xo_ledger *st, xo_collector *col, xo_rng *rng, prepared *out) {
  uint64_t nonce; int rc = XO_OK;
  if (!xo_ledger_get_nonce(st, tx->source, &nonce)) return XO_ERR_STATE;
  if (nonce != tx->nonce) return XO_ERR_INVALID_NONCE;
  xo_ledger_set_nonce(st, tx->source, tx->nonce);
  if (!verify_commitment_assets(tx)) return XO_ERR_FORMAT;
  uint32_t k = tx->type == XO_TX_TRANSFERS ? tx->count : 0;
  dtransfer *dt = malloc(sizeof(dtransfer) * (k + 1)); ge *nsc = malloc(sizeof(ge) * (tx->n_sc + 1)); uint8_t *bytes = NULL;
  for (uint32_t i = 0; i < k; i++) if (!ristretto_decode(&dt[i].commitment, tx->transfers[i].commitment) || !ristretto_decode(&dt[i].sender_handle, tx->transfers[i].sender_handle) ||
      !ristretto_decode(&dt[i].receiver_handle, tx->transfers[i].receiver_handle)) { rc = XO_ERR_DECOMPRESSION; goto done; }
  for (int i = 0; i < tx->n_sc; i++) if (!ristretto_decode(&nsc[i], tx->sc + 256 * i + 32)) { rc = XO_ERR_DECOMPRESSION; goto done; }
  ge src; if (!ristretto_decode(&src, tx->source)) { rc = XO_ERR_DECOMPRESSION; goto done; }
  prepare_transcript(&out->t, tx);
  size_t ms_index; size_t nbytes = xo_tx_to_bytes(tx, &bytes, &ms_index);
  if (!xo_sig_verify(tx->sig, bytes, nbytes, &src, NULL)) { rc = XO_ERR_SIGNATURE; goto done; }
  { const uint8_t *signers; int n_signers; uint8_t threshold;
    if (xo_ledger_get_multisig(st, tx->source, &signers, &n_signers, &threshold)) {
      if (tx->n_ms < 0) { rc = XO_ERR_FORMAT; goto done; }
      if (tx->n_ms == 0 || tx->n_ms != threshold) { rc = XO_ERR_FORMAT; goto done; }
      uint8_t hash[32]; xo_blake3(bytes, ms_index, hash);
      for (int i = 0; i < tx->n_ms; i++) {
        for (int j = 0; j < tx->n_ms; j++) if (i != j && tx->ms[65 * i] == tx->ms[65 * j]) { rc = XO_ERR_FORMAT; goto done; }
        int idx = tx->ms[65 * i];
        if (idx < n_signers) { ge sp; if (!ristretto_decode(&sp, signers + 32 * idx)) { rc = XO_ERR_DECOMPRESSION; goto done; } if (!xo_sig_verify(tx->ms + 65 * i + 1, hash, 32, &sp, NULL)) { rc = XO_ERR_SIGNATURE; goto done; } }
      }
    } else if (tx->n_ms >= 0) { rc = XO_ERR_FORMAT; goto done; } }
  for (int i = 0; i < tx->n_sc; i++) {
    const uint8_t *asset = tx->sc + 256 * i; uint8_t cur[64], enc[64]; ct_t c, o, n;
    if (!xo_ledger_get_balance(st, tx->source, asset, cur)) { rc = XO_ERR_STATE; goto done; }
    if (!ct_decode(&c, cur)) { rc = XO_ERR_DECOMPRESSION; goto done; }
    sender_output_ct(&o, tx, asset, dt); ge_sub(&n.C, &c.C, &o.C); ge_sub(&n.D, &c.D, &o.D);
    xo_transcript_append(&out->t, "dom-sep", "new-commitment-proof", 20); xo_transcript_append(&out->t, "new_source_commitment_asset", asset, 32); xo_transcript_append(&out->t, "new_source_commitment", asset + 32, 32);
    rc = xo_eq_proof_pre_verify(asset + 64, &src, &n.C, &n.D, &nsc[i], &out->t, col, rng); if (rc) goto done;
    ct_encode(enc, &n); xo_ledger_set_balance(st, tx->source, asset, enc);
    if (st->record_out) { ct_encode(enc, &o); xo_ledger_set_output(st, tx->source, asset, enc); }
  }
  if (tx->type == XO_TX_TRANSFERS) {
    for (uint32_t i = 0; i < k; i++) {
      const xo_transfer *t = &tx->transfers[i]; ge dest; uint8_t cur[64], enc[64]; ct_t c;
      if (!ristretto_decode(&dest, t->dest)) { rc = XO_ERR_DECOMPRESSION; goto done; }
      if (!xo_ledger_get_balance(st, t->dest, t->asset, cur)) { rc = XO_ERR_STATE; goto done; }
      if (!ct_decode(&c, cur)) { rc = XO_ERR_DECOMPRESSION; goto done; }
      ge_add(&c.C, &c.C, &dt[i].commitment); ge_add(&c.D, &c.D, &dt[i].receiver_handle); ct_encode(enc, &c); xo_ledger_set_balance(st, t->dest, t->asset, enc);
      xo_transcript_append(&out->t, "dom-sep", "transfer-proof", 14); xo_transcript_append(&out->t, "dest_pubkey", t->dest, 32); xo_transcript_append(&out->t, "amount_commitment", t->commitment, 32);
      xo_transcript_append(&out->t, "amount_sender_handle", t->sender_handle, 32); xo_transcript_append(&out->t, "amount_receiver_handle", t->receiver_handle, 32);
      rc = xo_validity_proof_pre_verify(t->proof, &dt[i].commitment, &dest, &src, &dt[i].receiver_handle, &dt[i].sender_handle, &out->t, col, rng); if (rc) goto done;
    }
  } else if (tx->type == XO_TX_BURN) {
    xo_transcript_append(&out->t, "dom-sep", "burn-proof", 10); xo_transcript_append(&out->t, "asset", tx->body, 32); xo_transcript_append_u64(&out->t, "amount", rd64(tx->body + 32));
  } else if (tx->type == XO_TX_MULTISIG) {
    uint32_t ns = tx->count, th = tx->aux;
    if (th > ns || (ns != 0 && th == 0)) { rc = XO_ERR_FORMAT; goto done; }
    for (uint32_t i = 0; i < ns; i++) for (uint32_t j = 0; j < ns; j++) if (i != j && !memcmp(tx->body + 32 * i, tx->body + 32 * j, 32)) { rc = XO_ERR_FORMAT; goto done; }
    for (uint32_t i = 0; i < ns; i++) if (!memcmp(tx->body + 32 * i, tx->source, 32)) { rc = XO_ERR_FORMAT; goto done; }
    xo_transcript_append(&out->t, "dom-sep", "multisig-proof", 14); xo_transcript_append_u64(&out->t, "threshold", th);
    for (uint32_t i = 0; i < ns; i++) xo_transcript_append(&out->t, "signer", tx->body + 32 * i, 32);
    xo_ledger_set_multisig(st, tx->source, tx->body, (int)ns, (uint8_t)th);
  }
  { size_t nc = (size_t)tx->n_sc + k, m = 1; while (m < nc) m <<= 1;  /* next_power_of_two(0) == 1 */
    out->m = (int)m; out->commit_enc = calloc(m, 32); out->commit_pts = malloc(sizeof(ge) * m);
    for (int i = 0; i < tx->n_sc; i++) { memcpy(out->commit_enc + 32 * i, tx->sc + 256 * i + 32, 32); out->commit_pts[i] = nsc[i]; }
    for (uint32_t i = 0; i < k; i++) { memcpy(out->commit_enc + 32 * (tx->n_sc + i), tx->transfers[i].commitment, 32); out->commit_pts[tx->n_sc + i] = dt[i].commitment; }
    for (size_t i = nc; i < m; i++) ge_identity(&out->commit_pts[i]); }
done:
  free(dt); free(nsc); free(bytes); return rc;
}
static void prepared_free(prepared *p) { free(p->commit_enc); free(p->commit_pts); }
/* partial != NULL: multi-GPU shard mode -- skip the two identity decisions and return the partial sums' encodings (sigma || range) */
int xo_verify_batch_ex(const uint8_t *const *blobs, const size_t *lens, size_t n, xo_ledger *state, xo_rng *rng, long *fail_index, uint8_t *partial) {
  xo_collector col; xo_collector_init(&col); prepared *prep = calloc(n ? n : 1, sizeof(prepared)); xo_tx *txs = calloc(n ? n : 1, sizeof(xo_tx)); int rc = XO_OK; size_t done = 0;
  if (fail_index) *fail_index = -1;
  for (size_t i = 0; i < n; i++) { rc = xo_tx_parse(&txs[i], blobs[i], lens[i]); if (rc) { if (fail_index) *fail_index = (long)i; n = i; goto out; } }
  for (size_t i = 0; i < n; i++) { rc = pre_verify(&txs[i], state, &col, rng, &prep[i]); done = i + 1; if (rc) { if (fail_index) *fail_index = (long)i; goto out; } }
  if (partial) memset(partial, 0, 64);
  if (!xo_collector_verify(&col, partial) && !partial) { rc = XO_ERR_GENERIC_PROOF; goto out; }
  { xo_rp_item *items = malloc(sizeof(xo_rp_item) * (n + 1));
    for (size_t i = 0; i < n; i++) { items[i].proof = txs[i].rp; items[i].len = txs[i].rp_len; items[i].t = &prep[i].t; items[i].commit_enc = prep[i].commit_enc; items[i].commit_pts = prep[i].commit_pts; items[i].m = prep[i].m; }
    rc = n ? xo_rp_verify_batch_ex(items, n, rng, partial ? partial + 32 : NULL, partial != NULL) : XO_OK; free(items); }
out:
  for (size_t i = 0; i < done; i++) prepared_free(&prep[i]); for (size_t i = 0; i < n; i++) xo_tx_free(&txs[i]);
  free(prep); free(txs); xo_collector_free(&col); return rc;
}
int xo_verify_batch(const uint8_t *const *blobs, const size_t *lens, size_t n, xo_ledger *state, xo_rng *rng, long *fail_index) { return xo_verify_batch_ex(blobs, lens, n, state, rng, fail_index, NULL); }
int xo_verify(const uint8_t *blob, size_t len, xo_ledger *state, xo_rng *rng) { long fi; return xo_verify_batch(&blob, &len, 1, state, rng, &fi); }
int xo_apply_without_verify(const uint8_t *blob, size_t len, xo_ledger *st) { /* src/tx/verify.rs:545-619 */
  xo_tx tx; int rc = xo_tx_parse(&tx, blob, len); if (rc) return rc;
  uint32_t k = tx.type == XO_TX_TRANSFERS ? tx.count : 0; dtransfer *dt = malloc(sizeof(dtransfer) * (k + 1));
  for (uint32_t i = 0; i < k; i++) if (!ristretto_decode(&dt[i].commitment, tx.transfers[i].commitment) || !ristretto_decode(&dt[i].sender_handle, tx.transfers[i].sender_handle) ||
      !ristretto_decode(&dt[i].receiver_handle, tx.transfers[i].receiver_handle)) { rc = XO_ERR_DECOMPRESSION; goto done; }
  for (int i = 0; i < tx.n_sc; i++) { const uint8_t *asset = tx.sc + 256 * i; uint8_t cur[64], enc[64]; ct_t c, o, n;
    if (!xo_ledger_get_balance(st, tx.source, asset, cur)) { rc = XO_ERR_STATE; goto done; } if (!ct_decode(&c, cur)) { rc = XO_ERR_DECOMPRESSION; goto done; }
    sender_output_ct(&o, &tx, asset, dt); ge_sub(&n.C, &c.C, &o.C); ge_sub(&n.D, &c.D, &o.D); ct_encode(enc, &n); xo_ledger_set_balance(st, tx.source, asset, enc); if (st->record_out) { ct_encode(enc, &o); xo_ledger_set_output(st, tx.source, asset, enc); } }
  for (uint32_t i = 0; i < k; i++) { const xo_transfer *t = &tx.transfers[i]; uint8_t cur[64], enc[64]; ct_t c;
    if (!xo_ledger_get_balance(st, t->dest, t->asset, cur)) { rc = XO_ERR_STATE; goto done; } if (!ct_decode(&c, cur)) { rc = XO_ERR_DECOMPRESSION; goto done; }
    ge_add(&c.C, &c.C, &dt[i].commitment); ge_add(&c.D, &c.D, &dt[i].receiver_handle); ct_encode(enc, &c); xo_ledger_set_balance(st, t->dest, t->asset, enc); }
  if (tx.type == XO_TX_MULTISIG) xo_ledger_set_multisig(st, tx.source, tx.body, (int)tx.count, (uint8_t)tx.aux);
done: free(dt); xo_tx_free(&tx); return rc;
}
/* ---------------------------------------------------------------- keys, encryption, builder */
void xo_pubkey_from_secret(const sc *sk, uint8_t pk[32], ge *P) { sc inv; sc_invert(&inv, sk); ge p; ge_scalarmult(&p, &inv, xo_H()); if (pk) ristretto_encode(pk, &p); if (P) *P = p; } /* src/elgamal.rs:102-107 */
void xo_keygen(xo_rng *rng, sc *sk, uint8_t pk[32]) { xo_rng_scalar(rng, sk); xo_pubkey_from_secret(sk, pk, NULL); }
void xo_encrypt(uint8_t ct[64], const ge *P, uint64_t amount, const sc *opening) { /* src/elgamal.rs:116-130,228-230,266-271 */
  sc x; sc_from_u64(&x, amount); sc s2[2] = { x, *opening }; ge p2[2] = { *xo_G(), *xo_H() }; ge C, D; ge_msm_vartime(&C, s2, p2, 2); ge_scalarmult(&D, opening, P);
  ristretto_encode(ct, &C); ristretto_encode(ct + 32, &D);
}
static uint64_t tx_cost(const xo_tx_spec *s, const uint8_t asset[32]) { /* src/tx/builder.rs:261-294 */
  uint64_t c = 0; if (!memcmp(asset, ZERO32, 32)) c += s->fee;
  if (s->type == XO_TX_TRANSFERS) { for (uint32_t i = 0; i < s->n_transfers; i++) if (!memcmp(s->transfers[i].asset, asset, 32)) c += s->transfers[i].amount; }
  else if (s->type == XO_TX_BURN) { if (!memcmp(s->burn_asset, asset, 32)) c += s->burn_amount; }
  else if (s->type == XO_TX_CALL) { for (uint32_t i = 0; i < s->n_call_assets; i++) if (!memcmp(s->call_assets + 32 * i, asset, 32)) c += s->call_amounts[i]; }
  return c;
}
size_t xo_tx_build(uint8_t **out, const xo_tx_spec *s, const sc *sk, const xo_ledger *state, xo_rng *rng, const uint8_t *ms_index, const sc *ms_sk, int n_ms) {
  ge P_src; uint8_t pk[32]; xo_pubkey_from_secret(sk, pk, &P_src);
  uint32_t k = s->type == XO_TX_TRANSFERS ? s->n_transfers : 0, a = s->n_assets; size_t nc = a + k, m = 1; while (m < nc) m <<= 1;
  sc *openings = calloc(m, sizeof(sc)); uint64_t *values = calloc(m, sizeof(uint64_t));
  ge *tC = malloc(sizeof(ge) * (k + 1)), *tDs = malloc(sizeof(ge) * (k + 1)), *tDr = malloc(sizeof(ge) * (k + 1)), *Pd = malloc(sizeof(ge) * (k + 1));
  uint8_t *tenc = malloc(96 * (k + 1));
  for (uint32_t i = 0; i < k; i++) { /* src/tx/builder.rs:329-357 */
    const xo_transfer_spec *t = &s->transfers[i]; if (!ristretto_decode(&Pd[i], t->dest)) return 0;
    sc r, x; xo_rng_scalar(rng, &r); sc_from_u64(&x, t->amount); sc s2[2] = { x, r }; ge p2[2] = { *xo_G(), *xo_H() }; ge_msm_vartime(&tC[i], s2, p2, 2);
    ge_scalarmult(&tDs[i], &r, &P_src); ge_scalarmult(&tDr[i], &r, &Pd[i]); openings[a + i] = r; values[a + i] = t->amount;
    ristretto_encode(tenc + 96 * i, &tC[i]); ristretto_encode(tenc + 96 * i + 32, &tDs[i]); ristretto_encode(tenc + 96 * i + 64, &tDr[i]);
  }
  xo_transcript tr; xo_transcript_init(&tr, "transaction-proof"); xo_transcript_append_u64(&tr, "version", s->version); xo_transcript_append(&tr, "source_pubkey", pk, 32);
  xo_transcript_append_u64(&tr, "fee", s->fee); xo_transcript_append_u64(&tr, "nonce", s->nonce);
  buf scb = {0};
  for (uint32_t i = 0; i < a; i++) { /* src/tx/builder.rs:381-421 */
    const uint8_t *asset = s->assets + 32 * i; uint64_t cost = tx_cost(s, asset); if (s->balances[i] < cost) return 0; uint64_t nb = s->balances[i] - cost;
    xo_rng_scalar(rng, &openings[i]); values[i] = nb;
    uint8_t cur[64]; ct_t c; if (!xo_ledger_get_balance(state, pk, asset, cur) || !ct_decode(&c, cur)) return 0;
    sc x; sc_from_u64(&x, nb); sc s2[2] = { x, openings[i] }; ge p2[2] = { *xo_G(), *xo_H() }; ge nC; ge_msm_vartime(&nC, s2, p2, 2); uint8_t nC_enc[32]; ristretto_encode(nC_enc, &nC);
    /* new source ciphertext = current - cost terms (src/tx/builder.rs:222-258) */
    sc amt; uint64_t plain = 0; if (!memcmp(asset, ZERO32, 32)) plain += s->fee;
    if (s->type == XO_TX_BURN && !memcmp(asset, s->burn_asset, 32)) plain += s->burn_amount;
    if (s->type == XO_TX_CALL) for (uint32_t j = 0; j < s->n_call_assets; j++) if (!memcmp(s->call_assets + 32 * j, asset, 32)) plain += s->call_amounts[j];
    sc_from_u64(&amt, plain); ge t; ge_scalarmult(&t, &amt, xo_G()); ge_sub(&c.C, &c.C, &t);
    for (uint32_t j = 0; j < k; j++) if (!memcmp(s->transfers[j].asset, asset, 32)) { ge_sub(&c.C, &c.C, &tC[j]); ge_sub(&c.D, &c.D, &tDs[j]); }
    xo_transcript_append(&tr, "dom-sep", "new-commitment-proof", 20); xo_transcript_append(&tr, "new_source_commitment_asset", asset, 32); xo_transcript_append(&tr, "new_source_commitment", nC_enc, 32);
    uint8_t proof[192]; xo_eq_proof_new(proof, sk, &P_src, &c.D, &openings[i], nb, &tr, rng);
    buf_put(&scb, asset, 32); buf_put(&scb, nC_enc, 32); buf_put(&scb, proof, 192);
  }
  buf body = {0}; uint32_t count = 0, aux = 0;
  switch (s->type) {
  case XO_TX_TRANSFERS: count = k;
    for (uint32_t i = 0; i < k; i++) { const xo_transfer_spec *t = &s->transfers[i];
      xo_transcript_append(&tr, "dom-sep", "transfer-proof", 14); xo_transcript_append(&tr, "dest_pubkey", t->dest, 32); xo_transcript_append(&tr, "amount_commitment", tenc + 96 * i, 32);
      xo_transcript_append(&tr, "amount_sender_handle", tenc + 96 * i + 32, 32); xo_transcript_append(&tr, "amount_receiver_handle", tenc + 96 * i + 64, 32);
      uint8_t proof[160]; xo_validity_proof_new(proof, &Pd[i], &P_src, t->amount, &openings[a + i], &tr, rng);
      buf_put(&body, t->asset, 32); buf_put(&body, t->dest, 32); buf_put(&body, tenc + 96 * i, 96); buf_put(&body, proof, 160);
      buf_le32(&body, t->has_extra ? t->extra_len : 0xFFFFFFFFu); if (t->has_extra) buf_put(&body, t->extra, t->extra_len); } break;
  case XO_TX_BURN: xo_transcript_append(&tr, "dom-sep", "burn-proof", 10); xo_transcript_append(&tr, "asset", s->burn_asset, 32); xo_transcript_append_u64(&tr, "amount", s->burn_amount);
    buf_put(&body, s->burn_asset, 32); buf_le64(&body, s->burn_amount); break;
  case XO_TX_CALL: count = s->n_call_assets; aux = s->n_params; buf_put(&body, s->contract, 32);
    for (uint32_t i = 0; i < count; i++) { buf_put(&body, s->call_assets + 32 * i, 32); buf_le64(&body, s->call_amounts[i]); } buf_put(&body, s->raw_tail, s->raw_tail_len); break;
  case XO_TX_DEPLOY: aux = s->raw_tail_len; buf_put(&body, s->raw_tail, s->raw_tail_len); break;
  case XO_TX_MULTISIG: count = s->n_signers; aux = s->threshold;
    xo_transcript_append(&tr, "dom-sep", "multisig-proof", 14); xo_transcript_append_u64(&tr, "threshold", s->threshold);
    for (uint32_t i = 0; i < count; i++) xo_transcript_append(&tr, "signer", s->signers + 32 * i, 32); buf_put(&body, s->signers, (size_t)count * 32); break;
  }
  size_t rp_len = xo_rp_size((int)m); uint8_t *rp = malloc(rp_len); xo_rp_prove(rp, values, openings, (int)m, &tr, rng);
  buf w = {0}; uint8_t hdr[64]; memset(hdr, 0, 64); hdr[0] = s->version; hdr[1] = s->type; hdr[2] = (uint8_t)a; hdr[3] = n_ms > 0 ? (uint8_t)n_ms : 0xFF;
  wr32(hdr + 4, count); wr32(hdr + 8, aux); wr32(hdr + 12, (uint32_t)rp_len); memcpy(hdr + 16, pk, 32); wr64(hdr + 48, s->fee); wr64(hdr + 56, s->nonce);
  buf_put(&w, hdr, 64); buf_put(&w, body.p, body.n); buf_put(&w, rp, rp_len); buf_put(&w, scb.p, scb.n);
  /* unsigned bytes -> multisig signatures over blake3(bytes) -> main signature over bytes incl. multisig section */
  uint8_t zsig[64] = {0}; buf tmp = {0}; buf_put(&tmp, w.p, w.n); tmp.p[3] = 0xFF; buf_put(&tmp, zsig, 64);
  xo_tx view; if (xo_tx_parse(&view, tmp.p, tmp.n)) return 0; uint8_t *bytes; size_t msi; size_t nb = xo_tx_to_bytes(&view, &bytes, &msi); xo_tx_free(&view);
  if (n_ms > 0) { uint8_t hash[32]; xo_blake3(bytes, nb, hash);
    for (int i = 0; i < n_ms; i++) { uint8_t spk[32], sg[64]; xo_pubkey_from_secret(&ms_sk[i], spk, NULL); xo_sign(sg, &ms_sk[i], spk, hash, 32, rng); buf_u8(&w, ms_index[i]); buf_put(&w, sg, 64); } }
  free(bytes); free(tmp.p); tmp = (buf){0}; buf_put(&tmp, w.p, w.n); buf_put(&tmp, zsig, 64);
  if (xo_tx_parse(&view, tmp.p, tmp.n)) return 0; nb = xo_tx_to_bytes(&view, &bytes, &msi); xo_tx_free(&view);
  uint8_t sig[64]; xo_sign(sig, sk, pk, bytes, nb, rng); buf_put(&w, sig, 64);
  free(bytes); free(tmp.p); free(body.p); free(scb.p); free(rp); free(openings); free(values); free(tC); free(tDs); free(tDr); free(Pd); free(tenc);
  *out = w.p; return w.n;
}
