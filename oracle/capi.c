/* oracle/capi.c -- TEST INFRASTRUCTURE.  Bulk helpers and synthetic-workload generators exported for ctypes
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs ONLY). */
#include "tx.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
void xo_init(int m_max) { xo_G(); xo_H(); xo_bp_ensure(m_max); }
/* --- primitives over byte arrays */
void xo_decode_batch(const uint8_t *enc, size_t n, uint8_t *ok, uint8_t *xy_out /* n*64 canonical x||y or NULL */) {
  for (size_t i = 0; i < n; i++) { ge p; ok[i] = (uint8_t)ristretto_decode(&p, enc + 32 * i); if (xy_out) { if (ok[i]) { fe_tobytes(xy_out + 64 * i, &p.X); fe_tobytes(xy_out + 64 * i + 32, &p.Y); } else memset(xy_out + 64 * i, 0, 64); } }
}
int xo_point_add(const uint8_t a[32], const uint8_t b[32], int sub, uint8_t out[32]) { ge p, q, r; if (!ristretto_decode(&p, a) || !ristretto_decode(&q, b)) return 0; if (sub) ge_sub(&r, &p, &q); else ge_add(&r, &p, &q); ristretto_encode(out, &r); return 1; }
int xo_scalarmult(const uint8_t s[32], const uint8_t p[32], uint8_t out[32]) { ge P, r; sc k; if (!ristretto_decode(&P, p)) return 0; sc_frombytes_mod_order(&k, s); ge_scalarmult(&r, &k, &P); ristretto_encode(out, &r); return 1; }
void xo_from_uniform(const uint8_t u[64], uint8_t out[32]) { ge p; ristretto_from_uniform(&p, u); ristretto_encode(out, &p); }
void xo_sc_reduce_wide(const uint8_t in[64], uint8_t out[32]) { sc s; sc_frombytes_wide(&s, in); sc_tobytes(out, &s); }
void xo_sc_op(int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { sc x, y, r; sc_frombytes_mod_order(&x, a); sc_frombytes_mod_order(&y, b);
  switch (op) { case 0: sc_add(&r, &x, &y); break; case 1: sc_sub(&r, &x, &y); break; case 2: sc_mul(&r, &x, &y); break; case 3: sc_invert(&r, &x); break; default: sc_neg(&r, &x); } sc_tobytes(out, &r); }
/* mode 0: dalek dispatch, 1: straus, 2: pippenger, 3: naive.  returns 0 if a point fails to decode */
int xo_msm(const uint8_t *scalars, const uint8_t *points, size_t n, int mode, uint8_t out[32]) {
  sc *s = malloc(sizeof(sc) * (n + 1)); ge *p = malloc(sizeof(ge) * (n + 1)); int ok = 1;
  for (size_t i = 0; i < n && ok; i++) { sc_frombytes_mod_order(&s[i], scalars + 32 * i); ok = ristretto_decode(&p[i], points + 32 * i); }
  if (ok) { ge r; if (mode == 1) ge_msm_straus(&r, s, p, n); else if (mode == 2) ge_msm_pippenger(&r, s, p, n); else if (mode == 3) ge_msm_naive(&r, s, p, n); else ge_msm_vartime(&r, s, p, n); ristretto_encode(out, &r); }
  free(s); free(p); return ok;
}
/* ciphertext add/sub, config 4 (src/elgamal.rs:322-342, src/tx/verify.rs:561-609): out = bal +/- delta */
void xo_ct_update(const uint8_t *bal, const uint8_t *delta, const uint8_t *sub, size_t n, uint8_t *out, uint8_t *ok) {
  for (size_t i = 0; i < n; i++) { ok[i] = 1; for (int h = 0; h < 2; h++) { if (!xo_point_add(bal + 64 * i + 32 * h, delta + 64 * i + 32 * h, sub[i], out + 64 * i + 32 * h)) { ok[i] = 0; memset(out + 64 * i, 0, 64); break; } } }
}
/* --- C2 inputs: SHAKE256("xhe-msm" || seed || i) -> 64 B scalar (wide reduce) || 64 B uniform (one-way map) */
typedef struct { uint64_t seed; size_t lo, hi; uint8_t *scalars, *points; int nbases; } msm_job;
static void *msm_gen_worker(void *arg) { msm_job *j = arg;
  for (size_t i = j->lo; i < j->hi; i++) { uint8_t in[23], o[128]; memcpy(in, "xhe-msm", 7); memcpy(in + 7, &j->seed, 8); uint64_t ii = i; memcpy(in + 15, &ii, 8); xo_shake256(in, 23, o, 128);
    xo_sc_reduce_wide(o, j->scalars + 32 * i); if (j->points) xo_from_uniform(o + 64, j->points + 32 * i); } return NULL; }
void xo_gen_msm_inputs(uint64_t seed, size_t n, uint8_t *scalars, uint8_t *points, int threads) {
  if (threads < 1) threads = 1; pthread_t th[64]; msm_job jobs[64]; if (threads > 64) threads = 64;
  for (int t = 0; t < threads; t++) { jobs[t] = (msm_job){ seed, n * t / threads, n * (t + 1) / threads, scalars, points, 0 }; pthread_create(&th[t], NULL, msm_gen_worker, &jobs[t]); }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}
/* expected value of sum s_i * B_{idx_i} with B_j = b_j G known multiples: (sum s_i b_idx_i) G -- exact at any n */
void xo_msm_expected_known_bases(const uint8_t *scalars, const uint32_t *idx, size_t n, const uint8_t *base_scalars, uint8_t out[32]) {
  sc acc; sc_0(&acc); for (size_t i = 0; i < n; i++) { sc s, b, t; sc_frombytes_mod_order(&s, scalars + 32 * i); sc_frombytes_mod_order(&b, base_scalars + 32 * idx[i]); sc_mul(&t, &s, &b); sc_add(&acc, &acc, &t); }
  ge r; ge_scalarmult_base(&r, &acc); ristretto_encode(out, &r);
}
/* --- C1/C3 batches: T transfer TXs, a assets, k transfers each */
typedef struct { uint8_t *blobs; size_t *offsets; size_t n; xo_ledger *ledger; } xo_batch;
typedef struct { uint64_t seed; size_t lo, hi, T; int a, k; uint8_t **blob; size_t *len; uint8_t *sender_pk, *sender_ct; uint8_t *recv_pk, *recv_ct; } mint_job;
static void asset_id(uint8_t out[32], int j) { memset(out, j == 0 ? 0 : 0x50 + j, 32); }
static void derive_key(uint64_t seed, const char *tag, uint64_t i, sc *sk) { uint8_t in[32], o[64]; memset(in, 0, 32); memcpy(in, "xhe-key", 7); in[7] = (uint8_t)tag[0]; memcpy(in + 8, &seed, 8); memcpy(in + 16, &i, 8); xo_shake256(in, 24, o, 64); sc_frombytes_wide(sk, o); if (sc_iszero(sk)) sc_1(sk); }
#define XO_BAL (1ULL << 40)
static void *mint_worker(void *arg) { mint_job *j = arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    uint8_t sd[24]; memcpy(sd, "mint-tx", 8); memcpy(sd + 8, &j->seed, 8); uint64_t ii = i; memcpy(sd + 16, &ii, 8); xo_rng rng; xo_rng_init(&rng, sd, 24);
    sc sk, rsk; derive_key(j->seed, "s", i, &sk); ge P; uint8_t pk[32]; xo_pubkey_from_secret(&sk, pk, &P); memcpy(j->sender_pk + 32 * i, pk, 32);
    derive_key(j->seed, "r", i, &rsk); ge RP; uint8_t rpk[32]; xo_pubkey_from_secret(&rsk, rpk, &RP); memcpy(j->recv_pk + 32 * i, rpk, 32);
    xo_ledger *tmp = xo_ledger_new(); uint8_t assets[8 * 32]; uint64_t bals[8];
    for (int q = 0; q < j->a; q++) { asset_id(assets + 32 * q, q); bals[q] = XO_BAL; sc op; xo_rng_scalar(&rng, &op); uint8_t ct[64]; xo_encrypt(ct, &P, XO_BAL, &op); memcpy(j->sender_ct + (i * j->a + q) * 64, ct, 64); xo_ledger_set_balance(tmp, pk, assets + 32 * q, ct);
      sc op2; xo_rng_scalar(&rng, &op2); xo_encrypt(ct, &RP, 0, &op2); memcpy(j->recv_ct + (i * j->a + q) * 64, ct, 64); }
    xo_transfer_spec *ts = calloc(j->k, sizeof *ts);
    for (int q = 0; q < j->k; q++) { asset_id(ts[q].asset, q % j->a); size_t r = (i * 7919 + 13 * q + 1) % j->T; (void)r; memcpy(ts[q].dest, rpk, 32); uint32_t amt; xo_rng_bytes(&rng, &amt, 4); ts[q].amount = amt; }
    xo_tx_spec sp; memset(&sp, 0, sizeof sp); sp.version = 1; sp.type = XO_TX_TRANSFERS; sp.fee = 3; sp.nonce = 0; sp.transfers = ts; sp.n_transfers = j->k; sp.assets = assets; sp.balances = bals; sp.n_assets = j->a;
    j->len[i] = xo_tx_build(&j->blob[i], &sp, &sk, tmp, &rng, NULL, NULL, 0); free(ts); xo_ledger_free(tmp);
  } return NULL; }
/* distinct senders (sender i, receiver i are fresh accounts; every TX touches its own balance keys) */
xo_batch *xo_mint_transfers(uint64_t seed, size_t T, int a, int k, int threads) {
  if (a < 1 || a > 8 || k < 0 || threads < 1) return NULL; if (threads > 64) threads = 64; size_t m = 1; while (m < (size_t)(a + k)) m <<= 1; xo_init((int)m);
  uint8_t **blob = calloc(T, sizeof *blob); size_t *len = calloc(T, sizeof *len); uint8_t *spk = malloc(32 * T + 1), *sct = malloc(64 * T * a + 1), *rpk = malloc(32 * T + 1), *rct = malloc(64 * T * a + 1);
  pthread_t th[64]; mint_job jobs[64];
  for (int t = 0; t < threads; t++) { jobs[t] = (mint_job){ seed, T * t / threads, T * (t + 1) / threads, T, a, k, blob, len, spk, sct, rpk, rct }; pthread_create(&th[t], NULL, mint_worker, &jobs[t]); }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  xo_batch *b = calloc(1, sizeof *b); b->n = T; b->offsets = calloc(T + 1, sizeof(size_t)); b->ledger = xo_ledger_new();
  for (size_t i = 0; i < T; i++) { if (!len[i]) return NULL; b->offsets[i + 1] = b->offsets[i] + len[i]; }
  b->blobs = malloc(b->offsets[T] + 1);
  for (size_t i = 0; i < T; i++) { memcpy(b->blobs + b->offsets[i], blob[i], len[i]); free(blob[i]); xo_ledger_set_nonce(b->ledger, spk + 32 * i, 0); xo_ledger_set_nonce(b->ledger, rpk + 32 * i, 0);
    for (int q = 0; q < a; q++) { uint8_t as[32]; asset_id(as, q); xo_ledger_set_balance(b->ledger, spk + 32 * i, as, sct + (i * a + q) * 64); xo_ledger_set_balance(b->ledger, rpk + 32 * i, as, rct + (i * a + q) * 64); } }
  free(blob); free(len); free(spk); free(sct); free(rpk); free(rct); return b;
}
/* benches/tx.rs:129-186 shape: one sender -> one receiver, amount 1, fee 3, nonce 0, balance 100000, chained via apply_without_verify */
xo_batch *xo_mint_chain(uint64_t seed, size_t T, int k) {
  size_t m = 1; while (m < (size_t)(1 + k)) m <<= 1; xo_init((int)m);
  uint8_t sd[16]; memcpy(sd, "mintchn", 8); memcpy(sd + 8, &seed, 8); xo_rng rng; xo_rng_init(&rng, sd, 16);
  sc sk, rsk; uint8_t pk[32], rpk[32]; ge P, RP; derive_key(seed, "S", 0, &sk); derive_key(seed, "R", 0, &rsk); xo_pubkey_from_secret(&sk, pk, &P); xo_pubkey_from_secret(&rsk, rpk, &RP);
  xo_batch *b = calloc(1, sizeof *b); b->n = T; b->offsets = calloc(T + 1, sizeof(size_t)); b->ledger = xo_ledger_new(); uint8_t native[32] = {0}, ct[64]; sc op;
  xo_rng_scalar(&rng, &op); xo_encrypt(ct, &P, 100000, &op); xo_ledger_set_balance(b->ledger, pk, native, ct); xo_rng_scalar(&rng, &op); xo_encrypt(ct, &RP, 0, &op); xo_ledger_set_balance(b->ledger, rpk, native, ct);
  xo_ledger_set_nonce(b->ledger, pk, 0); xo_ledger_set_nonce(b->ledger, rpk, 0);
  xo_ledger *prover = xo_ledger_clone(b->ledger); uint64_t bal = 100000; size_t cap = 0;
  for (size_t i = 0; i < T; i++) {
    xo_transfer_spec *ts = calloc(k ? k : 1, sizeof *ts); for (int q = 0; q < k; q++) { memcpy(ts[q].dest, rpk, 32); ts[q].amount = 1; }
    xo_tx_spec sp; memset(&sp, 0, sizeof sp); sp.version = 1; sp.type = XO_TX_TRANSFERS; sp.fee = 3; sp.transfers = ts; sp.n_transfers = k; sp.assets = native; sp.balances = &bal; sp.n_assets = 1;
    uint8_t *blob; size_t len = xo_tx_build(&blob, &sp, &sk, prover, &rng, NULL, NULL, 0); free(ts); if (!len) return NULL;
    xo_apply_without_verify(blob, len, prover); bal -= 3 + (uint64_t)k;
    if (b->offsets[i] + len > cap) { cap = (b->offsets[i] + len) * 2; b->blobs = realloc(b->blobs, cap); }
    memcpy(b->blobs + b->offsets[i], blob, len); b->offsets[i + 1] = b->offsets[i] + len; free(blob);
  }
  xo_ledger_free(prover); return b;
}
/* The same one-sender chain (benches/tx.rs:153-186) minted in PARALLEL, for 10k-transaction chains: every transaction draws its
 * randomness from its own stream (seed, i), so its transfer ciphertexts do not depend on the prover's state.  Pass 1 builds all
 * transactions against the initial ledger (their ciphertexts are final, their eq proofs are not); the running sender balance after
 * each of them follows by apply_without_verify (cheap, sequential); pass 2 rebuilds every transaction with the same randomness
 * against the balance it really meets. */
typedef struct { uint64_t seed; size_t lo, hi; int k; const sc *sk; const uint8_t *pk, *rpk; const uint8_t *pre_ct /* T x 64 or NULL */; const uint8_t *init_ct; uint8_t **blob; size_t *len; } chain_job;
static void *chain_worker(void *arg) { chain_job *j = arg; uint8_t native[32] = {0};
  for (size_t i = j->lo; i < j->hi; i++) {
    uint8_t sd[24]; memcpy(sd, "chain-tx", 8); memcpy(sd + 8, &j->seed, 8); uint64_t ii = i; memcpy(sd + 16, &ii, 8); xo_rng rng; xo_rng_init(&rng, sd, 24);
    xo_ledger *tmp = xo_ledger_new(); xo_ledger_set_balance(tmp, j->pk, native, j->pre_ct ? j->pre_ct + 64 * i : j->init_ct); xo_ledger_set_nonce(tmp, j->pk, 0);
    xo_transfer_spec *ts = calloc(j->k ? j->k : 1, sizeof *ts); for (int q = 0; q < j->k; q++) { memcpy(ts[q].dest, j->rpk, 32); ts[q].amount = 1; }
    uint64_t bal = 100000 - (uint64_t)i * (3 + (uint64_t)j->k);
    xo_tx_spec sp; memset(&sp, 0, sizeof sp); sp.version = 1; sp.type = XO_TX_TRANSFERS; sp.fee = 3; sp.transfers = ts; sp.n_transfers = j->k; sp.assets = native; sp.balances = &bal; sp.n_assets = 1;
    if (j->blob[i]) { free(j->blob[i]); j->blob[i] = NULL; }
    j->len[i] = xo_tx_build(&j->blob[i], &sp, j->sk, tmp, &rng, NULL, NULL, 0); free(ts); xo_ledger_free(tmp);
  } return NULL; }
xo_batch *xo_mint_chain_par(uint64_t seed, size_t T, int k, int threads) {
  if (threads < 1) threads = 1; if (threads > 64) threads = 64; if ((uint64_t)T * (3 + (uint64_t)k) > 100000) return NULL;
  size_t m = 1; while (m < (size_t)(1 + k)) m <<= 1; xo_init((int)m);
  uint8_t sd[16]; memcpy(sd, "mintchp", 8); memcpy(sd + 8, &seed, 8); xo_rng rng; xo_rng_init(&rng, sd, 16);
  sc sk, rsk; uint8_t pk[32], rpk[32]; ge P, RP; derive_key(seed, "S", 0, &sk); derive_key(seed, "R", 0, &rsk); xo_pubkey_from_secret(&sk, pk, &P); xo_pubkey_from_secret(&rsk, rpk, &RP);
  xo_batch *b = calloc(1, sizeof *b); b->n = T; b->offsets = calloc(T + 1, sizeof(size_t)); b->ledger = xo_ledger_new(); uint8_t native[32] = {0}, ct0[64], ct[64]; sc op;
  xo_rng_scalar(&rng, &op); xo_encrypt(ct0, &P, 100000, &op); xo_ledger_set_balance(b->ledger, pk, native, ct0); xo_rng_scalar(&rng, &op); xo_encrypt(ct, &RP, 0, &op); xo_ledger_set_balance(b->ledger, rpk, native, ct);
  xo_ledger_set_nonce(b->ledger, pk, 0); xo_ledger_set_nonce(b->ledger, rpk, 0);
  uint8_t **blob = calloc(T, sizeof *blob); size_t *len = calloc(T, sizeof *len); uint8_t *pre = malloc(64 * T + 64);
  pthread_t th[64]; chain_job jobs[64];
  for (int pass = 0; pass < 2; pass++) {
    for (int t = 0; t < threads; t++) { jobs[t] = (chain_job){ seed, T * t / threads, T * (t + 1) / threads, k, &sk, pk, rpk, pass ? pre : NULL, ct0, blob, len }; pthread_create(&th[t], NULL, chain_worker, &jobs[t]); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    if (pass == 0) {   /* the sender balance each transaction meets */
      xo_ledger *run = xo_ledger_clone(b->ledger);
      for (size_t i = 0; i < T; i++) { if (!len[i]) return NULL; xo_ledger_get_balance(run, pk, native, pre + 64 * i); xo_apply_without_verify(blob[i], len[i], run); }
      xo_ledger_free(run);
    }
  }
  for (size_t i = 0; i < T; i++) { if (!len[i]) return NULL; b->offsets[i + 1] = b->offsets[i] + len[i]; }
  b->blobs = malloc(b->offsets[T] + 1);
  for (size_t i = 0; i < T; i++) { memcpy(b->blobs + b->offsets[i], blob[i], len[i]); free(blob[i]); }
  free(blob); free(len); free(pre); return b;
}
/* Config 5: a mixed batch -- 60 % Transfers (k in 1..4, a in 1..2), 15 % Burn, 15 % CallContract (one asset, one
 * parameter), 5 % MultiSig set-ups, 5 % transfers from threshold-2 multisig accounts -- every transaction from its own fresh
 * sender, so the batch mints in parallel.  party_capacity 8 covers it (a + k <= 6). */
typedef struct { uint64_t seed; size_t lo, hi; uint8_t **blob; size_t *len; uint8_t *spk, *sct, *rpk, *rct, *ms /* T x (1 + 2 x 32): flag, two signer keys */; } mixed_job;
static int mixed_kind(size_t i) { unsigned r = (unsigned)((i * 2654435761u) >> 7) % 100; return r < 60 ? 0 : r < 75 ? 1 : r < 90 ? 2 : r < 95 ? 4 : 5; }   /* 5 = transfer from a multisig account */
static void *mixed_worker(void *arg) { mixed_job *j = arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    uint8_t sd[24]; memcpy(sd, "mixed-tx", 8); memcpy(sd + 8, &j->seed, 8); uint64_t ii = i; memcpy(sd + 16, &ii, 8); xo_rng rng; xo_rng_init(&rng, sd, 24);
    sc sk, rsk, c1, c2; derive_key(j->seed, "s", i, &sk); ge P; uint8_t pk[32]; xo_pubkey_from_secret(&sk, pk, &P); memcpy(j->spk + 32 * i, pk, 32);
    derive_key(j->seed, "r", i, &rsk); ge RP; uint8_t rpk[32]; xo_pubkey_from_secret(&rsk, rpk, &RP); memcpy(j->rpk + 32 * i, rpk, 32);
    const int kind = mixed_kind(i); uint32_t rnd; xo_rng_bytes(&rng, &rnd, 4);
    const int a = (kind == 0 || kind == 5) ? 1 + (int)(rnd & 1) : 1, k = (kind == 0 || kind == 5) ? 1 + (int)((rnd >> 1) & 3) : 0;
    xo_ledger *tmp = xo_ledger_new(); uint8_t assets[2 * 32]; uint64_t bals[2];
    for (int q = 0; q < 2; q++) { asset_id(assets + 32 * q, q); bals[q] = XO_BAL; sc op; xo_rng_scalar(&rng, &op); uint8_t ct[64]; xo_encrypt(ct, &P, XO_BAL, &op); memcpy(j->sct + (i * 2 + q) * 64, ct, 64); xo_ledger_set_balance(tmp, pk, assets + 32 * q, ct);
      sc op2; xo_rng_scalar(&rng, &op2); xo_encrypt(ct, &RP, 0, &op2); memcpy(j->rct + (i * 2 + q) * 64, ct, 64); }
    xo_tx_spec sp; memset(&sp, 0, sizeof sp); sp.version = 1; sp.fee = 3; sp.assets = assets; sp.balances = bals; sp.n_assets = (uint32_t)a;
    xo_transfer_spec ts[4]; memset(ts, 0, sizeof ts); uint64_t amt = 0; uint8_t tail[32]; uint8_t signers[64], msidx[2] = {0, 1}; sc mssk[2];
    j->ms[65 * i] = 0;
    if (kind == 0 || kind == 5) {
      sp.type = XO_TX_TRANSFERS; sp.transfers = ts; sp.n_transfers = (uint32_t)k;
      for (int q = 0; q < k; q++) { asset_id(ts[q].asset, q % a); memcpy(ts[q].dest, rpk, 32); uint32_t v; xo_rng_bytes(&rng, &v, 4); ts[q].amount = v; }
    } else if (kind == 1) { sp.type = XO_TX_BURN; memcpy(sp.burn_asset, assets, 32); sp.burn_amount = 1000 + (rnd & 0xffff);
    } else if (kind == 2) { sp.type = XO_TX_CALL; memset(sp.contract, 9, 32); amt = 40 + (rnd & 0xff); sp.call_assets = assets; sp.call_amounts = &amt; sp.n_call_assets = 1;
      uint32_t kl = 6, vl = 4; memcpy(tail, &kl, 4); memcpy(tail + 4, "method", 6); memcpy(tail + 10, &vl, 4); memcpy(tail + 14, "swap", 4); sp.raw_tail = tail; sp.raw_tail_len = 18; sp.n_params = 1;
    } else if (kind == 4) { sp.type = XO_TX_MULTISIG; derive_key(j->seed, "c", 2 * i, &c1); derive_key(j->seed, "c", 2 * i + 1, &c2); xo_pubkey_from_secret(&c1, signers, NULL); xo_pubkey_from_secret(&c2, signers + 32, NULL);
      sp.signers = signers; sp.n_signers = 2; sp.threshold = 1; }
    if (kind == 5) {   /* the account already has a threshold-2 multisig setting in the ledger: both co-signers sign */
      derive_key(j->seed, "c", 2 * i, &c1); derive_key(j->seed, "c", 2 * i + 1, &c2); xo_pubkey_from_secret(&c1, signers, NULL); xo_pubkey_from_secret(&c2, signers + 32, NULL);
      j->ms[65 * i] = 1; memcpy(j->ms + 65 * i + 1, signers, 64); mssk[0] = c1; mssk[1] = c2;
      xo_ledger_set_multisig(tmp, pk, signers, 2, 2);
      j->len[i] = xo_tx_build(&j->blob[i], &sp, &sk, tmp, &rng, msidx, mssk, 2);
    } else j->len[i] = xo_tx_build(&j->blob[i], &sp, &sk, tmp, &rng, NULL, NULL, 0);
    xo_ledger_free(tmp);
  } return NULL; }
xo_batch *xo_mint_mixed(uint64_t seed, size_t T, int threads) {
  if (threads < 1) threads = 1; if (threads > 64) threads = 64; xo_init(8);
  uint8_t **blob = calloc(T, sizeof *blob); size_t *len = calloc(T, sizeof *len); uint8_t *spk = malloc(32 * T + 1), *sct = malloc(128 * T + 1), *rpk = malloc(32 * T + 1), *rct = malloc(128 * T + 1), *ms = calloc(65 * T + 1, 1);
  pthread_t th[64]; mixed_job jobs[64];
  for (int t = 0; t < threads; t++) { jobs[t] = (mixed_job){ seed, T * t / threads, T * (t + 1) / threads, blob, len, spk, sct, rpk, rct, ms }; pthread_create(&th[t], NULL, mixed_worker, &jobs[t]); }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  xo_batch *b = calloc(1, sizeof *b); b->n = T; b->offsets = calloc(T + 1, sizeof(size_t)); b->ledger = xo_ledger_new();
  for (size_t i = 0; i < T; i++) { if (!len[i]) return NULL; b->offsets[i + 1] = b->offsets[i] + len[i]; }
  b->blobs = malloc(b->offsets[T] + 1);
  for (size_t i = 0; i < T; i++) { memcpy(b->blobs + b->offsets[i], blob[i], len[i]); free(blob[i]); xo_ledger_set_nonce(b->ledger, spk + 32 * i, 0); xo_ledger_set_nonce(b->ledger, rpk + 32 * i, 0);
    for (int q = 0; q < 2; q++) { uint8_t as[32]; asset_id(as, q); xo_ledger_set_balance(b->ledger, spk + 32 * i, as, sct + (i * 2 + q) * 64); xo_ledger_set_balance(b->ledger, rpk + 32 * i, as, rct + (i * 2 + q) * 64); }
    if (ms[65 * i]) xo_ledger_set_multisig(b->ledger, spk + 32 * i, ms + 65 * i + 1, 2, 2); }
  free(blob); free(len); free(spk); free(sct); free(rpk); free(rct); free(ms); return b;
}
/* the multisig settings of a batch's initial ledger as records pk[32] n[1] threshold[1] signers[n x 32] (for importing into another ledger) */
size_t xo_ledger_dump_multisig(const xo_ledger *l, uint8_t *out, size_t cap);
/* secret key of sender (which = 's') / receiver ('r') i of a batch minted by xo_mint_transfers / xo_mint_mixed with `seed`: lets a
 * test re-sign a deliberately broken transaction of a big minted batch (xo_resign) */
void xo_mint_secret(uint64_t seed, int which, uint64_t i, uint8_t sk_out[32]) { sc sk; char tag[2] = {(char)which, 0}; derive_key(seed, tag, i, &sk); sc_tobytes(sk_out, &sk); }
void xo_batch_free(xo_batch *b) { if (!b) return; free(b->blobs); free(b->offsets); xo_ledger_free(b->ledger); free(b); }
/* verify a whole batch against a clone of its ledger; returns verdict */
int xo_batch_verify(const xo_batch *b, uint64_t rng_seed, long *fail_index, xo_ledger **final_state) {
  const uint8_t **blobs = malloc(sizeof(*blobs) * (b->n + 1)); size_t *lens = malloc(sizeof(size_t) * (b->n + 1));
  for (size_t i = 0; i < b->n; i++) { blobs[i] = b->blobs + b->offsets[i]; lens[i] = b->offsets[i + 1] - b->offsets[i]; }
  xo_ledger *st = xo_ledger_clone(b->ledger); xo_rng rng; xo_rng_init(&rng, &rng_seed, 8);
  int rc = xo_verify_batch(blobs, lens, b->n, st, &rng, fail_index); if (final_state) *final_state = st; else xo_ledger_free(st); free(blobs); free(lens); return rc;
}
/* CPU baseline: `threads` independent verify_batch calls on cloned ledgers (benches/tx.rs:314-325); returns seconds */
typedef struct { const xo_batch *b; int rc; } vjob;
static void *verify_worker(void *arg) { vjob *j = arg; long fi; j->rc = xo_batch_verify(j->b, 1234, &fi, NULL); return NULL; }
double xo_batch_verify_timed(const xo_batch *b, int threads, int *rc_out) {
  pthread_t th[256]; vjob jobs[256]; if (threads > 256) threads = 256; struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) { jobs[t].b = b; pthread_create(&th[t], NULL, verify_worker, &jobs[t]); }
  int rc = 0; for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
  clock_gettime(CLOCK_MONOTONIC, &t1); if (rc_out) *rc_out = rc; return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
double xo_msm_timed(const uint8_t *scalars, const uint8_t *points, size_t n, uint8_t out[32]) {
  sc *s = malloc(sizeof(sc) * (n + 1)); ge *p = malloc(sizeof(ge) * (n + 1));
  for (size_t i = 0; i < n; i++) { sc_frombytes_mod_order(&s[i], scalars + 32 * i); ristretto_decode(&p[i], points + 32 * i); }
  struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0); ge r; ge_msm_vartime(&r, s, p, n); clock_gettime(CLOCK_MONOTONIC, &t1);
  ristretto_encode(out, &r); free(s); free(p); return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
/* re-sign a (possibly mutated) transaction in place: recompute to_bytes and the trailing 64-byte signature */
int xo_resign(uint8_t *blob, size_t len, const sc *sk, xo_rng *rng) {
  xo_tx tx; if (xo_tx_parse(&tx, blob, len)) return 0;
  uint8_t *bytes; size_t msi; size_t nb = xo_tx_to_bytes(&tx, &bytes, &msi); xo_tx_free(&tx);
  uint8_t pk[32]; xo_pubkey_from_secret(sk, pk, NULL);
  xo_sign(blob + len - 64, sk, pk, bytes, nb, rng); free(bytes); return 1;
}
/* first `count` transactions of a batch with a clone of its ledger (bounded CPU-baseline samples) */
xo_batch *xo_batch_slice(const xo_batch *b, size_t count) {
  if (count > b->n) count = b->n;
  xo_batch *s = calloc(1, sizeof *s); s->n = count; s->offsets = calloc(count + 1, sizeof(size_t)); memcpy(s->offsets, b->offsets, (count + 1) * sizeof(size_t));
  s->blobs = malloc(s->offsets[count] + 1); memcpy(s->blobs, b->blobs, s->offsets[count]); s->ledger = xo_ledger_clone(b->ledger); return s;
}
