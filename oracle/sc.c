/* oracle/sc.c -- TEST INFRASTRUCTURE.  See sc.h. */
#include "sc.h"
#include <string.h>
#include <stdlib.h>
typedef unsigned __int128 u128;
static const uint64_t SC_L[4] = { 0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0000000000000000ULL, 0x1000000000000000ULL };
static const uint64_t SC_R[4] = { 0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL };
static const uint64_t SC_RR[4] = { 0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL };
static const uint64_t SC_LFACTOR = 0xd2b51da312547e1bULL;
void sc_0(sc *r) { memset(r, 0, sizeof *r); }
void sc_1(sc *r) { memset(r, 0, sizeof *r); r->v[0] = 1; }
void sc_from_u64(sc *r, uint64_t x) { sc_0(r); r->v[0] = x; }
static int geq_l(const uint64_t a[4]) { for (int i = 3; i >= 0; i--) { if (a[i] > SC_L[i]) return 1; if (a[i] < SC_L[i]) return 0; } return 1; }
static void sub_l(uint64_t a[4]) { u128 b = 0; for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - SC_L[i] - (uint64_t)b; a[i] = (uint64_t)t; b = (t >> 64) & 1; } }
/* Montgomery product a*b*R^-1 mod l, inputs with a*b < l*R */
static void montmul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  uint64_t t[6] = {0,0,0,0,0,0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * SC_LFACTOR;
    c = (u128)m * SC_L[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * SC_L[j] + t[j]; t[j-1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  uint64_t o[4] = { t[0], t[1], t[2], t[3] };
  if (t[4] || geq_l(o)) sub_l(o);
  memcpy(r, o, 32);
}
int sc_frombytes_canonical(sc *r, const uint8_t s[32]) { memcpy(r->v, s, 32); return !geq_l(r->v); }
void sc_frombytes_mod_order(sc *r, const uint8_t s[32]) { uint64_t w[4]; memcpy(w, s, 32); montmul(r->v, w, SC_R); }
void sc_frombytes_wide(sc *r, const uint8_t s[64]) {
  uint64_t lo[4], hi[4]; memcpy(lo, s, 32); memcpy(hi, s + 32, 32);
  sc a, b; montmul(a.v, lo, SC_R); montmul(b.v, hi, SC_RR); sc_add(r, &a, &b);
}
void sc_tobytes(uint8_t s[32], const sc *a) { memcpy(s, a->v, 32); }
void sc_add(sc *r, const sc *a, const sc *b) {
  u128 c = 0; uint64_t o[4];
  for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; o[i] = (uint64_t)c; c >>= 64; }
  if (geq_l(o)) sub_l(o);
  memcpy(r->v, o, 32);
}
void sc_neg(sc *r, const sc *a) {
  if (sc_iszero(a)) { sc_0(r); return; }
  u128 b = 0; uint64_t o[4];
  for (int i = 0; i < 4; i++) { u128 t = (u128)SC_L[i] - a->v[i] - (uint64_t)b; o[i] = (uint64_t)t; b = (t >> 64) & 1; }
  memcpy(r->v, o, 32);
}
void sc_sub(sc *r, const sc *a, const sc *b) { sc n; sc_neg(&n, b); sc_add(r, a, &n); }
void sc_mul(sc *r, const sc *a, const sc *b) { uint64_t t[4]; montmul(t, a->v, b->v); montmul(r->v, t, SC_RR); }
void sc_muladd(sc *r, const sc *a, const sc *b, const sc *c) { sc t; sc_mul(&t, a, b); sc_add(r, &t, c); }
void sc_invert(sc *r, const sc *a) { /* a^(l-2), square-and-multiply in Montgomery form */
  uint64_t e[4]; memcpy(e, SC_L, 32); e[0] -= 2;
  uint64_t am[4], acc[4]; montmul(am, a->v, SC_RR); memcpy(acc, SC_R, 32);
  for (int i = 252; i >= 0; i--) { montmul(acc, acc, acc); if ((e[i >> 6] >> (i & 63)) & 1) montmul(acc, acc, am); }
  uint64_t one[4] = {1,0,0,0}; montmul(r->v, acc, one);
}
void sc_batch_invert(sc *xs, int n, sc *allinv) {
  if (n == 0) { if (allinv) sc_1(allinv); return; }
  sc *pre = (sc*)malloc(sizeof(sc) * n); sc acc; sc_1(&acc);
  for (int i = 0; i < n; i++) { pre[i] = acc; sc_mul(&acc, &acc, &xs[i]); }
  sc_invert(&acc, &acc); if (allinv) *allinv = acc;
  for (int i = n - 1; i >= 0; i--) { sc t; sc_mul(&t, &acc, &pre[i]); sc_mul(&acc, &acc, &xs[i]); xs[i] = t; }
  free(pre);
}
int sc_iszero(const sc *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
int sc_eq(const sc *a, const sc *b) { return memcmp(a->v, b->v, 32) == 0; }
