/* oracle/fe.c -- TEST INFRASTRUCTURE.  See fe.h. */
#include "fe.h"
#include "consts.h"
#include <string.h>
typedef unsigned __int128 u128;
#define M51 0x7ffffffffffffULL
void fe_0(fe *h) { memset(h, 0, sizeof *h); }
void fe_1(fe *h) { memset(h, 0, sizeof *h); h->v[0] = 1; }
static void fe_carry(fe *h) { /* weak reduction: limbs < 2^51 + small */
  uint64_t c;
  c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
  c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
  c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
  c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
  c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += c * 19;
}
void fe_add(fe *h, const fe *f, const fe *g) { for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i]; fe_carry(h); }
void fe_sub(fe *h, const fe *f, const fe *g) {
  /* add 16p before subtracting so limbs stay positive (inputs are weakly reduced: < 2^52) */
  h->v[0] = f->v[0] + 0x7ffffffffffed0ULL - g->v[0];
  for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 0x7ffffffffffff0ULL - g->v[i];
  fe_carry(h);
}
void fe_neg(fe *h, const fe *f) { fe z; fe_0(&z); fe_sub(h, &z, f); }
void fe_mul(fe *h, const fe *f, const fe *g) {
  const uint64_t *a = f->v, *b = g->v;
  uint64_t b1 = b[1] * 19, b2 = b[2] * 19, b3 = b[3] * 19, b4 = b[4] * 19;
  u128 c0 = (u128)a[0]*b[0] + (u128)a[4]*b1 + (u128)a[3]*b2 + (u128)a[2]*b3 + (u128)a[1]*b4;
  u128 c1 = (u128)a[1]*b[0] + (u128)a[0]*b[1] + (u128)a[4]*b2 + (u128)a[3]*b3 + (u128)a[2]*b4;
  u128 c2 = (u128)a[2]*b[0] + (u128)a[1]*b[1] + (u128)a[0]*b[2] + (u128)a[4]*b3 + (u128)a[3]*b4;
  u128 c3 = (u128)a[3]*b[0] + (u128)a[2]*b[1] + (u128)a[1]*b[2] + (u128)a[0]*b[3] + (u128)a[4]*b4;
  u128 c4 = (u128)a[4]*b[0] + (u128)a[3]*b[1] + (u128)a[2]*b[2] + (u128)a[1]*b[3] + (u128)a[0]*b[4];
  c1 += (uint64_t)(c0 >> 51); uint64_t r0 = (uint64_t)c0 & M51;
  c2 += (uint64_t)(c1 >> 51); uint64_t r1 = (uint64_t)c1 & M51;
  c3 += (uint64_t)(c2 >> 51); uint64_t r2 = (uint64_t)c2 & M51;
  c4 += (uint64_t)(c3 >> 51); uint64_t r3 = (uint64_t)c3 & M51;
  uint64_t carry = (uint64_t)(c4 >> 51); uint64_t r4 = (uint64_t)c4 & M51;
  r0 += carry * 19; r1 += r0 >> 51; r0 &= M51;
  h->v[0] = r0; h->v[1] = r1; h->v[2] = r2; h->v[3] = r3; h->v[4] = r4;
}
void fe_sq(fe *h, const fe *f) {
  const uint64_t *a = f->v;
  uint64_t a3_19 = a[3] * 19, a4_19 = a[4] * 19;
  u128 c0 = (u128)a[0]*a[0] + 2*((u128)a[1]*a4_19 + (u128)a[2]*a3_19);
  u128 c1 = (u128)a[3]*a3_19 + 2*((u128)a[0]*a[1] + (u128)a[2]*a4_19);
  u128 c2 = (u128)a[1]*a[1] + 2*((u128)a[0]*a[2] + (u128)a[4]*a3_19);
  u128 c3 = (u128)a[4]*a4_19 + 2*((u128)a[0]*a[3] + (u128)a[1]*a[2]);
  u128 c4 = (u128)a[2]*a[2] + 2*((u128)a[0]*a[4] + (u128)a[1]*a[3]);
  c1 += (uint64_t)(c0 >> 51); uint64_t r0 = (uint64_t)c0 & M51;
  c2 += (uint64_t)(c1 >> 51); uint64_t r1 = (uint64_t)c1 & M51;
  c3 += (uint64_t)(c2 >> 51); uint64_t r2 = (uint64_t)c2 & M51;
  c4 += (uint64_t)(c3 >> 51); uint64_t r3 = (uint64_t)c3 & M51;
  uint64_t carry = (uint64_t)(c4 >> 51); uint64_t r4 = (uint64_t)c4 & M51;
  r0 += carry * 19; r1 += r0 >> 51; r0 &= M51;
  h->v[0] = r0; h->v[1] = r1; h->v[2] = r2; h->v[3] = r3; h->v[4] = r4;
}
void fe_sqn(fe *h, const fe *f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }
/* t = f^(2^250-1) and f^11, the common prefix of the inversion / (p-5)/8 addition chains */
static void fe_pow_2_250_1(fe *t250, fe *f11, const fe *f) {
  fe t0, t1, t2, t3;
  fe_sq(&t0, f); fe_sqn(&t1, &t0, 2); fe_mul(&t1, f, &t1);      /* t0 = f^2, t1 = f^9 */
  fe_mul(&t0, &t0, &t1);                                         /* f^11 */
  *f11 = t0;
  fe_sq(&t2, &t0); fe_mul(&t1, &t1, &t2);                        /* f^31 = 2^5-1 */
  fe_sqn(&t2, &t1, 5); fe_mul(&t1, &t2, &t1);                    /* 2^10-1 */
  fe_sqn(&t2, &t1, 10); fe_mul(&t2, &t2, &t1);                   /* 2^20-1 */
  fe_sqn(&t3, &t2, 20); fe_mul(&t2, &t3, &t2);                   /* 2^40-1 */
  fe_sqn(&t2, &t2, 10); fe_mul(&t1, &t2, &t1);                   /* 2^50-1 */
  fe_sqn(&t2, &t1, 50); fe_mul(&t2, &t2, &t1);                   /* 2^100-1 */
  fe_sqn(&t3, &t2, 100); fe_mul(&t2, &t3, &t2);                  /* 2^200-1 */
  fe_sqn(&t2, &t2, 50); fe_mul(t250, &t2, &t1);                  /* 2^250-1 */
}
void fe_invert(fe *h, const fe *f) { fe t, f11; fe_pow_2_250_1(&t, &f11, f); fe_sqn(&t, &t, 5); fe_mul(h, &t, &f11); } /* 2^255-21 */
void fe_pow22523(fe *h, const fe *f) { fe t, f11; fe_pow_2_250_1(&t, &f11, f); fe_sqn(&t, &t, 2); fe_mul(h, &t, f); }   /* 2^252-3 */
void fe_frombytes(fe *h, const uint8_t s[32]) {
  uint64_t w[4]; memcpy(w, s, 32);
  h->v[0] = w[0] & M51;
  h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
  h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
  h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
  h->v[4] = (w[3] >> 12) & M51;
}
void fe_tobytes(uint8_t s[32], const fe *f) {
  fe t = *f; fe_carry(&t); fe_carry(&t);
  /* now t < 2^255 + small; compute q = floor((t + 19) / 2^255) */
  uint64_t q = (t.v[0] + 19) >> 51; q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
  t.v[0] += 19 * q;
  uint64_t c;
  c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
  c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
  c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
  c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
  t.v[4] &= M51;
  uint64_t w[4];
  w[0] = t.v[0] | (t.v[1] << 51);
  w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
  w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
  w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
  memcpy(s, w, 32);
}
int fe_isnegative(const fe *f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
int fe_iszero(const fe *f) { uint8_t s[32]; fe_tobytes(s, f); uint8_t r = 0; for (int i = 0; i < 32; i++) r |= s[i]; return r == 0; }
int fe_eq(const fe *f, const fe *g) { uint8_t a[32], b[32]; fe_tobytes(a, f); fe_tobytes(b, g); return memcmp(a, b, 32) == 0; }
void fe_cmov(fe *f, const fe *g, int b) { if (b) *f = *g; }
void fe_abs(fe *h, const fe *f) { if (fe_isnegative(f)) fe_neg(h, f); else *h = *f; }
int fe_sqrt_ratio_i(fe *r, const fe *u, const fe *v) {
  fe v3, v7, t, check, neg_u, neg_u_i, rp;
  fe_sq(&v3, v); fe_mul(&v3, &v3, v);          /* v^3 */
  fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);        /* v^7 */
  fe_mul(&t, u, &v7); fe_pow22523(&t, &t);     /* (u v^7)^((p-5)/8) */
  fe_mul(r, u, &v3); fe_mul(r, r, &t);
  fe_sq(&check, r); fe_mul(&check, &check, v);
  fe_neg(&neg_u, u); fe_mul(&neg_u_i, &neg_u, &FE_SQRT_M1);
  int correct = fe_eq(&check, u), flipped = fe_eq(&check, &neg_u), flipped_i = fe_eq(&check, &neg_u_i);
  fe_mul(&rp, r, &FE_SQRT_M1);
  if (flipped | flipped_i) *r = rp;
  fe_abs(r, r);
  return correct | flipped;
}
