/* oracle/ge.c -- TEST INFRASTRUCTURE.  See ge.h. */
#include "ge.h"
#include "consts.h"
#include <stdlib.h>
#include <string.h>
void ge_identity(ge *p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }
void ge_basepoint(ge *p) { p->X = FE_BX; p->Y = FE_BY; fe_1(&p->Z); p->T = FE_BT; }
void ge_to_pn(ge_pn *r, const ge *p) { fe_add(&r->YpX, &p->Y, &p->X); fe_sub(&r->YmX, &p->Y, &p->X); r->Z = p->Z; fe_mul(&r->T2d, &p->T, &FE_D2); }
static void completed_to_ext(ge *r, const fe *X, const fe *Y, const fe *Z, const fe *T) {
  fe x, y, z, t; fe_mul(&x, X, T); fe_mul(&y, Y, Z); fe_mul(&z, Z, T); fe_mul(&t, X, Y); r->X = x; r->Y = y; r->Z = z; r->T = t;
}
void ge_add_pn(ge *r, const ge *p, const ge_pn *q) {
  fe a, b, PP, MM, TT, ZZ, X, Y, Z, T;
  fe_add(&a, &p->Y, &p->X); fe_sub(&b, &p->Y, &p->X);
  fe_mul(&PP, &a, &q->YpX); fe_mul(&MM, &b, &q->YmX); fe_mul(&TT, &p->T, &q->T2d); fe_mul(&ZZ, &p->Z, &q->Z); fe_add(&ZZ, &ZZ, &ZZ);
  fe_sub(&X, &PP, &MM); fe_add(&Y, &PP, &MM); fe_add(&Z, &ZZ, &TT); fe_sub(&T, &ZZ, &TT);
  completed_to_ext(r, &X, &Y, &Z, &T);
}
void ge_sub_pn(ge *r, const ge *p, const ge_pn *q) {
  fe a, b, PM, MP, TT, ZZ, X, Y, Z, T;
  fe_add(&a, &p->Y, &p->X); fe_sub(&b, &p->Y, &p->X);
  fe_mul(&PM, &a, &q->YmX); fe_mul(&MP, &b, &q->YpX); fe_mul(&TT, &p->T, &q->T2d); fe_mul(&ZZ, &p->Z, &q->Z); fe_add(&ZZ, &ZZ, &ZZ);
  fe_sub(&X, &PM, &MP); fe_add(&Y, &PM, &MP); fe_sub(&Z, &ZZ, &TT); fe_add(&T, &ZZ, &TT);
  completed_to_ext(r, &X, &Y, &Z, &T);
}
void ge_add(ge *r, const ge *p, const ge *q) { ge_pn n; ge_to_pn(&n, q); ge_add_pn(r, p, &n); }
void ge_sub(ge *r, const ge *p, const ge *q) { ge_pn n; ge_to_pn(&n, q); ge_sub_pn(r, p, &n); }
void ge_neg(ge *r, const ge *p) { fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T); }
void ge_double(ge *r, const ge *p) {
  fe XX, YY, ZZ2, XpY, S, D, X, Z, T;
  fe_sq(&XX, &p->X); fe_sq(&YY, &p->Y); fe_sq(&ZZ2, &p->Z); fe_add(&ZZ2, &ZZ2, &ZZ2);
  fe_add(&XpY, &p->X, &p->Y); fe_sq(&XpY, &XpY);
  fe_add(&S, &YY, &XX); fe_sub(&D, &YY, &XX);
  fe_sub(&X, &XpY, &S); Z = D; fe_sub(&T, &ZZ2, &D);
  completed_to_ext(r, &X, &S, &Z, &T);
}
void ge_mul_pow2(ge *r, const ge *p, int k) { *r = *p; for (int i = 0; i < k; i++) ge_double(r, r); }
int ge_ristretto_eq(const ge *p, const ge *q) {
  fe a, b, c, d; fe_mul(&a, &p->X, &q->Y); fe_mul(&b, &p->Y, &q->X); fe_mul(&c, &p->X, &q->X); fe_mul(&d, &p->Y, &q->Y);
  return fe_eq(&a, &b) | fe_eq(&c, &d);
}
int ge_ristretto_is_identity(const ge *p) { return fe_iszero(&p->X) | fe_iszero(&p->Y); }
int ristretto_decode(ge *p, const uint8_t s_bytes[32]) {
  fe s, ss, u1, u2, u2s, v, I, dx, dy, one; uint8_t chk[32];
  fe_frombytes(&s, s_bytes); fe_tobytes(chk, &s);
  if (memcmp(chk, s_bytes, 32) != 0 || (s_bytes[0] & 1)) return 0;  /* non-canonical or negative */
  fe_1(&one); fe_sq(&ss, &s); fe_sub(&u1, &one, &ss); fe_add(&u2, &one, &ss); fe_sq(&u2s, &u2);
  fe_sq(&v, &u1); fe_mul(&v, &v, &FE_D); fe_neg(&v, &v); fe_sub(&v, &v, &u2s);      /* -(d u1^2) - u2^2 */
  fe t; fe_mul(&t, &v, &u2s);
  int ok = fe_sqrt_ratio_i(&I, &one, &t);
  fe_mul(&dx, &I, &u2); fe_mul(&dy, &I, &dx); fe_mul(&dy, &dy, &v);
  fe x, y, tt; fe_add(&x, &s, &s); fe_mul(&x, &x, &dx); fe_abs(&x, &x); fe_mul(&y, &u1, &dy); fe_mul(&tt, &x, &y);
  if (!ok || fe_isnegative(&tt) || fe_iszero(&y)) return 0;
  p->X = x; p->Y = y; fe_1(&p->Z); p->T = tt; return 1;
}
void ristretto_encode(uint8_t out[32], const ge *p) {
  fe u1, u2, t, I, den1, den2, zinv, ix, iy, ench, x, y, dinv, s, one;
  fe_1(&one);
  fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t); fe_mul(&u2, &p->X, &p->Y);
  fe_sq(&t, &u2); fe_mul(&t, &t, &u1); fe_sqrt_ratio_i(&I, &one, &t);
  fe_mul(&den1, &I, &u1); fe_mul(&den2, &I, &u2); fe_mul(&zinv, &den1, &den2); fe_mul(&zinv, &zinv, &p->T);
  fe_mul(&ix, &p->X, &FE_SQRT_M1); fe_mul(&iy, &p->Y, &FE_SQRT_M1); fe_mul(&ench, &den1, &FE_INVSQRT_A_MINUS_D);
  fe_mul(&t, &p->T, &zinv); int rotate = fe_isnegative(&t);
  x = p->X; y = p->Y; dinv = den2;
  if (rotate) { x = iy; y = ix; dinv = ench; }
  fe_mul(&t, &x, &zinv); if (fe_isnegative(&t)) fe_neg(&y, &y);
  fe_sub(&s, &p->Z, &y); fe_mul(&s, &s, &dinv); fe_abs(&s, &s); fe_tobytes(out, &s);
}
static void elligator(ge *p, const fe *t0) {
  fe r, u, v, s, sp, c, N, w0, w1, w2, w3, one, t; fe_1(&one);
  fe_sq(&r, t0); fe_mul(&r, &r, &FE_SQRT_M1);
  fe_add(&u, &r, &one); fe_mul(&u, &u, &FE_ONE_MINUS_D_SQ);
  fe_mul(&t, &r, &FE_D); fe_add(&t, &t, &one); fe_neg(&t, &t);          /* -1 - r d */
  fe_add(&v, &r, &FE_D); fe_mul(&v, &v, &t);
  int sq = fe_sqrt_ratio_i(&s, &u, &v);
  fe_mul(&sp, &s, t0); fe_abs(&sp, &sp); fe_neg(&sp, &sp);
  fe_neg(&c, &one);
  if (!sq) { s = sp; c = r; }
  fe_sub(&t, &r, &one); fe_mul(&N, &c, &t); fe_mul(&N, &N, &FE_D_MINUS_ONE_SQ); fe_sub(&N, &N, &v);
  fe_add(&w0, &s, &s); fe_mul(&w0, &w0, &v); fe_mul(&w1, &N, &FE_SQRT_AD_MINUS_ONE);
  fe_sq(&t, &s); fe_sub(&w2, &one, &t); fe_add(&w3, &one, &t);
  fe_mul(&p->X, &w0, &w3); fe_mul(&p->Y, &w2, &w1); fe_mul(&p->Z, &w1, &w3); fe_mul(&p->T, &w0, &w2);
}
void ristretto_from_uniform(ge *p, const uint8_t b[64]) {
  fe t0, t1; ge p0, p1; fe_frombytes(&t0, b); fe_frombytes(&t1, b + 32); elligator(&p0, &t0); elligator(&p1, &t1); ge_add(p, &p0, &p1);
}
/* signed radix-2^w digits, dalek Scalar::as_radix_2w (w in 4..8); returns digit count */
static int radix_2w(int8_t *digits_out_small, int16_t *digits_out, const sc *s, int w) {
  int digits_count = (256 + w - 1) / w; if (w == 8) digits_count += 1;
  int64_t radix = 1 << w, carry = 0; uint64_t mask = (uint64_t)radix - 1;
  for (int i = 0; i < digits_count; i++) {
    int bit = i * w; int64_t coef;
    if (bit >= 256) coef = 0;
    else { int idx = bit / 64, off = bit % 64; uint64_t b = s->v[idx] >> off; if (off + w > 64 && idx < 3) b |= s->v[idx + 1] << (64 - off); coef = (int64_t)(b & mask); }
    coef += carry; carry = (coef + (radix / 2)) >> w; coef -= carry << w;
    if (digits_out) digits_out[i] = (int16_t)coef;
    if (digits_out_small) digits_out_small[i] = (int8_t)coef;
  }
  /* dalek folds the final carry into the last digit for w < 8; for w = 8 the extra digit absorbs it */
  if (w < 8 && carry) { if (digits_out) digits_out[digits_count - 1] += (int16_t)(carry << w); if (digits_out_small) digits_out_small[digits_count - 1] += (int8_t)(carry << w); }
  return digits_count;
}
void ge_scalarmult(ge *r, const sc *s, const ge *p) {
  /* lookup table 1P..8P, radix-16 signed digits, 63 x (4 doublings + add): dalek variable_base::mul */
  ge_pn tab[8]; ge cur = *p; ge_to_pn(&tab[0], &cur);
  for (int i = 1; i < 8; i++) { ge_add_pn(&cur, p, &tab[i - 1]); ge_to_pn(&tab[i], &cur); }
  int8_t d[64]; { /* as_radix_16 */
    uint8_t b[32]; sc_tobytes(b, s);
    for (int i = 0; i < 32; i++) { d[2*i] = b[i] & 15; d[2*i+1] = (b[i] >> 4) & 15; }
    for (int i = 0; i < 63; i++) { int8_t c = (d[i] + 8) >> 4; d[i] -= c << 4; d[i+1] += c; }
  }
  ge acc; ge_identity(&acc);
  for (int i = 63; i >= 0; i--) {
    if (i != 63) ge_mul_pow2(&acc, &acc, 4);
    if (d[i] > 0) ge_add_pn(&acc, &acc, &tab[d[i] - 1]); else if (d[i] < 0) ge_sub_pn(&acc, &acc, &tab[-d[i] - 1]);
    else { ge_pn idn; ge id; ge_identity(&id); ge_to_pn(&idn, &id); ge_add_pn(&acc, &acc, &idn); } /* constant-time shape: always one add */
  }
  *r = acc;
}
void ge_scalarmult_base(ge *r, const sc *s) { ge b; ge_basepoint(&b); ge_scalarmult(r, s, &b); }
void ge_msm_naive(ge *r, const sc *s, const ge *p, size_t n) {
  ge acc, t; ge_identity(&acc); for (size_t i = 0; i < n; i++) { ge_scalarmult(&t, &s[i], &p[i]); ge_add(&acc, &acc, &t); } *r = acc;
}
/* width-w non-adjacent form, dalek Scalar::non_adjacent_form */
static void naf(int8_t out[256], const sc *s, int w) {
  memset(out, 0, 256); uint64_t x[5] = { s->v[0], s->v[1], s->v[2], s->v[3], 0 };
  uint64_t width = 1ULL << w, wmask = width - 1; int pos = 0; uint64_t carry = 0;
  while (pos < 256) {
    int idx = pos / 64, off = pos % 64; uint64_t bits;
    if (off < 64 - w) bits = x[idx] >> off; else bits = (x[idx] >> off) | (x[idx + 1] << (64 - off));
    uint64_t window = carry + (bits & wmask);
    if ((window & 1) == 0) { pos += 1; continue; }
    if (window < width / 2) { carry = 0; out[pos] = (int8_t)window; } else { carry = 1; out[pos] = (int8_t)((int64_t)window - (int64_t)width); }
    pos += w;
  }
}
void ge_msm_straus(ge *r, const sc *s, const ge *p, size_t n) {
  int8_t (*nafs)[256] = malloc(n * 256 + 1); ge_pn (*tabs)[8] = malloc(n * sizeof(ge_pn) * 8 + 1);
  for (size_t i = 0; i < n; i++) {
    naf(nafs[i], &s[i], 5);
    ge p2, cur = p[i]; ge_double(&p2, &p[i]); ge_pn p2n; ge_to_pn(&p2n, &p2); ge_to_pn(&tabs[i][0], &cur);
    for (int j = 1; j < 8; j++) { ge_add_pn(&cur, &cur, &p2n); ge_to_pn(&tabs[i][j], &cur); }   /* odd multiples 1,3,..,15 */
  }
  ge acc; ge_identity(&acc);
  for (int i = 255; i >= 0; i--) {
    ge_double(&acc, &acc);
    for (size_t k = 0; k < n; k++) { int8_t d = nafs[k][i]; if (d > 0) ge_add_pn(&acc, &acc, &tabs[k][d / 2]); else if (d < 0) ge_sub_pn(&acc, &acc, &tabs[k][(-d) / 2]); }
  }
  free(nafs); free(tabs); *r = acc;
}
void ge_msm_pippenger(ge *r, const sc *s, const ge *p, size_t n) {
  int w = n < 500 ? 6 : n < 800 ? 7 : 8;
  int nb = 1 << (w - 1), nd = (256 + w - 1) / w + (w == 8);
  int16_t *dig = malloc(sizeof(int16_t) * n * nd + 2); ge_pn *pn = malloc(sizeof(ge_pn) * n + 1); ge *buckets = malloc(sizeof(ge) * nb);
  for (size_t i = 0; i < n; i++) { radix_2w(NULL, dig + i * nd, &s[i], w); ge_to_pn(&pn[i], &p[i]); }
  ge total; ge_identity(&total);
  for (int c = nd - 1; c >= 0; c--) {
    for (int b = 0; b < nb; b++) ge_identity(&buckets[b]);
    for (size_t i = 0; i < n; i++) { int d = dig[i * nd + c]; if (d > 0) ge_add_pn(&buckets[d - 1], &buckets[d - 1], &pn[i]); else if (d < 0) ge_sub_pn(&buckets[-d - 1], &buckets[-d - 1], &pn[i]); }
    ge inter = buckets[nb - 1], sum = buckets[nb - 1];
    for (int b = nb - 2; b >= 0; b--) { ge_add(&inter, &inter, &buckets[b]); ge_add(&sum, &sum, &inter); }
    if (c != nd - 1) ge_mul_pow2(&total, &total, w);
    ge_add(&total, &total, &sum);
  }
  free(dig); free(pn); free(buckets); *r = total;
}
void ge_msm_vartime(ge *r, const sc *s, const ge *p, size_t n) { if (n < 190) ge_msm_straus(r, s, p, n); else ge_msm_pippenger(r, s, p, n); }
