/* oracle/fe.h -- TEST INFRASTRUCTURE.  GF(2^255-19) in radix 2^51 (five 64-bit limbs, unsigned __int128 products):
 * the same representation as curve25519-dalek's serial u64 backend (un-vendored dependency of the reference,
 * curve25519-dalek 5.0.0-pre.0 @8f6d2ded; call sites src/compressed.rs:28-34, src/elgamal.rs:283-370). */
#ifndef XO_FE_H
#define XO_FE_H
#include <stdint.h>
typedef struct { uint64_t v[5]; } fe;
void fe_0(fe *h); void fe_1(fe *h);
void fe_add(fe *h, const fe *f, const fe *g);
void fe_sub(fe *h, const fe *f, const fe *g);
void fe_neg(fe *h, const fe *f);
void fe_mul(fe *h, const fe *f, const fe *g);
void fe_sq(fe *h, const fe *f);
void fe_sqn(fe *h, const fe *f, int n);
void fe_invert(fe *h, const fe *f);
void fe_pow22523(fe *h, const fe *f);            /* f^((p-5)/8) */
void fe_frombytes(fe *h, const uint8_t s[32]);   /* ignores bit 255 */
void fe_tobytes(uint8_t s[32], const fe *h);     /* canonical */
int fe_isnegative(const fe *f);
int fe_iszero(const fe *f);
int fe_eq(const fe *f, const fe *g);
void fe_cmov(fe *f, const fe *g, int b);
void fe_abs(fe *h, const fe *f);
int fe_sqrt_ratio_i(fe *r, const fe *u, const fe *v); /* RFC 9496 4.2 SQRT_RATIO_M1; returns was_square */
#endif
