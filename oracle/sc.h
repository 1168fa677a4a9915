/* oracle/sc.h -- TEST INFRASTRUCTURE.  Scalars mod l = 2^252 + 27742317777372353535851937790883648493, four 64-bit
 * limbs, Montgomery reduction with R = 2^256.  Mirrors curve25519-dalek's `Scalar` semantics (two Montgomery
 * reductions per multiplication; wide reduction = lo*R/R + hi*R^2/R) used at src/proofs.rs:162-198,304-347,
 * src/transcript.rs:46-51, src/elgamal.rs:64. */
#ifndef XO_SC_H
#define XO_SC_H
#include <stdint.h>
typedef struct { uint64_t v[4]; } sc;
void sc_0(sc *r); void sc_1(sc *r);
void sc_from_u64(sc *r, uint64_t x);
int  sc_frombytes_canonical(sc *r, const uint8_t s[32]); /* 1 iff s < l */
void sc_frombytes_mod_order(sc *r, const uint8_t s[32]);
void sc_frombytes_wide(sc *r, const uint8_t s[64]);
void sc_tobytes(uint8_t s[32], const sc *a);
void sc_add(sc *r, const sc *a, const sc *b);
void sc_sub(sc *r, const sc *a, const sc *b);
void sc_neg(sc *r, const sc *a);
void sc_mul(sc *r, const sc *a, const sc *b);
void sc_muladd(sc *r, const sc *a, const sc *b, const sc *c); /* a*b + c */
void sc_invert(sc *r, const sc *a);
void sc_batch_invert(sc *xs, int n, sc *allinv); /* in place; allinv = prod of inverses (may be NULL) */
int  sc_iszero(const sc *a);
int  sc_eq(const sc *a, const sc *b);
#endif
