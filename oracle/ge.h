/* oracle/ge.h -- TEST INFRASTRUCTURE.  Edwards25519 extended-coordinate group ops, ristretto255 (RFC 9496)
 * decode / encode / one-way map, and the multiscalar multiplications of curve25519-dalek restated with the
 * same algorithm choices (Straus width-5 NAF below 190 points, Pippenger w = 6/7/8 signed radix-2^w above;
 * constant-time-shaped radix-16 variable-base multiplication).  Reference call sites: src/proofs.rs:50,62;
 * src/compressed.rs:19,30,46-47,58-60,73,80,95,102; src/elgamal.rs:22,39,61,283-370. */
#ifndef XO_GE_H
#define XO_GE_H
#include "fe.h"
#include "sc.h"
#include <stddef.h>
typedef struct { fe X, Y, Z, T; } ge;          /* extended (X:Y:Z:T), x=X/Z, y=Y/Z, xy=T/Z */
typedef struct { fe YpX, YmX, Z, T2d; } ge_pn; /* projective Niels */
void ge_identity(ge *p);
void ge_basepoint(ge *p);
void ge_add(ge *r, const ge *p, const ge *q);
void ge_sub(ge *r, const ge *p, const ge *q);
void ge_neg(ge *r, const ge *p);
void ge_double(ge *r, const ge *p);
void ge_to_pn(ge_pn *r, const ge *p);
void ge_add_pn(ge *r, const ge *p, const ge_pn *q);
void ge_sub_pn(ge *r, const ge *p, const ge_pn *q);
void ge_mul_pow2(ge *r, const ge *p, int k);
int  ge_ristretto_eq(const ge *p, const ge *q);      /* coset equality */
int  ge_ristretto_is_identity(const ge *p);
int  ristretto_decode(ge *p, const uint8_t s[32]);   /* 1 ok / 0 invalid */
void ristretto_encode(uint8_t s[32], const ge *p);
void ristretto_from_uniform(ge *p, const uint8_t b[64]);
void ge_scalarmult(ge *r, const sc *s, const ge *p);          /* dalek variable_base::mul shape */
void ge_scalarmult_base(ge *r, const sc *s);
void ge_msm_straus(ge *r, const sc *s, const ge *p, size_t n);
void ge_msm_pippenger(ge *r, const sc *s, const ge *p, size_t n);
void ge_msm_vartime(ge *r, const sc *s, const ge *p, size_t n); /* dalek dispatch: n < 190 ? straus : pippenger */
void ge_msm_naive(ge *r, const sc *s, const ge *p, size_t n);
#endif
