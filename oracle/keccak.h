/* oracle/keccak.h -- TEST INFRASTRUCTURE (CPU oracle; never linked into the product library).
 * Keccak-f[1600], SHA3-256/512, SHAKE256 (FIPS 202).  Stands in for the `sha3 0.10.8` crate the reference
 * uses at src/elgamal.rs:19-22,58-64 and for the Keccak permutation inside merlin 4.0.0. */
#ifndef XO_KECCAK_H
#define XO_KECCAK_H
#include <stddef.h>
#include <stdint.h>
void xo_keccak_f1600(uint64_t st[25]);
typedef struct { uint64_t st[25]; unsigned pos, rate; } xo_sponge;
void xo_sponge_init(xo_sponge *s, unsigned rate_bytes);
void xo_sponge_absorb(xo_sponge *s, const void *data, size_t n);
void xo_sponge_finish(xo_sponge *s, uint8_t domain);          /* pad10*1 with the domain suffix byte */
void xo_sponge_squeeze(xo_sponge *s, void *out, size_t n);
void xo_sha3_256(const void *m, size_t n, uint8_t out[32]);
void xo_sha3_512(const void *m, size_t n, uint8_t out[64]);
void xo_shake256(const void *m, size_t n, void *out, size_t outlen);
#endif
