/* oracle/merlin.h -- TEST INFRASTRUCTURE.  STROBE-128/Keccak-f[1600] as used by Merlin ("Merlin v1.0"),
 * restating the un-vendored `merlin 4.0.0 @ee9ef32e` crate the reference calls at
 * src/tx/verify.rs:152-156 and src/transcript.rs:37-111.  Pinned by the public Merlin KAT (tests/test_oracle_kat.py). */
#ifndef XO_MERLIN_H
#define XO_MERLIN_H
#include <stddef.h>
#include <stdint.h>
typedef struct { uint8_t st[200]; uint8_t pos, pos_begin, cur_flags; } xo_transcript;
void xo_transcript_init(xo_transcript *t, const char *label);
void xo_transcript_append(xo_transcript *t, const char *label, const void *msg, size_t n);
void xo_transcript_append_u64(xo_transcript *t, const char *label, uint64_t v);
void xo_transcript_challenge(xo_transcript *t, const char *label, void *out, size_t n);
extern uint64_t xo_keccak_count; /* permutation counter (SURVEY appendix C bookkeeping) */
#endif
