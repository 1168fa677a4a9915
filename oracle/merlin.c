/* oracle/merlin.c -- TEST INFRASTRUCTURE.  See merlin.h. */
#include "merlin.h"
#include "keccak.h"
#include <string.h>
#define STROBE_R 166
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
uint64_t xo_keccak_count = 0;
static void run_f(xo_transcript *s) {
  s->st[s->pos] ^= s->pos_begin; s->st[s->pos + 1] ^= 0x04; s->st[STROBE_R + 1] ^= 0x80;
  uint64_t w[25]; memcpy(w, s->st, 200); xo_keccak_f1600(w); memcpy(s->st, w, 200); xo_keccak_count++;
  s->pos = 0; s->pos_begin = 0;
}
static void absorb(xo_transcript *s, const uint8_t *d, size_t n) { while (n--) { s->st[s->pos++] ^= *d++; if (s->pos == STROBE_R) run_f(s); } }
static void squeeze(xo_transcript *s, uint8_t *d, size_t n) { while (n--) { *d++ = s->st[s->pos]; s->st[s->pos++] = 0; if (s->pos == STROBE_R) run_f(s); } }
static void begin_op(xo_transcript *s, uint8_t flags, int more) {
  if (more) return; /* continuation of the same op */
  uint8_t old_begin = s->pos_begin; s->pos_begin = s->pos + 1; s->cur_flags = flags;
  uint8_t hdr[2] = { old_begin, flags }; absorb(s, hdr, 2);
  if ((flags & (FLAG_C | FLAG_K)) && s->pos != 0) run_f(s);
}
static void meta_ad(xo_transcript *s, const void *d, size_t n, int more) { begin_op(s, FLAG_M | FLAG_A, more); absorb(s, (const uint8_t*)d, n); }
static void ad(xo_transcript *s, const void *d, size_t n, int more) { begin_op(s, FLAG_A, more); absorb(s, (const uint8_t*)d, n); }
static void prf(xo_transcript *s, void *d, size_t n, int more) { begin_op(s, FLAG_I | FLAG_A | FLAG_C, more); squeeze(s, (uint8_t*)d, n); }
static void strobe_init(xo_transcript *s, const char *proto) {
  memset(s, 0, sizeof *s);
  static const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
  memcpy(s->st, hdr, 6); memcpy(s->st + 6, "STROBEv1.0.2", 12);
  uint64_t w[25]; memcpy(w, s->st, 200); xo_keccak_f1600(w); memcpy(s->st, w, 200);
  meta_ad(s, proto, strlen(proto), 0);
}
void xo_transcript_append(xo_transcript *t, const char *label, const void *msg, size_t n) {
  uint32_t len = (uint32_t)n; uint8_t le[4] = { (uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24) };
  meta_ad(t, label, strlen(label), 0); meta_ad(t, le, 4, 1); ad(t, msg, n, 0);
}
void xo_transcript_init(xo_transcript *t, const char *label) { strobe_init(t, "Merlin v1.0"); xo_transcript_append(t, "dom-sep", label, strlen(label)); }
void xo_transcript_append_u64(xo_transcript *t, const char *label, uint64_t v) {
  uint8_t le[8]; for (int i = 0; i < 8; i++) le[i] = (uint8_t)(v >> (8 * i)); xo_transcript_append(t, label, le, 8);
}
void xo_transcript_challenge(xo_transcript *t, const char *label, void *out, size_t n) {
  uint32_t len = (uint32_t)n; uint8_t le[4] = { (uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24) };
  meta_ad(t, label, strlen(label), 0); meta_ad(t, le, 4, 1); prf(t, out, n, 0);
}
