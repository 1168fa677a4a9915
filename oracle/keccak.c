/* oracle/keccak.c -- TEST INFRASTRUCTURE.  See keccak.h. */
#include "keccak.h"
#include <string.h>
static const uint64_t RC[24] = {
  0x0000000000000001ULL,0x0000000000008082ULL,0x800000000000808aULL,0x8000000080008000ULL,0x000000000000808bULL,0x0000000080000001ULL,
  0x8000000080008081ULL,0x8000000000008009ULL,0x000000000000008aULL,0x0000000000000088ULL,0x0000000080008009ULL,0x000000008000000aULL,
  0x000000008000808bULL,0x800000000000008bULL,0x8000000000008089ULL,0x8000000000008003ULL,0x8000000000008002ULL,0x8000000000000080ULL,
  0x000000000000800aULL,0x800000008000000aULL,0x8000000080008081ULL,0x8000000000008080ULL,0x0000000080000001ULL,0x8000000080008008ULL};
static const int ROTC[24] = {1,3,6,10,15,21,28,36,45,55,2,14,27,41,56,8,25,43,62,18,39,61,20,44};
static const int PILN[24] = {10,7,11,17,18,3,5,16,8,21,24,4,15,23,19,13,12,2,20,14,22,9,6,1};
#define ROL(x,n) (((x) << (n)) | ((x) >> (64 - (n))))
void xo_keccak_f1600(uint64_t st[25]) {
  uint64_t bc[5], t;
  for (int r = 0; r < 24; r++) {
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i+5] ^ st[i+10] ^ st[i+15] ^ st[i+20];
    for (int i = 0; i < 5; i++) { t = bc[(i+4)%5] ^ ROL(bc[(i+1)%5], 1); for (int j = 0; j < 25; j += 5) st[j+i] ^= t; }
    t = st[1];
    for (int i = 0; i < 24; i++) { int j = PILN[i]; uint64_t b = st[j]; st[j] = ROL(t, ROTC[i]); t = b; }
    for (int j = 0; j < 25; j += 5) { for (int i = 0; i < 5; i++) bc[i] = st[j+i]; for (int i = 0; i < 5; i++) st[j+i] ^= (~bc[(i+1)%5]) & bc[(i+2)%5]; }
    st[0] ^= RC[r];
  }
}
void xo_sponge_init(xo_sponge *s, unsigned rate) { memset(s, 0, sizeof *s); s->rate = rate; }
void xo_sponge_absorb(xo_sponge *s, const void *data, size_t n) {
  const uint8_t *d = (const uint8_t*)data; uint8_t *b = (uint8_t*)s->st;
  while (n--) { b[s->pos++] ^= *d++; if (s->pos == s->rate) { xo_keccak_f1600(s->st); s->pos = 0; } }
}
void xo_sponge_finish(xo_sponge *s, uint8_t dom) {
  uint8_t *b = (uint8_t*)s->st; b[s->pos] ^= dom; b[s->rate-1] ^= 0x80; xo_keccak_f1600(s->st); s->pos = 0;
}
void xo_sponge_squeeze(xo_sponge *s, void *out, size_t n) {
  uint8_t *o = (uint8_t*)out; const uint8_t *b = (const uint8_t*)s->st;
  while (n--) { if (s->pos == s->rate) { xo_keccak_f1600(s->st); s->pos = 0; } *o++ = b[s->pos++]; }
}
void xo_sha3_256(const void *m, size_t n, uint8_t out[32]) { xo_sponge s; xo_sponge_init(&s,136); xo_sponge_absorb(&s,m,n); xo_sponge_finish(&s,0x06); xo_sponge_squeeze(&s,out,32); }
void xo_sha3_512(const void *m, size_t n, uint8_t out[64]) { xo_sponge s; xo_sponge_init(&s,72); xo_sponge_absorb(&s,m,n); xo_sponge_finish(&s,0x06); xo_sponge_squeeze(&s,out,64); }
void xo_shake256(const void *m, size_t n, void *out, size_t outlen) { xo_sponge s; xo_sponge_init(&s,136); xo_sponge_absorb(&s,m,n); xo_sponge_finish(&s,0x1f); xo_sponge_squeeze(&s,out,outlen); }
