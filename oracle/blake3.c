/* oracle/blake3.c -- TEST INFRASTRUCTURE.  BLAKE3 (default hash mode, 32-byte output) for the multisig message digest,
 * reference src/tx/verify.rs:267 and src/tx/builder.rs:192-195 (`blake3 1.8.2`).  Pinned against the python `blake3` wheel. */
#include <stdint.h>
#include <string.h>
#include <stddef.h>
static const uint32_t IV[8] = {0x6A09E667,0xBB67AE85,0x3C6EF372,0xA54FF53A,0x510E527F,0x9B05688C,0x1F83D9AB,0x5BE0CD19};
static const uint8_t PERM[16] = {2,6,3,10,7,0,4,13,1,11,12,5,9,14,15,8};
enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };
#define ROR(x,n) (((x) >> (n)) | ((x) << (32 - (n))))
#define G(a,b,c,d,mx,my) do { a = a + b + mx; d = ROR(d ^ a, 16); c = c + d; b = ROR(b ^ c, 12); a = a + b + my; d = ROR(d ^ a, 8); c = c + d; b = ROR(b ^ c, 7); } while (0)
static void compress(const uint32_t cv[8], const uint8_t block[64], uint8_t block_len, uint64_t counter, uint8_t flags, uint32_t out[16]) {
  uint32_t m[16], v[16]; for (int i = 0; i < 16; i++) m[i] = (uint32_t)block[4*i] | (uint32_t)block[4*i+1] << 8 | (uint32_t)block[4*i+2] << 16 | (uint32_t)block[4*i+3] << 24;
  for (int i = 0; i < 8; i++) v[i] = cv[i];
  v[8] = IV[0]; v[9] = IV[1]; v[10] = IV[2]; v[11] = IV[3]; v[12] = (uint32_t)counter; v[13] = (uint32_t)(counter >> 32); v[14] = block_len; v[15] = flags;
  for (int r = 0; r < 7; r++) {
    G(v[0],v[4],v[8],v[12],m[0],m[1]); G(v[1],v[5],v[9],v[13],m[2],m[3]); G(v[2],v[6],v[10],v[14],m[4],m[5]); G(v[3],v[7],v[11],v[15],m[6],m[7]);
    G(v[0],v[5],v[10],v[15],m[8],m[9]); G(v[1],v[6],v[11],v[12],m[10],m[11]); G(v[2],v[7],v[8],v[13],m[12],m[13]); G(v[3],v[4],v[9],v[14],m[14],m[15]);
    uint32_t p[16]; for (int i = 0; i < 16; i++) p[i] = m[PERM[i]]; memcpy(m, p, sizeof m);
  }
  for (int i = 0; i < 8; i++) { out[i] = v[i] ^ v[i+8]; out[i+8] = v[i+8] ^ cv[i]; }
}
/* chaining value of one chunk (<= 1024 bytes); if root, extra flag ROOT on its last block */
static void chunk_cv(const uint8_t *in, size_t n, uint64_t counter, int root, uint32_t cv_out[8]) {
  uint32_t cv[8], out[16]; memcpy(cv, IV, sizeof cv); size_t nblocks = n == 0 ? 1 : (n + 63) / 64;
  for (size_t b = 0; b < nblocks; b++) {
    uint8_t block[64]; memset(block, 0, 64); size_t take = n - b * 64 < 64 ? n - b * 64 : 64; memcpy(block, in + b * 64, take);
    uint8_t flags = 0; if (b == 0) flags |= CHUNK_START; if (b == nblocks - 1) { flags |= CHUNK_END; if (root) flags |= ROOT; }
    compress(cv, block, (uint8_t)take, counter, flags, out); memcpy(cv, out, sizeof cv);
  }
  memcpy(cv_out, cv, sizeof cv);
}
static void parent_cv(const uint32_t l[8], const uint32_t r[8], int root, uint32_t cv_out[8]) {
  uint8_t block[64]; for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) { block[4*i+k] = (uint8_t)(l[i] >> (8*k)); block[32+4*i+k] = (uint8_t)(r[i] >> (8*k)); }
  uint32_t out[16]; compress(IV, block, 64, 0, PARENT | (root ? ROOT : 0), out); memcpy(cv_out, out, 32);
}
/* left subtree takes the largest power-of-two number of chunks strictly less than the total */
static void subtree(const uint8_t *in, size_t n, uint64_t chunk0, int root, uint32_t cv_out[8]) {
  if (n <= 1024) { chunk_cv(in, n, chunk0, root, cv_out); return; }
  size_t chunks = (n + 1023) / 1024, left = 1; while (left * 2 < chunks) left *= 2;
  uint32_t l[8], r[8]; subtree(in, left * 1024, chunk0, 0, l); subtree(in + left * 1024, n - left * 1024, chunk0 + left, 0, r); parent_cv(l, r, root, cv_out);
}
void xo_blake3(const uint8_t *in, size_t n, uint8_t out[32]) { uint32_t cv[8]; subtree(in, n, 0, 1, cv); for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) out[4*i+k] = (uint8_t)(cv[i] >> (8*k)); }
