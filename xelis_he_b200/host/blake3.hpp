// host/blake3.hpp -- BLAKE3 (default mode, 32-byte digest) for the multisig message hash on the HOST side
// (reference src/tx/verify.rs:267: blake3::hash(&bytes[..multisig_index])).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace xhe_host {
namespace b3 {
static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const uint8_t SIGMA[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum : uint8_t { F_CHUNK_START = 1, F_CHUNK_END = 2, F_PARENT = 4, F_ROOT = 8 };
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void mix(uint32_t* v, int a, int b, int c, int d, uint32_t x, uint32_t y) {
  v[a] += v[b] + x; v[d] = rotr(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotr(v[b] ^ v[c], 12);
  v[a] += v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);  v[c] += v[d]; v[b] = rotr(v[b] ^ v[c], 7);
}
static inline void compress(const uint32_t cv[8], const uint8_t block[64], uint32_t len, uint64_t counter, uint32_t flags, uint32_t out[8]) {
  uint32_t m[16], v[16];
  memcpy(m, block, 64);
  memcpy(v, cv, 32); memcpy(v + 8, IV, 16);
  v[12] = (uint32_t)counter; v[13] = (uint32_t)(counter >> 32); v[14] = len; v[15] = flags;
  for (int round = 0;; round++) {
    mix(v, 0, 4, 8, 12, m[0], m[1]); mix(v, 1, 5, 9, 13, m[2], m[3]); mix(v, 2, 6, 10, 14, m[4], m[5]); mix(v, 3, 7, 11, 15, m[6], m[7]);
    mix(v, 0, 5, 10, 15, m[8], m[9]); mix(v, 1, 6, 11, 12, m[10], m[11]); mix(v, 2, 7, 8, 13, m[12], m[13]); mix(v, 3, 4, 9, 14, m[14], m[15]);
    if (round == 6) break;
    uint32_t t[16]; for (int i = 0; i < 16; i++) t[i] = m[SIGMA[i]]; memcpy(m, t, 64);
  }
  for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}
static inline void chunk(const uint8_t* in, size_t n, uint64_t index, bool root, uint32_t cv[8]) {
  memcpy(cv, IV, 32);
  size_t blocks = n ? (n + 63) / 64 : 1;
  for (size_t k = 0; k < blocks; k++) {
    uint8_t buf[64] = {0}; size_t take = n - 64 * k < 64 ? n - 64 * k : 64; memcpy(buf, in + 64 * k, take);
    uint32_t fl = (k == 0 ? F_CHUNK_START : 0) | (k + 1 == blocks ? (F_CHUNK_END | (root ? F_ROOT : 0)) : 0);
    compress(cv, buf, (uint32_t)take, index, fl, cv);
  }
}
static inline void tree(const uint8_t* in, size_t n, uint64_t first_chunk, bool root, uint32_t cv[8]) {
  if (n <= 1024) { chunk(in, n, first_chunk, root, cv); return; }
  size_t chunks = (n + 1023) / 1024, left = 1;
  while (2 * left < chunks) left *= 2;     // largest power of two strictly below the chunk count
  uint32_t lr[16];
  tree(in, 1024 * left, first_chunk, false, lr); tree(in + 1024 * left, n - 1024 * left, first_chunk + left, false, lr + 8);
  compress(IV, (const uint8_t*)lr, 64, 0, F_PARENT | (root ? F_ROOT : 0), cv);
}
}  // namespace b3
static inline void blake3(const uint8_t* in, size_t n, uint8_t out[32]) { uint32_t cv[8]; b3::tree(in, n, 0, true, cv); memcpy(out, cv, 32); }
}  // namespace xhe_host
