// host/keccak.hpp -- Keccak-f[1600], SHA3-512, SHAKE256 and the STROBE-128 / Merlin transcript for the HOST side of
// the product (north_star: "Rust host code keeps Merlin transcript / Fiat-Shamir challenge derivation").
// Stands in for the `sha3` and `merlin` crates at the reference call sites src/elgamal.rs:53-65,
// src/transcript.rs:37-111, src/tx/verify.rs:146-158.  Independent of oracle/ (which is test-only).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace xhe_host {

static inline uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// fully unrolled-by-lane Keccak-f[1600]
static inline void keccak_f1600(uint64_t s[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
      0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
      0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  for (int r = 0; r < 24; r++) {
    uint64_t c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20], c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21], c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22],
             c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23], c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
    uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
    uint64_t b[25];
    b[0] = s[0] ^ d0;              b[10] = rotl64(s[1] ^ d1, 1);   b[20] = rotl64(s[2] ^ d2, 62);  b[5] = rotl64(s[3] ^ d3, 28);   b[15] = rotl64(s[4] ^ d4, 27);
    b[16] = rotl64(s[5] ^ d0, 36); b[1] = rotl64(s[6] ^ d1, 44);   b[11] = rotl64(s[7] ^ d2, 6);   b[21] = rotl64(s[8] ^ d3, 55);  b[6] = rotl64(s[9] ^ d4, 20);
    b[7] = rotl64(s[10] ^ d0, 3);  b[17] = rotl64(s[11] ^ d1, 10); b[2] = rotl64(s[12] ^ d2, 43);  b[12] = rotl64(s[13] ^ d3, 25); b[22] = rotl64(s[14] ^ d4, 39);
    b[23] = rotl64(s[15] ^ d0, 41); b[8] = rotl64(s[16] ^ d1, 45); b[18] = rotl64(s[17] ^ d2, 15); b[3] = rotl64(s[18] ^ d3, 21);  b[13] = rotl64(s[19] ^ d4, 8);
    b[14] = rotl64(s[20] ^ d0, 18); b[24] = rotl64(s[21] ^ d1, 2); b[9] = rotl64(s[22] ^ d2, 61);  b[19] = rotl64(s[23] ^ d3, 56); b[4] = rotl64(s[24] ^ d4, 14);
    for (int y = 0; y < 25; y += 5) {
      s[y + 0] = b[y + 0] ^ (~b[y + 1] & b[y + 2]); s[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]); s[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
      s[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y + 0]); s[y + 4] = b[y + 4] ^ (~b[y + 0] & b[y + 1]);
    }
    s[0] ^= RC[r];
  }
}

struct Sponge {
  uint64_t st[25]; unsigned pos, rate;
  explicit Sponge(unsigned rate_bytes) : pos(0), rate(rate_bytes) { memset(st, 0, sizeof st); }
  void absorb(const void* data, size_t n) {
    const uint8_t* d = (const uint8_t*)data; uint8_t* b = (uint8_t*)st;
    while (n) {
      size_t take = rate - pos < n ? rate - pos : n;
      for (size_t i = 0; i < take; i++) b[pos + i] ^= d[i];
      pos += (unsigned)take; d += take; n -= take;
      if (pos == rate) { keccak_f1600(st); pos = 0; }
    }
  }
  void finish(uint8_t domain) { uint8_t* b = (uint8_t*)st; b[pos] ^= domain; b[rate - 1] ^= 0x80; keccak_f1600(st); pos = 0; }
  void squeeze(void* out, size_t n) {
    uint8_t* o = (uint8_t*)out; const uint8_t* b = (const uint8_t*)st;
    while (n) {
      if (pos == rate) { keccak_f1600(st); pos = 0; }
      size_t take = rate - pos < n ? rate - pos : n;
      memcpy(o, b + pos, take); pos += (unsigned)take; o += take; n -= take;
    }
  }
};
static inline void sha3_512(const void* m, size_t n, uint8_t out[64]) { Sponge s(72); s.absorb(m, n); s.finish(0x06); s.squeeze(out, 64); }
static inline void shake256(const void* m, size_t n, void* out, size_t outlen) { Sponge s(136); s.absorb(m, n); s.finish(0x1f); s.squeeze(out, outlen); }

// Merlin transcript ("Merlin v1.0" over STROBE-128, security 128 => rate 166)
class Transcript {
 public:
  explicit Transcript(const char* label) {
    memset(st_, 0, sizeof st_); pos_ = 0; pos_begin_ = 0; cur_flags_ = 0;
    static const uint8_t hdr[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(st_, hdr, 6); memcpy(st_ + 6, "STROBEv1.0.2", 12);
    permute();
    meta_ad("Merlin v1.0", 11, false);
    append("dom-sep", label, strlen(label));
  }
  void append(const char* label, const void* msg, size_t n) {
    uint8_t le[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    meta_ad(label, strlen(label), false); meta_ad(le, 4, true); ad(msg, n);
  }
  void append_u64(const char* label, uint64_t v) { uint8_t le[8]; for (int i = 0; i < 8; i++) le[i] = (uint8_t)(v >> (8 * i)); append(label, le, 8); }
  void challenge(const char* label, void* out, size_t n) {
    uint8_t le[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    meta_ad(label, strlen(label), false); meta_ad(le, 4, true);
    begin_op(FLAG_I | FLAG_A | FLAG_C);
    uint8_t* o = (uint8_t*)out;
    for (size_t i = 0; i < n; i++) { o[i] = st_[pos_]; st_[pos_++] = 0; if (pos_ == R) run_f(); }
  }
  uint64_t permutations = 0;

 private:
  enum { R = 166, FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
  alignas(8) uint8_t st_[200]; uint8_t pos_, pos_begin_, cur_flags_;
  void permute() { keccak_f1600(reinterpret_cast<uint64_t*>(st_)); permutations++; }
  void run_f() { st_[pos_] ^= pos_begin_; st_[pos_ + 1] ^= 0x04; st_[R + 1] ^= 0x80; permute(); pos_ = 0; pos_begin_ = 0; }
  void absorb(const uint8_t* d, size_t n) { for (size_t i = 0; i < n; i++) { st_[pos_++] ^= d[i]; if (pos_ == R) run_f(); } }
  void begin_op(uint8_t flags) {
    uint8_t old = pos_begin_; pos_begin_ = pos_ + 1; cur_flags_ = flags;
    uint8_t h[2] = {old, flags}; absorb(h, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
  }
  void meta_ad(const void* d, size_t n, bool more) { if (!more) begin_op(FLAG_M | FLAG_A); absorb((const uint8_t*)d, n); }
  void ad(const void* d, size_t n) { begin_op(FLAG_A); absorb((const uint8_t*)d, n); }
};

}  // namespace xhe_host
