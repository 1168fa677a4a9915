// host/scalar_host.hpp -- the little mod-l arithmetic the HOST side needs: 64-byte challenge / hash reduction
// (Scalar::from_bytes_mod_order_wide; reference src/transcript.rs:46-51, src/elgamal.rs:64) and the canonical check
// (Scalar::from_canonical_bytes, applied by serde in the reference).  Everything else about scalars runs on the device.
#pragma once
#include <stdint.h>
#include <string.h>

namespace xhe_host {

class ScalarL {
  typedef unsigned __int128 u128;
  static constexpr uint64_t L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
  static constexpr uint64_t R1[4] = {0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL};   // 2^256 mod l
  static constexpr uint64_t R2[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};   // 2^512 mod l
  static constexpr uint64_t NINV = 0xd2b51da312547e1bULL;                                                                          // -l^-1 mod 2^64
  static bool geq_l(const uint64_t a[4]) { for (int i = 3; i >= 0; i--) { if (a[i] != L[i]) return a[i] > L[i]; } return true; }
  // a * b / 2^256 mod l (a < 2^256, b < l)
  static void montmul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * NINV;
      c = ((u128)m * L[0] + t[0]) >> 64;
      for (int j = 1; j < 4; j++) { c += (u128)m * L[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_l(t)) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - L[i] - (uint64_t)bw; t[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
    memcpy(r, t, 32);
  }

 public:
  static bool is_canonical(const uint8_t s[32]) { uint64_t w[4]; memcpy(w, s, 32); return !geq_l(w); }
  // x mod l for a 512-bit little-endian x = lo + 2^256 hi:  lo*R/R + hi*R^2/R
  static void reduce_wide(const uint8_t in[64], uint8_t out[32]) {
    uint64_t lo[4], hi[4], a[4], b[4]; memcpy(lo, in, 32); memcpy(hi, in + 32, 32);
    montmul(a, lo, R1); montmul(b, hi, R2);
    u128 c = 0; uint64_t s[4];
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; s[i] = (uint64_t)c; c >>= 64; }
    if (geq_l(s)) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)s[i] - L[i] - (uint64_t)bw; s[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
    memcpy(out, s, 32);
  }
};

}  // namespace xhe_host
