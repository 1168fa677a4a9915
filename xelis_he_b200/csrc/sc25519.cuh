// sc25519.cuh -- scalars mod l = 2^252 + 27742317777372353535851937790883648493 in eight 32-bit limbs.
//
// Replaces curve25519-dalek's `Scalar` arithmetic used by the verification-weight expansion
// (reference src/proofs.rs:162-198,304-347; bulletproofs verification scalars).
// Montgomery form (R = 2^256) is used inside kernels: one CIOS product is 8 rows of (8 + 4 + 1) limb products --
// the m*l row only needs the four low limbs of l plus a shift, because l = 2^252 + c with c < 2^125.
#pragma once
#include "fe25519.cuh"

namespace xhe {

struct sc { uint32_t v[8]; };

#if defined(__CUDA_ARCH__)
#define XHE_CONST __device__ __constant__ const
#else
#define XHE_CONST static const
#endif
// l, R mod l, R^2 mod l, -l^-1 mod 2^32 (oracle/tools/gen_consts.py)
XHE_CONST uint32_t SC_L[8] = {0x5cf5d3ed, 0x5812631a, 0xa2f79cd6, 0x14def9de, 0x00000000, 0x00000000, 0x00000000, 0x10000000};
XHE_CONST uint32_t SC_R1[8] = {0x8d98951d, 0xd6ec3174, 0x737dcf70, 0xc6ef5bf4, 0xfffffffe, 0xffffffff, 0xffffffff, 0x0fffffff};
XHE_CONST uint32_t SC_RR[8] = {0x449c0f01, 0xa40611e3, 0x68859347, 0xd00e1ba7, 0x17f5be65, 0xceec73d2, 0x7c309a3d, 0x0399411b};
#define XHE_SC_LFACTOR 0x12547e1bu

XHE_HD sc sc_zero() { sc r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
XHE_HD sc sc_load_const(const uint32_t* c) { sc r; for (int i = 0; i < 8; i++) r.v[i] = c[i]; return r; }
XHE_HD sc sc_from_u64(uint64_t x) { sc r = sc_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }
XHE_HD sc sc_frombytes(const uint8_t* s) {
  sc r;
  for (int i = 0; i < 8; i++) r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
  return r;
}
XHE_HD void sc_tobytes(uint8_t* s, const sc& a) {
  for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)a.v[i]; s[4 * i + 1] = (uint8_t)(a.v[i] >> 8); s[4 * i + 2] = (uint8_t)(a.v[i] >> 16); s[4 * i + 3] = (uint8_t)(a.v[i] >> 24); }
}
XHE_HD bool sc_iszero(const sc& a) { uint32_t r = 0; for (int i = 0; i < 8; i++) r |= a.v[i]; return r == 0; }
XHE_HD bool sc_eq(const sc& a, const sc& b) { uint32_t r = 0; for (int i = 0; i < 8; i++) r |= a.v[i] ^ b.v[i]; return r == 0; }

// a >= l ?
XHE_HD bool sc_geq_l(const uint32_t* a) {
  uint32_t t[8];
  return sub8(t, a, SC_L) == 0;
}
XHE_HD bool sc_is_canonical(const sc& a) { return !sc_geq_l(a.v); }
// conditional subtract of l (input < 2l)
XHE_HD void sc_csub_l(uint32_t* a, uint32_t extra_hi) {
  uint32_t t[8];
  uint32_t bw = sub8(t, a, SC_L);
  bool take = extra_hi != 0 || bw == 0;
  for (int i = 0; i < 8; i++) a[i] = take ? t[i] : a[i];
}

XHE_HD sc sc_add(const sc& a, const sc& b) {  // inputs < l
  sc r;
  uint32_t c = add8(r.v, a.v, b.v);   // < 2l < 2^254: c is always 0
  sc_csub_l(r.v, c);
  return r;
}
XHE_HD sc sc_sub(const sc& a, const sc& b) {  // inputs < l
  sc r;
  uint32_t bw = sub8(r.v, a.v, b.v);
  uint32_t t[8];
  add8(t, r.v, SC_L);
  for (int i = 0; i < 8; i++) r.v[i] = bw ? t[i] : r.v[i];
  return r;
}
XHE_HD sc sc_neg(const sc& a) { return sc_sub(sc_zero(), a); }

// Montgomery reduction (t + M*l) / 2^256 of a 16-limb t < l*2^256; result < l.  M = sum m_i 2^(32 i) is fixed one
// limb at a time (m_i = limb_i * (-l^-1) mod 2^32).  Because l = 2^252 + c with c < 2^125, M*l splits into M*c (8 rows of
// four limb products, kept in their own even/odd slot arrays so that every chain's carry word is still untouched when
// the chain ends) and M << 252 (shifts only).  `carry` tracks what the cancelled low limbs pushed upward.
XHE_HD sc sc_redc(const uint32_t* t) {
  uint32_t re[14], ro[14], m[9];
#pragma unroll
  for (int i = 0; i < 14; i++) { re[i] = 0; ro[i] = 0; }
  const uint32_t c0 = 0x5cf5d3edu, c1 = 0x5812631au, c2 = 0xa2f79cd6u, c3 = 0x14def9deu;
  uint32_t carry = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    // limb i of t + M*c + (M << 252), before and after this row's contribution (which clears its low word)
    uint32_t fix = (i == 7) ? (uint32_t)(m[0] << 28) : 0u;
    uint64_t x = (uint64_t)t[i] + re[i] + carry + fix;
    if (i > 0) x += ro[i - 1];
    m[i] = (uint32_t)x * XHE_SC_LFACTOR;
    if ((i & 1) == 0) {
      mad2w(re + i, re[i + 4], c0, c2, m[i]);
      mad2w(ro + i, ro[i + 4], c1, c3, m[i]);
    } else {
      mad2w(ro + i - 1, ro[i + 3], c0, c2, m[i]);
      mad2w(re + i + 1, re[i + 5], c1, c3, m[i]);
    }
    uint64_t y = (uint64_t)t[i] + re[i] + carry + fix;
    if (i > 0) y += ro[i - 1];
    carry = (uint32_t)(y >> 32);
  }
  m[8] = 0;
  // limbs 8..15: t_hi + (M*c)_hi + (M << 252)_hi + carry
  uint32_t h[8], u[8], v[8], w[8], z[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    h[k] = (m[k] >> 4) | (m[k + 1] << 28);
    u[k] = 8 + k < 14 ? re[8 + k] : 0;
    v[k] = 7 + k < 14 ? ro[7 + k] : 0;
  }
  sc r;
  add8(w, t + 8, h);
  add8(z, u, v);
  add8(r.v, w, z);
  addw8(r.v, carry);      // the total is < 2l < 2^254: none of these carries out
  sc_csub_l(r.v, 0);
  return r;
}

// Montgomery product a*b*R^-1 mod l (R = 2^256) for a*b < l*R; result < l
XHE_HD sc sc_montmul(const sc& a, const sc& b) {
  uint32_t t[16];
  mul512(t, a.v, b.v);
  return sc_redc(t);
}
XHE_HD sc sc_montsq(const sc& a) {
  uint32_t t[16];
  sq512(t, a.v);
  return sc_redc(t);
}
XHE_HD sc sc_to_mont(const sc& a) { return sc_montmul(a, sc_load_const(SC_RR)); }         // a*R
XHE_HD sc sc_from_mont(const sc& a) { sc one = sc_zero(); one.v[0] = 1; return sc_montmul(a, one); }
XHE_HD sc sc_mul(const sc& a, const sc& b) { return sc_montmul(sc_montmul(a, b), sc_load_const(SC_RR)); }  // plain-form product
// reduce a 256-bit / 512-bit little-endian integer mod l (dalek from_bytes_mod_order / _wide)
XHE_HD sc sc_reduce256(const sc& a) { return sc_montmul(a, sc_load_const(SC_R1)); }
XHE_HD sc sc_reduce512(const sc& lo, const sc& hi) { return sc_add(sc_montmul(lo, sc_load_const(SC_R1)), sc_montmul(hi, sc_load_const(SC_RR))); }
// a^(l-2) for a in Montgomery form -> Montgomery form
XHE_HD sc sc_mont_invert(const sc& a) {
  sc acc = sc_load_const(SC_R1);
  for (int i = 252; i >= 0; i--) {
    acc = sc_montsq(acc);
    uint32_t w = SC_L[i >> 5] - ((i >> 5) == 0 ? 2u : 0u);  // (l-2): only limb 0 changes (0x5cf5d3ed - 2, no borrow)
    if ((w >> (i & 31)) & 1u) acc = sc_montmul(acc, a);
  }
  return acc;
}

}  // namespace xhe
