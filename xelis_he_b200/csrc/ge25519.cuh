// ge25519.cuh -- Edwards25519 group law in extended coordinates and ristretto255 (RFC 9496) for sm_100a.
//
// Replaces curve25519-dalek's EdwardsPoint / RistrettoPoint / CompressedRistretto (un-vendored dependency of the
// reference; call sites src/compressed.rs:17-106, src/elgamal.rs:39,283-370, src/proofs.rs:50,62,168-179,306-317).
// Layouts chosen for the GPU: decompressed inputs are affine (Z = 1) and kept as affine-Niels triples
// (y+x, y-x, 2dxy) = 96 B, so the Pippenger bucket add is a 7 M mixed addition; accumulators are extended (X:Y:Z:T).
#pragma once
#include "fe25519.cuh"

namespace xhe {

struct ge { fe X, Y, Z, T; };            // extended, x = X/Z, y = Y/Z, xy = T/Z
struct ge_aff { fe x, y; };              // affine (decompressed input)
struct ge_niels { fe ypx, ymx, t2d; };   // affine Niels: (y+x, y-x, 2d*x*y)

#if defined(__CUDA_ARCH__)
#define XHE_FECONST __device__ __constant__ const
#else
#define XHE_FECONST static const
#endif
// constants from oracle/tools/gen_consts.py (RFC 9496 section 4.1 values)
XHE_FECONST uint32_t FE_D[8] = {0x135978a3, 0x75eb4dca, 0x4141d8ab, 0x00700a4d, 0x7779e898, 0x8cc74079, 0x2b6ffe73, 0x52036cee};
XHE_FECONST uint32_t FE_D2[8] = {0x26b2f159, 0xebd69b94, 0x8283b156, 0x00e0149a, 0xeef3d130, 0x198e80f2, 0x56dffce7, 0x2406d9dc};
XHE_FECONST uint32_t FE_SQRT_M1[8] = {0x4a0ea0b0, 0xc4ee1b27, 0xad2fe478, 0x2f431806, 0x3dfbd7a7, 0x2b4d0099, 0x4fc1df0b, 0x2b832480};
XHE_FECONST uint32_t FE_SQRT_AD_MINUS_ONE[8] = {0x497b2e1b, 0x7e97f6a0, 0x1b7854bd, 0xaf9d8e0c, 0x31f5d1fd, 0x0f3cfcc9, 0x2b8348ac, 0x376931bf};
XHE_FECONST uint32_t FE_INVSQRT_A_MINUS_D[8] = {0x805d40ea, 0x99c8fdaa, 0x5a4172be, 0x9d2f1617, 0xfe01d840, 0x16c27b91, 0xcfaffca2, 0x786c8905};
XHE_FECONST uint32_t FE_ONE_MINUS_D_SQ[8] = {0x945fc176, 0xe27c09c1, 0xcd5e350f, 0x2c81a138, 0xbe70dfe4, 0x9994abdd, 0xb2b3e0d7, 0x029072a8};
XHE_FECONST uint32_t FE_D_MINUS_ONE_SQ[8] = {0x44ed4d20, 0x31ad5aaa, 0xb01e1999, 0xd29e4a2c, 0x529b4eeb, 0x4cdcd32f, 0xf66c2241, 0x5968b37a};
XHE_FECONST uint32_t FE_BX[8] = {0x8f25d51a, 0xc9562d60, 0x9525a7b2, 0x692cc760, 0xfdd6dc5c, 0xc0a4e231, 0xcd6e53fe, 0x216936d3};
XHE_FECONST uint32_t FE_BY[8] = {0x66666658, 0x66666666, 0x66666666, 0x66666666, 0x66666666, 0x66666666, 0x66666666, 0x66666666};

XHE_FECONST uint32_t FE_INV_D[8] = {0xcdc9f843, 0x25e0f276, 0x4279542e, 0x0b5dd698, 0xcdb9cf66, 0x2b162114, 0x14d5ce43, 0x40907ed2};

XHE_HD fe fe_const(const uint32_t* c) { fe r; for (int i = 0; i < 8; i++) r.v[i] = c[i]; return r; }

XHE_HD ge ge_identity() { ge r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }
XHE_HD ge ge_from_affine(const ge_aff& a) { ge r; r.X = a.x; r.Y = a.y; r.Z = fe_one(); r.T = fe_mul(a.x, a.y); return r; }
XHE_HD ge_aff ge_aff_identity() { ge_aff r; r.x = fe_zero(); r.y = fe_one(); return r; }
XHE_HD ge_niels niels_from_affine(const ge_aff& a) {
  ge_niels n; n.ypx = fe_add(a.y, a.x); n.ymx = fe_sub(a.y, a.x); n.t2d = fe_mul(fe_mul(a.x, a.y), fe_const(FE_D2)); return n;
}
XHE_HD ge_niels niels_identity() { ge_niels n; n.ypx = fe_one(); n.ymx = fe_one(); n.t2d = fe_zero(); return n; }
XHE_HD ge_niels niels_neg(const ge_niels& n) { ge_niels r; r.ypx = n.ymx; r.ymx = n.ypx; r.t2d = fe_neg(n.t2d); return r; }
XHE_HD ge_niels niels_cneg(const ge_niels& n, bool neg) {
  ge_niels r; r.ypx = fe_select(n.ypx, n.ymx, neg); r.ymx = fe_select(n.ymx, n.ypx, neg); r.t2d = fe_cneg(n.t2d, neg); return r;
}
// extended point (Z = 2) from an affine Niels triple: 1 M instead of a 7 M mixed add onto the identity
XHE_HD ge ge_from_niels(const ge_niels& n) {
  ge r; r.X = fe_sub(n.ypx, n.ymx); r.Y = fe_add(n.ypx, n.ymx); r.Z = fe_zero(); r.Z.v[0] = 2; r.T = fe_mul(n.t2d, fe_const(FE_INV_D)); return r;
}
XHE_HD ge ge_neg(const ge& p) { ge r; r.X = fe_neg(p.X); r.Y = p.Y; r.Z = p.Z; r.T = fe_neg(p.T); return r; }

// mixed addition, extended + affine Niels: 7 M
XHE_HD ge ge_madd(const ge& p, const ge_niels& q) {
  fe a = fe_mul(fe_sub(p.Y, p.X), q.ymx);
  fe b = fe_mul(fe_add(p.Y, p.X), q.ypx);
  fe c = fe_mul(p.T, q.t2d);
  fe d = fe_dbl(p.Z);
  fe e = fe_sub(b, a), f = fe_sub(d, c), g = fe_add(d, c), h = fe_add(b, a);
  ge r; r.X = fe_mul(e, f); r.Y = fe_mul(g, h); r.Z = fe_mul(f, g); r.T = fe_mul(e, h); return r;
}
// full addition, extended + extended: 9 M (8 M + the 2d multiply)
XHE_HD ge ge_add(const ge& p, const ge& q) {
  fe a = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
  fe b = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
  fe c = fe_mul(fe_mul(p.T, q.T), fe_const(FE_D2));
  fe d = fe_dbl(fe_mul(p.Z, q.Z));
  fe e = fe_sub(b, a), f = fe_sub(d, c), g = fe_add(d, c), h = fe_add(b, a);
  ge r; r.X = fe_mul(e, f); r.Y = fe_mul(g, h); r.Z = fe_mul(f, g); r.T = fe_mul(e, h); return r;
}
XHE_HD ge ge_sub(const ge& p, const ge& q) { return ge_add(p, ge_neg(q)); }
// doubling: 4 S + 4 M
XHE_HD ge ge_double(const ge& p) {
  fe xx = fe_sq(p.X), yy = fe_sq(p.Y), zz2 = fe_dbl(fe_sq(p.Z));
  fe xpy2 = fe_sq(fe_add(p.X, p.Y));
  fe s = fe_add(yy, xx), d = fe_sub(yy, xx);
  fe cx = fe_sub(xpy2, s), ct = fe_sub(zz2, d);
  ge r; r.X = fe_mul(cx, ct); r.Y = fe_mul(s, d); r.Z = fe_mul(d, ct); r.T = fe_mul(cx, s); return r;
}
// doubling whose result feeds another doubling: T is not an input of the doubling formulas, so T3 = E * H (one multiply of
// eight) is skipped; the returned T is stale and must not be read
XHE_HD ge ge_double_pz(const ge& p) {
  fe xx = fe_sq(p.X), yy = fe_sq(p.Y), zz2 = fe_dbl(fe_sq(p.Z));
  fe xpy2 = fe_sq(fe_add(p.X, p.Y));
  fe s = fe_add(yy, xx), d = fe_sub(yy, xx);
  fe cx = fe_sub(xpy2, s), ct = fe_sub(zz2, d);
  ge r; r.X = fe_mul(cx, ct); r.Y = fe_mul(s, d); r.Z = fe_mul(d, ct); r.T = p.T; return r;
}
// Ristretto coset identity test (reference src/proofs.rs:62: RistrettoPoint::is_identity): X == 0 || Y == 0
XHE_HD bool ge_ristretto_is_identity(const ge& p) { return fe_iszero(p.X) || fe_iszero(p.Y); }

// RFC 9496 4.2 SQRT_RATIO_M1: r = sqrt(u/v) (or sqrt(i*u/v)), returns was_square.  1 pow22523 = 252 S + 12 M
XHE_HD bool fe_sqrt_ratio_i(fe& r, const fe& u, const fe& v) {
  fe v3 = fe_mul(fe_sq(v), v);
  fe v7 = fe_mul(fe_sq(v3), v);
  fe rr = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
  fe check = fe_mul(v, fe_sq(rr));
  fe neg_u = fe_neg(u);
  fe i = fe_const(FE_SQRT_M1);
  bool correct = fe_eq(check, u), flipped = fe_eq(check, neg_u), flipped_i = fe_eq(check, fe_mul(neg_u, i));
  rr = fe_select(rr, fe_mul(rr, i), flipped || flipped_i);
  r = fe_abs(rr);
  return correct || flipped;
}
// invsqrt specialisation (u = 1): saves the u multiplies
XHE_HD bool fe_invsqrt(fe& r, const fe& v) { return fe_sqrt_ratio_i(r, fe_one(), v); }

// ristretto255 decode (RFC 9496 4.3.1).  Returns false for non-canonical / negative s, non-square, negative t, y == 0.
XHE_HD bool ristretto_decode(ge_aff& out, const uint8_t* bytes) {
  fe s = fe_frombytes(bytes);
  // canonical: re-encoding must match all 32 bytes (this also rejects bit 255 set)
  uint8_t chk[32];
  fe_tobytes(chk, s);
  uint32_t diff = 0;
  for (int i = 0; i < 32; i++) diff |= (uint32_t)(chk[i] ^ bytes[i]);
  bool ok = (diff == 0) && !(bytes[0] & 1);
  fe one = fe_one();
  fe ss = fe_sq(s);
  fe u1 = fe_sub(one, ss), u2 = fe_add(one, ss);
  fe u2s = fe_sq(u2);
  fe v = fe_sub(fe_neg(fe_mul(fe_const(FE_D), fe_sq(u1))), u2s);
  fe I;
  bool sq = fe_invsqrt(I, fe_mul(v, u2s));
  fe dx = fe_mul(I, u2);
  fe dy = fe_mul(fe_mul(I, dx), v);
  fe x = fe_abs(fe_mul(fe_dbl(s), dx));
  fe y = fe_mul(u1, dy);
  fe t = fe_mul(x, y);
  ok = ok && sq && !fe_isneg(t) && !fe_iszero(y);
  out.x = x; out.y = y;
  return ok;
}

// ristretto255 encode (RFC 9496 4.3.2).  Optionally returns 1/Z-normalised affine coordinates of the SAME Edwards
// point (z_inv falls out of the invsqrt), used to turn computed balance points into MSM inputs for free.
XHE_HD void ristretto_encode(uint8_t* out, const ge& p, ge_aff* aff_out = nullptr) {
  fe u1 = fe_mul(fe_add(p.Z, p.Y), fe_sub(p.Z, p.Y));
  fe u2 = fe_mul(p.X, p.Y);
  fe I;
  fe_invsqrt(I, fe_mul(u1, fe_sq(u2)));
  fe den1 = fe_mul(I, u1), den2 = fe_mul(I, u2);
  fe zinv = fe_mul(fe_mul(den1, den2), p.T);
  if (aff_out) {
    // z_inv == 1/Z whenever the point is not in the degenerate set X*Y == 0 (identity coset); callers that need
    // affine output for those points handle them separately (see ge_normalize).
    aff_out->x = fe_mul(p.X, zinv); aff_out->y = fe_mul(p.Y, zinv);
  }
  fe i = fe_const(FE_SQRT_M1);
  fe ix = fe_mul(p.X, i), iy = fe_mul(p.Y, i);
  fe ench = fe_mul(den1, fe_const(FE_INVSQRT_A_MINUS_D));
  bool rotate = fe_isneg(fe_mul(p.T, zinv));
  fe x = fe_select(p.X, iy, rotate), y = fe_select(p.Y, ix, rotate);
  fe dinv = fe_select(den2, ench, rotate);
  y = fe_cneg(y, fe_isneg(fe_mul(x, zinv)));
  fe s = fe_abs(fe_mul(dinv, fe_sub(p.Z, y)));
  fe_tobytes(out, s);
}

// Elligator map (RFC 9496 4.3.4 MAP) and the 64-byte one-way map
XHE_HD ge ristretto_elligator(const fe& t0) {
  fe one = fe_one(), d = fe_const(FE_D);
  fe r = fe_mul(fe_const(FE_SQRT_M1), fe_sq(t0));
  fe u = fe_mul(fe_add(r, one), fe_const(FE_ONE_MINUS_D_SQ));
  fe v = fe_mul(fe_neg(fe_add(fe_mul(r, d), one)), fe_add(r, d));
  fe s;
  bool sq = fe_sqrt_ratio_i(s, u, v);
  fe sp = fe_neg(fe_abs(fe_mul(s, t0)));
  s = fe_select(sp, s, sq);
  fe c = fe_select(r, fe_neg(one), sq);
  fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), fe_const(FE_D_MINUS_ONE_SQ)), v);
  fe w0 = fe_mul(fe_dbl(s), v), w1 = fe_mul(N, fe_const(FE_SQRT_AD_MINUS_ONE));
  fe s2 = fe_sq(s);
  fe w2 = fe_sub(one, s2), w3 = fe_add(one, s2);
  ge p; p.X = fe_mul(w0, w3); p.Y = fe_mul(w2, w1); p.Z = fe_mul(w1, w3); p.T = fe_mul(w0, w2); return p;
}
XHE_HD ge ristretto_from_uniform(const uint8_t* b64) { return ge_add(ristretto_elligator(fe_frombytes(b64)), ristretto_elligator(fe_frombytes(b64 + 32))); }

}  // namespace xhe
