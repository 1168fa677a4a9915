// oct.cuh -- warp-cooperative GF(2^255-19) and Edwards arithmetic for the SERIAL chains of the MSM tail (the c * W dependent
// doublings of the Horner recombination, msm.cu).  A field element is spread over the 8 lanes of an aligned group, one
// radix-2^32 limb per lane; a point (X : Y : Z : T) over the 4 groups of a warp -- ONE register per lane.  A field multiply is
// then 8 limb products per lane (every lane forms its own column k and column k + 8 of the schoolbook product from operands
// fetched with warp shuffles, folds 2^256 = 38 locally) and a three-step carry resolution across the lanes: two ripple steps
// (one shuffle each) that bring every limb below 2^32 + 2^21, and a carry-lookahead step over ballot masks (a lane generates
// if its limb overflowed, propagates if it is 0xffffffff; the carries are the carries of the integer sum of the two masks).
// The top limb is cut at bit 31 (2^255 = 19) in the ripple steps, so no carry can leave lane 7 in the last one.
//
// Measured purpose: a doubling on one lane is a ~2,500-cycle dependency chain, shared by the four lanes of a
// quad (quad.cuh) still ~1,300 ns / 2,550 cycles; the 255 doublings of a 2^20-point MSM were 0.33 ms of its 2.5.  Here the
// longest chain in a multiply is 8 dependent multiply-adds, and the four coordinates' products of a formula layer run
// side by side in the four groups.
//
// Values are ordinary `fe` values (8 limbs, in [0, 2^256), not necessarily below p): memory layout and semantics are those
// of fe25519.cuh / ge25519.cuh, so word l of a stored `ge` is lane l's register.  All 32 lanes of a warp must call these
// functions together.  Formulas: the same complete a = -1 formulas as ge_double / ge_add (ge25519.cuh).
#pragma once
#include "ge25519.cuh"

namespace xhe {

#define OCT_FULL 0xffffffffu

// carry-lookahead step: every lane holds s < 2^32 + 1 split as (vlo, g); lane 7 neither generates nor propagates
__device__ __forceinline__ uint32_t oct_lookahead(uint32_t vlo, uint32_t g) {
  const uint32_t lane = threadIdx.x & 31u, sh = lane & 24u;
  const uint32_t G = (__ballot_sync(OCT_FULL, g != 0u) >> sh) & 0xffu;
  const uint32_t P = (__ballot_sync(OCT_FULL, vlo == 0xffffffffu) >> sh) & 0xffu;
  const uint32_t x = G | P, carr = (x + G) ^ x ^ G;             // bit k = carry into lane k (G and P are disjoint)
  return vlo + ((carr >> (lane & 7u)) & 1u);
}
// second ripple step + lookahead: lane value v = lo + 2^32 hi, hi < 2^15
__device__ __forceinline__ uint32_t oct_norm2(uint32_t lo, uint32_t hi) {
  const uint32_t lane = threadIdx.x & 31u, k = lane & 7u;
  const bool top = k == 7u;
  const uint32_t w0 = top ? (lo & 0x7fffffffu) : lo;
  const uint32_t c = top ? (((hi << 1) | (lo >> 31)) * 19u) : hi;
  const uint32_t cin = __shfl_sync(OCT_FULL, c, (lane & 24u) | ((k + 7u) & 7u));      // from lane k - 1 of the group (lane 7 wraps to lane 0 with 2^255 = 19)
  const uint32_t s = w0 + cin;
  return oct_lookahead(s, s < w0 ? 1u : 0u);
}
__device__ __forceinline__ uint32_t oct_norm64(unsigned long long v) { return oct_norm2((uint32_t)v, (uint32_t)(v >> 32)); }      // v < 2^47
// first ripple step for a three-word lane value r = r0 + 2^32 r1 + 2^64 r2 < 2^73
__device__ __forceinline__ uint32_t oct_norm3(uint32_t r0, uint32_t r1, uint32_t r2) {
  const uint32_t lane = threadIdx.x & 31u, k = lane & 7u;
  const bool top = k == 7u;
  unsigned long long c = ((unsigned long long)r2 << 32) | r1;                          // < 2^41
  const uint32_t w0 = top ? (r0 & 0x7fffffffu) : r0;
  if (top) c = ((c << 1) | (r0 >> 31)) * 19ull;                                        // < 2^47
  const uint32_t src = (lane & 24u) | ((k + 7u) & 7u);
  const uint32_t clo = __shfl_sync(OCT_FULL, (uint32_t)c, src), chi = __shfl_sync(OCT_FULL, (uint32_t)(c >> 32), src);
  return oct_norm64((((unsigned long long)chi << 32) | clo) + w0);
}

// limb k of 4p = 2^257 - 76 as 33-bit lane constants (so that a - b + 4p is positive in every lane)
__device__ __forceinline__ unsigned long long oct_c4p() { return (threadIdx.x & 7u) ? 0x1fffffffeull : 0x1ffffffb4ull; }
__device__ __forceinline__ unsigned long long oct_c8p() { return (threadIdx.x & 7u) ? 0x3fffffffcull : 0x3ffffff68ull; }

__device__ __forceinline__ uint32_t oct_add(uint32_t a, uint32_t b) { return oct_norm64((unsigned long long)a + b); }
__device__ __forceinline__ uint32_t oct_sub(uint32_t a, uint32_t b) { return oct_norm64((unsigned long long)a + oct_c4p() - b); }

// Four independent lane values (each < 2^47) normalised side by side.  A warp issues in order and the compiler keeps warp-level
// primitives in program order, so four calls of oct_norm64 in a row would pay four shuffle + ballot round trips one after the
// other; here the four shuffles, then the eight ballots, are issued back to back and their latencies overlap.
__device__ __forceinline__ void oct_norm64_x4(const unsigned long long v[4], uint32_t out[4]) {
  const uint32_t lane = threadIdx.x & 31u, k = lane & 7u, sh = lane & 24u, src = sh | ((k + 7u) & 7u);
  const bool top = k == 7u;
  uint32_t w0[4], c[4], s[4], g[4], G[4], P[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t lo = (uint32_t)v[q], hi = (uint32_t)(v[q] >> 32);
    w0[q] = top ? (lo & 0x7fffffffu) : lo;
    c[q] = top ? (((hi << 1) | (lo >> 31)) * 19u) : hi;
  }
#pragma unroll
  for (int q = 0; q < 4; q++) c[q] = __shfl_sync(OCT_FULL, c[q], src);
#pragma unroll
  for (int q = 0; q < 4; q++) { s[q] = w0[q] + c[q]; g[q] = s[q] < w0[q] ? 1u : 0u; }
#pragma unroll
  for (int q = 0; q < 4; q++) { G[q] = __ballot_sync(OCT_FULL, g[q] != 0u); P[q] = __ballot_sync(OCT_FULL, s[q] == 0xffffffffu); }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t Gq = (G[q] >> sh) & 0xffu, Pq = (P[q] >> sh) & 0xffu, x = Gq | Pq, carr = (x + Gq) ^ x ^ Gq;
    out[q] = s[q] + ((carr >> k) & 1u);
  }
}

// product of the group's two field elements; a, b = this lane's limbs
__device__ __forceinline__ uint32_t oct_mul(uint32_t a, uint32_t b) {
  const uint32_t lane = threadIdx.x & 31u, base = lane & 24u, k = lane & 7u;
  uint32_t ai[8], bj[8];
#pragma unroll
  for (uint32_t i = 0; i < 8; i++) { ai[i] = __shfl_sync(OCT_FULL, a, base | i); bj[i] = __shfl_sync(OCT_FULL, b, base | ((k - i) & 7u)); }      // all sixteen in flight together
  uint32_t A[2] = {0u, 0u}, B[2] = {0u, 0u}, A2 = 0u, B2 = 0u;      // column k and column k + 8, three words each
#pragma unroll
  for (uint32_t i = 0; i < 8; i++) {
    const uint32_t m = (i <= k) ? 0xffffffffu : 0u;
    mad1w(A, A2, ai[i] & m, bj[i]);
    mad1w(B, B2, ai[i] & ~m, bj[i]);
  }
  // R = A + 38 B < 2^73
  uint32_t R[3] = {A[0], A[1], A2};
  mad1w(R, R[2], B[0], 38u);
  mad1w_nc(R + 1, B[1], 38u);
  R[2] += 38u * B2;
  return oct_norm3(R[0], R[1], R[2]);
}

// ---- points: lane 8 c + k holds limb k of coordinate c (X, Y, Z, T) ----
__device__ __forceinline__ uint32_t oct_identity() { const uint32_t lane = threadIdx.x & 31u; return (lane == 8u || lane == 16u) ? 1u : 0u; }

// doubling (ge_double's formulas with (X + Y)^2 - X^2 - Y^2 taken as 2 X Y): two multiply layers, one layer of linear combinations
__device__ __forceinline__ uint32_t oct_double(uint32_t p) {
  const uint32_t lane = threadIdx.x & 31u, k = lane & 7u, c = lane >> 3;
  const uint32_t x = __shfl_sync(OCT_FULL, p, k), y = __shfl_sync(OCT_FULL, p, 8u | k);
  // groups: X^2, Y^2, Z^2, X Y
  const uint32_t m1 = oct_mul(c == 3u ? x : p, c == 3u ? y : p);
  const unsigned long long xx = __shfl_sync(OCT_FULL, m1, k), yy = __shfl_sync(OCT_FULL, m1, 8u | k), zz = __shfl_sync(OCT_FULL, m1, 16u | k), xy = __shfl_sync(OCT_FULL, m1, 24u | k);
  // s = yy + xx ; d = yy - xx ; cx = 2 xy ; ct = 2 zz - d
  const unsigned long long v[4] = {yy + xx, yy + oct_c4p() - xx, 2ull * xy, 2ull * zz + xx + oct_c4p() - yy};
  uint32_t n[4]; oct_norm64_x4(v, n);
  const uint32_t s = n[0], d = n[1], cx = n[2], ct = n[3];
  // X3 = cx ct, Y3 = s d, Z3 = d ct, T3 = cx s
  const uint32_t l = c == 1u ? s : (c == 2u ? d : cx);
  const uint32_t r = c == 1u ? d : (c == 3u ? s : ct);
  return oct_mul(l, r);
}

// complete addition of two extended points, 9 M (ge_add): the 2d multiply is a layer of its own
__device__ __forceinline__ uint32_t oct_add_pt(uint32_t p, uint32_t q) {
  const uint32_t lane = threadIdx.x & 31u, k = lane & 7u, c = lane >> 3;
  const unsigned long long px = __shfl_sync(OCT_FULL, p, k), py = __shfl_sync(OCT_FULL, p, 8u | k), qx = __shfl_sync(OCT_FULL, q, k), qy = __shfl_sync(OCT_FULL, q, 8u | k);
  const unsigned long long v1[4] = {py + oct_c4p() - px, py + px, qy + oct_c4p() - qx, qy + qx};
  uint32_t n1[4]; oct_norm64_x4(v1, n1);
  const uint32_t pm = n1[0], pp = n1[1], qm = n1[2], qp = n1[3];
  // groups: (Y1 - X1)(Y2 - X2), (Y1 + X1)(Y2 + X2), Z1 Z2, T1 T2
  const uint32_t l1 = c == 0u ? pm : (c == 1u ? pp : p), r1 = c == 0u ? qm : (c == 1u ? qp : q);
  const uint32_t m1 = oct_mul(l1, r1);
  const uint32_t m2 = oct_mul(m1, FE_D2[k]);                       // group 3: 2d T1 T2 (the other groups' products are not used)
  const unsigned long long a = __shfl_sync(OCT_FULL, m1, k), b = __shfl_sync(OCT_FULL, m1, 8u | k), zz = __shfl_sync(OCT_FULL, m1, 16u | k), cc = __shfl_sync(OCT_FULL, m2, 24u | k);
  // e = b - a ; f = 2 zz - c ; g = 2 zz + c ; h = b + a
  const unsigned long long v2[4] = {b + oct_c4p() - a, 2ull * zz + oct_c4p() - cc, 2ull * zz + cc, b + a};
  uint32_t n2[4]; oct_norm64_x4(v2, n2);
  const uint32_t e = n2[0], f = n2[1], g = n2[2], h = n2[3];
  // X3 = e f, Y3 = g h, Z3 = f g, T3 = e h
  const uint32_t l = (c == 0u || c == 3u) ? e : (c == 1u ? g : f);
  const uint32_t r = c == 0u ? f : (c == 2u ? g : h);
  return oct_mul(l, r);
}

// the point as an ordinary `ge` (every lane gets the whole point)
__device__ __forceinline__ ge oct_to_ge(uint32_t p) {
  ge r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_sync(OCT_FULL, p, i); r.Y.v[i] = __shfl_sync(OCT_FULL, p, 8 + i);
    r.Z.v[i] = __shfl_sync(OCT_FULL, p, 16 + i); r.T.v[i] = __shfl_sync(OCT_FULL, p, 24 + i);
  }
  return r;
}

}  // namespace xhe
