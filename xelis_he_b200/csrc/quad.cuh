// quad.cuh -- quad-cooperative point arithmetic for LATENCY-bound chains (Horner tail of the MSM, per-signature
// double-scalar multiplication).  A single thread needs 3.2-4.4 k cycles per point operation (measured, tools/op_bench.py)
// because a field multiply is a ~490-cycle dependent chain; the twisted-Edwards formulas have two layers of four independent
// field products, so the four lanes of an aligned quad each take one product per layer and exchange results with warp
// shuffles.  Every lane of the quad holds the same point before and after the call.  All 32 lanes of the warp must call
// these functions together (full-mask shuffles); lanes whose result is not needed simply carry a copy.
#pragma once
#include "ge25519.cuh"

namespace xhe {

__device__ __forceinline__ fe quad_bcast(const fe& v, int src_in_quad) {
  fe r; const int src = (threadIdx.x & 28) | src_in_quad;   // lane index inside the warp of the quad's lane `src_in_quad`
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src);
  return r;
}
// branch-free per-lane operand selection (masks, so the compiler cannot turn it into divergent control flow that would
// make each lane class execute its own copy of the following multiply)
__device__ __forceinline__ fe quad_pick(const fe& a, const fe& b, const fe& c, const fe& d, int ql) {
  const uint32_t m0 = 0u - (uint32_t)(ql == 0), m1 = 0u - (uint32_t)(ql == 1), m2 = 0u - (uint32_t)(ql == 2), m3 = 0u - (uint32_t)(ql == 3);
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t v;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(v) : "r"(a.v[i]), "r"(m0), "r"(b.v[i] & m1));     // (a & m0) | (b & m1)
    r.v[i] = v | (c.v[i] & m2) | (d.v[i] & m3);
  }
  return r;
}

// doubling: layer 1 = X^2, Y^2, Z^2, (X+Y)^2 ; layer 2 = X3, Y3, Z3, T3
__device__ __forceinline__ ge quad_double(const ge& p) {
  const int ql = threadIdx.x & 3;
  fe in = quad_pick(p.X, p.Y, p.Z, fe_add(p.X, p.Y), ql);
  fe sq = fe_sq(in);
  fe xx = quad_bcast(sq, 0), yy = quad_bcast(sq, 1), zz2 = fe_dbl(quad_bcast(sq, 2)), xpy2 = quad_bcast(sq, 3);
  fe s = fe_add(yy, xx), d = fe_sub(yy, xx);
  fe cx = fe_sub(xpy2, s), ct = fe_sub(zz2, d);
  // X3 = cx*ct, Y3 = s*d, Z3 = d*ct, T3 = cx*s
  fe l = quad_pick(cx, s, d, cx, ql), r = quad_pick(ct, d, ct, s, ql);
  fe pr = fe_mul(l, r);
  ge o; o.X = quad_bcast(pr, 0); o.Y = quad_bcast(pr, 1); o.Z = quad_bcast(pr, 2); o.T = quad_bcast(pr, 3);
  return o;
}

// full addition of two extended points: layer 1 = (Y1-X1)(Y2-X2), (Y1+X1)(Y2+X2), T1*T2, Z1*Z2 ; lane 2 then scales by 2d ;
// layer 2 = E*F, G*H, F*G, E*H
__device__ __forceinline__ ge quad_add(const ge& p, const ge& q) {
  const int ql = threadIdx.x & 3;
  fe l1 = quad_pick(fe_sub(p.Y, p.X), fe_add(p.Y, p.X), p.T, p.Z, ql);
  fe r1 = quad_pick(fe_sub(q.Y, q.X), fe_add(q.Y, q.X), q.T, q.Z, ql);
  fe m1 = fe_mul(l1, r1);
  fe tt = quad_bcast(m1, 2);
  fe c = fe_mul(tt, fe_const(FE_D2));                 // computed by every lane (same latency as waiting for one lane to do it)
  fe a = quad_bcast(m1, 0), b = quad_bcast(m1, 1), d = fe_dbl(quad_bcast(m1, 3));
  fe e = fe_sub(b, a), f = fe_sub(d, c), g = fe_add(d, c), h = fe_add(b, a);
  fe l2 = quad_pick(e, g, f, e, ql), r2 = quad_pick(f, h, g, h, ql);
  fe pr = fe_mul(l2, r2);
  ge o; o.X = quad_bcast(pr, 0); o.Y = quad_bcast(pr, 1); o.Z = quad_bcast(pr, 2); o.T = quad_bcast(pr, 3);
  return o;
}

// mixed addition with an affine-Niels operand: layer 1 = (Y-X)*ymx, (Y+X)*ypx, T*t2d, (lane 3 idle) ; layer 2 as above
__device__ __forceinline__ ge quad_madd(const ge& p, const ge_niels& q) {
  const int ql = threadIdx.x & 3;
  fe l1 = quad_pick(fe_sub(p.Y, p.X), fe_add(p.Y, p.X), p.T, p.Z, ql);
  fe r1 = quad_pick(q.ymx, q.ypx, q.t2d, fe_one(), ql);
  fe m1 = fe_mul(l1, r1);
  fe a = quad_bcast(m1, 0), b = quad_bcast(m1, 1), c = quad_bcast(m1, 2), d = fe_dbl(p.Z);
  fe e = fe_sub(b, a), f = fe_sub(d, c), g = fe_add(d, c), h = fe_add(b, a);
  fe l2 = quad_pick(e, g, f, e, ql), r2 = quad_pick(f, h, g, h, ql);
  fe pr = fe_mul(l2, r2);
  ge o; o.X = quad_bcast(pr, 0); o.Y = quad_bcast(pr, 1); o.Z = quad_bcast(pr, 2); o.T = quad_bcast(pr, 3);
  return o;
}

}  // namespace xhe
