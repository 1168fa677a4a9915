// fe25519.cuh -- GF(2^255-19) for sm_100a in eight 32-bit limbs (radix 2^32, values kept in [0, 2^256)).
//
// Replaces curve25519-dalek's `FieldElement` (un-vendored dependency of the reference; used behind
// src/compressed.rs:28-34, src/elgamal.rs:283-370, src/proofs.rs:50).  Not a port: dalek uses 5x51-bit (u64) or
// 10x25.5-bit (u32) limbs; here products are 32x32->64 `IMAD.WIDE.U32` issued as even/odd column chains
// (mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32.X with the carry in a predicate), so a full
// 8x8 schoolbook product is 64 IMAD.WIDE + 16 IADD3 and the 2^256 = 38 fold is 8 more.  Work unit:
// M = 72 limb products (LP), S = 44 LP.
//
// Every carry chain lives inside ONE asm statement, so neither NVVM nor ptxas can separate a chain from its flag.
// When compiled for the host (tests/hostemu only -- never shipped) the same chains run as portable C, so the
// library logic above the chains is unit-tested on the CPU against Python big integers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XHE_HD __host__ __device__ __forceinline__
#else
#define XHE_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define XHE_ASM 1
#else
#define XHE_ASM 0
#endif

namespace xhe {

struct fe { uint32_t v[8]; };

// ---------------------------------------------------------------------------------------------------------------
// chain primitives: acc[2k], acc[2k+1] (aligned 64-bit slots) += a_k * b for k < N, carry chained across slots;
// the carry out of the last slot is ADDED to `co` (or dropped in the _nc forms, where the caller proves it is 0).
// ---------------------------------------------------------------------------------------------------------------
#if !XHE_ASM
inline uint32_t emu_madw(uint32_t* acc, const uint32_t* a, int n, uint32_t b) {
  uint64_t c = 0;
  for (int k = 0; k < n; k++) {
    uint64_t p = (uint64_t)a[k] * b;
    uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    acc[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  return (uint32_t)c;
}
#endif

XHE_HD void mad1w(uint32_t* acc, uint32_t& co, uint32_t a0, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(co)
      : "r"(a0), "r"(b));
#else
  const uint32_t a[1] = {a0};
  uint32_t c = emu_madw(acc, a, 1, b);
  co += c;
#endif
}

XHE_HD void mad1w_nc(uint32_t* acc, uint32_t a0, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
      "madc.hi.u32 %1, %2, %3, %1;"
      : "+r"(acc[0]), "+r"(acc[1])
      : "r"(a0), "r"(b));
#else
  const uint32_t a[1] = {a0};
  uint32_t c = emu_madw(acc, a, 1, b);
  (void)c;
#endif
}

XHE_HD void mad2w(uint32_t* acc, uint32_t& co, uint32_t a0, uint32_t a1, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(co)
      : "r"(a0), "r"(a1), "r"(b));
#else
  const uint32_t a[2] = {a0, a1};
  uint32_t c = emu_madw(acc, a, 2, b);
  co += c;
#endif
}

XHE_HD void mad2w_nc(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
      "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
      "madc.hi.u32 %3, %5, %6, %3;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
      : "r"(a0), "r"(a1), "r"(b));
#else
  const uint32_t a[2] = {a0, a1};
  uint32_t c = emu_madw(acc, a, 2, b);
  (void)c;
#endif
}

XHE_HD void mad3w(uint32_t* acc, uint32_t& co, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
      "addc.u32 %6, %6, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(co)
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
#else
  const uint32_t a[3] = {a0, a1, a2};
  uint32_t c = emu_madw(acc, a, 3, b);
  co += c;
#endif
}

XHE_HD void mad3w_nc(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %6, %9, %0;\n\t"
      "madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
      "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
      "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
      "madc.lo.cc.u32 %4, %8, %9, %4;\n\t"
      "madc.hi.u32 %5, %8, %9, %5;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
#else
  const uint32_t a[3] = {a0, a1, a2};
  uint32_t c = emu_madw(acc, a, 3, b);
  (void)c;
#endif
}

XHE_HD void mad4w(uint32_t* acc, uint32_t& co, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(co)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  uint32_t c = emu_madw(acc, a, 4, b);
  co += c;
#endif
}

XHE_HD void mad4w_nc(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  uint32_t c = emu_madw(acc, a, 4, b);
  (void)c;
#endif
}

// out[0..7] = {a0,a1,a2,a3} * b at slots (0,1)..(6,7) (no accumulate)
XHE_HD void mul4w(uint32_t* out, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if XHE_ASM
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;"
      : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5]), "=r"(out[6]), "=r"(out[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  for (int k = 0; k < 4; k++) { uint64_t p = (uint64_t)a[k] * b; out[2 * k] = (uint32_t)p; out[2 * k + 1] = (uint32_t)(p >> 32); }
#endif
}

// r = x + y + cin (8 limbs), returns carry-out.  cin in {0,1}.
XHE_HD uint32_t add8c(uint32_t* r, const uint32_t* x, const uint32_t* y, uint32_t cin) {
#if XHE_ASM
  uint32_t c;
  asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"   // CC <- cin
      "addc.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(cin));
  return c;
#else
  uint64_t c = cin;
  for (int i = 0; i < 8; i++) { c += (uint64_t)x[i] + y[i]; r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
#endif
}

XHE_HD uint32_t add8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
#if XHE_ASM
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
  return c;
#else
  return add8c(r, x, y, 0);
#endif
}

// r = x - y, returns borrow (1 if x < y)
XHE_HD uint32_t sub8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
#if XHE_ASM
  uint32_t b;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(b)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
  return b & 1u;  // 0 - 0 - borrow = 0xffffffff when the chain borrowed
#else
  uint64_t b = 0;
  for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)x[i] - y[i] - b; r[i] = (uint32_t)t; b = (t >> 32) & 1; }
  return (uint32_t)b;
#endif
}

// r[0..7] += u (one word) with full propagation, returns carry-out
XHE_HD uint32_t addw8(uint32_t* r, uint32_t u) {
#if XHE_ASM
  uint32_t c;
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(c)
      : "r"(u));
  return c;
#else
  uint64_t c = u;
  for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
#endif
}

// r[0..7] -= u with full propagation, returns borrow
XHE_HD uint32_t subw8(uint32_t* r, uint32_t u) {
#if XHE_ASM
  uint32_t b;
  asm("sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\t"
      "subc.cc.u32 %4, %4, 0;\n\t"
      "subc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\t"
      "subc.cc.u32 %7, %7, 0;\n\t"
      "subc.u32 %8, 0, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(b)
      : "r"(u));
  return b & 1u;
#else
  uint64_t b = u;
  for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)r[i] - b; r[i] = (uint32_t)t; b = (t >> 32) & 1; }
  return (uint32_t)b;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// field operations.  Invariant: every fe is a 256-bit integer in [0, 2^256) congruent to the element mod p.
// ---------------------------------------------------------------------------------------------------------------

XHE_HD fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
XHE_HD fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }

XHE_HD fe fe_add(const fe& a, const fe& b) {
  fe r;
  uint32_t c = add8(r.v, a.v, b.v);
  c = addw8(r.v, c * 38u);  // 2^256 = 38 (mod p)
  r.v[0] += c * 38u;        // after a second wrap r is tiny, so this cannot carry
  return r;
}

XHE_HD fe fe_sub(const fe& a, const fe& b) {
  fe r;
  uint32_t bw = sub8(r.v, a.v, b.v);
  bw = subw8(r.v, bw * 38u);
  r.v[0] -= bw * 38u;       // after a second wrap r is within 38 of 2^256, so this cannot borrow
  return r;
}

XHE_HD fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }
XHE_HD fe fe_dbl(const fe& a) { return fe_add(a, a); }

// t[0..15] = ev[0..15] + (od[0..14] << 32)
XHE_HD void merge16(uint32_t* t, const uint32_t* ev, const uint32_t* od) {
  t[0] = ev[0];
  uint32_t c = add8(t + 1, ev + 1, od);  // limbs 1..8
  uint32_t e2[8], o2[8], hi[8];
#pragma unroll
  for (int i = 0; i < 7; i++) { e2[i] = ev[9 + i]; o2[i] = od[8 + i]; }
  e2[7] = 0; o2[7] = 0;
  add8c(hi, e2, o2, c);
#pragma unroll
  for (int i = 0; i < 7; i++) t[9 + i] = hi[i];
}

// 16-limb product t -> fe: lo + 38*hi (nine limbs), then fold the ninth limb twice
XHE_HD fe fe_reduce512(const uint32_t* t) {
  uint32_t ev[9], od[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ev[i] = t[i];
  ev[8] = 0;
  mad4w(ev, ev[8], t[8], t[10], t[12], t[14], 38u);
  mul4w(od, t[9], t[11], t[13], t[15], 38u);
  fe r;
  r.v[0] = ev[0];
  add8(s, ev + 1, od);  // limbs 1..8; limb 8 = ev[8] + od[7] + carry <= 2^6, no overflow
#pragma unroll
  for (int i = 0; i < 7; i++) r.v[1 + i] = s[i];
  uint32_t c = addw8(r.v, s[7] * 38u);
  r.v[0] += c * 38u;
  return r;
}

// t[0..15] = a[0..7] * b[0..7] (full 512-bit product)
XHE_HD void mul512(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  // product a[j]*b[i] lands at limb i+j: even positions accumulate in ev[], odd in od[] (od[k] is limb k+1), so
  // every product is an aligned 64-bit slot and ptxas emits one IMAD.WIDE.U32(.X) per limb product.
  uint32_t ev[16], od[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    mad4w(ev + i, ev[i + 8], a[0], a[2], a[4], a[6], b[i]);
    mad4w(od + i, od[i + 8], a[1], a[3], a[5], a[7], b[i]);
    if (i + 10 < 16) mad4w(ev + i + 2, ev[i + 10], a[1], a[3], a[5], a[7], b[i + 1]);
    else mad4w_nc(ev + i + 2, a[1], a[3], a[5], a[7], b[i + 1]);
    mad4w(od + i, od[i + 8], a[0], a[2], a[4], a[6], b[i + 1]);
  }
  merge16(t, ev, od);
}

XHE_HD fe fe_mul(const fe& a, const fe& b) {
  uint32_t t[16];
  mul512(t, a.v, b.v);
  return fe_reduce512(t);
}

// t[0..15] = x[0..7]^2
XHE_HD void sq512(uint32_t* t, const uint32_t* x) {
  // 28 cross products x[i]*x[j] (i<j) in even/odd chains, doubled, plus the 8 squares: 36 IMAD.WIDE.
  uint32_t ev[16], od[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
  // limb position p = i+j; p even -> slot ev[p], p odd -> slot od[p-1]
  mad4w(od + 0, od[8], x[1], x[3], x[5], x[7], x[0]);   // p = 1,3,5,7
  mad3w(ev + 2, ev[8], x[2], x[4], x[6], x[0]);         // p = 2,4,6
  mad3w(od + 2, od[8], x[2], x[4], x[6], x[1]);         // p = 3,5,7
  mad3w(ev + 4, ev[10], x[3], x[5], x[7], x[1]);        // p = 4,6,8
  mad3w(od + 4, od[10], x[3], x[5], x[7], x[2]);        // p = 5,7,9
  mad2w(ev + 6, ev[10], x[4], x[6], x[2]);              // p = 6,8
  mad2w(od + 6, od[10], x[4], x[6], x[3]);              // p = 7,9
  mad2w(ev + 8, ev[12], x[5], x[7], x[3]);              // p = 8,10
  mad2w(od + 8, od[12], x[5], x[7], x[4]);              // p = 9,11
  mad1w(ev + 10, ev[12], x[6], x[4]);                   // p = 10
  mad1w(od + 10, od[12], x[6], x[5]);                   // p = 11
  mad1w(ev + 12, ev[14], x[7], x[5]);                   // p = 12
  mad1w_nc(od + 12, x[7], x[6]);                        // p = 13
  merge16(t, ev, od);
  // t = 2t (t < 2^511)
  {
    uint32_t lo[8], hi[8];
    uint32_t c = add8(lo, t, t);
    add8c(hi, t + 8, t + 8, c);
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = lo[i]; t[8 + i] = hi[i]; }
  }
  // + squares x[i]^2 at limb 2i
#if XHE_ASM
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
      "madc.hi.u32 %15, %23, %23, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]),
        "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
#else
  {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t p = (uint64_t)x[i] * x[i];
      uint64_t lo = (uint64_t)t[2 * i] + (uint32_t)p + c; t[2 * i] = (uint32_t)lo;
      uint64_t hi = (uint64_t)t[2 * i + 1] + (uint32_t)(p >> 32) + (lo >> 32); t[2 * i + 1] = (uint32_t)hi; c = hi >> 32;
    }
  }
#endif
}

XHE_HD fe fe_sq(const fe& a) {
  uint32_t t[16];
  sq512(t, a.v);
  return fe_reduce512(t);
}

XHE_HD fe fe_sqn(fe a, int n) {
  for (int i = 0; i < n; i++) a = fe_sq(a);
  return a;
}

// canonical representative in [0, p)
XHE_HD fe fe_freeze(const fe& a) {
  fe t = a;
  uint32_t b = t.v[7] >> 31;
  t.v[7] &= 0x7fffffffu;
  addw8(t.v, 19u * b);  // t < 2^255 + 19
  fe u = t;
  addw8(u.v, 19u);
  uint32_t q = u.v[7] >> 31;  // 1 iff t >= p
  addw8(t.v, 19u * q);
  t.v[7] &= 0x7fffffffu;
  return t;
}

XHE_HD bool fe_iszero(const fe& a) {
  fe t = fe_freeze(a);
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r |= t.v[i];
  return r == 0;
}
XHE_HD bool fe_isneg(const fe& a) { return fe_freeze(a).v[0] & 1u; }
XHE_HD bool fe_eq(const fe& a, const fe& b) { return fe_iszero(fe_sub(a, b)); }
XHE_HD fe fe_select(const fe& a, const fe& b, bool pick_b) {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = pick_b ? b.v[i] : a.v[i];
  return r;
}
XHE_HD fe fe_cneg(const fe& a, bool neg) { return fe_select(a, fe_neg(a), neg); }
XHE_HD fe fe_abs(const fe& a) { return fe_cneg(a, fe_isneg(a)); }

// bytes <-> fe (little-endian; frombytes ignores bit 255 like dalek's FieldElement::from_bytes)
XHE_HD fe fe_frombytes(const uint8_t* s) {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
  r.v[7] &= 0x7fffffffu;
  return r;
}
XHE_HD void fe_tobytes(uint8_t* s, const fe& a) {
  fe t = fe_freeze(a);
#pragma unroll
  for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)t.v[i]; s[4 * i + 1] = (uint8_t)(t.v[i] >> 8); s[4 * i + 2] = (uint8_t)(t.v[i] >> 16); s[4 * i + 3] = (uint8_t)(t.v[i] >> 24); }
}

// a^(2^250-1) and a^11: shared prefix of the inversion and (p-5)/8 chains (254 S + 11 M in total)
XHE_HD void fe_pow_2_250_1(fe& t250, fe& a11, const fe& a) {
  fe t0 = fe_sq(a);
  fe t1 = fe_mul(a, fe_sqn(t0, 2));     // a^9
  t0 = fe_mul(t0, t1);                  // a^11
  a11 = t0;
  t1 = fe_mul(t1, fe_sq(t0));           // 2^5-1
  t1 = fe_mul(fe_sqn(t1, 5), t1);       // 2^10-1
  fe t2 = fe_mul(fe_sqn(t1, 10), t1);   // 2^20-1
  t2 = fe_mul(fe_sqn(t2, 20), t2);      // 2^40-1
  t1 = fe_mul(fe_sqn(t2, 10), t1);      // 2^50-1
  t2 = fe_mul(fe_sqn(t1, 50), t1);      // 2^100-1
  t2 = fe_mul(fe_sqn(t2, 100), t2);     // 2^200-1
  t250 = fe_mul(fe_sqn(t2, 50), t1);    // 2^250-1
}
XHE_HD fe fe_invert(const fe& a) { fe t, a11; fe_pow_2_250_1(t, a11, a); return fe_mul(fe_sqn(t, 5), a11); }
XHE_HD fe fe_pow22523(const fe& a) { fe t, a11; fe_pow_2_250_1(t, a11, a); return fe_mul(fe_sqn(t, 2), a); }

}  // namespace xhe
