// kernels_point.cu -- per-point kernels: ristretto255 decode / encode / one-way map, ciphertext add/sub, selftest and
// the integer-pipe microbenchmarks.  One thread per point; these kernels are bound by the integer-multiply pipe
// (one invsqrt = 254 S + 11 M ~ 12k limb products per point against 32-128 B of traffic).
#include <stdlib.h>
#include "xhe_internal.cuh"
#include "quad.cuh"
using namespace xhe;

#define XHE_PT_THREADS 128

// K2: batched CompressedRistretto::decompress.  Replaces src/compressed.rs:28-34 et al.
__global__ void __launch_bounds__(XHE_PT_THREADS) k_decompress(const uint8_t* __restrict__ enc, size_t n, uint32_t* __restrict__ aff,
                                                               uint32_t* __restrict__ niels, uint8_t* __restrict__ ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_aff p;
  bool good = decode_words(p, enc + 32 * i);
  if (!good) p = ge_aff_identity();   // keep downstream arithmetic well-defined; the flag carries the verdict
  if (aff) { st_fe(aff + 16 * i, p.x); st_fe(aff + 16 * i + 8, p.y); }
  if (niels) st_niels(niels + 24 * i, niels_from_affine(p));
  ok[i] = good ? 1 : 0;
}

// K3: batched RistrettoPoint::compress from extended coordinates.  Replaces src/compressed.rs:17-21 et al.
__global__ void __launch_bounds__(XHE_PT_THREADS) k_compress_ext(const uint32_t* __restrict__ ext, size_t n, uint8_t* __restrict__ enc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge p; ld_ge(p, ext + 32 * i);
  encode_words(enc + 32 * i, p);
}
__global__ void __launch_bounds__(XHE_PT_THREADS) k_compress_aff(const uint32_t* __restrict__ aff, size_t n, uint8_t* __restrict__ enc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_aff a; ld_fe(a.x, aff + 16 * i); ld_fe(a.y, aff + 16 * i + 8);
  encode_words(enc + 32 * i, ge_from_affine(a));
}
// from canonical little-endian bytes x||y (host API convenience)
__global__ void __launch_bounds__(XHE_PT_THREADS) k_compress_xy_bytes(const uint8_t* __restrict__ xy, size_t n, uint8_t* __restrict__ enc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[32]; ge_aff a;
  ld_bytes32(b, xy + 64 * i); a.x = fe_frombytes(b); a.x.v[7] |= 0;  // fe_frombytes clears bit 255 (canonical inputs have it clear)
  ld_bytes32(b, xy + 64 * i + 32); a.y = fe_frombytes(b);
  encode_words(enc + 32 * i, ge_from_affine(a));
}
__global__ void __launch_bounds__(XHE_PT_THREADS) k_affine_to_bytes(const uint32_t* __restrict__ aff, size_t n, uint8_t* __restrict__ xy) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe x, y; ld_fe(x, aff + 16 * i); ld_fe(y, aff + 16 * i + 8);
  x = fe_freeze(x); y = fe_freeze(y);
  st_fe(reinterpret_cast<uint32_t*>(xy + 64 * i), x); st_fe(reinterpret_cast<uint32_t*>(xy + 64 * i + 32), y);
}

// K11 (part): ristretto255 one-way map of 64 uniform bytes -> encoding (+ optional affine Niels for generator tables)
__global__ void __launch_bounds__(XHE_PT_THREADS) k_from_uniform(const uint8_t* __restrict__ u64, size_t n, uint8_t* __restrict__ enc, uint32_t* __restrict__ niels) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[64];
  ld_bytes32(b, u64 + 64 * i); ld_bytes32(b + 32, u64 + 64 * i + 32);
  ge p = ristretto_from_uniform(b);
  if (enc) encode_words(enc + 32 * i, p);
  if (niels) {
    fe zi = fe_invert(p.Z);
    ge_aff a; a.x = fe_mul(p.X, zi); a.y = fe_mul(p.Y, zi);
    st_niels(niels + 24 * i, niels_from_affine(a));
  }
}

// K4 (compressed I/O): out = bal +/- delta on 32-byte encodings, one thread per POINT (2 per ciphertext).
// Replaces src/elgamal.rs:322-342 + (de)compression in src/tx/verify.rs:561-609.  3 invsqrt per thread.
__global__ void __launch_bounds__(XHE_PT_THREADS) k_ct_update(const uint8_t* __restrict__ bal, const uint8_t* __restrict__ delta, const uint8_t* __restrict__ sub,
                                                              size_t n_points, uint8_t* __restrict__ out, uint8_t* __restrict__ ok_pt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  ge_aff a, b;
  bool good = decode_words(a, bal + 32 * i);
  good = decode_words(b, delta + 32 * i) && good;
  ge_niels nb = niels_cneg(niels_from_affine(b), sub[i >> 1] != 0);
  ge r = ge_madd(ge_from_affine(a), nb);
  if (good) encode_words(out + 32 * i, r);
  else { reinterpret_cast<uint4*>(out + 32 * i)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(out + 32 * i)[1] = make_uint4(0, 0, 0, 0); }
  ok_pt[i] = good ? 1 : 0;
}
// a ciphertext is valid only if both halves decoded; an invalid one is returned as 64 zero bytes (oracle: xo_ct_update)
__global__ void k_and_pairs(const uint8_t* __restrict__ ok_pt, size_t n, uint8_t* __restrict__ ok, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t good = ok_pt[2 * i] & ok_pt[2 * i + 1];
  ok[i] = good;
  if (!good) { uint4 z = make_uint4(0, 0, 0, 0); uint4* o = reinterpret_cast<uint4*>(out + 64 * i); o[0] = z; o[1] = z; o[2] = z; o[3] = z; }
}

// K4 (resident): balances live on the device as extended points in coordinate-planar layout
// [X | Y | Z | T][2n points][8 limbs]; deltas as affine Niels planar [ypx | ymx | t2d][2n][8].  One thread per point:
// 7 loads + 4 stores of 32 B, fully coalesced (adjacent lanes touch adjacent 32-byte sectors).  HBM-bound:
// 128 B read + 128 B written + 96 B read = 352 B per point = 704 B per account.
template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_ct_update_resident(uint32_t* __restrict__ bal, const uint32_t* __restrict__ delta, const uint8_t* __restrict__ sub, size_t n_points) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  ge p; ge_niels q;
  ld_fe_rw(p.X, bal + 8 * i); ld_fe_rw(p.Y, bal + 8 * (n_points + i)); ld_fe_rw(p.Z, bal + 8 * (2 * n_points + i)); ld_fe_rw(p.T, bal + 8 * (3 * n_points + i));
  ld_fe(q.ypx, delta + 8 * i); ld_fe(q.ymx, delta + 8 * (n_points + i)); ld_fe(q.t2d, delta + 8 * (2 * n_points + i));
  ge r = ge_madd(p, niels_cneg(q, sub[i >> 1] != 0));
  st_fe(bal + 8 * i, r.X); st_fe(bal + 8 * (n_points + i), r.Y); st_fe(bal + 8 * (2 * n_points + i), r.Z); st_fe(bal + 8 * (3 * n_points + i), r.T);
}

// arithmetic selftest: same op codes as tests/hostemu/fe_emu.cpp::emu_fe_op
__global__ void k_selftest_fe(int op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, size_t n, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe x, y, r; ld_fe(x, a + 8 * i); ld_fe(y, b + 8 * i);
  switch (op) {
    case 0: r = fe_add(x, y); break; case 1: r = fe_sub(x, y); break; case 2: r = fe_mul(x, y); break; case 3: r = fe_sq(x); break;
    case 4: r = fe_freeze(x); break; case 5: r = fe_invert(x); break; case 6: r = fe_pow22523(x); break; case 7: r = fe_neg(x); break;
    case 8: { sc s, t; for (int k = 0; k < 8; k++) { s.v[k] = x.v[k]; t.v[k] = y.v[k]; } sc m = sc_mul(sc_reduce256(s), sc_reduce256(t)); for (int k = 0; k < 8; k++) r.v[k] = m.v[k]; break; }
    case 9: { sc s, t; for (int k = 0; k < 8; k++) { s.v[k] = x.v[k]; t.v[k] = y.v[k]; } sc m = sc_reduce512(s, t); for (int k = 0; k < 8; k++) r.v[k] = m.v[k]; break; }
    case 10: { sc s; for (int k = 0; k < 8; k++) s.v[k] = x.v[k]; sc m = sc_from_mont(sc_mont_invert(sc_to_mont(sc_reduce256(s)))); for (int k = 0; k < 8; k++) r.v[k] = m.v[k]; break; }
    default: r = fe_zero();
  }
  st_fe(out + 8 * i, r);
}

// integer-pipe microbenchmarks: 8 independent dependency chains per thread, ITER x 8 x 8 instructions of one kind.
// Every multiplicand changes from one instruction to the next use: ptxas hoists loop-invariant products and replaces a
// `mad.wide` whose product is invariant by IADD3 pairs (an earlier version of WHICH == 2 measured exactly that, 18.4 T/s
// of additions).  SASS of each variant is checked in tools/imad_probe.cu / DESIGN.md 4.1.
//   0: IMAD (32-bit low product)        1: IMAD.HI.U32
//   2: IMAD.WIDE.U32 Rd64, Ra, Rb, RZ  (two vector multiplicands, both result words live: the plain 32x32->64 product)
//   3: IMAD.WIDE.U32 / IMAD.WIDE.U32.X carry chains (what the radix-2^32 field multiply issues)
template <int WHICH>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
  uint32_t x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  unsigned long long w[8];
#pragma unroll
  for (int k = 0; k < 8; k++) w[k] = ((unsigned long long)(a * 2654435761u + k) << 32) | (a + k);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (WHICH == 0) {
        asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                     "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                     : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(b), "r"(a));
      } else if (WHICH == 1) {
        asm volatile("mad.hi.u32 %0, %0, %8, %9; mad.hi.u32 %1, %1, %8, %9; mad.hi.u32 %2, %2, %8, %9; mad.hi.u32 %3, %3, %8, %9;"
                     "mad.hi.u32 %4, %4, %8, %9; mad.hi.u32 %5, %5, %8, %9; mad.hi.u32 %6, %6, %8, %9; mad.hi.u32 %7, %7, %8, %9;"
                     : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(b), "r"(a));
      } else if (WHICH == 3) {
        // carry-chained wide mads (what the radix-2^32 field multiply issues): 4 chains of 2 slots
        asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.u32 %3, %8, %9, %3;"
                     "mad.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                     : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(b), "r"(a));
      } else {
        // w_k <- lo(w_k) * hi(w_{k+1}): 8 independent 64-bit products per round, every result word feeds a later multiply
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[k]) : "r"((uint32_t)w[k]), "r"((uint32_t)(w[(k + 1) & 7] >> 32)));
      }
    }
  }
  unsigned long long f = w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7];
  uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)f ^ (uint32_t)(f >> 32);
  if (r == 0x12345678u) out[0] = r;   // practically never: keeps the chains alive
}

// ---------------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

extern "C" int32_t xhe_decompress_dev(xhe_ctx* ctx, const void* d_enc, size_t n, void* d_affine, void* d_niels, void* d_ok) {
  if (!ctx || (n && (!d_enc || !d_ok))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  static const size_t dec_smem = []() { const char* e = getenv("XHE_DEC_SMEM"); size_t v = e ? (size_t)atol(e) : 0;     // residency limiter, see k_msm_accum_tiles
    if (v > 48 * 1024) cudaFuncSetAttribute(k_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); return v; }();
  k_decompress<<<blocks_for(n, XHE_PT_THREADS), XHE_PT_THREADS, dec_smem, ctx->stream>>>((const uint8_t*)d_enc, n, (uint32_t*)d_affine, (uint32_t*)d_niels, (uint8_t*)d_ok);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
extern "C" int32_t xhe_compress_dev(xhe_ctx* ctx, const void* d_ext, size_t n, void* d_enc) {
  if (!ctx || (n && (!d_ext || !d_enc))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  k_compress_ext<<<blocks_for(n, XHE_PT_THREADS), XHE_PT_THREADS, 0, ctx->stream>>>((const uint32_t*)d_ext, n, (uint8_t*)d_enc);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
extern "C" int32_t xhe_from_uniform_dev(xhe_ctx* ctx, const void* d_u, size_t n, void* d_enc) {
  if (!ctx || (n && (!d_u || !d_enc))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  k_from_uniform<<<blocks_for(n, XHE_PT_THREADS), XHE_PT_THREADS, 0, ctx->stream>>>((const uint8_t*)d_u, n, (uint8_t*)d_enc, nullptr);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_from_uniform_niels_dev(xhe_ctx* ctx, const void* d_u, size_t n, void* d_niels) {
  k_from_uniform<<<blocks_for(n, XHE_PT_THREADS), XHE_PT_THREADS, 0, ctx->stream>>>((const uint8_t*)d_u, n, nullptr, (uint32_t*)d_niels);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_compress_xy_bytes_dev(xhe_ctx* ctx, const void* d_xy, size_t n, void* d_enc) {
  k_compress_xy_bytes<<<blocks_for(n, XHE_PT_THREADS), XHE_PT_THREADS, 0, ctx->stream>>>((const uint8_t*)d_xy, n, (uint8_t*)d_enc);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_affine_to_bytes_dev(xhe_ctx* ctx, const void* d_aff, size_t n, void* d_xy) {
  k_affine_to_bytes<<<blocks_for(n, XHE_PT_THREADS), XHE_PT_THREADS, 0, ctx->stream>>>((const uint32_t*)d_aff, n, (uint8_t*)d_xy);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
extern "C" int32_t xhe_ct_update_dev(xhe_ctx* ctx, const void* d_bal, const void* d_delta, const void* d_sub, size_t n, void* d_out, void* d_ok) {
  if (!ctx || (n && (!d_bal || !d_delta || !d_sub || !d_out || !d_ok))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  // per-point flags are staged in the tail of the caller's d_ok?  no: use ctx scratch
  if (ctx->scratch_bytes < 2 * n) { if (ctx->d_scratch) cudaFree(ctx->d_scratch); XHE_CUDA_OK(ctx, cudaMalloc(&ctx->d_scratch, 2 * n + 4096)); ctx->scratch_bytes = 2 * n + 4096; }
  k_ct_update<<<blocks_for(2 * n, XHE_PT_THREADS), XHE_PT_THREADS, 0, ctx->stream>>>((const uint8_t*)d_bal, (const uint8_t*)d_delta, (const uint8_t*)d_sub, 2 * n, (uint8_t*)d_out, (uint8_t*)ctx->d_scratch);
  XHE_LAUNCHED(ctx);
  k_and_pairs<<<blocks_for(n, 256), 256, 0, ctx->stream>>>((const uint8_t*)ctx->d_scratch, n, (uint8_t*)d_ok, (uint8_t*)d_out);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
extern "C" int32_t xhe_ct_update_resident_dev(xhe_ctx* ctx, void* d_bal_ext, const void* d_delta_niels, const void* d_sub, size_t n) {
  if (!ctx || (n && (!d_bal_ext || !d_delta_niels || !d_sub))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  static const int tpb = getenv("XHE_CTRES_TPB") ? atoi(getenv("XHE_CTRES_TPB")) : 128;     // experiment knob: threads per block (x10 + min blocks/SM for the capped variants).  Measured on B200, 1 M accounts, L2 flushed (tools/ct_resident_bench.py): 256 -> 0.161 ms, 128 -> 0.156, 64 -> 0.155, 128 capped at 80 registers -> 0.163, at 64 registers -> 0.199
#define XHE_CTRES(T, M) k_ct_update_resident<T, M><<<blocks_for(2 * n, T), T, 0, ctx->stream>>>((uint32_t*)d_bal_ext, (const uint32_t*)d_delta_niels, (const uint8_t*)d_sub, 2 * n)
  if (tpb == 64) XHE_CTRES(64, 1); else if (tpb == 128) XHE_CTRES(128, 1); else if (tpb == 1286) XHE_CTRES(128, 6); else if (tpb == 1288) XHE_CTRES(128, 8); else XHE_CTRES(256, 1);
#undef XHE_CTRES
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
extern "C" int32_t xhe_selftest_fe(xhe_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
  if (!ctx || !a || !b || !out) return XHE_E_ARG;
  uint32_t *da, *db, *dout;
  XHE_CUDA_OK(ctx, cudaMalloc(&da, 32 * n)); XHE_CUDA_OK(ctx, cudaMalloc(&db, 32 * n)); XHE_CUDA_OK(ctx, cudaMalloc(&dout, 32 * n));
  XHE_CUDA_OK(ctx, cudaMemcpy(da, a, 32 * n, cudaMemcpyHostToDevice)); XHE_CUDA_OK(ctx, cudaMemcpy(db, b, 32 * n, cudaMemcpyHostToDevice));
  k_selftest_fe<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(op, da, db, n, dout);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  XHE_CUDA_OK(ctx, cudaMemcpy(out, dout, 32 * n, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout); return XHE_OK;
}
extern "C" int32_t xhe_measure_int_peak(xhe_ctx* ctx, int which, double* rate) {
  if (!ctx || !rate || which < 0 || which > 3) return XHE_E_ARG;
  uint32_t* d; XHE_CUDA_OK(ctx, cudaMalloc(&d, 64));
  const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, ctx->stream);
    if (which == 0) k_int_peak<0><<<blocks, threads, 0, ctx->stream>>>(d, 12345u + rep, iters);
    else if (which == 1) k_int_peak<1><<<blocks, threads, 0, ctx->stream>>>(d, 12345u + rep, iters);
    else if (which == 3) k_int_peak<3><<<blocks, threads, 0, ctx->stream>>>(d, 12345u + rep, iters);
    else k_int_peak<2><<<blocks, threads, 0, ctx->stream>>>(d, 12345u + rep, iters);
    XHE_LAUNCHED(ctx);
    cudaEventRecord(e1, ctx->stream); XHE_CUDA_OK(ctx, cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  *rate = (double)blocks * threads * iters * (which == 3 ? 32.0 : 64.0) / (best * 1e-3);   // which 3: 32 wide products per iteration
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); return XHE_OK;
}

// ---- micro-op timing (design evidence): cycles per field / point operation for `warps` resident warps per SM ------
template <int OP>
__global__ void k_bench_op(uint32_t* out, const uint32_t* in, int iters, unsigned long long* cycles) {
  fe a, b; ld_fe(a, in + 8 * (threadIdx.x & 31)); ld_fe(b, in + 8 * ((threadIdx.x + 7) & 31));
  ge p; p.X = a; p.Y = b; p.Z = fe_one(); p.T = fe_mul(a, b);
  ge_niels q; q.ypx = b; q.ymx = a; q.t2d = p.T;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) a = fe_mul(a, b);
    else if (OP == 1) a = fe_sq(a);
    else if (OP == 2) p = ge_double(p);
    else if (OP == 3) p = ge_madd(p, q);
    else if (OP == 4) p = ge_add(p, p);
    else if (OP == 5) a = fe_add(a, b);
    else if (OP == 6) p = quad_double(p);
    else if (OP == 7) p = quad_add(p, p);
    else if (OP == 8) p = quad_madd(p, q);
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  uint32_t r = 0;
  for (int i = 0; i < 8; i++) r ^= a.v[i] ^ p.X.v[i] ^ p.Y.v[i] ^ p.Z.v[i] ^ p.T.v[i];
  if (r == 0x12345u) out[0] = r;
}
extern "C" int32_t xhe_bench_op(xhe_ctx* ctx, int op, int threads_per_block, int blocks, int iters, double* cycles_per_op) {
  // cycles_per_op[0]: latency seen by warp 0 (clock64); cycles_per_op[1]: whole-kernel ns per (iteration x warp per SMSP)
  uint32_t *d_in, *d_out; unsigned long long* d_c;
  XHE_CUDA_OK(ctx, cudaMalloc(&d_in, 32 * 32)); XHE_CUDA_OK(ctx, cudaMalloc(&d_out, 64)); XHE_CUDA_OK(ctx, cudaMalloc(&d_c, 8));
  uint32_t h[256]; for (int i = 0; i < 256; i++) h[i] = 0x9e3779b9u * (i + 1) + 12345u;
  for (int i = 0; i < 32; i++) h[8 * i + 7] &= 0x7fffffffu;
  cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, ctx->stream);
    switch (op) {
      case 0: k_bench_op<0><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 1: k_bench_op<1><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 2: k_bench_op<2><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 3: k_bench_op<3><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 4: k_bench_op<4><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 6: k_bench_op<6><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 7: k_bench_op<7><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      case 8: k_bench_op<8><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
      default: k_bench_op<5><<<blocks, threads_per_block, 0, ctx->stream>>>(d_out, d_in, iters, d_c); break;
    }
    cudaEventRecord(e1, ctx->stream);
    XHE_LAUNCHED(ctx);
    XHE_CUDA_OK(ctx, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
  }
  unsigned long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  cycles_per_op[0] = (double)c / iters;
  double warp_ops_per_smsp = (double)blocks * (threads_per_block / 32) * iters / (ctx->sm_count * 4.0);
  cycles_per_op[1] = ms * 1e6 / warp_ops_per_smsp;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_c); return XHE_OK;
}

// CUDA loads kernels lazily (CUDA_MODULE_LOADING=LAZY is the default since 12.2), and loading one may need every running kernel
// to finish first: with the polling chain kernel of msm.cu in flight, the FIRST launch of any other kernel would wait for a
// kernel that is waiting for it.  Every kernel of this file is therefore loaded when the first context is created.
void xhe_preload_point() {
  const void* ks[] = {(const void*)k_decompress, (const void*)k_compress_ext, (const void*)k_compress_aff, (const void*)k_compress_xy_bytes, (const void*)k_affine_to_bytes, (const void*)k_from_uniform, (const void*)k_ct_update, (const void*)k_and_pairs, (const void*)k_ct_update_resident<128, 1>, (const void*)k_selftest_fe};      // (the microbenchmark kernels never run beside a batch)
  cudaFuncAttributes a;
  for (const void* k : ks) cudaFuncGetAttributes(&a, k);
}
