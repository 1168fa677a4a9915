// ecdlp.cu -- decoding of decrypted amounts: v in [0, 2^range_bits) with v * G == M.
//
// Replaces ECDLPInstance::decode / par_decode (reference src/elgamal.rs:67-92; the curve25519-dalek fork's `ecdlp` module
// behind them) and ElGamalSecretKey::decrypt (src/elgamal.rs:140-145: M = C - s * D).  Not a port of the fork's table files
// and CPU search: a baby-step / giant-step search laid out for the GPU,
//
//   * everything runs on 8 * M and 8 * G: a ristretto255 element is an Edwards point up to 4-torsion, 8 * (anything) is
//     torsion-free, and for torsion-free points the affine y coordinate identifies the point up to sign -- so a baby-step
//     table keyed by y serves r and -r at once, and a key costs a share of one batched inversion instead of the inverse
//     square root of a ristretto encoding;
//   * baby steps: j * G8 for j in [0, 2^l1], built on the device (a run of consecutive multiples per thread, mixed additions,
//     one inversion per run), stored in an open-addressing table of 8-byte slots (32-bit tag of y | j, sign of x);
//   * giant steps: v = i * 2^(l1+1) + r with |r| <= 2^l1, so Q_i = 8M - i * 2^(l1+1) * G8 has its y in the table; one WARP
//     per point, lane l takes i = l, l + 32, ..., a chunk of steps per batched inversion; every candidate is confirmed by
//     recomputing v * G8 (fixed-base) and comparing projectively, so a tag collision cannot produce a wrong amount.
//
// Work per decoded point at range 2^32, l1 = 22: 512 giant steps x (7 M mixed add + 3 M of the batched inversion + 1/8 of an
// inversion) ~ 40 k limb products more than a single ristretto decode; the table (2^22 entries, 64 MB of slots) is built once.
#include "xhe_internal.cuh"
#include <algorithm>
#include <vector>
using namespace xhe;

struct xhe_ecdlp {
  xhe_ctx* ctx = nullptr;
  uint32_t l1 = 0; size_t n_slots = 0;
  unsigned long long* d_slots = nullptr;     // tag32 << 32 | (j << 1 | x sign); 0 = empty
  uint32_t* d_consts = nullptr;              // niels(G8) | niels(-32 * S2) | ge(-k * S2), k = 0..31    (S2 = 2^(l1+1) * G8)
  const uint32_t* tabG = nullptr;            // the context's 8-bit fixed-base table of G (8 windows: u64 * G)
  void* d_tmp = nullptr; size_t tmp_bytes = 0;
};

extern "C" const uint32_t* xhe_internal_tabG(xhe_ctx* ctx);      // verify.cu

namespace {
#define ECDLP_RUN 16          // consecutive baby steps per thread (one inversion per run)
#define ECDLP_CHUNK 8         // giant steps per batched inversion

inline unsigned nblk(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ unsigned long long y_slot_hash(const fe& y) { return (((unsigned long long)y.v[1] << 32) | y.v[0]) * 0x9E3779B97F4A7C15ull; }
__device__ __forceinline__ uint32_t y_tag(const fe& y) { return (y.v[2] ^ (y.v[5] * 0x85EBCA6Bu)) | 1u; }

// v * G8 for a 64-bit v: fixed-base over the context's 8-bit table of G, then three doublings
__device__ __forceinline__ ge mul_g8(const uint32_t* __restrict__ tabG, unsigned long long v) {
  ge acc = ge_identity();
  for (int w = 0; w < 8; w++) {
    const uint32_t d = (uint32_t)(v >> (8 * w)) & 0xffu;
    if (d) { ge_niels q; ld_niels(q, tabG + 24 * ((size_t)w * 255 + (d - 1))); acc = ge_madd(acc, q); }
  }
  return ge_double(ge_double(ge_double(acc)));
}
__device__ __forceinline__ bool ge_equal(const ge& a, const ge& b) {      // projective equality of torsion-free points
  return fe_eq(fe_mul(a.X, b.Z), fe_mul(b.X, a.Z)) && fe_eq(fe_mul(a.Y, b.Z), fe_mul(b.Y, a.Z));
}

// constants of a table: niels(G8), niels(-32 S2), ge(-k S2) for k < 32
__global__ void k_ecdlp_consts(const uint32_t* __restrict__ tabG, uint32_t l1, uint32_t* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  const ge g8 = mul_g8(tabG, 1ull);
  { fe zi = fe_invert(g8.Z); ge_aff a; a.x = fe_mul(g8.X, zi); a.y = fe_mul(g8.Y, zi); st_niels(out, niels_from_affine(a)); }
  ge s2 = g8;
  for (uint32_t k = 0; k < l1 + 1; k++) s2 = ge_double(s2);
  const ge ns2 = ge_neg(s2);
  ge acc = ge_identity();
  for (int k = 0; k < 32; k++) { st_ge(out + 48 + 32 * k, acc); acc = ge_add(acc, ns2); }
  { fe zi = fe_invert(acc.Z); ge_aff a; a.x = fe_mul(acc.X, zi); a.y = fe_mul(acc.Y, zi); st_niels(out + 24, niels_from_affine(a)); }      // acc = -32 S2
}

// baby steps: thread t inserts j * G8 for j in [t * RUN, (t + 1) * RUN), j <= 2^l1
__global__ void __launch_bounds__(128) k_ecdlp_baby(const uint32_t* __restrict__ tabG, const uint32_t* __restrict__ consts, uint32_t l1, unsigned long long* __restrict__ slots, size_t n_slots) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, j0 = t * ECDLP_RUN, jmax = (size_t)1 << l1;
  if (j0 > jmax) return;
  ge_niels g8; ld_niels(g8, consts);
  ge p = mul_g8(tabG, (unsigned long long)j0);
  fe X[ECDLP_RUN], Y[ECDLP_RUN], Z[ECDLP_RUN], pre[ECDLP_RUN];
  fe run = fe_one();
#pragma unroll 1
  for (int k = 0; k < ECDLP_RUN; k++) { X[k] = p.X; Y[k] = p.Y; Z[k] = p.Z; pre[k] = run; run = fe_mul(run, p.Z); p = ge_madd(p, g8); }
  fe inv = fe_invert(run);                     // one inversion per run (Montgomery's trick)
#pragma unroll 1
  for (int k = ECDLP_RUN - 1; k >= 0; k--) {
    const fe zi = fe_mul(inv, pre[k]); inv = fe_mul(inv, Z[k]);
    const size_t j = j0 + k;
    if (j > jmax) continue;
    const fe y = fe_freeze(fe_mul(Y[k], zi)), x = fe_freeze(fe_mul(X[k], zi));
    const unsigned long long val = ((unsigned long long)y_tag(y) << 32) | ((unsigned long long)j << 1) | (x.v[0] & 1u);
    size_t s = (size_t)(y_slot_hash(y) >> 20) & (n_slots - 1);
    for (;;) { if (atomicCAS(slots + s, 0ull, val) == 0ull) break; s = (s + 1) & (n_slots - 1); }
  }
}

// decode: one warp per point.  pts = n x 32 words (affine x, y of the decoded input point; status 2 marks an invalid encoding).
__global__ void __launch_bounds__(128) k_ecdlp_decode(const uint32_t* __restrict__ pts_aff, const uint8_t* __restrict__ ok_in, size_t n, const uint32_t* __restrict__ tabG, const uint32_t* __restrict__ consts,
                                                     const unsigned long long* __restrict__ slots, size_t n_slots, uint32_t l1, uint32_t range_bits, long long* __restrict__ out_value, uint8_t* __restrict__ status) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const uint32_t lane = threadIdx.x & 31u;
  if (w >= n) return;
  if (!ok_in[w]) { if (lane == 0) { out_value[w] = -1; status[w] = 2; } return; }
  ge_aff a; ld_fe(a.x, pts_aff + 16 * w); ld_fe(a.y, pts_aff + 16 * w + 8);
  const ge m8 = ge_double(ge_double(ge_double(ge_from_affine(a))));                       // 8 M: torsion-free
  const unsigned long long limit = range_bits >= 63 ? ~0ull >> 1 : (1ull << range_bits);   // v < limit
  const unsigned long long stride = 1ull << (l1 + 1), n_giant = (limit >> (l1 + 1)) + 2;     // i in [0, n_giant)
  ge neg_k; ld_ge(neg_k, consts + 48 + 32 * lane);
  ge q = ge_add(m8, neg_k);                                                               // Q_lane = 8M - lane * S2
  ge_niels step; ld_niels(step, consts + 24);                                             // -32 S2
  long long found = -1;
  for (unsigned long long i0 = lane; ; i0 += 32ull * ECDLP_CHUNK) {
    // (warp-uniform loop: every lane walks the same number of chunks)
    const unsigned long long base = i0 - lane;
    if (base >= n_giant) break;
    fe Xs[ECDLP_CHUNK], Ys[ECDLP_CHUNK], Zs[ECDLP_CHUNK], pre[ECDLP_CHUNK];
    fe run = fe_one();
#pragma unroll 1
    for (int k = 0; k < ECDLP_CHUNK; k++) { Xs[k] = q.X; Ys[k] = q.Y; Zs[k] = q.Z; pre[k] = run; run = fe_mul(run, q.Z); q = ge_madd(q, step); }
    fe inv = fe_invert(run);
#pragma unroll 1
    for (int k = ECDLP_CHUNK - 1; k >= 0; k--) {
      const fe zi = fe_mul(inv, pre[k]); inv = fe_mul(inv, Zs[k]);
      const unsigned long long i = i0 + 32ull * (unsigned long long)k;
      if (i >= n_giant || found >= 0) continue;
      const fe y = fe_freeze(fe_mul(Ys[k], zi));
      const uint32_t tag = y_tag(y);
      size_t s = (size_t)(y_slot_hash(y) >> 20) & (n_slots - 1);
      for (;;) {
        const unsigned long long e = __ldg(slots + s);
        if (!e) break;
        if ((uint32_t)(e >> 32) == tag) {
          const unsigned long long j = (e & 0xffffffffull) >> 1; const uint32_t xs = (uint32_t)e & 1u;
          const fe x = fe_freeze(fe_mul(Xs[k], zi));
          const bool same_sign = (x.v[0] & 1u) == xs;                                       // Q = +j G8 or -j G8 (x of the identity / of the 2-torsion-free points: 0 only for j = 0)
          const long long r = same_sign ? (long long)j : -(long long)j;
          const long long v = (long long)(i * stride) + r;
          if (v >= 0 && (unsigned long long)v < limit && ge_equal(mul_g8(tabG, (unsigned long long)v), m8)) { found = v; break; }
          if (j != 0) {      // tag collision or the other sign (x == 0 cannot happen for j != 0): try the mirrored candidate too
            const long long v2 = (long long)(i * stride) - r;
            if (v2 >= 0 && (unsigned long long)v2 < limit && ge_equal(mul_g8(tabG, (unsigned long long)v2), m8)) { found = v2; break; }
          }
        }
        s = (s + 1) & (n_slots - 1);
      }
    }
    // any lane done?  (every lane must take part in the vote: the loop above has no early exit)
    if (__any_sync(0xffffffffu, found >= 0)) break;
  }
  // the amount is unique in range: take the smallest found value of the warp
  unsigned long long best = found >= 0 ? (unsigned long long)found : ~0ull;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) { const unsigned long long o = __shfl_down_sync(0xffffffffu, best, d); best = o < best ? o : best; }
  if (lane == 0) { if (best != ~0ull) { out_value[w] = (long long)best; status[w] = 1; } else { out_value[w] = -1; status[w] = 0; } }
}

// M = C - s * D per ciphertext (src/elgamal.rs:140-145): thread per ciphertext, 4-bit fixed windows over s; writes affine (x, y)
__global__ void __launch_bounds__(64) k_decrypt(const uint8_t* __restrict__ cts, size_t n, const uint32_t* __restrict__ sk_words, uint32_t* __restrict__ out_aff, uint8_t* __restrict__ ok) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_aff c, d;
  const bool good = decode_words(c, cts + 64 * i) & decode_words(d, cts + 64 * i + 32);
  if (!good) { ok[i] = 0; return; }
  sc s; for (int k = 0; k < 8; k++) s.v[k] = sk_words[k];
  const ge D = ge_from_affine(d);
  ge tab[8];                                   // 1..8 multiples of D (signed 4-bit digits)
  tab[0] = D; for (int k = 1; k < 8; k++) tab[k] = ge_add(tab[k - 1], D);
  // signed radix-16 recoding of s (s < l < 2^253: the top digit absorbs the carry)
  int8_t dig[65]; uint32_t carry = 0;
  for (int w = 0; w < 64; w++) { uint32_t v = ((s.v[w >> 3] >> ((w & 7) * 4)) & 15u) + carry; carry = v > 8u ? 1u : 0u; dig[w] = (int8_t)(carry ? (int)v - 16 : (int)v); }
  dig[64] = (int8_t)carry;
  ge acc = ge_identity();
#pragma unroll 1
  for (int w = 64; w >= 0; w--) {
    if (w != 64) { acc = ge_double_pz(acc); acc = ge_double_pz(acc); acc = ge_double_pz(acc); acc = ge_double(acc); }
    const int dg = dig[w];
    if (dg > 0) acc = ge_add(acc, tab[dg - 1]); else if (dg < 0) acc = ge_add(acc, ge_neg(tab[-dg - 1]));
  }
  const ge m = ge_add(ge_from_affine(c), ge_neg(acc));      // C - s D
  const fe zi = fe_invert(m.Z);
  st_fe(out_aff + 16 * i, fe_mul(m.X, zi)); st_fe(out_aff + 16 * i + 8, fe_mul(m.Y, zi));
  ok[i] = 1;
}
__global__ void __launch_bounds__(128) k_decode_points(const uint8_t* __restrict__ enc, size_t n, uint32_t* __restrict__ out_aff, uint8_t* __restrict__ ok) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_aff a; const bool good = decode_words(a, enc + 32 * i);
  if (good) { st_fe(out_aff + 16 * i, a.x); st_fe(out_aff + 16 * i + 8, a.y); }
  ok[i] = good ? 1 : 0;
}

int32_t ensure_tmp(xhe_ecdlp* t, size_t bytes) {
  if (t->tmp_bytes >= bytes) return XHE_OK;
  if (t->d_tmp) cudaFree(t->d_tmp);
  t->d_tmp = nullptr; t->tmp_bytes = 0;
  XHE_CUDA_OK(t->ctx, cudaMalloc(&t->d_tmp, bytes + bytes / 4 + 256));
  t->tmp_bytes = bytes + bytes / 4 + 256;
  return XHE_OK;
}

// common tail: affine points + validity flags are in the scratch; run the search and copy the answers out
int32_t run_decode(xhe_ecdlp* t, uint8_t* d_in, size_t in_bytes, size_t n, uint32_t range_bits, int64_t* out_value, uint8_t* status) {
  xhe_ctx* ctx = t->ctx; cudaStream_t st = ctx->stream;
  uint32_t* d_aff = (uint32_t*)(d_in + ((in_bytes + 255) & ~(size_t)255));
  uint8_t* d_ok = (uint8_t*)(d_aff + 16 * n);
  long long* d_val = (long long*)(d_ok + ((n + 255) & ~(size_t)255));
  uint8_t* d_status = (uint8_t*)(d_val + n);
  k_ecdlp_decode<<<nblk(32 * n, 128), 128, 0, st>>>(d_aff, d_ok, n, t->tabG, t->d_consts, t->d_slots, t->n_slots, t->l1, range_bits, d_val, d_status); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaGetLastError());
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(out_value, d_val, 8 * n, cudaMemcpyDeviceToHost, st));
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, st));
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return XHE_OK;
}
size_t scratch_need(size_t in_bytes, size_t n) { return ((in_bytes + 255) & ~(size_t)255) + 64 * n + ((n + 255) & ~(size_t)255) + 8 * n + n + 1024; }
}  // namespace

extern "C" int32_t xhe_ecdlp_create(xhe_ctx* ctx, uint32_t l1_bits, xhe_ecdlp** out) {
  if (!ctx || !out || l1_bits < 8 || l1_bits > 26) return XHE_E_ARG;
  const uint32_t* tabG = xhe_internal_tabG(ctx);
  if (!tabG) return XHE_E_CUDA;
  xhe_ecdlp* t = new xhe_ecdlp(); t->ctx = ctx; t->l1 = l1_bits; t->tabG = tabG;
  t->n_slots = (size_t)1 << (l1_bits + 1);                                  // load factor 1/2
  cudaStream_t st = ctx->stream;
  if (cudaMalloc(&t->d_slots, 8 * t->n_slots) != cudaSuccess || cudaMalloc(&t->d_consts, 4 * (48 + 32 * 32)) != cudaSuccess) { ctx->err = "ecdlp: table allocation failed"; cudaGetLastError(); if (t->d_slots) cudaFree(t->d_slots); delete t; return XHE_E_NOMEM; }
  XHE_CUDA_OK(ctx, cudaMemsetAsync(t->d_slots, 0, 8 * t->n_slots, st));
  k_ecdlp_consts<<<1, 32, 0, st>>>(tabG, l1_bits, t->d_consts); XHE_LAUNCHED(ctx);
  const size_t n_threads = (((size_t)1 << l1_bits) + ECDLP_RUN) / ECDLP_RUN + 1;
  k_ecdlp_baby<<<nblk(n_threads, 128), 128, 0, st>>>(tabG, t->d_consts, l1_bits, t->d_slots, t->n_slots); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaGetLastError());
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(st));
  *out = t;
  return XHE_OK;
}
extern "C" void xhe_ecdlp_destroy(xhe_ecdlp* t) {
  if (!t) return;
  if (t->d_slots) cudaFree(t->d_slots);
  if (t->d_consts) cudaFree(t->d_consts);
  if (t->d_tmp) cudaFree(t->d_tmp);
  delete t;
}
extern "C" size_t xhe_ecdlp_table_bytes(const xhe_ecdlp* t) { return t ? 8 * t->n_slots : 0; }

extern "C" int32_t xhe_ecdlp_decode(xhe_ecdlp* t, const uint8_t* points, size_t n, uint32_t range_bits, int64_t* out_value, uint8_t* status) {
  if (!t || (n && (!points || !out_value || !status)) || range_bits < 1 || range_bits > 62) return XHE_E_ARG;
  if (range_bits > t->l1 + 1 + 32) { t->ctx->err = "ecdlp: range too wide for this table (more than 2^32 giant steps)"; return XHE_E_ARG; }
  if (!n) return XHE_OK;
  xhe_ctx* ctx = t->ctx; cudaStream_t st = ctx->stream;
  int32_t rc = ensure_tmp(t, scratch_need(32 * n, n)); if (rc) return rc;
  uint8_t* d_in = (uint8_t*)t->d_tmp;
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(d_in, points, 32 * n, cudaMemcpyHostToDevice, st));
  uint32_t* d_aff = (uint32_t*)(d_in + ((32 * n + 255) & ~(size_t)255)); uint8_t* d_ok = (uint8_t*)(d_aff + 16 * n);
  k_decode_points<<<nblk(n, 128), 128, 0, st>>>(d_in, n, d_aff, d_ok); XHE_LAUNCHED(ctx);
  return run_decode(t, d_in, 32 * n, n, range_bits, out_value, status);
}

extern "C" int32_t xhe_decrypt_decode(xhe_ecdlp* t, const uint8_t sk[32], const uint8_t* cts, size_t n, uint32_t range_bits, int64_t* out_value, uint8_t* status) {
  if (!t || !sk || (n && (!cts || !out_value || !status)) || range_bits < 1 || range_bits > 62) return XHE_E_ARG;
  if (range_bits > t->l1 + 1 + 32) { t->ctx->err = "ecdlp: range too wide for this table (more than 2^32 giant steps)"; return XHE_E_ARG; }
  if (!n) return XHE_OK;
  xhe_ctx* ctx = t->ctx; cudaStream_t st = ctx->stream;
  int32_t rc = ensure_tmp(t, scratch_need(64 * n + 256, n)); if (rc) return rc;
  uint8_t* d_in = (uint8_t*)t->d_tmp;
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(d_in, cts, 64 * n, cudaMemcpyHostToDevice, st));
  uint32_t* d_sk = (uint32_t*)(d_in + 64 * n);                                // (inside the input region: scratch_need reserved 256 bytes for it)
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(d_sk, sk, 32, cudaMemcpyHostToDevice, st));
  uint32_t* d_aff = (uint32_t*)(d_in + ((64 * n + 256 + 255) & ~(size_t)255)); uint8_t* d_ok = (uint8_t*)(d_aff + 16 * n);
  k_decrypt<<<nblk(n, 64), 64, 0, st>>>(d_in, n, d_sk, d_aff, d_ok); XHE_LAUNCHED(ctx);
  rc = run_decode(t, d_in, 64 * n + 256, n, range_bits, out_value, status);
  cudaMemsetAsync(d_sk, 0, 32, st);                                          // the secret key does not stay in device scratch
  cudaStreamSynchronize(st);
  return rc;
}
