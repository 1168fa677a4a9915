// xhe_internal.cuh -- context, error plumbing and device memory layouts shared by the .cu files of libxhe_cuda.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>
#include <vector>
#include "../../include/xhe.h"
#include "ge25519.cuh"
#include "sc25519.cuh"

struct xhe_ctx {
  int device = 0;
  int sm_count = 148;
  uint32_t party_capacity = 0;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  std::string err;
  // generator tables (device): affine Niels, index 0 = G, 1 = H, then G_vec[party][64], H_vec[party][64]
  void* d_gens_niels = nullptr;   // (2 + 128 * party_capacity) * 96 B
  size_t n_gens = 0;
  // fixed-base table of the range-proof generators (verify.cu, k_fb_*): entry (g, j) = 2^(8j) * gens[g], affine Niels;
  // built on the first batch with range proofs when party_capacity <= XHE_FB_MAX_PARTIES, else the generic MSM is used
  void* d_fb_tab = nullptr; void* d_fb_dig = nullptr; void* d_fb_bsum = nullptr;
  void* d_scratch = nullptr; size_t scratch_bytes = 0;       // grow-only device scratch for host-buffer entry points
  void* h_pinned = nullptr; size_t pinned_bytes = 0;         // grow-only pinned staging
  void* d_commit = nullptr; size_t commit_bytes = 0;         // slots / ops of xhe_ledger_commit_batch
  // optional CUDA-event timing of the main kernels (bench.py roofline): accumulated since the last reset
  bool timing = false;
  bool serial = false;                                        // diagnostics: run the pipelines of xhe_batch_run back to back on one stream
  struct KernelTimer { const char* name; double ms = 0; uint64_t launches = 0; double units = 0; };
  KernelTimer timers[16];
  int n_timers = 0;
  struct Pending { int timer; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> ev_pool;                           // timing events, created when timing is switched on: nothing may be created while a chain kernel polls (msm.cu)
  // timeline of the last timed xhe_batch_run (start/end of every timed kernel relative to the start of the run)
  cudaEvent_t tl_base = nullptr; size_t tl_mark = 0;
  struct Span { const char* name; float t0, t1; };
  std::vector<Span> timeline;
  cudaEvent_t sync_ev = nullptr;                              // blocking-sync event: host waits yield the core instead of spinning
  uint32_t* h_res = nullptr;                                  // pinned 512-byte landing zone of a batch's result block
  void* h_small = nullptr;                                    // pinned (device-mapped) 7 KiB input block of xhe_sum_encodings
  void* resident = nullptr;                                   // DeviceBatch of the batch currently resident (verify.cu)
  cudaStream_t aux[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // side streams for the independent pipelines of xhe_batch_run
  cudaEvent_t ev[14] = {nullptr};
  // grouped MSM tail (msm.cu): per lane a reduction stream, a Horner stream and their events
  cudaStream_t msm_side[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  cudaEvent_t msm_ev[2][17] = {{nullptr}, {nullptr}};
};

// scoped timing of one kernel launch on ctx->stream (no-op unless ctx->timing)
struct XheTimed {
  xhe_ctx* ctx; int idx = -1; cudaEvent_t e0 = nullptr, e1 = nullptr;
  XheTimed(xhe_ctx* c, const char* name, double units) : ctx(c) {
    if (!c->timing) return;
    for (int i = 0; i < c->n_timers; i++) if (!strcmp(c->timers[i].name, name)) idx = i;
    if (idx < 0 && c->n_timers < 16) { idx = c->n_timers++; c->timers[idx].name = name; }
    if (idx < 0) return;
    if (c->ev_pool.size() < 2) { idx = -1; return; }          // pool exhausted: this launch goes untimed
    c->timers[idx].launches++; c->timers[idx].units += units;
    e0 = c->ev_pool.back(); c->ev_pool.pop_back(); e1 = c->ev_pool.back(); c->ev_pool.pop_back();
    cudaEventRecord(e0, c->stream);
  }
  ~XheTimed() { if (idx >= 0) { cudaEventRecord(e1, ctx->stream); ctx->pending.push_back({idx, e0, e1}); } }
};

#define XHE_CUDA_OK(ctx, call)                                                                       \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                              \
      return XHE_E_CUDA;                                                                             \
    }                                                                                                \
  } while (0)

#define XHE_LAUNCHED(ctx) do { (ctx)->launches++; } while (0)

// Wait for everything queued on `st`.  cudaStreamSynchronize spins on a core for the whole wait; with several batches in
// flight (one host thread per context) those spinning threads starve the host phases of the other batches, so the hot
// path waits on a blocking-sync event instead.
inline cudaError_t xhe_wait_stream(xhe_ctx* ctx, cudaStream_t st) {
  if (!ctx->sync_ev) { cudaError_t e = cudaEventCreateWithFlags(&ctx->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  cudaError_t e = cudaEventRecord(ctx->sync_ev, st); if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->sync_ev);
}

namespace xhe {

// 128-bit vector load/store of limb arrays (all device point arrays are 32-byte aligned)
__device__ __forceinline__ void ld_fe(fe& r, const uint32_t* p) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
}
__device__ __forceinline__ void ld_fe_rw(fe& r, const uint32_t* p) {   // plain (non-nc) load for buffers written in the same kernel
  uint4 a = *reinterpret_cast<const uint4*>(p), b = *(reinterpret_cast<const uint4*>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
}
__device__ __forceinline__ void st_fe(uint32_t* p, const fe& r) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void ld_bytes32(uint8_t* dst, const uint8_t* src) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(src)), b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; i++) { dst[4 * i] = (uint8_t)w[i]; dst[4 * i + 1] = (uint8_t)(w[i] >> 8); dst[4 * i + 2] = (uint8_t)(w[i] >> 16); dst[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
// ristretto decode straight from global memory words (no byte shuffling): returns ok
__device__ __forceinline__ bool decode_words(ge_aff& out, const uint8_t* enc32) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(enc32)), b = __ldg(reinterpret_cast<const uint4*>(enc32) + 1);
  uint8_t bytes[32];
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; i++) { bytes[4 * i] = (uint8_t)w[i]; bytes[4 * i + 1] = (uint8_t)(w[i] >> 8); bytes[4 * i + 2] = (uint8_t)(w[i] >> 16); bytes[4 * i + 3] = (uint8_t)(w[i] >> 24); }
  return ristretto_decode(out, bytes);
}
__device__ __forceinline__ void encode_words(uint8_t* enc32, const ge& p, ge_aff* aff = nullptr) {
  uint8_t bytes[32];
  ristretto_encode(bytes, p, aff);
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = (uint32_t)bytes[4 * i] | ((uint32_t)bytes[4 * i + 1] << 8) | ((uint32_t)bytes[4 * i + 2] << 16) | ((uint32_t)bytes[4 * i + 3] << 24);
  reinterpret_cast<uint4*>(enc32)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4*>(enc32)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
// encoding of a RESULT that is usually the identity (a verified batch's MSM): every representative of the identity coset
// (X == 0 or Y == 0) encodes to 32 zero bytes (RFC 9496 4.3.2 is constant on cosets), so the ~265-multiplication inverse
// square root -- a 60 us single-thread chain at the very end of a step -- runs only for results that are not the identity
__device__ __forceinline__ void encode_result_words(uint8_t* enc32, const ge& p) {
  if (ge_ristretto_is_identity(p)) { reinterpret_cast<uint4*>(enc32)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(enc32)[1] = make_uint4(0, 0, 0, 0); }
  else encode_words(enc32, p);
}
__device__ __forceinline__ void ld_niels(ge_niels& n, const uint32_t* p) { ld_fe(n.ypx, p); ld_fe(n.ymx, p + 8); ld_fe(n.t2d, p + 16); }
__device__ __forceinline__ void st_niels(uint32_t* p, const ge_niels& n) { st_fe(p, n.ypx); st_fe(p + 8, n.ymx); st_fe(p + 16, n.t2d); }
__device__ __forceinline__ void ld_ge(ge& g, const uint32_t* p) { ld_fe_rw(g.X, p); ld_fe_rw(g.Y, p + 8); ld_fe_rw(g.Z, p + 16); ld_fe_rw(g.T, p + 24); }
__device__ __forceinline__ void st_ge(uint32_t* p, const ge& g) { st_fe(p, g.X); st_fe(p + 8, g.Y); st_fe(p + 16, g.Z); st_fe(p + 24, g.T); }

}  // namespace xhe

// kernel launchers implemented across the .cu files
int32_t xhe_msm_sort(xhe_ctx* ctx, const void* d_scalars, size_t n, void* d_ws, size_t ws_bytes, void* d_bad_flag);
int32_t xhe_msm_finish(xhe_ctx* ctx, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id, void* d_out_ext, cudaEvent_t after_accum = nullptr, int lane = 0, int chain_mode = -1);
// the Horner chain of an MSM as a job for the polling chain kernel (msm.cu): one kernel can serve two MSMs
struct XheChainJob { const uint32_t* hnodes; uint32_t* ready; int G, Wg, W, c; uint8_t* out_enc; uint32_t* is_identity; uint32_t* out_ext; uint32_t* status; };
XheChainJob xhe_msm_chain_job(size_t n, void* d_ws, void* d_out_enc, void* d_is_id, void* d_out_ext);
int32_t xhe_msm_chain_reset(xhe_ctx* ctx, cudaStream_t st, const XheChainJob& j);
int32_t xhe_msm_chain_launch(xhe_ctx* ctx, cudaStream_t st, const XheChainJob& a, const XheChainJob& b, int wait);
bool xhe_msm_chain_enabled();
int32_t xhe_msm_side_init(xhe_ctx* ctx, int lane);
int32_t xhe_launch_msm(xhe_ctx* ctx, const void* d_scalars, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id);
