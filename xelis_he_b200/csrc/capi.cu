// capi.cu -- context management and the host-buffer entry points of the C ABI (include/xhe.h)
#include "xhe_internal.cuh"
#include "../host/keccak.hpp"
#include <vector>
#include <algorithm>

int32_t xhe_from_uniform_niels_dev(xhe_ctx* ctx, const void* d_u, size_t n, void* d_niels);
size_t xhe_preload_msm(); size_t xhe_preload_verify(); size_t xhe_preload_fs(); size_t xhe_preload_point(); size_t xhe_preload_ledger();
int32_t xhe_compress_xy_bytes_dev(xhe_ctx* ctx, const void* d_xy, size_t n, void* d_enc);
int32_t xhe_affine_to_bytes_dev(xhe_ctx* ctx, const void* d_aff, size_t n, void* d_xy);

namespace {
struct DevBuf {   // RAII device allocation for the synchronous host-buffer entry points
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
};
}  // namespace

extern "C" int32_t xhe_ctx_create(int device, uint32_t party_capacity, xhe_ctx** out) {
  if (!out || party_capacity > 512) return XHE_E_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return XHE_E_CUDA;   // no GPU: fail loudly, there is no fallback
  if (cudaSetDevice(device) != cudaSuccess) return XHE_E_CUDA;
  // Two things CUDA does lazily would make the first launch of a kernel wait for every RUNNING kernel -- fatal while the
  // polling chain kernel of msm.cu is in flight (it waits for exactly those launches; measured: the first batch of a context
  // timed out): loading the kernel (CUDA_MODULE_LOADING=LAZY), and growing the context's local-memory pool when a kernel needs
  // a larger per-thread stack than any before it (k_rp_prep: 3.5 KB against the 1 KB default).  Both happen here instead.
  { static bool loaded[64] = {false};
    if (device < 64 && !loaded[device]) {
      size_t frame = std::max(std::max(xhe_preload_msm(), xhe_preload_verify()), std::max(xhe_preload_fs(), std::max(xhe_preload_point(), xhe_preload_ledger())));
      size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitStackSize);
      size_t want = std::max<size_t>(4096, frame + 512);       // (the pool is this many bytes x the resident threads of the device: ~1.2 GB at 4 KB)
      if (cur < want) cudaDeviceSetLimit(cudaLimitStackSize, want);
      cudaGetLastError(); loaded[device] = true; } }
  xhe_ctx* ctx = new xhe_ctx();
  ctx->device = device; ctx->party_capacity = party_capacity;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  // generator table: G, H, G_vec[party][64], H_vec[party][64]  (src/proofs.rs:19-22, src/elgamal.rs:16-24)
  size_t n = 2 + 128 * (size_t)party_capacity;
  ctx->n_gens = n;
  std::vector<uint8_t> uni(64 * n);
  {
    // G: handled on device from its encoding; here slot 0 is filled afterwards.  H = one-way-map(SHA3-512(G_compressed)).
    static const uint8_t G_ENC[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                      0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    xhe_host::sha3_512(G_ENC, 32, &uni[64]);
    for (uint32_t which = 0; which < 2; which++)
      for (uint32_t j = 0; j < party_capacity; j++) {
        uint8_t label[5] = {(uint8_t)(which ? 'H' : 'G'), (uint8_t)j, (uint8_t)(j >> 8), (uint8_t)(j >> 16), (uint8_t)(j >> 24)};
        xhe_host::Sponge sp(136); sp.absorb("GeneratorsChain", 15); sp.absorb(label, 5); sp.finish(0x1f);
        sp.squeeze(&uni[64 * (2 + (size_t)which * 64 * party_capacity + 64 * (size_t)j)], 64 * 64);
      }
  }
  DevBuf du;
  if (du.alloc(64 * n) != cudaSuccess || cudaMalloc(&ctx->d_gens_niels, 96 * n) != cudaSuccess) { delete ctx; return XHE_E_NOMEM; }
  cudaMemcpy(du.p, uni.data(), 64 * n, cudaMemcpyHostToDevice);
  if (xhe_from_uniform_niels_dev(ctx, du.p, n, ctx->d_gens_niels) != XHE_OK) { delete ctx; return XHE_E_CUDA; }
  {
    // slot 0 <- basepoint G as affine Niels (decode its encoding on the device)
    DevBuf denc, dok;
    static const uint8_t G_ENC[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                      0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    denc.alloc(32); dok.alloc(1);
    cudaMemcpy(denc.p, G_ENC, 32, cudaMemcpyHostToDevice);
    xhe_decompress_dev(ctx, denc.p, 1, nullptr, ctx->d_gens_niels, dok.p);
    if (cudaDeviceSynchronize() != cudaSuccess) { ctx->err = cudaGetErrorString(cudaGetLastError()); delete ctx; return XHE_E_CUDA; }
  }
  *out = ctx;
  return XHE_OK;
}
extern "C" uint32_t xhe_ctx_party_capacity(const xhe_ctx* ctx) { return ctx ? ctx->party_capacity : 0; }
extern "C" void xhe_ctx_destroy(xhe_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_gens_niels) cudaFree(ctx->d_gens_niels);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_fb_tab) cudaFree(ctx->d_fb_tab);
  if (ctx->d_fb_dig) cudaFree(ctx->d_fb_dig);
  if (ctx->d_fb_bsum) cudaFree(ctx->d_fb_bsum);
  if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
  if (ctx->h_res) cudaFreeHost(ctx->h_res);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->d_commit) cudaFree(ctx->d_commit);
  for (auto& st : ctx->aux) if (st) cudaStreamDestroy(st);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto& l : ctx->msm_side) for (auto& st : l) if (st) cudaStreamDestroy(st);
  for (auto& l : ctx->msm_ev) for (auto& e : l) if (e) cudaEventDestroy(e);
  delete ctx;
}
extern "C" const char* xhe_last_error(const xhe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int32_t xhe_ctx_set_stream(xhe_ctx* ctx, void* s) { if (!ctx) return XHE_E_ARG; ctx->stream = (cudaStream_t)s; return XHE_OK; }
extern "C" int32_t xhe_ctx_sync(xhe_ctx* ctx) { if (!ctx) return XHE_E_ARG; XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream)); return XHE_OK; }
extern "C" uint64_t xhe_ctx_launch_count(const xhe_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int32_t xhe_ctx_set_serial(xhe_ctx* ctx, int serial) { if (!ctx) return XHE_E_ARG; ctx->serial = serial != 0; return XHE_OK; }
extern "C" int32_t xhe_ctx_timing(xhe_ctx* ctx, int enable) {
  if (!ctx) return XHE_E_ARG;
  for (auto& p : ctx->pending) { ctx->ev_pool.push_back(p.e0); ctx->ev_pool.push_back(p.e1); }
  ctx->pending.clear(); ctx->n_timers = 0; ctx->timing = enable != 0;
  if (enable) while (ctx->ev_pool.size() < 2048) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) break; ctx->ev_pool.push_back(e); }
  for (auto& t : ctx->timers) t = xhe_ctx::KernelTimer();
  return XHE_OK;
}
// collect: returns the number of timed kernels; fills names / total ms / launches / algorithmic units (limb products)
extern "C" int32_t xhe_ctx_timing_read(xhe_ctx* ctx, const char** names, double* ms, uint64_t* launches, double* units, int cap) {
  if (!ctx) return XHE_E_ARG;
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto* s : ctx->aux) if (s) XHE_CUDA_OK(ctx, cudaStreamSynchronize(s));
  ctx->timeline.clear();
  for (size_t i = 0; i < ctx->pending.size(); i++) {
    auto& p = ctx->pending[i];
    float t = 0; cudaEventElapsedTime(&t, p.e0, p.e1); ctx->timers[p.timer].ms += t;
    if (ctx->tl_base && i >= ctx->tl_mark) {
      float a = 0, b = 0;
      if (cudaEventElapsedTime(&a, ctx->tl_base, p.e0) == cudaSuccess && cudaEventElapsedTime(&b, ctx->tl_base, p.e1) == cudaSuccess) ctx->timeline.push_back({ctx->timers[p.timer].name, a, b});
    }
    ctx->ev_pool.push_back(p.e0); ctx->ev_pool.push_back(p.e1);
  }
  (void)cudaGetLastError();
  ctx->pending.clear(); ctx->tl_mark = 0;
  int n = ctx->n_timers < cap ? ctx->n_timers : cap;
  for (int i = 0; i < n; i++) { names[i] = ctx->timers[i].name; ms[i] = ctx->timers[i].ms; launches[i] = ctx->timers[i].launches; units[i] = ctx->timers[i].units; }
  return n;
}
// spans (ms from the start of the last timed xhe_batch_run) gathered by the last xhe_ctx_timing_read
extern "C" int32_t xhe_ctx_timeline(xhe_ctx* ctx, const char** names, float* t0, float* t1, int cap) {
  if (!ctx) return XHE_E_ARG;
  int n = (int)ctx->timeline.size() < cap ? (int)ctx->timeline.size() : cap;
  for (int i = 0; i < n; i++) { names[i] = ctx->timeline[i].name; t0[i] = ctx->timeline[i].t0; t1[i] = ctx->timeline[i].t1; }
  return n;
}
extern "C" const void* xhe_ctx_generators_dev(const xhe_ctx* ctx, size_t* n) { if (n) *n = ctx->n_gens; return ctx->d_gens_niels; }

#define H2D(dst, src, n) XHE_CUDA_OK(ctx, cudaMemcpyAsync((dst), (src), (n), cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, n) XHE_CUDA_OK(ctx, cudaMemcpyAsync((dst), (src), (n), cudaMemcpyDeviceToHost, ctx->stream))

extern "C" int32_t xhe_ristretto_decompress(xhe_ctx* ctx, const uint8_t* enc, size_t n, uint8_t* xy, uint8_t* ok) {
  if (!ctx || (n && (!enc || !ok))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  DevBuf denc, daff, dxy, dok;
  XHE_CUDA_OK(ctx, denc.alloc(32 * n)); XHE_CUDA_OK(ctx, daff.alloc(64 * n)); XHE_CUDA_OK(ctx, dok.alloc(n));
  H2D(denc.p, enc, 32 * n);
  int32_t rc = xhe_decompress_dev(ctx, denc.p, n, daff.p, nullptr, dok.p); if (rc) return rc;
  if (xy) { XHE_CUDA_OK(ctx, dxy.alloc(64 * n)); rc = xhe_affine_to_bytes_dev(ctx, daff.p, n, dxy.p); if (rc) return rc; D2H(xy, dxy.p, 64 * n); }
  D2H(ok, dok.p, n);
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return XHE_OK;
}
extern "C" int32_t xhe_ristretto_compress(xhe_ctx* ctx, const uint8_t* xy, size_t n, uint8_t* enc) {
  if (!ctx || (n && (!xy || !enc))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  DevBuf dxy, denc;
  XHE_CUDA_OK(ctx, dxy.alloc(64 * n)); XHE_CUDA_OK(ctx, denc.alloc(32 * n));
  H2D(dxy.p, xy, 64 * n);
  int32_t rc = xhe_compress_xy_bytes_dev(ctx, dxy.p, n, denc.p); if (rc) return rc;
  D2H(enc, denc.p, 32 * n);
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return XHE_OK;
}
extern "C" int32_t xhe_ristretto_from_uniform(xhe_ctx* ctx, const uint8_t* u, size_t n, uint8_t* enc) {
  if (!ctx || (n && (!u || !enc))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  DevBuf du, denc;
  XHE_CUDA_OK(ctx, du.alloc(64 * n)); XHE_CUDA_OK(ctx, denc.alloc(32 * n));
  H2D(du.p, u, 64 * n);
  int32_t rc = xhe_from_uniform_dev(ctx, du.p, n, denc.p); if (rc) return rc;
  D2H(enc, denc.p, 32 * n);
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return XHE_OK;
}
extern "C" int32_t xhe_ct_update(xhe_ctx* ctx, const uint8_t* bal, const uint8_t* delta, const uint8_t* sub, size_t n, uint8_t* out, uint8_t* ok) {
  if (!ctx || (n && (!bal || !delta || !sub || !out || !ok))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  DevBuf dbal, ddelta, dsub, dout, dok;
  XHE_CUDA_OK(ctx, dbal.alloc(64 * n)); XHE_CUDA_OK(ctx, ddelta.alloc(64 * n)); XHE_CUDA_OK(ctx, dsub.alloc(n)); XHE_CUDA_OK(ctx, dout.alloc(64 * n)); XHE_CUDA_OK(ctx, dok.alloc(n));
  H2D(dbal.p, bal, 64 * n); H2D(ddelta.p, delta, 64 * n); H2D(dsub.p, sub, n);
  int32_t rc = xhe_ct_update_dev(ctx, dbal.p, ddelta.p, dsub.p, n, dout.p, dok.p); if (rc) return rc;
  D2H(out, dout.p, 64 * n); D2H(ok, dok.p, n);
  XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return XHE_OK;
}
