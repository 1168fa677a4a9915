// ledger.cu -- device-resident ledger: a state backend for BlockchainVerificationState
// (reference src/tx/verify.rs:25-77) that keeps the balances ON the device as decompressed extended points, so that the
// balance algebra of src/elgamal.rs:322-342 / src/tx/verify.rs:574,602 runs on resident data -- a dense update of every
// account is the HBM-bound kernel of config 4 -- and compression happens only when somebody asks for the bytes.
//
// Layout: coordinate-planar extended points [X | Y | Z | T][2 * capacity][8 words], two points (commitment, handle) per
// (account, asset) slot, the layout k_ct_update_resident streams at 0.74 of the HBM copy peak.  The (account || asset) ->
// slot index lives on the host (the open-addressing table of the host layer): lookups are the caller's, arithmetic is the
// device's.  One ledger belongs to one context (its stream orders every operation).
#include "xhe_internal.cuh"
#include "../host/verifier.hpp"
#include <algorithm>
#include <vector>
using namespace xhe;

struct xhe_ledger {
  xhe_ctx* ctx = nullptr;
  size_t cap = 0;                       // slots
  uint32_t* d_bal = nullptr;            // 4 planes x 2 cap points x 8 words
  xhe_host::FlatTable<64, uint32_t> index;
  std::vector<uint8_t> slot_ok;         // 0 for a slot whose stored ciphertext did not decode
  // grow-only staging
  uint32_t* d_snap = nullptr;           // snapshot of the table (xhe_ledger_snapshot)
  void* d_tmp = nullptr; size_t tmp_bytes = 0;
  void* h_tmp = nullptr; size_t h_bytes = 0;
};

namespace {
__device__ __forceinline__ void ld_planar(ge& p, const uint32_t* bal, size_t stride, size_t i) {
  ld_fe_rw(p.X, bal + 8 * i); ld_fe_rw(p.Y, bal + 8 * (stride + i)); ld_fe_rw(p.Z, bal + 8 * (2 * stride + i)); ld_fe_rw(p.T, bal + 8 * (3 * stride + i));
}
__device__ __forceinline__ void st_planar(uint32_t* bal, size_t stride, size_t i, const ge& p) {
  st_fe(bal + 8 * i, p.X); st_fe(bal + 8 * (stride + i), p.Y); st_fe(bal + 8 * (2 * stride + i), p.Z); st_fe(bal + 8 * (3 * stride + i), p.T);
}
// decode n ciphertexts (2 n encodings) into the slots: one thread per point
__global__ void __launch_bounds__(128) k_ledger_store(const uint8_t* __restrict__ enc, const uint32_t* __restrict__ slots, size_t n_points, size_t stride, uint32_t* __restrict__ bal, uint8_t* __restrict__ ok_pt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  ge_aff a; bool good = decode_words(a, enc + 32 * i);
  if (!good) a = ge_aff_identity();
  st_planar(bal, stride, 2 * (size_t)slots[i >> 1] + (i & 1), ge_from_affine(a));
  ok_pt[i] = good ? 1 : 0;
}
// bal[slot] +/- delta, delta given as a compressed ciphertext (decode + 7 M mixed addition): integer-bound
__global__ void __launch_bounds__(128) k_ledger_update(const uint8_t* __restrict__ delta, const uint32_t* __restrict__ slots, const uint8_t* __restrict__ sub, size_t n_points, size_t stride,
                                                       uint32_t* __restrict__ bal, uint8_t* __restrict__ ok_pt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  ge_aff d; bool good = decode_words(d, delta + 32 * i);
  ok_pt[i] = good ? 1 : 0;
  if (!good) return;                                   // an ill-formed delta leaves the balance alone (the caller gets the flag)
  const size_t p = 2 * (size_t)slots[i >> 1] + (i & 1);
  ge b; ld_planar(b, bal, stride, p);
  st_planar(bal, stride, p, ge_madd(b, niels_cneg(niels_from_affine(d), sub[i >> 1] != 0)));
}
// the same for EVERY slot [0, n_slots), deltas already resident as planar affine Niels [ypx | ymx | t2d][2 n_slots][8]: HBM-bound
__global__ void __launch_bounds__(128) k_ledger_update_dense(uint32_t* __restrict__ bal, size_t stride, const uint32_t* __restrict__ delta, const uint8_t* __restrict__ sub, size_t n_points) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  ge p; ge_niels q;
  ld_planar(p, bal, stride, i);
  ld_fe(q.ypx, delta + 8 * i); ld_fe(q.ymx, delta + 8 * (n_points + i)); ld_fe(q.t2d, delta + 8 * (2 * n_points + i));
  st_planar(bal, stride, i, ge_madd(p, niels_cneg(q, sub[i >> 1] != 0)));
}
// compressed export on demand: one encode (inverse square root) per point
__global__ void __launch_bounds__(128) k_ledger_export(const uint32_t* __restrict__ slots, size_t n_points, size_t stride, const uint32_t* __restrict__ bal, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  ge p; ld_planar(p, bal, stride, 2 * (size_t)slots[i >> 1] + (i & 1));
  encode_words(out + 32 * i, p);
}
inline unsigned nblk(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

int32_t reserve(xhe_ledger* l, size_t dev_bytes, size_t host_bytes) {
  xhe_ctx* ctx = l->ctx;
  if (l->tmp_bytes < dev_bytes) { if (l->d_tmp) cudaFree(l->d_tmp); l->d_tmp = nullptr; l->tmp_bytes = 0; XHE_CUDA_OK(ctx, cudaMalloc(&l->d_tmp, dev_bytes + dev_bytes / 4)); l->tmp_bytes = dev_bytes + dev_bytes / 4; }
  if (l->h_bytes < host_bytes) { if (l->h_tmp) cudaFreeHost(l->h_tmp); l->h_tmp = nullptr; l->h_bytes = 0; XHE_CUDA_OK(ctx, cudaHostAlloc(&l->h_tmp, host_bytes + host_bytes / 4, cudaHostAllocDefault)); l->h_bytes = host_bytes + host_bytes / 4; }
  return XHE_OK;
}
}  // namespace

extern "C" int32_t xhe_ledger_create(xhe_ctx* ctx, size_t capacity, xhe_ledger** out) {
  if (!ctx || !out || capacity == 0 || capacity > ((size_t)1 << 30)) return XHE_E_ARG;
  xhe_ledger* l = new xhe_ledger(); l->ctx = ctx; l->cap = capacity;
  if (cudaMalloc(&l->d_bal, 4 * 2 * capacity * 32) != cudaSuccess) { delete l; ctx->err = "ledger: out of device memory"; return XHE_E_NOMEM; }
  l->index.reserve(capacity); l->slot_ok.reserve(capacity);
  *out = l;
  return XHE_OK;
}
extern "C" void xhe_ledger_destroy(xhe_ledger* l) {
  if (!l) return;
  cudaSetDevice(l->ctx->device);
  cudaFree(l->d_bal); if (l->d_snap) cudaFree(l->d_snap); if (l->d_tmp) cudaFree(l->d_tmp); if (l->h_tmp) cudaFreeHost(l->h_tmp);
  delete l;
}
extern "C" size_t xhe_ledger_size(const xhe_ledger* l) { return l ? l->index.size() : 0; }
extern "C" void* xhe_ledger_device_table(const xhe_ledger* l, size_t* plane_stride_points) { if (plane_stride_points) *plane_stride_points = l ? 2 * l->cap : 0; return l ? l->d_bal : nullptr; }

// insert / overwrite n balances (get_account_balance's backing store, src/tx/verify.rs:29-34): keys n x 64 (account || asset),
// cts n x 64 compressed.  ok[i] = 0: that ciphertext does not decode (stored as identity, exported as "not found").
extern "C" int32_t xhe_ledger_load(xhe_ledger* l, const uint8_t* keys, const uint8_t* cts, size_t n, uint8_t* ok) {
  if (!l || (n && (!keys || !cts))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  xhe_ctx* ctx = l->ctx; cudaStream_t st = ctx->stream;
  int32_t rc = reserve(l, 64 * n + 4 * n + 2 * n + 256, 64 * n + 4 * n + 2 * n + 256); if (rc) return rc;
  uint8_t* h = (uint8_t*)l->h_tmp; uint32_t* h_slots = (uint32_t*)(h + 64 * n);
  for (size_t i = 0; i < n; i++) {
    bool fresh = false; uint32_t* v = l->index.insert(keys + 64 * i, &fresh);
    if (fresh) { if (l->slot_ok.size() >= l->cap) { ctx->err = "ledger: capacity exceeded"; return XHE_E_NOMEM; } *v = (uint32_t)l->slot_ok.size(); l->slot_ok.push_back(1); }
    h_slots[i] = *v;
  }
  memcpy(h, cts, 64 * n);
  uint8_t* d = (uint8_t*)l->d_tmp; uint32_t* d_slots = (uint32_t*)(d + 64 * n); uint8_t* d_ok = d + 64 * n + 4 * n;
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(d, h, 64 * n + 4 * n, cudaMemcpyHostToDevice, st));
  k_ledger_store<<<nblk(2 * n, 128), 128, 0, st>>>(d, d_slots, 2 * n, 2 * l->cap, l->d_bal, d_ok); XHE_LAUNCHED(ctx);
  uint8_t* h_ok = h + 64 * n + 4 * n;
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(h_ok, d_ok, 2 * n, cudaMemcpyDeviceToHost, st));
  XHE_CUDA_OK(ctx, xhe_wait_stream(ctx, st));
  for (size_t i = 0; i < n; i++) { uint8_t good = h_ok[2 * i] & h_ok[2 * i + 1]; l->slot_ok[h_slots[i]] = good; if (ok) ok[i] = good; }
  return XHE_OK;
}

// ElGamalCiphertext Add / Sub in place (src/elgamal.rs:322-342; what apply_without_verify does per balance, src/tx/verify.rs:574,602):
// bal[key_i] = bal[key_i] +/- delta_i.  A key that occurs several times is updated in order.  status[i]: 0 applied, 1 unknown
// key, 2 ill-formed delta (balance untouched).
extern "C" int32_t xhe_ledger_update(xhe_ledger* l, const uint8_t* keys, const uint8_t* deltas, const uint8_t* sub, size_t n, uint8_t* status) {
  if (!l || (n && (!keys || !deltas || !sub))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  xhe_ctx* ctx = l->ctx; cudaStream_t st = ctx->stream;
  int32_t rc = reserve(l, 64 * n + 4 * n + n + 2 * n + 256, 64 * n + 4 * n + n + 2 * n + 256); if (rc) return rc;
  // two updates of one slot must not race: the batch is cut into rounds in which every slot occurs once (usually one round)
  std::vector<uint32_t> order; order.reserve(n);
  std::vector<uint32_t> slot_of(n); std::vector<uint8_t> stat(n, 0);
  for (size_t i = 0; i < n; i++) { const uint32_t* v = l->index.find(keys + 64 * i); if (!v || !l->slot_ok[*v]) { stat[i] = 1; continue; } slot_of[i] = *v; order.push_back((uint32_t)i); }
  std::vector<uint32_t> round_of(l->slot_ok.size(), 0), todo = order, next;
  uint32_t round = 1;
  while (!todo.empty()) {
    std::vector<uint32_t> now; next.clear();
    for (uint32_t i : todo) { if (round_of[slot_of[i]] == round) next.push_back(i); else { round_of[slot_of[i]] = round; now.push_back(i); } }
    const size_t m = now.size();
    uint8_t* h = (uint8_t*)l->h_tmp; uint32_t* h_slots = (uint32_t*)(h + 64 * m); uint8_t* h_sub = h + 64 * m + 4 * m;
    for (size_t j = 0; j < m; j++) { memcpy(h + 64 * j, deltas + 64 * (size_t)now[j], 64); h_slots[j] = slot_of[now[j]]; h_sub[j] = sub[now[j]]; }
    uint8_t* d = (uint8_t*)l->d_tmp; uint32_t* d_slots = (uint32_t*)(d + 64 * m); uint8_t* d_sub = d + 64 * m + 4 * m; uint8_t* d_ok = d_sub + ((m + 15) & ~(size_t)15);
    XHE_CUDA_OK(ctx, cudaMemcpyAsync(d, h, 64 * m + 4 * m + m, cudaMemcpyHostToDevice, st));
    k_ledger_update<<<nblk(2 * m, 128), 128, 0, st>>>(d, d_slots, d_sub, 2 * m, 2 * l->cap, l->d_bal, d_ok); XHE_LAUNCHED(ctx);
    uint8_t* h_ok = h + 64 * m + 4 * m + ((m + 15) & ~(size_t)15);
    XHE_CUDA_OK(ctx, cudaMemcpyAsync(h_ok, d_ok, 2 * m, cudaMemcpyDeviceToHost, st));
    XHE_CUDA_OK(ctx, xhe_wait_stream(ctx, st));
    for (size_t j = 0; j < m; j++) if (!(h_ok[2 * j] & h_ok[2 * j + 1])) stat[now[j]] = 2;
    todo.swap(next); round++;
  }
  if (status) memcpy(status, stat.data(), n);
  return XHE_OK;
}

// the same algebra for EVERY slot [0, size) with deltas that are already on the device as planar affine Niels
// [ypx | ymx | t2d][2 size][8 words] (config 4: ciphertext add/sub over 1 M accounts): asynchronous, HBM-bound
extern "C" int32_t xhe_ledger_update_dense_dev(xhe_ledger* l, const void* d_delta_niels_planar, const void* d_sub) {
  if (!l || !d_delta_niels_planar || !d_sub) return XHE_E_ARG;
  const size_t n = l->index.size(); if (!n) return XHE_OK;
  xhe_ctx* ctx = l->ctx;
  k_ledger_update_dense<<<nblk(2 * n, 128), 128, 0, ctx->stream>>>(l->d_bal, 2 * l->cap, (const uint32_t*)d_delta_niels_planar, (const uint8_t*)d_sub, 2 * n); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaGetLastError());
  return XHE_OK;
}

// compressed export on demand (what get_account_balance hands out, src/tx/verify.rs:29-34): found[i] = 0 for an unknown key
extern "C" int32_t xhe_ledger_export(xhe_ledger* l, const uint8_t* keys, size_t n, uint8_t* out_cts, uint8_t* found) {
  if (!l || (n && (!keys || !out_cts))) return XHE_E_ARG;
  if (!n) return XHE_OK;
  xhe_ctx* ctx = l->ctx; cudaStream_t st = ctx->stream;
  int32_t rc = reserve(l, 64 * n + 4 * n + 256, 64 * n + 4 * n + 256); if (rc) return rc;
  uint8_t* h = (uint8_t*)l->h_tmp; uint32_t* h_slots = (uint32_t*)(h + 64 * n);
  std::vector<uint8_t> fnd(n, 1);
  for (size_t i = 0; i < n; i++) { const uint32_t* v = l->index.find(keys + 64 * i); if (!v || !l->slot_ok[*v]) { fnd[i] = 0; h_slots[i] = 0; } else h_slots[i] = *v; }
  uint8_t* d = (uint8_t*)l->d_tmp; uint32_t* d_slots = (uint32_t*)(d + 64 * n);
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(d_slots, h_slots, 4 * n, cudaMemcpyHostToDevice, st));
  k_ledger_export<<<nblk(2 * n, 128), 128, 0, st>>>(d_slots, 2 * n, 2 * l->cap, l->d_bal, d); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(h, d, 64 * n, cudaMemcpyDeviceToHost, st));
  XHE_CUDA_OK(ctx, xhe_wait_stream(ctx, st));
  for (size_t i = 0; i < n; i++) { if (fnd[i]) memcpy(out_cts + 64 * i, h + 64 * i, 64); else memset(out_cts + 64 * i, 0, 64); }
  if (found) memcpy(found, fnd.data(), n);
  return XHE_OK;
}

extern "C" uint32_t xhe_ledger_slot(const xhe_ledger* l, const uint8_t key64[64]) {
  if (!l || !key64) return 0xFFFFFFFFu;
  const uint32_t* v = l->index.find(key64);
  return (v && l->slot_ok[*v]) ? *v : 0xFFFFFFFFu;
}
extern "C" int32_t xhe_ledger_snapshot(xhe_ledger* l) {
  if (!l) return XHE_E_ARG;
  xhe_ctx* ctx = l->ctx; const size_t bytes = 4 * 2 * l->cap * 32;
  if (!l->d_snap) XHE_CUDA_OK(ctx, cudaMalloc(&l->d_snap, bytes));
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(l->d_snap, l->d_bal, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return XHE_OK;
}
extern "C" int32_t xhe_ledger_restore(xhe_ledger* l) {
  if (!l || !l->d_snap) return XHE_E_ARG;
  xhe_ctx* ctx = l->ctx;
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(l->d_bal, l->d_snap, 4 * 2 * l->cap * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  return XHE_OK;
}

size_t xhe_preload_ledger() {
  const void* ks[] = {(const void*)k_ledger_store, (const void*)k_ledger_update, (const void*)k_ledger_update_dense, (const void*)k_ledger_export};
  cudaFuncAttributes a; size_t mx = 0;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) == cudaSuccess && a.localSizeBytes > mx) mx = a.localSizeBytes;
  return mx;
}
