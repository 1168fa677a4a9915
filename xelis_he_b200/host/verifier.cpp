// host/verifier.cpp -- see verifier.hpp.  HOST side of Transaction::verify_batch / apply_without_verify.
#include "verifier.hpp"
#include "blake3.hpp"
#include "keccak.hpp"
#include "scalar_host.hpp"

#include <algorithm>
#include <chrono>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <thread>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>

namespace xhe_host {

// grow-only page-locked vector: contents of the big per-batch arrays are built directly in pinned memory so the
// uploads in xhe_batch_prepare are asynchronous DMA transfers; one set is cached per ctx and reused across batches
template <class T>
struct PinnedVec {
  T* p = nullptr; size_t n = 0, cap = 0;
  ~PinnedVec() { if (p) { if (was_pageable) free(p); else cudaFreeHost(p); } }
  PinnedVec() {}
  PinnedVec(const PinnedVec&) = delete; PinnedVec& operator=(const PinnedVec&) = delete;
  void reserve(size_t c) {
    if (c <= cap) return;
    size_t nc = cap ? cap : 1024; while (nc < c) nc *= 2;
    T* q = nullptr;
    if (cudaHostAlloc((void**)&q, nc * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); q = (T*)malloc(nc * sizeof(T)); pageable = true; }
    if (n) memcpy(q, p, n * sizeof(T));
    if (p) { if (was_pageable) free(p); else cudaFreeHost(p); }
    p = q; cap = nc; was_pageable = pageable;
  }
  void clear() { n = 0; }
  size_t size() const { return n; }
  T* data() { return p; } const T* data() const { return p; }
  T& operator[](size_t i) { return p[i]; } const T& operator[](size_t i) const { return p[i]; }
  T& back() { return p[n - 1]; }
  void push_back(const T& v) { if (n == cap) reserve(n + 1); p[n++] = v; }
  void append(const T* src, size_t k) { if (n + k > cap) reserve(n + k); memcpy(p + n, src, k * sizeof(T)); n += k; }
  void resize(size_t k) { if (k > cap) reserve(k); if (k > n) memset(p + n, 0, (k - n) * sizeof(T)); n = k; }
  void assign(size_t k, T v) { resize(0); reserve(k); for (size_t i = 0; i < k; i++) p[i] = v; n = k; }
  bool pageable = false, was_pageable = false;
};

static inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
static inline void put_be64(std::vector<uint8_t>& o, uint64_t v) { for (int i = 7; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i))); }
static inline void put(std::vector<uint8_t>& o, const uint8_t* p, size_t n) { o.insert(o.end(), p, p + n); }
static const uint8_t ZERO32[32] = {0};
static inline bool is_zero32(const uint8_t* p) { return memcmp(p, ZERO32, 32) == 0; }
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------------------------------
// wire format (spec: oracle/tx.h header comment -- shared as a format, not as code)
// ------------------------------------------------------------------------------------------------------------------
int TxView::parse(const uint8_t* b, size_t n) {
  blob = b; len = n; transfers.clear();
  body = rp = sc = ms = sig = source = nullptr; body_len = 0; n_ms = -1; n_sc = 0; count = aux = rp_len = 0;      // (a view may be reused for another transaction)
  if (n < 128) return XHE_ERR_PARSE;
  version = b[0]; type = b[1]; n_sc = b[2]; n_ms = b[3] == 0xFF ? -1 : b[3];
  count = rd32(b + 4); aux = rd32(b + 8); rp_len = rd32(b + 12); source = b + 16; fee = rd64(b + 48); nonce = rd64(b + 56);
  if (type > 4) return XHE_ERR_PARSE;
  size_t off = 64;
  auto need = [&](size_t k) { return n - off >= k; };
  body = b + off;
  switch (type) {
    case 0:
      if (count > 65535) return XHE_ERR_PARSE;
      transfers.resize(count);
      for (uint32_t i = 0; i < count; i++) {
        if (!need(324)) return XHE_ERR_PARSE;
        TransferView& t = transfers[i]; const uint8_t* p = b + off;
        t.asset = p; t.dest = p + 32; t.commitment = p + 64; t.sender_handle = p + 96; t.receiver_handle = p + 128; t.proof = p + 160; off += 320;
        uint32_t el = rd32(b + off); off += 4; t.has_extra = el != 0xFFFFFFFFu; t.extra_len = t.has_extra ? el : 0;
        if (!need(t.extra_len)) return XHE_ERR_PARSE;
        t.extra = b + off; off += t.extra_len;
        if (!ScalarL::is_canonical(t.proof + 96) || !ScalarL::is_canonical(t.proof + 128)) return XHE_ERR_PARSE;
      }
      break;
    case 1: if (!need(40)) return XHE_ERR_PARSE; off += 40; break;
    case 2:
      if (!need(32) || count > 65535 || aux > 65535) return XHE_ERR_PARSE; off += 32;
      if (!need((size_t)count * 40)) return XHE_ERR_PARSE; off += (size_t)count * 40;
      for (uint32_t i = 0; i < aux * 2; i++) { if (!need(4)) return XHE_ERR_PARSE; uint32_t l = rd32(b + off); off += 4; if (!need(l)) return XHE_ERR_PARSE; off += l; }
      break;
    case 3: if (!need(aux)) return XHE_ERR_PARSE; off += aux; break;
    case 4: if (count > 255 || aux > 255 || !need((size_t)count * 32)) return XHE_ERR_PARSE; off += (size_t)count * 32; break;
  }
  body_len = off - 64;
  if (!need(rp_len)) return XHE_ERR_PARSE; rp = b + off; off += rp_len;
  if (!need((size_t)n_sc * 256)) return XHE_ERR_PARSE; sc = b + off; off += (size_t)n_sc * 256;
  for (int i = 0; i < n_sc; i++) for (int k = 0; k < 3; k++) if (!ScalarL::is_canonical(sc + 256 * i + 160 + 32 * k)) return XHE_ERR_PARSE;
  if (n_ms > 0) { if (!need((size_t)n_ms * 65)) return XHE_ERR_PARSE; ms = b + off; off += (size_t)n_ms * 65; }
  if (!need(64)) return XHE_ERR_PARSE; sig = b + off; off += 64;
  if (off != n) return XHE_ERR_PARSE;
  // RangeProof::from_bytes structure + canonical scalars (FormatError at deserialisation time in the reference)
  if (rp_len % 32 || rp_len < 9 * 32 || ((rp_len / 32 - 9) & 1) || (rp_len / 32 - 9) / 2 >= 32) return XHE_ERR_PARSE;
  if (!ScalarL::is_canonical(rp + 128) || !ScalarL::is_canonical(rp + 160) || !ScalarL::is_canonical(rp + 192) || !ScalarL::is_canonical(rp + rp_len - 64) || !ScalarL::is_canonical(rp + rp_len - 32)) return XHE_ERR_PARSE;
  if (!ScalarL::is_canonical(sig) || !ScalarL::is_canonical(sig + 32)) return XHE_ERR_PARSE;
  for (int i = 0; i < n_ms; i++) if (!ScalarL::is_canonical(ms + 65 * i + 1) || !ScalarL::is_canonical(ms + 65 * i + 33)) return XHE_ERR_PARSE;
  return XHE_OK;
}

void TxView::to_bytes(std::vector<uint8_t>& o, size_t* multisig_index) const {
  o.clear(); o.reserve(len);
  o.push_back(version); put(o, source, 32); put_be64(o, fee); put_be64(o, nonce);
  const uint8_t* p = body;
  switch (type) {
    case 0: for (const TransferView& t : transfers) { put(o, t.asset, 160); if (t.has_extra) put(o, t.extra, t.extra_len); put(o, t.proof, 160); } break;
    case 1: put(o, p, 32); put_be64(o, rd64(p + 32)); break;
    case 2: put(o, p, 32); p += 32; for (uint32_t i = 0; i < count; i++) { put(o, p, 32); put_be64(o, rd64(p + 32)); p += 40; }
      for (uint32_t i = 0; i < aux * 2; i++) { uint32_t l = rd32(p); put(o, p + 4, l); p += 4 + l; } break;
    case 3: put(o, p, aux); break;
    case 4: o.push_back((uint8_t)aux); put(o, p, (size_t)count * 32); break;
  }
  put(o, rp, rp_len);
  put(o, sc, (size_t)n_sc * 256);
  if (multisig_index) *multisig_index = o.size();
  if (n_ms > 0) put(o, ms, (size_t)n_ms * 65);
}

// ------------------------------------------------------------------------------------------------------------------
// batch assembly
// ------------------------------------------------------------------------------------------------------------------
namespace {

const uint32_t OPREF = 0x80000000u;   // point reference to the output of a balance-chain op (fixed up once n_points is known)

enum CheckKind : uint8_t { CK_HOST = 0, CK_POINT = 1, CK_SIG = 2 };
struct Check { CheckKind kind; int32_t err; uint32_t a; };

struct SigEntry { uint32_t tx; bool is_multisig; const uint8_t* sig; uint32_t pk; const uint8_t* pk_enc; };
struct StateUpdate { Bytes32 account, asset; Role role; uint32_t op_c, op_d; bool output = false; /* set_output_ciphertext instead of update_account_balance */ };
inline bool apply_update(VerificationState& state, const uint8_t* account, const uint8_t* asset, Role role, bool output, const uint8_t ct[64]) {
  return output ? state.set_output_ciphertext(account, asset, ct) : state.update_account_balance(account, asset, ct, role);
}
const long long OUT_PREV_C = -1 - XHE_OP_PLUS_AMOUNT, OUT_PREV_D = -1;   // output-ciphertext ops start from the identity (point 0); the commitment half ADDS amount*G
struct Chain { long last_c = -1, last_d = -1; uint32_t length = 0; uint32_t slot = 0xFFFFFFFFu; /* device-resident ledger slot of this key, if any */ };

struct TxPlan {
  uint32_t check_begin = 0, check_end = 0;
  bool proofs = false;                  // true when the tx reached the sigma / range collection stage
  uint32_t eq_begin = 0, val_begin = 0; // first eq / validity proof slot
  int32_t rp_slot = -1;                 // range-proof slot or -1
  uint32_t rp_chal_begin = 0;
  uint32_t sig_begin = 0, sig_end = 0;
  bool rp_structural_fail = false;
};

struct HostCache {   // page-locked staging reused across batches (one per ctx)
  PinnedVec<uint8_t> points, eq_scalars, val_scalars, rp_scalars, rp_challenges, fs_blob, sig_s, sig_e;
  PinnedVec<uint32_t> eq_points, val_points, rp_points, op_terms, fs_plan;
  PinnedVec<uint64_t> fs_off;
};
std::mutex g_cache_mu;
std::unordered_map<xhe_ctx*, HostCache*> g_cache;
HostCache& cache_for(xhe_ctx* ctx) {
  std::lock_guard<std::mutex> g(g_cache_mu);
  HostCache*& c = g_cache[ctx];
  if (!c) c = new HostCache();
  return *c;
}

// 2^64 * G as a ristretto255 encoding (checked against the oracle and libsodium in tests/test_host_cpu.py)
const uint8_t ENC_G_2_64[32] = {0xc8, 0x93, 0xf5, 0x39, 0x19, 0x0e, 0xa3, 0x9d, 0x3b, 0x95, 0x98, 0x4c, 0x6e, 0xcf, 0x75, 0x51,
                                0xbb, 0xd9, 0x2b, 0x3b, 0x7a, 0x63, 0x51, 0x26, 0x14, 0xd6, 0x08, 0x40, 0xea, 0x02, 0x0d, 0x2d};
struct Builder {
  PinnedVec<uint8_t>& points;           // 32 B each; index 0 = identity
  std::vector<Check> checks;
  std::vector<SigEntry> sigs;
  std::vector<long long> op_prev; std::vector<uint32_t> op_term_off; PinnedVec<uint32_t>& op_terms; std::vector<uint64_t> op_amount;
  PinnedVec<uint32_t>&eq_points, &val_points; PinnedVec<uint8_t>&eq_scalars, &val_scalars;
  std::vector<uint32_t> rp_m, rp_point_off, rp_chal_off; PinnedVec<uint32_t>& rp_points; PinnedVec<uint8_t>&rp_scalars, &rp_challenges;
  std::vector<StateUpdate> updates;
  std::unordered_map<Ct64, Chain, KeyHash> chains;
  uint32_t max_chain = 1;
  explicit Builder(HostCache& H) : points(H.points), op_terms(H.op_terms), eq_points(H.eq_points), val_points(H.val_points), eq_scalars(H.eq_scalars), val_scalars(H.val_scalars),
                                   rp_points(H.rp_points), rp_scalars(H.rp_scalars), rp_challenges(H.rp_challenges) {
    points.assign(32, 0); op_terms.clear(); eq_points.clear(); val_points.clear(); eq_scalars.clear(); val_scalars.clear(); rp_points.clear(); rp_scalars.clear(); rp_challenges.clear();
    op_term_off.push_back(0); rp_point_off.push_back(0); rp_chal_off.push_back(0);
  }
  uint32_t add_point(const uint8_t* enc) { uint32_t i = (uint32_t)(points.size() / 32); points.append(enc, 32); return i; }
  uint32_t g64 = 0; uint32_t g_2_64() { if (!g64) g64 = add_point(ENC_G_2_64); return g64; }   // term for plain amounts that reach 2^64
  uint32_t add_op(long long prev, uint64_t amount) { op_prev.push_back(prev); op_amount.push_back(amount); op_term_off.push_back((uint32_t)op_terms.size()); return (uint32_t)op_prev.size() - 1; }
  void close_op() { op_term_off.back() = (uint32_t)op_terms.size(); }
};

struct Rng {   // SHAKE256(seed || tx index) stream for the per-proof random batch factors (reference: Scalar::random)
  Sponge sp;
  Rng(const uint8_t* seed, size_t n, uint64_t idx) : sp(136) { sp.absorb("xhe-batch-factors", 17); sp.absorb(seed, n); sp.absorb(&idx, 8); sp.finish(0x1f); }
  void scalar(uint8_t out[32]) { sp.squeeze(out, 32); out[31] &= 0x0f; }   // uniform in [0, 2^252) -- canonical
};

void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& fn) {
  if (threads <= 1 || n < 2) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) { size_t lo = n * t / threads, hi = n * (t + 1) / threads; if (lo < hi) th.emplace_back(fn, lo, hi, t); }
  for (auto& x : th) x.join();
}

bool has_commitment_for(const TxView& tx, const uint8_t* asset) { for (int i = 0; i < tx.n_sc; i++) if (!memcmp(tx.sc + 256 * i, asset, 32)) return true; return false; }
bool verify_commitment_assets(const TxView& tx) {   // src/tx/verify.rs:161-199
  if (!has_commitment_for(tx, ZERO32)) return false;
  for (int i = 0; i < tx.n_sc; i++) for (int j = 0; j < tx.n_sc; j++) if (i != j && !memcmp(tx.sc + 256 * i, tx.sc + 256 * j, 32)) return false;
  if (tx.type == 0) { for (const TransferView& t : tx.transfers) if (!has_commitment_for(tx, t.asset)) return false; }
  else if (tx.type == 1) return has_commitment_for(tx, tx.body);
  else if (tx.type == 2) { for (uint32_t i = 0; i < tx.count; i++) if (!has_commitment_for(tx, tx.body + 32 + 40 * i)) return false; }
  return true;
}
// the `Scalar::from(..)` parts of get_sender_output_ct, src/tx/verify.rs:107-144.  The reference adds them as Scalars, so
// fee + amount can reach 2^64: the low 64 bits are returned and *carry says whether one 2^64 * G term has to be added
// (at most two u64 summands per asset, so the carry is 0 or 1).
uint64_t plain_output_amount(const TxView& tx, const uint8_t* asset, bool* carry) {
  uint64_t a = 0, b = 0;
  if (is_zero32(asset)) a = tx.fee;
  if (tx.type == 1) { if (!memcmp(asset, tx.body, 32)) b = rd64(tx.body + 32); }
  else if (tx.type == 2) { const uint8_t* hit = nullptr; for (uint32_t i = 0; i < tx.count; i++) if (!memcmp(asset, tx.body + 32 + 40 * i, 32)) hit = tx.body + 32 + 40 * i; if (hit) b = rd64(hit + 32); }
  uint64_t s = a + b;
  *carry = s < a;
  return s;
}

// resolve the (account, asset) balance chain: returns prev references for the commitment / handle ops, registering the
// initial balance points on first touch.  *loaded = point index of the first half if it was read from state now.
Chain* resolve_chain(Builder& B, VerificationState& st, const uint8_t* account, const uint8_t* asset, Role role, long long* prev_c, long long* prev_d, int64_t* loaded) {
  Ct64 key = MockLedger::key(account, asset);
  auto ins = B.chains.try_emplace(key);      // one hash per touch
  Chain& c = ins.first->second;
  *loaded = -1;
  if (ins.second) {
    uint8_t ct[64];
    if (!st.get_account_balance(account, asset, role, ct)) { B.chains.erase(ins.first); return nullptr; }
    uint32_t ic = B.add_point(ct), id = B.add_point(ct + 32);
    c.last_c = -(1 + (long long)ic); c.last_d = -(1 + (long long)id); c.length = 0;
    *loaded = ic;
  }
  *prev_c = c.last_c; *prev_d = c.last_d;
  return &c;
}
inline void advance_chain(Builder& B, Chain* c, uint32_t op_c, uint32_t op_d) {
  c->last_c = op_c; c->last_d = op_d; c->length++;
  if (c->length > B.max_chain) B.max_chain = c->length;
}

// the two ops of get_sender_output_ct(asset) (src/tx/verify.rs:107-144) and the state call that hands the result over
template <typename Idx>
inline void push_output_ops(Builder& B, const TxView& tx, const uint8_t* asset, const Idx& iC, const Idx& iDs, uint32_t k) {
  bool carry; uint64_t amount = plain_output_amount(tx, asset, &carry);
  uint32_t oc = B.add_op(OUT_PREV_C, amount);
  for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) B.op_terms.push_back(iC[t]);
  if (carry) B.op_terms.push_back(B.g_2_64());
  B.close_op();
  uint32_t od = B.add_op(OUT_PREV_D, 0);
  for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) B.op_terms.push_back(iDs[t]);
  B.close_op();
  StateUpdate u; memcpy(u.account.data(), tx.source, 32); memcpy(u.asset.data(), asset, 32); u.role = Sender; u.op_c = oc; u.op_d = od; u.output = true; B.updates.push_back(u);
}

}  // namespace

// Writes to the state that a batch wants to make besides balances: update_account_nonce (src/tx/verify.rs:219-221) and
// set_multisig_for_account (src/tx/verify.rs:426).  They are STAGED while the batch is walked -- later transactions of the
// batch read the multisig settings through the overlay -- and applied together with the balance updates once the batch is
// accepted, so a rejected (or dropped) batch leaves the state untouched.  `foreign` entries come from earlier shards of a
// sharded batch: visible to this shard's transactions, committed by the rank that owns them.
namespace {
struct Staged {
  struct Ms { Bytes32 account; std::vector<uint8_t> signers; uint8_t threshold; bool foreign; };
  std::vector<std::pair<Bytes32, uint64_t>> nonces;
  std::vector<Ms> multisig;
  std::unordered_map<Bytes32, size_t, Key32Hash> ms_index;      // account -> latest entry of `multisig`
  void set_nonce(const uint8_t* account, uint64_t nonce) { Bytes32 k; memcpy(k.data(), account, 32); nonces.emplace_back(k, nonce); }
  void set_multisig(const uint8_t* account, const uint8_t* signers, size_t n, uint8_t threshold, bool foreign) {
    Ms m; memcpy(m.account.data(), account, 32); m.signers.assign(signers, signers + 32 * n); m.threshold = threshold; m.foreign = foreign;
    ms_index[m.account] = multisig.size(); multisig.push_back(std::move(m));
  }
  // get_multisig_for_account through the overlay
  bool get_multisig(VerificationState& state, const uint8_t* account, std::vector<Bytes32>* signers, uint8_t* threshold, bool* present) const {
    if (!ms_index.empty()) {
      Bytes32 k; memcpy(k.data(), account, 32);
      auto it = ms_index.find(k);
      if (it != ms_index.end()) {
        const Ms& m = multisig[it->second]; const size_t n = m.signers.size() / 32;
        *present = n != 0;                       // an empty signer list deletes the setting (src/lib.rs:186-193)
        if (n) { signers->resize(n); for (size_t i = 0; i < n; i++) memcpy((*signers)[i].data(), &m.signers[32 * i], 32); *threshold = m.threshold; }
        return true;
      }
    }
    return state.get_multisig_for_account(account, signers, threshold, present);
  }
  int apply(VerificationState& state) const {
    for (const auto& nn : nonces) if (!state.update_account_nonce(nn.first.data(), nn.second)) return XHE_ERR_STATE;
    for (const Ms& m : multisig) if (!m.foreign && !state.set_multisig_for_account(m.account.data(), m.signers.data(), m.signers.size() / 32, m.threshold)) return XHE_ERR_STATE;
    return XHE_OK;
  }
};
// MultiSig payload rules (src/tx/verify.rs:401-418)
bool multisig_payload_ok(const TxView& tx) {
  const uint32_t ns = tx.count, th = tx.aux;
  if (th > ns || (ns != 0 && th == 0)) return false;
  for (uint32_t x = 0; x < ns; x++) for (uint32_t y = 0; y < ns; y++) if (x != y && !memcmp(tx.body + 32 * x, tx.body + 32 * y, 32)) return false;
  for (uint32_t x = 0; x < ns; x++) if (!memcmp(tx.body + 32 * x, tx.source, 32)) return false;
  return true;
}
// random seed of the batch factors: SHA3-512(personalisation || 32 bytes of OS entropy); the personalisation alone only when
// the caller asked for a replayable run (tests)
bool make_seed(const BatchOptions& opt, uint8_t seed[32]) {
  Sponge sp(72);
  if (opt.rng_seed && opt.rng_seed_len) sp.absorb(opt.rng_seed, opt.rng_seed_len);
  if (!(opt.deterministic_seed && opt.rng_seed && opt.rng_seed_len)) {
    uint8_t os[32]; FILE* f = fopen("/dev/urandom", "rb");
    if (!f || fread(os, 1, 32, f) != 32) { if (f) fclose(f); return false; }
    fclose(f); sp.absorb(os, 32);
  }
  sp.finish(0x06); uint8_t h[64]; sp.squeeze(h, 64); memcpy(seed, h, 32);
  return true;
}

// shard-mode state updates waiting for the cross-rank decision (one slot per ctx)
struct Pending { std::vector<StateUpdate> updates; std::vector<uint8_t> op_out; Staged staged; };
std::mutex g_pending_mu;
std::unordered_map<xhe_ctx*, Pending> g_pending;
}  // namespace

static int apply_pending(const Pending& P, VerificationState& state) {
  int rc = P.staged.apply(state); if (rc) return rc;
  const size_t n = P.updates.size();
  for (size_t j = 0; j < n; j++) {
    const StateUpdate& u = P.updates[j];
    if (j + 8 < n) state.prefetch_balance(P.updates[j + 8].account.data(), P.updates[j + 8].asset.data());
    uint8_t ct[64]; memcpy(ct, &P.op_out[32 * (size_t)u.op_c], 32); memcpy(ct + 32, &P.op_out[32 * (size_t)u.op_d], 32);
    if (!apply_update(state, u.account.data(), u.asset.data(), u.role, u.output, ct)) return XHE_ERR_STATE;
  }
  return XHE_OK;
}
// detach the held-back updates of ctx's last shard-mode batch (so the ctx can take the next batch before the cross-rank
// decision arrives); nullptr if there are none
static Pending* take_pending(xhe_ctx* ctx) {
  std::lock_guard<std::mutex> g(g_pending_mu);          // the lock only covers the hand-over, not the state walk
  auto it = g_pending.find(ctx);
  if (it == g_pending.end()) return nullptr;
  Pending* P = new Pending(std::move(it->second)); g_pending.erase(it);
  return P;
}
int commit_pending(xhe_ctx* ctx, VerificationState& state) {
  Pending* P = take_pending(ctx);
  if (!P) return XHE_E_ARG;
  int rc = apply_pending(*P, state);
  delete P;
  return rc;
}
// the (key, new balance) pairs a shard-mode batch is holding back, as 128-byte records account || asset || ciphertext in
// update order -- what a rank sends to its peers so that every replica of the state ends up identical (distributed.py)
size_t export_pending(void* pending, uint8_t* out, size_t cap) {
  Pending* P = (Pending*)pending; if (!P) return 0;
  size_t k = 0;
  for (const StateUpdate& u : P->updates) {
    if (u.output) continue;
    if ((k + 1) * 128 <= cap) { uint8_t* o = out + 128 * k; memcpy(o, u.account.data(), 32); memcpy(o + 32, u.asset.data(), 32); memcpy(o + 64, &P->op_out[32 * (size_t)u.op_c], 32); memcpy(o + 96, &P->op_out[32 * (size_t)u.op_d], 32); }
    k++;
  }
  return k;
}

// ------------------------------------------------------------------------------------------------------------------
// Transaction::verify_batch
// ------------------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------------------
// fast path: the host reads transaction HEADERS only (shape, source, fee, nonce, asset ids), resolves state and balance
// chains and computes prefix sums; the device builds every table from the uploaded wire blobs (k_layout), replays the
// transcripts and checks the signatures.  Optimistic: returns 1 when the batch is accepted (state applied / partials out),
// 0 when ANYTHING is unusual or failed -- the caller then runs the exact path, which reproduces the reference's verdict and
// error precedence.  Multisig accounts / multisig transactions always take the exact path.
// ------------------------------------------------------------------------------------------------------------------
// Device admission.  Several batches in flight (one context and one host thread each) overlap the host phase of one batch with
// the device phase of others -- but only if the device phases END at different times: batches whose kernels all share the GPU
// at once also finish together, their threads then run their next host phases together while the GPU idles, and the pipeline
// could degenerate into a convoy.  XHE_DEVICE_SLOTS = k admits at most k batches into the device phase at a time, first come
// first served.  Measured on B200 (10 k batches, 6 in flight, 16 host threads): no limit 2.78 M TX/s, k = 3: 2.20 M, k = 2:
// 2.41 M, k = 1: 2.11 M -- the batches do not convoy, the host phases are the limit; the knob stays off (0) by default.
struct DeviceSlots {
  std::mutex mu; std::condition_variable cv; int free_slots; unsigned long long next_ticket = 0, serving = 0;
  DeviceSlots() { const char* e = getenv("XHE_DEVICE_SLOTS"); free_slots = e ? atoi(e) : 0; }
  struct Hold { DeviceSlots* d; ~Hold() { if (d) { std::lock_guard<std::mutex> g(d->mu); d->free_slots++; d->cv.notify_all(); } } };
  Hold acquire() {
    if (limit_off) return Hold{nullptr};
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long t = next_ticket++;
    cv.wait(lk, [&] { return serving == t && free_slots > 0; });
    serving++; free_slots--; cv.notify_all();
    return Hold{this};
  }
  bool limit_off = false;
};
static DeviceSlots& device_slots() { static DeviceSlots* d = [] { DeviceSlots* x = new DeviceSlots(); x->limit_off = x->free_slots <= 0; return x; }(); return *d; }

// (account, asset) -> tail of its balance chain inside the batch, for the fast path's walk.  The walk makes one lookup per key a
// transaction moves, so the table is kept small enough to live in the core's cache: a slot is the key's 32-bit tag and the
// index of a densely packed entry (8 bytes; 512 KB at 30 k keys) and the entries hold POINTERS to the key bytes inside the
// batch (which outlive the walk) instead of copies.  (The generic FlatTable<64, Chain> this replaces was a 16 MB array of
// 128-byte entries: a cache miss per lookup.)
class ChainTable {
 public:
  struct Entry { const uint8_t *account, *asset; Chain c; };
  void reset(size_t keys) { size_t cap = 1024; while (cap < 2 * keys + 16) cap <<= 1; if (cap != slots_.size()) slots_.assign(cap, 0); else std::fill(slots_.begin(), slots_.end(), 0); entries_.clear(); entries_.reserve(keys + 16); }
  static uint64_t digest(const uint8_t* account, const uint8_t* asset) { uint64_t h, g; memcpy(&h, account + 5, 8); memcpy(&g, asset + 5, 8); return (h * 0x9E3779B97F4A7C15ull ^ g) * 0xD6E8FEB86659FD93ull; }
  Chain* insert(const uint8_t* account, const uint8_t* asset, bool* fresh) {
    if (2 * (entries_.size() + 1) > slots_.size()) grow();
    const uint64_t h = digest(account, asset); const size_t mask = slots_.size() - 1; const uint32_t tag = (uint32_t)(h >> 32) | 1u;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
      const uint64_t s = slots_[i];
      if (!s) { slots_[i] = ((uint64_t)tag << 32) | (uint32_t)entries_.size(); entries_.push_back(Entry{account, asset, Chain()}); *fresh = true; return &entries_.back().c; }
      if ((uint32_t)(s >> 32) == tag) { Entry& e = entries_[(uint32_t)s]; if (!memcmp(e.account, account, 32) && !memcmp(e.asset, asset, 32)) { *fresh = false; return &e.c; } }
    }
  }
  void prefetch(const uint8_t* account, const uint8_t* asset) const { if (!slots_.empty()) __builtin_prefetch(&slots_[(size_t)digest(account, asset) & (slots_.size() - 1)]); }
  size_t size() const { return entries_.size(); }
  template <typename F> void for_each(F f) const { for (const Entry& e : entries_) f(e.c); }
 private:
  void grow() { std::vector<uint64_t> ns(std::max<size_t>(1024, 2 * slots_.size()), 0); const size_t mask = ns.size() - 1;
    for (size_t q = 0; q < entries_.size(); q++) { const uint64_t h = digest(entries_[q].account, entries_[q].asset); size_t i = (size_t)h & mask; while (ns[i]) i = (i + 1) & mask; ns[i] = ((uint64_t)((uint32_t)(h >> 32) | 1u) << 32) | (uint32_t)q; }
    slots_.swap(ns); }
  std::vector<uint64_t> slots_; std::vector<Entry> entries_;
};

struct FastCache {
  PinnedVec<uint8_t> blob, region_b, op_out, tx_flags; PinnedVec<uint64_t> off; PinnedVec<uint32_t> plan, terms, term_off, rp_m, rp_pt_off, rp_ch_off; PinnedVec<long long> prev; PinnedVec<uint64_t> amount;
  ChainTable chains;                 // (account, asset) -> tail of its balance chain inside the batch
  std::vector<TxView> txs;           // parsed views of the current batch (reused)
};
static std::unordered_map<xhe_ctx*, FastCache*> g_fast_cache;
static FastCache& fast_cache_for(xhe_ctx* ctx) { std::lock_guard<std::mutex> g(g_cache_mu); FastCache*& c = g_fast_cache[ctx]; if (!c) c = new FastCache(); return *c; }

static int verify_batch_exact(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n, VerificationState& state, const BatchOptions& opt, long* fail_index, BatchTimings* tm);

// The (account, asset) keys an xtx1 blob touches, read straight from its framing without validating it: f(account, asset).
// Used to find, among the EARLIER shards of a sharded batch, the few transactions whose balance chains a shard must follow.
template <class F>
static bool scan_keys(const uint8_t* b, size_t n, F&& f) {
  if (n < 128) return false;
  const uint8_t type = b[1], n_sc = b[2]; const uint32_t count = rd32(b + 4), aux = rd32(b + 8), rp_len = rd32(b + 12);
  size_t off = 64;
  if (type == 0) { for (uint32_t i = 0; i < count; i++) { if (off + 324 > n) return false; f(b + off + 32, b + off); uint32_t el = rd32(b + off + 320); off += 324 + (el == 0xFFFFFFFFu ? 0 : (size_t)el); } }
  else if (type == 1) off += 40;
  else if (type == 2) { off += 32 + 40 * (size_t)count; for (uint32_t i = 0; i < 2 * aux; i++) { if (off + 4 > n) return false; off += 4 + (size_t)rd32(b + off); } }
  else if (type == 3) off += aux;
  else if (type == 4) off += 32 * (size_t)count;
  else return false;
  off += rp_len;
  if (off + 256 * (size_t)n_sc > n) return false;
  for (uint32_t q = 0; q < n_sc; q++) f(b + 16, b + off + 256 * (size_t)q);
  return true;
}
// ---- key-digest index (built once per batch, when its transactions are received: xheh_batch_index_build) ---------------
// digest of an (account, asset) key, and of "account sets its multisig" -- 64-bit, collisions only cost a spurious candidate
static inline uint64_t key_digest(const uint8_t* account, const uint8_t* asset) { Ct64 kk = MockLedger::key(account, asset); return FlatTable<64, uint8_t>::hash(kk.data()) | 1u; }
static inline uint64_t multisig_digest(const uint8_t* source) { return (FlatTable<32, uint8_t>::hash(source) ^ 0x6d756c7469736967ull) | 1u; }
struct BatchIndex { std::vector<uint32_t> off; std::vector<uint64_t> dig; size_t n = 0; };      // digests of transaction i: dig[off[i] .. off[i+1])
static BatchIndex* build_index(const uint8_t* const* blobs, const size_t* lens, size_t n, int threads) {
  BatchIndex* ix = new BatchIndex(); ix->n = n; ix->off.assign(n + 1, 0);
  threads = std::max(1, threads);
  std::vector<uint32_t> cnt(n, 0);
  auto each = [&](size_t i, auto&& f) {
    scan_keys(blobs[i], lens[i], [&](const uint8_t* account, const uint8_t* asset) { f(key_digest(account, asset)); });
    if (lens[i] >= 128 && blobs[i][1] == 4) f(multisig_digest(blobs[i] + 16));
  };
  parallel_for(n, threads, [&](size_t a, size_t b, int) { for (size_t i = a; i < b; i++) { uint32_t c = 0; each(i, [&](uint64_t) { c++; }); cnt[i] = c; } });
  for (size_t i = 0; i < n; i++) ix->off[i + 1] = ix->off[i] + cnt[i];
  ix->dig.resize(ix->off[n]);
  parallel_for(n, threads, [&](size_t a, size_t b, int) { for (size_t i = a; i < b; i++) { uint64_t* d = &ix->dig[ix->off[i]]; each(i, [&](uint64_t h) { *d++ = h; }); } });
  return ix;
}
struct DigestSet {      // open addressing over non-zero 64-bit digests
  std::vector<uint64_t> slot; size_t mask = 0;
  void reserve(size_t n) { size_t cap = 64; while (cap < 4 * n) cap <<= 1; slot.assign(cap, 0); mask = cap - 1; }
  void insert(uint64_t h) { for (size_t i = (size_t)(h >> 17) & mask;; i = (i + 1) & mask) { if (slot[i] == h) return; if (!slot[i]) { slot[i] = h; return; } } }
  bool has(uint64_t h) const { for (size_t i = (size_t)(h >> 17) & mask;; i = (i + 1) & mask) { if (slot[i] == h) return true; if (!slot[i]) return false; } }
};

// The (account, asset) keys a shard moves and its senders.  Their 64-bit digests are always built (one small open-addressing
// set, reused by the calling thread from batch to batch); the exact tables only when an earlier transaction's digest matches --
// in a batch whose shards are independent nothing bigger than that set is allocated or touched.
struct ShardKeys {
  const TxView* txs; size_t n; DigestSet& dig; FlatTable<64, uint8_t> keys; FlatTable<32, uint8_t> sources; bool exact = false;
  static DigestSet& tl_set() { static thread_local DigestSet s; return s; }
  ShardKeys(const TxView* t, size_t count) : txs(t), n(count), dig(tl_set()) {
    size_t nk = 0; for (size_t j = 0; j < n; j++) nk += 1 + txs[j].n_sc + txs[j].transfers.size();
    dig.reserve(nk);
    for (size_t j = 0; j < n; j++) {
      const TxView& tx = txs[j]; dig.insert(multisig_digest(tx.source));
      for (uint32_t q = 0; q < tx.n_sc; q++) dig.insert(key_digest(tx.source, tx.sc + 256 * q));
      for (const TransferView& tr : tx.transfers) dig.insert(key_digest(tr.dest, tr.asset));
    }
  }
  void build_exact() {
    if (exact) return;
    size_t nk = 0; for (size_t j = 0; j < n; j++) nk += txs[j].n_sc + txs[j].transfers.size();
    keys.reserve(nk + 8); sources.reserve(n + 8);
    for (size_t j = 0; j < n; j++) {
      const TxView& tx = txs[j]; sources.insert(tx.source);
      for (uint32_t q = 0; q < tx.n_sc; q++) keys.insert(MockLedger::key(tx.source, tx.sc + 256 * q).data());
      for (const TransferView& tr : tx.transfers) keys.insert(MockLedger::key(tr.dest, tr.asset).data());
    }
    exact = true;
  }
};

// indices i < lo of the transactions that move one of the shard's keys, or set the multisig of one of its senders (in batch
// order).  With the batch's key-digest index the earlier shards' BYTES are not read at all: their digests (a few 8-byte words
// per transaction, sequential) are looked up in the shard's digest set; without it the digests are computed from the blobs'
// framing.  Only candidates are confirmed against the exact keys.  K.exact is set iff the result is non-empty.
static std::vector<size_t> foreign_hits(const uint8_t* const* blobs, const size_t* lens, size_t lo, ShardKeys& K, int threads, const BatchIndex* ix = nullptr) {
  std::vector<std::vector<size_t>> part(std::max(1, threads));
  const DigestSet& mine = K.dig;
  if (ix && ix->n >= lo) {
    parallel_for(lo, threads, [&](size_t a, size_t b, int tid) {
      const uint64_t* d = ix->dig.data();
      for (size_t i = a; i < b; i++) {
        bool cand = false;
        for (uint32_t q = ix->off[i]; q < ix->off[i + 1] && !cand; q++) cand = mine.has(d[q]);
        if (cand) part[tid].push_back(i);
      }
    });
  } else {
    parallel_for(lo, threads, [&](size_t a, size_t b, int tid) {
      for (size_t i = a; i < b; i++) {
        if (i + 12 < b) { const uint8_t* nb = blobs[i + 12]; __builtin_prefetch(nb); __builtin_prefetch(nb + 64); if (lens[i + 12] >= 384) { __builtin_prefetch(nb + lens[i + 12] - 64 - 256); __builtin_prefetch(nb + lens[i + 12] - 64 - 192); } }      // header, first transfer, source commitment
        bool cand = false;
        scan_keys(blobs[i], lens[i], [&](const uint8_t* account, const uint8_t* asset) { if (!cand && mine.has(key_digest(account, asset))) cand = true; });
        if (!cand && lens[i] >= 128 && blobs[i][1] == 4 && mine.has(multisig_digest(blobs[i] + 16))) cand = true;
        if (cand) part[tid].push_back(i);
      }
    });
  }
  std::vector<size_t> out;
  for (auto& p : part) out.insert(out.end(), p.begin(), p.end());
  if (out.empty()) return out;
  K.build_exact();
  std::vector<size_t> hits;
  for (size_t i : out) {
    bool hit = false;
    scan_keys(blobs[i], lens[i], [&](const uint8_t* account, const uint8_t* asset) { if (!hit && K.keys.find(MockLedger::key(account, asset).data())) hit = true; });
    if (!hit && lens[i] >= 128 && blobs[i][1] == 4 && K.sources.find(blobs[i] + 16)) hit = true;
    if (hit) hits.push_back(i);
  }
  std::sort(hits.begin(), hits.end());
  return hits;
}

// the reference's verdict for ONE transaction of the batch (it is known to be the first one that fails a per-transaction
// check): the exact path over the one-transaction shard [i, i+1) -- earlier transactions only advance its balance chains
static int exact_verdict_of(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n, VerificationState& state, const BatchOptions& opt, size_t i) {
  BatchOptions o = opt; o.fast_path = false; o.shard_lo = i; o.shard_hi = i + 1; o.apply_state = false;
  uint8_t partial[64]; o.partial_out = partial;
  long fi = -1;
  int code = verify_batch_exact(ctx, blobs, lens, n, state, o, &fi, nullptr);
  delete take_pending(ctx);                                   // shard mode parks the updates: not wanted here
  return code;
}

// returns 1: accepted (state applied, or partials / pending updates out in shard mode); 2: decided, *rc_out = verdict code and
// *fail_out = failing transaction (index into the whole batch, -1 for the two batch-level checks); 0: not decided here, the
// caller runs the exact path (*rc_out < 0: infrastructure error).
static int verify_batch_fast(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n_total, VerificationState& state, const BatchOptions& opt, BatchTimings* tm, int* rc_out, long* fail_out) {
  double t0 = now_ms();
  *rc_out = XHE_OK; *fail_out = -1;
  const size_t lo = std::min(opt.shard_lo, n_total), hi = std::min(opt.shard_hi, n_total);
  if (hi <= lo) return 0;
  const size_t n = hi - lo;                                   // transactions of this shard; index j <-> batch index lo + j
  int threads = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::thread::hardware_concurrency());
  uint8_t seed[32];
  if (!make_seed(opt, seed)) return 0;
  FastCache& F = fast_cache_for(ctx);
  // this shard's transactions: txs[j] = blobs[lo + j].  The views are kept from batch to batch (their transfer lists keep their
  // capacity: no allocation per transaction after the first batch)
  if (F.txs.size() < n) F.txs.resize(n);
  std::vector<TxView>& txs = F.txs; std::vector<int> parse_rc(n, 0);
  parallel_for(n, threads, [&](size_t a, size_t b, int) { for (size_t j = a; j < b; j++) parse_rc[j] = txs[j].parse(blobs[lo + j], lens[lo + j]); });
  for (size_t j = 0; j < n; j++) if (parse_rc[j]) return 0;
  double t1 = now_ms();
  const uint32_t party_capacity = opt.host_dry_run ? 512u : xhe_ctx_party_capacity(ctx);
  F.off.resize(n + 1); F.plan.resize(8 * n); F.rp_m.resize(n); F.rp_pt_off.resize(n + 1); F.rp_ch_off.resize(n + 1);
  F.region_b.clear(); F.terms.clear(); F.term_off.clear(); F.prev.clear(); F.amount.clear();
  F.term_off.push_back(0);
  ChainTable& chains = F.chains; { size_t nk = 0; for (size_t j = 0; j < n; j++) nk += txs[j].n_sc + txs[j].transfers.size(); chains.reset(nk); }
  struct Upd { const uint8_t *account, *asset; Role role; uint32_t op_c; bool output; };
  std::vector<Upd> updates; updates.reserve(3 * n);
  Staged staged; staged.nonces.reserve(n);
  const bool want_out = state.wants_output_ciphertexts(); std::vector<size_t> out_slots;
  const long long RB = (long long)1 << 40;     // marks "region B slot j" until the region-A size is known
  uint32_t pt = 1, n_eq = 0, n_val = 0, max_chain = 1; uint64_t off = 0;
  F.rp_pt_off[0] = 0; F.rp_ch_off[0] = 0;
  std::vector<Bytes32> signers;
  std::vector<size_t> rb_terms;                // positions in F.terms that hold a region-B slot (foreign ciphertext points)
  // a state whose balances live on this context's device: chains start from ledger slots, nothing is uploaded
  // or decompressed for them, and the accepted outputs go back into the slots on the device.  (Shard mode keeps the compressed
  // route: its updates wait for the cross-rank decision while the context moves on to the next batch.)
  xhe_ledger* dl = opt.partial_out ? nullptr : state.device_ledger();
  auto touch = [&](const uint8_t* account, const uint8_t* asset, Role role, long long* pc, long long* pd) -> Chain* {
    bool fresh = false;
    Chain& c = *chains.insert(account, asset, &fresh);
    if (fresh && dl) {
      uint32_t slot;
      if (!state.device_slot(account, asset, &slot)) return nullptr;
      c.last_c = -(1 + 2 * (long long)slot) - XHE_OP_FROM_LEDGER; c.last_d = -(1 + 2 * (long long)slot + 1) - XHE_OP_FROM_LEDGER; c.length = 0; c.slot = slot;
    } else if (fresh) {
      uint8_t ct[64];
      if (!state.get_account_balance(account, asset, role, ct)) return nullptr;
      long long j = (long long)(F.region_b.size() / 32); F.region_b.append(ct, 64);
      c.last_c = -(RB + j); c.last_d = -(RB + j + 1); c.length = 0;
    }
    *pc = c.last_c; *pd = c.last_d; return &c;
  };
  auto push_op = [&](long long prev, uint64_t amount) { F.prev.push_back(prev); F.amount.push_back(amount); return (uint32_t)F.prev.size() - 1; };
  auto rb_term = [&](const uint8_t* enc, bool neg) { uint32_t j = (uint32_t)(F.region_b.size() / 32); F.region_b.append(enc, 32); rb_terms.push_back(F.terms.size()); F.terms.push_back(j | (neg ? 0x80000000u : 0u)); };
  // ---- earlier shards of a sharded batch: follow the balance chains (and multisig settings) this shard's transactions depend
  // on.  The group operations of a foreign transaction on a shared (account, asset) are replayed here (no proofs: its own
  // rank verifies them); its ciphertext points travel in region B.
  if (lo > 0) {
    ShardKeys K(txs.data(), n); FlatTable<64, uint8_t>& keys = K.keys; FlatTable<32, uint8_t>& sources = K.sources;      // (the exact tables are filled iff there is a hit)
    TxView ftx;
    for (size_t i : foreign_hits(blobs, lens, lo, K, threads, (opt.key_index && ((const BatchIndex*)opt.key_index)->n == n_total) ? (const BatchIndex*)opt.key_index : nullptr)) {
      if (ftx.parse(blobs[i], lens[i])) continue;
      const TxView& tx = ftx; const uint32_t k = tx.n_transfers();
      if (tx.type == 4 && sources.find(tx.source)) return 0;                    // multisig setting for one of our senders: exact path
      for (uint32_t q = 0; q < tx.n_sc; q++) {
        const uint8_t* asset = tx.sc + 256 * q;
        if (!keys.find(MockLedger::key(tx.source, asset).data())) continue;
        long long pc, pd; Chain* ch = touch(tx.source, asset, Sender, &pc, &pd); if (!ch) return 0;
        bool carry; const uint64_t amount = plain_output_amount(tx, asset, &carry); if (carry) return 0;
        uint32_t oc = push_op(pc, amount);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) rb_term(tx.transfers[t].commitment, true);
        F.term_off.push_back((uint32_t)F.terms.size());
        uint32_t od = push_op(pd, 0);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) rb_term(tx.transfers[t].sender_handle, true);
        F.term_off.push_back((uint32_t)F.terms.size());
        ch->last_c = oc; ch->last_d = od; if (++ch->length > max_chain) max_chain = ch->length;
      }
      for (uint32_t t = 0; t < k; t++) {
        const TransferView& tr = tx.transfers[t];
        if (!keys.find(MockLedger::key(tr.dest, tr.asset).data())) continue;
        long long pc, pd; Chain* ch = touch(tr.dest, tr.asset, Receiver, &pc, &pd); if (!ch) return 0;
        uint32_t oc = push_op(pc, 0); rb_term(tr.commitment, false); F.term_off.push_back((uint32_t)F.terms.size());
        uint32_t od = push_op(pd, 0); rb_term(tr.receiver_handle, false); F.term_off.push_back((uint32_t)F.terms.size());
        ch->last_c = oc; ch->last_d = od; if (++ch->length > max_chain) max_chain = ch->length;
      }
    }
  }
  // the walk below is one dependent hash lookup after another: announce the lookups a few transactions ahead
  const size_t AHEAD = 12;
  auto announce = [&](const TxView& tx) {
    state.prefetch_account(tx.source);
    for (uint32_t q = 0; q < tx.n_sc; q++) { state.prefetch_balance(tx.source, tx.sc + 256 * q); chains.prefetch(tx.source, tx.sc + 256 * q); }
    for (uint32_t t = 0; t < tx.n_transfers(); t++) { const TransferView& tr = tx.transfers[t]; state.prefetch_balance(tr.dest, tr.asset); chains.prefetch(tr.dest, tr.asset); }
  };
  for (size_t j = 0; j < n && j < AHEAD; j++) announce(txs[j]);
  // A transaction that fails one of the HOST-side checks ends the walk (n_run transactions go to the device): the reference
  // stops at the first failing transaction, so whatever comes after it cannot change the verdict.
  size_t n_run = n; bool host_fail = false;
  for (size_t j = 0; j < n; j++) {
    const TxView& tx = txs[j];
    if (j + AHEAD < n) announce(txs[j + AHEAD]);
    if (tx.type == 4 || tx.n_ms >= 0) return 0;                               // multisig: exact path
    uint64_t nonce;
    if (!state.get_account_nonce(tx.source, &nonce) || nonce != tx.nonce || !verify_commitment_assets(tx)) { n_run = j; host_fail = true; break; }
    { uint8_t th; bool present = false; if (!staged.get_multisig(state, tx.source, &signers, &th, &present) || present) return 0; }
    const uint32_t k = tx.n_transfers(), a = tx.n_sc, lg = (tx.rp_len / 32 - 9) / 2;
    uint32_t m = 1, lg_need = 6; while (m < a + k) { m <<= 1; lg_need++; }
    if (lg != lg_need || m > party_capacity) return 0;            // structural range-proof failure / larger proofs: the exact path decides
    const size_t n_prev0 = F.prev.size(), n_terms0 = F.terms.size(), n_rb0 = F.region_b.size(), n_upd0 = updates.size();
    uint32_t* P = &F.plan[8 * j];
    P[0] = n_eq; P[1] = n_val; P[2] = (uint32_t)j; P[3] = F.rp_ch_off[j]; P[4] = (uint32_t)j; P[5] = 1; P[6] = pt; P[7] = (uint32_t)F.prev.size();
    F.rp_m[j] = m; F.rp_pt_off[j + 1] = F.rp_pt_off[j] + 4 + 2 * lg + m; F.rp_ch_off[j + 1] = F.rp_ch_off[j] + 4 + lg;
    const uint32_t iT = pt + 1;
    bool state_fail = false;
    for (uint32_t q = 0; q < a; q++) {
      const uint8_t* asset = tx.sc + 256 * q; long long pc, pd;
      Chain* ch = touch(tx.source, asset, Sender, &pc, &pd); if (!ch) { state_fail = true; break; }
      bool carry; const uint64_t amount = plain_output_amount(tx, asset, &carry);
      if (carry) return 0;                                                      // fee + amount >= 2^64: exact path (adds the 2^64 * G term)
      uint32_t oc = push_op(pc, amount);
      for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) F.terms.push_back((iT + 3 * t) | 0x80000000u);
      F.term_off.push_back((uint32_t)F.terms.size());
      uint32_t od = push_op(pd, 0);
      for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) F.terms.push_back((iT + 3 * t + 1) | 0x80000000u);
      F.term_off.push_back((uint32_t)F.terms.size());
      ch->last_c = oc; ch->last_d = od; if (++ch->length > max_chain) max_chain = ch->length;
      updates.push_back({tx.source, asset, Sender, oc, false});
      if (want_out) { out_slots.push_back(updates.size()); updates.push_back({tx.source, asset, Sender, 0, true}); }   // reference order: update, then set_output (339-340)
    }
    for (uint32_t t = 0; t < k && !state_fail; t++) {
      const TransferView& tr = tx.transfers[t]; long long pc, pd;
      Chain* ch = touch(tr.dest, tr.asset, Receiver, &pc, &pd); if (!ch) { state_fail = true; break; }
      uint32_t oc = push_op(pc, 0); F.terms.push_back(iT + 3 * t); F.term_off.push_back((uint32_t)F.terms.size());
      uint32_t od = push_op(pd, 0); F.terms.push_back(iT + 3 * t + 2); F.term_off.push_back((uint32_t)F.terms.size());
      ch->last_c = oc; ch->last_d = od; if (++ch->length > max_chain) max_chain = ch->length;
      updates.push_back({tr.dest, tr.asset, Receiver, oc, false});
    }
    if (state_fail) {   // a balance lookup failed: this transaction fails (the exact path names the error)
      if (F.prev.size() != n_prev0 || updates.size() != n_upd0) return 0;      // ... but it had already advanced a chain: let the exact path sort it out
      (void)n_terms0; (void)n_rb0;
      n_run = j; host_fail = true; break;
    }
    if (want_out) {   // output ciphertexts: appended after the ops k_layout indexes (P[7] + 2q are the sender ops)
      for (uint32_t q = 0; q < a; q++) {
        const uint8_t* asset = tx.sc + 256 * q;
        bool carry; uint32_t oc = push_op(OUT_PREV_C, plain_output_amount(tx, asset, &carry));   // carry == false: checked at the sender op above
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) F.terms.push_back(iT + 3 * t);
        F.term_off.push_back((uint32_t)F.terms.size());
        push_op(OUT_PREV_D, 0);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) F.terms.push_back(iT + 3 * t + 1);
        F.term_off.push_back((uint32_t)F.terms.size());
        updates[out_slots[q]].op_c = oc;
      }
      out_slots.clear();
    }
    staged.set_nonce(tx.source, tx.nonce);                                      // update_account_nonce(source, nonce), src/tx/verify.rs:219-221
    F.off[j] = off; off += (lens[lo + j] + 15) & ~(size_t)15;
    pt += 1 + 3 * k + a + 3 * a + k + 3 * k + 4 + 2 * lg;
    n_eq += a; n_val += k;
  }
  const bool shard = opt.partial_out != nullptr;
  if (n_run == 0) {   // the first transaction of the shard already fails a host-side check
    int code = exact_verdict_of(ctx, blobs, lens, n_total, state, opt, lo);
    if (code <= 0) { *rc_out = code < 0 ? code : XHE_OK; return 0; }
    if (shard) { memset(opt.partial_out, 0, 64); std::lock_guard<std::mutex> g(g_pending_mu); g_pending[ctx] = Pending(); }
    *rc_out = code; *fail_out = (long)lo; return 2;
  }
  F.off[n_run] = off;
  const uint32_t n_a_end = pt, n_rb = (uint32_t)(F.region_b.size() / 32), n_points = n_a_end + n_rb;
  for (size_t q = 0; q < F.prev.size(); q++) if (F.prev[q] <= -RB && F.prev[q] > -XHE_OP_FROM_LEDGER) F.prev[q] = -(1 + (long long)n_a_end + (-F.prev[q] - RB));
  for (size_t pos : rb_terms) F.terms[pos] = (F.terms[pos] & 0x80000000u) | (n_a_end + (F.terms[pos] & 0x7fffffffu));
  double t2 = now_ms();
  // Zero-copy upload: when the caller's transactions already sit back to back (16-byte aligned, the layout
  // the device reads) in ONE page-locked buffer -- xheh_blob_arena_* builds such a buffer -- the DMA reads them where they
  // are; otherwise they are gathered into this context's pinned staging buffer first.
  const uint8_t* blob_base = nullptr;
  {
    bool contiguous = (((uintptr_t)blobs[lo]) & 15) == 0;
    for (size_t j = 0; contiguous && j + 1 < n_run; j++) contiguous = blobs[lo + j + 1] == blobs[lo + j] + ((lens[lo + j] + 15) & ~(size_t)15);
    if (contiguous) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, blobs[lo]) == cudaSuccess && at.type == cudaMemoryTypeHost) blob_base = blobs[lo]; else cudaGetLastError();
    }
  }
  if (!blob_base) {
    F.blob.resize(off);
    parallel_for(n_run, threads, [&](size_t a, size_t b, int) { for (size_t j = a; j < b; j++) memcpy(&F.blob[F.off[j]], blobs[lo + j], lens[lo + j]); });
    blob_base = F.blob.data();
  }
  double t3 = now_ms();
  if (opt.host_dry_run) {   // diagnostics (host-phase profiling without a device): nothing is verified, the call reports an error code
    if (tm) { tm->parse_ms = t1 - t0; tm->resolve_ms = t2 - t1; tm->transcript_ms = t3 - t2; tm->total_ms = t3 - t0; tm->used_fast_path = true; }
    *rc_out = XHE_E_ARG; return 0;
  }
  xhe_batch xb; memset(&xb, 0, sizeof xb); xb.struct_size = (uint32_t)sizeof xb;
  xb.n_tx = (uint32_t)n_run; xb.n_points = n_points; xb.n_sigs = (uint32_t)n_run;
  xb.n_ops = (uint32_t)F.prev.size(); xb.op_prev = (const int64_t*)F.prev.data(); xb.op_term_off = F.term_off.data(); xb.op_terms = F.terms.data(); xb.op_amount = F.amount.data(); xb.max_chain = max_chain;
  xb.n_eq = n_eq; xb.n_val = n_val; xb.n_rp = (uint32_t)n_run; xb.rp_m = F.rp_m.data(); xb.rp_point_off = F.rp_pt_off.data(); xb.rp_chal_off = F.rp_ch_off.data();
  xb.fs_blobs = blob_base; xb.fs_blob_off = F.off.data(); xb.fs_plan = F.plan.data(); memcpy(xb.fs_seed, seed, 32); xb.fs_index_base = lo;
  xb.layout_on_device = 1; xb.n_region_b = n_rb; xb.region_b = F.region_b.data(); xb.ledger = dl;
  PinnedVec<uint8_t>& op_out = F.op_out; op_out.n = 0; op_out.reserve(32 * (size_t)xb.n_ops + 64); op_out.n = 32 * (size_t)xb.n_ops;
  PinnedVec<uint8_t>& txf = F.tx_flags; txf.n = 0; txf.reserve(n_run + 64); txf.n = n_run;
  xhe_verdict v; memset(&v, 0, sizeof v); v.struct_size = (uint32_t)sizeof v; v.op_out = (dl && !want_out) ? nullptr : op_out.data(); v.tx_flags = txf.data();      // resident ledger: no balance comes back
  int32_t rc; { DeviceSlots::Hold slot = device_slots().acquire(); rc = xhe_verify_batch(ctx, &xb, &v); }
  double t4 = now_ms();
  if (tm) { tm->parse_ms = t1 - t0; tm->resolve_ms = t2 - t1; tm->transcript_ms = t3 - t2; tm->device_ms = t4 - t3; tm->total_ms = t4 - t0; }
  if (rc != XHE_OK) { *rc_out = rc; return 0; }
  // ---- verdict.  Per-transaction anomalies first (the first failing transaction in batch order, src/tx/verify.rs:492-498),
  // then the two batch-level checks.  The failing transaction's error is the exact path's verdict on that ONE transaction.
  long first_bad = -1;
  if (v.device_flags & 8u) return 0;                                            // a balance read from the state does not decode: exact path
  if (v.device_flags & 7u) { for (size_t j = 0; j < n_run; j++) if (txf[j] & 7u) { first_bad = (long)j; break; } if (first_bad < 0) return 0; }
  if (first_bad < 0 && host_fail) first_bad = (long)n_run;
  if (first_bad >= 0) {
    int code = exact_verdict_of(ctx, blobs, lens, n_total, state, opt, lo + (size_t)first_bad);
    if (code <= 0) { *rc_out = code < 0 ? code : XHE_OK; return 0; }            // the exact path disagrees: let it decide the whole batch
    if (shard) { memcpy(opt.partial_out, v.sigma_enc, 32); memcpy(opt.partial_out + 32, v.range_enc, 32); std::lock_guard<std::mutex> g(g_pending_mu); g_pending[ctx] = Pending(); }
    *rc_out = code; *fail_out = (long)lo + first_bad;
    if (tm) { tm->used_fast_path = true; tm->finish_ms = now_ms() - t4; tm->total_ms = now_ms() - t0; }
    return 2;
  }
  const bool rp_structural = (v.device_flags & 16u) != 0;      // an identity-encoded range-proof point: RangeProof, but only after the sigma check
  if (!shard) {
    int code = !v.sigma_is_identity ? XHE_ERR_GENERIC_PROOF : ((rp_structural || !v.range_is_identity) ? XHE_ERR_RANGE_PROOF : XHE_OK);      // src/tx/verify.rs:500-502, 504-514
    if (code) { *rc_out = code; *fail_out = -1; if (tm) { tm->used_fast_path = true; tm->finish_ms = now_ms() - t4; tm->total_ms = now_ms() - t0; } return 2; }
  } else if (rp_structural) {   // the joint decision (distributed.decide) places a shard's structural failure after the summed sigma check
    memcpy(opt.partial_out, v.sigma_enc, 32); memcpy(opt.partial_out + 32, v.range_enc, 32);
    { std::lock_guard<std::mutex> g(g_pending_mu); g_pending[ctx] = Pending(); }
    *rc_out = XHE_ERR_RANGE_PROOF; *fail_out = -1; return 2;
  }
  if (shard) {
    memcpy(opt.partial_out, v.sigma_enc, 32); memcpy(opt.partial_out + 32, v.range_enc, 32);
    Pending Pn; Pn.op_out.assign(op_out.data(), op_out.data() + op_out.size()); Pn.updates.reserve(updates.size()); Pn.staged = std::move(staged);
    for (const Upd& u : updates) { StateUpdate su; memcpy(su.account.data(), u.account, 32); memcpy(su.asset.data(), u.asset, 32); su.role = u.role; su.op_c = u.op_c; su.op_d = u.op_c + 1; su.output = u.output; Pn.updates.push_back(su); }
    std::lock_guard<std::mutex> g(g_pending_mu);
    g_pending[ctx] = std::move(Pn);
  } else if (opt.apply_state && dl) {
    // update_account_balance on the device: the LAST op of every chain is that key's new balance (src/tx/verify.rs:329-336,367-374)
    if (staged.apply(state) != XHE_OK) { *rc_out = XHE_ERR_STATE; *fail_out = -1; return 2; }
    std::vector<uint32_t> slots, ops; slots.reserve(chains.size()); ops.reserve(chains.size());
    chains.for_each([&](const Chain& c) { if (c.slot != 0xFFFFFFFFu && c.last_c >= 0) { slots.push_back(c.slot); ops.push_back((uint32_t)c.last_c); } });
    int32_t crc = xhe_ledger_commit_batch(dl, ctx, slots.data(), ops.data(), slots.size());
    if (crc != XHE_OK) { *rc_out = crc; *fail_out = -1; return crc < 0 ? 0 : 2; }
    if (want_out) for (const Upd& u : updates) if (u.output) { uint8_t ct[64]; memcpy(ct, &op_out[32 * (size_t)u.op_c], 64); if (!state.set_output_ciphertext(u.account, u.asset, ct)) { *rc_out = XHE_ERR_STATE; *fail_out = -1; return 2; } }
  } else if (opt.apply_state) {
    if (staged.apply(state) != XHE_OK) { *rc_out = XHE_ERR_STATE; *fail_out = -1; return 2; }
    for (size_t j = 0; j < updates.size(); j++) {
      const Upd& u = updates[j];
      if (j + 8 < updates.size()) state.prefetch_balance(updates[j + 8].account, updates[j + 8].asset);
      uint8_t ct[64]; memcpy(ct, &op_out[32 * (size_t)u.op_c], 64);       // commitment op and handle op are adjacent
      if (!apply_update(state, u.account, u.asset, u.role, u.output, ct)) { *rc_out = XHE_ERR_STATE; *fail_out = -1; return 2; }
    }
  }
  double t5 = now_ms();
  if (tm) { tm->used_fast_path = true; tm->parse_ms = t1 - t0; tm->resolve_ms = t2 - t1; tm->transcript_ms = t3 - t2; tm->device_ms = t4 - t3; tm->finish_ms = t5 - t4; tm->total_ms = t5 - t0; tm->keccak_f = 0; }
  return 1;
}

// Where should the Merlin transcripts of this batch run?  On the device every transaction's transcript is ONE thread's sequential
// chain of Keccak permutations (about 16 us each at one warp per sub-partition); on the host a permutation takes well under a
// microsecond and transactions spread over the threads.  Thousands of ordinary transactions hide the device latency (10 k
// a1k1: 24 permutations each, 0.4 ms); a handful of transactions with hundreds of transfers each -- the reference's
// 16 x 255-transfer bench, benches/tx.rs:231-233: ~1,300 permutations per transaction -- do not (21 ms against < 1 ms).
static bool transcripts_favor_host(const uint8_t* const* blobs, const size_t* lens, size_t lo, size_t hi, int threads) {
  double max_p = 0, sum_p = 0;
  for (size_t i = lo; i < hi; i++) {
    if (lens[i] < 128) continue;
    const uint8_t* b = blobs[i]; const double k = b[1] == 0 ? (double)rd32(b + 4) : 0.0, a = b[2];
    const double p = 10.0 + 14.0 * a + 5.5 * k + 16.0;       // header, per source commitment, per transfer, range proof
    sum_p += p; if (p > max_p) max_p = p;
  }
  const double device_us = 16.0 * max_p, host_us = 0.7 * sum_p / std::max(1, threads);
  return device_us > 3000.0 && device_us > 2.0 * host_us;
}

int verify_batch(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n, VerificationState& state, const BatchOptions& opt, long* fail_index, BatchTimings* tm) {
  if (opt.fast_path && !opt.host_dry_run) {
    const size_t lo = std::min(opt.shard_lo, n), hi = std::min(opt.shard_hi, n);
    const int threads = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (transcripts_favor_host(blobs, lens, lo, hi, threads)) {      // few, very long transcripts: north_star's split (host Merlin, exact path)
      BatchOptions o = opt; o.fast_path = false; o.device_fiat_shamir = false;
      return verify_batch_exact(ctx, blobs, lens, n, state, o, fail_index, tm);
    }
  }
  if (opt.fast_path) {
    int rc = XHE_OK; long fi = -1;
    int how = verify_batch_fast(ctx, blobs, lens, n, state, opt, tm, &rc, &fi);
    if (how == 1) { if (fail_index) *fail_index = -1; return XHE_OK; }
    if (how == 2) { if (fail_index) *fail_index = fi; return rc; }
    if (rc != XHE_OK) { if (fail_index) *fail_index = -1; return rc; }       // infrastructure error: report it
  }
  return verify_batch_exact(ctx, blobs, lens, n, state, opt, fail_index, tm);
}

static int verify_batch_exact(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n_total, VerificationState& state, const BatchOptions& opt, long* fail_index, BatchTimings* tm) {
  double t0 = now_ms();
  if (fail_index) *fail_index = -1;
  const bool want_out = state.wants_output_ciphertexts();
  int threads = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::thread::hardware_concurrency());
  uint8_t seed[32];
  if (!make_seed(opt, seed)) return XHE_E_ARG;
  // this call verifies transactions [lo, hi) of the batch (everything unless shard_lo / shard_hi say otherwise); batch index
  // i <-> local index i - lo in `plan` and in the device batch
  const size_t lo = std::min(opt.shard_lo, n_total), hi = std::min(opt.shard_hi, n_total);
  const uint32_t party_capacity = xhe_ctx_party_capacity(ctx);

  // ---- parse (parallel)
  const size_t nsh = hi - lo;
  std::vector<TxView> txs(nsh); std::vector<int> parse_rc(nsh, 0);  // this shard's transactions: txs[i - lo] = blobs[i]
  parallel_for(nsh, threads, [&](size_t a, size_t b, int) { for (size_t j = a; j < b; j++) parse_rc[j] = txs[j].parse(blobs[lo + j], lens[lo + j]); });
  size_t n_live = hi; int parse_err = XHE_OK;
  for (size_t i = lo; i < hi; i++) if (parse_rc[i - lo]) { n_live = i; parse_err = parse_rc[i - lo]; break; }
  double t1 = now_ms();

  // ---- phase A: sequential state resolution and batch layout (mirrors pre_verify's order, src/tx/verify.rs:203-485)
  HostCache& HC = cache_for(ctx);
  Builder B(HC); std::vector<TxPlan> plan(n_live > lo ? n_live - lo : 0);
  Staged staged;
  {
    size_t tk = 0, ta = 0, tlg = 0;
    for (size_t i = lo; i < n_live; i++) { tk += txs[i - lo].n_transfers(); ta += txs[i - lo].n_sc; tlg += (txs[i - lo].rp_len / 32 - 9) / 2; }
    size_t nl = n_live - std::min(lo, n_live), npts = 1 + nl * 5 + tk * 9 + ta * 6 + 2 * tlg + 8;
    B.points.reserve(32 * npts); B.checks.reserve(npts + 2 * nl); B.sigs.reserve(nl + 8);
    B.op_prev.reserve(2 * (ta + tk)); B.op_amount.reserve(2 * (ta + tk)); B.op_term_off.reserve(2 * (ta + tk) + 1); B.op_terms.reserve(2 * (ta + 2 * tk));
    B.eq_points.reserve(7 * ta); B.eq_scalars.reserve(192 * ta); B.val_points.reserve(8 * tk); B.val_scalars.reserve(160 * tk);
    B.rp_m.reserve(nl); B.rp_point_off.reserve(nl + 1); B.rp_chal_off.reserve(nl + 1); B.rp_points.reserve(4 * nl + 2 * tlg + 2 * (ta + tk)); B.rp_scalars.reserve(224 * nl);
    B.rp_challenges.reserve(32 * (4 * nl + tlg)); B.updates.reserve(ta + tk); B.chains.reserve(2 * (ta + tk));
  }
  // ---- earlier shards of a sharded batch: an (account, asset) balance this shard reads may have been moved by
  // transactions [0, lo) -- a sender with transactions in both shards, a receiver credited there and spending here
  // (src/lib.rs:908-921), two shards crediting one receiver.  Their GROUP operations on the shared keys are replayed here in
  // batch order as ordinary balance-chain ops (no proofs: the owning rank verifies those), so this shard's proofs meet the
  // same ciphertexts as in the reference's sequential walk (src/tx/verify.rs:301-336,354-374).  MultiSig settings of earlier
  // shards are visible through the overlay.  If a foreign transaction is invalid its own rank rejects the batch.
  if (lo > 0 && n_live > lo) {
    ShardKeys K(txs.data(), n_live - lo); FlatTable<64, uint8_t>& keys = K.keys; FlatTable<32, uint8_t>& sources = K.sources;      // (the exact tables are filled iff there is a hit)
    // the earlier transactions are only SCANNED for these keys (scan_keys reads the framing, nothing else); the few that touch
    // one are parsed.  The points of a foreign transaction must outlive this loop: they are copied into the point table.
    TxView ftx;
    for (size_t i : foreign_hits(blobs, lens, lo, K, threads, (opt.key_index && ((const BatchIndex*)opt.key_index)->n == n_total) ? (const BatchIndex*)opt.key_index : nullptr)) {
      if (ftx.parse(blobs[i], lens[i])) continue;
      const TxView& tx = ftx; const uint32_t k = tx.n_transfers();
      if (tx.type == 4 && sources.find(tx.source) && multisig_payload_ok(tx)) staged.set_multisig(tx.source, tx.body, tx.count, (uint8_t)tx.aux, true);
      for (uint32_t q = 0; q < tx.n_sc; q++) {
        const uint8_t* asset = tx.sc + 256 * q;
        if (!keys.find(MockLedger::key(tx.source, asset).data())) continue;
        long long pc, pd; int64_t loaded; Chain* ch = resolve_chain(B, state, tx.source, asset, Sender, &pc, &pd, &loaded); if (!ch) continue;
        bool carry; const uint64_t amount = plain_output_amount(tx, asset, &carry);
        uint32_t oc = B.add_op(pc, amount);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) { uint32_t ip = B.add_point(tx.transfers[t].commitment); B.op_terms.push_back(ip | 0x80000000u); }
        if (carry) B.op_terms.push_back(B.g_2_64() | 0x80000000u);
        B.close_op();
        uint32_t od = B.add_op(pd, 0);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) { uint32_t ip = B.add_point(tx.transfers[t].sender_handle); B.op_terms.push_back(ip | 0x80000000u); }
        B.close_op();
        advance_chain(B, ch, oc, od);
      }
      for (uint32_t t = 0; t < k; t++) {
        const TransferView& tr = tx.transfers[t];
        if (!keys.find(MockLedger::key(tr.dest, tr.asset).data())) continue;
        long long pc, pd; int64_t loaded; Chain* ch = resolve_chain(B, state, tr.dest, tr.asset, Receiver, &pc, &pd, &loaded); if (!ch) continue;
        uint32_t oc = B.add_op(pc, 0); { uint32_t ip = B.add_point(tr.commitment); B.op_terms.push_back(ip); } B.close_op();
        uint32_t od = B.add_op(pd, 0); { uint32_t ip = B.add_point(tr.receiver_handle); B.op_terms.push_back(ip); } B.close_op();
        advance_chain(B, ch, oc, od);
      }
    }
  }
  std::vector<uint32_t> iC, iDs, iDr, iN, Ls, Rs; std::vector<Bytes32> signers;
  size_t n_reached = n_live;    // txs after the first host-side hard error are never reached by the reference
  long capacity_fail = -1;
  for (size_t i = lo; i < n_live; i++) {
    const TxView& tx = txs[i - lo]; TxPlan& P = plan[i - lo];
    P.check_begin = (uint32_t)B.checks.size(); P.sig_begin = (uint32_t)B.sigs.size();
    auto host_fail = [&](int err) { B.checks.push_back({CK_HOST, err, 0}); };
    auto finish = [&]() { P.check_end = (uint32_t)B.checks.size(); P.sig_end = (uint32_t)B.sigs.size(); };
    bool stop = false;
    do {
      uint64_t nonce;
      if (!state.get_account_nonce(tx.source, &nonce)) { host_fail(XHE_ERR_STATE); stop = true; break; }
      if (nonce != tx.nonce) { host_fail(XHE_ERR_INVALID_NONCE); stop = true; break; }
      staged.set_nonce(tx.source, tx.nonce);                     // update_account_nonce(source, nonce): applied with the batch (src/tx/verify.rs:219-221)
      if (!verify_commitment_assets(tx)) { host_fail(XHE_ERR_FORMAT); stop = true; break; }
      const uint32_t k = tx.n_transfers(), a = tx.n_sc;
      iC.resize(k); iDs.resize(k); iDr.resize(k); iN.resize(a);
      for (uint32_t t = 0; t < k; t++) {
        iC[t] = B.add_point(tx.transfers[t].commitment); iDs[t] = B.add_point(tx.transfers[t].sender_handle); iDr[t] = B.add_point(tx.transfers[t].receiver_handle);
        B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, iC[t]}); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, iDs[t]}); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, iDr[t]});
      }
      for (uint32_t q = 0; q < a; q++) { iN[q] = B.add_point(tx.sc + 256 * q + 32); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, iN[q]}); }
      uint32_t iSrc = B.add_point(tx.source); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, iSrc});
      // 0. signature (src/tx/verify.rs:253-256)
      B.checks.push_back({CK_SIG, XHE_ERR_SIGNATURE, (uint32_t)B.sigs.size()});
      B.sigs.push_back({(uint32_t)i, false, tx.sig, iSrc, tx.source});
      // multisig rules (259-292)
      {
        uint8_t threshold = 0; bool present = false;
        if (!staged.get_multisig(state, tx.source, &signers, &threshold, &present)) { host_fail(XHE_ERR_STATE); stop = true; break; }
        if (present) {
          if (tx.n_ms < 0) { host_fail(XHE_ERR_FORMAT); stop = true; break; }
          if (tx.n_ms == 0 || tx.n_ms != (int)threshold) { host_fail(XHE_ERR_FORMAT); stop = true; break; }
          for (int s = 0; s < tx.n_ms && !stop; s++) {
            for (int s2 = 0; s2 < tx.n_ms; s2++) if (s != s2 && tx.ms[65 * s] == tx.ms[65 * s2]) { host_fail(XHE_ERR_FORMAT); stop = true; break; }
            if (stop) break;
            uint32_t idx = tx.ms[65 * s];
            if (idx < signers.size()) {
              // the signer key must outlive this call: copy it into the point table and point at that copy
              uint32_t ip = B.add_point(signers[idx].data());
              B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, ip});
              B.checks.push_back({CK_SIG, XHE_ERR_SIGNATURE, (uint32_t)B.sigs.size()});
              B.sigs.push_back({(uint32_t)i, true, tx.ms + 65 * s + 1, ip, nullptr});
            }
          }
          if (stop) break;
        } else if (tx.n_ms >= 0) { host_fail(XHE_ERR_FORMAT); stop = true; break; }
      }
      // 1. commitment equality proofs + sender balance updates (296-341)
      P.eq_begin = (uint32_t)(B.eq_points.size() / 7); P.val_begin = (uint32_t)(B.val_points.size() / 8);
      for (uint32_t q = 0; q < a && !stop; q++) {
        const uint8_t* asset = tx.sc + 256 * q; const uint8_t* proof = asset + 64;
        long long pc, pd; int64_t loaded;
        Chain* ch = resolve_chain(B, state, tx.source, asset, Sender, &pc, &pd, &loaded);
        if (!ch) { host_fail(XHE_ERR_STATE); stop = true; break; }
        if (loaded >= 0) { B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, (uint32_t)loaded}); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, (uint32_t)loaded + 1}); }
        bool carry; const uint64_t amount = plain_output_amount(tx, asset, &carry);
        uint32_t oc = B.add_op(pc, amount);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) B.op_terms.push_back(iC[t] | 0x80000000u);
        if (carry) B.op_terms.push_back(B.g_2_64() | 0x80000000u);
        B.close_op();
        uint32_t od = B.add_op(pd, 0);
        for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) B.op_terms.push_back(iDs[t] | 0x80000000u);
        B.close_op();
        advance_chain(B, ch, oc, od);
        StateUpdate u; memcpy(u.account.data(), tx.source, 32); memcpy(u.asset.data(), asset, 32); u.role = Sender; u.op_c = oc; u.op_d = od; B.updates.push_back(u);
        if (want_out) push_output_ops(B, tx, asset, iC, iDs, k);
        // eq proof: Y identity check (src/transcript.rs:73-84) then Y decompression (src/proofs.rs:168-179)
        if (is_zero32(proof) || is_zero32(proof + 32) || is_zero32(proof + 64)) { host_fail(XHE_ERR_TRANSCRIPT); stop = true; break; }
        uint32_t y0 = B.add_point(proof), y1 = B.add_point(proof + 32), y2 = B.add_point(proof + 64);
        for (uint32_t y : {y0, y1, y2}) B.checks.push_back({CK_POINT, XHE_ERR_COMMITMENT_EQ_PROOF, y});
        // P_src, Y0, D_src, C_src, Y1, C_dst, Y2
        for (uint32_t p : {iSrc, y0, OPREF | od, OPREF | oc, y1, iN[q], y2}) B.eq_points.push_back(p);
        size_t so = B.eq_scalars.size(); B.eq_scalars.resize(so + 192);
        memcpy(&B.eq_scalars[so], proof + 96, 96);    // z_s, z_x, z_r ; c, w, bf are filled by the transcript phase
      }
      if (stop) break;
      // 2. transfers: receiver balance updates + ciphertext validity proofs (344-394)
      if (tx.type == 0) {
        for (uint32_t t = 0; t < k && !stop; t++) {
          const TransferView& tr = tx.transfers[t];
          uint32_t iDest = B.add_point(tr.dest); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, iDest});
          long long pc, pd; int64_t loaded;
          Chain* ch = resolve_chain(B, state, tr.dest, tr.asset, Receiver, &pc, &pd, &loaded);
          if (!ch) { host_fail(XHE_ERR_STATE); stop = true; break; }
          if (loaded >= 0) { B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, (uint32_t)loaded}); B.checks.push_back({CK_POINT, XHE_ERR_DECOMPRESSION, (uint32_t)loaded + 1}); }
          uint32_t oc = B.add_op(pc, 0); B.op_terms.push_back(iC[t]); B.close_op();
          uint32_t od = B.add_op(pd, 0); B.op_terms.push_back(iDr[t]); B.close_op();
          advance_chain(B, ch, oc, od);
          StateUpdate u; memcpy(u.account.data(), tr.dest, 32); memcpy(u.asset.data(), tr.asset, 32); u.role = Receiver; u.op_c = oc; u.op_d = od; B.updates.push_back(u);
          const uint8_t* proof = tr.proof;
          if (is_zero32(proof) || is_zero32(proof + 32) || is_zero32(proof + 64)) { host_fail(XHE_ERR_TRANSCRIPT); stop = true; break; }
          uint32_t y0 = B.add_point(proof), y1 = B.add_point(proof + 32), y2 = B.add_point(proof + 64);
          for (uint32_t y : {y0, y1, y2}) B.checks.push_back({CK_POINT, XHE_ERR_CT_VALIDITY_PROOF, y});
          // C, Y0, P_dest, D_dest, Y1, P_src, D_src, Y2
          for (uint32_t p : {iC[t], y0, iDest, iDr[t], y1, iSrc, iDs[t], y2}) B.val_points.push_back(p);
          size_t so = B.val_scalars.size(); B.val_scalars.resize(so + 160);
          memcpy(&B.val_scalars[so], proof + 96, 64);   // z_r, z_x
        }
        if (stop) break;
      } else if (tx.type == 4) {   // MultiSig setup (401-428)
        if (!multisig_payload_ok(tx)) { host_fail(XHE_ERR_FORMAT); stop = true; break; }
        staged.set_multisig(tx.source, tx.body, tx.count, (uint8_t)tx.aux, false);      // set_multisig_for_account: applied with the batch (src/tx/verify.rs:426)
      }
      // 3. range proof view (434-478, 504-510): commitments = new source commitments ++ transfer commitments ++ identity duds
      {
        uint32_t nc = a + k, m = 1; while (m < nc) m <<= 1;
        uint32_t lg = (tx.rp_len / 32 - 9) / 2, lg_need = 6; { uint32_t mm = m; while (mm > 1) { mm >>= 1; lg_need++; } }
        // more parties than bulletproofs' BP_GENS holds (BulletproofGens::new(64, 512), src/proofs.rs:20): the reference's batch
        // verifier answers with a RangeProof error, after the sigma check.  More than THIS context was created for is an
        // operator error, reported as such with the transaction's index before anything runs on the device.
        bool structural = (lg != lg_need) || m > 512;
        if (!structural && m > party_capacity) { capacity_fail = (long)i; stop = true; break; }
        const uint8_t* rp = tx.rp;
        for (int q = 0; q < 4 && !structural; q++) if (is_zero32(rp + 32 * q)) structural = true;
        for (uint32_t q = 0; q < 2 * lg && !structural; q++) if (is_zero32(rp + 224 + 32 * q)) structural = true;
        if (structural) { P.rp_structural_fail = true; }
        else {
          P.rp_slot = (int32_t)B.rp_m.size(); B.rp_m.push_back(m);
          for (int q = 0; q < 4; q++) B.rp_points.push_back(B.add_point(rp + 32 * q));
          Ls.resize(lg); Rs.resize(lg);
          for (uint32_t q = 0; q < lg; q++) { Ls[q] = B.add_point(rp + 224 + 64 * q); Rs[q] = B.add_point(rp + 224 + 64 * q + 32); }
          for (uint32_t q = 0; q < lg; q++) B.rp_points.push_back(Ls[q]);
          for (uint32_t q = 0; q < lg; q++) B.rp_points.push_back(Rs[q]);
          for (uint32_t q = 0; q < a; q++) B.rp_points.push_back(iN[q]);
          for (uint32_t q = 0; q < k; q++) B.rp_points.push_back(iC[q]);
          for (uint32_t q = nc; q < m; q++) B.rp_points.push_back(0);
          B.rp_point_off.push_back((uint32_t)B.rp_points.size());
          size_t so = B.rp_scalars.size(); B.rp_scalars.resize(so + 224);
          memcpy(&B.rp_scalars[so], rp + 128, 96);                        // t_x, t_x_blinding, e_blinding
          memcpy(&B.rp_scalars[so + 96], rp + tx.rp_len - 64, 64);       // a, b
          P.rp_chal_begin = B.rp_chal_off.back();
          B.rp_chal_off.push_back(P.rp_chal_begin + 4 + lg);
          B.rp_challenges.resize(32 * (size_t)B.rp_chal_off.back());
        }
      }
      P.proofs = true;
    } while (false);
    finish();
    if (stop) { n_reached = i + 1; plan.resize(n_reached - lo); break; }
  }
  if (capacity_fail >= 0) {
    if (fail_index) *fail_index = capacity_fail;
    return XHE_E_CAPACITY;
  }
  const size_t n_loc = n_reached > lo ? n_reached - lo : 0;      // transactions of this shard that reach the device
  double t2 = now_ms();

  // ---- phase B: Merlin transcripts / Fiat-Shamir challenges, message hashes, random batch factors (parallel over txs)
  const size_t n_sigs = B.sigs.size();
  const bool dev_fs = opt.device_fiat_shamir || opt.fast_path;
  std::vector<Sponge> sig_sponge(n_sigs, Sponge(72));
  std::vector<uint64_t> perms(threads > 0 ? threads : 1, 0);
  parallel_for(n_loc, threads, [&](size_t ja, size_t jb, int tid) {
    std::vector<uint8_t> bytes; uint64_t kf = 0;
    for (size_t j = ja; j < jb; j++) {
      const size_t i = lo + j;
      const TxView& tx = txs[j]; const TxPlan& P = plan[j];
      // signature message hashes: SHA3-512(pk || message || r) -- absorb everything but r now (src/elgamal.rs:53-65)
      // (device Fiat-Shamir mode: only multisig co-signatures, which need BLAKE3, are still hashed here)
      if (P.sig_end > P.sig_begin && (!dev_fs || P.sig_end - P.sig_begin > 1)) {
        size_t msi; tx.to_bytes(bytes, &msi);
        uint8_t h32[32]; bool have_hash = false;
        for (uint32_t s = P.sig_begin; s < P.sig_end; s++) {
          const SigEntry& e = B.sigs[s]; Sponge& sp = sig_sponge[s];
          if (!e.is_multisig) { if (dev_fs) continue; sp.absorb(tx.source, 32); sp.absorb(bytes.data(), bytes.size()); }
          else { if (!have_hash) { blake3(bytes.data(), msi, h32); have_hash = true; } sp.absorb(&B.points[32 * (size_t)e.pk], 32); sp.absorb(h32, 32); }
        }
      }
      if (!P.proofs || dev_fs) continue;
      Rng rng(seed, 32, i);                 // i = index in the WHOLE batch: two shards never share a factor stream
      Transcript T("transaction-proof");    // prepare_transcript, src/tx/verify.rs:146-158
      T.append_u64("version", tx.version); T.append("source_pubkey", tx.source, 32); T.append_u64("fee", tx.fee); T.append_u64("nonce", tx.nonce);
      auto challenge = [&](const char* label, uint8_t out[32]) { uint8_t b64[64]; T.challenge(label, b64, 64); ScalarL::reduce_wide(b64, out); };
      for (uint32_t q = 0; q < tx.n_sc; q++) {   // src/tx/verify.rs:315-318 + src/proofs.rs:142-161
        const uint8_t* asset = tx.sc + 256 * q; const uint8_t* proof = asset + 64;
        T.append("dom-sep", "new-commitment-proof", 20); T.append("new_source_commitment_asset", asset, 32); T.append("new_source_commitment", asset + 32, 32);
        T.append("dom-sep", "equality-proof", 14); T.append("Y_0", proof, 32); T.append("Y_1", proof + 32, 32); T.append("Y_2", proof + 64, 32);
        uint8_t* sc = &B.eq_scalars[192 * (size_t)(P.eq_begin + q)];
        challenge("c", sc + 96);
        T.append("z_s", proof + 96, 32); T.append("z_x", proof + 128, 32); T.append("z_r", proof + 160, 32);
        challenge("w", sc + 128); rng.scalar(sc + 160);
      }
      if (tx.type == 0) {
        for (uint32_t t = 0; t < tx.count; t++) {   // src/tx/verify.rs:378-383 + src/proofs.rs:291-302
          const TransferView& tr = tx.transfers[t];
          T.append("dom-sep", "transfer-proof", 14); T.append("dest_pubkey", tr.dest, 32); T.append("amount_commitment", tr.commitment, 32);
          T.append("amount_sender_handle", tr.sender_handle, 32); T.append("amount_receiver_handle", tr.receiver_handle, 32);
          T.append("dom-sep", "validity-proof", 14); T.append("Y_0", tr.proof, 32); T.append("Y_1", tr.proof + 32, 32); T.append("Y_2", tr.proof + 64, 32);
          uint8_t* sc = &B.val_scalars[160 * (size_t)(P.val_begin + t)];
          challenge("c", sc + 64);
          T.append("z_r", tr.proof + 96, 32); T.append("z_x", tr.proof + 128, 32);
          challenge("w", sc + 96); rng.scalar(sc + 128);
        }
      } else if (tx.type == 1) {
        T.append("dom-sep", "burn-proof", 10); T.append("asset", tx.body, 32); T.append_u64("amount", rd64(tx.body + 32));
      } else if (tx.type == 4) {
        T.append("dom-sep", "multisig-proof", 14); T.append_u64("threshold", tx.aux);
        for (uint32_t s = 0; s < tx.count; s++) T.append("signer", tx.body + 32 * s, 32);
      }
      if (P.rp_slot >= 0) {   // bulletproofs verification transcript
        uint32_t m = B.rp_m[P.rp_slot], lg = (tx.rp_len / 32 - 9) / 2; const uint8_t* rp = tx.rp;
        T.append("dom-sep", "rangeproof v1", 13); T.append_u64("n", 64); T.append_u64("m", m);
        for (uint32_t q = 0; q < tx.n_sc; q++) T.append("V", tx.sc + 256 * q + 32, 32);
        for (uint32_t q = 0; q < tx.n_transfers(); q++) T.append("V", tx.transfers[q].commitment, 32);
        for (uint32_t q = tx.n_sc + tx.n_transfers(); q < m; q++) T.append("V", ZERO32, 32);
        uint8_t* ch = &B.rp_challenges[32 * (size_t)P.rp_chal_begin];
        T.append("A", rp, 32); T.append("S", rp + 32, 32);
        challenge("y", ch); challenge("z", ch + 32);
        T.append("T_1", rp + 64, 32); T.append("T_2", rp + 96, 32);
        challenge("x", ch + 64);
        T.append("t_x", rp + 128, 32); T.append("t_x_blinding", rp + 160, 32); T.append("e_blinding", rp + 192, 32);
        challenge("w", ch + 96);
        T.append("dom-sep", "ipp v1", 6); T.append_u64("n", 64ull * m);
        for (uint32_t q = 0; q < lg; q++) { T.append("L", rp + 224 + 64 * q, 32); T.append("R", rp + 224 + 64 * q + 32, 32); challenge("u", ch + 128 + 32 * q); }
        uint8_t* sc = &B.rp_scalars[224 * (size_t)P.rp_slot];
        rng.scalar(sc + 160); rng.scalar(sc + 192);     // c (intra-proof weight), rho (batch factor)
      }
      kf += T.permutations;
    }
    perms[tid] += kf;
  });
  double t3 = now_ms();

  // ---- phase C: device
  const uint32_t n_points = (uint32_t)(B.points.size() / 32);
  for (size_t q = 0; q < B.eq_points.size(); q++) if (B.eq_points[q] & OPREF) B.eq_points[q] = n_points + (B.eq_points[q] & ~OPREF);
  PinnedVec<uint8_t>&sig_s = HC.sig_s, &sig_e = HC.sig_e; sig_s.resize(32 * n_sigs + 1); sig_e.resize(32 * n_sigs + 1); std::vector<uint32_t> sig_pk(n_sigs + 1);
  for (size_t s = 0; s < n_sigs; s++) { memcpy(&sig_s[32 * s], B.sigs[s].sig, 32); memcpy(&sig_e[32 * s], B.sigs[s].sig + 32, 32); sig_pk[s] = B.sigs[s].pk; }
  xhe_batch xb; memset(&xb, 0, sizeof xb); xb.struct_size = (uint32_t)sizeof xb;
  xb.n_tx = (uint32_t)n_loc; xb.n_points = n_points; xb.points = B.points.data();
  xb.n_sigs = (uint32_t)n_sigs; xb.sig_s = sig_s.data(); xb.sig_e = sig_e.data(); xb.sig_pk = sig_pk.data();
  xb.n_ops = (uint32_t)B.op_prev.size(); xb.op_prev = (const int64_t*)B.op_prev.data(); xb.op_term_off = B.op_term_off.data(); xb.op_terms = B.op_terms.data(); xb.op_amount = B.op_amount.data(); xb.max_chain = B.max_chain;
  xb.n_eq = (uint32_t)(B.eq_points.size() / 7); xb.eq_points = B.eq_points.data(); xb.eq_scalars = B.eq_scalars.data();
  xb.n_val = (uint32_t)(B.val_points.size() / 8); xb.val_points = B.val_points.data(); xb.val_scalars = B.val_scalars.data();
  xb.n_rp = (uint32_t)B.rp_m.size(); xb.rp_m = B.rp_m.data(); xb.rp_point_off = B.rp_point_off.data(); xb.rp_points = B.rp_points.data();
  xb.rp_scalars = B.rp_scalars.data(); xb.rp_chal_off = B.rp_chal_off.data(); xb.rp_challenges = B.rp_challenges.data();
  PinnedVec<uint8_t>& fs_blob = HC.fs_blob; PinnedVec<uint64_t>& fs_off = HC.fs_off; PinnedVec<uint32_t>& fs_plan = HC.fs_plan;
  if (dev_fs && n_loc) {
    fs_off.resize(n_loc + 1); fs_off[0] = 0;
    for (size_t j = 0; j < n_loc; j++) fs_off[j + 1] = fs_off[j] + ((lens[lo + j] + 15) & ~(size_t)15);
    fs_blob.resize(fs_off[n_loc]); fs_plan.resize(6 * n_loc);
    parallel_for(n_loc, threads, [&](size_t ja, size_t jb, int) {
      for (size_t j = ja; j < jb; j++) {
        memcpy(&fs_blob[fs_off[j]], blobs[lo + j], lens[lo + j]);
        const TxPlan& P = plan[j]; uint32_t* w = &fs_plan[6 * j];
        w[0] = P.eq_begin; w[1] = P.val_begin; w[2] = P.rp_slot >= 0 ? (uint32_t)P.rp_slot : 0xFFFFFFFFu; w[3] = P.rp_chal_begin;
        w[4] = P.sig_end > P.sig_begin ? P.sig_begin : 0xFFFFFFFFu; w[5] = P.proofs ? 1u : 0u;
      }
    });
    xb.fs_blobs = fs_blob.data(); xb.fs_blob_off = fs_off.data(); xb.fs_plan = fs_plan.data(); memcpy(xb.fs_seed, seed, 32); xb.fs_index_base = lo;
  }
  std::vector<uint8_t> point_ok(n_points + 1), sig_r(32 * n_sigs + 1), op_out(32 * (size_t)xb.n_ops + 1), sig_ok_dev(n_sigs + 1, 0);
  xhe_verdict v; memset(&v, 0, sizeof v); v.struct_size = (uint32_t)sizeof v; v.point_ok = point_ok.data(); v.sig_r = sig_r.data(); v.op_out = op_out.data(); v.sig_ok = sig_ok_dev.data();
  int32_t rc; { DeviceSlots::Hold slot = device_slots().acquire(); rc = xhe_verify_batch(ctx, &xb, &v); }
  double t4 = now_ms();
  if (rc != XHE_OK) { if (tm) { tm->parse_ms = t1 - t0; tm->resolve_ms = t2 - t1; tm->transcript_ms = t3 - t2; tm->device_ms = t4 - t3; tm->total_ms = t4 - t0; } return rc; }

  // ---- phase D: verdict with the reference's precedence
  std::vector<uint8_t> sig_ok(n_sigs + 1, 0);
  parallel_for(n_sigs, threads, [&](size_t lo, size_t hi, int) {
    for (size_t s = lo; s < hi; s++) {
      if (dev_fs && !B.sigs[s].is_multisig) { sig_ok[s] = sig_ok_dev[s]; continue; }
      Sponge sp = sig_sponge[s]; sp.absorb(&sig_r[32 * s], 32); sp.finish(0x06);
      uint8_t h[64], e2[32]; sp.squeeze(h, 64); ScalarL::reduce_wide(h, e2);
      sig_ok[s] = memcmp(e2, B.sigs[s].sig + 32, 32) == 0;
    }
  });
  int verdict = XHE_OK; long bad_tx = -1;
  for (size_t j = 0; j < n_loc && verdict == XHE_OK; j++) {
    for (uint32_t c = plan[j].check_begin; c < plan[j].check_end; c++) {
      const Check& ck = B.checks[c]; bool fail = false;
      if (ck.kind == CK_HOST) fail = true; else if (ck.kind == CK_POINT) fail = !point_ok[ck.a]; else fail = !sig_ok[ck.a];
      if (fail) { verdict = ck.err; bad_tx = (long)(lo + j); break; }
    }
  }
  if (verdict == XHE_OK && parse_err != XHE_OK) { verdict = parse_err; bad_tx = (long)n_live; }
  const bool shard = opt.partial_out != nullptr;
  if (shard) { memcpy(opt.partial_out, v.sigma_enc, 32); memcpy(opt.partial_out + 32, v.range_enc, 32); }
  if (verdict == XHE_OK && !shard && !v.sigma_is_identity) verdict = XHE_ERR_GENERIC_PROOF;            // src/tx/verify.rs:500-502
  if (verdict == XHE_OK) { for (size_t j = 0; j < n_loc; j++) if (plan[j].rp_structural_fail) verdict = XHE_ERR_RANGE_PROOF; }
  if (verdict == XHE_OK && (v.device_flags & 16u)) verdict = XHE_ERR_RANGE_PROOF;                       // a zero folding challenge (k_rp_prep: the scaled equation would hold trivially)
  if (verdict == XHE_OK && !shard && !v.range_is_identity) verdict = XHE_ERR_RANGE_PROOF;               // src/tx/verify.rs:504-514
  if (shard) {
    Pending Pn; Pn.updates = B.updates; Pn.op_out.assign(op_out.data(), op_out.data() + op_out.size()); Pn.staged = std::move(staged);
    std::lock_guard<std::mutex> g(g_pending_mu);
    g_pending[ctx] = std::move(Pn);
  } else if (verdict == XHE_OK && opt.apply_state) {
    if (staged.apply(state) != XHE_OK) verdict = XHE_ERR_STATE;
    if (verdict == XHE_OK) for (const StateUpdate& u : B.updates) {
      uint8_t ct[64]; memcpy(ct, &op_out[32 * (size_t)u.op_c], 32); memcpy(ct + 32, &op_out[32 * (size_t)u.op_d], 32);
      if (!apply_update(state, u.account.data(), u.asset.data(), u.role, u.output, ct)) { verdict = XHE_ERR_STATE; break; }
    }
  }
  if (fail_index) *fail_index = bad_tx;
  double t5 = now_ms();
  if (tm) { tm->parse_ms = t1 - t0; tm->resolve_ms = t2 - t1; tm->transcript_ms = t3 - t2; tm->device_ms = t4 - t3; tm->finish_ms = t5 - t4; tm->total_ms = t5 - t0; tm->keccak_f = 0; for (uint64_t p : perms) tm->keccak_f += p; }
  return verdict;
}

// ------------------------------------------------------------------------------------------------------------------
// Transaction::apply_without_verify over a list of txs (src/tx/verify.rs:545-619)
// ------------------------------------------------------------------------------------------------------------------
int apply_without_verify(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n, VerificationState& state) {
  std::vector<TxView> txs(n);
  for (size_t i = 0; i < n; i++) { int rc = txs[i].parse(blobs[i], lens[i]); if (rc) return rc; }
  Builder B(cache_for(ctx)); std::vector<uint32_t> need_ok;
  const bool want_out = state.wants_output_ciphertexts();
  for (size_t i = 0; i < n; i++) {
    const TxView& tx = txs[i]; const uint32_t k = tx.n_transfers();
    std::vector<uint32_t> iC(k), iDs(k), iDr(k);
    for (uint32_t t = 0; t < k; t++) { iC[t] = B.add_point(tx.transfers[t].commitment); iDs[t] = B.add_point(tx.transfers[t].sender_handle); iDr[t] = B.add_point(tx.transfers[t].receiver_handle); need_ok.push_back(iC[t]); need_ok.push_back(iDs[t]); need_ok.push_back(iDr[t]); }
    for (uint32_t q = 0; q < tx.n_sc; q++) {
      const uint8_t* asset = tx.sc + 256 * q; long long pc, pd; int64_t loaded;
      Chain* ch = resolve_chain(B, state, tx.source, asset, Sender, &pc, &pd, &loaded); if (!ch) return XHE_ERR_STATE;
      if (loaded >= 0) { need_ok.push_back((uint32_t)loaded); need_ok.push_back((uint32_t)loaded + 1); }
      bool carry; const uint64_t amount = plain_output_amount(tx, asset, &carry);
      uint32_t oc = B.add_op(pc, amount);
      for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) B.op_terms.push_back(iC[t] | 0x80000000u);
      if (carry) B.op_terms.push_back(B.g_2_64() | 0x80000000u);
      B.close_op();
      uint32_t od = B.add_op(pd, 0);
      for (uint32_t t = 0; t < k; t++) if (!memcmp(asset, tx.transfers[t].asset, 32)) B.op_terms.push_back(iDs[t] | 0x80000000u);
      B.close_op();
      advance_chain(B, ch, oc, od);
      StateUpdate u; memcpy(u.account.data(), tx.source, 32); memcpy(u.asset.data(), asset, 32); u.role = Sender; u.op_c = oc; u.op_d = od; B.updates.push_back(u);
      if (want_out) push_output_ops(B, tx, asset, iC, iDs, k);
    }
    for (uint32_t t = 0; t < k; t++) {
      const TransferView& tr = tx.transfers[t]; long long pc, pd; int64_t loaded;
      Chain* ch = resolve_chain(B, state, tr.dest, tr.asset, Receiver, &pc, &pd, &loaded); if (!ch) return XHE_ERR_STATE;
      if (loaded >= 0) { need_ok.push_back((uint32_t)loaded); need_ok.push_back((uint32_t)loaded + 1); }
      uint32_t oc = B.add_op(pc, 0); B.op_terms.push_back(iC[t]); B.close_op();
      uint32_t od = B.add_op(pd, 0); B.op_terms.push_back(iDr[t]); B.close_op();
      advance_chain(B, ch, oc, od);
      StateUpdate u; memcpy(u.account.data(), tr.dest, 32); memcpy(u.asset.data(), tr.asset, 32); u.role = Receiver; u.op_c = oc; u.op_d = od; B.updates.push_back(u);
    }
    if (tx.type == 4) state.set_multisig_for_account(tx.source, tx.body, tx.count, (uint8_t)tx.aux);
  }
  const uint32_t n_points = (uint32_t)(B.points.size() / 32);
  xhe_batch xb; memset(&xb, 0, sizeof xb); xb.struct_size = (uint32_t)sizeof xb;
  xb.n_tx = (uint32_t)n; xb.n_points = n_points; xb.points = B.points.data();
  xb.n_ops = (uint32_t)B.op_prev.size(); xb.op_prev = (const int64_t*)B.op_prev.data(); xb.op_term_off = B.op_term_off.data(); xb.op_terms = B.op_terms.data(); xb.op_amount = B.op_amount.data(); xb.max_chain = B.max_chain;
  uint32_t zero_off = 0; xb.rp_point_off = &zero_off; xb.rp_chal_off = &zero_off;
  std::vector<uint8_t> point_ok(n_points + 1), op_out(32 * (size_t)xb.n_ops + 1);
  xhe_verdict v; memset(&v, 0, sizeof v); v.struct_size = (uint32_t)sizeof v; v.point_ok = point_ok.data(); v.op_out = op_out.data();
  int32_t rc = xhe_verify_batch(ctx, &xb, &v); if (rc) return rc;
  for (uint32_t p : need_ok) if (!point_ok[p]) return XHE_ERR_DECOMPRESSION;   // "ill-formed ciphertext" (the reference panics)
  for (const StateUpdate& u : B.updates) {
    uint8_t ct[64]; memcpy(ct, &op_out[32 * (size_t)u.op_c], 32); memcpy(ct + 32, &op_out[32 * (size_t)u.op_d], 32);
    if (!apply_update(state, u.account.data(), u.asset.data(), u.role, u.output, ct)) return XHE_ERR_STATE;
  }
  return XHE_OK;
}

}  // namespace xhe_host

// ------------------------------------------------------------------------------------------------------------------
// C exports for the Python binding / tests (the mock ledger plays the role of src/lib.rs::mock::Ledger)
// ------------------------------------------------------------------------------------------------------------------
using namespace xhe_host;
extern "C" {
// (every ledger handle is a VerificationState*: the mock and the device-resident state are passed to the same verify entry points)
void* xheh_ledger_new() { return static_cast<VerificationState*>(new MockLedger()); }
// device-resident state: balances in an xhe_ledger on ctx's device
void* xheh_dledger_new(xhe_ctx* ctx, size_t capacity) { DeviceLedgerState* d = new DeviceLedgerState(ctx, capacity); if (!d->led) { delete d; return nullptr; } return static_cast<VerificationState*>(d); }
void xheh_dledger_free(void* l) { delete static_cast<DeviceLedgerState*>((VerificationState*)l); }
int32_t xheh_dledger_import(void* l, const uint8_t* recs, size_t n) { return static_cast<DeviceLedgerState*>((VerificationState*)l)->import_records(recs, n) ? XHE_OK : XHE_ERR_STATE; }
size_t xheh_dledger_export(void* l, uint8_t* out, size_t cap) { return static_cast<DeviceLedgerState*>((VerificationState*)l)->export_all(out, cap); }
void xheh_dledger_set_multisig(void* l, const uint8_t* pk, const uint8_t* signers, size_t n, uint8_t threshold) { static_cast<DeviceLedgerState*>((VerificationState*)l)->set_multisig_for_account(pk, signers, n, threshold); }
int32_t xheh_dledger_snapshot(void* l) { return xhe_ledger_snapshot(static_cast<DeviceLedgerState*>((VerificationState*)l)->led); }
int32_t xheh_dledger_restore(void* l) { return xhe_ledger_restore(static_cast<DeviceLedgerState*>((VerificationState*)l)->led); }
void* xheh_ledger_clone(const void* l) { return static_cast<VerificationState*>(new MockLedger(*static_cast<const MockLedger*>((const VerificationState*)l))); }
void xheh_ledger_free(void* l) { delete static_cast<MockLedger*>((VerificationState*)l); }
void xheh_ledger_set_balance(void* l, const uint8_t* pk, const uint8_t* asset, const uint8_t* ct) { (static_cast<MockLedger*>((VerificationState*)l))->set_balance(pk, asset, ct); }
int xheh_ledger_get_balance(void* l, const uint8_t* pk, const uint8_t* asset, uint8_t* ct) { return (static_cast<MockLedger*>((VerificationState*)l))->get_account_balance(pk, asset, Sender, ct) ? 1 : 0; }
void xheh_ledger_set_nonce(void* l, const uint8_t* pk, uint64_t nonce) { (static_cast<MockLedger*>((VerificationState*)l))->set_nonce(pk, nonce); }
void xheh_ledger_set_multisig(void* l, const uint8_t* pk, const uint8_t* signers, size_t n, uint8_t threshold) { (static_cast<MockLedger*>((VerificationState*)l))->set_multisig_for_account(pk, signers, n, threshold); }
int xheh_ledger_has_multisig(void* l, const uint8_t* pk) { std::vector<Bytes32> s; uint8_t t; bool p; (static_cast<MockLedger*>((VerificationState*)l))->get_multisig_for_account(pk, &s, &t, &p); return p ? 1 : 0; }
size_t xheh_ledger_size(void* l) { return (static_cast<MockLedger*>((VerificationState*)l))->balances.size(); }
// output ciphertexts (set_output_ciphertext, src/tx/verify.rs:60-66): off by default like the reference mock, which drops them
void xheh_ledger_record_outputs(void* l, int on) { (static_cast<MockLedger*>((VerificationState*)l))->record_outputs = on != 0; }
size_t xheh_ledger_outputs_size(void* l) { return (static_cast<MockLedger*>((VerificationState*)l))->outputs.size(); }
size_t xheh_ledger_export_outputs(void* l, uint8_t* out, size_t cap) { MockLedger* L = static_cast<MockLedger*>((VerificationState*)l); size_t i = 0; L->outputs.for_each([&](const uint8_t* k, const Ct64& v) { if ((i + 1) * 128 <= cap) { memcpy(out + 128 * i, k, 64); memcpy(out + 128 * i + 64, v.data(), 64); } i++; }); return i; }
// bulk import of records (pk[32] asset[32] ct[64]) and nonce-0 accounts
void xheh_ledger_import(void* l, const uint8_t* recs, size_t n) { MockLedger* L = static_cast<MockLedger*>((VerificationState*)l); L->balances.reserve(L->balances.size() + n); L->nonces.reserve(L->nonces.size() + n);
  for (size_t i = 0; i < n; i++) { const uint8_t* r = recs + 128 * i; L->set_balance(r, r + 32, r + 64); bool fresh = false; uint64_t* v = L->nonces.insert(r, &fresh); if (fresh) *v = 0; } }
size_t xheh_ledger_export(void* l, uint8_t* out, size_t cap) { MockLedger* L = static_cast<MockLedger*>((VerificationState*)l); size_t i = 0; L->balances.for_each([&](const uint8_t* k, const Ct64& v) { if ((i + 1) * 128 <= cap) { memcpy(out + 128 * i, k, 64); memcpy(out + 128 * i + 64, v.data(), 64); } i++; }); return i; }
// timings: parse, resolve, transcript, device, finish, total (ms), keccak permutations
int32_t xheh_verify_batch(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, const uint8_t* seed, size_t seed_len, int threads, long* fail_index, double* timings7) {
  BatchOptions opt; opt.threads = threads; opt.rng_seed = seed; opt.rng_seed_len = seed_len;
  BatchTimings tm;
  int rc = verify_batch(ctx, blobs, lens, n, *(VerificationState*)ledger, opt, fail_index, &tm);
  if (timings7) { timings7[0] = tm.parse_ms; timings7[1] = tm.resolve_ms; timings7[2] = tm.transcript_ms; timings7[3] = tm.device_ms; timings7[4] = tm.finish_ms; timings7[5] = tm.total_ms; timings7[6] = (double)tm.keccak_f; }
  return rc;
}
int32_t xheh_verify_batch_partial(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, const uint8_t* seed, size_t seed_len, int threads, long* fail_index, double* timings7, uint8_t* partial64) {
  BatchOptions opt; opt.threads = threads; opt.rng_seed = seed; opt.rng_seed_len = seed_len; opt.partial_out = partial64;
  BatchTimings tm;
  int rc = verify_batch(ctx, blobs, lens, n, *(VerificationState*)ledger, opt, fail_index, &tm);
  if (timings7) { timings7[0] = tm.parse_ms; timings7[1] = tm.resolve_ms; timings7[2] = tm.transcript_ms; timings7[3] = tm.device_ms; timings7[4] = tm.finish_ms; timings7[5] = tm.total_ms; timings7[6] = (double)tm.keccak_f; }
  return rc;
}
// general entry: flags bit 0 = device-side Fiat-Shamir, bit 1 = shard mode (partial64 must be non-null), bit 2 = fast path,
// bit 3 = replayable batch factors (tests only: the seed is used without OS entropy)
int32_t xheh_verify_batch_ex(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, const uint8_t* seed, size_t seed_len, int threads, uint32_t flags, long* fail_index, double* timings7, uint8_t* partial64) {
  BatchOptions opt; opt.threads = threads; opt.rng_seed = seed; opt.rng_seed_len = seed_len; opt.device_fiat_shamir = (flags & 1u) != 0; opt.partial_out = (flags & 2u) ? partial64 : nullptr; opt.fast_path = (flags & 4u) != 0;
  opt.deterministic_seed = (flags & 8u) != 0; opt.host_dry_run = (flags & 16u) != 0;
  BatchTimings tm;
  int rc = verify_batch(ctx, blobs, lens, n, *(VerificationState*)ledger, opt, fail_index, &tm);
  if (timings7) { timings7[0] = tm.parse_ms; timings7[1] = tm.resolve_ms; timings7[2] = tm.transcript_ms; timings7[3] = tm.device_ms; timings7[4] = tm.finish_ms; timings7[5] = tm.total_ms; timings7[6] = tm.used_fast_path ? -1.0 : (double)tm.keccak_f; }
  return rc;
}
// one rank's share of a sharded batch: blobs = the WHOLE batch, this call verifies [lo, hi) and follows the
// balance chains that start in [0, lo).  Always shard mode: partial64 receives the partial sigma / range encodings, the
// state updates wait for xheh_commit_pending / xheh_take_pending, *fail_index is an index into the whole batch.
int32_t xheh_verify_batch_shard(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, size_t lo, size_t hi, const uint8_t* seed, size_t seed_len, int threads, uint32_t flags,
                                long* fail_index, double* timings7, uint8_t* partial64) {
  if (!partial64 || lo > hi) return XHE_E_ARG;
  BatchOptions opt; opt.threads = threads; opt.rng_seed = seed; opt.rng_seed_len = seed_len; opt.device_fiat_shamir = (flags & 1u) != 0; opt.partial_out = partial64; opt.fast_path = (flags & 4u) != 0;
  opt.deterministic_seed = (flags & 8u) != 0; opt.host_dry_run = (flags & 16u) != 0; opt.shard_lo = lo; opt.shard_hi = hi;
  BatchTimings tm;
  int rc = verify_batch(ctx, blobs, lens, n, *(VerificationState*)ledger, opt, fail_index, &tm);
  if (timings7) { timings7[0] = tm.parse_ms; timings7[1] = tm.resolve_ms; timings7[2] = tm.transcript_ms; timings7[3] = tm.device_ms; timings7[4] = tm.finish_ms; timings7[5] = tm.total_ms; timings7[6] = tm.used_fast_path ? -1.0 : (double)tm.keccak_f; }
  return rc;
}
// the same with the batch's key-digest index (xheh_batch_index_build): the earlier shards' bytes are not scanned
int32_t xheh_verify_batch_shard_ix(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n, size_t lo, size_t hi, const uint8_t* seed, size_t seed_len, int threads, uint32_t flags,
                                   long* fail_index, double* timings7, uint8_t* partial64, const void* index) {
  if (!partial64 || lo > hi) return XHE_E_ARG;
  BatchOptions opt; opt.threads = threads; opt.rng_seed = seed; opt.rng_seed_len = seed_len; opt.device_fiat_shamir = (flags & 1u) != 0; opt.partial_out = partial64; opt.fast_path = (flags & 4u) != 0;
  opt.deterministic_seed = (flags & 8u) != 0; opt.host_dry_run = (flags & 16u) != 0; opt.shard_lo = lo; opt.shard_hi = hi; opt.key_index = index;
  BatchTimings tm;
  int rc = verify_batch(ctx, blobs, lens, n, *(VerificationState*)ledger, opt, fail_index, &tm);
  if (timings7) { timings7[0] = tm.parse_ms; timings7[1] = tm.resolve_ms; timings7[2] = tm.transcript_ms; timings7[3] = tm.device_ms; timings7[4] = tm.finish_ms; timings7[5] = tm.total_ms; timings7[6] = tm.used_fast_path ? -1.0 : (double)tm.keccak_f; }
  return rc;
}
// key-digest index of a batch: per transaction, 64-bit digests of the (account, asset) balances it moves and of a multisig
// setting it makes.  Built once, where the transactions are received and framed (next to the blob arena); a shard-mode call
// that gets it finds the earlier transactions its own depend on without reading the earlier shards' bytes.
void* xheh_batch_index_build(const uint8_t* const* blobs, const size_t* lens, size_t n, int threads) { return build_index(blobs, lens, n, threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency())); }
void xheh_batch_index_free(void* index) { delete (BatchIndex*)index; }
size_t xheh_batch_index_bytes(const void* index) { const BatchIndex* ix = (const BatchIndex*)index; return ix ? 4 * ix->off.size() + 8 * ix->dig.size() : 0; }
// the transactions of [0, lo) that shard [lo, hi) depends on (shared balances, multisig settings of its senders), in batch
// order; index may be NULL (the blobs are scanned).  Returns the count (out receives at most cap of them); -1: a transaction of the shard does not parse.
long xheh_shard_dependencies(const uint8_t* const* blobs, const size_t* lens, size_t n, size_t lo, size_t hi, const void* index, int threads, size_t* out, size_t cap) {
  lo = std::min(lo, n); hi = std::min(hi, n);
  std::vector<TxView> txs(hi - lo);
  for (size_t j = lo; j < hi; j++) if (txs[j - lo].parse(blobs[j], lens[j])) return -1;
  ShardKeys K(txs.data(), txs.size());
  const BatchIndex* ix = (const BatchIndex*)index;
  std::vector<size_t> hits = foreign_hits(blobs, lens, lo, K, std::max(1, threads), (ix && ix->n == n) ? ix : nullptr);
  for (size_t i = 0; i < hits.size() && i < cap; i++) out[i] = hits[i];
  return (long)hits.size();
}
int32_t xheh_commit_pending(xhe_ctx* ctx, void* ledger) { return commit_pending(ctx, *(VerificationState*)ledger); }
// detached form: take the held-back updates now, commit (or drop) them when the cross-rank decision is known
void* xheh_take_pending(xhe_ctx* ctx) { return take_pending(ctx); }
int32_t xheh_commit_taken(void* pending, void* ledger) { if (!pending) return XHE_E_ARG; Pending* P = (Pending*)pending; int rc = apply_pending(*P, *(VerificationState*)ledger); delete P; return rc; }
void xheh_drop_taken(void* pending) { delete (Pending*)pending; }
// a page-locked buffer for a batch's transactions laid out the way the device reads them (back to back, each padded to 16
// bytes): transactions that live there are uploaded without the gather into staging (zero-copy input)
void* xheh_blob_arena_alloc(size_t bytes) { void* p = nullptr; if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void xheh_blob_arena_free(void* p) { if (p) cudaFreeHost(p); }
// the balance updates a detached shard-mode batch holds, as 128-byte records (account, asset, new ciphertext) in update order
size_t xheh_export_taken(void* pending, uint8_t* out, size_t cap) { return export_pending(pending, out, cap); }
// apply such records to a ledger (a peer rank's updates): update_account_balance per record, in order
int32_t xheh_ledger_apply_records(void* l, const uint8_t* recs, size_t n) { MockLedger* L = static_cast<MockLedger*>((VerificationState*)l); for (size_t i = 0; i < n; i++) { const uint8_t* r = recs + 128 * i; if (!L->update_account_balance(r, r + 32, r + 64, Receiver)) return XHE_ERR_STATE; } return XHE_OK; }
int32_t xheh_apply_without_verify(xhe_ctx* ctx, void* ledger, const uint8_t* const* blobs, const size_t* lens, size_t n) { return apply_without_verify(ctx, blobs, lens, n, *(VerificationState*)ledger); }
// host-only helpers exposed for CPU tests of the host logic
void xheh_merlin_test(const char* proto, const char* label, const uint8_t* msg, size_t n, const char* chal_label, uint8_t* out, size_t outlen) { Transcript t(proto); t.append(label, msg, n); t.challenge(chal_label, out, outlen); }
void xheh_sha3_512(const uint8_t* m, size_t n, uint8_t* out) { sha3_512(m, n, out); }
void xheh_shake256(const uint8_t* m, size_t n, uint8_t* out, size_t outlen) { shake256(m, n, out, outlen); }
void xheh_blake3(const uint8_t* m, size_t n, uint8_t* out) { blake3(m, n, out); }
void xheh_reduce_wide(const uint8_t* in, uint8_t* out) { ScalarL::reduce_wide(in, out); }
void xheh_const_g_2_64(uint8_t* out32) { memcpy(out32, xhe_host::ENC_G_2_64, 32); }
int32_t xheh_tx_to_bytes(const uint8_t* blob, size_t len, uint8_t* out, size_t cap, size_t* out_len, size_t* ms_index) {
  TxView tx; int rc = tx.parse(blob, len); if (rc) return rc; std::vector<uint8_t> b; tx.to_bytes(b, ms_index); *out_len = b.size(); if (b.size() <= cap) memcpy(out, b.data(), b.size()); return XHE_OK; }
}
