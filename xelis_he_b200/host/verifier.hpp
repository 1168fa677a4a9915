// host/verifier.hpp -- HOST side of the drop-in: the C++ mirror of the reference's verification API.
//
//   reference (Rust)                                          here
//   trait BlockchainVerificationState  src/tx/verify.rs:25-77   class xhe_host::VerificationState
//   mock::Ledger                       src/lib.rs:106-201       class xhe_host::MockLedger
//   Transaction (serde)                src/tx/mod.rs:102-119    xtx1 wire blobs (format: oracle/tx.h header comment) + TxView
//   Transaction::to_bytes              src/tx/verify.rs:623-688 TxView::to_bytes
//   Transaction::verify_batch          src/tx/verify.rs:487-517 xhe_host::verify_batch
//   Transaction::apply_without_verify  src/tx/verify.rs:545-619 xhe_host::apply_without_verify
//
// The host keeps what north_star assigns to it: parsing, state lookups, Merlin transcripts / Fiat-Shamir challenges,
// SHA3-512 / BLAKE3 message hashes and the verdict precedence; every group / field / scalar-expansion operation is
// delegated to the device through the C ABI (include/xhe.h).  In a Rust deployment this file is the part that stays Rust.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <array>
#include <string>
#include <algorithm>
#include <unordered_map>
#include <vector>
#include "../../include/xhe.h"

namespace xhe_host {

typedef std::array<uint8_t, 32> Bytes32;
typedef std::array<uint8_t, 64> Ct64;   // CompressedCiphertext (src/compressed.rs:37-41): commitment || handle
enum Role { Sender = 0, Receiver = 1 };

struct VerificationState {               // src/tx/verify.rs:25-77; every call returns false for Self::Error
  virtual ~VerificationState() {}
  virtual bool get_account_balance(const uint8_t account[32], const uint8_t asset[32], Role role, uint8_t out_ct[64]) = 0;
  virtual bool update_account_balance(const uint8_t account[32], const uint8_t asset[32], const uint8_t new_ct[64], Role role) = 0;
  virtual bool get_account_nonce(const uint8_t account[32], uint64_t* nonce) = 0;
  virtual bool update_account_nonce(const uint8_t account[32], uint64_t nonce) = 0;
  virtual bool set_multisig_for_account(const uint8_t account[32], const uint8_t* signers, size_t n, uint8_t threshold) = 0;
  virtual bool get_multisig_for_account(const uint8_t account[32], std::vector<Bytes32>* signers, uint8_t* threshold, bool* present) = 0;
  // set_output_ciphertext (src/tx/verify.rs:60-66, called at 339-340 and 582): the ciphertext get_sender_output_ct
  // produced for (source, asset), handed over COMPRESSED (64 B; dalek's RistrettoPoint is opaque to a C ABI).  It costs
  // two extra encodings per (tx, asset) on the device, so it is computed only for states that ask for it.
  virtual bool wants_output_ciphertexts() const { return false; }
  virtual bool set_output_ciphertext(const uint8_t /*account*/[32], const uint8_t /*asset*/[32], const uint8_t /*ct*/[64]) { return true; }
  // optional hints (not part of the reference trait): the batch front end announces the lookups it is about to make
  virtual void prefetch_account(const uint8_t[32]) const {}
  virtual void prefetch_balance(const uint8_t[32], const uint8_t[32]) const {}
  // optional device-resident backend (not part of the reference trait): a state whose balances live in an
  // xhe_ledger on the verifying context's device hands out the ledger and the slot of a key.  The fast path then reads the
  // balance where it is (no upload, no decompression) and commits the update on the device (no download, no re-encoding on
  // the host side); get_account_balance / update_account_balance stay available as the compressed view of the same data.
  virtual xhe_ledger* device_ledger() { return nullptr; }
  virtual bool device_slot(const uint8_t /*account*/[32], const uint8_t /*asset*/[32], uint32_t* /*slot*/) { return false; }
};

struct KeyHash { size_t operator()(const Ct64& k) const { uint64_t h; memcpy(&h, k.data() + 5, 8); uint64_t g; memcpy(&g, k.data() + 37, 8); return (size_t)(h * 0x9E3779B97F4A7C15ull ^ g); } };
struct Key32Hash { size_t operator()(const Bytes32& k) const { uint64_t h; memcpy(&h, k.data() + 5, 8); return (size_t)(h * 0x9E3779B97F4A7C15ull); } };

// Open-addressing table with inline keys (linear probing, one tag byte per slot, no erase): a lookup touches one tag
// line and one entry, and the slot of a key can be prefetched from its hash alone -- what the batch front end needs
// when it walks 10^4 transactions' accounts in order.
template <size_t KB, typename V>
class FlatTable {
 public:
  struct alignas(KB >= 64 ? 64 : 8) Entry { std::array<uint8_t, KB> key; V val; };
  static uint64_t hash(const uint8_t* k) {     // keys are (concatenations of) compressed points / asset hashes: already uniform
    uint64_t h; memcpy(&h, k + 5, 8);
    if (KB >= 64) { uint64_t g; memcpy(&g, k + 37, 8); h = h * 0x9E3779B97F4A7C15ull ^ g; }
    return h * 0xD6E8FEB86659FD93ull;
  }
  size_t size() const { return count_; }
  void reserve(size_t n) { size_t cap = 16; while (cap < 2 * n) cap <<= 1; if (cap > tags_.size()) rehash(cap); }
  V* find(const uint8_t* k) {
    if (tags_.empty()) return nullptr;
    const uint64_t h = hash(k); const size_t mask = tags_.size() - 1; const uint8_t tag = (uint8_t)(h >> 57) | 0x80;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
      const uint8_t c = tags_[i];
      if (c == 0) return nullptr;
      if (c == tag && !memcmp(entries_[i].key.data(), k, KB)) return &entries_[i].val;
    }
  }
  const V* find(const uint8_t* k) const { return const_cast<FlatTable*>(this)->find(k); }
  // returns the value slot; *inserted tells whether the key was new (its value is then value-initialised)
  V* insert(const uint8_t* k, bool* inserted = nullptr) {
    if (2 * (count_ + 1) > tags_.size()) rehash(tags_.empty() ? 16 : 2 * tags_.size());
    const uint64_t h = hash(k); const size_t mask = tags_.size() - 1; const uint8_t tag = (uint8_t)(h >> 57) | 0x80;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
      const uint8_t c = tags_[i];
      if (c == 0) { tags_[i] = tag; memcpy(entries_[i].key.data(), k, KB); entries_[i].val = V(); count_++; if (inserted) *inserted = true; return &entries_[i].val; }
      if (c == tag && !memcmp(entries_[i].key.data(), k, KB)) { if (inserted) *inserted = false; return &entries_[i].val; }
    }
  }
  void prefetch(const uint8_t* k) const {
    if (tags_.empty()) return;
    const size_t i = (size_t)hash(k) & (tags_.size() - 1);
    __builtin_prefetch(&tags_[i]); const char* e = (const char*)&entries_[i]; __builtin_prefetch(e); if (sizeof(Entry) > 64) __builtin_prefetch(e + 64);
  }
  void clear() { std::fill(tags_.begin(), tags_.end(), 0); count_ = 0; }
  template <typename F> void for_each(F f) const { for (size_t i = 0; i < tags_.size(); i++) if (tags_[i]) f(entries_[i].key.data(), entries_[i].val); }
 private:
  void rehash(size_t cap) {
    std::vector<uint8_t> ot; std::vector<Entry> oe; ot.swap(tags_); oe.swap(entries_);
    tags_.assign(cap, 0); entries_.resize(cap); count_ = 0;
    for (size_t i = 0; i < ot.size(); i++) if (ot[i]) *insert(oe[i].key.data()) = oe[i].val;
  }
  std::vector<uint8_t> tags_; std::vector<Entry> entries_; size_t count_ = 0;
};

class MockLedger : public VerificationState {   // src/lib.rs:106-201
 public:
  FlatTable<64, Ct64> balances;                 // key = account || asset
  FlatTable<32, uint64_t> nonces;
  std::unordered_map<Bytes32, std::pair<std::vector<Bytes32>, uint8_t>, Key32Hash> multisig;
  bool record_outputs = false;                  // the reference mock drops output ciphertexts (src/lib.rs:166-175); tests keep the last one per key
  FlatTable<64, Ct64> outputs;
  bool wants_output_ciphertexts() const override { return record_outputs; }
  bool set_output_ciphertext(const uint8_t account[32], const uint8_t asset[32], const uint8_t ct[64]) override { memcpy(outputs.insert(key(account, asset).data())->data(), ct, 64); return true; }
  static Ct64 key(const uint8_t a[32], const uint8_t b[32]) { Ct64 k; memcpy(k.data(), a, 32); memcpy(k.data() + 32, b, 32); return k; }
  void set_balance(const uint8_t account[32], const uint8_t asset[32], const uint8_t ct[64]) { memcpy(balances.insert(key(account, asset).data())->data(), ct, 64); }
  void set_nonce(const uint8_t account[32], uint64_t nonce) { *nonces.insert(account) = nonce; }
  bool get_account_balance(const uint8_t account[32], const uint8_t asset[32], Role, uint8_t out_ct[64]) override {
    const Ct64* v = balances.find(key(account, asset).data()); if (!v) return false; memcpy(out_ct, v->data(), 64); return true; }
  bool update_account_balance(const uint8_t account[32], const uint8_t asset[32], const uint8_t new_ct[64], Role) override {
    Ct64* v = balances.find(key(account, asset).data()); if (!v) return false; memcpy(v->data(), new_ct, 64); return true; }
  bool get_account_nonce(const uint8_t account[32], uint64_t* nonce) override { const uint64_t* v = nonces.find(account); if (!v) return false; *nonce = *v; return true; }
  bool update_account_nonce(const uint8_t account[32], uint64_t nonce) override { uint64_t* v = nonces.find(account); if (!v) return false; *v = nonce; return true; }
  bool set_multisig_for_account(const uint8_t account[32], const uint8_t* signers, size_t n, uint8_t threshold) override {
    Bytes32 k; memcpy(k.data(), account, 32);
    if (n == 0) { multisig.erase(k); return true; }
    std::vector<Bytes32> v(n); for (size_t i = 0; i < n; i++) memcpy(v[i].data(), signers + 32 * i, 32);
    multisig[k] = std::make_pair(v, threshold); return true; }
  bool get_multisig_for_account(const uint8_t account[32], std::vector<Bytes32>* signers, uint8_t* threshold, bool* present) override {
    if (multisig.empty()) { *present = false; return true; }
    Bytes32 k; memcpy(k.data(), account, 32); auto it = multisig.find(k); *present = it != multisig.end();
    if (*present) { *signers = it->second.first; *threshold = it->second.second; } return true; }
  void prefetch_account(const uint8_t account[32]) const override { nonces.prefetch(account); }
  void prefetch_balance(const uint8_t account[32], const uint8_t asset[32]) const override { balances.prefetch(key(account, asset).data()); }
};

// BlockchainVerificationState over a device-resident ledger (include/xhe.h, xhe_ledger_*): balances are decompressed points in
// HBM, nonces and multisig settings stay in host tables like the mock's.
class DeviceLedgerState : public VerificationState {
 public:
  xhe_ctx* ctx; xhe_ledger* led = nullptr;
  FlatTable<32, uint64_t> nonces;
  std::unordered_map<Bytes32, std::pair<std::vector<Bytes32>, uint8_t>, Key32Hash> multisig;
  std::vector<Ct64> keys;                       // insertion order = slot order
  DeviceLedgerState(xhe_ctx* c, size_t capacity) : ctx(c) { if (xhe_ledger_create(c, capacity, &led) != XHE_OK) led = nullptr; }
  ~DeviceLedgerState() override { if (led) xhe_ledger_destroy(led); }
  bool import_records(const uint8_t* recs, size_t n) {      // pk[32] asset[32] ct[64]; accounts get nonce 0
    std::vector<uint8_t> k(64 * n), c(64 * n), ok(n);
    for (size_t i = 0; i < n; i++) { memcpy(&k[64 * i], recs + 128 * i, 64); memcpy(&c[64 * i], recs + 128 * i + 64, 64); }
    const size_t before = xhe_ledger_size(led);
    if (xhe_ledger_load(led, k.data(), c.data(), n, ok.data()) != XHE_OK) return false;
    for (size_t i = 0; i < n; i++) { bool fresh = false; uint64_t* v = nonces.insert(recs + 128 * i, &fresh); if (fresh) *v = 0; }
    if (xhe_ledger_size(led) != before) { keys.clear(); }   // rebuilt lazily by export_all
    for (size_t i = 0; i < n; i++) { Ct64 kk; memcpy(kk.data(), &k[64 * i], 64); all_keys_.insert(kk.data()); }
    return true;
  }
  size_t export_all(uint8_t* out, size_t cap) {             // records pk asset ct of every balance
    std::vector<uint8_t> k; all_keys_.for_each([&](const uint8_t* key, const uint8_t&) { k.insert(k.end(), key, key + 64); });
    const size_t n = k.size() / 64;
    if (out && 128 * n <= cap) { std::vector<uint8_t> c(64 * n + 64), f(n + 1); xhe_ledger_export(led, k.data(), n, c.data(), f.data()); for (size_t i = 0; i < n; i++) { memcpy(out + 128 * i, &k[64 * i], 64); memcpy(out + 128 * i + 64, &c[64 * i], 64); } }
    return n;
  }
  bool get_account_balance(const uint8_t account[32], const uint8_t asset[32], Role, uint8_t out_ct[64]) override {
    Ct64 k = MockLedger::key(account, asset); uint8_t f = 0; return xhe_ledger_export(led, k.data(), 1, out_ct, &f) == XHE_OK && f; }
  bool update_account_balance(const uint8_t account[32], const uint8_t asset[32], const uint8_t new_ct[64], Role) override {
    Ct64 k = MockLedger::key(account, asset); if (xhe_ledger_slot(led, k.data()) == 0xFFFFFFFFu) return false; uint8_t ok = 0; return xhe_ledger_load(led, k.data(), new_ct, 1, &ok) == XHE_OK && ok; }
  bool get_account_nonce(const uint8_t account[32], uint64_t* nonce) override { const uint64_t* v = nonces.find(account); if (!v) return false; *nonce = *v; return true; }
  bool update_account_nonce(const uint8_t account[32], uint64_t nonce) override { uint64_t* v = nonces.find(account); if (!v) return false; *v = nonce; return true; }
  bool set_multisig_for_account(const uint8_t account[32], const uint8_t* signers, size_t n, uint8_t threshold) override {
    Bytes32 k; memcpy(k.data(), account, 32);
    if (n == 0) { multisig.erase(k); return true; }
    std::vector<Bytes32> v(n); for (size_t i = 0; i < n; i++) memcpy(v[i].data(), signers + 32 * i, 32);
    multisig[k] = std::make_pair(v, threshold); return true; }
  bool get_multisig_for_account(const uint8_t account[32], std::vector<Bytes32>* signers, uint8_t* threshold, bool* present) override {
    if (multisig.empty()) { *present = false; return true; }
    Bytes32 k; memcpy(k.data(), account, 32); auto it = multisig.find(k); *present = it != multisig.end();
    if (*present) { *signers = it->second.first; *threshold = it->second.second; } return true; }
  void prefetch_account(const uint8_t account[32]) const override { nonces.prefetch(account); }
  xhe_ledger* device_ledger() override { return led; }
  bool device_slot(const uint8_t account[32], const uint8_t asset[32], uint32_t* slot) override { Ct64 k = MockLedger::key(account, asset); uint32_t s = xhe_ledger_slot(led, k.data()); if (s == 0xFFFFFFFFu) return false; *slot = s; return true; }
 private:
  FlatTable<64, uint8_t> all_keys_;
};

struct TransferView { const uint8_t *asset, *dest, *commitment, *sender_handle, *receiver_handle, *proof, *extra; uint32_t extra_len; bool has_extra; };
struct TxView {   // parsed xtx1 blob (pointers into the caller's buffer)
  const uint8_t* blob = nullptr; size_t len = 0;
  uint8_t version = 0, type = 0, n_sc = 0; int n_ms = -1; uint32_t count = 0, aux = 0, rp_len = 0; const uint8_t* source = nullptr; uint64_t fee = 0, nonce = 0;
  std::vector<TransferView> transfers; const uint8_t *body = nullptr, *rp = nullptr, *sc = nullptr, *ms = nullptr, *sig = nullptr; size_t body_len = 0;
  int parse(const uint8_t* blob, size_t len);                     // XHE_OK / XHE_ERR_PARSE
  void to_bytes(std::vector<uint8_t>& out, size_t* multisig_index) const;   // src/tx/verify.rs:623-688
  uint32_t n_transfers() const { return type == 0 ? count : 0; }
};

struct BatchTimings { bool used_fast_path = false; double parse_ms = 0, resolve_ms = 0, transcript_ms = 0, device_ms = 0, finish_ms = 0, total_ms = 0; uint64_t keccak_f = 0; };
struct BatchOptions { int threads = 0;
                      /* Personalisation of the per-proof random batch factors (reference: Scalar::random per proof, src/proofs.rs:181,326).
                       * The factors MUST be unpredictable to whoever produced the proofs, so 32 bytes of OS entropy are always folded in;
                       * deterministic_seed (tests only) uses rng_seed alone so that a run can be replayed. */
                      const uint8_t* rng_seed = nullptr; size_t rng_seed_len = 0; bool deterministic_seed = false;
                      bool apply_state = true;
                      uint8_t* partial_out = nullptr; /* 64 B: shard mode, see verify_batch */
                      /* shard mode on the WHOLE batch: this rank verifies transactions [shard_lo, min(shard_hi, n)) of blobs[0..n) and reads
                       * the earlier ones only to advance the balance chains (and multisig settings) its own transactions depend on */
                      size_t shard_lo = 0, shard_hi = (size_t)-1;
                      const void* key_index = nullptr; /* xheh_batch_index_build of the same batch (optional): earlier shards are searched through their key digests */
                      bool device_fiat_shamir = false; /* transcripts, batch factors and main-signature hashes on the GPU */
                      bool host_dry_run = false; /* diagnostics: run the host phases of the fast path only (no device, nothing verified; the call returns XHE_E_ARG with the timings filled) */
                      bool fast_path = false; /* optimistic device-layout path first (implies device Fiat-Shamir); the exact path decides on any failure */ };

// Transaction::verify_batch.  Returns XHE_OK or the verdict code; *fail_index = first failing tx (-1 for the two
// batch-level MSM checks, as in the reference where those errors carry no tx).
// Nothing is written to `state` before the verdict is known: nonce and multisig writes are staged like the balance updates
// and applied together on accept (the reference mutates as it goes and tells callers to discard the state of a failed batch).
// Shard mode (opt.partial_out != nullptr, multi-GPU): the two identity decisions are NOT taken here; the encodings of this
// shard's partial sigma / range sums are returned (sigma || range) for the caller to combine across ranks, and the state
// updates are held back until commit_pending().  With shard_lo / shard_hi the call takes the WHOLE batch: *fail_index is
// then an index into the whole batch, and balance chains that start in earlier shards are followed.
int verify_batch(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n, VerificationState& state, const BatchOptions& opt, long* fail_index, BatchTimings* timings);
// apply the balance updates a shard-mode verify_batch held back (after the cross-rank decision accepted the batch)
int commit_pending(xhe_ctx* ctx, VerificationState& state);
size_t export_pending(void* pending, uint8_t* out, size_t cap);
// Transaction::apply_without_verify for a list of txs applied in order (balance updates only; config 4 shape)
int apply_without_verify(xhe_ctx* ctx, const uint8_t* const* blobs, const size_t* lens, size_t n, VerificationState& state);

}  // namespace xhe_host
