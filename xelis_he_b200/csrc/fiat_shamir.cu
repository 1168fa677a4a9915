// fiat_shamir.cu -- device-side Fiat-Shamir: the Merlin transcripts of src/tx/verify.rs:146-158,315-326,
// 378-393,397-399,420-424 + src/proofs.rs:142-161,291-302 + the bulletproofs verification transcript, the
// per-proof random batch factors, and the SHA3-512 signature hash of src/elgamal.rs:53-65, one thread per transaction.
// Challenges are reduced mod l and written straight into the device scalar arrays the weight kernels read, so in this mode
// the host neither hashes nor uploads challenges.  Host mode (north_star's division of labour) remains the default of the
// C ABI; both modes produce identical bytes (tests/test_gpu_verify.py::test_device_fiat_shamir_matches_host).
#include "xhe_internal.cuh"
using namespace xhe;

namespace {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// (kept out of line, like the sponge operations below: one thread per transaction walks ~100 transcript operations, and with
// everything inlined the kernel grew to ~120k instructions and stalled on instruction fetch)
__device__ __noinline__ void keccak_f1600(uint64_t* st) {
  const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
                           0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                           0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
                           0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  uint64_t s[25];
#pragma unroll
  for (int i = 0; i < 25; i++) s[i] = st[i];
#pragma unroll 1
  for (int r = 0; r < 24; r++) {
    uint64_t c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20], c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21], c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22],
             c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23], c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
    uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
    uint64_t b[25];
    b[0] = s[0] ^ d0;               b[10] = rotl64(s[1] ^ d1, 1);   b[20] = rotl64(s[2] ^ d2, 62);  b[5] = rotl64(s[3] ^ d3, 28);   b[15] = rotl64(s[4] ^ d4, 27);
    b[16] = rotl64(s[5] ^ d0, 36);  b[1] = rotl64(s[6] ^ d1, 44);   b[11] = rotl64(s[7] ^ d2, 6);   b[21] = rotl64(s[8] ^ d3, 55);  b[6] = rotl64(s[9] ^ d4, 20);
    b[7] = rotl64(s[10] ^ d0, 3);   b[17] = rotl64(s[11] ^ d1, 10); b[2] = rotl64(s[12] ^ d2, 43);  b[12] = rotl64(s[13] ^ d3, 25); b[22] = rotl64(s[14] ^ d4, 39);
    b[23] = rotl64(s[15] ^ d0, 41); b[8] = rotl64(s[16] ^ d1, 45);  b[18] = rotl64(s[17] ^ d2, 15); b[3] = rotl64(s[18] ^ d3, 21);  b[13] = rotl64(s[19] ^ d4, 8);
    b[14] = rotl64(s[20] ^ d0, 18); b[24] = rotl64(s[21] ^ d1, 2);  b[9] = rotl64(s[22] ^ d2, 61);  b[19] = rotl64(s[23] ^ d3, 56);  b[4] = rotl64(s[24] ^ d4, 14);
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
      s[y + 0] = b[y + 0] ^ (~b[y + 1] & b[y + 2]); s[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]); s[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
      s[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y + 0]); s[y + 4] = b[y + 4] ^ (~b[y + 0] & b[y + 1]);
    }
    s[0] ^= RC[r];
  }
#pragma unroll
  for (int i = 0; i < 25; i++) st[i] = s[i];
}

// up to eight message bytes as one little-endian word: the loads are independent of each other (one memory round trip),
// word loads when the pointer allows it
__device__ __forceinline__ uint64_t load_le(const uint8_t* d, uint32_t nb) {
  if (nb == 8 && (((uintptr_t)d) & 3) == 0) { const uint32_t* w = (const uint32_t*)d; return (uint64_t)w[0] | ((uint64_t)w[1] << 32); }
  uint64_t v = 0;
#pragma unroll
  for (uint32_t i = 0; i < 8; i++) if (i < nb) v |= (uint64_t)d[i] << (8 * i);
  return v;
}

struct Sponge {   // sponge over a local-memory state; absorbs and squeezes up to eight bytes per state access
  uint64_t st[25]; uint32_t pos, rate;
  __device__ void init(uint32_t r) { for (int i = 0; i < 25; i++) st[i] = 0; pos = 0; rate = r; }
  __device__ __forceinline__ void xor_byte(uint32_t p, uint8_t v) { st[p >> 3] ^= (uint64_t)v << ((p & 7) * 8); }
  __device__ __forceinline__ uint8_t get_byte(uint32_t p) const { return (uint8_t)(st[p >> 3] >> ((p & 7) * 8)); }
  // XOR the nb <= 8 low bytes of w (upper bytes zero) into the state at byte position p
  __device__ __forceinline__ void xor_word(uint32_t p, uint64_t w, uint32_t nb) {
    const uint32_t o = p & 7, sh = o * 8;
    st[p >> 3] ^= w << sh;
    if (o + nb > 8) st[(p >> 3) + 1] ^= w >> (64 - sh);
  }
  // the nb <= 8 state bytes at position p as a little-endian word
  __device__ __forceinline__ uint64_t get_word(uint32_t p, uint32_t nb) const {
    const uint32_t o = p & 7, sh = o * 8;
    uint64_t v = st[p >> 3] >> sh;
    if (o + nb > 8) v |= st[(p >> 3) + 1] << (64 - sh);
    return nb < 8 ? v & ((1ull << (8 * nb)) - 1) : v;
  }
  __device__ __noinline__ void absorb(const uint8_t* d, uint32_t n) {
    while (n) {
      uint32_t nb = min(min(8u, n), rate - pos);
      xor_word(pos, load_le(d, nb), nb);
      pos += nb; d += nb; n -= nb;
      if (pos == rate) { keccak_f1600(st); pos = 0; }
    }
  }
  __device__ void finish(uint8_t dom) { xor_byte(pos, dom); xor_byte(rate - 1, 0x80); keccak_f1600(st); pos = 0; }
  __device__ __noinline__ void squeeze(uint8_t* o, uint32_t n) {
    while (n) {
      if (pos == rate) { keccak_f1600(st); pos = 0; }
      uint32_t nb = min(min(8u, n), rate - pos);
      uint64_t v = get_word(pos, nb);
      for (uint32_t i = 0; i < nb; i++) o[i] = (uint8_t)(v >> (8 * i));
      pos += nb; o += nb; n -= nb;
    }
  }
};

struct Merlin {   // STROBE-128 / "Merlin v1.0"
  Sponge s; uint8_t pos_begin;
  enum { R = 166, F_I = 1, F_A = 2, F_C = 4, F_M = 16, F_K = 32 };
  __device__ __noinline__ void run_f() { s.xor_byte(s.pos, pos_begin); s.xor_byte(s.pos + 1, 0x04); s.xor_byte(R + 1, 0x80); keccak_f1600(s.st); s.pos = 0; pos_begin = 0; }
  __device__ __noinline__ void absorb(const uint8_t* d, uint32_t n) {
    while (n) {
      uint32_t nb = min(min(8u, n), (uint32_t)R - s.pos);
      s.xor_word(s.pos, load_le(d, nb), nb);
      s.pos += nb; d += nb; n -= nb;
      if (s.pos == R) run_f();
    }
  }
  __device__ __noinline__ void begin_op(uint8_t flags) {
    uint8_t h[2] = {pos_begin, flags}; pos_begin = (uint8_t)(s.pos + 1);
    absorb(h, 2);
    if ((flags & (F_C | F_K)) && s.pos != 0) run_f();
  }
  __device__ void meta_ad(const uint8_t* d, uint32_t n, bool more) { if (!more) begin_op(F_M | F_A); absorb(d, n); }
  __device__ void init(const char* label, uint32_t llen) {
    s.init(200); pos_begin = 0;
    const uint8_t hdr[18] = {1, R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    for (int i = 0; i < 18; i++) s.xor_byte(i, hdr[i]);
    keccak_f1600(s.st);
    const uint8_t m[11] = {'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0'};
    meta_ad(m, 11, false);
    const uint8_t ds[7] = {'d', 'o', 'm', '-', 's', 'e', 'p'};
    append(ds, 7, (const uint8_t*)label, llen);
  }
  __device__ __noinline__ void append(const uint8_t* label, uint32_t llen, const uint8_t* msg, uint32_t n) {
    uint8_t le[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    meta_ad(label, llen, false); meta_ad(le, 4, true);
    begin_op(F_A); absorb(msg, n);
  }
  __device__ void append_u64(const uint8_t* label, uint32_t llen, uint64_t v) { uint8_t le[8]; for (int i = 0; i < 8; i++) le[i] = (uint8_t)(v >> (8 * i)); append(label, llen, le, 8); }
  // 64 challenge bytes reduced mod l (ProtocolTranscript::challenge_scalar, src/transcript.rs:46-51) -> 8 words
  __device__ __noinline__ void challenge_scalar(const uint8_t* label, uint32_t llen, uint32_t* out8) {
    uint8_t le[4] = {64, 0, 0, 0};
    meta_ad(label, llen, false); meta_ad(le, 4, true);
    begin_op(F_I | F_A | F_C);
    // PRF output: the state bytes are read out and cleared (STROBE squeeze with the C flag), eight at a time
    uint8_t b[64]; uint32_t got = 0;
    while (got < 64) {
      uint32_t nb = min(8u, min(64u - got, (uint32_t)R - s.pos));
      uint64_t v = s.get_word(s.pos, nb);
      s.xor_word(s.pos, v, nb);
      for (uint32_t i = 0; i < nb; i++) b[got + i] = (uint8_t)(v >> (8 * i));
      got += nb; s.pos += nb;
      if (s.pos == R) run_f();
    }
    sc lo = sc_frombytes(b), hi = sc_frombytes(b + 32);
    sc r = sc_reduce512(lo, hi);
    for (int i = 0; i < 8; i++) out8[i] = r.v[i];
  }
};

#define LBL(str) (const uint8_t*)(str), (uint32_t)(sizeof(str) - 1)
__device__ __forceinline__ uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
__device__ __forceinline__ uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }

// plan words per tx: 0 eq_begin, 1 val_begin, 2 rp_slot (0xffffffff none), 3 rp_chal_begin, 4 main signature slot (0xffffffff none), 5 flags (bit0: proofs stage reached)
__global__ void __launch_bounds__(64) k_fiat_shamir(const uint8_t* __restrict__ blobs, const unsigned long long* __restrict__ blob_off, const uint32_t* __restrict__ plan, uint32_t plan_stride, uint32_t n_tx,
                                                    const uint8_t* __restrict__ seed32, unsigned long long index_base, uint32_t* __restrict__ eq_sc, uint32_t* __restrict__ val_sc, uint32_t* __restrict__ rp_sc,
                                                    uint32_t* __restrict__ rp_chal, const uint32_t* __restrict__ rp_m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tx) return;
  const uint32_t* P = plan + plan_stride * (size_t)i;
  if (!(P[5] & 1u)) return;
  const uint8_t* b = blobs + blob_off[i];
  const uint8_t version = b[0], type = b[1], n_sc = b[2];
  const uint32_t count = rd32(b + 4), aux = rd32(b + 8), rp_len = rd32(b + 12);
  const uint8_t* source = b + 16; const uint64_t fee = rd64(b + 48), nonce = rd64(b + 56);
  // locate the sections (same framing as host TxView::parse)
  const uint8_t* body = b + 64; const uint8_t* p = body;
  if (type == 0) { for (uint32_t t = 0; t < count; t++) { uint32_t el = rd32(p + 320); p += 324 + (el == 0xFFFFFFFFu ? 0 : el); } }
  else if (type == 1) p += 40;
  else if (type == 2) { p += 32 + 40 * (size_t)count; for (uint32_t q = 0; q < 2 * aux; q++) p += 4 + rd32(p); }
  else if (type == 3) p += aux;
  else p += 32 * (size_t)count;
  const uint8_t* rp = p; const uint8_t* scs = rp + rp_len;
  // per-proof random batch factors: SHAKE256("xhe-batch-factors" || seed || index of the tx in the whole batch), 32 bytes each, top nibble cleared
  Sponge rng; rng.init(136);
  { const char tag[] = "xhe-batch-factors"; rng.absorb((const uint8_t*)tag, 17); rng.absorb(seed32, 32); unsigned long long idx = index_base + i; rng.absorb((const uint8_t*)&idx, 8); rng.finish(0x1f); }
  auto rnd_scalar = [&](uint32_t* out8) { uint8_t r[32]; rng.squeeze(r, 32); r[31] &= 0x0f; sc v = sc_frombytes(r); for (int q = 0; q < 8; q++) out8[q] = v.v[q]; };
  Merlin T; T.init("transaction-proof", 17);
  T.append_u64(LBL("version"), version); T.append(LBL("source_pubkey"), source, 32); T.append_u64(LBL("fee"), fee); T.append_u64(LBL("nonce"), nonce);
  for (uint32_t q = 0; q < n_sc; q++) {
    const uint8_t* asset = scs + 256 * (size_t)q; const uint8_t* proof = asset + 64;
    T.append(LBL("dom-sep"), LBL("new-commitment-proof")); T.append(LBL("new_source_commitment_asset"), asset, 32); T.append(LBL("new_source_commitment"), asset + 32, 32);
    T.append(LBL("dom-sep"), LBL("equality-proof")); T.append(LBL("Y_0"), proof, 32); T.append(LBL("Y_1"), proof + 32, 32); T.append(LBL("Y_2"), proof + 64, 32);
    uint32_t* o = eq_sc + 48 * (size_t)(P[0] + q);
    T.challenge_scalar(LBL("c"), o + 24);
    T.append(LBL("z_s"), proof + 96, 32); T.append(LBL("z_x"), proof + 128, 32); T.append(LBL("z_r"), proof + 160, 32);
    T.challenge_scalar(LBL("w"), o + 32); rnd_scalar(o + 40);
  }
  if (type == 0) {
    const uint8_t* tp = body;
    for (uint32_t t = 0; t < count; t++) {
      uint32_t el = rd32(tp + 320); const uint8_t* proof = tp + 160;
      T.append(LBL("dom-sep"), LBL("transfer-proof")); T.append(LBL("dest_pubkey"), tp + 32, 32); T.append(LBL("amount_commitment"), tp + 64, 32);
      T.append(LBL("amount_sender_handle"), tp + 96, 32); T.append(LBL("amount_receiver_handle"), tp + 128, 32);
      T.append(LBL("dom-sep"), LBL("validity-proof")); T.append(LBL("Y_0"), proof, 32); T.append(LBL("Y_1"), proof + 32, 32); T.append(LBL("Y_2"), proof + 64, 32);
      uint32_t* o = val_sc + 40 * (size_t)(P[1] + t);
      T.challenge_scalar(LBL("c"), o + 16);
      T.append(LBL("z_r"), proof + 96, 32); T.append(LBL("z_x"), proof + 128, 32);
      T.challenge_scalar(LBL("w"), o + 24); rnd_scalar(o + 32);
      tp += 324 + (el == 0xFFFFFFFFu ? 0 : el);
    }
  } else if (type == 1) {
    T.append(LBL("dom-sep"), LBL("burn-proof")); T.append(LBL("asset"), body, 32); T.append_u64(LBL("amount"), rd64(body + 32));
  } else if (type == 4) {
    T.append(LBL("dom-sep"), LBL("multisig-proof")); T.append_u64(LBL("threshold"), aux);
    for (uint32_t q = 0; q < count; q++) T.append(LBL("signer"), body + 32 * (size_t)q, 32);
  }
  if (P[2] != 0xFFFFFFFFu) {
    const uint32_t m = rp_m[P[2]], lg = (rp_len / 32 - 9) / 2;
    T.append(LBL("dom-sep"), LBL("rangeproof v1")); T.append_u64(LBL("n"), 64); T.append_u64(LBL("m"), m);
    for (uint32_t q = 0; q < n_sc; q++) T.append(LBL("V"), scs + 256 * (size_t)q + 32, 32);
    uint32_t nv = n_sc;
    if (type == 0) { const uint8_t* tp = body; for (uint32_t t = 0; t < count; t++) { T.append(LBL("V"), tp + 64, 32); uint32_t el = rd32(tp + 320); tp += 324 + (el == 0xFFFFFFFFu ? 0 : el); nv++; } }
    { uint8_t z[32]; for (int q = 0; q < 32; q++) z[q] = 0; for (; nv < m; nv++) T.append(LBL("V"), z, 32); }
    uint32_t* ch = rp_chal + 8 * (size_t)P[3];
    T.append(LBL("A"), rp, 32); T.append(LBL("S"), rp + 32, 32);
    T.challenge_scalar(LBL("y"), ch); T.challenge_scalar(LBL("z"), ch + 8);
    T.append(LBL("T_1"), rp + 64, 32); T.append(LBL("T_2"), rp + 96, 32);
    T.challenge_scalar(LBL("x"), ch + 16);
    T.append(LBL("t_x"), rp + 128, 32); T.append(LBL("t_x_blinding"), rp + 160, 32); T.append(LBL("e_blinding"), rp + 192, 32);
    T.challenge_scalar(LBL("w"), ch + 24);
    T.append(LBL("dom-sep"), LBL("ipp v1")); T.append_u64(LBL("n"), 64ull * m);
    for (uint32_t q = 0; q < lg; q++) { T.append(LBL("L"), rp + 224 + 64 * (size_t)q, 32); T.append(LBL("R"), rp + 224 + 64 * (size_t)q + 32, 32); T.challenge_scalar(LBL("u"), ch + 32 + 8 * q); }
    uint32_t* o = rp_sc + 56 * (size_t)P[2];
    rnd_scalar(o + 40); rnd_scalar(o + 48);
  }
}

// main signature check: e' = SHA3-512(pk || to_bytes(tx) || r) mod l ; ok = (e' == e)    (src/elgamal.rs:38-42,53-65; to_bytes: src/tx/verify.rs:623-688)
// r is the LAST thing absorbed, so the check is split: k_sig_hash_prefix absorbs pk || to_bytes(tx) (about 19 of the 20
// permutations of a one-transfer transaction) as soon as the blobs are on the device, beside the signature group operations;
// k_sig_hash_final only absorbs r, pads and compares.  state = 25 lanes + the byte position, 26 x 8 bytes per signature slot.
__global__ void __launch_bounds__(64) k_sig_hash_prefix(const uint8_t* __restrict__ blobs, const unsigned long long* __restrict__ blob_off, const uint32_t* __restrict__ plan, uint32_t plan_stride, uint32_t n_tx,
                                                        unsigned long long* __restrict__ state) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tx) return;
  uint32_t slot = plan[plan_stride * (size_t)i + 4];
  if (slot == 0xFFFFFFFFu) return;
  const uint8_t* b = blobs + blob_off[i];
  const uint8_t type = b[1], n_sc = b[2]; const int n_ms = b[3] == 0xFF ? -1 : b[3];
  const uint32_t count = rd32(b + 4), aux = rd32(b + 8), rp_len = rd32(b + 12);
  Sponge h; h.init(72);
  h.absorb(b + 16, 32);                                  // pk (compressed source)
  h.absorb(b, 1); h.absorb(b + 16, 32);                  // version, source
  { uint8_t be[16]; for (int q = 0; q < 8; q++) { be[q] = b[48 + 7 - q]; be[8 + q] = b[56 + 7 - q]; } h.absorb(be, 16); }   // fee, nonce big-endian
  const uint8_t* p = b + 64;
  if (type == 0) { for (uint32_t t = 0; t < count; t++) { uint32_t el = rd32(p + 320); h.absorb(p, 160); if (el != 0xFFFFFFFFu) h.absorb(p + 324, el); h.absorb(p + 160, 160); p += 324 + (el == 0xFFFFFFFFu ? 0 : el); } }
  else if (type == 1) { h.absorb(p, 32); uint8_t be[8]; for (int q = 0; q < 8; q++) be[q] = p[32 + 7 - q]; h.absorb(be, 8); p += 40; }
  else if (type == 2) { h.absorb(p, 32); p += 32; for (uint32_t q = 0; q < count; q++) { h.absorb(p, 32); uint8_t be[8]; for (int z = 0; z < 8; z++) be[z] = p[32 + 7 - z]; h.absorb(be, 8); p += 40; }
                        for (uint32_t q = 0; q < 2 * aux; q++) { uint32_t l = rd32(p); h.absorb(p + 4, l); p += 4 + l; } }
  else if (type == 3) { h.absorb(p, aux); p += aux; }
  else { uint8_t th = (uint8_t)aux; h.absorb(&th, 1); h.absorb(p, 32 * count); p += 32 * (size_t)count; }
  h.absorb(p, rp_len); p += rp_len;
  h.absorb(p, 256 * (uint32_t)n_sc); p += 256 * (size_t)n_sc;
  if (n_ms > 0) h.absorb(p, 65 * (uint32_t)n_ms);
  unsigned long long* o = state + 26 * (size_t)slot;
  for (int q = 0; q < 25; q++) o[q] = h.st[q];
  o[25] = h.pos;
}
__global__ void __launch_bounds__(64) k_sig_hash_final(const uint32_t* __restrict__ plan, uint32_t plan_stride, uint32_t n_tx, const unsigned long long* __restrict__ state,
                                                       const uint8_t* __restrict__ sig_r, const uint32_t* __restrict__ sig_e, uint8_t* __restrict__ sig_ok) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tx) return;
  uint32_t slot = plan[plan_stride * (size_t)i + 4];
  if (slot == 0xFFFFFFFFu) return;
  Sponge h; h.rate = 72;
  const unsigned long long* in = state + 26 * (size_t)slot;
  for (int q = 0; q < 25; q++) h.st[q] = in[q];
  h.pos = (uint32_t)in[25];
  h.absorb(sig_r + 32 * (size_t)slot, 32);
  h.finish(0x06);
  uint8_t d[64]; h.squeeze(d, 64);
  sc e2 = sc_reduce512(sc_frombytes(d), sc_frombytes(d + 32));
  bool ok = true;
  for (int q = 0; q < 8; q++) ok = ok && (e2.v[q] == sig_e[8 * (size_t)slot + q]);
  sig_ok[slot] = ok ? 1 : 0;
}

}  // namespace

int32_t xhe_launch_fiat_shamir(xhe_ctx* ctx, const uint8_t* d_blobs, const unsigned long long* d_off, const uint32_t* d_plan, uint32_t plan_stride, uint32_t n_tx, const uint8_t* d_seed, unsigned long long index_base,
                               uint32_t* d_eq_sc, uint32_t* d_val_sc, uint32_t* d_rp_sc, uint32_t* d_rp_chal, const uint32_t* d_rp_m) {
  if (!n_tx) return XHE_OK;
  XheTimed t(ctx, "k_fiat_shamir", 0);
  k_fiat_shamir<<<(n_tx + 63) / 64, 64, 0, ctx->stream>>>(d_blobs, d_off, d_plan, plan_stride, n_tx, d_seed, index_base, d_eq_sc, d_val_sc, d_rp_sc, d_rp_chal, d_rp_m);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_launch_sig_hash_prefix(xhe_ctx* ctx, const uint8_t* d_blobs, const unsigned long long* d_off, const uint32_t* d_plan, uint32_t plan_stride, uint32_t n_tx, unsigned long long* d_state) {
  if (!n_tx) return XHE_OK;
  XheTimed t(ctx, "k_sig_hash_prefix", 0);
  k_sig_hash_prefix<<<(n_tx + 63) / 64, 64, 0, ctx->stream>>>(d_blobs, d_off, d_plan, plan_stride, n_tx, d_state);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_launch_sig_hash_final(xhe_ctx* ctx, const uint32_t* d_plan, uint32_t plan_stride, uint32_t n_tx, const unsigned long long* d_state, const uint8_t* d_sig_r, const uint32_t* d_sig_e, uint8_t* d_sig_ok) {
  if (!n_tx) return XHE_OK;
  XheTimed t(ctx, "k_sig_hash", 0);
  k_sig_hash_final<<<(n_tx + 63) / 64, 64, 0, ctx->stream>>>(d_plan, plan_stride, n_tx, d_state, d_sig_r, d_sig_e, d_sig_ok);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// fast path: device-side batch layout (SURVEY.md 8 f.2 in spirit).  The host uploads the raw xtx1 blobs plus 8 plan words
// per transaction (prefix sums over the transactions' shapes); this kernel builds the point table and every per-proof array
// from the blob bytes, so the host never touches proof bytes.  Per-tx point layout (region A, starting at plan[6]):
//   [source] [k x (C, D_sender, D_receiver)] [a x new commitment] [a x (Y0,Y1,Y2)] [k x dest] [k x (Y0,Y1,Y2)] [A,S,T1,T2] [L x lg] [R x lg]
// plan words: 0 eq_begin, 1 val_begin, 2 rp slot, 3 rp challenge offset, 4 signature slot, 5 flags, 6 point base, 7 first balance-op index
// An all-zero Y encoding (TranscriptError::IdentityPoint, src/transcript.rs:73-84) raises bit 0 of *viol and of the transaction's
// flag byte; an all-zero A / S / T / L / R (rejected inside bulletproofs' verifier) raises bit 4 of *viol / bit 3 of the flag byte.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ bool copy32_is_zero(uint8_t* dst, const uint8_t* src) {
  uint32_t acc = 0;
  if ((((uintptr_t)src) & 3) == 0) {
    const uint32_t* s4 = (const uint32_t*)src; uint32_t* d4 = (uint32_t*)dst;
#pragma unroll
    for (int i = 0; i < 8; i++) { uint32_t w = s4[i]; d4[i] = w; acc |= w; }
  } else {
    for (int i = 0; i < 32; i++) { uint8_t b = src[i]; dst[i] = b; acc |= b; }
  }
  return acc == 0;
}
__device__ __forceinline__ void copy_words(uint32_t* dst, const uint8_t* src, int nwords) {
  for (int i = 0; i < nwords; i++) dst[i] = rd32(src + 4 * i);
}

__global__ void __launch_bounds__(64) k_layout(const uint8_t* __restrict__ blobs, const unsigned long long* __restrict__ blob_off, const uint32_t* __restrict__ plan, uint32_t n_tx,
                                               uint32_t n_points, uint8_t* __restrict__ enc, uint32_t* __restrict__ sig_idx /* eq then val point indices */, uint32_t n_eq_total,
                                               uint32_t* __restrict__ eq_sc, uint32_t* __restrict__ val_sc, uint32_t* __restrict__ rp_sc, uint32_t* __restrict__ range_idx,
                                               const uint32_t* __restrict__ rp_pt_off, uint32_t* __restrict__ sig_s, uint32_t* __restrict__ sig_e, uint32_t* __restrict__ sig_pk,
                                               uint32_t* __restrict__ viol, uint8_t* __restrict__ tx_flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tx) return;
  const uint32_t* P = plan + 8 * (size_t)i;
  const uint8_t* b = blobs + blob_off[i];
  const uint8_t type = b[1]; const uint32_t a = b[2];
  const uint32_t count = rd32(b + 4), aux = rd32(b + 8), rp_len = rd32(b + 12);
  const uint32_t k = type == 0 ? count : 0, lg = (rp_len / 32 - 9) / 2;
  const uint8_t* body = b + 64; const uint8_t* p = body;
  if (type == 0) { for (uint32_t t = 0; t < count; t++) { uint32_t el = rd32(p + 320); p += 324 + (el == 0xFFFFFFFFu ? 0 : el); } }
  else if (type == 1) p += 40;
  else if (type == 2) { p += 32 + 40 * (size_t)count; for (uint32_t q = 0; q < 2 * aux; q++) p += 4 + rd32(p); }
  else if (type == 3) p += aux;
  else p += 32 * (size_t)count;
  const uint8_t* rp = p; const uint8_t* scs = rp + rp_len; const uint8_t* sig = scs + 256 * (size_t)a + (b[3] == 0xFF ? 0 : 65 * (size_t)b[3]);
  const uint32_t base = P[6];
  const uint32_t iSrc = base, iT = base + 1, iN = iT + 3 * k, iEqY = iN + a, iDest = iEqY + 3 * a, iValY = iDest + k, iRp = iValY + 3 * k;
  bool bad = false;
  copy32_is_zero(enc + 32 * (size_t)iSrc, b + 16);
  { const uint8_t* tp = body;
    for (uint32_t t = 0; t < k; t++) {
      uint32_t el = rd32(tp + 320);
      copy32_is_zero(enc + 32 * (size_t)(iT + 3 * t), tp + 64); copy32_is_zero(enc + 32 * (size_t)(iT + 3 * t + 1), tp + 96); copy32_is_zero(enc + 32 * (size_t)(iT + 3 * t + 2), tp + 128);
      copy32_is_zero(enc + 32 * (size_t)(iDest + t), tp + 32);
      const uint8_t* proof = tp + 160;
      for (int y = 0; y < 3; y++) bad |= copy32_is_zero(enc + 32 * (size_t)(iValY + 3 * t + y), proof + 32 * y);
      // validity proof t: points C, Y0, P_dest, D_dest, Y1, P_src, D_src, Y2 ; scalars z_r, z_x (c, w, bf come from the transcript kernel)
      uint32_t* vi = sig_idx + 7 * (size_t)n_eq_total + 8 * (size_t)(P[1] + t);
      vi[0] = iT + 3 * t; vi[1] = iValY + 3 * t; vi[2] = iDest + t; vi[3] = iT + 3 * t + 2; vi[4] = iValY + 3 * t + 1; vi[5] = iSrc; vi[6] = iT + 3 * t + 1; vi[7] = iValY + 3 * t + 2;
      copy_words(val_sc + 40 * (size_t)(P[1] + t), proof + 96, 16);
      tp += 324 + (el == 0xFFFFFFFFu ? 0 : el);
    } }
  for (uint32_t q = 0; q < a; q++) {
    const uint8_t* sc = scs + 256 * (size_t)q; const uint8_t* proof = sc + 64;
    copy32_is_zero(enc + 32 * (size_t)(iN + q), sc + 32);
    for (int y = 0; y < 3; y++) bad |= copy32_is_zero(enc + 32 * (size_t)(iEqY + 3 * q + y), proof + 32 * y);
    // eq proof q: P_src, Y0, D_src, C_src, Y1, C_dst, Y2 ; balance-chain outputs live at n_points + op index (C op, then D op per touch)
    uint32_t* ei = sig_idx + 7 * (size_t)(P[0] + q); const uint32_t op = P[7] + 2 * q;
    ei[0] = iSrc; ei[1] = iEqY + 3 * q; ei[2] = n_points + op + 1; ei[3] = n_points + op; ei[4] = iEqY + 3 * q + 1; ei[5] = iN + q; ei[6] = iEqY + 3 * q + 2;
    copy_words(eq_sc + 48 * (size_t)(P[0] + q), proof + 96, 24);
  }
  bool bad_rp = false;   // an identity-encoded A / S / T / L / R fails inside the range-proof batch (RangeProof, after the sigma check), not at this transaction
  { // range proof: points A,S,T1,T2,L[lg],R[lg],V[m]; scalars t_x,t_x_blinding,e_blinding,a,b
    for (int q = 0; q < 4; q++) bad_rp |= copy32_is_zero(enc + 32 * (size_t)(iRp + q), rp + 32 * q);
    for (uint32_t q = 0; q < lg; q++) { bad_rp |= copy32_is_zero(enc + 32 * (size_t)(iRp + 4 + q), rp + 224 + 64 * (size_t)q); bad_rp |= copy32_is_zero(enc + 32 * (size_t)(iRp + 4 + lg + q), rp + 224 + 64 * (size_t)q + 32); }
    uint32_t* ri = range_idx + rp_pt_off[P[2]]; uint32_t m = 1; while (m < a + k) m <<= 1;
    for (uint32_t q = 0; q < 4 + 2 * lg; q++) ri[q] = iRp + q;
    for (uint32_t q = 0; q < a; q++) ri[4 + 2 * lg + q] = iN + q;
    for (uint32_t q = 0; q < k; q++) ri[4 + 2 * lg + a + q] = iT + 3 * q;
    for (uint32_t q = a + k; q < m; q++) ri[4 + 2 * lg + q] = 0;
    uint32_t* rs = rp_sc + 56 * (size_t)P[2];
    copy_words(rs, rp + 128, 24); copy_words(rs + 24, rp + rp_len - 64, 16);
  }
  copy_words(sig_s + 8 * (size_t)P[4], sig, 8); copy_words(sig_e + 8 * (size_t)P[4], sig + 32, 8); sig_pk[P[4]] = iSrc;
  tx_flags[i] = (bad ? 1 : 0) | (bad_rp ? 8 : 0);
  if (bad) atomicOr(viol, 1u);
  if (bad_rp) atomicOr(viol, 16u);
}

// per-transaction anomaly bits of the device-layout path (same meaning as xhe_verdict.device_flags bits 0-2): bit 0 was set by
// k_layout; bit 1 = one of the transaction's own points (region A) failed to decompress; bit 2 = its signature hash mismatched
__global__ void __launch_bounds__(128) k_tx_flags(const uint32_t* __restrict__ plan, uint32_t n_tx, uint32_t n_a_end, const uint8_t* __restrict__ pt_ok, const uint8_t* __restrict__ sig_ok, uint8_t* __restrict__ tx_flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tx) return;
  const uint32_t lo = plan[8 * (size_t)i + 6], hi = i + 1 < n_tx ? plan[8 * (size_t)(i + 1) + 6] : n_a_end;
  uint8_t f = tx_flags[i];
  for (uint32_t p = lo; p < hi; p++) if (!pt_ok[p]) { f |= 2; break; }
  if (!sig_ok[plan[8 * (size_t)i + 4]]) f |= 4;
  tx_flags[i] = f;
}
// OR of (flag byte == 0) over n bytes into bit `bit` of *viol
__global__ void __launch_bounds__(256) k_any_zero(const uint8_t* __restrict__ flags, uint32_t n, uint32_t bit, uint32_t* __restrict__ viol) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = i < n && flags[i] == 0;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(viol, 1u << bit);
}
}  // namespace

int32_t xhe_launch_layout(xhe_ctx* ctx, const uint8_t* d_blobs, const unsigned long long* d_off, const uint32_t* d_plan, uint32_t n_tx, uint32_t n_points, uint8_t* d_enc,
                          uint32_t* d_sig_idx, uint32_t n_eq, uint32_t* d_eq_sc, uint32_t* d_val_sc, uint32_t* d_rp_sc, uint32_t* d_range_idx, const uint32_t* d_rp_pt_off,
                          uint32_t* d_sig_s, uint32_t* d_sig_e, uint32_t* d_sig_pk, uint32_t* d_viol, uint8_t* d_tx_flags) {
  if (!n_tx) return XHE_OK;
  XheTimed t(ctx, "k_layout", 0);
  k_layout<<<(n_tx + 63) / 64, 64, 0, ctx->stream>>>(d_blobs, d_off, d_plan, n_tx, n_points, d_enc, d_sig_idx, n_eq, d_eq_sc, d_val_sc, d_rp_sc, d_range_idx, d_rp_pt_off, d_sig_s, d_sig_e, d_sig_pk, d_viol, d_tx_flags);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_launch_tx_flags(xhe_ctx* ctx, const uint32_t* d_plan, uint32_t n_tx, uint32_t n_a_end, const uint8_t* d_pt_ok, const uint8_t* d_sig_ok, uint8_t* d_tx_flags) {
  if (!n_tx) return XHE_OK;
  k_tx_flags<<<(n_tx + 127) / 128, 128, 0, ctx->stream>>>(d_plan, n_tx, n_a_end, d_pt_ok, d_sig_ok, d_tx_flags);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}
int32_t xhe_launch_any_zero(xhe_ctx* ctx, const uint8_t* d_flags, uint32_t n, uint32_t bit, uint32_t* d_viol) {
  if (!n) return XHE_OK;
  k_any_zero<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_flags, n, bit, d_viol);
  XHE_LAUNCHED(ctx); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK;
}

// CUDA loads kernels lazily (CUDA_MODULE_LOADING=LAZY is the default since 12.2), and loading one may need every running kernel
// to finish first: with the polling chain kernel of msm.cu in flight, the FIRST launch of any other kernel would wait for a
// kernel that is waiting for it.  Every kernel of this file is therefore loaded when the first context is created.
size_t xhe_preload_fs() {      // returns the largest per-thread local-memory frame among them
  const void* ks[] = {(const void*)k_fiat_shamir, (const void*)k_sig_hash_prefix, (const void*)k_sig_hash_final, (const void*)k_layout, (const void*)k_tx_flags, (const void*)k_any_zero};
  cudaFuncAttributes a; size_t mx = 0;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) == cudaSuccess && a.localSizeBytes > mx) mx = a.localSizeBytes;
  return mx;
}
