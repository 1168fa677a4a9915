// msm.cu -- K6/K7: variable-time multiscalar multiplication (signed-digit Pippenger) for sm_100a.
//
// Replaces RistrettoPoint::vartime_multiscalar_mul + is_identity (reference src/proofs.rs:49-67; dalek's
// Straus/Pippenger behind it) and the MSM inside bulletproofs' verify_batch (src/tx/verify.rs:504-514).
// Not a port of dalek's w <= 8 column loop: the whole scalar is recoded into W = ceil(254/c) signed c-bit digits
// (c chosen per n, up to 16+), every (window, bucket) pair is an independent work item, and the pipeline is
//
//   k_msm_count       thread/point : recode scalar -> W signed digits, histogram (window,bucket) sizes      [L2 atomics]
//   scan              3 kernels    : exclusive prefix sum over the W*2^(c-1) bucket sizes
//   k_msm_scatter     thread/point : counting-sort point indices (sign in bit 31) into bucket order          [L2 atomics, 8 in flight]
//   k_msm_tile_runs   thread/tile  : the sorted list is cut into fixed tiles of 32 entries; count the bucket runs per tile
//   k_msm_accum_tiles thread/tile  : gather 96-byte affine-Niels points (6 x LDG.128, next point prefetched),
//                                    7 M mixed additions into a register-resident extended accumulator      [HOT: IMAD pipe]
//                                    -- also records where each bucket's partial sums start / how many there are
//   k_msm_fold_heavy  block/bucket : buckets with many partials (under-filled top window, skewed scalars)
//   k_msm_bucket_seg  CTA/256 buckets: thread/bucket merge of the partials, then quad/4 buckets running sum -> (run, wsum) node
//   k_msm_nodes32     CTA/32 nodes : shared-memory suffix scan + tree with quad-cooperative point arithmetic (quad.cuh),
//                                    repeated until one node per window
//   k_msm_horner_g    1 warp       : Horner over the windows of one window GROUP, chained through a 128-byte accumulator
//
// Windows are sorted most-significant first and processed in G groups: group g's accumulation runs on the caller's stream
// while the latency-bound reduction + Horner segment of group g-1 (its c * windows dependent doublings) runs on a
// high-priority side stream, so only the last group's tail is exposed.
//
// Algorithmic work: n*W mixed adds of 7 M = 504 limb products each dominate.
#include <stdlib.h>
#include <stddef.h>
#include "xhe_internal.cuh"
#include "quad.cuh"
#include "oct.cuh"
#include <algorithm>
#include <vector>
#include <string.h>
using namespace xhe;

namespace {

#define MSM_TILE 32   // entries per accumulation work item
#define MSM_SEG 4     // buckets per level-1 running-sum segment (one quad)
#define MSM_SEG_LOG 2
#define MSM_MAX_GROUPS 8
#define NODES_SEQ_R 8    // children per parent in the work-efficient fold levels (k_msm_nodes_seq)
#define XHE_ACCUM_SMEM_DEFAULT 0

struct MsmPlan {
  int c, W;             // window bits, windows
  uint32_t B;           // buckets per window = 2^(c-1)
  size_t total_buckets; // W * B
  int G, Wg;            // window groups (most significant first) and windows per group (the last group may be shorter)
  int seg_log;          // log2 of the buckets one quad folds in the first reduction level (2 or 4)
  // workspace offsets (bytes)
  size_t n_tiles, max_runs;
  size_t off_counts, off_offsets, off_cursor, off_blocksums, off_list, off_tileg0, off_runs, off_runoff, off_part, off_pstart, off_pcount, off_heavy, off_nodes_a, off_nodes_b, off_hnodes, off_hacc, off_ready, off_flag, total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

MsmPlan make_plan(size_t n) {
  MsmPlan p;
  // cost model in units of one mixed add: W * (n + K * 2^(c-1)); c in [4, 18].  K = what one bucket costs in the reduction
  // stages relative to one accumulated entry.  Measured on B200 (tools/msm_bench.py sweep, profiles/r01_msm_window_sweep.md):
  // below 2^19 points the latency of the per-window stages dominates and fewer, larger windows win (K = 3 picks c = 13 at
  // 2^16, 15 at 2^18); above, the reduction's ~1.35 ns per bucket against ~0.1 ns per entry makes K = 12 the better fit
  // (c = 15 at 2^20, 16 at 2^22).  Buckets must stay below ~100 entries on average: longer ones take the heavy-bucket path.
  static const double K_env = getenv("XHE_MSM_BUCKET_COST") ? atof(getenv("XHE_MSM_BUCKET_COST")) : 0.0;
  const double K = K_env > 0 ? K_env : (n >= ((size_t)1 << 19) ? 12.0 : 3.0);
  double best = 1e300; int bc = 8;
  for (int c = 4; c <= 18; c++) {
    int W = (254 + c - 1) / c;
    double cost = (double)W * ((double)n + K * (double)(1u << (c - 1)));
    if (cost < best) { best = cost; bc = c; }
  }
  if (n <= 4096 && bc < 8) bc = 8;   // tiny inputs are pure latency: fewer windows shorten the per-window stages and the Horner chain
  p.c = bc; p.W = (254 + bc - 1) / bc; p.B = 1u << (bc - 1); p.total_buckets = (size_t)p.W * p.B;
  // window groups: the reduction + Horner segment of a group overlaps the accumulation of the next one.  Small inputs are
  // pure latency (one launch wave per group at best): no split.
  // Measured on B200 (profiles/r02_msm_sweep.md): splitting never paid -- a group's launch of the accumulation is a partial
  // wave at 150 k points (four launches of 183 blocks took 2.2x one launch of 732), and at 2^20..2^22 the reduction kernels
  // that run beside the next group's accumulation slow it by as much as they hide (2.43 ms at G = 1, 2.50-2.55 at G = 3..4).
  // One group is the default; XHE_MSM_GROUPS keeps the experiment reproducible.
  static const int g_env = getenv("XHE_MSM_GROUPS") ? atoi(getenv("XHE_MSM_GROUPS")) : 1;
  int G = std::max(1, std::min(g_env, MSM_MAX_GROUPS));
  if (n * (size_t)p.W < ((size_t)1 << 17)) G = 1;
  p.Wg = (p.W + G - 1) / G; p.G = (p.W + p.Wg - 1) / p.Wg;
  // first reduction level: a quad folds 4 buckets (5 dependent point operations: small MSMs are latency-bound there) or 16 (30
  // operations, but a quarter of the nodes reach the scan kernels, which spend 10 operations per node: large MSMs are
  // throughput-bound there -- 2^20 points: 0.39 ms -> see profiles/r02_msm_sweep.md)
  static const int seg_env = getenv("XHE_MSM_SEG_LOG") ? atoi(getenv("XHE_MSM_SEG_LOG")) : 0;
  p.seg_log = seg_env == 2 || seg_env == 4 ? seg_env : 2;      // with the work-efficient fold levels (k_msm_nodes_seq) the short first level wins at every size: see DESIGN.md 4.4
  if ((1u << p.seg_log) > p.B) p.seg_log = 2;
  size_t o = 0;
  p.off_counts = o; o = align_up(o + 4 * (p.total_buckets + 1), 256);
  p.off_offsets = o; o = align_up(o + 4 * (p.total_buckets + 1), 256);
  p.off_cursor = o; o = align_up(o + 4 * (p.total_buckets + 1), 256);
  p.off_blocksums = o; o = align_up(o + 4 * 4096, 256);
  size_t nw = n * (size_t)p.W;
  p.n_tiles = (nw + MSM_TILE - 1) / MSM_TILE;
  p.max_runs = p.n_tiles + std::min(p.total_buckets, nw) + 1;
  p.off_list = o; o = align_up(o + 4 * nw + 4, 256);
  p.off_tileg0 = o; o = align_up(o + 4 * (p.n_tiles + 1), 256);       // tile_g0: bucket of each tile's first entry
  p.off_runs = o; o = align_up(o + 4 * (p.n_tiles + 1), 256);
  p.off_runoff = o; o = align_up(o + 4 * (p.n_tiles + 1), 256);
  p.off_part = o; o = align_up(o + 128 * p.max_runs, 256);
  p.off_pstart = o; o = align_up(o + 4 * p.total_buckets, 256);
  p.off_pcount = o; o = align_up(o + 4 * p.total_buckets, 256);
  p.off_heavy = o; o = align_up(o + 4 * (p.total_buckets + 8 * (MSM_MAX_GROUPS + 1)), 256);      // per group: [count, bucket ids...]
  size_t nodes1 = (p.total_buckets + MSM_SEG - 1) / MSM_SEG + 64;
  p.off_nodes_a = o; o = align_up(o + 256 * nodes1, 256);
  p.off_nodes_b = o; o = align_up(o + 256 * ((nodes1 + NODES_SEQ_R - 1) / NODES_SEQ_R + (size_t)p.W + 64), 256);      // first fold level: radix 8
  p.off_hnodes = o; o = align_up(o + 256 * ((size_t)p.W + 1), 256);      // one final node per window: input of the Horner stream
  p.off_hacc = o; o = align_up(o + 128 * (MSM_MAX_GROUPS + 1), 256);
  p.off_ready = o; o = align_up(o + 64, 256);          // ready[g]: final nodes of group g written; ready[8]: chain status
  p.off_flag = o; o = align_up(o + 64, 256);
  p.total = o;
  return p;
}

// ---- digit recoding ----------------------------------------------------------------------------------------------
// signed radix-2^c digits d_w in [-2^(c-1), 2^(c-1)], sum d_w 2^(c w) = s, for s < 2^253 and c*W >= 254 (the top digit
// absorbs the final carry without overflow).  The sort key of digit d of window w is (W-1-w) * B + |d| - 1: windows are
// laid out MOST SIGNIFICANT FIRST, so the reduction / Horner pipeline can start on the top windows while the lower ones
// are still being accumulated.
__device__ __forceinline__ uint32_t limb_of(const uint32_t s[8], int limb) {      // s[limb] without dynamic register indexing
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) r = (limb == k) ? s[k] : r;
  return r;
}
__device__ __forceinline__ int32_t next_digit(const uint32_t s[8], int w, int c, uint32_t& carry) {
  const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
  int bit = w * c, limb = bit >> 5, sh = bit & 31;
  uint32_t v = 0;
  if (limb < 8) {
    v = limb_of(s, limb) >> sh;
    if (sh + c > 32 && limb + 1 < 8) v |= limb_of(s, limb + 1) << (32 - sh);
  }
  v = (v & mask) + carry;
  carry = v > half ? 1u : 0u;            // v in [0, 2^c]; digits above half become negative with a carry
  return carry ? (int32_t)v - (int32_t)(1u << c) : (int32_t)v;
}
__device__ __forceinline__ void ld_scalar(uint32_t s[8], const uint32_t* __restrict__ scalars, size_t i) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(scalars + 8 * i)), b = __ldg(reinterpret_cast<const uint4*>(scalars + 8 * i) + 1);
  s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

__global__ void __launch_bounds__(256) k_msm_count(const uint32_t* __restrict__ scalars, size_t n, int c, int W, uint32_t B, uint32_t* __restrict__ counts, uint32_t* __restrict__ bad_flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8]; ld_scalar(s, scalars, i);
  if (sc_geq_l(s)) { atomicOr(bad_flag, 1u); return; }   // non-canonical scalar: reported as a bad argument, contributes nothing
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) {
    int32_t d = next_digit(s, w, c, carry);
    if (d != 0) atomicAdd(&counts[(size_t)(W - 1 - w) * B + (uint32_t)((d < 0 ? -d : d) - 1)], 1u);
  }
}

// counting-sort scatter.  The position of an entry is a RETURNING atomic on its bucket's cursor; eight of them are issued
// before the first dependent store, so a thread has eight L2 round trips in flight instead of one.
#define SCAT_CH 8
__global__ void __launch_bounds__(256) k_msm_scatter(const uint32_t* __restrict__ scalars, size_t n, int c, int W, uint32_t B, uint32_t* __restrict__ cursor, uint32_t* __restrict__ list) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8]; ld_scalar(s, scalars, i);
  if (sc_geq_l(s)) return;
  uint32_t carry = 0;
  for (int w0 = 0; w0 < W; w0 += SCAT_CH) {
    uint32_t key[SCAT_CH], val[SCAT_CH], pos[SCAT_CH];
#pragma unroll
    for (int j = 0; j < SCAT_CH; j++) {
      key[j] = 0xffffffffu; val[j] = 0;
      if (w0 + j < W) {
        int32_t d = next_digit(s, w0 + j, c, carry);
        if (d != 0) { key[j] = (uint32_t)(W - 1 - (w0 + j)) * B + (uint32_t)((d < 0 ? -d : d) - 1); val[j] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u); }
      }
    }
#pragma unroll
    for (int j = 0; j < SCAT_CH; j++) pos[j] = key[j] != 0xffffffffu ? atomicAdd(&cursor[key[j]], 1u) : 0u;
#pragma unroll
    for (int j = 0; j < SCAT_CH; j++) if (key[j] != 0xffffffffu) list[pos[j]] = val[j];     // the bucket of a position follows from the offsets: no second array
  }
}

// ---- exclusive scan over m uint32 (m up to 4096 * 2048) --------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8   // per thread -> 2048 per block
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(const uint32_t* __restrict__ in, size_t m, uint32_t* __restrict__ out, uint32_t* __restrict__ blocksums) {
  __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
  size_t base = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < m ? in[base + k] : 0u; sum += v[k]; }
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += warp_tot[w];
  uint32_t run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < m) out[base + k] = run; run += v[k]; }
  if (threadIdx.x == SCAN_THREADS - 1) blocksums[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t* __restrict__ blocksums, int nb, uint32_t* __restrict__ grand_total) {
  // single block, nb <= 4096: serial-in-thread chunks + one warp-level pass
  __shared__ uint32_t part[1024];
  int per = (nb + 1023) / 1024;
  int lo = threadIdx.x * per, hi = min(nb, lo + per);
  uint32_t s = 0;
  for (int i = lo; i < hi; i++) s += blocksums[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < 1024; i++) { uint32_t t = part[i]; part[i] = run; run += t; } *grand_total = run; }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int i = lo; i < hi; i++) { uint32_t t = blocksums[i]; blocksums[i] = run; run += t; }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, size_t m, const uint32_t* __restrict__ blocksums, uint32_t* __restrict__ copy) {
  size_t base = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  uint32_t add = blocksums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < m) { uint32_t v = out[base + k] + add; out[base + k] = v; if (copy) copy[base + k] = v; }
}

// ---- balanced accumulation: fixed-size tiles of the bucket-sorted entry list ----------------------------------------
// Bucket g owns the positions [offsets[g], offsets[g+1]) of the sorted list (offsets[m] = number of entries).
// bucket_of: the bucket that contains position pos, searched in (lo_hint, m); requires offsets[lo_hint] <= pos.
__device__ __forceinline__ uint32_t bucket_of(const uint32_t* __restrict__ offsets, uint32_t m, uint32_t pos, uint32_t lo_hint) {
  uint32_t lo = lo_hint + 1, hi = m;                 // smallest idx in [lo, hi] with offsets[idx] > pos
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(offsets + mid) > pos) hi = mid; else lo = mid + 1; }
  return lo - 1;
}
// the bucket that starts at position pos, given that bucket g ends there: usually g + 1; a binary search skips empty buckets
__device__ __forceinline__ uint32_t next_bucket(const uint32_t* __restrict__ offsets, uint32_t m, uint32_t pos, uint32_t g) {
  return __ldg(offsets + g + 2) > pos ? g + 1 : bucket_of(offsets, m, pos, g + 1);
}
// runs[t] = number of maximal same-bucket runs inside tile t (a run is one partial sum); tile_g0[t] = bucket of its first entry
__global__ void __launch_bounds__(256) k_msm_tile_runs(const uint32_t* __restrict__ offsets, uint32_t m, size_t n_tiles, uint32_t* __restrict__ runs, uint32_t* __restrict__ tile_g0) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  uint32_t N = __ldg(offsets + m);
  size_t start = t * MSM_TILE;
  uint32_t cnt = 0;
  if (t < n_tiles && start < N) {
    uint32_t end = (uint32_t)min((size_t)N, start + MSM_TILE);
    uint32_t g = bucket_of(offsets, m, (uint32_t)start, 0);
    tile_g0[t] = g; cnt = 1;
    uint32_t nb = __ldg(offsets + g + 1);
    while (nb < end) { g = next_bucket(offsets, m, nb, g); nb = __ldg(offsets + g + 1); cnt++; }
  }
  runs[t] = cnt;
}

// HOT: each thread walks MSM_TILE consecutive entries, gathers the 96-byte affine-Niels points (6 x LDG.128, the next
// point prefetched under the current addition) and accumulates 7 M mixed additions in registers; a partial sum is
// flushed whenever the bucket id changes.  Every lane does the same number of additions: no divergence on bucket size.
// One launch per window group: a thread works only if its tile STARTS in a bucket of [klo, khi).  The partial sums of a
// bucket occupy consecutive slots in bucket order; the run that contains the bucket's first entry records the first slot
// (pstart), every flush counts (pcount), and the flush that makes a bucket "heavy" appends it to its group's list.
#define HEAVY_PARTIALS 12
struct AccumOut { uint32_t *part, *pstart, *pcount, *heavy; uint32_t keys_per_group; };
__device__ __forceinline__ void flush_run(const AccumOut& o, uint32_t slot, uint32_t g, const ge& acc) {
  st_ge(o.part + 32 * (size_t)slot, acc);
  if (atomicAdd(&o.pcount[g], 1u) == HEAVY_PARTIALS) {
    uint32_t grp = g / o.keys_per_group; uint32_t* h = o.heavy + (size_t)grp * o.keys_per_group + 8 * grp;
    h[1 + atomicAdd(&h[0], 1u)] = g;
  }
}
// SMEM_ACC is the A/B for north_star's "shared-memory bucket accumulation": the same kernel with the running bucket sum kept in
// the CTA's shared memory (128 B per thread, read and written around every addition) instead of registers.  XHE_MSM_SMEM_ACC=1
// selects it; the measured difference is in DESIGN.md 4.4.
__device__ __forceinline__ void lds_ge(ge& g, const uint32_t* p) {      // volatile 128-bit shared-memory accesses: the compiler may not keep the sum in registers
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  uint32_t* w = &g.X.v[0];
#pragma unroll
  for (int i = 0; i < 8; i++) asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4 * i]), "=r"(w[4 * i + 1]), "=r"(w[4 * i + 2]), "=r"(w[4 * i + 3]) : "r"(a + 16 * i));
}
__device__ __forceinline__ void sts_ge(uint32_t* p, const ge& g) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  const uint32_t* w = &g.X.v[0];
#pragma unroll
  for (int i = 0; i < 8; i++) asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a + 16 * i), "r"(w[4 * i]), "r"(w[4 * i + 1]), "r"(w[4 * i + 2]), "r"(w[4 * i + 3]) : "memory");
}
static_assert(sizeof(ge) == 128 && offsetof(ge, Y) == 32 && offsetof(ge, Z) == 64 && offsetof(ge, T) == 96, "ge is four contiguous 8-word field elements");
template <int MINB, bool SMEM_ACC>
__global__ void __launch_bounds__(128, MINB) k_msm_accum_tiles(const uint32_t* __restrict__ niels, const uint32_t* __restrict__ list, const uint32_t* __restrict__ offsets, uint32_t m,
                                                              const uint32_t* __restrict__ tile_g0, const uint32_t* __restrict__ run_off, size_t n_tiles,
                                                              uint32_t klo, uint32_t khi, AccumOut out) {
  __shared__ __align__(16) uint32_t sacc[SMEM_ACC ? 128 * 32 : 4];
  uint32_t* my = sacc + (SMEM_ACC ? 32 * threadIdx.x : 0);
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  uint32_t N = __ldg(offsets + m);
  size_t start = t * MSM_TILE;
  if (start >= N) return;
  uint32_t g = __ldg(tile_g0 + t);
  if (g < klo || g >= khi) return;
  uint32_t cnt = (uint32_t)(min((size_t)N, start + MSM_TILE) - start);
  uint32_t slot = run_off[t];
  uint32_t nb = __ldg(offsets + g + 1);          // nb: first position past the current bucket
  if (__ldg(offsets + g) == (uint32_t)start) out.pstart[g] = slot;
  uint32_t e = __ldg(list + start);
  ge_niels q; ld_niels(q, niels + 24 * (size_t)(e & 0x7fffffffu));
  ge acc = ge_from_niels(niels_cneg(q, (e >> 31) != 0));
  if (SMEM_ACC) sts_ge(my, acc);
  if (cnt > 1) { e = __ldg(list + start + 1); ld_niels(q, niels + 24 * (size_t)(e & 0x7fffffffu)); }
  for (uint32_t j = 1; j < cnt; j++) {
    const uint32_t pos = (uint32_t)start + j;
    ge_niels cur = niels_cneg(q, (e >> 31) != 0);
    if (j + 1 < cnt) { e = __ldg(list + start + j + 1); ld_niels(q, niels + 24 * (size_t)(e & 0x7fffffffu)); }
    if (SMEM_ACC) lds_ge(acc, my);
    if (pos == nb) {   // bucket boundary: flush the finished run, restart from this point
      flush_run(out, slot, g, acc); slot++;
      g = next_bucket(offsets, m, pos, g); nb = __ldg(offsets + g + 1);
      out.pstart[g] = slot;                       // pos is this bucket's first entry
      acc = ge_from_niels(cur);
    } else {
      acc = ge_madd(acc, cur);
    }
    if (SMEM_ACC) sts_ge(my, acc);
  }
  if (SMEM_ACC) lds_ge(acc, my);
  flush_run(out, slot, g, acc);
}

__global__ void k_msm_zero_heads(uint32_t* __restrict__ heavy, uint32_t keys_per_group, int G, uint32_t* __restrict__ ready) {
  if ((int)threadIdx.x < G) heavy[(size_t)threadIdx.x * keys_per_group + 8 * threadIdx.x] = 0;
  if (threadIdx.x < 16) ready[threadIdx.x] = 0;
}

// buckets whose partial list is long (under-filled top window, skewed scalars) are folded first, one WARP per bucket: the
// lanes stride over the partial sums, then a shuffle tree.  (Round 1 used a block per bucket with a 7-round shared-memory
// tree: at 2^22 points the top window alone has 8,192 buckets of ~17 partials, and that fold took 0.19 ms at 2^20.)
#define FOLD_THREADS 128
__device__ __forceinline__ ge shfl_down_ge(const ge& p, int d) {
  ge r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], d); r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], d);
    r.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], d); r.T.v[i] = __shfl_down_sync(0xffffffffu, p.T.v[i], d);
  }
  return r;
}
__global__ void __launch_bounds__(FOLD_THREADS) k_msm_fold_heavy(uint32_t* __restrict__ part, const uint32_t* __restrict__ pstart, uint32_t* __restrict__ pcount, const uint32_t* __restrict__ heavy) {
  const uint32_t nh = heavy[0], lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t h = warp; h < nh; h += n_warps) {
    const uint32_t g = heavy[1 + h], ps = pstart[g], pc = pcount[g];
    ge acc = ge_identity(), s;
    for (uint32_t j = lane; j < pc; j += 32) { ld_ge(s, part + 32 * (size_t)(ps + j)); acc = ge_add(acc, s); }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) acc = ge_add(acc, shfl_down_ge(acc, d));
    if (lane == 0) { st_ge(part + 32 * (size_t)ps, acc); pcount[g] = 1; }
  }
}

// ---- bucket reduction -----------------------------------------------------------------------------------------------
// node = (run, wsum): run = sum of the bucket sums in its range, wsum = sum (b - lo) * S_b (weights relative to the range).
// These stages have few work items and long dependent chains, so every point operation is shared by the four lanes of a
// quad (quad.cuh); each lane of a quad holds the same point.
__device__ __forceinline__ void st_ge_quad(uint32_t* p, const ge& g) {       // lane ql of the quad stores coordinate ql
  const int ql = threadIdx.x & 3;
  st_fe(p + 8 * ql, quad_pick(g.X, g.Y, g.Z, g.T, ql));
}
// level 1: one CTA per 256 consecutive bucket keys.  Phase 1: thread/bucket adds up the bucket's partial sums (at most
// HEAVY_PARTIALS after the fold).  Phase 2: quad/L buckets running sum -> node of width L (L = 4: all 64 quads; L = 16: 16 quads).
template <int L>
__global__ void __launch_bounds__(256) k_msm_bucket_seg(const uint32_t* __restrict__ part, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pcount,
                                                        uint32_t klo, uint32_t khi, uint32_t* __restrict__ nodes) {
  __shared__ __align__(16) uint32_t sm[256 * 32];
  const uint32_t k = klo + blockIdx.x * 256 + threadIdx.x;
  ge v = ge_identity();
  if (k < khi) {
    const uint32_t pc = pcount[k];
    if (pc) {
      const uint32_t ps = pstart[k];
      ld_ge(v, part + 32 * (size_t)ps);
      for (uint32_t j = 1; j < pc; j++) { ge s; ld_ge(s, part + 32 * (size_t)(ps + j)); v = ge_add(v, s); }
    }
  }
  st_ge(sm + 32 * threadIdx.x, v);
  __syncthreads();
  const uint32_t q = threadIdx.x >> 2;
  if (q >= 256 / L) return;                      // (whole warps: 256 / L quads = 64 or 16, a multiple of 8)
  ge run, wsum, s;
  ld_ge(run, sm + 32 * (L * q + L - 1)); wsum = run;
#pragma unroll 1
  for (int r = L - 2; r >= 1; r--) { ld_ge(s, sm + 32 * (L * q + r)); run = quad_add(run, s); wsum = quad_add(wsum, run); }
  ld_ge(s, sm + 32 * (L * q)); run = quad_add(run, s);
  const uint32_t kq = klo + blockIdx.x * 256 + L * q;
  if (kq < khi) { uint32_t* o = nodes + 64 * (size_t)((kq - klo) / L); st_ge_quad(o, run); st_ge_quad(o + 32, wsum); }
}

// one CTA (32 quads) folds 32 consecutive child nodes (each of width 2^child_log buckets) of ONE window into a parent:
//   run' = sum run_c ; wsum' = sum wsum_c + 2^child_log * sum_c c * run_c , and sum_c c * run_c = sum_{c >= 1} suffix_c.
// Suffix sums by a Hillis-Steele scan through shared memory (5 rounds), then the 31 suffixes and the 32 wsums are summed by
// two trees that run side by side on quads 0-15 / 16-31 (5 rounds).
__global__ void __launch_bounds__(128) k_msm_nodes32(const uint32_t* __restrict__ in, uint32_t children_per_window, uint32_t parents_per_window, int child_log, uint32_t* __restrict__ out,
                                                     uint32_t* __restrict__ ready /* NULL, or the counter of finished final nodes the chain kernel polls */) {
  __shared__ __align__(16) uint32_t sA[32 * 32], sW[32 * 32];
  const uint32_t w = blockIdx.x / parents_per_window, pidx = blockIdx.x % parents_per_window;
  const uint32_t q = threadIdx.x >> 2, child = pidx * 32 + q;
  ge run, wsum;
  if (child < children_per_window) { const uint32_t* p = in + 64 * ((size_t)w * children_per_window + child); ld_ge(run, p); ld_ge(wsum, p + 32); }
  else { run = ge_identity(); wsum = ge_identity(); }
  ge suf = run;
#pragma unroll 1
  for (uint32_t d = 1; d < 32; d <<= 1) {
    st_ge_quad(sA + 32 * q, suf);
    __syncthreads();
    ge t = ge_identity();
    if (q + d < 32) ld_ge(t, sA + 32 * (q + d));
    __syncthreads();
    suf = quad_add(suf, t);
  }
  { ge A = suf; if (q == 0) A = ge_identity(); st_ge_quad(sA + 32 * q, A); st_ge_quad(sW + 32 * q, wsum); }
  __syncthreads();
  uint32_t* arr = q < 16 ? sA : sW; const uint32_t i = q & 15;
#pragma unroll 1
  for (uint32_t st = 16; st >= 1; st >>= 1) {
    ge a = ge_identity(), b = ge_identity();
    if (i < st) { ld_ge(a, arr + 32 * i); ld_ge(b, arr + 32 * (i + st)); }
    ge r = quad_add(a, b);
    __syncthreads();
    if (i < st) st_ge_quad(arr + 32 * i, r);
    __syncthreads();
  }
  if (threadIdx.x < 32) {     // warp 0 (its eight quads compute the same thing; quad 0 stores)
    ge At, Wt; ld_ge(At, sA); ld_ge(Wt, sW);
    for (int k = 0; k < child_log; k++) At = quad_double(At);
    Wt = quad_add(Wt, At);
    if (q == 0) { uint32_t* o = out + 64 * ((size_t)w * parents_per_window + pidx); st_ge_quad(o, suf); st_ge_quad(o + 32, Wt); if (ready) __threadfence(); }
    __syncwarp();
    if (ready && threadIdx.x == 0) atomicAdd(ready, 1u);      // release: the four lanes' stores are fenced before the count moves
  }
}

// The same fold for R = 8 children per parent, WORK-efficient: a pair of quads per parent walks the children from the last to
// the first -- quad 0 keeps run' (a running sum) and sum_{c >= 1} suffix_c (the running sum added up), quad 1 sums the
// wsum_c -- 3 additions per child instead of the ~6 of the parallel scan above, at 2 R dependent additions.  Used for the
// wide middle levels (thousands of nodes per window), where the scan kernel was throughput-bound (2^20 points: 174 us for the
// first level); the last level (<= 32 nodes per window) keeps the parallel form.
__global__ void __launch_bounds__(128) k_msm_nodes_seq(const uint32_t* __restrict__ in, uint32_t children_per_window, uint32_t parents_per_window, uint32_t n_parents, int child_log, uint32_t* __restrict__ out) {
  const uint32_t pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, role = (threadIdx.x >> 2) & 1u;      // role 0: run / suffix sums, role 1: wsum
  const bool live = pair < n_parents;                                                                    // (whole warps stay in the shuffles)
  const uint32_t w = live ? pair / parents_per_window : 0u, pidx = live ? pair % parents_per_window : 0u;
  ge acc1 = ge_identity(), acc2 = ge_identity();
#pragma unroll 1
  for (int c = NODES_SEQ_R - 1; c >= 0; c--) {
    const uint32_t child = pidx * NODES_SEQ_R + (uint32_t)c;
    ge x = ge_identity();
    if (live && child < children_per_window) ld_ge(x, in + 64 * ((size_t)w * children_per_window + child) + 32 * role);
    acc1 = quad_add(acc1, x);
    if (c >= 1) acc2 = quad_add(acc2, acc1);
  }
#pragma unroll 1
  for (int k = 0; k < child_log; k++) acc2 = quad_double(acc2);
  // quad 0 takes the other quad's sum of the wsums
  ge W;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    W.X.v[i] = __shfl_down_sync(0xffffffffu, acc1.X.v[i], 4); W.Y.v[i] = __shfl_down_sync(0xffffffffu, acc1.Y.v[i], 4);
    W.Z.v[i] = __shfl_down_sync(0xffffffffu, acc1.Z.v[i], 4); W.T.v[i] = __shfl_down_sync(0xffffffffu, acc1.T.v[i], 4);
  }
  const ge Wt = quad_add(acc2, W);
  if (live && role == 0u) { uint32_t* o = out + 64 * ((size_t)w * parents_per_window + pidx); st_ge_quad(o, acc1); st_ge_quad(o + 32, Wt); }
}

// Horner segment over the n_w windows of one group (most significant first): acc = 2^c * acc + S_w, S_w = wsum + run
// (bucket b carries multiplier b+1).  acc_in = the accumulator the previous group left (NULL for the first group, whose
// first window needs no doublings).  c * n_w sequential doublings: a pure latency chain, so one quad shares every point
// operation (the eight quads of the warp first form the S_w in parallel).  The last group also encodes the result.
__global__ void __launch_bounds__(32) k_msm_horner_g(const uint32_t* __restrict__ nodes, int n_w, int c, const uint32_t* __restrict__ acc_in, uint32_t* __restrict__ acc_out, int last,
                                                    uint8_t* __restrict__ out_enc, uint32_t* __restrict__ is_identity, uint32_t* __restrict__ out_ext) {
  __shared__ __align__(16) uint32_t sS[32 * 32];
  const int q = threadIdx.x >> 2;
  for (int base = 0; base < n_w; base += 8) {
    const int j = base + q;
    ge run = ge_identity(), wsum = ge_identity();
    if (j < n_w) { ld_ge(run, nodes + 64 * (size_t)j); ld_ge(wsum, nodes + 64 * (size_t)j + 32); }
    ge S = quad_add(run, wsum);
    if (j < n_w) st_ge_quad(sS + 32 * j, S);
  }
  __syncwarp();
  ge acc = ge_identity();
  if (acc_in) ld_ge(acc, acc_in);
  for (int w = 0; w < n_w; w++) {
    if (acc_in || w) for (int k = 0; k < c; k++) acc = quad_double(acc);
    ge S; ld_ge(S, sS + 32 * w);
    acc = quad_add(acc, S);
  }
  if (threadIdx.x != 0) return;
  if (acc_out) st_ge(acc_out, acc);
  if (!last) return;
  if (out_ext) {   // canonical coordinates so any consumer (other ranks, the CPU oracle) can read them
    st_fe(out_ext, fe_freeze(acc.X)); st_fe(out_ext + 8, fe_freeze(acc.Y)); st_fe(out_ext + 16, fe_freeze(acc.Z)); st_fe(out_ext + 24, fe_freeze(acc.T));
  }
  if (out_enc) encode_result_words(out_enc, acc);
  if (is_identity) *is_identity = ge_ristretto_is_identity(acc) ? 1u : 0u;
}
// The same recombination with warp-cooperative arithmetic (oct.cuh): warp w first forms S_w = run_w + wsum_w, then warp 0 runs
// the chain acc = 2^c acc + S_w with the accumulator spread over its 32 lanes (one limb of one coordinate per lane).  The
// default for a single window group; the quad kernel above remains for window groups / the polling-chain experiment.
#define HORNER_OCT_THREADS 256
__global__ void __launch_bounds__(HORNER_OCT_THREADS) k_msm_horner_oct(const uint32_t* __restrict__ nodes, int n_w, int c, uint8_t* __restrict__ out_enc, uint32_t* __restrict__ is_identity, uint32_t* __restrict__ out_ext) {
  __shared__ uint32_t sS[32 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int w = warp; w < n_w; w += HORNER_OCT_THREADS / 32) sS[32 * w + lane] = oct_add_pt(__ldg(nodes + 64 * (size_t)w + lane), __ldg(nodes + 64 * (size_t)w + 32 + lane));
  __syncthreads();
  if (warp != 0) return;
  uint32_t acc = sS[lane];                                  // window 0 = most significant: no doublings before it
  for (int w = 1; w < n_w; w++) {
#pragma unroll 1
    for (int k = 0; k < c; k++) acc = oct_double(acc);
    acc = oct_add_pt(acc, sS[32 * w + lane]);
  }
  const ge r = oct_to_ge(acc);
  if (lane != 0) return;
  if (out_ext) { st_fe(out_ext, fe_freeze(r.X)); st_fe(out_ext + 8, fe_freeze(r.Y)); st_fe(out_ext + 16, fe_freeze(r.Z)); st_fe(out_ext + 24, fe_freeze(r.T)); }
  if (out_enc) encode_result_words(out_enc, r);
  if (is_identity) *is_identity = ge_ristretto_is_identity(r) ? 1u : 0u;
}
// self-test of oct.cuh (xhe_selftest_oct): op 0 mul, 1 add, 2 sub on field elements (8 lanes each); 3 doubling, 4 addition on points (a warp each)
__global__ void __launch_bounds__(128) k_selftest_oct(int op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, size_t n, uint32_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (op < 3) {
    const size_t i = t >> 3; const bool live = i < n;           // every lane of the warp takes part in the shuffles
    const uint32_t x = live ? a[t] : 0u, y = live ? b[t] : 0u;
    const uint32_t r = op == 0 ? oct_mul(x, y) : (op == 1 ? oct_add(x, y) : oct_sub(x, y));
    if (live) out[t] = r;
  } else {
    const size_t i = t >> 5; const bool live = i < n;
    const uint32_t x = live ? a[t] : oct_identity(), y = live ? b[t] : oct_identity();
    const uint32_t r = op == 3 ? oct_double(x) : oct_add_pt(x, y);
    if (live) out[t] = r;
  }
}

// The same chain as ONE kernel that is launched BEFORE its inputs exist and polls for them (experiment, XHE_MSM_CHAIN=1):
// a latency-bound chain that shares a sub-partition with the warps of a throughput-bound kernel gets 1/(N+1) of the
// multiplier pipe and stretches N-fold (measured: the per-group Horner kernels above gained nothing beside the
// accumulation).  This CTA is sized to own its SM (512 threads x 128 registers = the whole register file; 14 warps park
// at the final barrier), so the one or two chains it serves (warp 0 / warp 1 sit on different sub-partitions) run at their
// isolated speed.  The producers (last level of k_msm_nodes32) count finished nodes in ready[g]; the polling is bounded
// (2 s): on a timeout the status word is raised and the caller reports XHE_E_CUDA instead of hanging.
struct ChainJob {
  const uint32_t* hnodes; uint32_t* ready; int G, Wg, W, c;
  uint8_t* out_enc; uint32_t* is_identity; uint32_t* out_ext; uint32_t* status;
};
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void chain_run(const ChainJob& J, uint32_t* sS, int wait) {
  const int lane = threadIdx.x & 31, q = lane >> 2;
  ge acc = ge_identity();
  bool ok = true; uint32_t diag = 0;
  for (int g = 0; g < J.G && ok; g++) {
    const int w0 = g * J.Wg, nw = min(J.Wg, J.W - w0);
    if (wait) {
      const unsigned long long t0 = globaltimer_ns();
      for (;;) {
        uint32_t st = 0;
        if (lane == 0) { uint32_t seen = *(volatile uint32_t*)(J.ready + g); st = seen >= (uint32_t)nw ? 1u : (globaltimer_ns() - t0 > 2000000000ull ? 2u : 0u);
                         if (st == 2u) diag = 0x80000000u | ((uint32_t)(threadIdx.x >> 5) << 28) | ((uint32_t)g << 24) | ((uint32_t)nw << 12) | (seen & 0xfffu); }
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st == 1u) break;
        if (st == 2u) { ok = false; break; }
        __nanosleep(200);
      }
      if (!ok) break;
      __threadfence();
    }
    for (int base = 0; base < nw; base += 8) {
      const int j = base + q;
      ge run = ge_identity(), wsum = ge_identity();
      if (j < nw) { ld_ge(run, J.hnodes + 64 * (size_t)(w0 + j)); ld_ge(wsum, J.hnodes + 64 * (size_t)(w0 + j) + 32); }
      ge S = quad_add(run, wsum);
      if (j < nw) st_ge_quad(sS + 32 * j, S);
    }
    __syncwarp();
    for (int w = 0; w < nw; w++) {
      if (g || w) for (int k = 0; k < J.c; k++) acc = quad_double(acc);
      ge S; ld_ge(S, sS + 32 * w);
      acc = quad_add(acc, S);
    }
    __syncwarp();
  }
  if (lane != 0) return;
  if (!ok) { *J.status = diag; if (J.is_identity) *J.is_identity = 2u; return; }      // status: which chain, which group, expected and seen node counts
  if (J.out_ext) { st_fe(J.out_ext, fe_freeze(acc.X)); st_fe(J.out_ext + 8, fe_freeze(acc.Y)); st_fe(J.out_ext + 16, fe_freeze(acc.Z)); st_fe(J.out_ext + 24, fe_freeze(acc.T)); }
  if (J.out_enc) encode_result_words(J.out_enc, acc);
  if (J.is_identity) *J.is_identity = ge_ristretto_is_identity(acc) ? 1u : 0u;
}
__global__ void __launch_bounds__(512, 1) k_msm_chain(ChainJob j0, ChainJob j1, int wait) {
  __shared__ __align__(16) uint32_t sS[2][32 * 32];
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && j0.W > 0) chain_run(j0, sS[0], wait);
  if (warp == 1 && j1.W > 0) chain_run(j1, sS[1], wait);
  __syncthreads();
}
__global__ void k_msm_empty(uint8_t* __restrict__ out_enc, uint32_t* __restrict__ is_identity, uint32_t* __restrict__ out_ext) {
  if (out_enc) { reinterpret_cast<uint4*>(out_enc)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(out_enc)[1] = make_uint4(0, 0, 0, 0); }
  if (is_identity) *is_identity = 1u;
  if (out_ext) st_ge(out_ext, ge_identity());
}

int g_accum_variant = getenv("XHE_MSM_SMEM_ACC") && atoi(getenv("XHE_MSM_SMEM_ACC")) ? 104 : 4;   // resident 128-thread blocks per SM the hot kernel is compiled for (4 -> 128 regs, 6 -> 80, 8 -> 64); 104 = the shared-memory A/B
inline unsigned nblk(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" void xhe_msm_set_variant(int v) { g_accum_variant = v; }
extern "C" size_t xhe_msm_workspace_bytes(const xhe_ctx*, size_t n) { return make_plan(n).total; }
extern "C" int32_t xhe_msm_plan(size_t n, int* c, int* W) { MsmPlan p = make_plan(n); if (c) *c = p.c; if (W) *W = p.W; return XHE_OK; }

// The pipeline in two phases, so that a caller whose scalars are ready before its points (xhe_batch_run) can overlap them:
//   xhe_msm_sort   -- steps 1-4: needs only the scalars (digit recoding, counting sort into bucket order, tile runs)
//   xhe_msm_finish -- accumulation + reduction + Horner: needs the points; d_out_ext (optional) = the un-normalised
//                     extended result (128 B)
namespace {
struct MsmPtrs {
  uint32_t *counts, *offsets, *cursor, *blocksums, *list, *tile_g0, *runs, *run_off, *part, *pstart, *pcount, *heavy, *nodes_a, *nodes_b, *hnodes, *hacc, *ready, *flag;
};
inline MsmPtrs msm_ptrs(const MsmPlan& p, void* d_ws, void* d_bad_flag) {
  uint8_t* ws = (uint8_t*)d_ws;
  MsmPtrs q;
  q.counts = (uint32_t*)(ws + p.off_counts); q.offsets = (uint32_t*)(ws + p.off_offsets); q.cursor = (uint32_t*)(ws + p.off_cursor);
  q.blocksums = (uint32_t*)(ws + p.off_blocksums); q.list = (uint32_t*)(ws + p.off_list); q.tile_g0 = (uint32_t*)(ws + p.off_tileg0);
  q.runs = (uint32_t*)(ws + p.off_runs); q.run_off = (uint32_t*)(ws + p.off_runoff); q.part = (uint32_t*)(ws + p.off_part);
  q.pstart = (uint32_t*)(ws + p.off_pstart); q.pcount = (uint32_t*)(ws + p.off_pcount); q.heavy = (uint32_t*)(ws + p.off_heavy);
  q.nodes_a = (uint32_t*)(ws + p.off_nodes_a); q.nodes_b = (uint32_t*)(ws + p.off_nodes_b); q.hnodes = (uint32_t*)(ws + p.off_hnodes); q.hacc = (uint32_t*)(ws + p.off_hacc); q.ready = (uint32_t*)(ws + p.off_ready);
  q.flag = d_bad_flag ? (uint32_t*)d_bad_flag : (uint32_t*)(ws + p.off_flag);
  return q;
}
}  // namespace

int32_t xhe_msm_sort(xhe_ctx* ctx, const void* d_scalars, size_t n, void* d_ws, size_t ws_bytes, void* d_bad_flag) {
  if (!ctx) return XHE_E_ARG;
  if (n == 0) return XHE_OK;
  if (!d_scalars || !d_ws) return XHE_E_ARG;
  if (n >= (1ull << 31)) return XHE_E_ARG;
  cudaStream_t st = ctx->stream;
  MsmPlan p = make_plan(n);
  if (ws_bytes < p.total) { ctx->err = "msm workspace too small"; return XHE_E_ARG; }
  MsmPtrs q = msm_ptrs(p, d_ws, d_bad_flag);
  const size_t m = p.total_buckets;
  XHE_CUDA_OK(ctx, cudaMemsetAsync(q.counts, 0, 4 * (m + 1), st));
  XHE_CUDA_OK(ctx, cudaMemsetAsync(q.pcount, 0, 4 * m, st));
  if (!d_bad_flag) XHE_CUDA_OK(ctx, cudaMemsetAsync(q.flag, 0, 4, st));
  k_msm_zero_heads<<<1, 32, 0, st>>>(q.heavy, (uint32_t)(p.Wg * p.B), p.G, q.ready); XHE_LAUNCHED(ctx);
  k_msm_count<<<nblk(n, 256), 256, 0, st>>>((const uint32_t*)d_scalars, n, p.c, p.W, p.B, q.counts, q.flag); XHE_LAUNCHED(ctx);
  unsigned nb = nblk(m, SCAN_THREADS * SCAN_ITEMS);
  if (nb > 4096) { ctx->err = "msm: too many buckets"; return XHE_E_ARG; }
  k_scan_blocks<<<nb, SCAN_THREADS, 0, st>>>(q.counts, m, q.offsets, q.blocksums); XHE_LAUNCHED(ctx);
  k_scan_totals<<<1, 1024, 0, st>>>(q.blocksums, (int)nb, q.offsets + m); XHE_LAUNCHED(ctx);
  k_scan_add<<<nb, SCAN_THREADS, 0, st>>>(q.offsets, m, q.blocksums, q.cursor); XHE_LAUNCHED(ctx);
  k_msm_scatter<<<nblk(n, 256), 256, 0, st>>>((const uint32_t*)d_scalars, n, p.c, p.W, p.B, q.cursor, q.list); XHE_LAUNCHED(ctx);
  // offsets[m] = total number of non-zero digits (device-side); tiles beyond it are empty
  k_msm_tile_runs<<<nblk(p.n_tiles + 1, 256), 256, 0, st>>>(q.offsets, (uint32_t)m, p.n_tiles, q.runs, q.tile_g0); XHE_LAUNCHED(ctx);
  unsigned nbt = nblk(p.n_tiles + 1, SCAN_THREADS * SCAN_ITEMS);
  if (nbt > 4096) { ctx->err = "msm: too many tiles"; return XHE_E_ARG; }
  k_scan_blocks<<<nbt, SCAN_THREADS, 0, st>>>(q.runs, p.n_tiles + 1, q.run_off, q.blocksums); XHE_LAUNCHED(ctx);
  k_scan_totals<<<1, 1024, 0, st>>>(q.blocksums, (int)nbt, (uint32_t*)((uint8_t*)d_ws + p.off_flag) + 2); XHE_LAUNCHED(ctx);
  k_scan_add<<<nbt, SCAN_THREADS, 0, st>>>(q.run_off, p.n_tiles + 1, q.blocksums, nullptr); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaGetLastError());
  return XHE_OK;
}

// side streams / events of the grouped tail (lane 0: the caller's first MSM in flight, lane 1: a second one, as in
// xhe_batch_run where the sigma and the range MSM run side by side)
int32_t xhe_msm_side_init(xhe_ctx* ctx, int lane) {
  if (ctx->msm_side[lane][0]) return XHE_OK;
  int lo_pri = 0, hi_pri = 0; XHE_CUDA_OK(ctx, cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
  for (int i = 0; i < 2; i++) XHE_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->msm_side[lane][i], cudaStreamNonBlocking, hi_pri));   // latency chains first
  for (auto& e : ctx->msm_ev[lane]) XHE_CUDA_OK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return XHE_OK;
}

static XheChainJob to_job(const ChainJob& j) { XheChainJob o; o.hnodes = j.hnodes; o.ready = j.ready; o.G = j.G; o.Wg = j.Wg; o.W = j.W; o.c = j.c; o.out_enc = j.out_enc; o.is_identity = j.is_identity; o.out_ext = j.out_ext; o.status = j.status; return o; }
static ChainJob from_job(const XheChainJob& j) { ChainJob o; o.hnodes = j.hnodes; o.ready = j.ready; o.G = j.G; o.Wg = j.Wg; o.W = j.W; o.c = j.c; o.out_enc = j.out_enc; o.is_identity = j.is_identity; o.out_ext = j.out_ext; o.status = j.status; return o; }
// the chain job of the MSM over n points whose workspace is d_ws (W = 0: nothing to do)
XheChainJob xhe_msm_chain_job(size_t n, void* d_ws, void* d_out_enc, void* d_is_id, void* d_out_ext) {
  ChainJob j; memset(&j, 0, sizeof j);
  if (n && d_ws) {
    MsmPlan p = make_plan(n); MsmPtrs q = msm_ptrs(p, d_ws, nullptr);
    j.hnodes = q.hnodes; j.ready = q.ready; j.G = p.G; j.Wg = p.Wg; j.W = p.W; j.c = p.c; j.status = q.ready + 8;
    j.out_enc = (uint8_t*)d_out_enc; j.is_identity = (uint32_t*)d_is_id; j.out_ext = (uint32_t*)d_out_ext;
  }
  return to_job(j);
}
int32_t xhe_msm_chain_reset(xhe_ctx* ctx, cudaStream_t st, const XheChainJob& j) {
  if (j.W > 0) XHE_CUDA_OK(ctx, cudaMemsetAsync(j.ready, 0, 64, st));
  return XHE_OK;
}
int32_t xhe_msm_chain_launch(xhe_ctx* ctx, cudaStream_t st, const XheChainJob& a, const XheChainJob& b, int wait) {
  if (a.W <= 0 && b.W <= 0) return XHE_OK;
  k_msm_chain<<<1, 512, 0, st>>>(from_job(a), from_job(b), wait); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaGetLastError());
  return XHE_OK;
}
// OFF by default (XHE_MSM_CHAIN=1 switches it on).  Measured on B200, 10 k a1k1 batch: the chain itself runs at its isolated
// speed, but the step did not get faster (3.49 ms against 3.32 ms: the reduction kernels that feed it are what the contention
// stretches), and a polling kernel makes every implicit device synchronisation elsewhere in the process (cudaMalloc, pinned
// allocations, a kernel that grows the local-memory pool) a 2-second stall -- DESIGN.md 4.4 has the numbers.
bool xhe_msm_chain_enabled() { static const bool on = getenv("XHE_MSM_CHAIN") && atoi(getenv("XHE_MSM_CHAIN")) != 0; return on; }

// chain_mode: 0 = per-group Horner kernels behind events (also what the serial diagnostics mode uses);
//             1 = this call launches its own polling chain kernel first;
//             2 = the caller has launched a chain kernel that serves this MSM (xhe_batch_run: one kernel for both MSMs) and joins it
int32_t xhe_msm_finish(xhe_ctx* ctx, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id, void* d_out_ext, cudaEvent_t after_accum, int lane, int chain_mode) {
  if (!ctx || lane < 0 || lane > 1) return XHE_E_ARG;
  cudaStream_t st = ctx->stream;
  if (n == 0) { k_msm_empty<<<1, 1, 0, st>>>((uint8_t*)d_out_enc, (uint32_t*)d_is_id, (uint32_t*)d_out_ext); XHE_LAUNCHED(ctx); if (after_accum) XHE_CUDA_OK(ctx, cudaEventRecord(after_accum, st)); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK; }
  if (!d_niels || !d_ws) return XHE_E_ARG;
  MsmPlan p = make_plan(n);
  if (ws_bytes < p.total) { ctx->err = "msm workspace too small"; return XHE_E_ARG; }
  if (p.Wg > 32) { ctx->err = "msm: window group too large"; return XHE_E_ARG; }
  MsmPtrs q = msm_ptrs(p, d_ws, nullptr);
  const size_t m = p.total_buckets;
  if (chain_mode < 0) chain_mode = xhe_msm_chain_enabled() ? 1 : 0;
  if (ctx->serial) chain_mode = 0;
  // Residency limiter: the hot kernel takes the whole register file at 4 blocks of 128 threads x 128 registers per SM, so no
  // block of a concurrently running latency-bound kernel (signatures, reduction tails of the other MSM) can start on an SM
  // while a wave is resident.  The field multiply saturates the multiplier pipe from 2 warps per sub-partition, so a
  // dynamic shared-memory request that caps residency at 3 blocks per SM costs the kernel little and leaves a quarter of
  // the registers to the other streams (XHE_ACCUM_SMEM overrides; 0 = no cap).
  static const size_t accum_smem = []() { const char* e = getenv("XHE_ACCUM_SMEM"); size_t v = e ? (size_t)atol(e) : (size_t)XHE_ACCUM_SMEM_DEFAULT;
    if (v > 48 * 1024) { cudaFuncSetAttribute(k_msm_accum_tiles<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); cudaFuncSetAttribute(k_msm_accum_tiles<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); cudaFuncSetAttribute(k_msm_accum_tiles<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); }
    return v; }();
  // The grouped tail: accumulation launches of the window groups back to back on the caller's stream; behind each, on a
  // high-priority side stream, the group's reduction (s_red); the Horner chain on s_hor (chain_mode 0 / 1) or in the
  // caller's chain kernel (chain_mode 2).
  const bool split = !ctx->serial && (p.G > 1 || chain_mode == 1);
  cudaStream_t s_red = st, s_hor = st;
  cudaEvent_t* ev = nullptr;
  if (split || chain_mode == 2) { int32_t rc = xhe_msm_side_init(ctx, lane); if (rc) return rc; ev = ctx->msm_ev[lane]; }
  if (split) { s_red = ctx->msm_side[lane][0]; s_hor = ctx->msm_side[lane][1]; }
  if (chain_mode == 1) {      // the polling chain goes first: it must own an SM before the accumulation fills the machine
    XheChainJob job = xhe_msm_chain_job(n, d_ws, d_out_enc, d_is_id, d_out_ext), none; memset(&none, 0, sizeof none);
    XHE_CUDA_OK(ctx, cudaEventRecord(ev[2 * MSM_MAX_GROUPS], st)); XHE_CUDA_OK(ctx, cudaStreamWaitEvent(s_hor, ev[2 * MSM_MAX_GROUPS], 0));      // after the sort (which zeroed ready[])
    int32_t rc = xhe_msm_chain_launch(ctx, s_hor, job, none, 1); if (rc) return rc;
  }
  const uint32_t kpg = (uint32_t)(p.Wg * p.B);
  AccumOut ao{q.part, q.pstart, q.pcount, q.heavy, kpg};
  { XheTimed timed(ctx, "k_msm_accum_tiles", 504.0 * (double)n * 16.0);      // canonical units: 16 windows, whatever W the plan picked
    for (int g = 0; g < p.G; g++) {
      const uint32_t klo = (uint32_t)g * kpg, khi = (uint32_t)std::min<size_t>(m, (size_t)(g + 1) * kpg);
      switch (g_accum_variant) {
        case 6: k_msm_accum_tiles<6, false><<<nblk(p.n_tiles, 128), 128, accum_smem, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, klo, khi, ao); break;
        case 8: k_msm_accum_tiles<8, false><<<nblk(p.n_tiles, 128), 128, accum_smem, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, klo, khi, ao); break;
        case 104: k_msm_accum_tiles<4, true><<<nblk(p.n_tiles, 128), 128, 0, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, klo, khi, ao); break;     // A/B: sum in shared memory
        default: k_msm_accum_tiles<4, false><<<nblk(p.n_tiles, 128), 128, accum_smem, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, klo, khi, ao); break;
      }
      XHE_LAUNCHED(ctx);
      if (s_red != st) XHE_CUDA_OK(ctx, cudaEventRecord(ev[g], st));
    } }
  if (after_accum) XHE_CUDA_OK(ctx, cudaEventRecord(after_accum, st));      // the throughput-bound part of this MSM is over: what follows is latency-bound
  for (int g = 0; g < p.G; g++) {
    const uint32_t klo = (uint32_t)g * kpg, khi = (uint32_t)std::min<size_t>(m, (size_t)(g + 1) * kpg);
    const uint32_t nw = (khi - klo) / p.B;
    if (s_red != st) XHE_CUDA_OK(ctx, cudaStreamWaitEvent(s_red, ev[g], 0));
    const uint32_t* heavy_g = q.heavy + (size_t)g * kpg + 8 * g;
    k_msm_fold_heavy<<<(unsigned)std::min<size_t>((khi - klo + 3) / 4, 8 * (size_t)ctx->sm_count), FOLD_THREADS, 0, s_red>>>(q.part, q.pstart, q.pcount, heavy_g); XHE_LAUNCHED(ctx);
    if (p.seg_log == 4) k_msm_bucket_seg<16><<<nblk(khi - klo, 256), 256, 0, s_red>>>(q.part, q.pstart, q.pcount, klo, khi, q.nodes_a);
    else k_msm_bucket_seg<4><<<nblk(khi - klo, 256), 256, 0, s_red>>>(q.part, q.pstart, q.pcount, klo, khi, q.nodes_a);
    XHE_LAUNCHED(ctx);
    uint32_t per_window = (uint32_t)(p.B >> p.seg_log); int child_log = p.seg_log;
    uint32_t *cur = q.nodes_a, *nxt = q.nodes_b;
    uint32_t* hn = q.hnodes + 64 * (size_t)(klo / p.B);
    static const bool nodes_scan_only = getenv("XHE_MSM_NODES_SCAN") != nullptr && atoi(getenv("XHE_MSM_NODES_SCAN")) != 0;      // A/B: parallel-scan levels only
    while (per_window > 32 && !nodes_scan_only) {      // wide levels: work-efficient radix-8 folds
      const uint32_t parents = (per_window + NODES_SEQ_R - 1) / NODES_SEQ_R, n_par = parents * nw;
      k_msm_nodes_seq<<<nblk(8 * (size_t)n_par, 128), 128, 0, s_red>>>(cur, per_window, parents, n_par, child_log, nxt); XHE_LAUNCHED(ctx);
      std::swap(cur, nxt); per_window = parents; child_log += 3;
    }
    while (per_window > 1) {
      const uint32_t parents = (per_window + 31) / 32;
      uint32_t* dst = parents == 1 ? hn : nxt;
      k_msm_nodes32<<<parents * nw, 128, 0, s_red>>>(cur, per_window, parents, child_log, dst, (parents == 1 && chain_mode) ? q.ready + g : nullptr); XHE_LAUNCHED(ctx);
      std::swap(cur, nxt); per_window = parents; child_log += 5;
    }
    if (chain_mode == 0) {
      if (s_hor != s_red) { XHE_CUDA_OK(ctx, cudaEventRecord(ev[MSM_MAX_GROUPS + g], s_red)); XHE_CUDA_OK(ctx, cudaStreamWaitEvent(s_hor, ev[MSM_MAX_GROUPS + g], 0)); }
      const bool last = g == p.G - 1;
      static const bool quad_horner = getenv("XHE_MSM_QUAD_HORNER") != nullptr && atoi(getenv("XHE_MSM_QUAD_HORNER")) != 0;      // A/B: the round-1 chain
      if (p.G == 1 && !quad_horner) { k_msm_horner_oct<<<1, HORNER_OCT_THREADS, 0, s_hor>>>(hn, (int)nw, p.c, (uint8_t*)d_out_enc, (uint32_t*)d_is_id, (uint32_t*)d_out_ext); XHE_LAUNCHED(ctx); }
      else k_msm_horner_g<<<1, 32, 0, s_hor>>>(hn, (int)nw, p.c, g ? q.hacc + 32 * (size_t)(g - 1) : nullptr, q.hacc + 32 * (size_t)g, last ? 1 : 0,
                                          (uint8_t*)d_out_enc, (uint32_t*)d_is_id, (uint32_t*)d_out_ext); XHE_LAUNCHED(ctx);
    }
  }
  // joins come LAST: every producer is already queued when a wait on the chain kernel enters a hardware queue
  if (s_red != st && chain_mode != 0) { XHE_CUDA_OK(ctx, cudaEventRecord(ev[2 * MSM_MAX_GROUPS - 1], s_red)); XHE_CUDA_OK(ctx, cudaStreamWaitEvent(st, ev[2 * MSM_MAX_GROUPS - 1], 0)); }
  if (s_hor != st && chain_mode != 2) { XHE_CUDA_OK(ctx, cudaEventRecord(ev[2 * MSM_MAX_GROUPS], s_hor)); XHE_CUDA_OK(ctx, cudaStreamWaitEvent(st, ev[2 * MSM_MAX_GROUPS], 0)); }
  XHE_CUDA_OK(ctx, cudaGetLastError());
  return XHE_OK;
}

int32_t xhe_launch_msm_ex(xhe_ctx* ctx, const void* d_scalars, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id, void* d_out_ext, void* d_bad_flag) {
  if (n && (!d_scalars || !d_niels || !d_ws)) return XHE_E_ARG;
  int32_t rc = xhe_msm_sort(ctx, d_scalars, n, d_ws, ws_bytes, d_bad_flag); if (rc) return rc;
  return xhe_msm_finish(ctx, d_niels, n, d_ws, ws_bytes, d_out_enc, d_is_id, d_out_ext, nullptr, 0, -1);
}
int32_t xhe_launch_msm(xhe_ctx* ctx, const void* d_scalars, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id) {
  return xhe_launch_msm_ex(ctx, d_scalars, d_niels, n, d_ws, ws_bytes, d_out_enc, d_is_id, nullptr, nullptr);
}
extern "C" int32_t xhe_msm_dev(xhe_ctx* ctx, const void* d_scalars, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc32, void* d_is_identity_u32) {
  return xhe_launch_msm(ctx, d_scalars, d_niels, n, d_ws, ws_bytes, d_out_enc32, d_is_identity_u32);
}
// host-buffer entry point: scalars + compressed points -> encoding of sum s_i P_i and the Ristretto identity verdict
extern "C" int32_t xhe_msm_vartime(xhe_ctx* ctx, const uint8_t* scalars, const uint8_t* enc_points, size_t n, uint8_t out_enc[32], int32_t* is_identity) {
  if (!ctx || !out_enc || (n && (!scalars || !enc_points))) return XHE_E_ARG;
  size_t wsb = make_plan(n).total;
  void *d_s = nullptr, *d_e = nullptr, *d_n = nullptr, *d_ok = nullptr, *d_ws = nullptr, *d_out = nullptr;
  int32_t rc = XHE_OK;
  auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_e); cudaFree(d_n); cudaFree(d_ok); cudaFree(d_ws); cudaFree(d_out); };
#define TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return XHE_E_CUDA; } } while (0)
  TRY(cudaMalloc(&d_s, 32 * n + 32)); TRY(cudaMalloc(&d_e, 32 * n + 32)); TRY(cudaMalloc(&d_n, 96 * n + 96)); TRY(cudaMalloc(&d_ok, n + 4)); TRY(cudaMalloc(&d_ws, wsb)); TRY(cudaMalloc(&d_out, 256));
  TRY(cudaMemcpyAsync(d_s, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream)); TRY(cudaMemcpyAsync(d_e, enc_points, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  TRY(cudaMemsetAsync((uint8_t*)d_out + 64, 0, 4, ctx->stream));
  rc = xhe_decompress_dev(ctx, d_e, n, nullptr, d_n, d_ok); if (rc) { cleanup(); return rc; }
  rc = xhe_launch_msm_ex(ctx, d_s, d_n, n, d_ws, wsb, d_out, (uint8_t*)d_out + 32, nullptr, (uint8_t*)d_out + 64); if (rc) { cleanup(); return rc; }
  uint8_t host[72]; std::vector<uint8_t> ok(n);
  TRY(cudaMemcpyAsync(host, d_out, 68, cudaMemcpyDeviceToHost, ctx->stream)); if (n) TRY(cudaMemcpyAsync(ok.data(), d_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
  TRY(cudaStreamSynchronize(ctx->stream));
  cleanup();
#undef TRY
  uint32_t bad; memcpy(&bad, host + 64, 4);
  if (bad) { ctx->err = "msm: non-canonical scalar"; return XHE_E_ARG; }
  for (size_t i = 0; i < n; i++) if (!ok[i]) { ctx->err = "msm: invalid point encoding at index " + std::to_string(i); return XHE_E_ARG; }
  memcpy(out_enc, host, 32);
  { uint32_t f; memcpy(&f, host + 32, 4); if (f == 2u) { ctx->err = "msm: the Horner chain kernel timed out waiting for its inputs"; return XHE_E_CUDA; } if (is_identity) *is_identity = (int32_t)f; }
  return XHE_OK;
}

// self-test of the warp-cooperative arithmetic (oct.cuh) against host big integers: op 0 mul, 1 add, 2 sub over n field elements
// (8 words each); op 3 doubling, 4 addition over n extended points (32 words each: X, Y, Z, T)
extern "C" int32_t xhe_selftest_oct(xhe_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
  if (!ctx || !a || !b || !out || op < 0 || op > 4) return XHE_E_ARG;
  const size_t words = (op < 3 ? 8 : 32) * n;
  uint32_t *da = nullptr, *db = nullptr, *dout = nullptr;
  XHE_CUDA_OK(ctx, cudaMalloc(&da, 4 * words + 16)); XHE_CUDA_OK(ctx, cudaMalloc(&db, 4 * words + 16)); XHE_CUDA_OK(ctx, cudaMalloc(&dout, 4 * words + 16));
  XHE_CUDA_OK(ctx, cudaMemcpyAsync(da, a, 4 * words, cudaMemcpyHostToDevice, ctx->stream)); XHE_CUDA_OK(ctx, cudaMemcpyAsync(db, b, 4 * words, cudaMemcpyHostToDevice, ctx->stream));
  if (words) { k_selftest_oct<<<nblk(words, 128), 128, 0, ctx->stream>>>(op, da, db, n, dout); XHE_LAUNCHED(ctx); }
  XHE_CUDA_OK(ctx, cudaGetLastError()); XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  XHE_CUDA_OK(ctx, cudaMemcpy(out, dout, 4 * words, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout); return XHE_OK;
}

// latency of the warp-cooperative operations (design evidence, tools/op_bench.py): one warp, `iters` dependent operations
template <int OP>
__global__ void __launch_bounds__(32) k_bench_oct(uint32_t* out, int iters, unsigned long long* cycles) {
  uint32_t p = 0x9e3779b9u * (threadIdx.x + 1) + 12345u, q = p * 2654435761u + 7u;
  if ((threadIdx.x & 7u) == 7u) { p &= 0x7fffffffu; q &= 0x7fffffffu; }
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
    if (OP == 0) p = oct_mul(p, q);
    else if (OP == 1) p = oct_double(p);
    else if (OP == 2) p = oct_add_pt(p, q);
    else p = oct_add(p, q);
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  if (p == 0x12345u) out[0] = p;
}
extern "C" int32_t xhe_bench_oct(xhe_ctx* ctx, int op, int iters, double* cycles_per_op) {
  if (!ctx || !cycles_per_op || op < 0 || op > 3 || iters <= 0) return XHE_E_ARG;
  uint32_t* d_out; unsigned long long* d_c;
  XHE_CUDA_OK(ctx, cudaMalloc(&d_out, 64)); XHE_CUDA_OK(ctx, cudaMalloc(&d_c, 8));
  for (int rep = 0; rep < 2; rep++) {
    switch (op) {
      case 0: k_bench_oct<0><<<1, 32, 0, ctx->stream>>>(d_out, iters, d_c); break;
      case 1: k_bench_oct<1><<<1, 32, 0, ctx->stream>>>(d_out, iters, d_c); break;
      case 2: k_bench_oct<2><<<1, 32, 0, ctx->stream>>>(d_out, iters, d_c); break;
      default: k_bench_oct<3><<<1, 32, 0, ctx->stream>>>(d_out, iters, d_c); break;
    }
    XHE_LAUNCHED(ctx);
    XHE_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  unsigned long long c; XHE_CUDA_OK(ctx, cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost));
  *cycles_per_op = (double)c / iters;
  cudaFree(d_out); cudaFree(d_c); return XHE_OK;
}

// CUDA loads kernels lazily (CUDA_MODULE_LOADING=LAZY is the default since 12.2), and loading one may need every running kernel
// to finish first: with the polling chain kernel of msm.cu in flight, the FIRST launch of any other kernel would wait for a
// kernel that is waiting for it.  Every kernel of this file is therefore loaded when the first context is created.
size_t xhe_preload_msm() {      // returns the largest per-thread local-memory frame among them
  const void* ks[] = {(const void*)k_msm_count, (const void*)k_msm_scatter, (const void*)k_scan_blocks, (const void*)k_scan_totals, (const void*)k_scan_add, (const void*)k_msm_tile_runs, (const void*)k_msm_accum_tiles<4, false>, (const void*)k_msm_accum_tiles<6, false>, (const void*)k_msm_accum_tiles<8, false>, (const void*)k_msm_accum_tiles<4, true>, (const void*)k_msm_zero_heads, (const void*)k_msm_fold_heavy, (const void*)k_msm_bucket_seg<4>, (const void*)k_msm_bucket_seg<16>, (const void*)k_msm_nodes32, (const void*)k_msm_nodes_seq, (const void*)k_msm_horner_g, (const void*)k_msm_horner_oct, (const void*)k_selftest_oct, (const void*)k_msm_chain, (const void*)k_msm_empty};
  cudaFuncAttributes a; size_t mx = 0;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) == cudaSuccess && a.localSizeBytes > mx) mx = a.localSizeBytes;
  return mx;
}
