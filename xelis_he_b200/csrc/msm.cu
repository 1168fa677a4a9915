// msm.cu -- K6/K7: variable-time multiscalar multiplication (signed-digit Pippenger) for sm_100a.
//
// Replaces RistrettoPoint::vartime_multiscalar_mul + is_identity (reference src/proofs.rs:49-67; dalek's
// Straus/Pippenger behind it) and the MSM inside bulletproofs' verify_batch (src/tx/verify.rs:504-514).
// Not a port of dalek's w <= 8 column loop: the whole scalar is recoded into W = ceil(254/c) signed c-bit digits
// (c chosen per n, up to 16+), every (window, bucket) pair is an independent work item, and the pipeline is
//
//   k_msm_count     thread/point : recode scalar -> W signed digits, histogram (window,bucket) sizes      [L2 atomics]
//   scan            3 kernels    : exclusive prefix sum over the W*2^(c-1) bucket sizes
//   k_msm_scatter   thread/point : counting-sort point indices (sign in bit 31) into bucket order          [L2 atomics]
//   k_msm_sizeperm  thread/bucket: order buckets by size (descending) so the 32 lanes of a warp loop equally
//   k_msm_accum     thread/bucket: gather 96-byte affine-Niels points (6 x LDG.128, next point prefetched),
//                                  7 M mixed additions into a register-resident extended accumulator        [HOT: IMAD pipe]
//   k_msm_seg       thread/8 buckets: running-sum reduction of 8 consecutive buckets -> (run, wsum) node
//   k_msm_nodes     warp/32 nodes: warp-shuffle suffix-scan + tree reduction of nodes (repeated until 1 node/window)
//   k_msm_horner    1 thread     : sum_w 2^(c w) S_w by Horner, ristretto encode, identity flag
//
// Algorithmic work: n*W mixed adds of 7 M = 504 limb products each dominate.
#include <stdlib.h>
#include "xhe_internal.cuh"
#include "quad.cuh"
#include <algorithm>
#include <vector>
#include <string.h>
using namespace xhe;

namespace {

#define MSM_TILE 32   // entries per accumulation work item
#define XHE_ACCUM_SMEM_DEFAULT 0

struct MsmPlan {
  int c, W;             // window bits, windows
  uint32_t B;           // buckets per window = 2^(c-1)
  size_t total_buckets; // W * B
  int seg_log;          // log2 of buckets per level-1 segment
  // workspace offsets (bytes)
  size_t n_tiles, max_runs;
  size_t off_counts, off_offsets, off_cursor, off_blocksums, off_list, off_tileg0, off_runs, off_runoff, off_part, off_partg, off_pstart, off_pcount, off_heavy, off_nodes_a, off_nodes_b, off_flag, total;
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
  static const int seg_env = getenv("XHE_MSM_SEG_LOG") ? atoi(getenv("XHE_MSM_SEG_LOG")) : 3;   // buckets per running-sum thread = 2^seg_log
  p.seg_log = std::min(seg_env, bc - 1);
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
  p.off_partg = o; o = align_up(o + 4 * p.max_runs, 256);
  p.off_pstart = o; o = align_up(o + 4 * p.total_buckets, 256);
  p.off_pcount = o; o = align_up(o + 4 * p.total_buckets, 256);
  p.off_heavy = o; o = align_up(o + 4 * (p.total_buckets + 4), 256);
  size_t nodes1 = p.total_buckets >> p.seg_log;
  p.off_nodes_a = o; o = align_up(o + 256 * nodes1, 256);
  p.off_nodes_b = o; o = align_up(o + 256 * ((nodes1 + 31) / 32 + (size_t)p.W), 256);
  p.off_flag = o; o = align_up(o + 64, 256);
  p.total = o;
  return p;
}

// ---- digit recoding ----------------------------------------------------------------------------------------------
// signed radix-2^c digits d_w in [-2^(c-1), 2^(c-1)], sum d_w 2^(c w) = s, for s < 2^253 and c*W >= 254 (the top digit
// absorbs the final carry without overflow).
template <typename F>
__device__ __forceinline__ void for_each_digit(const uint32_t s[8], int c, int W, F&& f) {
  uint32_t carry = 0;
  const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
  for (int w = 0; w < W; w++) {
    int bit = w * c, limb = bit >> 5, sh = bit & 31;
    uint32_t v = 0;
    if (limb < 8) {
      v = s[limb] >> sh;
      if (sh + c > 32 && limb + 1 < 8) v |= s[limb + 1] << (32 - sh);
    }
    v = (v & mask) + carry;
    carry = v > half ? 1u : 0u;          // v in [0, 2^c]; digits above half become negative with a carry
    int32_t d = carry ? (int32_t)v - (int32_t)(1u << c) : (int32_t)v;
    f(w, d);
  }
}

__global__ void __launch_bounds__(256) k_msm_count(const uint32_t* __restrict__ scalars, size_t n, int c, int W, uint32_t B, uint32_t* __restrict__ counts, uint32_t* __restrict__ bad_flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  { uint4 a = __ldg(reinterpret_cast<const uint4*>(scalars + 8 * i)), b = __ldg(reinterpret_cast<const uint4*>(scalars + 8 * i) + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w; }
  if (sc_geq_l(s)) { atomicOr(bad_flag, 1u); return; }   // non-canonical scalar: reported as a bad argument, contributes nothing
  for_each_digit(s, c, W, [&](int w, int32_t d) {
    if (d != 0) atomicAdd(&counts[(size_t)w * B + (uint32_t)((d < 0 ? -d : d) - 1)], 1u);
  });
}

__global__ void __launch_bounds__(256) k_msm_scatter(const uint32_t* __restrict__ scalars, size_t n, int c, int W, uint32_t B, uint32_t* __restrict__ cursor, uint32_t* __restrict__ list) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  { uint4 a = __ldg(reinterpret_cast<const uint4*>(scalars + 8 * i)), b = __ldg(reinterpret_cast<const uint4*>(scalars + 8 * i) + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w; }
  if (sc_geq_l(s)) return;
  for_each_digit(s, c, W, [&](int w, int32_t d) {
    if (d != 0) {
      uint32_t g = (uint32_t)w * B + (uint32_t)((d < 0 ? -d : d) - 1);
      uint32_t pos = atomicAdd(&cursor[g], 1u);
      list[pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);     // the bucket of a position follows from the offsets: no second array
    }
  });
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
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_accum_tiles(const uint32_t* __restrict__ niels, const uint32_t* __restrict__ list, const uint32_t* __restrict__ offsets, uint32_t m,
                                                              const uint32_t* __restrict__ tile_g0, const uint32_t* __restrict__ run_off, size_t n_tiles,
                                                              uint32_t* __restrict__ part, uint32_t* __restrict__ part_g) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  uint32_t N = __ldg(offsets + m);
  size_t start = t * MSM_TILE;
  if (start >= N) return;
  uint32_t cnt = (uint32_t)(min((size_t)N, start + MSM_TILE) - start);
  uint32_t slot = run_off[t];
  uint32_t g = __ldg(tile_g0 + t), nb = __ldg(offsets + g + 1);          // nb: first position past the current bucket
  uint32_t e = __ldg(list + start);
  ge_niels q; ld_niels(q, niels + 24 * (size_t)(e & 0x7fffffffu));
  ge acc = ge_from_niels(niels_cneg(q, (e >> 31) != 0));
  if (cnt > 1) { e = __ldg(list + start + 1); ld_niels(q, niels + 24 * (size_t)(e & 0x7fffffffu)); }
  for (uint32_t j = 1; j < cnt; j++) {
    const uint32_t pos = (uint32_t)start + j;
    ge_niels cur = niels_cneg(q, (e >> 31) != 0);
    if (j + 1 < cnt) { e = __ldg(list + start + j + 1); ld_niels(q, niels + 24 * (size_t)(e & 0x7fffffffu)); }
    if (pos == nb) {   // bucket boundary: flush the finished run, restart from this point
      st_ge(part + 32 * (size_t)slot, acc); part_g[slot] = g; slot++;
      g = next_bucket(offsets, m, pos, g); nb = __ldg(offsets + g + 1);
      acc = ge_from_niels(cur);
    } else {
      acc = ge_madd(acc, cur);
    }
  }
  st_ge(part + 32 * (size_t)slot, acc); part_g[slot] = g;
}

// partial sums are in bucket order; record where each bucket's partials start and how many there are
__global__ void __launch_bounds__(256) k_msm_bucket_index(const uint32_t* __restrict__ part_g, const uint32_t* __restrict__ total_runs, size_t max_runs, uint32_t* __restrict__ pstart, uint32_t* __restrict__ pcount) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= max_runs || s >= *total_runs) return;
  uint32_t g = part_g[s];
  atomicAdd(&pcount[g], 1u);
  if (s == 0 || part_g[s - 1] != g) pstart[g] = (uint32_t)s;
}

// buckets whose partial list is long (under-filled top window, skewed scalars) are folded by a whole block first
#define HEAVY_PARTIALS 6
__global__ void __launch_bounds__(256) k_msm_find_heavy(const uint32_t* __restrict__ pcount, size_t m, uint32_t* __restrict__ heavy /* [0] = count, [1..] = bucket ids */) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < m && pcount[g] > HEAVY_PARTIALS) heavy[1 + atomicAdd(&heavy[0], 1u)] = (uint32_t)g;
}
#define FOLD_THREADS 128
__global__ void __launch_bounds__(FOLD_THREADS) k_msm_fold_heavy(uint32_t* __restrict__ part, const uint32_t* __restrict__ pstart, uint32_t* __restrict__ pcount, const uint32_t* __restrict__ heavy) {
  __shared__ uint32_t sm[FOLD_THREADS * 32];
  uint32_t nh = heavy[0];
  for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
    uint32_t g = heavy[1 + h], ps = pstart[g], pc = pcount[g];
    ge acc = ge_identity(), s;
    for (uint32_t j = threadIdx.x; j < pc; j += FOLD_THREADS) { ld_ge(s, part + 32 * (size_t)(ps + j)); acc = ge_add(acc, s); }
    st_ge(sm + 32 * threadIdx.x, acc);
    __syncthreads();
    for (int stride = FOLD_THREADS / 2; stride >= 1; stride >>= 1) {
      if ((int)threadIdx.x < stride && threadIdx.x + stride < min(pc, (uint32_t)FOLD_THREADS)) {
        ge a, b; ld_ge(a, sm + 32 * threadIdx.x); ld_ge(b, sm + 32 * (threadIdx.x + stride));
        st_ge(sm + 32 * threadIdx.x, ge_add(a, b));
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) { ge r; ld_ge(r, sm); st_ge(part + 32 * (size_t)ps, r); pcount[g] = 1; }
    __syncthreads();
  }
}

// ---- bucket reduction -----------------------------------------------------------------------------------------------
// node = (run, wsum): run = sum of the bucket sums in its range, wsum = sum (b - lo) * S_b (weights relative to the range)
__global__ void __launch_bounds__(128) k_msm_seg(const uint32_t* __restrict__ part, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pcount,
                                                 size_t n_nodes, int seg_log, uint32_t* __restrict__ nodes) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes) return;
  const int L = 1 << seg_log;
  size_t b0 = t << seg_log;
  ge run = ge_identity(), wsum = ge_identity(), s;
  for (int r = L - 1; r >= 0; r--) {
    uint32_t pc = pcount[b0 + r];
    if (pc) {
      uint32_t ps = pstart[b0 + r];
      for (uint32_t j = 0; j < pc; j++) { ld_ge(s, part + 32 * (size_t)(ps + j)); run = ge_add(run, s); }
    }
    if (r >= 1) wsum = ge_add(wsum, run);
  }
  st_ge(nodes + 64 * t, run); st_ge(nodes + 64 * t + 32, wsum);
}

__device__ __forceinline__ ge shfl_down_ge(const ge& p, int d) {
  ge r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], d); r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], d);
    r.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], d); r.T.v[i] = __shfl_down_sync(0xffffffffu, p.T.v[i], d);
  }
  return r;
}
// one warp folds 32 consecutive child nodes (each of width 2^child_log buckets) of ONE window into a parent node.
// nodes_per_window_in children per window; parents per window = ceil(children / 32).
__global__ void __launch_bounds__(128) k_msm_nodes(const uint32_t* __restrict__ in, uint32_t children_per_window, uint32_t parents_per_window, int W, int child_log, uint32_t* __restrict__ out) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= parents_per_window * (uint32_t)W) return;
  uint32_t w = warp / parents_per_window, pidx = warp % parents_per_window;
  uint32_t child = pidx * 32 + lane;
  ge run, wsum;
  if (child < children_per_window) { const uint32_t* p = in + 64 * ((size_t)w * children_per_window + child); ld_ge(run, p); ld_ge(wsum, p + 32); }
  else { run = ge_identity(); wsum = ge_identity(); }
  // suffix sums of run: suf_c = sum_{j >= c} run_j
  ge suf = run;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) { ge t = shfl_down_ge(suf, d); ge a = ge_add(suf, t); bool take = lane + d < 32; suf.X = fe_select(suf.X, a.X, take); suf.Y = fe_select(suf.Y, a.Y, take); suf.Z = fe_select(suf.Z, a.Z, take); suf.T = fe_select(suf.T, a.T, take); }
  // A = sum_{c >= 1} suf_c = sum_c c * run_c ;  Ws = sum_c wsum_c  (both by tree reduction towards lane 0)
  ge A = suf;
  if (lane == 0) A = ge_identity();
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) { A = ge_add(A, shfl_down_ge(A, d)); wsum = ge_add(wsum, shfl_down_ge(wsum, d)); }
  if (lane == 0) {
    for (int k = 0; k < child_log; k++) A = ge_double(A);
    wsum = ge_add(wsum, A);
    uint32_t* o = out + 64 * ((size_t)w * parents_per_window + pidx);
    st_ge(o, suf); st_ge(o + 32, wsum);
  }
}

// final: one node per window.  S_w = wsum + run (bucket b carries multiplier b+1); result = sum_w 2^(c w) S_w by Horner.
// 254 sequential doublings: a pure latency chain, so one quad (4 lanes) shares every point operation (quad.cuh).
__global__ void __launch_bounds__(32) k_msm_horner(const uint32_t* __restrict__ nodes, int W, int c, uint8_t* __restrict__ out_enc, uint32_t* __restrict__ is_identity, uint32_t* __restrict__ out_ext) {
  ge acc = ge_identity();
  for (int w = W - 1; w >= 0; w--) {
    if (w != W - 1) for (int k = 0; k < c; k++) acc = quad_double(acc);
    ge run, wsum; ld_ge(run, nodes + 64 * (size_t)w); ld_ge(wsum, nodes + 64 * (size_t)w + 32);
    acc = quad_add(acc, quad_add(run, wsum));
  }
  if (threadIdx.x != 0) return;
  if (out_ext) {   // canonical coordinates so any consumer (other ranks, the CPU oracle) can read them
    st_fe(out_ext, fe_freeze(acc.X)); st_fe(out_ext + 8, fe_freeze(acc.Y)); st_fe(out_ext + 16, fe_freeze(acc.Z)); st_fe(out_ext + 24, fe_freeze(acc.T));
  }
  if (out_enc) encode_words(out_enc, acc);
  if (is_identity) *is_identity = ge_ristretto_is_identity(acc) ? 1u : 0u;
}
__global__ void k_msm_empty(uint8_t* __restrict__ out_enc, uint32_t* __restrict__ is_identity, uint32_t* __restrict__ out_ext) {
  if (out_enc) { reinterpret_cast<uint4*>(out_enc)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(out_enc)[1] = make_uint4(0, 0, 0, 0); }
  if (is_identity) *is_identity = 1u;
  if (out_ext) st_ge(out_ext, ge_identity());
}

int g_accum_variant = 4;   // resident 128-thread blocks per SM the hot kernel is compiled for (4 -> 128 regs, 6 -> 80, 8 -> 64)
inline unsigned nblk(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" void xhe_msm_set_variant(int v) { g_accum_variant = v; }
extern "C" size_t xhe_msm_workspace_bytes(const xhe_ctx*, size_t n) { return make_plan(n).total; }
extern "C" int32_t xhe_msm_plan(size_t n, int* c, int* W) { MsmPlan p = make_plan(n); if (c) *c = p.c; if (W) *W = p.W; return XHE_OK; }

// The pipeline in two phases, so that a caller whose scalars are ready before its points (xhe_batch_run) can overlap them:
//   xhe_msm_sort   -- steps 1-4: needs only the scalars (digit recoding, counting sort into bucket order, tile runs)
//   xhe_msm_finish -- steps 5-9: needs the points; d_out_ext (optional) = the un-normalised extended result (128 B)
namespace {
struct MsmPtrs {
  uint32_t *counts, *offsets, *cursor, *blocksums, *list, *tile_g0, *runs, *run_off, *part, *part_g, *pstart, *pcount, *heavy, *nodes_a, *nodes_b, *flag;
};
inline MsmPtrs msm_ptrs(const MsmPlan& p, void* d_ws, void* d_bad_flag) {
  uint8_t* ws = (uint8_t*)d_ws;
  MsmPtrs q;
  q.counts = (uint32_t*)(ws + p.off_counts); q.offsets = (uint32_t*)(ws + p.off_offsets); q.cursor = (uint32_t*)(ws + p.off_cursor);
  q.blocksums = (uint32_t*)(ws + p.off_blocksums); q.list = (uint32_t*)(ws + p.off_list); q.tile_g0 = (uint32_t*)(ws + p.off_tileg0);
  q.runs = (uint32_t*)(ws + p.off_runs); q.run_off = (uint32_t*)(ws + p.off_runoff); q.part = (uint32_t*)(ws + p.off_part); q.part_g = (uint32_t*)(ws + p.off_partg);
  q.pstart = (uint32_t*)(ws + p.off_pstart); q.pcount = (uint32_t*)(ws + p.off_pcount); q.heavy = (uint32_t*)(ws + p.off_heavy);
  q.nodes_a = (uint32_t*)(ws + p.off_nodes_a); q.nodes_b = (uint32_t*)(ws + p.off_nodes_b);
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
  XHE_CUDA_OK(ctx, cudaMemsetAsync(q.heavy, 0, 4, st));
  if (!d_bad_flag) XHE_CUDA_OK(ctx, cudaMemsetAsync(q.flag, 0, 4, st));
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

int32_t xhe_msm_finish(xhe_ctx* ctx, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id, void* d_out_ext, cudaEvent_t after_accum) {
  if (!ctx) return XHE_E_ARG;
  cudaStream_t st = ctx->stream;
  if (n == 0) { k_msm_empty<<<1, 1, 0, st>>>((uint8_t*)d_out_enc, (uint32_t*)d_is_id, (uint32_t*)d_out_ext); XHE_LAUNCHED(ctx); if (after_accum) XHE_CUDA_OK(ctx, cudaEventRecord(after_accum, st)); XHE_CUDA_OK(ctx, cudaGetLastError()); return XHE_OK; }
  if (!d_niels || !d_ws) return XHE_E_ARG;
  MsmPlan p = make_plan(n);
  if (ws_bytes < p.total) { ctx->err = "msm workspace too small"; return XHE_E_ARG; }
  MsmPtrs q = msm_ptrs(p, d_ws, nullptr);
  const size_t m = p.total_buckets;
  // Residency limiter: the hot kernel takes the whole register file at 4 blocks of 128 threads x 128 registers per SM, so no
  // block of a concurrently running latency-bound kernel (signatures, reduction tails of the other MSM) can start on an SM
  // while a wave is resident.  The field multiply saturates the multiplier pipe from 2 warps per sub-partition, so a
  // dynamic shared-memory request that caps residency at 3 blocks per SM costs the kernel little and leaves a quarter of
  // the registers to the other streams (XHE_ACCUM_SMEM overrides; 0 = no cap).
  static const size_t accum_smem = []() { const char* e = getenv("XHE_ACCUM_SMEM"); size_t v = e ? (size_t)atol(e) : (size_t)XHE_ACCUM_SMEM_DEFAULT;
    if (v > 48 * 1024) { cudaFuncSetAttribute(k_msm_accum_tiles<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); cudaFuncSetAttribute(k_msm_accum_tiles<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); cudaFuncSetAttribute(k_msm_accum_tiles<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v); }
    return v; }();
  { XheTimed timed(ctx, "k_msm_accum_tiles", 504.0 * (double)n * p.W);
  switch (g_accum_variant) {
    case 6: k_msm_accum_tiles<6><<<nblk(p.n_tiles, 128), 128, accum_smem, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, q.part, q.part_g); break;
    case 8: k_msm_accum_tiles<8><<<nblk(p.n_tiles, 128), 128, accum_smem, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, q.part, q.part_g); break;
    default: k_msm_accum_tiles<4><<<nblk(p.n_tiles, 128), 128, accum_smem, st>>>((const uint32_t*)d_niels, q.list, q.offsets, (uint32_t)m, q.tile_g0, q.run_off, p.n_tiles, q.part, q.part_g); break;
  } }
  XHE_LAUNCHED(ctx);
  if (after_accum) XHE_CUDA_OK(ctx, cudaEventRecord(after_accum, st));      // the throughput-bound part of this MSM is over: what follows is latency-bound
  k_msm_bucket_index<<<nblk(p.max_runs, 256), 256, 0, st>>>(q.part_g, q.run_off + p.n_tiles, p.max_runs, q.pstart, q.pcount); XHE_LAUNCHED(ctx);
  k_msm_find_heavy<<<nblk(m, 256), 256, 0, st>>>(q.pcount, m, q.heavy); XHE_LAUNCHED(ctx);
  k_msm_fold_heavy<<<std::min<size_t>(m, 2 * (size_t)ctx->sm_count), FOLD_THREADS, 0, st>>>(q.part, q.pstart, q.pcount, q.heavy); XHE_LAUNCHED(ctx);
  size_t n_nodes = m >> p.seg_log;
  k_msm_seg<<<nblk(n_nodes, 128), 128, 0, st>>>(q.part, q.pstart, q.pcount, n_nodes, p.seg_log, q.nodes_a); XHE_LAUNCHED(ctx);
  uint32_t per_window = (uint32_t)(p.B >> p.seg_log); int child_log = p.seg_log;
  uint32_t *cur = q.nodes_a, *nxt = q.nodes_b;
  while (per_window > 1) {
    uint32_t parents = (per_window + 31) / 32;
    size_t warps = (size_t)parents * p.W;
    k_msm_nodes<<<nblk(warps * 32, 128), 128, 0, st>>>(cur, per_window, parents, p.W, child_log, nxt); XHE_LAUNCHED(ctx);
    std::swap(cur, nxt); per_window = parents; child_log += 5;
  }
  k_msm_horner<<<1, 32, 0, st>>>(cur, p.W, p.c, (uint8_t*)d_out_enc, (uint32_t*)d_is_id, (uint32_t*)d_out_ext); XHE_LAUNCHED(ctx);
  XHE_CUDA_OK(ctx, cudaGetLastError());
  return XHE_OK;
}

int32_t xhe_launch_msm_ex(xhe_ctx* ctx, const void* d_scalars, const void* d_niels, size_t n, void* d_ws, size_t ws_bytes, void* d_out_enc, void* d_is_id, void* d_out_ext, void* d_bad_flag) {
  if (n && (!d_scalars || !d_niels || !d_ws)) return XHE_E_ARG;
  int32_t rc = xhe_msm_sort(ctx, d_scalars, n, d_ws, ws_bytes, d_bad_flag); if (rc) return rc;
  return xhe_msm_finish(ctx, d_niels, n, d_ws, ws_bytes, d_out_enc, d_is_id, d_out_ext, nullptr);
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
  if (is_identity) { uint32_t f; memcpy(&f, host + 32, 4); *is_identity = (int32_t)f; }
  return XHE_OK;
}
