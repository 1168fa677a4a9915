// imad_probe.cu -- design evidence for the field layer (not part of the product): which forms of the 32x32->64
// multiply-add issue at which rate on sm_100a, whether the ALU pipe (IADD3 carry chains) overlaps the multiplier pipe,
// and what that means for a radix-2^32 field multiply whose partial products are split between
//   chain form  : IMAD.WIDE.U32(.X) with the carry in a predicate (0 ALU instructions, half-rate multiplier slots) and
//   product form: IMAD.WIDE.U32 Rd, Ra, Rb, RZ (2 register reads) + IADD3(.X) carry chains on the ALU pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I xelis_he_b200/csrc -o tools/imad_probe tools/imad_probe.cu
// Run  : tools/imad_probe > gpurun_out/imad_probe.json
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fe25519.cuh"
#define xhe xhe29
#include "experimental/fe25519_r29.cuh.txt"   // the radix-2^29 experiment (9 limbs, carry-free 64-bit columns)
#undef xhe
using namespace xhe;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

// ---------------------------------------------------------------------------------------------------------------
// raw instruction rates: 8 independent chains per thread
// ---------------------------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) k_raw(uint32_t* out, const uint32_t* in, int iters) {
  uint32_t a = in[threadIdx.x & 31] | 1u, bu = in[40] | 1u;   // a: per-thread, bu: warp-uniform
  uint32_t av[8], bv[8];
  unsigned long long w[8];
  uint32_t lo[8], hi[8];
  const double dA = 1.0 + (double)(in[1] & 0xff) * 1e-9, dB = (double)(in[2] & 0xff) * 1e-3;
#pragma unroll
  for (int k = 0; k < 8; k++) { av[k] = in[(threadIdx.x + k) & 63] | 1u; bv[k] = in[(threadIdx.x + 3 * k + 1) & 63] | 1u; w[k] = a + k; lo[k] = a + k; hi[k] = k; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (WHICH == 0) {          // accumulate form, one multiplicand warp-uniform, the other = low word of the neighbouring chain
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((uint32_t)w[(k + 1) & 7]), "r"(bu));
      } else if (WHICH == 1) {   // accumulate form, two vector multiplicands
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((uint32_t)w[(k + 1) & 7]), "r"(bv[k]));
      } else if (WHICH == 2) {   // accumulate form, one vector multiplicand shared by 8 consecutive instructions (.reuse candidate)
        a += (uint32_t)w[7];
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a), "r"(bv[k]));
      } else if (WHICH == 3) {   // product form (addend RZ), two vector multiplicands; both result words stay live
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[k]) : "r"((uint32_t)w[k]), "r"((uint32_t)(w[(k + 1) & 7] >> 32)));
      } else if (WHICH == 7) {   // product form, one multiplicand shared by 8 consecutive instructions
        a += (uint32_t)w[7];
#pragma unroll
        for (int k = 0; k < 8; k++) { unsigned long long t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)(w[k] >> 32)), "r"(a)); w[k] = t | 1ull; }
      } else if (WHICH == 4) {   // product form + a two-word carry add per product: do the pipes overlap?
#pragma unroll
        for (int k = 0; k < 8; k++) {
          unsigned long long t;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(lo[k]), "r"(bv[k]));
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[k]), "+r"(hi[k]) : "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
        }
      } else if (WHICH == 5) {   // carry-chained wide mads (what the current field multiply issues)
        asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.u32 %3, %10, %9, %3;"
                     "mad.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                     : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]) : "r"(av[0]), "r"(bv[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]));
        asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.u32 %3, %10, %9, %3;"
                     "mad.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                     : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7]) : "r"(av[4]), "r"(bv[1]), "r"(av[5]), "r"(av[6]), "r"(av[7]));
      } else if (WHICH == 6) {   // ALU only: 8-word carry chains (IADD3 / IADD3.X)
        asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;"
                     "addc.cc.u32 %4, %4, %12; addc.cc.u32 %5, %5, %13; addc.cc.u32 %6, %6, %14; addc.u32 %7, %7, %15;"
                     : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7])
                     : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]));
        asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;"
                     "addc.cc.u32 %4, %4, %12; addc.cc.u32 %5, %5, %13; addc.cc.u32 %6, %6, %14; addc.u32 %7, %7, %15;"
                     : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7])
                     : "r"(av[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]), "r"(av[4]), "r"(av[5]), "r"(av[6]), "r"(av[7]));
      } else if (WHICH == 10) {  // DFMA (FP64 pipe), 8 dependent chains: the instruction a 52-bit-limb multiply would use (not taken: north_star)
#pragma unroll
        for (int k = 0; k < 8; k++) { double x = __longlong_as_double((long long)w[k]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(dA), "d"(dB)); w[k] = (unsigned long long)__double_as_longlong(x); }
      } else if (WHICH == 9) {   // accumulate form written in C (ptxas keeps IMAD.WIDE Rd, Ra, Rb, Rc64 here)
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] += (unsigned long long)(uint32_t)w[(k + 1) & 7] * bv[k];
      } else if (WHICH == 8) {   // chain form and ALU chains side by side (1 multiply : 1 add)
        asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.u32 %3, %10, %9, %3;"
                     "mad.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                     : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]) : "r"(av[0]), "r"(bv[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]));
        asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %6; addc.u32 %3, %3, %7;"
                     : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]) : "r"(av[4]), "r"(av[5]), "r"(av[6]), "r"(av[7]));
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) r ^= (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32) ^ lo[k] ^ hi[k];
  if (r == 0x12345678u) out[0] = r;
}

// ---------------------------------------------------------------------------------------------------------------
// field multiply / square variants, cycles per warp-operation per SMSP at 8 warps per SMSP
//   0: current radix-2^32 multiply   1: current square
//   2: radix-2^29 multiply as plain C (ptxas keeps accumulate-form IMAD.WIDE)
//   3: radix-2^29 multiply with the products issued row by row (ptxas turns them into IMAD.WIDE ..., RZ + 3-input IADD3)
// (a radix-2^32 variant that splits chains into mul.wide + add.cc was tried here too: ptxas fuses it back into
//  IMAD.WIDE.U32.X, the SASS is identical to variant 0)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void madw29(unsigned long long& c, uint32_t a, uint32_t b) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b)); }
__device__ __forceinline__ xhe29::fe fe_mul_row29(const xhe29::fe& a, const xhe29::fe& b) {
  uint64_t c[17];
#pragma unroll
  for (int k = 0; k < 17; k++) c[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
#pragma unroll
    for (int j = 0; j < 9; j++) madw29(*(unsigned long long*)&c[i + j], a.v[i], b.v[j]);
  }
  return xhe29::fe_reduce_cols(c);
}
template <int V>
__global__ void __launch_bounds__(256) k_mul(uint32_t* out, const uint32_t* in, int iters) {
  if (V < 2) {
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.v[i] = in[(threadIdx.x * 8 + i) & 255]; b.v[i] = in[(threadIdx.x * 8 + i + 77) & 255]; }
    for (int it = 0; it < iters; it++) a = V == 0 ? fe_mul(a, b) : fe_sq(a);
#pragma unroll
    for (int i = 0; i < 8; i++) out[(blockIdx.x * blockDim.x + threadIdx.x) * 8 + i] = a.v[i];
  } else {
    uint32_t wa[8], wb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { wa[i] = in[(threadIdx.x * 8 + i) & 255]; wb[i] = in[(threadIdx.x * 8 + i + 77) & 255]; }
    wa[7] &= 0x7fffffffu; wb[7] &= 0x7fffffffu;
    xhe29::fe a = xhe29::fe_unpack(wa), b = xhe29::fe_unpack(wb);
    for (int it = 0; it < iters; it++) a = V == 2 ? xhe29::fe_mul(a, b) : fe_mul_row29(a, b);
    xhe29::fe_pack(wa, a);
#pragma unroll
    for (int i = 0; i < 8; i++) out[(blockIdx.x * blockDim.x + threadIdx.x) * 8 + i] = wa[i];
  }
}

static double g_clock_hz;
template <typename F>
static float time_kernel(F launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  g_clock_hz = khz * 1e3;
  uint32_t h[256]; for (int i = 0; i < 256; i++) h[i] = 0x9e3779b9u * (i + 1) + 12345u;
  for (int i = 0; i < 32; i++) h[8 * i + 7] &= 0x7fffffffu;
  uint32_t *d_in, *d_out; CK(cudaMalloc(&d_in, sizeof h)); CK(cudaMalloc(&d_out, (size_t)sms * 8 * 256 * 32)); CK(cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f,\n \"raw_T_per_s\": {", prop.name, sms, g_clock_hz / 1e6);
  const int iters = 2048, blocks = sms * 8, threads = 256;
  const char* names[11] = {"acc_uniform_b", "acc_two_vector", "acc_shared_a", "prod_rz_two_vector", "prod_rz_plus_2_iadd3", "carry_chain", "iadd3_chain", "prod_rz_shared_a", "carry_chain_plus_iadd3", "acc_c_code", "dfma_f64"};
  // multiply (or add) instructions per thread per iteration
  const double per_iter[11] = {32, 32, 32, 32, 32, 32, 64, 32, 16, 32, 32};
  for (int w = 0; w < 11; w++) {
    float ms = 0;
    switch (w) {
      case 0: ms = time_kernel([&] { k_raw<0><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 1: ms = time_kernel([&] { k_raw<1><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 2: ms = time_kernel([&] { k_raw<2><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 3: ms = time_kernel([&] { k_raw<3><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 4: ms = time_kernel([&] { k_raw<4><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 5: ms = time_kernel([&] { k_raw<5><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 6: ms = time_kernel([&] { k_raw<6><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 7: ms = time_kernel([&] { k_raw<7><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 8: ms = time_kernel([&] { k_raw<8><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 9: ms = time_kernel([&] { k_raw<9><<<blocks, threads>>>(d_out, d_in, iters); }); break;
      case 10: ms = time_kernel([&] { k_raw<10><<<blocks, threads>>>(d_out, d_in, iters); }); break;
    }
    double rate = (double)blocks * threads * iters * per_iter[w] / (ms * 1e-3);
    double cyc = ms * 1e-3 * g_clock_hz / ((double)blocks * (threads / 32) / (sms * 4.0) * iters * per_iter[w]);   // cycles per warp instruction per SMSP
    printf("%s\"%s\": [%.3f, %.2f]", w ? ", " : "", names[w], rate / 1e12, cyc);
  }
  printf("},\n \"fe_mul_cycles_per_warp_op_per_smsp\": {");
  const int miters = 1000, mblocks = sms * 4;
  auto report = [&](const char* name, float ms, bool first) {
    double cyc = ms * 1e-3 * g_clock_hz / ((double)mblocks * (threads / 32) / (sms * 4.0) * miters);
    printf("%s\"%s\": %.1f", first ? "" : ", ", name, cyc);
  };
#define RUN(NAME, V, FIRST) report(NAME, time_kernel([&] { k_mul<V><<<mblocks, threads>>>(d_out, d_in, miters); }), FIRST)
  RUN("r32_mul_current", 0, true);
  RUN("r32_sq_current", 1, false);
  RUN("r29_mul_c", 2, false);
  RUN("r29_mul_rows", 3, false);
  printf("}}\n");
  // the two radix-2^29 variants must agree with each other
  uint32_t *h0 = (uint32_t*)malloc(256 * 32), *h1 = (uint32_t*)malloc(256 * 32);
  k_mul<2><<<1, 256>>>(d_out, d_in, 3); CK(cudaMemcpy(h0, d_out, 256 * 32, cudaMemcpyDeviceToHost));
  k_mul<3><<<1, 256>>>(d_out, d_in, 3); CK(cudaMemcpy(h1, d_out, 256 * 32, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 256 * 8; i++) if (h0[i] != h1[i]) { fprintf(stderr, "MISMATCH at %d\n", i); return 1; }
  fprintf(stderr, "radix-2^29 variants agree\n");
  return 0;
}
