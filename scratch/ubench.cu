// Integer-pipe microbenchmark for B200: establishes the roofline denominator for the RSA kernel.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CHK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

// 0: mad.lo.u32 (IMAD), 8 independent chains
__global__ void k_imad_lo(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[8]; uint32_t b = seed | 1, c = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 1: mad.wide.u32 (IMAD.WIDE), 8 independent chains (64-bit accumulators)
__global__ void k_imad_wide(uint32_t* out, uint32_t seed, long long* cyc) {
  unsigned long long x[8]; uint32_t b = seed | 1, a = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 2: carry-chained IMAD.WIDE.U32.X: 2 chains of 4 (lo.cc/hi.cc pairs), as the Montgomery row uses
__global__ void k_imad_wide_x(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t e[9], o[9]; uint32_t b = seed | 1, a = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 9; i++) { e[i] = threadIdx.x + i * seed; o[i] = threadIdx.x * 3 + i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS / 2; it++) {
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
    asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                 "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                 "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                 "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.cc.u32 %7, %8, %9, %7;"
                 : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]) : "r"(a), "r"(b));
    asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                 "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                 "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                 "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.cc.u32 %7, %8, %9, %7;"
                 : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]) : "r"(a), "r"(b));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 9; i++) s ^= e[i] ^ o[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 3: mad.hi.u32
__global__ void k_imad_hi(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[8]; uint32_t b = seed | 0x80000001u, c = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * seed + 0x80000000u;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 4: IADD3 (add.u32 with 3 inputs -> lop/iadd3)
__global__ void k_iadd3(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[8]; uint32_t b = seed | 1, c = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[i]) : "r"(b), "r"(c));
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 5: shfl.sync.idx throughput
__global__ void k_shfl(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * seed;
  int src = (threadIdx.x & 28);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __shfl_sync(0xffffffffu, x[i], src + (i & 3));
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 6: mixed: 8 IMAD.WIDE + 8 IADD3 per iteration (dual-issue check)
__global__ void k_mix_wide_iadd(uint32_t* out, uint32_t seed, long long* cyc) {
  unsigned long long x[8]; uint32_t y[8]; uint32_t b = seed | 1, a = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i * seed; y[i] = threadIdx.x ^ i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
      asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(y[i]) : "r"(b), "r"(a));
    }
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int i = 0; i < 8; i++) s ^= x[i] ^ y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 7: mixed: 8 IMAD.WIDE + 2 SHFL per iteration
__global__ void k_mix_wide_shfl(uint32_t* out, uint32_t seed, long long* cyc) {
  unsigned long long x[8]; uint32_t y[2]; uint32_t b = seed | 1, a = seed ^ 0x9e3779b9u;
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i * seed; }
  y[0] = threadIdx.x; y[1] = seed;
  int src = (threadIdx.x & 28);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
    y[0] = __shfl_sync(0xffffffffu, y[0], src + 1);
    y[1] = __shfl_sync(0xffffffffu, y[1], src + 2);
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int i = 0; i < 8; i++) s ^= x[i]; s ^= y[0] ^ y[1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 8: 64-bit multiply mul.lo.u64 chains (to see how IMAD 64 lowers)
typedef void (*kern_t)(uint32_t*, uint32_t, long long*);
struct K { const char* name; kern_t f; double ops_per_iter; };

int main() {
  int dev = 0; cudaDeviceProp p; CHK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  K ks[] = {{"imad_lo", k_imad_lo, 8}, {"imad_wide", k_imad_wide, 8}, {"imad_wide_x_chain", k_imad_wide_x, 8},
            {"imad_hi", k_imad_hi, 8}, {"iadd3x2", k_iadd3, 16}, {"shfl_idx", k_shfl, 8},
            {"mix_wide(8)+iadd(16)", k_mix_wide_iadd, 8}, {"mix_wide(8)+shfl(2)", k_mix_wide_shfl, 8}};
  uint32_t* out; long long* cyc;
  int tpb = 1024, bps = 2; int blocks = sms * bps;
  CHK(cudaMalloc(&out, (size_t)blocks * tpb * 4)); CHK(cudaMalloc(&cyc, blocks * 8));
  long long* hc = (long long*)malloc(blocks * 8);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_prop\": %d, \"results\": [\n", p.name, sms, p.clockRate);
  int nk = sizeof(ks) / sizeof(ks[0]);
  for (int tp = 0; tp < 2; tp++) {
    tpb = tp == 0 ? 1024 : 256; bps = tp == 0 ? 2 : 8; blocks = sms * bps;
  for (int i = 0; i < nk; i++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) ks[i].f<<<blocks, tpb>>>(out, 12345u + w, cyc);
    CHK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    int reps = 10;
    for (int r = 0; r < reps; r++) ks[i].f<<<blocks, tpb>>>(out, 777u + r, cyc);
    cudaEventRecord(e1); CHK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    CHK(cudaMemcpy(hc, cyc, blocks * 8, cudaMemcpyDeviceToHost));
    double avgc = 0; for (int b = 0; b < blocks; b++) avgc += hc[b]; avgc /= blocks;
    double lane_ops = (double)blocks * tpb * ITERS * ks[i].ops_per_iter;  // per launch
    double tops = lane_ops * reps / (ms * 1e-3) / 1e12;
    // per-SM per-clk: threads resident per SM = tpb*bps, all concurrently
    double per_sm_clk = (double)tpb * bps * ITERS * ks[i].ops_per_iter / avgc;
    printf("  {\"kernel\": \"%s\", \"tpb\": %d, \"blocks_per_sm\": %d, \"T_lane_ops_per_s\": %.3f, \"lane_ops_per_clk_per_sm\": %.2f, \"ms_per_launch\": %.4f, \"eff_mhz\": %.0f}%s\n",
           ks[i].name, tpb, bps, tops, per_sm_clk, ms / reps, avgc / (ms / reps * 1e-3) / 1e6, (i == nk - 1 && tp == 1) ? "" : ",");
  }}
  printf("]}\n");
  return 0;
}
