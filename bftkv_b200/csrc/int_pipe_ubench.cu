// Integer-pipe microbenchmark v2 (single wave, occupancy-aware, warmed clocks)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define ITERS 2048
#define CHK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);}}while(0)

// carry chain of 4 wide MACs: acc[0..7] += a[0,2,4,6]*b (pairs), carry out dropped into acc[8]
#define CHAIN4(acc, a, b) \
  asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;" \
      "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;" \
      "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;" \
      "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7; addc.u32 %8, %8, 0;" \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]) \
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b))

template <int NCH>
__global__ void __launch_bounds__(256) k_chain(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t acc[NCH][9], a[4]; uint32_t b = (seed | 1) + 2 * threadIdx.x;
  for (int i = 0; i < 4; i++) a[i] = (seed ^ 0x9e3779b9u) * (i + 1) + threadIdx.x;
  for (int c = 0; c < NCH; c++) for (int i = 0; i < 9; i++) acc[c][i] = threadIdx.x + i * seed + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) CHAIN4(acc[c], a, b);
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 9; i++) s ^= acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// independent IMAD.WIDE (no carry), NCH*4 accumulators, distinct a regs
template <int NCH>
__global__ void __launch_bounds__(256) k_wide(uint32_t* out, uint32_t seed, long long* cyc) {
  unsigned long long acc[NCH][4]; uint32_t a[4]; uint32_t b = (seed | 1) + 2 * threadIdx.x;
  for (int i = 0; i < 4; i++) a[i] = (seed ^ 0x9e3779b9u) * (i + 1) + threadIdx.x;
  for (int c = 0; c < NCH; c++) for (int i = 0; i < 4; i++) acc[c][i] = threadIdx.x + i * seed + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
      for (int i = 0; i < 4; i++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c][i]) : "r"(a[i]), "r"(b));
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 4; i++) s ^= acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// IMAD.LO with 3 distinct regs
template <int NCH>
__global__ void __launch_bounds__(256) k_lo(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t acc[NCH][4]; uint32_t a[4]; uint32_t b = (seed | 1) + 2 * threadIdx.x;
  for (int i = 0; i < 4; i++) a[i] = (seed ^ 0x9e3779b9u) * (i + 1) + threadIdx.x;
  for (int c = 0; c < NCH; c++) for (int i = 0; i < 4; i++) acc[c][i] = threadIdx.x + i * seed + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
      for (int i = 0; i < 4; i++) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c][i]) : "r"(a[i]), "r"(b));
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 4; i++) s ^= acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// wide + carry handled on ALU pipe: (hi,lo) = a*b + acc ; then add.cc chain on separate IADD3s
// pattern: p64 = a*b (mul.wide) ; acc_j += lo(p) (add.cc) ; acc_j+1 += hi(p) + c (addc.cc)  -> 1 IMAD.WIDE + 2 IADD3 per MAC
template <int NCH>
__global__ void __launch_bounds__(256) k_wide_add(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t acc[NCH][9], a[4]; uint32_t b = (seed | 1) + 2 * threadIdx.x;
  for (int i = 0; i < 4; i++) a[i] = (seed ^ 0x9e3779b9u) * (i + 1) + threadIdx.x;
  for (int c = 0; c < NCH; c++) for (int i = 0; i < 9; i++) acc[c][i] = threadIdx.x + i * seed + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      unsigned long long p0, p1, p2, p3;
      asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[0]), "r"(b));
      asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(a[1]), "r"(b));
      asm("mul.wide.u32 %0, %1, %2;" : "=l"(p2) : "r"(a[2]), "r"(b));
      asm("mul.wide.u32 %0, %1, %2;" : "=l"(p3) : "r"(a[3]), "r"(b));
      asm("add.cc.u32 %0, %0, %9; addc.cc.u32 %1, %1, %10; addc.cc.u32 %2, %2, %11; addc.cc.u32 %3, %3, %12;"
          "addc.cc.u32 %4, %4, %13; addc.cc.u32 %5, %5, %14; addc.cc.u32 %6, %6, %15; addc.cc.u32 %7, %7, %16; addc.u32 %8, %8, 0;"
          : "+r"(acc[c][0]), "+r"(acc[c][1]), "+r"(acc[c][2]), "+r"(acc[c][3]), "+r"(acc[c][4]), "+r"(acc[c][5]), "+r"(acc[c][6]), "+r"(acc[c][7]), "+r"(acc[c][8])
          : "r"((uint32_t)p0), "r"((uint32_t)(p0 >> 32)), "r"((uint32_t)p1), "r"((uint32_t)(p1 >> 32)), "r"((uint32_t)p2), "r"((uint32_t)(p2 >> 32)), "r"((uint32_t)p3), "r"((uint32_t)(p3 >> 32)));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 9; i++) s ^= acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// chain + shfl + lop mix approximating the real loop: per 8 wide-X MACs: 1 shfl, 2 iadd3, 1 imad.lo
template <int NCH>
__global__ void __launch_bounds__(256) k_real(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t acc[NCH][9], a[4]; uint32_t b = (seed | 1) + 2 * threadIdx.x;
  for (int i = 0; i < 4; i++) a[i] = (seed ^ 0x9e3779b9u) * (i + 1) + threadIdx.x;
  for (int c = 0; c < NCH; c++) for (int i = 0; i < 9; i++) acc[c][i] = threadIdx.x + i * seed + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      CHAIN4(acc[c], a, b);
      b = __shfl_sync(0xffffffffu, acc[c][0] * seed, (threadIdx.x & 24));
      CHAIN4(acc[c], a, b);
      acc[c][8] += acc[c][1] + b;
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 9; i++) s ^= acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(uint32_t*, uint32_t, long long*);
struct K { const char* name; kern_t f; double macs_per_iter; };
static uint32_t* out; static long long* cyc; static long long* hc; static int sms;

static void run(const K& k, int maxblk, int first, int last) {
  cudaFuncAttributes fa; CHK(cudaFuncGetAttributes(&fa, k.f));
  int occ = 0; CHK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.f, 256, 0));
  for (int bps = 1; bps <= occ; bps *= 2) {
    if (maxblk && bps > maxblk) break;
    int blocks = sms * bps;
    for (int w = 0; w < 2; w++) k.f<<<blocks, 256>>>(out, 12345u + w, cyc);
    CHK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int reps = 5;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) k.f<<<blocks, 256>>>(out, 777u + r, cyc);
    cudaEventRecord(e1); CHK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    CHK(cudaMemcpy(hc, cyc, blocks * 8, cudaMemcpyDeviceToHost));
    double avgc = 0; for (int b = 0; b < blocks; b++) avgc += hc[b]; avgc /= blocks;
    double macs = (double)blocks * 256 * ITERS * k.macs_per_iter;
    printf("  {\"kernel\": \"%s\", \"regs\": %d, \"warps_per_sm\": %d, \"Tmac_per_s\": %.3f, \"mac_per_clk_per_sm\": %.2f, \"ms\": %.4f, \"sm_mhz\": %.0f},\n",
           k.name, fa.numRegs, bps * 8, macs * reps / (ms * 1e-3) / 1e12, 256.0 * bps * ITERS * k.macs_per_iter / avgc, ms / reps, avgc / (ms / reps * 1e-3) / 1e6);
  }
}

int main() {
  cudaDeviceProp p; CHK(cudaGetDeviceProperties(&p, 0)); sms = p.multiProcessorCount;
  CHK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4)); CHK(cudaMalloc(&cyc, sms * 8 * 8)); hc = (long long*)malloc(sms * 8 * 8);
  // warm clocks ~1s
  { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms = 0; cudaEventRecord(e0);
    while (ms < 1000) { for (int i = 0; i < 20; i++) k_wide<2><<<sms * 8, 256>>>(out, 1, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); } }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
  K ks[] = {
    {"imad_lo_x8", k_lo<2>, 8}, {"imad_lo_x16", k_lo<4>, 16},
    {"imad_wide_x8", k_wide<2>, 8}, {"imad_wide_x16", k_wide<4>, 16},
    {"wideX_chain4_x1", k_chain<1>, 4}, {"wideX_chain4_x2", k_chain<2>, 8}, {"wideX_chain4_x4", k_chain<4>, 16},
    {"mulwide+iadd3_x1", k_wide_add<1>, 4}, {"mulwide+iadd3_x2", k_wide_add<2>, 8}, {"mulwide+iadd3_x4", k_wide_add<4>, 16},
    {"real_mix_x1", k_real<1>, 8}, {"real_mix_x2", k_real<2>, 16},
  };
  for (auto& k : ks) run(k, 0, 0, 0);
  printf("  {}\n]}\n");
  return 0;
}
