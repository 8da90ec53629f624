#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// DFMA peak and the "Emmart" 52-bit limb product inner loop (2 DFMA + 2 int64 adds per limb product)
__global__ void __launch_bounds__(256) dfma_peak(double* out, double a0, double b0, int iters) {
  double acc[16];
  double a = a0 + threadIdx.x, b = b0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = (double)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = __fma_rz(a, b, acc[i]);
    a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// per limb product: hi = fma_rz(a,b,C1); lo = fma_rz(a,b,C2-hi); acc_hi += bits(hi); acc_lo += bits(lo)
__global__ void __launch_bounds__(256) limb_loop(unsigned long long* out, double a0, double b0, int iters) {
  const double C1 = 20282409603651670423947251286016.0;           // 2^104
  const double C2 = 20282409603651674927546878656512.0;           // 2^104 + 2^52
  long long acc[16];
  double a[8];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = a0 + threadIdx.x + i;
  double b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const double hi = __fma_rz(a[i], b, C1);
      const double sub = C2 - hi;
      const double lo = __fma_rz(a[i], b, sub);
      acc[i + 1] += __double_as_longlong(hi);
      acc[i] += __double_as_longlong(lo);
    }
    b = __shfl_sync(0xffffffffu, b, (threadIdx.x + 1) & 31) + 1.0;
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += (unsigned long long)acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sm * 8, block = 256, iters = 1 << 14;
  double* d; cudaMalloc(&d, (size_t)grid * block * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); dfma_peak<<<grid, block>>>(d, 1.5, 1.0000001, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("dfma_peak: %.3f ms  %.2f T DFMA/s\n", ms, (double)grid * block * iters * 16 / ms / 1e9);
  }
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); limb_loop<<<grid, block>>>((unsigned long long*)d, 4503599627370495.0 - 1000, 4503599627370001.0, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("limb_loop: %.3f ms  %.2f T limb-products(52x52)/s  = %.2f T 32x32-equivalents/s\n", ms, (double)grid * block * iters * 8 / ms / 1e9,
           (double)grid * block * iters * 8 / ms / 1e9 * (52.0 * 52.0) / 1024.0);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
