#include <cstdint>
// does ptxas fuse mad.lo.cc / madc.hi.cc pairs into IMAD.WIDE.U32(.X)?
__global__ void k(uint32_t* out, const uint32_t* in) {
  uint32_t a[8], acc[10];
  uint32_t bi = in[100 + threadIdx.x];
  for (int i = 0; i < 8; i++) a[i] = in[i * 32 + threadIdx.x];
  for (int i = 0; i < 10; i++) acc[i] = in[i * 32 + threadIdx.x + 1000];
  // even chain
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(bi));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[2]), "+r"(acc[3]) : "r"(a[2]), "r"(bi));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[4]), "+r"(acc[5]) : "r"(a[4]), "r"(bi));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[6]), "+r"(acc[7]) : "r"(a[6]), "r"(bi));
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[8]));
  for (int i = 0; i < 10; i++) out[i * 32 + threadIdx.x] = acc[i];
}
