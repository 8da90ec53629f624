// Host build of the Ed25519 arithmetic of bftkv_b200/csrc/ed25519.cuh (the same __host__ __device__
// code the kernel runs), so field / group / scalar code is unit-tested on the CPU.  Product code
// compiled for the host; not an oracle.
#include "../../bftkv_b200/csrc/ed25519.cuh"
#include <cstring>
using namespace bftq::ed;
extern "C" {
int ed_verify_core_host(const uint8_t* sig, const uint8_t* pk, const uint8_t* k64) {
  uint32_t k[8];
  sc_reduce64(k, k64);
  return verify_core(sig, pk, k) ? 1 : 0;
}
void ed_fe_mul_host(const uint8_t* a, const uint8_t* b, uint8_t* out) { fe x, y, z; fe_frombytes(x, a); fe_frombytes(y, b); fe_mul(z, x, y); fe_tobytes(out, z); }
void ed_fe_invert_host(const uint8_t* a, uint8_t* out) { fe x, z; fe_frombytes(x, a); fe_invert(z, x); fe_tobytes(out, z); }
void ed_fe_addsubmul_host(const uint8_t* a, const uint8_t* b, uint8_t* out) {   // (a+b)*(a-b) exercises uncarried inputs
  fe x, y, s, d, z; fe_frombytes(x, a); fe_frombytes(y, b); fe_add(s, x, y); fe_sub(d, x, y); fe_mul(z, s, d); fe_tobytes(out, z);
}
int ed_point_roundtrip_host(const uint8_t* in, uint8_t* out) { ge p; if (!ge_frombytes(p, in)) return 0; ge q; ge_dbl(q, p); ge r; ge_add(r, q, p); ge_tobytes(out, r); return 1; }
void ed_sc_reduce_host(const uint8_t* in64, uint8_t* out32) { uint32_t k[8]; sc_reduce64(k, in64); memcpy(out32, k, 32); }
// Windowed verification (per-key tables built on the host by the same code the table kernel runs).
int ed_verify_windowed_host(const uint8_t* sig, const uint8_t* pk, const uint8_t* k64) {
  static gec tabB[kEdTableEntries];
  static bool haveB = false;
  if (!haveB) { ge B; ge_basepoint(B); for (int w = 0; w < kEdWindows; w++) ge_window_multiples(tabB + w * kEdMultiples, B, w); haveB = true; }
  static gec tabA[kEdTableEntries];
  static uint8_t last_pk[32];
  static int last_ok = -1;
  if (last_ok < 0 || memcmp(last_pk, pk, 32) != 0) {
    ge A;
    last_ok = ge_frombytes(A, pk) ? 1 : 0;
    memcpy(last_pk, pk, 32);
    if (last_ok) { ge nA; ge_neg(nA, A); for (int w = 0; w < kEdWindows; w++) ge_window_multiples(tabA + w * kEdMultiples, nA, w); }
  }
  if (!last_ok) return 0;
  uint32_t k[8];
  sc_reduce64(k, k64);
  return verify_windowed(sig, k, tabB, tabA) ? 1 : 0;
}
void ed_signed_digits_host(const uint8_t* s32, int8_t* out64) {
  uint32_t w[8]; memcpy(w, s32, 32);
  int8_t e[64]; sc_signed_digits(e, w); memcpy(out64, e, 64);
}
}
