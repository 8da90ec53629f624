// Host build of the Ed25519 arithmetic of bftkv_b200/csrc/ed25519.cuh + ed25519_fast.cuh (the same __host__ __device__
// code the kernel runs), so field / group / scalar code is unit-tested on the CPU.  Product code
// compiled for the host; not an oracle.
#include "../../bftkv_b200/csrc/ed25519_fast.cuh"
#include <vector>
#include <cstring>
using namespace bftq::ed;
extern "C" {
int ed_verify_core_host(const uint8_t* sig, const uint8_t* pk, const uint8_t* k64) {
  uint32_t k[8];
  sc_reduce64(k, k64);
  return verify_core(sig, pk, k) ? 1 : 0;
}
int ed_verify_core_fast_host(const uint8_t* sig, const uint8_t* pk, const uint8_t* k64) {
  uint32_t k[8];
  sc_reduce64(k, k64);
  return verify_core_fast(sig, pk, k) ? 1 : 0;
}
void ed_fe_mul_host(const uint8_t* a, const uint8_t* b, uint8_t* out) { fe x, y, z; fe_frombytes(x, a); fe_frombytes(y, b); fe_mul(z, x, y); fe_tobytes(out, z); }
void ed_fe_invert_host(const uint8_t* a, uint8_t* out) { fe x, z; fe_frombytes(x, a); fe_invert(z, x); fe_tobytes(out, z); }
void ed_fe_addsubmul_host(const uint8_t* a, const uint8_t* b, uint8_t* out) {   // (a+b)*(a-b) exercises uncarried inputs
  fe x, y, s, d, z; fe_frombytes(x, a); fe_frombytes(y, b); fe_add(s, x, y); fe_sub(d, x, y); fe_mul(z, s, d); fe_tobytes(out, z);
}
int ed_point_roundtrip_host(const uint8_t* in, uint8_t* out) { ge p; if (!ge_frombytes(p, in)) return 0; ge q; ge_dbl(q, p); ge r; ge_add(r, q, p); ge_tobytes(out, r); return 1; }
void ed_sc_reduce_host(const uint8_t* in64, uint8_t* out32) { uint32_t k[8]; sc_reduce64(k, in64); memcpy(out32, k, 32); }
// ---- fast path (ed25519_fast.cuh): the same functions the kernels run, thread loops emulated ------------------------------
void ed_fex_mul_host(const uint8_t* a, const uint8_t* b, uint8_t* out) { fe x, y, z; fe_frombytes(x, a); fe_frombytes(y, b); fex_mul(z, x, y); fe_tobytes(out, z); }
void ed_fex_sq_host(const uint8_t* a, uint8_t* out) { fe x, z; fe_frombytes(x, a); fex_sq(z, x); fe_tobytes(out, z); }
// (a+b)^2 and (a+b)*(a-b): uncarried operands at the documented bounds; out = 64 bytes
void ed_fex_uncarried_host(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  fe x0, y0, x, y, s, d, z, one = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  fe_frombytes(x0, a); fe_frombytes(y0, b); fex_mul(x, x0, one); fex_mul(y, y0, one);       // carried representatives
  fex_add(s, x, y); fex_sub(d, x, y);
  fex_sq(z, s); fe_tobytes(out, z); fex_mul(z, s, d); fe_tobytes(out + 32, z);
}
void ed_fex_pow_host(const uint8_t* a, int inverse, uint8_t* out) { fe x, z; fe_frombytes(x, a); fex_pow_chain(z, x, inverse != 0); fe_tobytes(out, z); }
void ed_fex_towords_host(const uint8_t* a, const uint8_t* b, uint8_t* out) {      // canonical words of a*b
  fe x, y, z; fe_frombytes(x, a); fe_frombytes(y, b); fex_mul(z, x, y); uint32_t w[8]; fex_towords(w, z); memcpy(out, w, 32);
}
void ed_sc_reduce512_host(const uint8_t* in64, uint8_t* out32) { uint32_t x[16], k[8]; memcpy(x, in64, 64); sc_reduce512(k, x); memcpy(out32, k, 32); }
// signed radix-2^W digits of a scalar (W = 10: 26 digits, W = 12: 22 digits)
void ed_digits_host(const uint8_t* s32, int wbits, int32_t* out) {
  uint32_t w[kFxScalarWords] = {0}; memcpy(w, s32, 32); uint32_t c = 0;
  if (wbits == 10) for (int i = 0; i < FxA::windows; i++) out[i] = sc_digit<10>(w, 1, i, c);
  else for (int i = 0; i < FxB::windows; i++) out[i] = sc_digit<12>(w, 1, i, c);
}
int ed_gex_roundtrip_host(const uint8_t* in, uint8_t* out) {                      // encode(3 P) via gex_dbl / gex_add / pow chain
  gex p; if (!gex_frombytes(p, in)) return 0;
  gex q; gex_dbl(q, p); gex r; gex_add(r, q, p);
  fe zi; fex_pow_chain(zi, r.Z, true);
  uint32_t zero[8] = {0}; (void)zero;
  fe x, y; fex_mul(x, r.X, zi); fex_mul(y, r.Y, zi);
  uint32_t wx[8], wy[8]; fex_towords(wx, x); fex_towords(wy, y); wy[7] ^= (wx[0] & 1u) << 31; memcpy(out, wy, 32);
  return 1;
}
// Table of one point exactly as the two table kernels build it (bases, then chunks of eight multiples).
static void fx_build_table(std::vector<gea>& tab, const gex& P, int nw, int wbits) {
  const int multiples = 1 << (wbits - 1);
  tab.resize((size_t)nw * multiples);
  std::vector<gex> bases(nw);
  fx_window_bases(bases.data(), P, nw, wbits);
  for (int w = 0; w < nw; w++)
    for (int c = 0; c < multiples / kFxChunk; c++) fx_window_chunk(tab.data() + (size_t)w * multiples + c * kFxChunk, bases[w], c);
}
// table entry j of window w of a key's radix-2^10 table (base = 1: the base point's radix-2^12 table) as the affine
// point's encoding (checked against big-integer arithmetic by the test)
int ed_fx_table_entry_host(const uint8_t* pk, int base, int window, int j, uint8_t* out_ypx, uint8_t* out_ymx, uint8_t* out_xy2d) {
  static std::vector<gea> tab, tabB; static uint8_t last[32]; static bool have = false;
  if (base) { if (tabB.empty()) { gex B; gex_basepoint(B); fx_build_table(tabB, B, FxB::windows, kFxWB); } }
  else if (!have || memcmp(last, pk, 32) != 0) { gex A; if (!gex_frombytes(A, pk)) return 0; fx_build_table(tab, A, FxA::windows, kFxWA); memcpy(last, pk, 32); have = true; }
  const gea& e = base ? tabB[(size_t)window * FxB::multiples + j - 1] : tab[(size_t)window * FxA::multiples + j - 1];
  fe_tobytes(out_ypx, e.ypx); fe_tobytes(out_ymx, e.ymx); fe_tobytes(out_xy2d, e.xy2d);
  for (int i = 0; i < 10; i++) {                    // entries must be carried: safe as fex_mul's second operand
    const int lim = (i & 1) ? (1 << 24) + (1 << 18) : (1 << 25) + (1 << 19);
    if (e.ypx[i] > lim || e.ypx[i] < -lim || e.ymx[i] > lim || e.ymx[i] < -lim || e.xy2d[i] > lim || e.xy2d[i] < -lim) return -1;
  }
  return 1;
}
// The fast verification: accumulate (kernel 1) + finish with a simultaneous inversion (kernel 2, here over one result).
int ed_verify_fast_host(const uint8_t* sig, const uint8_t* pk, const uint8_t* k64) {
  static std::vector<gea> tabB, tabA;
  static uint8_t last_pk[32];
  static int last_ok = -1;
  if (tabB.empty()) { gex B; gex_basepoint(B); fx_build_table(tabB, B, FxB::windows, kFxWB); }
  if (last_ok < 0 || memcmp(last_pk, pk, 32) != 0) {
    gex A;
    last_ok = gex_frombytes(A, pk) ? 1 : 0;
    memcpy(last_pk, pk, 32);
    if (last_ok) { for (int i = 0; i < 10; i++) { A.X[i] = -A.X[i]; A.T[i] = -A.T[i]; } fx_build_table(tabA, A, FxA::windows, kFxWA); }
  }
  if (!last_ok) return 0;
  uint32_t x[16], k[8], s[8], r[8], s9[kFxScalarWords] = {0}, k9[kFxScalarWords] = {0};
  memcpy(x, k64, 64); sc_reduce512(k, x);
  memcpy(r, sig, 32); memcpy(s, sig + 32, 32);
  if (!sc_words_canonical(s)) return 0;
  memcpy(s9, s, 32); memcpy(k9, k, 32);
  gex p;
  fx_accumulate(p, s9, k9, 1, tabB.data(), tabA.data());
  fe zi; fex_pow_chain(zi, p.Z, true);
  return fx_encodes_to(p.X, p.Y, zi, r) ? 1 : 0;
}
// max |limb| seen on the accumulator after a verification's additions (bounds check of the lazy-carry discipline)
int ed_fast_limb_bound_host(const uint8_t* sig, const uint8_t* pk, const uint8_t* k64) {
  static std::vector<gea> tabB, tabA;
  if (tabB.empty()) { gex B; gex_basepoint(B); fx_build_table(tabB, B, FxB::windows, kFxWB); }
  gex A; if (!gex_frombytes(A, pk)) return -1;
  for (int i = 0; i < 10; i++) { A.X[i] = -A.X[i]; A.T[i] = -A.T[i]; }
  fx_build_table(tabA, A, FxA::windows, kFxWA);
  uint32_t x[16], k[8], s9[kFxScalarWords] = {0}, k9[kFxScalarWords] = {0};
  memcpy(x, k64, 64); sc_reduce512(k, x); memcpy(s9, sig + 32, 32); memcpy(k9, k, 32);
  gex p; fx_accumulate(p, s9, k9, 1, tabB.data(), tabA.data());
  int m = 0;
  for (int i = 0; i < 10; i++) for (const int32_t* c : {p.X, p.Y, p.Z, p.T}) { const int v = c[i] < 0 ? -c[i] : c[i]; if (v > m) m = v; }
  return m;
}
}
