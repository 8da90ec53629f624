// K1b — batched Ed25519 signature verification (RFC 8032, pure Ed25519) for sm_100a.
//
// BASELINE config 4 ("Ed25519 PGP keys").  NOTE: the reference itself cannot do this — its OpenPGP
// library (golang.org/x/crypto/openpgp @53104e6ec876) has no EdDSA (public-key algorithm 22) and
// skips such keys, so there is no reference behaviour to be bit-identical with.  The
// semantics here are RFC 8032 §5.1.7 as Go's crypto/ed25519 / ref10 / OpenSSL implement it:
//   reject S >= L;  decode A (reject y >= p, reject non-square, reject x = 0 with sign bit set);
//   k = SHA-512(R || A || M) mod L;  accept iff  encode([S]B - [k]A) == R  (byte compare).
// In OpenPGP (RFC 4880bis / GnuPG) M is the 32-byte v4 signature digest, so messages are a fixed 32
// bytes here.  Parity oracle for this kernel: libsodium (pynacl) and OpenSSL (`cryptography`).
//
// One thread per signature; field elements are 10 limbs in radix 2^25.5 (26/25-bit alternating,
// signed) so that every product is one 32x32->64 IMAD.WIDE and ten of them are summed without
// carries — the same lazy-carry idea as K1's radix-2^28 kernel, in the form curve25519 code has
// used since ref10.  Double-scalar multiplication is Shamir's trick over the table {B, -A, B-A}
// (253 doublings + ~190 additions, ~3.7 k field multiplications, ~0.4 M MACs per verify).
// Everything below is __host__ __device__ so the arithmetic is unit-tested on the CPU as well.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define BFTQ_HD __host__ __device__ inline
#define BFTQ_HD_NOINLINE __host__ __device__ __noinline__
#define BFTQ_ED_CONST __device__ __constant__ const
#else
#define BFTQ_HD inline
#define BFTQ_HD_NOINLINE inline
#define BFTQ_ED_CONST static const
#endif

namespace bftq { namespace ed {

typedef int32_t fe[10];

// constants (tools/gen_ed25519_consts.py)
#ifdef __CUDA_ARCH__
#define BFTQ_ED_TAB(name) d_##name
#else
#define BFTQ_ED_TAB(name) h_##name
#endif
#define BFTQ_ED_DEF(name, ...)                                      \
  static const int32_t h_##name[10] = __VA_ARGS__;                  \
  BFTQ_ED_DEVCONST(name, __VA_ARGS__)
#ifdef __CUDACC__
#define BFTQ_ED_DEVCONST(name, ...) __device__ __constant__ const int32_t d_##name[10] = __VA_ARGS__;
#else
#define BFTQ_ED_DEVCONST(name, ...)
#endif
BFTQ_ED_DEF(kD, {56195235, 13857412, 51736253, 6949390, 114729, 24766616, 60832955, 30306712, 48412415, 21499315})
BFTQ_ED_DEF(kD2, {45281625, 27714825, 36363642, 13898781, 229458, 15978800, 54557047, 27058993, 29715967, 9444199})
BFTQ_ED_DEF(kSqrtM1, {34513072, 25610706, 9377949, 3500415, 12389472, 33281959, 41962654, 31548777, 326685, 11406482})
BFTQ_ED_DEF(kBx, {52811034, 25909283, 16144682, 17082669, 27570973, 30858332, 40966398, 8378388, 20764389, 8758491})
BFTQ_ED_DEF(kBy, {40265304, 26843545, 13421772, 20132659, 26843545, 6710886, 53687091, 13421772, 40265318, 26843545})
BFTQ_ED_DEF(kBt, {28827043, 27438313, 39759291, 244362, 8635006, 11264893, 19351346, 13413597, 16611511, 27139452})
// group order L = 2^252 + 27742317777372353535851937790883648493, little-endian 32-bit words
#define BFTQ_ED_L {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0x0u, 0x0u, 0x0u, 0x10000000u}

// ---- field arithmetic mod p = 2^255 - 19 -------------------------------------------------------
BFTQ_HD void fe_copy(fe h, const fe f) { for (int i = 0; i < 10; i++) h[i] = f[i]; }
BFTQ_HD void fe_0(fe h) { for (int i = 0; i < 10; i++) h[i] = 0; }
BFTQ_HD void fe_1(fe h) { fe_0(h); h[0] = 1; }
BFTQ_HD void fe_add(fe h, const fe f, const fe g) { for (int i = 0; i < 10; i++) h[i] = f[i] + g[i]; }
BFTQ_HD void fe_sub(fe h, const fe f, const fe g) { for (int i = 0; i < 10; i++) h[i] = f[i] - g[i]; }
BFTQ_HD void fe_neg(fe h, const fe f) { for (int i = 0; i < 10; i++) h[i] = -f[i]; }

// Signed carry chain: brings every limb back to |h_even| <= 2^25, |h_odd| <= 2^24 (two passes: the
// first pass feeds 19*carry9 into limb 0).
BFTQ_HD void fe_carry(fe h, int64_t (&t)[10]) {
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const int b = (i & 1) ? 25 : 26;
      const int64_t c = (t[i] + ((int64_t)1 << (b - 1))) >> b;
      t[i] -= c * ((int64_t)1 << b);
      if (i < 9) t[i + 1] += c; else t[0] += 19 * c;
    }
  }
  // limb 0 may have picked up 19 * (tiny carry): one more step 0 -> 1 keeps the bound
  const int64_t c0 = (t[0] + ((int64_t)1 << 25)) >> 26;
  t[0] -= c0 * ((int64_t)1 << 26);
  t[1] += c0;
#pragma unroll
  for (int i = 0; i < 10; i++) h[i] = (int32_t)t[i];
}

// h = f * g.  Inputs bounded by ~2^27 per limb (a few additions of carried elements).
BFTQ_HD_NOINLINE void fe_mul(fe h, const fe f, const fe g) {
  int32_t g19[10], f2[10];
#pragma unroll
  for (int i = 0; i < 10; i++) { g19[i] = 19 * g[i]; f2[i] = 2 * f[i]; }
  int64_t t[10];
#pragma unroll
  for (int k = 0; k < 10; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {
#pragma unroll
    for (int j = 0; j < 10; j++) {
      const int32_t fi = ((i & 1) && (j & 1)) ? f2[i] : f[i];
      const int k = i + j;
      if (k < 10) t[k] += (int64_t)fi * g[j];
      else t[k - 10] += (int64_t)fi * g19[j];
    }
  }
  fe_carry(h, t);
}
BFTQ_HD void fe_sq(fe h, const fe f) { fe_mul(h, f, f); }

// Canonical 32-byte little-endian encoding.
BFTQ_HD_NOINLINE void fe_tobytes(uint8_t* s, const fe hin) {
  int64_t t[10];
  for (int i = 0; i < 10; i++) t[i] = hin[i];
  fe hh;
  fe_carry(hh, t);
  int32_t h[10];
  for (int i = 0; i < 10; i++) h[i] = hh[i];
  // q = floor((h + 19) / 2^255) in {-1?,0,1}: compute as ref10 does
  int32_t q = (19 * h[9] + (1 << 24)) >> 25;
  for (int i = 0; i < 10; i++) q = (h[i] + q) >> ((i & 1) ? 25 : 26);
  h[0] += 19 * q;
  for (int i = 0; i < 10; i++) {
    const int b = (i & 1) ? 25 : 26;
    const int32_t c = h[i] >> b;
    if (i < 9) h[i + 1] += c;
    h[i] -= c * (1 << b);
  }
  // pack 26/25-bit limbs
  uint32_t w[8];
  uint64_t acc = 0; int bits = 0, wi = 0;
  for (int i = 0; i < 10; i++) {
    acc |= (uint64_t)(uint32_t)h[i] << bits;
    bits += (i & 1) ? 25 : 26;
    while (bits >= 32 && wi < 8) { w[wi++] = (uint32_t)acc; acc >>= 32; bits -= 32; }
  }
  if (wi < 8) w[wi++] = (uint32_t)acc;
  for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)w[i]; s[4 * i + 1] = (uint8_t)(w[i] >> 8); s[4 * i + 2] = (uint8_t)(w[i] >> 16); s[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
// Loads 255 bits (the top bit is ignored).  Returns false when the value is >= p (non-canonical).
BFTQ_HD bool fe_frombytes(fe h, const uint8_t* s) {
  uint32_t w[8];
  for (int i = 0; i < 8; i++) w[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
  w[7] &= 0x7fffffffu;
  bool ge_p = w[7] == 0x7fffffffu && w[6] == 0xffffffffu && w[5] == 0xffffffffu && w[4] == 0xffffffffu && w[3] == 0xffffffffu &&
              w[2] == 0xffffffffu && w[1] == 0xffffffffu && w[0] >= 0xffffffedu;
  int pos = 0;
  for (int i = 0; i < 10; i++) {
    const int b = (i & 1) ? 25 : 26;
    const int wi = pos >> 5, sh = pos & 31;
    uint64_t v = w[wi];
    if (wi + 1 < 8) v |= (uint64_t)w[wi + 1] << 32;
    h[i] = (int32_t)((v >> sh) & (((uint64_t)1 << b) - 1));
    pos += b;
  }
  return !ge_p;
}
BFTQ_HD bool fe_isnegative(const fe f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
BFTQ_HD bool fe_iszero(const fe f) { uint8_t s[32]; fe_tobytes(s, f); uint8_t r = 0; for (int i = 0; i < 32; i++) r |= s[i]; return r == 0; }

// z^(2^252 - 3) = z^((p-5)/8)
BFTQ_HD_NOINLINE void fe_pow22523(fe out, const fe z) {
  fe t0, t1, t2;
  fe_sq(t0, z);
  fe_sq(t1, t0); fe_sq(t1, t1);
  fe_mul(t1, z, t1);
  fe_mul(t0, t0, t1);
  fe_sq(t0, t0);
  fe_mul(t0, t1, t0);                                   // z^31 = z^(2^5 - 1)
  fe_sq(t1, t0); for (int i = 1; i < 5; i++) fe_sq(t1, t1);
  fe_mul(t0, t1, t0);                                   // 2^10 - 1
  fe_sq(t1, t0); for (int i = 1; i < 10; i++) fe_sq(t1, t1);
  fe_mul(t1, t1, t0);                                   // 2^20 - 1
  fe_sq(t2, t1); for (int i = 1; i < 20; i++) fe_sq(t2, t2);
  fe_mul(t1, t2, t1);                                   // 2^40 - 1
  fe_sq(t1, t1); for (int i = 1; i < 10; i++) fe_sq(t1, t1);
  fe_mul(t0, t1, t0);                                   // 2^50 - 1
  fe_sq(t1, t0); for (int i = 1; i < 50; i++) fe_sq(t1, t1);
  fe_mul(t1, t1, t0);                                   // 2^100 - 1
  fe_sq(t2, t1); for (int i = 1; i < 100; i++) fe_sq(t2, t2);
  fe_mul(t1, t2, t1);                                   // 2^200 - 1
  fe_sq(t1, t1); for (int i = 1; i < 50; i++) fe_sq(t1, t1);
  fe_mul(t0, t1, t0);                                   // 2^250 - 1
  fe_sq(t0, t0); fe_sq(t0, t0);                         // 2^252 - 4
  fe_mul(out, t0, z);                                   // 2^252 - 3
}
// z^(p-2) = z^(2^255 - 21)
BFTQ_HD_NOINLINE void fe_invert(fe out, const fe z) {
  fe t0, t1, t2, t3;
  fe_sq(t0, z);                                         // 2
  fe_sq(t1, t0); fe_sq(t1, t1);                         // 8
  fe_mul(t1, z, t1);                                    // 9
  fe_mul(t0, t0, t1);                                   // 11
  fe_sq(t2, t0);                                        // 22
  fe_mul(t1, t1, t2);                                   // 31 = 2^5 - 1
  fe_sq(t2, t1); for (int i = 1; i < 5; i++) fe_sq(t2, t2);
  fe_mul(t1, t2, t1);                                   // 2^10 - 1
  fe_sq(t2, t1); for (int i = 1; i < 10; i++) fe_sq(t2, t2);
  fe_mul(t2, t2, t1);                                   // 2^20 - 1
  fe_sq(t3, t2); for (int i = 1; i < 20; i++) fe_sq(t3, t3);
  fe_mul(t2, t3, t2);                                   // 2^40 - 1
  fe_sq(t2, t2); for (int i = 1; i < 10; i++) fe_sq(t2, t2);
  fe_mul(t1, t2, t1);                                   // 2^50 - 1
  fe_sq(t2, t1); for (int i = 1; i < 50; i++) fe_sq(t2, t2);
  fe_mul(t2, t2, t1);                                   // 2^100 - 1
  fe_sq(t3, t2); for (int i = 1; i < 100; i++) fe_sq(t3, t3);
  fe_mul(t2, t3, t2);                                   // 2^200 - 1
  fe_sq(t2, t2); for (int i = 1; i < 50; i++) fe_sq(t2, t2);
  fe_mul(t1, t2, t1);                                   // 2^250 - 1
  fe_sq(t1, t1); for (int i = 1; i < 5; i++) fe_sq(t1, t1);   // 2^255 - 32
  fe_mul(out, t1, t0);                                  // 2^255 - 21
}

// ---- group: extended twisted Edwards coordinates (X:Y:Z:T), a = -1 ------------------------------
struct ge { fe X, Y, Z, T; };

BFTQ_HD void ge_identity(ge& p) { fe_0(p.X); fe_1(p.Y); fe_1(p.Z); fe_0(p.T); }
BFTQ_HD void ge_neg(ge& r, const ge& p) { fe_neg(r.X, p.X); fe_copy(r.Y, p.Y); fe_copy(r.Z, p.Z); fe_neg(r.T, p.T); }

// RFC 8032 §5.1.4 addition (unified).
BFTQ_HD_NOINLINE void ge_add(ge& r, const ge& p, const ge& q) {
  fe a, b, c, d, e, f, g, h, t;
  fe_sub(a, p.Y, p.X); fe_sub(t, q.Y, q.X); fe_mul(a, a, t);
  fe_add(b, p.Y, p.X); fe_add(t, q.Y, q.X); fe_mul(b, b, t);
  fe_mul(c, p.T, q.T); fe_mul(c, c, BFTQ_ED_TAB(kD2));
  fe_mul(d, p.Z, q.Z); fe_add(d, d, d);
  fe_sub(e, b, a); fe_sub(f, d, c); fe_add(g, d, c); fe_add(h, b, a);
  fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.T, e, h); fe_mul(r.Z, f, g);
}
// RFC 8032 §5.1.4 doubling.
BFTQ_HD_NOINLINE void ge_dbl(ge& r, const ge& p) {
  fe a, b, c, e, f, g, h, t;
  fe_sq(a, p.X); fe_sq(b, p.Y);
  fe_sq(c, p.Z); fe_add(c, c, c);
  fe_add(h, a, b);
  fe_add(t, p.X, p.Y); fe_sq(t, t); fe_sub(e, h, t);
  fe_sub(g, a, b);
  fe_add(f, c, g);
  // f = c + g can reach 2^27 per limb: keep it the FIRST operand (the second one is pre-multiplied by 19 in 32 bits)
  fe_mul(r.X, f, e); fe_mul(r.Y, g, h); fe_mul(r.T, e, h); fe_mul(r.Z, f, g);
}
// RFC 8032 §5.1.3 decoding.  false = not a curve point / non-canonical.
BFTQ_HD bool ge_frombytes(ge& p, const uint8_t* s) {
  const int sign = s[31] >> 7;
  if (!fe_frombytes(p.Y, s)) return false;
  fe u, v, v3, x, vxx, chk;
  fe_1(p.Z);
  fe_sq(u, p.Y);
  fe_mul(v, u, BFTQ_ED_TAB(kD));
  fe_sub(u, u, p.Z);                       // u = y^2 - 1
  fe_add(v, v, p.Z);                       // v = d y^2 + 1
  fe_sq(v3, v); fe_mul(v3, v3, v);         // v^3
  fe_sq(x, v3); fe_mul(x, x, v); fe_mul(x, x, u);   // u v^7
  fe_pow22523(x, x);
  fe_mul(x, x, v3); fe_mul(x, x, u);       // x = u v^3 (u v^7)^((p-5)/8)
  fe_sq(vxx, x); fe_mul(vxx, vxx, v);
  fe_sub(chk, vxx, u);
  if (!fe_iszero(chk)) {
    fe_add(chk, vxx, u);
    if (!fe_iszero(chk)) return false;
    fe_mul(x, x, BFTQ_ED_TAB(kSqrtM1));
  }
  if (fe_iszero(x) && sign) return false;
  if ((int)fe_isnegative(x) != sign) fe_neg(x, x);
  fe_copy(p.X, x);
  fe_mul(p.T, p.X, p.Y);
  return true;
}
BFTQ_HD void ge_tobytes(uint8_t* s, const ge& p) {
  fe zi, x, y;
  fe_invert(zi, p.Z);
  fe_mul(x, p.X, zi); fe_mul(y, p.Y, zi);
  fe_tobytes(s, y);
  s[31] ^= (uint8_t)(fe_isnegative(x) << 7);
}

// ---- scalars mod L ------------------------------------------------------------------------------
// out (8 words) = the 64-byte little-endian value `in` mod L (bitwise shift-subtract; 512 steps).
BFTQ_HD_NOINLINE void sc_reduce64(uint32_t (&out)[8], const uint8_t* in) {
  const uint32_t Lw[8] = BFTQ_ED_L;
  uint32_t r[9];
  for (int i = 0; i < 9; i++) r[i] = 0;
  for (int byte = 63; byte >= 0; byte--) {
    for (int bit = 7; bit >= 0; bit--) {
      uint32_t c = (in[byte] >> bit) & 1u;
      for (int i = 0; i < 9; i++) { const uint32_t nc = r[i] >> 31; r[i] = (r[i] << 1) | c; c = nc; }
      bool ge = r[8] != 0;
      if (!ge) { ge = true; for (int i = 7; i >= 0; i--) { if (r[i] != Lw[i]) { ge = r[i] > Lw[i]; break; } } }
      if (ge) {
        uint32_t br = 0;
        for (int i = 0; i < 8; i++) { const uint64_t d = (uint64_t)r[i] - Lw[i] - br; r[i] = (uint32_t)d; br = (uint32_t)(d >> 63); }
        r[8] -= br;
      }
    }
  }
  for (int i = 0; i < 8; i++) out[i] = r[i];
}
// s (32 bytes little-endian) < L ?
BFTQ_HD bool sc_is_canonical(const uint8_t* s, uint32_t (&w)[8]) {
  const uint32_t Lw[8] = BFTQ_ED_L;
  for (int i = 0; i < 8; i++) w[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
  for (int i = 7; i >= 0; i--) { if (w[i] != Lw[i]) return w[i] < Lw[i]; }
  return false;
}

// ---- verification given k = H(R || A || M) already reduced mod L --------------------------------
// sig = R (32) || S (32);  pk = A (32).  Returns true iff encode([S]B - [k]A) == R.
BFTQ_HD bool verify_core(const uint8_t* sig, const uint8_t* pk, const uint32_t (&k)[8]) {
  uint32_t s[8];
  if (!sc_is_canonical(sig + 32, s)) return false;                 // S >= L
  ge tab[3];                                                        // B, -A, B - A
  if (!ge_frombytes(tab[1], pk)) return false;
  ge_neg(tab[1], tab[1]);
  fe_copy(tab[0].X, BFTQ_ED_TAB(kBx)); fe_copy(tab[0].Y, BFTQ_ED_TAB(kBy)); fe_1(tab[0].Z); fe_copy(tab[0].T, BFTQ_ED_TAB(kBt));
  ge_add(tab[2], tab[0], tab[1]);
  ge p;
  ge_identity(p);
  for (int bit = 252; bit >= 0; bit--) {
    ge_dbl(p, p);
    const int idx = (int)((s[bit >> 5] >> (bit & 31)) & 1u) | (int)(((k[bit >> 5] >> (bit & 31)) & 1u) << 1);
    if (idx) { ge q; ge_add(q, p, tab[idx - 1]); p = q; }
  }
  uint8_t enc[32];
  ge_tobytes(enc, p);
  uint8_t diff = 0;
  for (int i = 0; i < 32; i++) diff |= enc[i] ^ sig[i];
  return diff == 0;
}


// ---- per-key window tables: [S]B - [k]A without a single doubling -------------------------------------------
// A batch of OpenPGP signatures is signed by a handful of keys (BASELINE configs[3]: 262 144 signatures, 15 keys),
// so the doublings of the double-scalar multiplication can be paid once per KEY instead of once per signature:
// for the base point and for every -A the table holds j * 16^i * P (i = 0..63, j = 1..8) in cached form, the
// scalars are recoded to 64 signed radix-16 digits, and a verification is at most 128 table additions
// (~1 000 field products) instead of 253 doublings + ~190 additions (~3 500).
struct gec { fe YpX, YmX, Z, T2d; };            // (Y+X, Y-X, Z, 2dT), every limb carried: safe as fe_mul's SECOND operand
constexpr int kEdWindows = 64, kEdMultiples = 8;
constexpr int kEdTableEntries = kEdWindows * kEdMultiples;

BFTQ_HD void fe_carried_add(fe h, const fe f, const fe g) { int64_t t[10]; for (int i = 0; i < 10; i++) t[i] = (int64_t)f[i] + g[i]; fe_carry(h, t); }
BFTQ_HD void fe_carried_sub(fe h, const fe f, const fe g) { int64_t t[10]; for (int i = 0; i < 10; i++) t[i] = (int64_t)f[i] - g[i]; fe_carry(h, t); }
BFTQ_HD void ge_to_cached(gec& c, const ge& p) {
  fe_carried_add(c.YpX, p.Y, p.X); fe_carried_sub(c.YmX, p.Y, p.X);
  fe_copy(c.Z, p.Z);
  fe_mul(c.T2d, p.T, BFTQ_ED_TAB(kD2));
}
// r = p + q (neg: p - q).  r may alias p.
BFTQ_HD_NOINLINE void ge_add_cached(ge& r, const ge& p, const gec& q, const bool neg) {
  fe a, b, c, d, e, f, g, h, t;
  fe_add(t, p.Y, p.X); fe_mul(a, t, neg ? q.YmX : q.YpX);
  fe_sub(t, p.Y, p.X); fe_mul(b, t, neg ? q.YpX : q.YmX);
  fe_mul(c, p.T, q.T2d);
  if (neg) fe_neg(c, c);
  fe_mul(d, p.Z, q.Z); fe_add(d, d, d);
  fe_sub(e, a, b); fe_add(h, a, b); fe_add(g, d, c); fe_sub(f, d, c);
  fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.T, e, h); fe_mul(r.Z, f, g);
}
// Window i of P's table: out[j-1] = j * 16^i * P, j = 1..8.
BFTQ_HD void ge_window_multiples(gec* out, const ge& P, const int window) {
  ge base = P;
  for (int t = 0; t < 4 * window; t++) { ge d; ge_dbl(d, base); base = d; }
  ge m = base;
  ge_to_cached(out[0], m);
  for (int j = 1; j < kEdMultiples; j++) { ge n; ge_add(n, m, base); m = n; ge_to_cached(out[j], m); }
}
BFTQ_HD void ge_basepoint(ge& b) { fe_copy(b.X, BFTQ_ED_TAB(kBx)); fe_copy(b.Y, BFTQ_ED_TAB(kBy)); fe_1(b.Z); fe_copy(b.T, BFTQ_ED_TAB(kBt)); }
// 64 signed radix-16 digits in [-8, 8] of a scalar < 2^253 given as 8 little-endian words (ref10's recoding).
BFTQ_HD void sc_signed_digits(int8_t (&e)[64], const uint32_t (&w)[8]) {
  for (int i = 0; i < 64; i++) e[i] = (int8_t)((w[i >> 3] >> (4 * (i & 7))) & 15u);
  int carry = 0;
  for (int i = 0; i < 63; i++) {
    e[i] = (int8_t)(e[i] + carry);
    carry = (e[i] + 8) >> 4;
    e[i] = (int8_t)(e[i] - (carry << 4));
  }
  e[63] = (int8_t)(e[63] + carry);
}
// encode([S]B - [k]A) == R with the two window tables (tabB for B, tabNegA for -A).  S canonical is checked here;
// the caller has checked that A decodes (its table exists).
BFTQ_HD bool verify_windowed(const uint8_t* sig, const uint32_t (&k)[8], const gec* tabB, const gec* tabNegA) {
  uint32_t s[8];
  if (!sc_is_canonical(sig + 32, s)) return false;                 // S >= L
  int8_t es[64], ek[64];
  sc_signed_digits(es, s);
  sc_signed_digits(ek, k);
  ge p;
  ge_identity(p);
  for (int i = 0; i < kEdWindows; i++) {
    const int ds = es[i], dk = ek[i];
    if (ds) { const gec q = tabB[i * kEdMultiples + (ds < 0 ? -ds : ds) - 1]; ge_add_cached(p, p, q, ds < 0); }
    if (dk) { const gec q = tabNegA[i * kEdMultiples + (dk < 0 ? -dk : dk) - 1]; ge_add_cached(p, p, q, dk < 0); }
  }
  uint8_t enc[32];
  ge_tobytes(enc, p);
  uint8_t diff = 0;
  for (int i = 0; i < 32; i++) diff |= enc[i] ^ sig[i];
  return diff == 0;
}

}}  // namespace bftq::ed

#ifdef __CUDACC__
#include "pgp_digest.cuh"
namespace bftq {
// One thread per signature.  status: 0 valid, 1 invalid, 4 key index out of range.
__global__ void __launch_bounds__(128)
ed25519_verify_kernel(const uint8_t* __restrict__ pubkeys, const uint32_t n_keys, const uint32_t* __restrict__ key_idx,
                      const uint8_t* __restrict__ sig, const uint8_t* __restrict__ msg, const uint64_t n_items,
                      uint8_t* __restrict__ status) {
  const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  const uint32_t kidx = __ldg(key_idx + item);
  if (kidx >= n_keys) { status[item] = 4; return; }
  uint8_t s[64], a[32];
  for (int i = 0; i < 64; i++) s[i] = __ldg(sig + item * 64 + i);
  for (int i = 0; i < 32; i++) a[i] = __ldg(pubkeys + (uint64_t)kidx * 32 + i);
  // k = SHA-512(R || A || M): 96 bytes = one padded block
  uint64_t w[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t r = 0, aa = 0, m = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      r = (r << 8) | s[8 * i + b];
      aa = (aa << 8) | a[8 * i + b];
      m = (m << 8) | (uint64_t)__ldg(msg + item * 32 + 8 * i + b);
    }
    w[i] = r; w[4 + i] = aa; w[8 + i] = m;
  }
  w[12] = 0x8000000000000000ull; w[13] = 0; w[14] = 0;
  w[15] = 96 * 8;
  uint64_t h[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  sha512_compress(h, w);
  uint8_t dg[64];
  for (int i = 0; i < 8; i++) for (int b = 0; b < 8; b++) dg[8 * i + b] = (uint8_t)(h[i] >> (56 - 8 * b));
  uint32_t k[8];
  ed::sc_reduce64(k, dg);
  status[item] = ed::verify_core(s, a, k) ? 0 : 1;
}

// k = SHA-512(R || A || M) mod L for one item (96 bytes = one padded block).
__device__ __forceinline__ void ed25519_hram(uint32_t (&k)[8], const uint8_t* s, const uint8_t* a, const uint8_t* m) {
  uint64_t w[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t r = 0, aa = 0, mm = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) { r = (r << 8) | s[8 * i + b]; aa = (aa << 8) | a[8 * i + b]; mm = (mm << 8) | (uint64_t)__ldg(m + 8 * i + b); }
    w[i] = r; w[4 + i] = aa; w[8 + i] = mm;
  }
  w[12] = 0x8000000000000000ull; w[13] = 0; w[14] = 0;
  w[15] = 96 * 8;
  uint64_t h[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  sha512_compress(h, w);
  uint8_t dg[64];
  for (int i = 0; i < 8; i++) for (int b = 0; b < 8; b++) dg[8 * i + b] = (uint8_t)(h[i] >> (56 - 8 * b));
  ed::sc_reduce64(k, dg);
}

// Window tables: block = one point (keys 0..n_keys-1: -A_key; block n_keys: the base point), thread = one window.
// key_ok[key] = 0 when A does not decode (RFC 8032 §5.1.3): every signature under that key is invalid.
__global__ void __launch_bounds__(64)
ed25519_table_kernel(const uint8_t* __restrict__ pubkeys, const uint32_t n_keys, ed::gec* __restrict__ tables, uint8_t* __restrict__ key_ok) {
  const uint32_t key = blockIdx.x;
  const int window = threadIdx.x;
  ed::ge P;
  if (key == n_keys) ed::ge_basepoint(P);
  else {
    uint8_t a[32];
    for (int i = 0; i < 32; i++) a[i] = __ldg(pubkeys + (uint64_t)key * 32 + i);
    const bool ok = ed::ge_frombytes(P, a);
    if (window == 0) key_ok[key] = ok ? 1 : 0;
    if (!ok) return;
    ed::ge_neg(P, P);
  }
  ed::ge_window_multiples(tables + ((size_t)key * ed::kEdWindows + window) * ed::kEdMultiples, P, window);
}

// One thread per signature against the window tables: no doublings (see ed::verify_windowed).
__global__ void __launch_bounds__(128)
ed25519_verify_windowed_kernel(const uint8_t* __restrict__ pubkeys, const uint32_t n_keys, const uint32_t* __restrict__ key_idx,
                               const uint8_t* __restrict__ sig, const uint8_t* __restrict__ msg, const uint64_t n_items,
                               const ed::gec* __restrict__ tables, const uint8_t* __restrict__ key_ok, uint8_t* __restrict__ status) {
  const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  const uint32_t kidx = __ldg(key_idx + item);
  if (kidx >= n_keys) { status[item] = 4; return; }
  if (!key_ok[kidx]) { status[item] = 1; return; }
  uint8_t s[64], a[32];
  for (int i = 0; i < 64; i++) s[i] = __ldg(sig + item * 64 + i);
  for (int i = 0; i < 32; i++) a[i] = __ldg(pubkeys + (uint64_t)kidx * 32 + i);
  uint32_t k[8];
  ed25519_hram(k, s, a, msg + item * 32);
  status[item] = ed::verify_windowed(s, k, tables + (size_t)n_keys * ed::kEdTableEntries, tables + (size_t)kidx * ed::kEdTableEntries) ? 0 : 1;
}
}  // namespace bftq
#endif
