// K1 (fast path) — RSA-2048 PKCS#1 v1.5 batch verify, radix 2^32 with IMAD.WIDE.U32.X carry chains.
//
// Same contract as rsa_verify.cuh (the radix-2^28 kernel, kept for moduli of 2041..2047 bits); this
// variant requires the modulus to have exactly 2048 bits, which is what every RSA-2048 key
// generator (gpg, OpenSSL, Go) produces.
//
// Why: on B200 every 64-bit-result integer multiply (IMAD.WIDE, IMAD.WIDE.X, IMAD.HI) occupies the
// FMA-heavy pipe for 4 cycles, carry chain or not (profiles/).  The carry-free
// radix-2^28 form therefore buys nothing per instruction and pays (76/64)^2 = 1.41x more of them.
// Here a signature is owned by 4 lanes x 16 32-bit limbs; mad.lo.cc/madc.hi.cc pairs compile to one
// IMAD.WIDE.U32.X each, and the classic even/odd column split keeps every chain's carry inside the
// instruction stream:
//   E[k] sits at limb position k     (even-aligned pairs (0,1),(2,3),...)
//   O[k] sits at limb position k+1   (odd-aligned pairs (1,2),(3,4),...)
// One round consumes TWO limbs of b (offsets 0 and 1): at offset 0 even limbs of the operand feed E
// and odd ones feed O, at offset 1 it is the other way round, so nothing ever has to be added across
// the two alignments inside a round.  After the round the number is shifted down two limbs, which is
// pure register renaming for E and O; the only cross terms are the 1-bit carry out of position 1
// (fed into the next round's first chain as its carry-in), the upper half of the one O pair the cut
// goes through (kept as a lone pending limb Z at the new position 0) and the two low limbs each lane
// hands to the lane below.  The scheme is checked limb-for-limb by tools/emu_r32.py.  Montgomery products are "almost" reduced (< R = 2^2048): one conditional
// subtraction of n when the result overflowed 2^2048, decided by the carry out of the top lane.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rsa_verify.cuh"

namespace bftq {
namespace r32 {

constexpr int T = 4;      // lanes per number
// W = 32-bit limbs per lane: 16 for 2048-bit numbers (the RSA kernel), 8 for 1024-bit ones (modexp).

// ---- carry-chain building blocks ----------------------------------------------------------------
// One asm statement per chain (8 mad.lo.cc/madc.hi.cc pairs + the carry limb), so ptxas sees the
// whole chain at once and keeps the carry in a predicate: each pair becomes one IMAD.WIDE.U32.X.
// P(lo,hi) are accumulator limbs, X the operand limbs, m the multiplier.
#define BFTQ_CHAIN8_BODY(first)                                                              \
  first " %0, %18, %26, %0;  madc.hi.cc.u32 %1, %18, %26, %1;"                                \
  "madc.lo.cc.u32 %2, %19, %26, %2;  madc.hi.cc.u32 %3, %19, %26, %3;"                        \
  "madc.lo.cc.u32 %4, %20, %26, %4;  madc.hi.cc.u32 %5, %20, %26, %5;"                        \
  "madc.lo.cc.u32 %6, %21, %26, %6;  madc.hi.cc.u32 %7, %21, %26, %7;"                        \
  "madc.lo.cc.u32 %8, %22, %26, %8;  madc.hi.cc.u32 %9, %22, %26, %9;"                        \
  "madc.lo.cc.u32 %10, %23, %26, %10; madc.hi.cc.u32 %11, %23, %26, %11;"                     \
  "madc.lo.cc.u32 %12, %24, %26, %12; madc.hi.cc.u32 %13, %24, %26, %13;"                     \
  "madc.lo.cc.u32 %14, %25, %26, %14; madc.hi.cc.u32 %15, %25, %26, %15;"                     \
  "addc.cc.u32 %16, %16, 0; addc.u32 %17, %17, 0;"
#define BFTQ_CHAIN8_OPS(p, c0, c1, x)                                                                                     \
  : "+r"(p[0]), "+r"(p[1]), "+r"(p[2]), "+r"(p[3]), "+r"(p[4]), "+r"(p[5]), "+r"(p[6]), "+r"(p[7]), "+r"(p[8]), "+r"(p[9]), \
    "+r"(p[10]), "+r"(p[11]), "+r"(p[12]), "+r"(p[13]), "+r"(p[14]), "+r"(p[15]), "+r"(c0), "+r"(c1)                      \
  : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(x[8]), "r"(x[10]), "r"(x[12]), "r"(x[14]), "r"(m)

#define BFTQ_CHAIN4_BODY(first)                                                              \
  first " %0, %10, %14, %0;  madc.hi.cc.u32 %1, %10, %14, %1;"                                \
  "madc.lo.cc.u32 %2, %11, %14, %2;  madc.hi.cc.u32 %3, %11, %14, %3;"                        \
  "madc.lo.cc.u32 %4, %12, %14, %4;  madc.hi.cc.u32 %5, %12, %14, %5;"                        \
  "madc.lo.cc.u32 %6, %13, %14, %6;  madc.hi.cc.u32 %7, %13, %14, %7;"                        \
  "addc.cc.u32 %8, %8, 0; addc.u32 %9, %9, 0;"
#define BFTQ_CHAIN4_OPS(p, c0, c1, x)                                                                                     \
  : "+r"(p[0]), "+r"(p[1]), "+r"(p[2]), "+r"(p[3]), "+r"(p[4]), "+r"(p[5]), "+r"(p[6]), "+r"(p[7]), "+r"(c0), "+r"(c1)      \
  : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(m)

// E/O accumulators of one lane: positions 0..W+2 (+1 scratch so every chain has two carry limbs).
template <int W>
struct Acc {
  uint32_t E[W + 4];
  uint32_t O[W + 2];
};

// One carry chain over W/2 (lo,hi) pairs starting at p[0], operand limbs x[0], x[2], ..., carry limbs c0, c1.
template <int W> struct Chain;
template <> struct Chain<16> {
  static __device__ __forceinline__ void run(uint32_t* p, uint32_t& c0, uint32_t& c1, const uint32_t* x, const uint32_t m) {
    asm volatile(BFTQ_CHAIN8_BODY("mad.lo.cc.u32") BFTQ_CHAIN8_OPS(p, c0, c1, x));
  }
  static __device__ __forceinline__ void run_cin(uint32_t* p, uint32_t& c0, uint32_t& c1, const uint32_t* x, const uint32_t m, const uint32_t cin) {
    asm volatile("{ .reg .u32 t; add.cc.u32 t, %27, 0xffffffff;" BFTQ_CHAIN8_BODY("madc.lo.cc.u32") "}"
                 BFTQ_CHAIN8_OPS(p, c0, c1, x), "r"(cin));
  }
};
template <> struct Chain<8> {
  static __device__ __forceinline__ void run(uint32_t* p, uint32_t& c0, uint32_t& c1, const uint32_t* x, const uint32_t m) {
    asm volatile(BFTQ_CHAIN4_BODY("mad.lo.cc.u32") BFTQ_CHAIN4_OPS(p, c0, c1, x));
  }
  static __device__ __forceinline__ void run_cin(uint32_t* p, uint32_t& c0, uint32_t& c1, const uint32_t* x, const uint32_t m, const uint32_t cin) {
    asm volatile("{ .reg .u32 t; add.cc.u32 t, %15, 0xffffffff;" BFTQ_CHAIN4_BODY("madc.lo.cc.u32") "}"
                 BFTQ_CHAIN4_OPS(p, c0, c1, x), "r"(cin));
  }
};

// offset 0: even limbs -> E pairs (k,k+1), odd limbs -> O pairs (O[k-1],O[k]).  `cin` (0/1) enters
// the even chain at position 0.
template <int W>
__device__ __forceinline__ void mac_off0_even(Acc<W>& A, const uint32_t (&x)[W], const uint32_t m, const uint32_t cin) {
  Chain<W>::run_cin(A.E, A.E[W], A.E[W + 1], x, m, cin);
}
template <int W>
__device__ __forceinline__ void mac_off0_even_nocin(Acc<W>& A, const uint32_t (&x)[W], const uint32_t m) {
  Chain<W>::run(A.E, A.E[W], A.E[W + 1], x, m);
}
template <int W>
__device__ __forceinline__ void mac_off0_odd(Acc<W>& A, const uint32_t (&x)[W], const uint32_t m) {
  Chain<W>::run(A.O, A.O[W], A.O[W + 1], x + 1, m);          // x[1], x[3], ...
}
// offset 1: even limbs -> O pairs (O[k],O[k+1]), odd limbs -> E pairs (E[k+1],E[k+2]).
template <int W>
__device__ __forceinline__ void mac_off1_even(Acc<W>& A, const uint32_t (&x)[W], const uint32_t m) {
  Chain<W>::run(A.O, A.O[W], A.O[W + 1], x, m);
}
template <int W>
__device__ __forceinline__ void mac_off1_odd(Acc<W>& A, const uint32_t (&x)[W], const uint32_t m) {
  Chain<W>::run(A.E + 2, A.E[W + 2], A.E[W + 3], x + 1, m);
}

// Ripple-add a 32-bit value into v[0..15]; returns the carry out (0/1).
template <int W>
__device__ __forceinline__ uint32_t ripple_add(uint32_t (&v)[W], const uint32_t x) {
  uint32_t c;
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(v[0]) : "r"(x));
#pragma unroll
  for (int k = 1; k < W; k++) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(v[k]));
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
}
// v -= x (one limb), returns borrow out (0/1).
template <int W>
__device__ __forceinline__ uint32_t ripple_sub(uint32_t (&v)[W], const uint32_t x) {
  uint32_t b;
  asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(v[0]) : "r"(x));
#pragma unroll
  for (int k = 1; k < W; k++) asm volatile("subc.cc.u32 %0, %0, 0;" : "+r"(v[k]));
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(b));
  return b & 1u;
}
// d = v - n (limb-wise with borrow chain), returns borrow out (0/1).
template <int W>
__device__ __forceinline__ uint32_t sub_n(uint32_t (&d)[W], const uint32_t (&v)[W], const uint32_t (&n)[W]) {
  uint32_t b;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d[0]) : "r"(v[0]), "r"(n[0]));
#pragma unroll
  for (int k = 1; k < W; k++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d[k]) : "r"(v[k]), "r"(n[k]));
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(b));
  return b & 1u;
}

// Carry (or borrow) into each lane of a 4-lane group from per-lane generate / propagate flags:
// in[r] = gen[r-1] | (prop[r-1] & in[r-1]), in[0] = 0.  Returns this lane's carry-in and, in `out`,
// the carry out of the top lane.  gbits/pbits are ballots already shifted to the group's bit 0.
__device__ __forceinline__ uint32_t lane_carry_in(const uint32_t gbits, const uint32_t pbits, const int r, uint32_t& out) {
  const uint32_t c1 = gbits & 1u;
  const uint32_t c2 = ((gbits >> 1) & 1u) | (((pbits >> 1) & 1u) & c1);
  const uint32_t c3 = ((gbits >> 2) & 1u) | (((pbits >> 2) & 1u) & c2);
  out = ((gbits >> 3) & 1u) | (((pbits >> 3) & 1u) & c3);
  return r == 0 ? 0u : (r == 1 ? c1 : (r == 2 ? c2 : c3));
}

// The tail of a Montgomery product: merges the E / O accumulators and the pending carries into W limbs per lane,
// resolves the carries across the four lanes and subtracts n once when the result overflowed 2^(128 W).
template <int W>
__device__ __forceinline__ void mont_finish(uint32_t (&out)[W], Acc<W>& A, const uint32_t cin, const uint32_t Z, const uint32_t (&n)[W],
                                            const int r, const int gbase) {
  // ---- merge E, O and the pending carry into 16 limbs + overflow ------------------------------
  uint32_t v[W], hi;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(v[0]) : "r"(A.E[0]), "r"(Z));
#pragma unroll
  for (int k = 1; k < W; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(v[k]) : "r"(A.E[k]), "r"(A.O[k - 1]));
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(hi) : "r"(A.E[W]), "r"(A.O[W - 1]));
  // the lane's overflow (a few units) belongs to the lane above; the top lane's is bit 2048+
  uint32_t from_below = __shfl_up_sync(kFull, hi, 1, T);
  if (r == 0) from_below = 0u;
  const uint32_t g = ripple_add(v, from_below + cin);
  bool ones = true;
#pragma unroll
  for (int k = 0; k < W; k++) ones = ones && (v[k] == 0xffffffffu);
  const uint32_t gb = __ballot_sync(kFull, g != 0u) >> gbase;
  const uint32_t pb = __ballot_sync(kFull, ones) >> gbase;
  uint32_t ctop;
  const uint32_t ci = lane_carry_in(gb, pb, r, ctop);
  ripple_add(v, ci);
  const uint32_t top_hi = __shfl_sync(kFull, hi, gbase + T - 1);
  const bool overflow = (top_hi + ctop) != 0u;          // result >= 2^2048: subtract n once
  // ---- conditional subtraction ---------------------------------------------------------------
  if (__any_sync(kFull, overflow)) {
    uint32_t d[W];
    const uint32_t bo = sub_n(d, v, n);
    bool zeros = true;
#pragma unroll
    for (int k = 0; k < W; k++) zeros = zeros && (d[k] == 0u);
    const uint32_t bgb = __ballot_sync(kFull, bo != 0u) >> gbase;
    const uint32_t bpb = __ballot_sync(kFull, zeros) >> gbase;
    uint32_t btop;
    const uint32_t bi = lane_carry_in(bgb, bpb, r, btop);
    ripple_sub(d, bi);
    if (overflow) {
#pragma unroll
      for (int k = 0; k < W; k++) v[k] = d[k];
    }
  }
#pragma unroll
  for (int k = 0; k < W; k++) out[k] = v[k];
}

// out = a * b * R^-1 mod n with R = 2^(128 W), out < R ("almost Montgomery").  a, b < R as W limbs/lane.
// All 32 lanes of the warp must call this together.
template <int W, bool STEP_SYNC = false>
__device__ __forceinline__ void mont_mul(uint32_t (&out)[W], const uint32_t (&a)[W], const uint32_t (&b)[W],
                                         const uint32_t (&n)[W], const uint32_t n0inv, const int r, const int gbase) {
  Acc<W> A;
#pragma unroll
  for (int k = 0; k < W + 4; k++) A.E[k] = 0u;
#pragma unroll
  for (int k = 0; k < W + 2; k++) A.O[k] = 0u;
  uint32_t cin = 0u;      // 1-bit carry pending at position 0
  uint32_t Z = 0u;        // odd-side limb pending at position 0 (the upper half of the O pair the shift cut)
#ifndef BFTQ_MUL_UNROLL
#define BFTQ_MUL_UNROLL 1
#endif
  constexpr int kMulUnroll = BFTQ_MUL_UNROLL;      // owner steps per loop body (code size x this)
#pragma unroll kMulUnroll
  for (int owner = 0; owner < T; owner++) {
    if (STEP_SYNC) __syncthreads();             // keep the block's warps in phase (see BFTQ_K1_SYNC)
    const int src = gbase + owner;
#pragma unroll
    for (int jj = 0; jj < W; jj += 2) {
      const uint32_t b0 = __shfl_sync(kFull, b[jj], src);
      const uint32_t b1 = __shfl_sync(kFull, b[jj + 1], src);
      // ---- offset 0 -------------------------------------------------------------------------
      mac_off0_even(A, a, b0, cin);
      uint32_t q0 = (A.E[0] + Z) * n0inv;
      q0 = __shfl_sync(kFull, q0, gbase);
      mac_off0_odd(A, a, b0);
      mac_off1_even(A, a, b1);
      mac_off1_odd(A, a, b1);
      mac_off0_even_nocin(A, n, q0);
      mac_off0_odd(A, n, q0);
      // ---- offset 1 -------------------------------------------------------------------------
      // position 0 = E[0] + Z: zero mod 2^32 in lane 0, its carry moves into position 1
      const uint64_t s0 = (uint64_t)A.E[0] + Z;
      const uint32_t p0 = (uint32_t)s0, c0 = (uint32_t)(s0 >> 32);
      uint32_t q1 = (A.E[1] + A.O[0] + c0) * n0inv;
      q1 = __shfl_sync(kFull, q1, gbase);
      mac_off1_even(A, n, q1);
      mac_off1_odd(A, n, q1);
      // positions 0 and 1 leave the lane: lane 0's are zero, the others' go to the lane below
      const uint64_t s1 = (uint64_t)A.E[1] + A.O[0] + c0;
      const uint32_t p1 = (uint32_t)s1;
      cin = (uint32_t)(s1 >> 32);
      Z = A.O[1];
      uint32_t r0 = __shfl_down_sync(kFull, p0, 1, T);
      uint32_t r1 = __shfl_down_sync(kFull, p1, 1, T);
      if (r == T - 1) { r0 = 0u; r1 = 0u; }
      // shift down two limbs (register renaming)
#pragma unroll
      for (int k = 0; k < W + 2; k++) A.E[k] = A.E[k + 2];
      A.E[W + 2] = 0u; A.E[W + 3] = 0u;
#pragma unroll
      for (int k = 0; k < W; k++) A.O[k] = A.O[k + 2];
      A.O[W] = 0u; A.O[W + 1] = 0u;
      asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, 0; addc.u32 %3, %3, 0;"
                   : "+r"(A.E[W - 2]), "+r"(A.E[W - 1]), "+r"(A.E[W]), "+r"(A.E[W + 1]) : "r"(r0), "r"(r1));
    }
  }
  mont_finish(out, A, cin, Z, n, r, gbase);
}

// x >= n ?  (lane-distributed compare)
template <int W>
__device__ __forceinline__ bool group_ge(const uint32_t (&x)[W], const uint32_t (&n)[W], const int gbase) {
  bool gt = false, lt = false;
#pragma unroll
  for (int j = W - 1; j >= 0; j--) {
    if (!gt && !lt) { gt = x[j] > n[j]; lt = x[j] < n[j]; }
  }
  const uint32_t gmask = ((1u << T) - 1u) << gbase;
  const uint32_t gtb = __ballot_sync(kFull, gt) & gmask;
  const uint32_t ltb = __ballot_sync(kFull, lt) & gmask;
  return gtb >= ltb;
}

// x -= n when x >= n (x < 2n on entry).
template <int W>
__device__ __forceinline__ void cond_sub(uint32_t (&x)[W], const uint32_t (&n)[W], const int r, const int gbase) {
  const bool ge = group_ge(x, n, gbase);
  uint32_t d[W];
  const uint32_t bo = sub_n(d, x, n);
  bool zeros = true;
#pragma unroll
  for (int k = 0; k < W; k++) zeros = zeros && (d[k] == 0u);
  const uint32_t bgb = __ballot_sync(kFull, bo != 0u) >> gbase;
  const uint32_t bpb = __ballot_sync(kFull, zeros) >> gbase;
  uint32_t btop;
  const uint32_t bi = lane_carry_in(bgb, bpb, r, btop);
  ripple_sub(d, bi);
  if (ge) {
#pragma unroll
    for (int k = 0; k < W; k++) x[k] = d[k];
  }
}

}  // namespace r32
}  // namespace bftq
#include "rsa_square_r32.cuh"
namespace bftq {
namespace r32 {

struct RsaKey32 {               // per key, radix 2^32 little-endian words
  uint32_t n[64];
  uint32_t r2[64];              // 2^4096 mod n
  uint32_t n0inv;               // -n^-1 mod 2^32
  uint32_t e;
  uint32_t nbits;
  uint32_t pad;
};

// In-block barriers that keep the four warps of a block in phase (see the kernel's task loop).  Measured on B200
// (gpurun_out -> profiles/k1_sync_variants_r02.txt, 65 536 / 524 288 items, two streams): 0 = none 53.2 / 46.9 M verifies/s,
// 1 = once per task 56.0 / 57.0, 2 = after every Montgomery product 58.0 / 58.0, 3 = 2 + every owner step 58.0 / 58.0.
// Warps that drift apart fetch different parts of the 15 KB hot loop and evict each other from the instruction caches
// (the long-batch rate of the barrier-free kernel DROPS, 44.8 M/s on one stream); a barrier per product costs 18
// bar.sync per 2.3 M instructions.
#ifndef BFTQ_K1_SYNC
#define BFTQ_K1_SYNC 2
#endif
// 1 = the exponentiation runs as one loop over a four-op program (one squaring and one product instance in the kernel:
// 66 KB of SASS instead of 116 KB).  Measured with BFTQ_K1_SYNC 2 (profiles/k1_variants_r02c.txt, two streams, 65 536 items):
// straight-line 58.1 M verifies/s, unified 58.7, unified + two owner steps of the squaring per loop body
// (BFTQ_SQR_UNROLL 2, the default now) 60.6, + BFTQ_MUL_UNROLL 2 60.8 (not taken: code size for 0.2 %).
#ifndef BFTQ_K1_UNIFIED
#define BFTQ_K1_UNIFIED 1
#endif
// SQ: the squarings of the exponentiation go through mont_sqr (rsa_square_r32.cuh) instead of mont_mul(y, y).
template <int BLOCK, int MIN_BLOCKS, bool SQ>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
rsa_verify_r32_kernel(const RsaKey32* __restrict__ keys, const uint32_t nkeys, const uint32_t* __restrict__ key_idx,
                      const uint8_t* __restrict__ sig, const uint8_t* __restrict__ digest, const uint32_t hash_alg,
                      const uint64_t n_items, const uint32_t flags, const uint8_t* __restrict__ pre_status,
                      uint8_t* __restrict__ status) {
  constexpr int W = 16;
  constexpr int kGroupsPerWarp = 32 / T;
  // s*R mod n is only needed again for exponents with interior 1 bits (never for 65537): park it in
  // shared memory instead of 16 registers.
  __shared__ uint32_t xm_s[W][BLOCK];
  __shared__ int nbmax_s;
  constexpr bool kStepSync = BFTQ_K1_SYNC >= 3;
  const int lane = threadIdx.x & 31;
  const int r = lane & (T - 1);
  const int gbase = lane & ~(T - 1);
  const int plen = c_hash_prefix[hash_alg].len;
  const int dlen = c_hash_prefix[hash_alg].dlen;
  const uint64_t warp_global = (uint64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
  const uint64_t warps_total = (uint64_t)gridDim.x * (BLOCK / 32);
  const uint32_t gmask = ((1u << T) - 1u) << gbase;

  // The trip count is uniform over the block (a warp whose eight signatures lie behind the end computes on clamped items
  // and stores nothing), so the warps of a block may meet at barriers: BFTQ_K1_SYNC 1 = once per task, 2 = after every
  // Montgomery product.  Warps that stay in phase fetch the same instructions at the same time (the hot loop is 15 KB).
  for (uint64_t bbase = (uint64_t)blockIdx.x * (BLOCK / 32) * kGroupsPerWarp; bbase < n_items; bbase += warps_total * kGroupsPerWarp) {
    const uint64_t wbase = bbase + (uint64_t)(threadIdx.x >> 5) * kGroupsPerWarp;
    if (BFTQ_K1_SYNC >= 1) __syncthreads();
    const uint64_t item_raw = wbase + (uint64_t)(lane / T);
    const bool valid = item_raw < n_items;
    const uint64_t item = valid ? item_raw : (n_items - 1);
    uint32_t kidx = __ldg(key_idx + item);
    const bool known = kidx < nkeys;
    if (!known) kidx = 0u;
    const RsaKey32* __restrict__ key = keys + kidx;

    uint32_t nd[W], y[W], t[W];
#pragma unroll
    for (int j = 0; j < W; j++) nd[j] = __ldg(&key->n[r * W + j]);
    const uint32_t n0inv = __ldg(&key->n0inv);
    const uint32_t e = __ldg(&key->e);
    const uint8_t* sp = sig + item * (uint64_t)kRsaBytes;
    bool s_ge_n;
#pragma unroll
    for (int j = 0; j < W; j++) y[j] = be_word(sp, r * W + j);
    s_ge_n = group_ge(y, nd, gbase);
    const int nb = 32 - __clz(e);
    int nbmax = nb;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nbmax = max(nbmax, __shfl_xor_sync(kFull, nbmax, o));
    if (BFTQ_K1_SYNC >= 2) {                                 // barriers inside the exponent loop: its trip count must be uniform over the block
      if (threadIdx.x == 0) nbmax_s = 0;
      __syncthreads();
      if (lane == 0) atomicMax(&nbmax_s, nbmax);
      __syncthreads();
      nbmax = nbmax_s;
    }
#if BFTQ_K1_UNIFIED
    // The whole exponentiation as ONE loop over a small program, so that the kernel holds a single instance of the
    // squaring and a single instance of the general product (the straight-line form below has two and three: 75 KB of
    // hot code per task instead of 31 KB; the instruction caches hold 32 KB).
    //   op 0: y = s * R^2 (to Montgomery form; also parked as xm)      op 1: the squaring of `bit`
    //   op 2: y *= xm after the squaring of an interior 1 bit            op 3: the last product, by the PLAIN s (or 1)
    int op = 0, bit = nbmax - 2;
#pragma unroll 1
    for (;;) {
      if (op == 1) {
        if (SQ) mont_sqr<kStepSync>(t, y, nd, n0inv, r, gbase); else mont_mul<W, kStepSync>(t, y, y, nd, n0inv, r, gbase);
        if (BFTQ_K1_SYNC >= 2) __syncthreads();              // every warp of the block runs nbmax - 1 squarings
        if (bit <= nb - 2) {
#pragma unroll
          for (int j = 0; j < W; j++) y[j] = t[j];
        }
        const bool mul = bit >= 1 && bit <= nb - 2 && ((e >> bit) & 1u);
        if (__any_sync(kFull, mul)) op = 2;
        else { bit--; op = bit >= 0 ? 1 : 3; }
      } else {
        uint32_t bop[W];
        if (op == 0) {
#pragma unroll
          for (int j = 0; j < W; j++) bop[j] = __ldg(&key->r2[r * W + j]);
        } else if (op == 2) {
#pragma unroll
          for (int j = 0; j < W; j++) bop[j] = xm_s[j][threadIdx.x];
        } else {                                              // plain s (bit 0 set) or plain 1: leaves Montgomery form
#pragma unroll
          for (int j = 0; j < W; j++) bop[j] = ((e & 1u) && nb >= 2) ? be_word(sp, r * W + j) : ((r == 0 && j == 0) ? 1u : 0u);
        }
        mont_mul(t, y, bop, nd, n0inv, r, gbase);
        if (op == 3) break;
        if (op == 0) {
#pragma unroll
          for (int j = 0; j < W; j++) { y[j] = t[j]; xm_s[j][threadIdx.x] = t[j]; }
          op = bit >= 0 ? 1 : 3;
        } else {
          if (bit <= nb - 2 && ((e >> bit) & 1u)) {
#pragma unroll
            for (int j = 0; j < W; j++) y[j] = t[j];
          }
          bit--;
          op = bit >= 0 ? 1 : 3;
        }
      }
    }
#else
    {
      uint32_t r2[W];
#pragma unroll
      for (int j = 0; j < W; j++) r2[j] = __ldg(&key->r2[r * W + j]);
      mont_mul(t, y, r2, nd, n0inv, r, gbase);              // s * R mod n (almost reduced)
#pragma unroll
      for (int j = 0; j < W; j++) y[j] = t[j];
    }
#pragma unroll
    for (int j = 0; j < W; j++) xm_s[j][threadIdx.x] = y[j];
#pragma unroll 1
    for (int bit = nbmax - 2; bit >= 1; bit--) {
      const bool active = bit <= nb - 2;
      if (SQ) mont_sqr<kStepSync>(t, y, nd, n0inv, r, gbase); else mont_mul<W, kStepSync>(t, y, y, nd, n0inv, r, gbase);
      if (BFTQ_K1_SYNC >= 2) __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < W; j++) y[j] = t[j];
      }
      const bool mul = active && ((e >> bit) & 1u);
      if (__any_sync(kFull, mul)) {
        uint32_t xm[W];
#pragma unroll
        for (int j = 0; j < W; j++) xm[j] = xm_s[j][threadIdx.x];
        mont_mul(t, y, xm, nd, n0inv, r, gbase);
        if (mul) {
#pragma unroll
          for (int j = 0; j < W; j++) y[j] = t[j];
        }
      }
    }
    if (__any_sync(kFull, nb >= 2)) {
      if (SQ) mont_sqr(t, y, nd, n0inv, r, gbase); else mont_mul(t, y, y, nd, n0inv, r, gbase);
      if (nb >= 2) {
#pragma unroll
        for (int j = 0; j < W; j++) y[j] = t[j];
      }
    }
    {
      uint32_t m1[W];                                       // plain s (bit 0 set) or plain 1
#pragma unroll
      for (int j = 0; j < W; j++) m1[j] = ((e & 1u) && nb >= 2) ? be_word(sp, r * W + j) : ((r == 0 && j == 0) ? 1u : 0u);
      mont_mul(t, y, m1, nd, n0inv, r, gbase);            // plain operand: leaves Montgomery form
    }
#endif
    cond_sub(t, nd, r, gbase);                            // t < 2^2048 < 2n  ->  t mod n

    const uint8_t* dp = digest + item * (uint64_t)dlen;
    bool eq = true;
#pragma unroll
    for (int j = 0; j < W; j++) eq = eq && (em_word(r * W + j, dp, plen, dlen, hash_alg) == t[j]);
    const uint32_t eqb = __ballot_sync(kFull, eq) & gmask;
    if (valid && r == 0) {
      uint8_t st = (eqb == gmask) ? (uint8_t)0 : (uint8_t)1;
      if ((flags & 1u) && s_ge_n) st = 1;
      if (__ldg(&key->nbits) != 2048u) st = 1;             // not this kernel's key class
      if (!known) st = 4;
      if (pre_status != nullptr) {
        const uint8_t pre = __ldg(pre_status + item_raw);
        if (pre != 0) st = pre;
      }
      status[item_raw] = st;
    }
  }
}

}  // namespace r32
}  // namespace bftq
