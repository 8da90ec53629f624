// K1's dedicated Montgomery SQUARING (radix 2^32, 4 lanes x 16 limbs) — included by rsa_verify_r32.cuh.
//
// 16 of the 18 Montgomery products of an e = 65537 verification are squarings, and a square needs only half of the
// limb products a x a of a general product (the n x q half of the CIOS loop stays as it is).  In the lane-distributed
// layout the lanes run the rounds of mont_mul in lock-step, so skipping "the lower triangle" only pays when every lane
// skips the same amount in the same round.  This tiling does that.  Row J = 16*Y + j (owner lane Y broadcasts limb
// a_J); lane X multiplies a_J by
//     X <  Y :  2 * (A_X with limbs <  j zeroed)             the pairs (i in block X, J) with i_loc >= j
//     X >  Y :  2 * (A_X with limbs <= j zeroed)             the pairs (i in block X, J) with i_loc >  j
//     X == Y :  a_j  +  2 * (A_X with limbs <= j zeroed)     the diagonal term once, the rest of the row twice
// so every unordered limb pair {i, J} of different blocks is met exactly once — in row J by the lane that owns i, or
// in row i by the lane that owns J — and in round j EVERY lane multiplies the limbs j..15 of its own block: the
// lock-step rounds shrink together, 136 limb products per lane and step instead of 256 (544 + 1024 IMAD.WIDE per lane
// and squaring instead of 2048; -20.8 % over the whole verification).  A contribution to position p is always made in
// a row <= p, i.e. before the reduction eliminates that position, so the interleaved reduction of mont_mul is kept
// unchanged; no shared memory, no second pass.
// The doubled operand is the lane-local a2 = 2 * A_X (17 limbs, a2[16] = the bit shifted out): a row uses a2[k] for
// k >= j + 2, two patched limbs at k = j and j + 1, and the bit a2[16] as an addend of the chain's first carry limb
// (free).  The pending 1-bit carry `cin`, which mont_mul feeds into the a x b chain at slot 0, enters with the n x q0
// chain here (the a x a chains no longer start at slot 0).  tools/emu_sq.py emulates this file limb for limb.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bftq {
namespace r32 {

// (lo, hi) += x * m + carry, as one IMAD.WIDE.U32(.X): `first` starts a carry chain, the others continue it.
__device__ __forceinline__ void mad_pair_first(uint32_t& lo, uint32_t& hi, const uint32_t x, const uint32_t m) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(m));
}
__device__ __forceinline__ void mad_pair_next(uint32_t& lo, uint32_t& hi, const uint32_t x, const uint32_t m) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(m));
}
// N (lo, hi) pairs starting at p[0], operand limbs x[0], x[2], ..., then the carry limbs: c0 += t + carry, c1 += carry.
template <int N>
__device__ __forceinline__ void chain_n(uint32_t* p, uint32_t& c0, uint32_t& c1, const uint32_t* x, const uint32_t m, const uint32_t t) {
  if (N == 0) {
    asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, 0;" : "+r"(c0), "+r"(c1) : "r"(t));
    return;
  }
  mad_pair_first(p[0], p[1], x[0], m);
#pragma unroll
  for (int i = 1; i < N; i++) mad_pair_next(p[2 * i], p[2 * i + 1], x[2 * i], m);
  asm volatile("addc.cc.u32 %0, %0, %2; addc.u32 %1, %1, 0;" : "+r"(c0), "+r"(c1) : "r"(t));
}

// One iteration of the squaring loop: the rows JJ and JJ + 1 of owner lane `owner` (JJ even, compile-time), then the
// reduction by two limbs exactly as in mont_mul.
template <int JJ>
__device__ __forceinline__ void sqr_iter(Acc<16>& A, uint32_t& cin, uint32_t& Z, const uint32_t (&a2)[17], const uint32_t (&n)[16],
                                         const uint32_t n0inv, const int r, const int gbase, const int owner) {
  constexpr int W = 16;
  const uint32_t lt1 = r < owner ? ~1u : 0u, nm = r < owner ? ~0u : ~1u, eqm = r == owner ? ~0u : 0u;
  // this lane's own limbs JJ and JJ + 1 (undoubled); the owner's are the round's multipliers
  const uint32_t aj0 = __funnelshift_r(a2[JJ], a2[JJ + 1], 1);
  const uint32_t aj1 = __funnelshift_r(a2[JJ + 1], a2[JJ + 2], 1);
  const uint32_t b0 = __shfl_sync(kFull, aj0, gbase + owner);
  const uint32_t b1 = __shfl_sync(kFull, aj1, gbase + owner);
  // row operands (window slot k <-> own limb k); only the first two limbs of a row differ from a2
  uint32_t m0[18], m1[18];
#pragma unroll
  for (int k = 0; k < 17; k++) { m0[k] = a2[k]; m1[k] = a2[k]; }
  m0[17] = 0u; m1[17] = 0u;
  // lt1 = lt ? ~1 : 0 (a doubled limb without the bit shifted in from the limb below), nm = lt ? ~0 : ~1, eqm = eq ? ~0 : 0
  m0[JJ] = (a2[JJ] & lt1) | (aj0 & eqm);
  m0[JJ + 1] = a2[JJ + 1] & nm;
  m1[JJ + 1] = (a2[JJ + 1] & lt1) | (aj1 & eqm);
  uint32_t top1 = a2[16];
  if (JJ + 2 < W) m1[JJ + 2] = a2[JJ + 2] & nm;
  else top1 = a2[16] & (lt1 >> 1);                    // row 15: limb 15 is doubled only for the lanes below the owner
  const uint32_t t0 = (0u - a2[16]) & b0;            // a2[16] (0/1) x b: lands on the even chain's first carry limb
  const uint32_t t1 = (0u - top1) & b1;
  // ---- offset 0 ---------------------------------------------------------------------------------------------
  chain_n<(W - JJ) / 2>(A.E + JJ, A.E[W], A.E[W + 1], m0 + JJ, b0, t0);                 // even limbs >= JJ      -> E pairs (k, k+1)
  uint32_t q0 = (A.E[0] + Z + cin) * n0inv;
  q0 = __shfl_sync(kFull, q0, gbase);
  chain_n<(W - JJ) / 2>(A.O + JJ, A.O[W], A.O[W + 1], m0 + JJ + 1, b0, 0u);             // odd limbs  >= JJ + 1  -> O pairs (k-1, k)
  chain_n<(W - JJ - 2) / 2>(A.O + JJ + 2, A.O[W], A.O[W + 1], m1 + JJ + 2, b1, t1);     // even limbs >= JJ + 2  -> O pairs (k, k+1)
  chain_n<(W - JJ) / 2>(A.E + JJ + 2, A.E[W + 2], A.E[W + 3], m1 + JJ + 1, b1, 0u);     // odd limbs  >= JJ + 1  -> E pairs (k+1, k+2)
  mac_off0_even(A, n, q0, cin);
  mac_off0_odd(A, n, q0);
  // ---- offset 1 ---------------------------------------------------------------------------------------------
  const uint64_t s0 = (uint64_t)A.E[0] + Z;
  const uint32_t p0 = (uint32_t)s0, c0 = (uint32_t)(s0 >> 32);
  uint32_t q1 = (A.E[1] + A.O[0] + c0) * n0inv;
  q1 = __shfl_sync(kFull, q1, gbase);
  mac_off1_even(A, n, q1);
  mac_off1_odd(A, n, q1);
  const uint64_t s1 = (uint64_t)A.E[1] + A.O[0] + c0;
  const uint32_t p1 = (uint32_t)s1;
  cin = (uint32_t)(s1 >> 32);
  Z = A.O[1];
  uint32_t r0 = __shfl_down_sync(kFull, p0, 1, T);
  uint32_t r1 = __shfl_down_sync(kFull, p1, 1, T);
  if (r == T - 1) { r0 = 0u; r1 = 0u; }
#pragma unroll
  for (int k = 0; k < W + 2; k++) A.E[k] = A.E[k + 2];
  A.E[W + 2] = 0u; A.E[W + 3] = 0u;
#pragma unroll
  for (int k = 0; k < W; k++) A.O[k] = A.O[k + 2];
  A.O[W] = 0u; A.O[W + 1] = 0u;
  asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, 0; addc.u32 %3, %3, 0;"
               : "+r"(A.E[W - 2]), "+r"(A.E[W - 1]), "+r"(A.E[W]), "+r"(A.E[W + 1]) : "r"(r0), "r"(r1));
}

// out = a * a * R^-1 mod n, out < R ("almost Montgomery"), R = 2^2048.  All 32 lanes of the warp call this together.
template <bool STEP_SYNC = false>
__device__ __forceinline__ void mont_sqr(uint32_t (&out)[16], const uint32_t (&a)[16], const uint32_t (&n)[16], const uint32_t n0inv,
                                         const int r, const int gbase) {
  constexpr int W = 16;
  uint32_t a2[17];
  a2[0] = a[0] << 1;
#pragma unroll
  for (int k = 1; k < W; k++) a2[k] = __funnelshift_l(a[k - 1], a[k], 1);
  a2[16] = a[W - 1] >> 31;
  Acc<W> A;
#pragma unroll
  for (int k = 0; k < W + 4; k++) A.E[k] = 0u;
#pragma unroll
  for (int k = 0; k < W + 2; k++) A.O[k] = 0u;
  uint32_t cin = 0u, Z = 0u;
#ifndef BFTQ_SQR_UNROLL
#define BFTQ_SQR_UNROLL 2
#endif
  constexpr int kSqrUnroll = BFTQ_SQR_UNROLL;      // owner steps per loop body (code size x this)
#pragma unroll kSqrUnroll
  for (int owner = 0; owner < T; owner++) {
    if (STEP_SYNC) __syncthreads();             // keep the block's warps in phase (see BFTQ_K1_SYNC in rsa_verify_r32.cuh)
    sqr_iter<0>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<2>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<4>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<6>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<8>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<10>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<12>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
    sqr_iter<14>(A, cin, Z, a2, n, n0inv, r, gbase, owner);
  }
  mont_finish(out, A, cin, Z, n, r, gbase);
}

}  // namespace r32
}  // namespace bftq
