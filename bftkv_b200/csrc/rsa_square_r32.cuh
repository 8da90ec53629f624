// EXPERIMENTAL (opt-in: BFTQ_RSA_KERNEL=r32sq; not validated on a GPU yet) — K1 with a dedicated Montgomery SQUARING.
//
// 16 of the 18 products of an e = 65537 verification are squarings, and a square needs only half of the limb
// products of a general product.  In the lane-distributed layout of rsa_verify_r32.cuh (4 lanes x 16 limbs) the lower
// triangle cannot simply be skipped inside the interleaved loop — idle lanes still occupy the FMA-heavy pipe — so the
// square is computed FIRST, in three lock-step product slots per lane
//     slot 0   A_L * A_L                                 (256 limb products here; 136 with inner symmetry, later)
//     slot 1   A_L * A_{(L+1) mod 4}                     (256)  the four pairs at cyclic distance 1
//     slot 2   A_{L&1} * half_L(A_{(L&1)+2})             (128)  the two pairs at distance 2, split between two lanes
// (lanes 2 and 3 hold a copy of blocks 0 and 1 for slot 2), the slot results are scattered through shared memory into
// a^2 as eight 16-limb units with a static table (tools/emu_sq.py prints it; at most five contributions per unit, the
// cross terms added twice), lane r keeping units r and 4 + r, and then the REDUCTION half of mont_mul's loop runs with
// the high units fed in at the top, two limbs per round.  FMA-heavy work per squaring: 640 + 1024 IMAD.WIDE per lane
// instead of 2048.  tools/emu_sq.py emulates this file limb for limb (block_mul's two chains per row, the scatter
// table, the carry normalisation, q0 with the pending carry, the feed) against big-int arithmetic.
#pragma once
#include "rsa_verify_r32.cuh"

namespace bftq {
namespace r32 {

constexpr int kSqWords = 88;                 // shared-memory words per lane: slot 0 (32) | slot 1 (32) | slot 2 (24)

// One contribution to a 16-limb unit of a^2: positions [dst, dst + len) of the unit += words [word + dst, ...) of
// lane `lane`'s column (word = slot base + offset - dst, so the word for position p is word + p), once or twice.
struct SqContrib { int8_t lane, word, dst, len, twice; };
__constant__ SqContrib c_sq_tab[8][5] = {
  /* unit 0 */ {{0, 0, 0, 16, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
  /* unit 1 */ {{0, 16, 0, 16, 0}, {0, 32, 0, 16, 1}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
  /* unit 2 */ {{0, 48, 0, 16, 1}, {0, 64, 0, 16, 1}, {1, 0, 0, 16, 0}, {2, 56, 8, 8, 1}, {0, 0, 0, 0, 0}},
  /* unit 3 */ {{0, 80, 0, 8, 1}, {1, 16, 0, 16, 0}, {1, 32, 0, 16, 1}, {2, 72, 0, 16, 1}, {3, 32, 0, 16, 1}},
  /* unit 4 */ {{1, 48, 0, 16, 1}, {1, 64, 0, 16, 1}, {2, 0, 0, 16, 0}, {3, 48, 0, 16, 1}, {3, 56, 8, 8, 1}},
  /* unit 5 */ {{1, 80, 0, 8, 1}, {2, 16, 0, 16, 0}, {2, 32, 0, 16, 1}, {3, 72, 0, 16, 1}, {0, 0, 0, 0, 0}},
  /* unit 6 */ {{2, 48, 0, 16, 1}, {3, 0, 0, 16, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
  /* unit 7 */ {{3, 16, 0, 16, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
};

// acc[0 .. NR+16) = x (16 limbs) * rows (NR limbs): per row an even chain (x0, x2, ...) at offset i and an odd chain
// (x1, x3, ...) at offset i + 1 into ONE accumulator; acc[NR+16 .. NR+19) are carry scratch and end up zero.
template <int NR, typename RowFn>
__device__ __forceinline__ void block_mul(uint32_t (&acc)[NR + 19], const uint32_t (&x)[16], RowFn row) {
#pragma unroll
  for (int k = 0; k < NR + 19; k++) acc[k] = 0u;
#pragma unroll
  for (int i = 0; i < NR; i++) {
    const uint32_t m = row(i);
    Chain<16>::run(acc + i, acc[i + 16], acc[i + 17], x, m);
    Chain<16>::run(acc + i + 1, acc[i + 17], acc[i + 18], x + 1, m);
  }
}

// v (16 limbs) += w (16 limbs), carry out into ov.
__device__ __forceinline__ void add16(uint32_t (&v)[16], uint32_t& ov, const uint32_t (&w)[16]) {
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(v[0]) : "r"(w[0]));
#pragma unroll
  for (int k = 1; k < 16; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k]));
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(ov));
}

// Carry normalisation of four 16-limb units held by the four lanes of a group (unit of lane r above unit of lane
// r - 1): every lane's overflow word goes into the lane above, `into_first` enters lane 0; returns the overflow of
// the top unit.  Same generate / propagate resolution as mont_mul's tail.
__device__ __forceinline__ uint32_t normalise4(uint32_t (&v)[16], const uint32_t hi, const uint32_t into_first, const int r, const int gbase) {
  uint32_t from_below = __shfl_up_sync(kFull, hi, 1, T);
  if (r == 0) from_below = into_first;
  const uint32_t g = ripple_add(v, from_below);
  bool ones = true;
#pragma unroll
  for (int k = 0; k < 16; k++) ones = ones && (v[k] == 0xffffffffu);
  const uint32_t gb = __ballot_sync(kFull, g != 0u) >> gbase;
  const uint32_t pb = __ballot_sync(kFull, ones) >> gbase;
  uint32_t ctop;
  const uint32_t ci = lane_carry_in(gb, pb, r, ctop);
  ripple_add(v, ci);
  const uint32_t top_hi = __shfl_sync(kFull, hi, gbase + T - 1);
  return top_hi + ctop;
}

// out = a * a * R^-1 mod n, out < R ("almost Montgomery"), R = 2^2048.  All 32 lanes of the warp call this together.
// sm = this block's shared-memory scratch (kSqWords * blockDim.x words), column = threadIdx.x.
__device__ __forceinline__ void mont_sqr(uint32_t (&out)[16], const uint32_t (&a)[16], const uint32_t (&n)[16], const uint32_t n0inv,
                                         const int r, const int gbase, uint32_t* __restrict__ sm) {
  constexpr int W = 16;
  const int bdim = blockDim.x;
  const int col = threadIdx.x;
  const int gcol = col - r;                                   // column of lane 0 of this group
  // ---- the three product slots, each stored to shared memory as soon as it is complete -----------------------
  {
    uint32_t acc[16 + 19];
    block_mul<16>(acc, a, [&](int i) { return a[i]; });                                           // slot 0
#pragma unroll
    for (int k = 0; k < 32; k++) sm[(0 + k) * bdim + col] = acc[k];
    const int next = gbase + ((r + 1) & 3);
    block_mul<16>(acc, a, [&](int i) { return __shfl_sync(kFull, a[i], next); });                  // slot 1
#pragma unroll
    for (int k = 0; k < 32; k++) sm[(32 + k) * bdim + col] = acc[k];
  }
  {
    uint32_t cpy[16];
#pragma unroll
    for (int k = 0; k < 16; k++) cpy[k] = __shfl_sync(kFull, a[k], gbase + (r & 1));
    const int src = gbase + (r | 2);
    uint32_t acc[8 + 19];
    block_mul<8>(acc, cpy, [&](int i) {                                                            // slot 2
      const uint32_t lo = __shfl_sync(kFull, a[i], src), hi = __shfl_sync(kFull, a[8 + i], src);
      return r < 2 ? lo : hi;
    });
#pragma unroll
    for (int k = 0; k < 24; k++) sm[(64 + k) * bdim + col] = acc[k];
  }
  __syncwarp();
  // ---- unit sums: lane r builds units r (low half of a^2) and 4 + r (high half) from the table ----------------
  uint32_t tlo[16], thi[16], ovlo = 0u, ovhi = 0u;
#pragma unroll
  for (int k = 0; k < 16; k++) { tlo[k] = 0u; thi[k] = 0u; }
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int u = half * 4 + r;
#pragma unroll 1
    for (int c = 0; c < 5; c++) {
      const SqContrib e = c_sq_tab[u][c];
      uint32_t w[16];
#pragma unroll
      for (int p = 0; p < 16; p++) {
        const bool in = p >= e.dst && p < e.dst + e.len;
        w[p] = in ? sm[(e.word + p) * bdim + gcol + e.lane] : 0u;
      }
      if (half == 0) add16(tlo, ovlo, w); else add16(thi, ovhi, w);
      const uint32_t again = e.twice ? 0xffffffffu : 0u;
#pragma unroll
      for (int p = 0; p < 16; p++) w[p] &= again;
      if (half == 0) add16(tlo, ovlo, w); else add16(thi, ovhi, w);
    }
  }
  __syncwarp();                                               // the scratch may be overwritten by the next squaring
  const uint32_t top_lo = normalise4(tlo, ovlo, 0u, r, gbase);
  normalise4(thi, ovhi, top_lo, r, gbase);                    // a^2 < 2^4096: the top overflow is zero
  // ---- reduction: mont_mul's loop without the a x b chains; the high units enter at the top lane --------------
  Acc<W> A;
#pragma unroll
  for (int k = 0; k < W; k++) A.E[k] = tlo[k];
#pragma unroll
  for (int k = W; k < W + 4; k++) A.E[k] = 0u;
#pragma unroll
  for (int k = 0; k < W + 2; k++) A.O[k] = 0u;
  uint32_t cin = 0u, Z = 0u;
#pragma unroll 1
  for (int owner = 0; owner < T; owner++) {
    const int src = gbase + owner;
#pragma unroll
    for (int jj = 0; jj < W; jj += 2) {
      const uint32_t f0 = __shfl_sync(kFull, thi[jj], src);
      const uint32_t f1 = __shfl_sync(kFull, thi[jj + 1], src);
      uint32_t q0 = (A.E[0] + Z + cin) * n0inv;
      q0 = __shfl_sync(kFull, q0, gbase);
      mac_off0_even(A, n, q0, cin);
      mac_off0_odd(A, n, q0);
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
      if (r == T - 1) { r0 = f0; r1 = f1; }
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
  // ---- merge E, O and the pending carry; conditional subtraction: as mont_mul ----------------------------------
  uint32_t v[W], hi;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(v[0]) : "r"(A.E[0]), "r"(Z));
#pragma unroll
  for (int k = 1; k < W; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(v[k]) : "r"(A.E[k]), "r"(A.O[k - 1]));
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(hi) : "r"(A.E[W]), "r"(A.O[W - 1]));
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
  const bool overflow = (top_hi + ctop) != 0u;
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

// rsa_verify_r32_kernel (rsa_verify_r32.cuh) with the squarings of the exponentiation going through mont_sqr; everything
// else — loads, the first and last products, the comparison with EM, the status rules — is that kernel's code.
// Dynamic shared memory: kSqWords * BLOCK * 4 bytes.
template <int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
rsa_verify_r32sq_kernel(const RsaKey32* __restrict__ keys, const uint32_t nkeys, const uint32_t* __restrict__ key_idx,
                      const uint8_t* __restrict__ sig, const uint8_t* __restrict__ digest, const uint32_t hash_alg,
                      const uint64_t n_items, const uint32_t flags, const uint8_t* __restrict__ pre_status,
                      uint8_t* __restrict__ status) {
  constexpr int W = 16;
  constexpr int kGroupsPerWarp = 32 / T;
  // s*R mod n is only needed again for exponents with interior 1 bits (never for 65537): park it in
  // shared memory instead of 16 registers.
  __shared__ uint32_t xm_s[W][BLOCK];
  extern __shared__ uint32_t sq_sm[];                 // kSqWords * BLOCK words of scratch for mont_sqr (dynamic)
  const int lane = threadIdx.x & 31;
  const int r = lane & (T - 1);
  const int gbase = lane & ~(T - 1);
  const int plen = c_hash_prefix[hash_alg].len;
  const int dlen = c_hash_prefix[hash_alg].dlen;
  const uint64_t warp_global = (uint64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
  const uint64_t warps_total = (uint64_t)gridDim.x * (BLOCK / 32);
  const uint32_t gmask = ((1u << T) - 1u) << gbase;

  for (uint64_t wbase = warp_global * kGroupsPerWarp; wbase < n_items; wbase += warps_total * kGroupsPerWarp) {
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
    {
      uint32_t xs[W], r2[W];
#pragma unroll
      for (int j = 0; j < W; j++) xs[j] = be_word(sp, r * W + j);
      s_ge_n = group_ge(xs, nd, gbase);
#pragma unroll
      for (int j = 0; j < W; j++) r2[j] = __ldg(&key->r2[r * W + j]);
      mont_mul(y, xs, r2, nd, n0inv, r, gbase);           // s * R mod n (almost reduced)
    }
#pragma unroll
    for (int j = 0; j < W; j++) xm_s[j][threadIdx.x] = y[j];
    const int nb = 32 - __clz(e);
    int nbmax = nb;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nbmax = max(nbmax, __shfl_xor_sync(kFull, nbmax, o));
#pragma unroll 1
    for (int bit = nbmax - 2; bit >= 1; bit--) {
      const bool active = bit <= nb - 2;
      mont_sqr(t, y, nd, n0inv, r, gbase, sq_sm);
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
      mont_sqr(t, y, nd, n0inv, r, gbase, sq_sm);
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
