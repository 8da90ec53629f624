// K2 — batched web-of-trust quorum tally for sm_100a.
//
// Replaces quorum/wotqs/wotqs.go:144-206 (IsQuorum / IsThreshold / IsSufficient / Reject over
// `intersection`, which counts duplicates of the INPUT list) as driven per response by
// protocol/client.go:74,77,111,153 and, for reads, protocol/client.go:181-205
// (isThreshold + maxTimestampedValue over the buckets m[t][value]).
//
// One warp per operation: lanes load the operation's responders (key index, status[, t, value
// id]), a ballot per clique gives the valid / failed member masks and popc() is the cardinality
// the predicates compare with (f, min, threshold, suff).  HBM-bound and tiny: 5..17 B per
// responder in, 1..5 B per operation out.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bftq {

constexpr int kMaxQc = 8;            // cliques per quorum (the reference's fixtures have 1..3)

struct QuorumDev {
  int32_t nqc;
  int32_t f[kMaxQc], min[kMaxQc], threshold[kMaxQc], suff[kMaxQc];
  uint32_t nkeys_words;              // words per membership bitmap
  const uint32_t* member_bits;       // nqc x nkeys_words bitmap over key indices
};

__device__ __forceinline__ bool is_member(const QuorumDev& q, int c, uint32_t key) {
  const uint32_t w = key >> 5;
  if (w >= q.nkeys_words) return false;
  return (__ldg(q.member_bits + (size_t)c * q.nkeys_words + w) >> (key & 31)) & 1u;
}

// out_bits[i]: bit0 IsQuorum, bit1 IsThreshold, bit2 IsSufficient (all over responders with
// status == 0), bit3 Reject (over responders with status != 0 — the `failure` list of
// protocol/client.go:77,113,263).
__global__ void __launch_bounds__(256)
tally_kernel(const QuorumDev q, const uint32_t* __restrict__ op_off, const uint32_t* __restrict__ key_idx,
             const uint8_t* __restrict__ status, const uint64_t n_ops, uint8_t* __restrict__ out_bits) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t op = warp; op < n_ops; op += nwarps) {
    const uint32_t lo = __ldg(op_off + op), hi = __ldg(op_off + op + 1);
    int cnt[kMaxQc], bad[kMaxQc];
#pragma unroll
    for (int c = 0; c < kMaxQc; c++) { cnt[c] = 0; bad[c] = 0; }
    for (uint32_t base = lo; base < hi; base += 32) {
      const uint32_t p = base + lane;
      const bool have = p < hi;
      const uint32_t k = have ? __ldg(key_idx + p) : 0xffffffffu;
      const bool ok = have && __ldg(status + p) == 0;
#pragma unroll
      for (int c = 0; c < kMaxQc; c++) {
        if (c < q.nqc) {
          const bool m = have && is_member(q, c, k);
          cnt[c] += __popc(__ballot_sync(0xffffffffu, m && ok));
          bad[c] += __popc(__ballot_sync(0xffffffffu, m && !ok));
        }
      }
    }
    if (lane == 0) {
      bool is_q = q.nqc > 0, is_t = q.nqc > 0, is_s = false, rej = true;
#pragma unroll
      for (int c = 0; c < kMaxQc; c++) {
        if (c < q.nqc) {
          if (q.f[c] > 0 && cnt[c] < q.min[c]) is_q = false;
          if (q.threshold[c] > 0 && cnt[c] < q.threshold[c]) is_t = false;
          if (q.suff[c] > 0 && cnt[c] >= q.suff[c]) is_s = true;
          if (q.f[c] == 0 || bad[c] <= q.f[c]) rej = false;
        }
      }
      out_bits[op] = (uint8_t)((is_q ? 1 : 0) | (is_t ? 2 : 0) | (is_s ? 4 : 0) | (rej ? 8 : 0));
    }
  }
}

// Read tally (protocol/client.go:189-205,207-230): responders with status == 0 are bucketed by
// (t, value id); only the bucket set of the MAXIMUM t is inspected; the first value (in responder
// order — Go iterates a map there, any order is legal) whose responders pass IsThreshold wins.
// out_winner[op] = index (within the op) of the first responder of the winning bucket, or
// 0xffffffff ("errInProgress").  Operations may have at most 32 responders.
__global__ void __launch_bounds__(256)
read_tally_kernel(const QuorumDev q, const uint32_t* __restrict__ op_off, const uint32_t* __restrict__ key_idx,
                  const uint8_t* __restrict__ status, const uint64_t* __restrict__ ts, const uint32_t* __restrict__ value_id,
                  const uint64_t n_ops, uint32_t* __restrict__ out_winner, uint8_t* __restrict__ out_bits,
                  uint8_t* __restrict__ out_decision, uint32_t* __restrict__ out_decided_at) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t op = warp; op < n_ops; op += nwarps) {
    const uint32_t lo = __ldg(op_off + op), hi = __ldg(op_off + op + 1);
    const uint32_t p = lo + lane;
    const bool have = p < hi && lane < 32;
    const uint32_t k = have ? __ldg(key_idx + p) : 0xffffffffu;
    const bool ok = have && __ldg(status + p) == 0;
    const uint64_t t = ok ? __ldg(ts + p) : 0ull;
    const uint32_t v = ok ? __ldg(value_id + p) : 0xffffffffu;
    // max t over the successful responses
    uint64_t maxt = t;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, maxt, o);
      maxt = other > maxt ? other : maxt;
    }
    const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
    const bool at_max = ok && t == maxt;
    const uint32_t same = __match_any_sync(0xffffffffu, at_max ? v : (0x80000000u | (uint32_t)lane)) & __ballot_sync(0xffffffffu, at_max);
    bool pass = at_max && q.nqc > 0;
    int rej_all = 1;
#pragma unroll
    for (int c = 0; c < kMaxQc; c++) {
      if (c < q.nqc) {
        const bool m = have && is_member(q, c, k);
        const uint32_t mm = __ballot_sync(0xffffffffu, m);
        if (q.threshold[c] > 0 && __popc(same & mm) < q.threshold[c]) pass = false;
        const int bad = __popc(mm & ~okmask);
        if (q.f[c] == 0 || bad <= q.f[c]) rej_all = 0;
      }
    }
    const uint32_t winners = __ballot_sync(0xffffffffu, pass);
    if (out_decision != nullptr) {
      // ---- Client.Read's decision in ARRIVAL order (protocol/client.go:250-268): the multicast callback runs once per
      // response; a good one is bucketed and maxTimestampedValue asked, a failed one joins `failure` and q.Reject is
      // asked, and the first decisive response fixes the result.  Only the bucket the newest response joined can have
      // become decisive (the others were inspected before, under the same or a smaller max t), so every lane j can
      // judge "would the callback decide at response j" on its own from the responses 0..j:
      //   value   : ok_j, t_j is the max t of the good responses 0..j, and the bucket (t_j, v_j) among 0..j passes IsThreshold
      //   reject  : response j failed and the failed responses 0..j pass Reject
      const uint32_t upto = 0xffffffffu >> (31 - lane);                      // lanes 0..lane
      uint64_t pmax = t;                                                      // inclusive prefix maximum of t over the good lanes (t = 0 elsewhere)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t other = __shfl_up_sync(0xffffffffu, pmax, o);
        if (lane >= o && other > pmax) pmax = other;
      }
      const uint32_t same_t = __match_any_sync(0xffffffffu, t);
      const uint32_t same_v = __match_any_sync(0xffffffffu, v);
      const uint32_t bucket = same_t & same_v & okmask & upto;                // responses 0..j in the bucket of response j
      bool dv = ok && t == pmax && q.nqc > 0, dr = have && !ok;
#pragma unroll
      for (int c = 0; c < kMaxQc; c++) {
        if (c < q.nqc) {
          const uint32_t mm = __ballot_sync(0xffffffffu, have && is_member(q, c, k));
          if (q.threshold[c] > 0 && __popc(bucket & mm) < q.threshold[c]) dv = false;
          const int bad = __popc(mm & ~okmask & upto);
          if (q.f[c] == 0 || bad <= q.f[c]) dr = false;
        }
      }
      const uint32_t dvm = __ballot_sync(0xffffffffu, dv), drm = __ballot_sync(0xffffffffu, dr);
      const uint32_t any = dvm | drm;
      const int d = any ? __ffs(any) - 1 : 0;
      const uint32_t bucket_d = __shfl_sync(0xffffffffu, bucket, d);
      if (lane == 0) {
        uint8_t dec = 2; uint32_t w = 0xffffffffu, at = hi - lo;
        if (any) {
          at = (uint32_t)d + 1;
          if ((dvm >> d) & 1u) { dec = 0; w = (uint32_t)(__ffs(bucket_d) - 1); } else dec = 1;
        }
        out_decision[op] = dec;
        out_decided_at[op] = at;
        out_winner[op] = w;
        out_bits[op] = (uint8_t)((winners != 0u && okmask != 0u ? 2 : 0) | (rej_all ? 8 : 0));
      }
      continue;
    }
    if (lane == 0) {
      uint32_t w = 0xffffffffu;
      if (okmask != 0u && winners != 0u) {
        const int first = __ffs(winners) - 1;                    // a lane of the winning bucket
        w = (uint32_t)first;
      }
      out_winner[op] = w;
      out_bits[op] = (uint8_t)((w != 0xffffffffu ? 2 : 0) | (rej_all ? 8 : 0));
    }
    // `first` is the lowest lane among qualifying lanes; lanes of one bucket qualify together,
    // so it is the first responder of the first qualifying bucket in responder order.
  }
}

}  // namespace bftq
