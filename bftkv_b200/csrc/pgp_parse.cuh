// K0 — OpenPGP signature packets parsed ON THE GPU (fast path of the packer) fused with K4's digest.
//
// Signature.Verify's batch form (crypto/pgp/crypto_pgp.go:319-330) spends, per item, ~0.3 us of host CPU on
// packet parsing, keyring lookup and tuple composition; at 47 M verifies/s that is 14 host cores.  Almost every
// bftkv signature has one shape — SignaturePacket.Data is ONE definite-length v4 RSA/SHA-256 binary signature
// packet whose issuer has exactly one usable key in the keyring — so this kernel takes the raw bytes as the caller
// handed them over, and per item (one thread):
//   1. parses the packet with pgp_fastparse.hpp (same rules as pgp_host.hpp, fuzzed against it on the host),
//   2. looks the issuer up in the call's issuer table (EntityList.KeysByIdUsage(id, KeyFlagSign) precomputed
//      per key id on the host: one usable RSA key of the 2048-bit class -> key index; anything else -> host),
//   3. left-pads the signature MPI to the key size (x/crypto padToKeySize) into K1's input layout,
//   4. hashes  signed bytes || hashed area || 04 FF len32  (SHA-256) and applies the 16-bit hash-tag check,
// and leaves key index / padded signature / digest / pre-status exactly as the host packer would have composed
// them, for K1.  Whatever is not that shape is flagged and goes through the host packer afterwards — a fallback
// is always safe; the flag, not a guess, decides.  The host's share drops to copying the raw bytes into staging.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "pgp_digest.cuh"
#include "pgp_fastparse.hpp"

namespace bftq {

struct IssuerEntry {
  uint64_t key_id;
  uint32_t key_idx;        // engine key table index (kind 0)
  uint16_t kbytes;         // pub.Size() of that key
  uint8_t algo;            // the key's public-key algorithm (1, 2 or 3)
  uint8_t kind;            // 0: exactly one usable RSA key, size class 256 -> decide here; 1: leave to the host packer
};

constexpr uint8_t kParseDecided = 0, kParseHost = 1;

// Bytes of  data || hashed || trailer(6)  as one stream, then SHA-256 padding.
struct FastSrc {
  const uint8_t* dp; const uint8_t* hp; uint32_t dlen, hlen, total;
  __device__ __forceinline__ uint32_t byte(uint32_t pos) const {
    if (pos < dlen) return __ldg(dp + pos);
    uint32_t q = pos - dlen;
    if (q < hlen) return __ldg(hp + q);
    q -= hlen;
    if (q < 6) return q == 0 ? 0x04u : q == 1 ? 0xffu : ((hlen >> (8 * (5 - q))) & 0xffu);
    return pos == total ? 0x80u : 0u;
  }
};

__global__ void __launch_bounds__(128)
pgp_parse_digest_kernel(const uint8_t* __restrict__ tbs_blob, const uint64_t* __restrict__ tbs_off, const uint64_t tbs_base,
                        const uint8_t* __restrict__ sig_blob, const uint64_t* __restrict__ sig_off, const uint64_t sig_base, const uint32_t n_items,
                        const IssuerEntry* __restrict__ issuers, const uint32_t n_issuers,
                        uint32_t* __restrict__ out_key_idx, uint8_t* __restrict__ out_sig /* n x 256 */,
                        uint8_t* __restrict__ out_digest /* n x 32 */, uint8_t* __restrict__ out_pre, uint8_t* __restrict__ out_where,
                        // collective form (all nullable): item s is ONE packet of a multi-signature stream — its signed bytes are
                        // tbs number data_idx[s], its packet ends at sig_end[s] (not at the next item's start), and the signer the
                        // quorum tally counts (the dense index of the issuer's entity) goes to out_signer[s]
                        const uint32_t* __restrict__ data_idx = nullptr, const uint64_t* __restrict__ sig_end = nullptr,
                        const uint32_t* __restrict__ hit_signer = nullptr, uint32_t* __restrict__ out_signer = nullptr) {
  const uint32_t item_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = item_raw < n_items;
  const uint32_t item = live ? item_raw : n_items - 1;        // idle lanes shadow the last item (the warp copies together)
  const int lane = threadIdx.x & 31;
  // offsets are the caller's own (relative to its whole blob): the chunk's blobs start at tbs_base / sig_base
  const uint32_t di = data_idx != nullptr ? data_idx[item] : item;
  const uint64_t s0 = tbs_off[di] - tbs_base, s1 = tbs_off[di + 1] - tbs_base, g0 = sig_off[item] - sig_base,
                 g1 = (sig_end != nullptr ? sig_end[item] : sig_off[item + 1]) - sig_base;
  const uint8_t* sg = sig_blob + g0;
  // ---- per item: parse, issuer lookup ---------------------------------------------------------------
  uint8_t where = kParseDecided, pre = 0;
  uint32_t kidx = 0, signer = 0xffffffffu;
  bool copy = false, hash = false;
  fastparse::FastSig f;
  f.mpi_off = 0; f.mpi_len = 0; f.hashed_off = 0; f.hashed_len = 0; f.tag = 0;
  if (s1 - s0 > 0x3fffffffull || fastparse::parse(sg, (size_t)(g1 - g0), f) != fastparse::kFast) where = kParseHost;
  else if (f.hash_id != 8 || f.sig_type != 0x00) where = kParseHost;      // other digests / text mode: host packer (K4 has them all)
  else {
    int hit = -1;
    for (uint32_t i = 0; i < n_issuers; i++) if (issuers[i].key_id == f.issuer) { hit = (int)i; break; }
    if (hit < 0) pre = 4;                                                  // BFTQ_ST_UNKNOWN_SIGNER: the stream ends -> ErrUnknownIssuer
    else {
      const IssuerEntry en = issuers[hit];
      if (en.kind != 0) where = kParseHost;
      else {
        kidx = en.key_idx;
        if (hit_signer != nullptr) signer = hit_signer[hit];
        hash = true;
        if (en.algo != f.pk_algo) pre = 1;                                 // "public key and signature use different algorithms"
        if (f.mpi_len > en.kbytes) { if (!pre) pre = 1; }                  // len(sig) != k
        else copy = true;
      }
    }
  }
  if (where == kParseHost) pre = 6;                                        // BFTQ_ST_MISSING: K1 leaves the item alone
  if (!live) copy = false;
  // ---- per warp: left-pad the signature MPIs into K1's layout (x/crypto padToKeySize), one item at a time,
  // the 32 lanes writing 4 consecutive bytes each: coalesced stores instead of 256 byte-stores per thread ----
  const uint64_t src_pos = g0 + f.mpi_off;
  for (int j = 0; j < 32; j++) {
    if (!__shfl_sync(0xffffffffu, (int)copy, j)) continue;
    const uint32_t it = __shfl_sync(0xffffffffu, item, j);
    const uint32_t len = __shfl_sync(0xffffffffu, f.mpi_len, j);
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)src_pos, j), hi = __shfl_sync(0xffffffffu, (uint32_t)(src_pos >> 32), j);
    const uint8_t* src = sig_blob + (((uint64_t)hi << 32) | lo);
    const uint32_t padn = 256u - len;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out_sig + (size_t)it * 256);
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const uint32_t b0 = (uint32_t)(t * 32 + lane) * 4u;
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t idx = b0 + b;
        const uint32_t byte = idx >= padn ? (uint32_t)__ldg(src + (idx - padn)) : 0u;
        v |= byte << (8 * b);
      }
      dst[t * 32 + lane] = v;
    }
  }
  // ---- per item: digest (K4's SHA-256 arm over three segments) + x/crypto's 16-bit quick check ----------
  if (hash && live) {
    FastSrc src{tbs_blob + s0, sg + f.hashed_off, (uint32_t)(s1 - s0), f.hashed_len, 0};
    src.total = src.dlen + src.hlen + 6;
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    const uint32_t nblocks = (src.total + 9 + 63) / 64;
    for (uint32_t blk = 0; blk < nblocks; blk++) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) v = (v << 8) | src.byte(blk * 64 + 4 * i + b);
        w[i] = v;
      }
      if (blk == nblocks - 1) { w[14] = 0u; w[15] = src.total * 8u; }
      sha256_compress(h, w);
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(out_digest + (size_t)item * 32);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = __byte_perm(h[i], 0, 0x0123);        // big-endian bytes
    if (!pre && (uint16_t)(h[0] >> 16) != f.tag) pre = 2;                   // BFTQ_ST_HASH_TAG
  }
  if (live) { out_key_idx[item] = kidx; out_pre[item] = pre; out_where[item] = where; if (out_signer != nullptr) out_signer[item] = signer; }
}

}  // namespace bftq
