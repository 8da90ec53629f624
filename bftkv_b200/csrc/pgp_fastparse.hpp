// Device-side fast path of the OpenPGP packer: the same parsing rules as pgp_host.hpp (read_packet +
// parse_signature + parse_subpackets), restricted to the shape almost every bftkv signature has — ONE
// definite-length v4 RSA signature packet that fills the whole SignaturePacket.Data — and written so that it
// compiles for the device (K0, pgp_parse.cuh) and for the host (tests/harness/fastparse_host.cpp fuzzes it
// against pgp_host.hpp).  Contract: kFast means "pgp_host.hpp parses this stream to exactly these fields";
// anything else — other versions, partial lengths, several packets, text mode, odd subpackets, anything the
// host parser would reject or skip — is kFallback and goes through the host packer unchanged.  A fallback is
// always safe; a wrong kFast is a parity bug, hence the fuzz harness.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define BFTQ_FP_HD __host__ __device__ __forceinline__
#else
#define BFTQ_FP_HD inline
#endif

namespace bftq { namespace fastparse {

enum : int { kFast = 0, kFallback = 1 };

struct FastSig {
  uint8_t sig_type, pk_algo, hash_id;
  uint16_t tag;                    // big-endian: first digest byte in the high half
  uint64_t issuer;
  uint32_t hashed_off, hashed_len; // bytes hashed after the data (version .. end of hashed area), offsets into the stream
  uint32_t mpi_off, mpi_len;       // RSA signature MPI bytes as stored
};

// pgp::parse_subpackets, same accept / reject decisions.
BFTQ_FP_HD int subpackets(const uint8_t* a, uint32_t n, bool hashed, bool& has_ctime, bool& has_issuer, uint64_t& issuer) {
  uint32_t p = 0;
  while (p < n) {
    uint32_t l;
    const uint8_t o = a[p];
    if (o < 192) { l = o; p += 1; }
    else if (o < 255) { if (p + 2 > n) return kFallback; l = ((uint32_t)(o - 192) << 8) + a[p + 1] + 192; p += 2; }
    else { if (p + 5 > n) return kFallback; l = ((uint32_t)a[p + 1] << 24) | ((uint32_t)a[p + 2] << 16) | ((uint32_t)a[p + 3] << 8) | a[p + 4]; p += 5; }
    if (l == 0 || l > n - p) return kFallback;
    const int typ = a[p] & 0x7f;
    const bool critical = (a[p] & 0x80) != 0;
    const uint8_t* sub = a + p + 1;
    const uint32_t sl = l - 1;
    p += l;
    switch (typ) {
      case 2: if (!hashed) break; if (sl != 4) return kFallback; has_ctime = true; break;
      case 3: case 9: if (hashed && sl != 4) return kFallback; break;
      case 16:
        if (sl != 8) return kFallback;
        has_issuer = true; issuer = 0;
        for (int i = 0; i < 8; i++) issuer = (issuer << 8) | sub[i];
        break;
      case 27: if (!hashed) break; if (sl == 0) return kFallback; break;
      case 25: if (hashed && sl != 1) return kFallback; break;
      case 29: if (hashed && sl == 0) return kFallback; break;
      case 11: case 21: case 22: case 30: case 32: break;
      default: if (critical) return kFallback; break;
    }
  }
  return kFast;
}

BFTQ_FP_HD int parse(const uint8_t* d, size_t n, FastSig& out) {
  if (n < 2 || n > 0x7fffffffu) return kFallback;
  const uint8_t hdr = d[0];
  if (!(hdr & 0x80)) return kFallback;
  uint32_t p, l; int tag;
  if (!(hdr & 0x40)) {                         // old format, definite length only
    tag = (hdr & 0x3f) >> 2;
    const int lt = hdr & 3;
    if (lt == 3) return kFallback;
    const uint32_t nl = 1u << lt;
    if (1 + nl > n) return kFallback;
    l = 0;
    for (uint32_t i = 0; i < nl; i++) l = (l << 8) | d[1 + i];
    p = 1 + nl;
  } else {                                     // new format, no partial lengths
    tag = hdr & 0x3f;
    const uint8_t o = d[1];
    if (o < 192) { l = o; p = 2; }
    else if (o < 224) { if (n < 3) return kFallback; l = ((uint32_t)(o - 192) << 8) + d[2] + 192; p = 3; }
    else if (o == 255) { if (n < 6) return kFallback; l = ((uint32_t)d[2] << 24) | ((uint32_t)d[3] << 16) | ((uint32_t)d[4] << 8) | d[5]; p = 6; }
    else return kFallback;
  }
  if (l != (uint32_t)n - p) return kFallback;  // the one packet is the whole stream
  if (tag != 2) return kFallback;
  const uint8_t* b = d + p;
  const uint32_t bn = l;
  if (bn < 6 || b[0] != 4) return kFallback;
  out.sig_type = b[1]; out.pk_algo = b[2]; out.hash_id = b[3];
  if (out.pk_algo != 1 && out.pk_algo != 3) return kFallback;
  switch (out.hash_id) { case 1: case 2: case 3: case 8: case 9: case 10: case 11: break; default: return kFallback; }   // pgp::hash_digest_len
  const uint32_t hl = ((uint32_t)b[4] << 8) | b[5];
  if (6 + hl + 2 > bn) return kFallback;
  bool has_ctime = false, has_issuer = false;
  uint64_t issuer = 0;
  if (subpackets(b + 6, hl, true, has_ctime, has_issuer, issuer)) return kFallback;
  uint32_t q = 6 + hl;
  const uint32_t ul = ((uint32_t)b[q] << 8) | b[q + 1];
  q += 2;
  if (q + ul + 2 > bn) return kFallback;
  if (subpackets(b + q, ul, false, has_ctime, has_issuer, issuer)) return kFallback;
  q += ul;
  if (!has_ctime || !has_issuer) return kFallback;
  out.tag = (uint16_t)(((uint16_t)b[q] << 8) | b[q + 1]);
  q += 2;
  if (q + 2 > bn) return kFallback;
  const uint32_t bits = ((uint32_t)b[q] << 8) | b[q + 1];
  const uint32_t len = (bits + 7) / 8;
  if (len > bn - q - 2) return kFallback;
  out.issuer = issuer;
  out.hashed_off = p; out.hashed_len = 6 + hl;
  out.mpi_off = p + q + 2; out.mpi_len = len;
  return kFast;
}

}}  // namespace bftq::fastparse
