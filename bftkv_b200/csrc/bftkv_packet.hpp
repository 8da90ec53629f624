// bftkv's own packet format <x, v, t, sig, ss, auth> (packet/packet.go:35-115,192-235) as far as Client.processResponse
// (protocol/client.go:207-230) needs it: does packet.Parse fail, and if not, what are the value bytes and the timestamp
// the response is bucketed by.  Written so that it compiles for the device (the read-path kernels) and for the host
// (the fallback packer); oracle/packet_oracle.py parse() is the same walk.
//
// Go's readers distinguish a clean end (io.EOF: zero bytes left when a read starts) from a short read
// (io.ErrUnexpectedEOF).  packet.Parse forgives io.EOF at every field after the variable — including the sub-fields of a
// signature — and stops there; a short read, or any error while reading the variable, fails the packet.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define BFTQ_PK_HD __host__ __device__ __forceinline__
#else
#define BFTQ_PK_HD inline
#endif

namespace bftq { namespace pkt {

struct View {
  bool err;               // packet.Parse returned an error
  uint64_t t;             // timestamp (0 when the packet ends before it)
  uint32_t value_off, value_len;   // the value chunk inside the packet (len 0: nil / empty, both bucket as "")
};

enum : int { kGot = 0, kEof = 1, kShort = 2 };

// binary.Read of `width` (<= 8) big-endian bytes
BFTQ_PK_HD int rd(const uint8_t* p, uint64_t n, uint64_t& pos, int width, uint64_t& v) {
  const uint64_t avail = n - pos;
  if (avail == 0) return kEof;
  if (avail < (uint64_t)width) return kShort;
  v = 0;
  for (int i = 0; i < width; i++) v = (v << 8) | p[pos + i];
  pos += (uint64_t)width;
  return kGot;
}
// packet.ReadChunk: u64 length, then that many bytes (io.ReadFull: io.EOF when nothing is left, short otherwise)
BFTQ_PK_HD int chunk(const uint8_t* p, uint64_t n, uint64_t& pos, uint64_t& off, uint64_t& len) {
  uint64_t l;
  const int rc = rd(p, n, pos, 8, l);
  if (rc) return rc;
  off = pos; len = 0;
  if (l == 0) return kGot;
  const uint64_t avail = n - pos;
  if (avail == 0) return kEof;
  if (avail < l) return kShort;            // (a length beyond any allocation makes the reference panic; here it fails)
  len = l; pos += l;
  return kGot;
}
// packet.readSignature
BFTQ_PK_HD int signature(const uint8_t* p, uint64_t n, uint64_t& pos) {
  uint64_t v, o, l;
  int rc;
  if ((rc = rd(p, n, pos, 1, v))) return rc;       // Type
  if ((rc = rd(p, n, pos, 4, v))) return rc;       // Version
  if ((rc = rd(p, n, pos, 1, v))) return rc;       // Completed
  if ((rc = chunk(p, n, pos, o, l))) return rc;    // Data
  return chunk(p, n, pos, o, l);                   // Cert
}

// packet.Parse over p[0 .. n), n > 0 (processResponse does not parse an empty answer: it buckets ("", 0)).
BFTQ_PK_HD View parse(const uint8_t* p, uint64_t n) {
  View out{false, 0, 0, 0};
  uint64_t pos = 0, off = 0, len = 0, t = 0;
  if (chunk(p, n, pos, off, len) != kGot) { out.err = true; return out; }      // variable: every error counts, io.EOF too
  int rc = chunk(p, n, pos, off, len);                                           // value
  if (rc == kShort) { out.err = true; return out; }
  if (rc == kEof) return out;
  out.value_off = (uint32_t)off; out.value_len = (uint32_t)len;
  rc = rd(p, n, pos, 8, t);                                                      // timestamp
  if (rc == kShort) { out.err = true; return out; }
  if (rc == kEof) return out;
  out.t = t;
  for (int s = 0; s < 2; s++) {                                                  // sig, ss
    rc = signature(p, n, pos);
    if (rc == kShort) { out.err = true; return out; }
    if (rc == kEof) return out;
  }
  rc = chunk(p, n, pos, off, len);                                               // auth
  if (rc == kShort) out.err = true;
  return out;
}

}}  // namespace bftq::pkt
