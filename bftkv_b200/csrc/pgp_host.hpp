// Host-side OpenPGP packer for libbftq: turns what bftkv hands to crypto.Signature /
// crypto.CollectiveSignature (raw concatenated OpenPGP signature packets + the signed bytes +
// a keyring of OpenPGP public-key blocks) into flat (key index, signature, digest-suffix) tuples
// for the CUDA kernels, and folds the per-tuple results back into the reference's decisions.
//
// It restates ONLY parsing / bookkeeping — no hashing, no modular arithmetic happens here:
//   crypto/pgp/crypto_pgp.go:195-197  getKeyring = secring ++ keyring
//   crypto/pgp/crypto_pgp.go:319-344  Signature.Verify / VerifyWithCertificate loop semantics
//   crypto/pgp/crypto_pgp.go:373-390  Signature.Signers
//   crypto/pgp/crypto_pgp.go:485-515  CollectiveSignature.Verify / Combine
// plus the accept/reject rules of golang.org/x/crypto/openpgp @53104e6ec876 those lines rely on
// (packet framing, v4 signature packets and subpackets, EntityList.KeysByIdUsage,
// CheckDetachedSignature's unknown-issuer skipping), restated from RFC 4880 and the module's
// published behaviour — its source is not in the reference tree.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace bftq { namespace pgp {

enum : int { kOk = 0, kEof = 1, kStructural = 2, kUnsupported = 3 };
constexpr uint8_t kKeyFlagCertify = 0x01, kKeyFlagSign = 0x02;

inline bool known_tag(int t) {   // packet.Read's switch; everything else is UnknownPacketTypeError (skipped)
  switch (t) { case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 8: case 9: case 11: case 13: case 14: case 17: case 18: return true; }
  return false;
}
inline int hash_digest_len(int id) {
  switch (id) { case 1: return 16; case 2: return 20; case 3: return 20; case 8: return 32; case 9: return 48; case 10: return 64; case 11: return 28; }
  return 0;
}

struct Reader {
  const uint8_t* d; size_t len; size_t pos;
  size_t remaining() const { return len - pos; }
};

// One packet.  `body` points into the input unless the packet used partial lengths, in which case
// it points into `scratch`.  A truncated length or body exhausts the reader (x/crypto's readFull runs into
// io.ErrUnexpectedEOF with everything consumed); a tag byte without the MSB consumes exactly that ONE byte
// (packet.readHeader reads one byte and returns StructuralError), so CollectiveSignature.Verify's loop
// `for r.Len() > 0` (crypto_pgp.go:489) resynchronises on the next byte: a stray byte from one Byzantine
// responder must not hide the honest signatures behind it.
inline int read_packet(Reader& r, int& tag, const uint8_t*& body, size_t& blen, std::vector<uint8_t>& scratch) {
  if (r.pos >= r.len) return kEof;
  const uint8_t* d = r.d;
  const size_t n = r.len;
  const uint8_t hdr = d[r.pos];
  auto bad = [&]() { r.pos = n; return (int)kStructural; };
  if (!(hdr & 0x80)) { r.pos += 1; return (int)kStructural; }
  size_t p = r.pos + 1;
  if (!(hdr & 0x40)) {                       // old format
    tag = (hdr & 0x3f) >> 2;
    const int lt = hdr & 3;
    if (lt == 3) { body = d + p; blen = n - p; r.pos = n; return kOk; }
    const size_t nl = (size_t)1 << lt;
    if (p + nl > n) return bad();
    size_t l = 0;
    for (size_t i = 0; i < nl; i++) l = (l << 8) | d[p + i];
    p += nl;
    if (l > n - p) return bad();
    body = d + p; blen = l; r.pos = p + l;
    return kOk;
  }
  tag = hdr & 0x3f;
  bool used_partial = false;
  scratch.clear();
  for (;;) {
    if (p >= n) return bad();
    const uint8_t o = d[p];
    size_t l; bool partial = false;
    if (o < 192) { l = o; p += 1; }
    else if (o < 224) { if (p + 2 > n) return bad(); l = ((size_t)(o - 192) << 8) + d[p + 1] + 192; p += 2; }
    else if (o == 255) { if (p + 5 > n) return bad(); l = ((size_t)d[p + 1] << 24) | ((size_t)d[p + 2] << 16) | ((size_t)d[p + 3] << 8) | d[p + 4]; p += 5; }
    else { l = (size_t)1 << (o & 0x1f); p += 1; partial = true; }
    if (l > n - p) return bad();
    if (partial || used_partial) { scratch.insert(scratch.end(), d + p, d + p + l); used_partial = true; }
    if (!partial) {
      if (used_partial) { body = scratch.data(); blen = scratch.size(); }
      else { body = d + p; blen = l; }
      r.pos = p + l;
      return kOk;
    }
    p += l;
  }
}

// Tag of the packet the reader stands at == literal data (11)?  (peek only)
inline bool tag_is_literal(const Reader& r) {
  if (r.pos >= r.len) return false;
  const uint8_t hdr = r.d[r.pos];
  if (!(hdr & 0x80)) return false;
  return ((hdr & 0x40) ? (hdr & 0x3f) : ((hdr & 0x3f) >> 2)) == 11;
}

struct Span {
  const uint8_t* p = nullptr; size_t n = 0;
  const uint8_t* data() const { return p; }
  size_t size() const { return n; }
};

struct SigPacket {
  int version = 0;
  uint8_t sig_type = 0, pk_algo = 0, hash_id = 0;
  bool has_issuer = false; uint64_t issuer = 0;
  bool has_ctime = false;
  bool flags_valid = false, flag_certify = false, flag_sign = false;
  int is_primary_id = -1;                 // -1 absent
  bool has_revocation_reason = false;
  uint8_t hash_tag[2] = {0, 0};
  // Views into the packet body handed to parse_signature (valid until the reader moves on): the packer
  // parses tens of millions of these per second, so nothing here allocates.
  Span hashed;                            // bytes hashed after the data: v4 = version..end of hashed area, v3 = type + time
  uint8_t trailer[6] = {0, 0, 0, 0, 0, 0};
  uint8_t trailer_len = 0;                // v4: 04 FF len32 follows the hashed area; v3: nothing
  Span mpi;                               // RSA signature MPI bytes (as stored)
  Span r, s;                              // DSA / ECDSA signature MPIs, leading zeros stripped
  size_t suffix_size() const { return hashed.n + trailer_len; }
  void write_suffix(uint8_t* out) const { if (hashed.n) memcpy(out, hashed.p, hashed.n); memcpy(out + hashed.n, trailer, trailer_len); }
};

inline int parse_subpackets(const uint8_t* a, size_t n, SigPacket& s, bool hashed) {
  size_t p = 0;
  while (p < n) {
    size_t l;
    const uint8_t o = a[p];
    if (o < 192) { l = o; p += 1; }
    else if (o < 255) { if (p + 2 > n) return kStructural; l = ((size_t)(o - 192) << 8) + a[p + 1] + 192; p += 2; }
    else { if (p + 5 > n) return kStructural; l = ((size_t)a[p + 1] << 24) | ((size_t)a[p + 2] << 16) | ((size_t)a[p + 3] << 8) | a[p + 4]; p += 5; }
    if (l == 0 || l > n - p) return kStructural;
    const int typ = a[p] & 0x7f;
    const bool critical = a[p] & 0x80;
    const uint8_t* sub = a + p + 1;
    const size_t sl = l - 1;
    p += l;
    switch (typ) {
      case 2: if (!hashed) break; if (sl != 4) return kStructural; s.has_ctime = true; break;
      case 3: case 9: if (hashed && sl != 4) return kStructural; break;
      case 16: if (sl != 8) return kStructural; s.has_issuer = true; s.issuer = 0; for (int i = 0; i < 8; i++) s.issuer = (s.issuer << 8) | sub[i]; break;
      case 27: if (!hashed) break; if (sl == 0) return kStructural; s.flags_valid = true; s.flag_certify = sub[0] & kKeyFlagCertify; s.flag_sign = sub[0] & kKeyFlagSign; break;
      case 25: if (hashed) { if (sl != 1) return kStructural; s.is_primary_id = sub[0] > 0; } break;
      case 29: if (hashed) { if (sl == 0) return kStructural; s.has_revocation_reason = true; } break;
      case 11: case 21: case 22: case 30: case 32: break;
      default: if (critical) return kUnsupported; break;
    }
  }
  return kOk;
}

inline int read_mpi(const uint8_t* b, size_t n, size_t& p, const uint8_t*& data, size_t& len, unsigned& bits) {
  if (p + 2 > n) return kStructural;
  bits = ((unsigned)b[p] << 8) | b[p + 1];
  len = (bits + 7) / 8;
  if (len > n - p - 2) return kStructural;
  data = b + p + 2;
  p += 2 + len;
  return kOk;
}

// packet.Signature.parse for v4; v2/v3 are recognised (so the stream stays in sync) and marked.
inline int parse_signature(const uint8_t* b, size_t n, SigPacket& s) {
  if (n < 1) return kStructural;
  s = SigPacket();
  s.version = b[0];
  if (s.version < 4) {                       // SignatureV3.parse
    if (n < 19 || (b[0] != 2 && b[0] != 3) || b[1] != 5) return kUnsupported;
    s.version = 3; s.sig_type = b[2]; s.pk_algo = b[15]; s.hash_id = b[16];
    s.has_issuer = true; s.issuer = 0; for (int i = 0; i < 8; i++) s.issuer = (s.issuer << 8) | b[7 + i];
    s.has_ctime = true;
    if (s.pk_algo != 1 && s.pk_algo != 3 && s.pk_algo != 17) return kUnsupported;
    if (!hash_digest_len(s.hash_id)) return kUnsupported;
    s.hashed = Span{b + 2, 5};
    s.hash_tag[0] = b[17]; s.hash_tag[1] = b[18];
    size_t p = 19; const uint8_t* md; size_t ml; unsigned bits;
    if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
    if (s.pk_algo == 1 || s.pk_algo == 3) s.mpi = Span{md, ml};
    else {                                                         // DSA: r, s
      while (ml && *md == 0) { md++; ml--; }
      s.r = Span{md, ml};
      if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
      while (ml && *md == 0) { md++; ml--; }
      s.s = Span{md, ml};
    }
    return kOk;
  }
  if (s.version != 4) return kUnsupported;
  if (n < 6) return kStructural;
  s.sig_type = b[1]; s.pk_algo = b[2]; s.hash_id = b[3];
  if (s.pk_algo != 1 && s.pk_algo != 3 && s.pk_algo != 17 && s.pk_algo != 19) return kUnsupported;
  if (!hash_digest_len(s.hash_id)) return kUnsupported;
  const size_t hl = ((size_t)b[4] << 8) | b[5];
  if (6 + hl + 2 > n) return kStructural;
  const size_t l = 6 + hl;
  s.hashed = Span{b, l};
  s.trailer[0] = 0x04; s.trailer[1] = 0xff; s.trailer[2] = (uint8_t)(l >> 24); s.trailer[3] = (uint8_t)(l >> 16);
  s.trailer[4] = (uint8_t)(l >> 8); s.trailer[5] = (uint8_t)l;
  s.trailer_len = 6;
  int rc = parse_subpackets(b + 6, hl, s, true);
  if (rc) return rc;
  size_t p = 6 + hl;
  const size_t ul = ((size_t)b[p] << 8) | b[p + 1];
  p += 2;
  if (p + ul + 2 > n) return kStructural;
  rc = parse_subpackets(b + p, ul, s, false);
  if (rc) return rc;
  p += ul;
  if (!s.has_ctime) return kStructural;      // "no creation time in signature"
  s.hash_tag[0] = b[p]; s.hash_tag[1] = b[p + 1];
  p += 2;
  const uint8_t* md; size_t ml; unsigned bits;
  if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
  if (s.pk_algo == 1 || s.pk_algo == 3) s.mpi = Span{md, ml};
  else {                                                           // DSA/ECDSA: r and s must be well formed
    while (ml && *md == 0) { md++; ml--; }
    s.r = Span{md, ml};
    if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
    while (ml && *md == 0) { md++; ml--; }
    s.s = Span{md, ml};
  }
  return kOk;
}

struct PubKey {
  uint8_t algo = 0;
  uint64_t key_id = 0;
  std::vector<uint8_t> n_be;     // RSA modulus, stripped
  uint32_t e = 0;
  unsigned nbits = 0;
  int32_t table_idx = -1;        // index in the engine's key table, -1: not verifiable on the device
  std::vector<uint8_t> ec_xy;    // ECDSA on P-256: X || Y (64 bytes); empty for other curves
  std::vector<uint8_t> dsa[4];   // DSA: p, q, g, y (stripped big-endian)
};
inline void strip_assign(std::vector<uint8_t>& out, const uint8_t* md, size_t ml) {
  while (ml && *md == 0) { md++; ml--; }
  out.assign(md, md + ml);
}

// SHA-1 only for key ids (fingerprint of the public-key packet) — identification, not verification.
inline void sha1(const uint8_t* m, size_t len, uint8_t out[20]) {
  uint32_t h[5] = {0x67452301, 0xEFCDAB89, 0x98BADCFE, 0x10325476, 0xC3D2E1F0};
  std::vector<uint8_t> buf(m, m + len);
  buf.push_back(0x80);
  while (buf.size() % 64 != 56) buf.push_back(0);
  const uint64_t bits = (uint64_t)len * 8;
  for (int i = 7; i >= 0; i--) buf.push_back((uint8_t)(bits >> (8 * i)));
  for (size_t o = 0; o < buf.size(); o += 64) {
    uint32_t w[80];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)buf[o + 4 * i] << 24) | ((uint32_t)buf[o + 4 * i + 1] << 16) | ((uint32_t)buf[o + 4 * i + 2] << 8) | buf[o + 4 * i + 3];
    for (int i = 16; i < 80; i++) { uint32_t x = w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16]; w[i] = (x << 1) | (x >> 31); }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    for (int i = 0; i < 80; i++) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDC; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6; }
      const uint32_t t = ((a << 5) | (a >> 27)) + f + e + k + w[i];
      e = d; d = c; c = (b << 30) | (b >> 2); b = a; a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
  }
  for (int i = 0; i < 5; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}

inline int parse_public_key(const uint8_t* b, size_t n, PubKey& k) {
  if (n < 6 || b[0] != 4) return kUnsupported;
  k = PubKey();
  k.algo = b[5];
  if (k.algo == 1 || k.algo == 2 || k.algo == 3) {
    size_t p = 6; const uint8_t* md; size_t ml; unsigned bits;
    if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
    while (ml && *md == 0) { md++; ml--; }
    k.n_be.assign(md, md + ml);
    k.nbits = ml ? (unsigned)(8 * (ml - 1) + (32 - __builtin_clz((unsigned)md[0]))) : 0;
    if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
    if (ml > 3) return kUnsupported;          // "large public exponent"
    k.e = 0; for (size_t i = 0; i < ml; i++) k.e = (k.e << 8) | md[i];
  } else if (k.algo == 19) {                     // parseECDSA: OID then the uncompressed point as one MPI
    if (n < 7) return kStructural;
    const size_t ol = b[6];
    if (ol == 0 || ol == 0xff || 7 + ol > n) return kStructural;
    static const uint8_t p256_oid[8] = {0x2A, 0x86, 0x48, 0xCE, 0x3D, 0x03, 0x01, 0x07};
    const bool p256 = ol == 8 && memcmp(b + 7, p256_oid, 8) == 0;
    size_t p = 7 + ol; const uint8_t* md; size_t ml; unsigned bits;
    if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
    if (p256 && ml == 65 && md[0] == 4) k.ec_xy.assign(md + 1, md + 65);
    k.nbits = 256;
  } else if (k.algo == 17) {                     // parseDSA: p, q, g, y
    size_t p = 6; const uint8_t* md; size_t ml; unsigned bits;
    for (int i = 0; i < 4; i++) {
      if (read_mpi(b, n, p, md, ml, bits)) return kStructural;
      strip_assign(k.dsa[i], md, ml);
    }
    const auto& P = k.dsa[0];
    k.nbits = P.empty() ? 0 : (unsigned)(8 * (P.size() - 1) + (32 - __builtin_clz((unsigned)P[0])));
  } else if (k.algo != 16 && k.algo != 18) {
    return kUnsupported;
  }
  std::vector<uint8_t> fp(3 + n);
  fp[0] = 0x99; fp[1] = (uint8_t)(n >> 8); fp[2] = (uint8_t)n;
  memcpy(fp.data() + 3, b, n);
  uint8_t dg[20];
  sha1(fp.data(), fp.size(), dg);
  k.key_id = 0; for (int i = 12; i < 20; i++) k.key_id = (k.key_id << 8) | dg[i];
  return kOk;
}

struct SelfSig { bool present = false, flags_valid = false, flag_sign = false, flag_certify = false, revocation_reason = false; };
struct Subkey { PubKey key; SelfSig sig; };
struct Entity {
  PubKey primary;
  SelfSig self;
  bool revoked = false;
  std::vector<Subkey> subkeys;
  std::vector<uint64_t> certifiers;      // issuers of third-party certifications (crypto_pgp.go:80-88)
};

inline SelfSig to_selfsig(const SigPacket& s) {
  SelfSig r; r.present = true; r.flags_valid = s.flags_valid; r.flag_sign = s.flag_sign; r.flag_certify = s.flag_certify;
  r.revocation_reason = s.has_revocation_reason; return r;
}

// Groups a serialized key block into entities (openpgp.ReadKeyRing / ReadEntity structure; the
// self-signature crypto checks x/crypto performs at LOAD time are not on the per-signature path).
inline void read_entities(const uint8_t* d, size_t n, std::vector<Entity>& out) {
  Reader r{d, n, 0};
  std::vector<uint8_t> scratch;
  int cur = -1; int last = 0; /* 0 none, 1 key, 2 uid, 3 subkey */ int uid_count = 0; bool have_self = false;
  for (;;) {
    int tag; const uint8_t* body; size_t bl;
    const int rc = read_packet(r, tag, body, bl, scratch);
    if (rc) break;
    if (tag == 6) {
      Entity e;
      if (parse_public_key(body, bl, e.primary) == kOk) { out.push_back(e); cur = (int)out.size() - 1; last = 1; uid_count = 0; have_self = false; }
      else { cur = -1; last = 0; }
    } else if (cur < 0) {
      continue;
    } else if (tag == 13) { last = 2; uid_count++; }
    else if (tag == 14) {
      Subkey sk;
      if (parse_public_key(body, bl, sk.key) == kOk) { out[cur].subkeys.push_back(sk); last = 3; } else last = 0;
    } else if (tag == 2) {
      SigPacket s;
      if (parse_signature(body, bl, s) != kOk || last == 0) continue;
      Entity& e = out[cur];
      if (last == 1) { if (s.sig_type == 0x20) e.revoked = true; }
      else if (last == 2) {
        if (s.sig_type >= 0x10 && s.sig_type <= 0x13) {
          if (s.has_issuer && s.issuer == e.primary.key_id) {
            if (!have_self || (s.is_primary_id == 1 && uid_count > 1)) { e.self = to_selfsig(s); have_self = true; }
          } else if (s.has_issuer) e.certifiers.push_back(s.issuer);
        }
      } else if (last == 3) { if (s.sig_type == 0x18) e.subkeys.back().sig = to_selfsig(s); }
    }
  }
}

struct KeyRef { const Entity* entity; const PubKey* key; };

// EntityList.KeysByIdUsage(id, KeyFlagSign) over secring ++ keyring.
inline void keys_by_id_usage(const std::vector<const std::vector<Entity>*>& rings, uint64_t id, uint8_t usage, std::vector<KeyRef>& out) {
  out.clear();
  for (auto* ring : rings)
    for (const Entity& e : *ring) {
      auto consider = [&](const PubKey& k, const SelfSig& ss) {
        if (e.revoked) return;
        if (ss.present && ss.revocation_reason) return;
        if (ss.present && ss.flags_valid && usage) {
          const uint8_t have = (ss.flag_certify ? kKeyFlagCertify : 0) | (ss.flag_sign ? kKeyFlagSign : 0);
          if ((have & usage) != usage) return;
        }
        out.push_back(KeyRef{&e, &k});
      };
      if (e.primary.key_id == id) consider(e.primary, e.self);
      for (const Subkey& sk : e.subkeys) if (sk.key.key_id == id) consider(sk.key, sk.sig);
    }
}

// One openpgp.CheckDetachedSignature call's worth of parsing on a shared reader.
// kOk: `sig` holds the first signature packet with a known issuer, `keys` its candidate keys.
// kEof: ErrUnknownIssuer (stream ended).  Otherwise a structural / unsupported error (that packet
// has been consumed).
inline int next_known_signature(Reader& r, const std::vector<const std::vector<Entity>*>& rings, SigPacket& sig,
                                std::vector<KeyRef>& keys, std::vector<uint8_t>& scratch) {
  for (;;) {
    int tag; const uint8_t* body; size_t bl;
    for (;;) {
      const int rc = read_packet(r, tag, body, bl, scratch);
      if (rc) return rc;
      if (known_tag(tag)) break;
    }
    if (tag != 2) return kStructural;                    // "non signature packet found"
    const int rc = parse_signature(body, bl, sig);
    if (rc) return rc;
    if (!sig.has_issuer) return kStructural;             // "signature doesn't have an issuer"
    keys_by_id_usage(rings, sig.issuer, kKeyFlagSign, keys);
    if (!keys.empty()) return kOk;
  }
}

}}  // namespace bftq::pgp

// ---- transport messages: PGPMessage.Decrypt's signature half (crypto_pgp.go:453-471) -------------------------------
// openpgp.ReadMessage's readSignedMessage walk over the packet stream a SymmetricallyEncrypted packet decrypts to:
// [compressed] one-pass signature (tag 4), literal data (tag 11), signature (tag 2).  Restated from the published
// golang.org/x/crypto/openpgp/read.go @53104e6ec876 (not in the reference tree); oracle/pgp_oracle.py message_verify
// is the same walk in Python.
namespace bftq { namespace pgp {

enum : int { kMsgOk = 0, kMsgReadFailed = 1 /* ReadMessage error -> ErrDecryptionFailed */, kMsgNotSigned = 2 /* ErrInvalidTransportSecurityData */,
              kMsgCompressed = 3 /* compressed data packet: not handled here */, kMsgBodyFailed = 4 /* the literal body ends early: ReadAll's error */ };

struct MessageHead {
  bool has_ops = false;
  uint8_t ops_sig_type = 0, ops_hash = 0, ops_pk_algo = 0;
  uint64_t ops_key_id = 0;
  bool binary = false;
  const uint8_t* name = nullptr; size_t name_len = 0;       // literal FileName (views into the packet body)
  const uint8_t* body = nullptr; size_t body_len = 0;       // literal data, de-chunked
};

// Walks to the literal data packet.  On kMsgOk the reader stands behind the literal packet; `scratch` owns the body when
// the packet used partial lengths (what Go's own writer always emits), so it must outlive `h`.
inline int read_message_head(Reader& r, MessageHead& h, std::vector<uint8_t>& scratch) {
  h = MessageHead();
  std::vector<uint8_t> tmp;
  for (;;) {
    int tag; const uint8_t* body; size_t bl;
    const bool at_literal = tag_is_literal(r);
    const int rc = read_packet(r, tag, body, bl, at_literal ? scratch : tmp);
    // a literal packet that ends early is discovered by ioutil.ReadAll(UnverifiedBody), after the IsSigned check
    if (rc) return rc != kEof && at_literal ? (h.has_ops ? kMsgBodyFailed : kMsgNotSigned) : kMsgReadFailed;   // EOF: packets.Next() -> io.EOF -> ReadMessage fails
    if (!known_tag(tag)) continue;
    if (tag == 8) return kMsgCompressed;                     // compressed data: not handled (bftkv's own Encrypt never compresses)
    if (tag == 4) {
      if (bl < 13) return kMsgReadFailed;
      if (body[0] != 3) return kMsgReadFailed;               // UnsupportedError("one-pass-signature packet version")
      if (!hash_digest_len(body[2])) return kMsgReadFailed;  // UnsupportedError("hash function")
      if (!body[12]) return kMsgReadFailed;                  // UnsupportedError("nested signatures")
      if (body[2] == 3) return kMsgReadFailed;               // hashForSignature: crypto.RIPEMD160 is not linked into bftkv
      if (body[1] != 0x00 && body[1] != 0x01) return kMsgReadFailed;   // hashForSignature: unsupported signature type
      h.has_ops = true; h.ops_sig_type = body[1]; h.ops_hash = body[2]; h.ops_pk_algo = body[3];
      h.ops_key_id = 0; for (int i = 0; i < 8; i++) h.ops_key_id = (h.ops_key_id << 8) | body[4 + i];
    } else if (tag == 2) {
      SigPacket s;
      if (parse_signature(body, bl, s)) return kMsgReadFailed;          // parsed by packet.Read, ignored by the switch
    } else if (tag == 11) {
      if (bl < 2 || bl < (size_t)2 + body[1] + 4) return kMsgReadFailed;
      h.binary = body[0] == 'b';
      h.name = body + 2; h.name_len = body[1];
      h.body = body + 6 + body[1]; h.body_len = bl - 6 - body[1];
      return h.has_ops ? kMsgOk : kMsgNotSigned;
    }
  }
}

// encoding/base64 StdEncoding.DecodeString (Go 1.13): CR / LF skipped anywhere, padding mandatory, nothing but CR / LF
// after it, trailing bits unchecked.  Returns false on CorruptInputError.  out may hold up to 3 * n / 4 bytes.
inline bool go_base64_std_decode(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  out.clear();
  auto val = [](uint8_t c) -> int {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
  };
  uint8_t q[4] = {0, 0, 0, 0}; int nq = 0; bool closed = false;
  size_t i = 0;
  auto flush = [&](int have) { // have = data characters of the quantum (2..4)
    const int v0 = val(q[0]), v1 = val(q[1]), v2 = have > 2 ? val(q[2]) : 0, v3 = have > 3 ? val(q[3]) : 0;
    out.push_back((uint8_t)((v0 << 2) | (v1 >> 4)));
    if (have > 2) out.push_back((uint8_t)(((v1 & 15) << 4) | (v2 >> 2)));
    if (have > 3) out.push_back((uint8_t)(((v2 & 3) << 6) | v3));
  };
  for (; i < n; i++) {
    const uint8_t c = src[i];
    if (c == '\r' || c == '\n') continue;
    if (closed) return false;                                  // data after the padding
    if (c == '=') {
      if (nq < 2) return false;
      if (nq == 2) {                                           // 'xx==' : the next significant byte must be '='
        size_t j = i + 1;
        while (j < n && (src[j] == '\r' || src[j] == '\n')) j++;
        if (j >= n || src[j] != '=') return false;
        i = j;
      }
      flush(nq);
      nq = 0; closed = true;
      continue;
    }
    if (val(c) < 0) return false;
    q[nq++] = c;
    if (nq == 4) { flush(4); nq = 0; }
  }
  return nq == 0;
}

}}  // namespace bftq::pgp
