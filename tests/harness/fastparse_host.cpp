// Fuzz harness (host build) for bftkv_b200/csrc/pgp_fastparse.hpp: whenever the fast parser says kFast, the
// reference-shaped host parser (pgp_host.hpp: read_packet + parse_signature) must parse the same stream to the
// same fields, with the packet filling the stream.  Input: a file of length-prefixed seed streams.  Output:
// "<streams tried> <fast> <mismatches>".
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../bftkv_b200/csrc/pgp_fastparse.hpp"
#include "../../bftkv_b200/csrc/pgp_host.hpp"

using namespace bftq;

static bool agrees(const std::vector<uint8_t>& s) {
  fastparse::FastSig f;
  if (fastparse::parse(s.data(), s.size(), f) != fastparse::kFast) return true;     // a fallback is always fine
  pgp::Reader r{s.data(), s.size(), 0};
  std::vector<uint8_t> scratch;
  int tag; const uint8_t* body; size_t bl;
  if (pgp::read_packet(r, tag, body, bl, scratch) != pgp::kOk) return false;
  if (!pgp::known_tag(tag) || tag != 2 || r.remaining() != 0) return false;
  pgp::SigPacket sp;
  if (pgp::parse_signature(body, bl, sp) != pgp::kOk) return false;
  if (sp.version != 4 || !sp.has_issuer) return false;
  if (sp.sig_type != f.sig_type || sp.pk_algo != f.pk_algo || sp.hash_id != f.hash_id || sp.issuer != f.issuer) return false;
  if (((sp.hash_tag[0] << 8) | sp.hash_tag[1]) != f.tag) return false;
  if (sp.hashed.p != s.data() + f.hashed_off || sp.hashed.n != f.hashed_len || sp.trailer_len != 6) return false;
  if (sp.mpi.p != s.data() + f.mpi_off || sp.mpi.n != f.mpi_len) return false;
  const size_t l = f.hashed_len;
  const uint8_t tr[6] = {0x04, 0xff, (uint8_t)(l >> 24), (uint8_t)(l >> 16), (uint8_t)(l >> 8), (uint8_t)l};
  for (int i = 0; i < 6; i++) if (sp.trailer[i] != tr[i]) return false;
  return true;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const long iters = atol(argv[2]);
  std::vector<std::vector<uint8_t>> seeds;
  for (;;) {
    uint8_t h[4];
    if (fread(h, 1, 4, f) != 4) break;
    const size_t n = ((size_t)h[0] << 24) | (h[1] << 16) | (h[2] << 8) | h[3];
    std::vector<uint8_t> s(n);
    if (n && fread(s.data(), 1, n, f) != n) break;
    seeds.push_back(s);
  }
  fclose(f);
  std::mt19937_64 rng(0xBF7C);
  long tried = 0, fast = 0, bad = 0;
  auto test = [&](const std::vector<uint8_t>& s) {
    tried++;
    fastparse::FastSig fs;
    if (fastparse::parse(s.data(), s.size(), fs) == fastparse::kFast) fast++;
    if (!agrees(s)) { bad++; if (bad < 5) { fprintf(stderr, "mismatch on:"); for (uint8_t c : s) fprintf(stderr, " %02x", c); fprintf(stderr, "\n"); } }
  };
  for (auto& s : seeds) test(s);
  for (long it = 0; it < iters && !seeds.empty(); it++) {
    std::vector<uint8_t> s = seeds[rng() % seeds.size()];
    const int kind = rng() % 8;
    if (kind == 0 && !s.empty()) s.resize(rng() % s.size());                                   // truncate
    else if (kind == 1) { const auto& o = seeds[rng() % seeds.size()]; s.insert(s.end(), o.begin(), o.end()); }   // two packets
    else if (kind == 2) s.push_back((uint8_t)rng());                                                // trailing byte
    else {
      const int flips = 1 + rng() % 3;
      for (int k = 0; k < flips && !s.empty(); k++) {
        // bias towards the header and the subpacket areas, where the structure lives
        const size_t pos = (rng() % 3) ? rng() % std::min<size_t>(s.size(), 40) : rng() % s.size();
        if (rng() % 2) s[pos] ^= (uint8_t)(1u << (rng() % 8)); else s[pos] = (uint8_t)rng();
      }
    }
    test(s);
  }
  printf("%ld %ld %ld\n", tried, fast, bad);
  return bad ? 1 : 0;
}
