// ASAN/UBSAN fuzz driver for the host OpenPGP parser (bftkv_b200/csrc/pgp_host.hpp):
//   g++ -O1 -g -fsanitize=address,undefined -o /tmp/asan/fuzz tools/fuzz_pgp_host.cpp   (inputs: see DESIGN.md "Robustness")
#include "/root/repo/bftkv_b200/csrc/pgp_host.hpp"
#include <cstdio>
#include <random>
using namespace bftq::pgp;
int main() {
  std::vector<uint8_t> kb; { FILE* f = fopen("/tmp/asan/keys.bin", "rb"); int c; while ((c = fgetc(f)) != EOF) kb.push_back(c); fclose(f); }
  std::vector<Entity> ents; read_entities(kb.data(), kb.size(), ents);
  std::vector<std::vector<uint8_t>> sigs; { FILE* f = fopen("/tmp/asan/sigs.bin", "rb"); uint32_t l; while (fread(&l, 4, 1, f) == 1) { std::vector<uint8_t> s(l); fread(s.data(), 1, l, f); sigs.push_back(s); } fclose(f); }
  printf("ents %zu sigs %zu\n", ents.size(), sigs.size());
  std::mt19937 rng(1);
  std::vector<const std::vector<Entity>*> rings = {&ents};
  size_t calls = 0;
  for (int t = 0; t < 300000; t++) {
    std::vector<uint8_t> d;
    int n = rng() % 4;
    for (int i = 0; i < n; i++) { auto& s = sigs[rng() % sigs.size()]; d.insert(d.end(), s.begin(), s.end()); }
    if (rng() % 5 == 0) { size_t l = rng() % 500; d.insert(d.end(), kb.begin(), kb.begin() + l); }
    if (!d.empty()) { int m = rng() % 5; for (int i = 0; i < m; i++) d[rng() % d.size()] ^= 1 << (rng() % 8); if (rng() % 4 == 0) d.resize(rng() % d.size()); }
    // exact-size heap copy so ASAN sees any overread
    uint8_t* buf = new uint8_t[d.size() ? d.size() : 1];
    if (!d.empty()) memcpy(buf, d.data(), d.size());
    Reader r{buf, d.size(), 0};
    std::vector<uint8_t> scratch; std::vector<KeyRef> keys; SigPacket sp;
    while (r.remaining() > 0) { int rc = next_known_signature(r, rings, sp, keys, scratch); if (rc == kOk) calls++; else if (rc == kEof) break; }
    std::vector<Entity> e2; read_entities(buf, d.size(), e2);
    delete[] buf;
  }
  printf("ok calls=%zu\n", calls);
}
