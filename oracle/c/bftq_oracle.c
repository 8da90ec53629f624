/*
 * bftq_oracle.c — CPU restatement of the bftkv signature-verify + quorum-tally arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Not product code: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, as the checker or as
 * the timed CPU baseline.  libbftq.so never links or calls it.
 *
 * What it follows (file:line relative to the yahoo/bftkv reference tree; the arithmetic itself
 * lives in un-vendored golang.org/x/crypto/openpgp @ v0.0.0-20191227163750-53104e6ec876 and the
 * Go 1.13 standard library, so it is restated from their published behaviour):
 *   orc_rsa_verify_*    crypto/pgp/crypto_pgp.go:324,338,490 -> openpgp.CheckDetachedSignature
 *                       -> packet.PublicKey.VerifySignature -> rsa.VerifyPKCS1v15:
 *                         c = big.Int(sig); m = c^e mod n (big.Int.Exp, no range check on c);
 *                         em = leftPad(m, k); em[0]==0, em[1]==1, em[2..k-tLen-2]==0xff,
 *                         em[k-tLen-1]==0, em[k-tLen..]==prefix||hashed.
 *                       DigestInfo prefixes: crypto/threshold/rsa/rsa.go:345-354 (copy of Go's).
 *   orc_tally_*         quorum/wotqs/wotqs.go:144-206 (IsQuorum/IsThreshold/IsSufficient/Reject,
 *                       intersection counting duplicates of the input list).
 *   orc_sha256          FIPS 180-4 (what x/crypto's hashForSignature instantiates for hash id 8).
 *
 * Big numbers: own 64-bit-limb Montgomery code (no OpenSSL/GMP), one pthread per requested core.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define L 32 /* 2048 bits in 64-bit limbs */

/* ------------------------------------------------------------------ SHA-256 ---------------- */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha256_block(uint32_t h[8], const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = hh + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
    uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
/* SHA-256 of the concatenation of up to 3 segments (tbs || hashed-suffix || trailer). */
void orc_sha256_3(const uint8_t* a, size_t al, const uint8_t* b, size_t bl, const uint8_t* c, size_t cl, uint8_t out[32]) {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint8_t buf[64];
  size_t fill = 0;
  uint64_t total = (uint64_t)al + bl + cl;
  const uint8_t* segs[3] = {a, b, c};
  size_t lens[3] = {al, bl, cl};
  for (int s = 0; s < 3; s++)
    for (size_t i = 0; i < lens[s]; i++) {
      buf[fill++] = segs[s][i];
      if (fill == 64) { sha256_block(h, buf); fill = 0; }
    }
  buf[fill++] = 0x80;
  if (fill > 56) { memset(buf + fill, 0, 64 - fill); sha256_block(h, buf); fill = 0; }
  memset(buf + fill, 0, 56 - fill);
  uint64_t bits = total * 8;
  for (int i = 0; i < 8; i++) buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
  sha256_block(h, buf);
  for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}
void orc_sha256(const uint8_t* m, size_t len, uint8_t out[32]) { orc_sha256_3(m, len, 0, 0, 0, 0, out); }

/* ------------------------------------------------------------------ 2048-bit Montgomery ---- */
typedef struct { uint64_t n[L], r2[L], one[L], n0inv; } mont_ctx;

static int geq(const uint64_t* a, const uint64_t* b) {
  for (int i = L - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
  return 1;
}
static void subn(uint64_t* a, const uint64_t* b) {
  uint64_t br = 0;
  for (int i = 0; i < L; i++) { u128 d = (u128)a[i] - b[i] - br; a[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
static void from_be(uint64_t* a, const uint8_t* be) {
  for (int i = 0; i < L; i++) { uint64_t v = 0; for (int b = 0; b < 8; b++) v = (v << 8) | be[256 - 8 * (i + 1) + b]; a[i] = v; }
}
static void to_be(const uint64_t* a, uint8_t* be) {
  for (int i = 0; i < L; i++) for (int b = 0; b < 8; b++) be[256 - 8 * (i + 1) + b] = (uint8_t)(a[i] >> (56 - 8 * b));
}
/* r = a*b*2^-2048 mod n (CIOS), inputs < n, output < n */
static void montmul(uint64_t* r, const uint64_t* a, const uint64_t* b, const mont_ctx* c) {
  uint64_t t[L + 2];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < L; i++) {
    u128 carry = 0;
    for (int j = 0; j < L; j++) { u128 v = (u128)a[j] * b[i] + t[j] + (uint64_t)carry; t[j] = (uint64_t)v; carry = v >> 64; }
    u128 v = (u128)t[L] + (uint64_t)carry; t[L] = (uint64_t)v; t[L + 1] = (uint64_t)(v >> 64);
    uint64_t m = t[0] * c->n0inv;
    v = (u128)m * c->n[0] + t[0]; carry = v >> 64;
    for (int j = 1; j < L; j++) { v = (u128)m * c->n[j] + t[j] + (uint64_t)carry; t[j - 1] = (uint64_t)v; carry = v >> 64; }
    v = (u128)t[L] + (uint64_t)carry; t[L - 1] = (uint64_t)v; t[L] = t[L + 1] + (uint64_t)(v >> 64);
  }
  if (t[L] || geq(t, c->n)) subn(t, c->n);
  memcpy(r, t, L * 8);
}
static int mont_init(mont_ctx* c, const uint8_t* n_be) {
  from_be(c->n, n_be);
  if (!(c->n[0] & 1)) return -1;
  uint64_t inv = c->n[0];
  for (int i = 0; i < 6; i++) inv *= 2 - c->n[0] * inv;
  c->n0inv = 0 - inv;
  /* x = 2^2047 mod n, then double up to 2^2048 (=R mod n) and 2^4096 (=R^2 mod n) */
  uint64_t x[L];
  memset(x, 0, sizeof(x));
  x[L - 1] = 1ull << 63;
  int top = 0;
  for (int i = L - 1; i >= 0 && !top; i--) if (c->n[i]) top = 1;
  if (!top) return -1;
  while (geq(x, c->n)) subn(x, c->n);
  for (int e = 2047; e < 4096; e++) {
    uint64_t hi = x[L - 1] >> 63;
    for (int i = L - 1; i > 0; i--) x[i] = (x[i] << 1) | (x[i - 1] >> 63);
    x[0] <<= 1;
    if (hi || geq(x, c->n)) subn(x, c->n);
    if (e + 1 == 2048) memcpy(c->one, x, sizeof(x));
  }
  memcpy(c->r2, x, sizeof(x));
  return 0;
}
/* out = base^e mod n, base any 2048-bit value (reduced first, as big.Int.Exp does implicitly) */
static void modexp_u32(uint64_t* out, const uint64_t* base, uint32_t e, const mont_ctx* c) {
  uint64_t b[L], bm[L], y[L], onep[L];
  memcpy(b, base, sizeof(b));
  while (geq(b, c->n)) subn(b, c->n);
  montmul(bm, b, c->r2, c);
  if (e == 0) {
    memcpy(y, c->one, sizeof(y));
  } else {
    /* left-to-right square and multiply from the top set bit (17 bits for e = 65537) */
    int top = 31;
    while (!((e >> top) & 1)) top--;
    memcpy(y, bm, sizeof(y));
    for (int bit = top - 1; bit >= 0; bit--) {
      montmul(y, y, y, c);
      if ((e >> bit) & 1) montmul(y, y, bm, c);
    }
  }
  memset(onep, 0, sizeof(onep));
  onep[0] = 1;
  montmul(out, y, onep, c);
}

static const uint8_t PFX_MD5[] = {0x30, 0x20, 0x30, 0x0c, 0x06, 0x08, 0x2a, 0x86, 0x48, 0x86, 0xf7, 0x0d, 0x02, 0x05, 0x05, 0x00, 0x04, 0x10};
static const uint8_t PFX_SHA1[] = {0x30, 0x21, 0x30, 0x09, 0x06, 0x05, 0x2b, 0x0e, 0x03, 0x02, 0x1a, 0x05, 0x00, 0x04, 0x14};
static const uint8_t PFX_RMD160[] = {0x30, 0x20, 0x30, 0x08, 0x06, 0x06, 0x28, 0xcf, 0x06, 0x03, 0x00, 0x31, 0x04, 0x14};
static const uint8_t PFX_SHA224[] = {0x30, 0x2d, 0x30, 0x0d, 0x06, 0x09, 0x60, 0x86, 0x48, 0x01, 0x65, 0x03, 0x04, 0x02, 0x04, 0x05, 0x00, 0x04, 0x1c};
static const uint8_t PFX_SHA256[] = {0x30, 0x31, 0x30, 0x0d, 0x06, 0x09, 0x60, 0x86, 0x48, 0x01, 0x65, 0x03, 0x04, 0x02, 0x01, 0x05, 0x00, 0x04, 0x20};
static const uint8_t PFX_SHA384[] = {0x30, 0x41, 0x30, 0x0d, 0x06, 0x09, 0x60, 0x86, 0x48, 0x01, 0x65, 0x03, 0x04, 0x02, 0x02, 0x05, 0x00, 0x04, 0x30};
static const uint8_t PFX_SHA512[] = {0x30, 0x51, 0x30, 0x0d, 0x06, 0x09, 0x60, 0x86, 0x48, 0x01, 0x65, 0x03, 0x04, 0x02, 0x03, 0x05, 0x00, 0x04, 0x40};
static int hash_info(uint32_t id, const uint8_t** pfx, int* plen, int* dlen) {
  switch (id) {
    case 1: *pfx = PFX_MD5; *plen = sizeof(PFX_MD5); *dlen = 16; return 0;
    case 2: *pfx = PFX_SHA1; *plen = sizeof(PFX_SHA1); *dlen = 20; return 0;
    case 3: *pfx = PFX_RMD160; *plen = sizeof(PFX_RMD160); *dlen = 20; return 0;
    case 8: *pfx = PFX_SHA256; *plen = sizeof(PFX_SHA256); *dlen = 32; return 0;
    case 9: *pfx = PFX_SHA384; *plen = sizeof(PFX_SHA384); *dlen = 48; return 0;
    case 10: *pfx = PFX_SHA512; *plen = sizeof(PFX_SHA512); *dlen = 64; return 0;
    case 11: *pfx = PFX_SHA224; *plen = sizeof(PFX_SHA224); *dlen = 28; return 0;
  }
  return -1;
}
int orc_hash_dlen(uint32_t id) { const uint8_t* p; int pl, dl; return hash_info(id, &p, &pl, &dl) ? 0 : dl; }

/* rsa.VerifyPKCS1v15 for k = 256.  Returns 0 = nil error, 1 = ErrVerification.
 * strict_range != 0 additionally rejects s >= n (Go >= 1.20 behaviour; not the pinned Go 1.13). */
static int verify_one(const mont_ctx* c, uint32_t e, const uint8_t* sig_be, const uint8_t* digest, uint32_t hash_alg, int strict_range) {
  const uint8_t* pfx; int plen, dlen;
  if (hash_info(hash_alg, &pfx, &plen, &dlen)) return 1;
  const int k = 256, tlen = plen + dlen;
  if (k < tlen + 11) return 1;
  uint64_t s[L], m[L];
  from_be(s, sig_be);
  if (strict_range && geq(s, c->n)) return 1;
  modexp_u32(m, s, e, c);
  uint8_t em[256];
  to_be(m, em);
  int ok = em[0] == 0;
  ok &= em[1] == 1;
  ok &= memcmp(em + k - dlen, digest, dlen) == 0;
  ok &= memcmp(em + k - tlen, pfx, plen) == 0;
  ok &= em[k - tlen - 1] == 0;
  for (int i = 2; i < k - tlen - 1; i++) ok &= em[i] == 0xff;
  return ok ? 0 : 1;
}

typedef struct {
  const mont_ctx* ctx; const uint32_t* exps; uint32_t nkeys;
  const uint32_t* key_idx; const uint8_t* sig; const uint8_t* digest; uint32_t hash_alg; int dlen;
  uint64_t lo, hi; int strict; uint8_t* status;
} job_t;
static void* worker(void* p) {
  job_t* j = (job_t*)p;
  for (uint64_t i = j->lo; i < j->hi; i++) {
    uint32_t k = j->key_idx[i];
    if (k >= j->nkeys) { j->status[i] = 4; continue; } /* unknown signer */
    j->status[i] = (uint8_t)verify_one(&j->ctx[k], j->exps[k], j->sig + i * 256, j->digest + i * (uint64_t)j->dlen, j->hash_alg, j->strict);
  }
  return 0;
}
/* Batch verify on `threads` host threads.  Returns 0 or -1 (bad key / hash id). */
int orc_rsa_verify_batch(const uint8_t* keys_n_be, const uint32_t* keys_e, uint32_t nkeys, const uint32_t* key_idx,
                         const uint8_t* sig_be, const uint8_t* digest, uint32_t hash_alg, uint64_t n_items, int strict_range,
                         int threads, uint8_t* status) {
  int dlen = orc_hash_dlen(hash_alg);
  if (!dlen) return -1;
  mont_ctx* ctx = (mont_ctx*)malloc(sizeof(mont_ctx) * (nkeys ? nkeys : 1));
  for (uint32_t k = 0; k < nkeys; k++) if (mont_init(&ctx[k], keys_n_be + (size_t)k * 256)) { free(ctx); return -1; }
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n_items) threads = n_items ? (int)n_items : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){ctx, keys_e, nkeys, key_idx, sig_be, digest, hash_alg, dlen, n_items * t / threads, n_items * (t + 1) / threads, strict_range, status};
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
  free(th); free(jobs); free(ctx);
  return 0;
}

/* ------------------------------------------------------------------ wotqs tally ------------ */
/* Quorum descriptor: nqc cliques; clique c = members[moff[c] .. moff[c+1]) with (f,min,threshold,suff)
 * in params[4c..4c+3].  Responders of op i = signer_id[ooff[i] .. ooff[i+1]) filtered by status==0.
 * Output bits: 1 IsQuorum, 2 IsThreshold, 4 IsSufficient, 8 Reject(of the FAILED responders, status!=0).
 * wotqs.go:144-206; duplicates in the responder list count each time (intersection iterates s1). */
void orc_tally_batch(const int32_t* params, const uint32_t* moff, const uint64_t* members, uint32_t nqc,
                     const uint32_t* ooff, const uint64_t* signer_id, const uint8_t* status, uint64_t n_ops, uint8_t* out) {
  for (uint64_t i = 0; i < n_ops; i++) {
    int is_q = nqc > 0, is_t = nqc > 0, is_s = 0, rej = 1;
    for (uint32_t c = 0; c < nqc; c++) {
      int f = params[4 * c], mn = params[4 * c + 1], th = params[4 * c + 2], sf = params[4 * c + 3];
      int cnt = 0, fail = 0;
      for (uint32_t p = ooff[i]; p < ooff[i + 1]; p++) {
        int in = 0;
        for (uint32_t m = moff[c]; m < moff[c + 1]; m++) if (members[m] == signer_id[p]) { in = 1; break; }
        if (in) { if (status[p] == 0) cnt++; else fail++; }
      }
      if (f > 0 && cnt < mn) is_q = 0;
      if (th > 0 && cnt < th) is_t = 0;
      if (sf > 0 && cnt >= sf) is_s = 1;
      if (f == 0 || fail <= f) rej = 0;
    }
    out[i] = (uint8_t)(is_q | (is_t << 1) | (is_s << 2) | (rej << 3));
  }
}

/* ------------------------------------------------------------------ read decision ---------- */
/* Client.Read's multicast callback (protocol/client.go:250-268) over the responders of op i in the order given
 * (= arrival order): a responder with status != 0 joins `failure` and q.Reject(failure) is tested (wotqs.go:178-185);
 * one with status == 0 is bucketed by (ts, value_id) (processResponse :207-230) and maxTimestampedValue (:189-205) is
 * asked: only the buckets of the maximum t so far, the first value (insertion order) whose responders pass IsThreshold
 * (wotqs.go:157-167, duplicates counted).  The first decisive response fixes the result:
 *   decision 0 value (winner = index inside the op of the first responder of the winning bucket), 1 rejected,
 *   2 exhausted (ErrInsufficientNumberOfResponses); decided_at = responses consumed (1-based; count for exhausted). */
static int in_clique(const uint32_t* moff, const uint64_t* members, uint32_t c, uint64_t id) {
  for (uint32_t m = moff[c]; m < moff[c + 1]; m++) if (members[m] == id) return 1;
  return 0;
}
void orc_read_decide_batch(const int32_t* params, const uint32_t* moff, const uint64_t* members, uint32_t nqc,
                           const uint32_t* ooff, const uint64_t* signer_id, const uint8_t* status, const uint64_t* ts,
                           const uint32_t* value_id, uint64_t n_ops, uint8_t* decision, uint32_t* winner, uint32_t* decided_at) {
  for (uint64_t i = 0; i < n_ops; i++) {
    const uint32_t lo = ooff[i], hi = ooff[i + 1];
    uint8_t dec = 2; uint32_t win = 0xffffffffu, at = hi - lo;
    for (uint32_t k = lo; k < hi && dec == 2; k++) {
      if (status[k] == 0) {
        uint64_t maxt = 0;                                   /* for t, vl := range m { if t >= maxt } */
        for (uint32_t p = lo; p <= k; p++) if (status[p] == 0 && ts[p] >= maxt) maxt = ts[p];
        /* values of the max-t bucket set in insertion order */
        for (uint32_t p = lo; p <= k && dec == 2; p++) {
          if (status[p] != 0 || ts[p] != maxt) continue;
          int first = 1;
          for (uint32_t p2 = lo; p2 < p; p2++) if (status[p2] == 0 && ts[p2] == maxt && value_id[p2] == value_id[p]) { first = 0; break; }
          if (!first) continue;
          int is_t = nqc > 0;
          for (uint32_t c = 0; c < nqc; c++) {
            int th = params[4 * c + 2], cnt = 0;
            for (uint32_t p2 = p; p2 <= k; p2++)
              if (status[p2] == 0 && ts[p2] == maxt && value_id[p2] == value_id[p] && in_clique(moff, members, c, signer_id[p2])) cnt++;
            if (th > 0 && cnt < th) is_t = 0;
          }
          if (is_t) { dec = 0; win = p - lo; at = k - lo + 1; }
        }
      } else {
        int rej = 1;
        for (uint32_t c = 0; c < nqc; c++) {
          int f = params[4 * c], fail = 0;
          for (uint32_t p = lo; p <= k; p++) if (status[p] != 0 && in_clique(moff, members, c, signer_id[p])) fail++;
          if (f == 0 || fail <= f) rej = 0;
        }
        if (rej) { dec = 1; at = k - lo + 1; }
      }
    }
    decision[i] = dec; winner[i] = win; decided_at[i] = at;
  }
}
