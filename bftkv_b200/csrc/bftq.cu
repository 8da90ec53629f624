// libbftq.so — C ABI (include/bftq.h) over the sm_100a kernels.
// Host side: engine life cycle, key table (per-key Montgomery constants), launch plumbing,
// pinned staging.  No CPU verification path exists here on purpose.
#include "../../include/bftq.h"
#include "rsa_verify.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(BFTQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));         \
  } while (0)

// ---- tiny host big-number helpers (2048-bit, 32 x u64 limbs, little-endian) -------------------
struct U2048 { uint64_t w[32]; };

bool ge(const U2048& a, const U2048& b) {
  for (int i = 31; i >= 0; i--) { if (a.w[i] != b.w[i]) return a.w[i] > b.w[i]; }
  return true;
}
void sub(U2048& a, const U2048& b) {
  unsigned __int128 br = 0;
  for (int i = 0; i < 32; i++) {
    unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - (uint64_t)br;
    a.w[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
// a = 2a mod n   (a < n on entry)
void dbl_mod(U2048& a, const U2048& n) {
  uint64_t top = a.w[31] >> 63;
  for (int i = 31; i > 0; i--) a.w[i] = (a.w[i] << 1) | (a.w[i - 1] >> 63);
  a.w[0] <<= 1;
  if (top || ge(a, n)) sub(a, n);
}
int bitlen(const U2048& a) {
  for (int i = 31; i >= 0; i--) if (a.w[i]) return 64 * i + 64 - __builtin_clzll(a.w[i]);
  return 0;
}
void from_be(U2048& a, const uint8_t* be) {   // 256 bytes big-endian
  for (int i = 0; i < 32; i++) {
    uint64_t v = 0;
    for (int b = 0; b < 8; b++) v = (v << 8) | be[256 - 8 * (i + 1) + b];
    a.w[i] = v;
  }
}
void to_digits(const U2048& a, uint32_t* d, int nd) {
  for (int i = 0; i < nd; i++) {
    int o = 28 * i;
    uint32_t v = 0;
    if (o < 2048) {
      int wi = o >> 6, sh = o & 63;
      unsigned __int128 t = a.w[wi];
      if (wi + 1 < 32) t |= (unsigned __int128)a.w[wi + 1] << 64;
      v = (uint32_t)(t >> sh) & bftq::kDigitMask;
    }
    d[i] = v;
  }
}

struct StagingSlot {
  cudaStream_t stream = nullptr;
  uint8_t* h_pinned = nullptr;  size_t h_cap = 0;
  uint8_t* d_buf = nullptr;     size_t d_cap = 0;
  bool busy = false;
};

}  // namespace

struct bftq_engine {
  int device = 0;
  int sm_count = 0;
  std::mutex mu;
  std::vector<bftq::RsaKeyDev> h_keys;
  bftq::RsaKeyDev* d_keys = nullptr;
  size_t d_keys_cap = 0;
  std::vector<StagingSlot*> slots;
  bftq_stats_t stats{};
  int rsa_t = 4;          // lanes per signature (env BFTQ_RSA_T)
  int rsa_block = 128;
};

namespace {

struct SlotLease {
  bftq_engine* e; StagingSlot* s;
  ~SlotLease() { std::lock_guard<std::mutex> g(e->mu); s->busy = false; }
};

int acquire_slot(bftq_engine* e, size_t h_bytes, size_t d_bytes, StagingSlot** out) {
  StagingSlot* s = nullptr;
  {
    std::lock_guard<std::mutex> g(e->mu);
    for (auto* c : e->slots) if (!c->busy) { s = c; break; }
    if (!s) { s = new StagingSlot(); e->slots.push_back(s); }
    s->busy = true;
  }
  CU(cudaSetDevice(e->device));
  if (!s->stream) CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (s->h_cap < h_bytes) {
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    s->h_pinned = nullptr; s->h_cap = 0;
    CU(cudaHostAlloc((void**)&s->h_pinned, h_bytes, cudaHostAllocDefault));
    s->h_cap = h_bytes;
  }
  if (s->d_cap < d_bytes) {
    if (s->d_buf) cudaFree(s->d_buf);
    s->d_buf = nullptr; s->d_cap = 0;
    CU(cudaMalloc((void**)&s->d_buf, d_bytes));
    s->d_cap = d_bytes;
  }
  *out = s;
  return BFTQ_OK;
}

template <int T, int W, int BLOCK>
int launch_rsa(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig, const uint8_t* d_digest,
               uint32_t hash_alg, uint64_t n_items, uint32_t flags, uint8_t* d_status, cudaStream_t st) {
  auto kern = bftq::rsa_verify_kernel<T, W, BLOCK>;
  static thread_local int occ_cache = 0;
  int occ = occ_cache;
  if (!occ) {
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, 0));
    if (occ < 1) occ = 1;
    occ_cache = occ;
  }
  const uint64_t per_block = (uint64_t)(BLOCK / 32) * (32 / T);
  uint64_t need = (n_items + per_block - 1) / per_block;
  uint64_t grid = std::min<uint64_t>(need, (uint64_t)e->sm_count * occ);
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, BLOCK, 0, st>>>(e->d_keys, (uint32_t)e->h_keys.size(), d_key_idx, d_sig, d_digest, hash_alg,
                                         n_items, flags, d_status);
  CU(cudaGetLastError());
  return BFTQ_OK;
}

int launch_rsa_any(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig, const uint8_t* d_digest,
                   uint32_t hash_alg, uint64_t n_items, uint32_t flags, uint8_t* d_status, cudaStream_t st) {
  {
    std::lock_guard<std::mutex> g(e->mu);
    e->stats.launches += 1;
    e->stats.items += n_items;
  }
  switch (e->rsa_t) {
    case 8: return launch_rsa<8, 10, 128>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_status, st);
    default: return launch_rsa<4, 19, 128>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_status, st);
  }
}

// ---- integer-pipe peak micro-benchmark ---------------------------------------------------------
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
  unsigned long long acc[16];
  uint32_t a[4];
  const uint32_t b = (seed | 1u) + 2u * threadIdx.x;
  for (int i = 0; i < 4; i++) a[i] = (seed ^ 0x9e3779b9u) * (i + 1) + threadIdx.x;
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i * seed;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i & 3]), "r"(b));
  }
  unsigned long long s = 0;
  for (int i = 0; i < 16; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

}  // namespace

extern "C" {

int bftq_version(void) { return BFTQ_VERSION; }
const char* bftq_last_error(void) { return g_last_error.c_str(); }

int bftq_init(int device, bftq_engine** out) {
  if (!out) return fail(BFTQ_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(BFTQ_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(BFTQ_ERR_INVALID_ARG, "device ordinal out of range");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(BFTQ_ERR_NO_DEVICE, std::string("device is not sm_100-class: ") + prop.name);
  auto* e = new bftq_engine();
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  if (const char* t = getenv("BFTQ_RSA_T")) {
    int v = atoi(t);
    if (v == 4 || v == 8) e->rsa_t = v;
  }
  *out = e;
  return BFTQ_OK;
}

void bftq_shutdown(bftq_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (auto* s : e->slots) {
    if (s->stream) { cudaStreamSynchronize(s->stream); cudaStreamDestroy(s->stream); }
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->d_buf) cudaFree(s->d_buf);
    delete s;
  }
  if (e->d_keys) cudaFree(e->d_keys);
  delete e;
}

int bftq_device_sm_count(bftq_engine* e) { return e ? e->sm_count : BFTQ_ERR_INVALID_ARG; }
int bftq_key_count(bftq_engine* e) {
  if (!e) return BFTQ_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(e->mu);
  return (int)e->h_keys.size();
}

int bftq_register_rsa_keys(bftq_engine* e, const uint8_t* n_be, const uint32_t* exps, uint32_t count,
                           uint32_t* first_index) {
  if (!e || !n_be || !exps) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::vector<bftq::RsaKeyDev> fresh(count);
  for (uint32_t k = 0; k < count; k++) {
    U2048 n;
    from_be(n, n_be + (size_t)k * 256);
    const int nb = bitlen(n);
    if (nb < 2041 || nb > 2048 || !(n.w[0] & 1))
      return fail(BFTQ_ERR_UNSUPPORTED_KEY, "modulus must be odd and 2041..2048 bits (key " + std::to_string(k) + ")");
    if (exps[k] == 0) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "public exponent 0");
    bftq::RsaKeyDev& kd = fresh[k];
    memset(&kd, 0, sizeof(kd));
    to_digits(n, kd.n, bftq::kMaxDigits);
    // -n^-1 mod 2^28 by Newton iteration on the low word.
    uint32_t n0 = (uint32_t)n.w[0], inv = n0;
    for (int i = 0; i < 5; i++) inv *= 2u - n0 * inv;
    kd.n0inv = (0u - inv) & bftq::kDigitMask;
    kd.e = exps[k];
    kd.nbits = (uint32_t)nb;
    // R^2 mod n for each digit layout: start from 2^2047 mod n and keep doubling.
    U2048 x;
    memset(&x, 0, sizeof(x));
    x.w[31] = 1ull << 63;                    // 2^2047
    if (ge(x, n)) sub(x, n);                 // n >= 2^2040 so one subtraction may not suffice...
    while (ge(x, n)) sub(x, n);
    int exp2 = 2047;
    for (int layout = 0; layout < bftq::kNumLayouts; layout++) {
      const int target = 2 * 28 * bftq::layout_digits(layout);
      while (exp2 < target) { dbl_mod(x, n); exp2++; }
      to_digits(x, kd.r2[layout], bftq::kMaxDigits);
    }
  }
  std::lock_guard<std::mutex> g(e->mu);
  CU(cudaSetDevice(e->device));
  const size_t old = e->h_keys.size();
  e->h_keys.insert(e->h_keys.end(), fresh.begin(), fresh.end());
  if (e->h_keys.size() > e->d_keys_cap) {
    // Kernels in flight may still read the old table: synchronise before replacing it.
    CU(cudaDeviceSynchronize());
    size_t cap = std::max<size_t>(64, e->h_keys.size() * 2);
    bftq::RsaKeyDev* nd = nullptr;
    CU(cudaMalloc((void**)&nd, cap * sizeof(bftq::RsaKeyDev)));
    if (e->d_keys) cudaFree(e->d_keys);
    e->d_keys = nd;
    e->d_keys_cap = cap;
    CU(cudaMemcpy(e->d_keys, e->h_keys.data(), e->h_keys.size() * sizeof(bftq::RsaKeyDev), cudaMemcpyHostToDevice));
  } else {
    CU(cudaMemcpy(e->d_keys + old, e->h_keys.data() + old, count * sizeof(bftq::RsaKeyDev), cudaMemcpyHostToDevice));
  }
  if (first_index) *first_index = (uint32_t)old;
  return BFTQ_OK;
}

int bftq_rsa_verify_batch_dev(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig_be,
                              const uint8_t* d_digest, uint32_t hash_alg, uint64_t n_items, uint32_t flags,
                              uint8_t* d_status, void* cuda_stream) {
  if (!e || !d_key_idx || !d_sig_be || !d_digest || !d_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (bftq::host_hash_dlen(hash_alg) == 0) return fail(BFTQ_ERR_INVALID_ARG, "unknown hash algorithm id");
  if (n_items == 0) return BFTQ_OK;
  if (!e->d_keys) return fail(BFTQ_ERR_INVALID_ARG, "no keys registered");
  CU(cudaSetDevice(e->device));
  return launch_rsa_any(e, d_key_idx, d_sig_be, d_digest, hash_alg, n_items, flags, d_status, (cudaStream_t)cuda_stream);
}

int bftq_rsa_verify_batch(bftq_engine* e, const uint32_t* key_idx, const uint8_t* sig_be, const uint8_t* digest,
                          uint32_t hash_alg, uint64_t n_items, uint32_t flags, uint8_t* out_status) {
  if (!e || !key_idx || !sig_be || !digest || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  const int dlen = bftq::host_hash_dlen(hash_alg);
  if (dlen == 0) return fail(BFTQ_ERR_INVALID_ARG, "unknown hash algorithm id");
  if (n_items == 0) return BFTQ_OK;
  if (!e->d_keys) return fail(BFTQ_ERR_INVALID_ARG, "no keys registered");
  // device layout of one staging buffer: [sig | digest | key_idx | status]
  const size_t sig_b = (size_t)n_items * 256, dig_b = (size_t)n_items * dlen, idx_b = (size_t)n_items * 4;
  const size_t off_dig = sig_b, off_idx = (off_dig + dig_b + 15) & ~(size_t)15, off_st = off_idx + idx_b;
  const size_t total = off_st + n_items;
  StagingSlot* s = nullptr;
  int rc = acquire_slot(e, total, total, &s);
  if (rc) return rc;
  SlotLease lease{e, s};
  // Pinned caller buffers go straight over PCIe; pageable ones are staged through pinned memory.
  auto h2d = [&](size_t off, const void* src, size_t bytes) -> int {
    cudaPointerAttributes at;
    bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const void* from = src;
    if (!pinned) { memcpy(s->h_pinned + off, src, bytes); from = s->h_pinned + off; }
    CU(cudaMemcpyAsync(s->d_buf + off, from, bytes, cudaMemcpyHostToDevice, s->stream));
    return BFTQ_OK;
  };
  if ((rc = h2d(0, sig_be, sig_b))) return rc;
  if ((rc = h2d(off_dig, digest, dig_b))) return rc;
  if ((rc = h2d(off_idx, key_idx, idx_b))) return rc;
  rc = launch_rsa_any(e, (const uint32_t*)(s->d_buf + off_idx), s->d_buf, s->d_buf + off_dig, hash_alg, n_items, flags,
                      s->d_buf + off_st, s->stream);
  if (rc) return rc;
  {
    cudaPointerAttributes at;
    bool pinned = cudaPointerGetAttributes(&at, out_status) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
      CU(cudaMemcpyAsync(out_status, s->d_buf + off_st, n_items, cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
    } else {
      CU(cudaMemcpyAsync(s->h_pinned + off_st, s->d_buf + off_st, n_items, cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
      memcpy(out_status, s->h_pinned + off_st, n_items);
    }
  }
  {
    std::lock_guard<std::mutex> g(e->mu);
    e->stats.h2d_bytes += sig_b + dig_b + idx_b;
    e->stats.d2h_bytes += n_items;
  }
  return BFTQ_OK;
}

int bftq_stats(bftq_engine* e, bftq_stats_t* out) {
  if (!e || !out) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> g(e->mu);
  *out = e->stats;
  return BFTQ_OK;
}

int bftq_measure_int_peak(bftq_engine* e, double* macs_per_second) {
  if (!e || !macs_per_second) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  CU(cudaSetDevice(e->device));
  const int blocks = e->sm_count * 8, iters = 4096;
  uint32_t* d = nullptr;
  CU(cudaMalloc((void**)&d, (size_t)blocks * 256 * 4));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  for (int w = 0; w < 20; w++) int_peak_kernel<<<blocks, 256>>>(d, 1234u + w, iters);   // warm clocks
  CU(cudaDeviceSynchronize());
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CU(cudaEventRecord(e0));
    for (int i = 0; i < 4; i++) int_peak_kernel<<<blocks, 256>>>(d, 99u + i, iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    double rate = 4.0 * blocks * 256.0 * iters * 16.0 / (ms * 1e-3);
    best = std::max(best, rate);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  {
    std::lock_guard<std::mutex> g(e->mu);
    e->stats.launches += 40;
  }
  *macs_per_second = best;
  return BFTQ_OK;
}

}  // extern "C"
