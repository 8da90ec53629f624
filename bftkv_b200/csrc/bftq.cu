// libbftq.so — C ABI (include/bftq.h) over the sm_100a kernels.
// Host side: engine life cycle, key table (per-key Montgomery constants), launch plumbing,
// pinned staging.  No CPU verification path exists here on purpose.
#include "../../include/bftq.h"
#include "rsa_verify.cuh"
#include "rsa_verify_r32.cuh"
#include "tally.cuh"
#include "lagrange.cuh"
#include "modexp.cuh"
#include "ed25519.cuh"
#include "p256.cuh"
#include "dsa_verify.cuh"
#include "pgp_digest.cuh"
#include "pgp_parse.cuh"
#include "pgp_host.hpp"
#include "wotqs_host.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <tuple>
#include <memory>
#include <thread>
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

// ---- tiny host big-number helpers (up to 4096 bit, 64 x u64 limbs, little-endian) ---------------
constexpr int kHL = 64;
struct UBig { uint64_t w[kHL]; };

bool ge(const UBig& a, const UBig& b) {
  for (int i = kHL - 1; i >= 0; i--) { if (a.w[i] != b.w[i]) return a.w[i] > b.w[i]; }
  return true;
}
void sub(UBig& a, const UBig& b) {
  unsigned __int128 br = 0;
  for (int i = 0; i < kHL; i++) {
    unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - (uint64_t)br;
    a.w[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
// a = 2a mod n   (a < n on entry)
void dbl_mod(UBig& a, const UBig& n) {
  uint64_t top = a.w[kHL - 1] >> 63;
  for (int i = kHL - 1; i > 0; i--) a.w[i] = (a.w[i] << 1) | (a.w[i - 1] >> 63);
  a.w[0] <<= 1;
  if (top || ge(a, n)) sub(a, n);
}
int bitlen(const UBig& a) {
  for (int i = kHL - 1; i >= 0; i--) if (a.w[i]) return 64 * i + 64 - __builtin_clzll(a.w[i]);
  return 0;
}
void from_be(UBig& a, const uint8_t* be, size_t len) {   // len bytes big-endian, len <= 512
  memset(&a, 0, sizeof(a));
  for (size_t i = 0; i < len; i++) {
    const size_t bi = len - 1 - i;                        // little-endian byte number
    a.w[bi >> 3] |= (uint64_t)be[i] << (8 * (bi & 7));
  }
}
void to_digits(const UBig& a, uint32_t* d, int nd) {
  for (int i = 0; i < nd; i++) {
    int o = 28 * i;
    uint32_t v = 0;
    if (o < 64 * kHL) {
      int wi = o >> 6, sh = o & 63;
      unsigned __int128 t = a.w[wi];
      if (wi + 1 < kHL) t |= (unsigned __int128)a.w[wi + 1] << 64;
      v = (uint32_t)(t >> sh) & bftq::kDigitMask;
    }
    d[i] = v;
  }
}

struct StagingSlot {
  cudaStream_t stream = nullptr;
  uint8_t* h_pinned = nullptr;  size_t h_cap = 0;
  uint8_t* d_buf = nullptr;     size_t d_cap = 0;
  cudaEvent_t done = nullptr;   // blocking-sync event: a waiting packer thread sleeps instead of spinning
  bool busy = false;
};

}  // namespace

// Persistent host workers for the packet-level entry points.  A batch call posts a job (its worker body) and
// works on it itself; idle pool threads join the oldest job that still has chunks to hand out.  Threads are
// created once (spawning 16 threads per call cost 0.5 ms of a 2.6 ms batch) and keep their per-thread caches.
struct PackerPool {
  struct Job {
    std::function<void()> body;
    std::function<bool()> has_work;          // false once every chunk has been handed out
    unsigned max_joiners = 0, joined = 0;    // pool threads allowed to / that did join (under PackerPool::mu)
    unsigned active = 0;                     // pool threads currently inside body()
  };
  std::mutex mu;
  std::condition_variable cv_jobs, cv_idle;
  std::deque<std::shared_ptr<Job>> jobs;
  std::vector<std::thread> threads;
  bool stop = false;

  void ensure(unsigned n) {                  // grow to n threads (under mu)
    while (threads.size() < n) threads.emplace_back([this] { loop(); });
  }
  void loop() {
    std::unique_lock<std::mutex> l(mu);
    for (;;) {
      std::shared_ptr<Job> j;
      cv_jobs.wait(l, [&] {
        if (stop) return true;
        for (auto& c : jobs) if (c->joined < c->max_joiners && c->has_work()) { j = c; return true; }
        return false;
      });
      if (stop) return;
      j->joined++; j->active++;
      l.unlock();
      j->body();
      l.lock();
      j->active--;
      cv_idle.notify_all();
    }
  }
  // Runs body on the calling thread and on up to `helpers` pool threads; returns when all of them are out.
  void run(unsigned helpers, std::function<void()> body, std::function<bool()> has_work) {
    auto j = std::make_shared<Job>();
    j->body = body; j->has_work = std::move(has_work); j->max_joiners = helpers;
    if (helpers) {
      std::lock_guard<std::mutex> l(mu);
      ensure(helpers);
      jobs.push_back(j);
    }
    if (helpers) cv_jobs.notify_all();
    body();
    if (!helpers) return;
    std::unique_lock<std::mutex> l(mu);
    for (auto it = jobs.begin(); it != jobs.end(); ++it) if (*it == j) { jobs.erase(it); break; }
    cv_idle.wait(l, [&] { return j->active == 0; });
  }
  ~PackerPool() {
    { std::lock_guard<std::mutex> l(mu); stop = true; }
    cv_jobs.notify_all();
    for (auto& t : threads) t.join();
  }
};

struct bftq_engine {
  int device = 0;
  int sm_count = 0;
  std::mutex mu;
  std::vector<bftq::RsaKeyDev> h_keys;
  bftq::RsaKeyDev* d_keys = nullptr;
  size_t d_keys_cap = 0;
  std::vector<bftq::r32::RsaKey32> h_keys32;     // radix-2^32 constants of the same keys
  bftq::r32::RsaKey32* d_keys32 = nullptr;
  bool all_2048 = true;                          // every registered modulus has exactly 2048 bits
  int rsa_kernel = 0;                            // 0 auto, 28 force radix-2^28, 32 force radix-2^32 (env BFTQ_RSA_KERNEL)
  std::vector<StagingSlot*> slots;
  bftq_stats_t stats{};
  std::map<std::string, uint32_t> key_lookup;   // (modulus bytes || e) -> key table index
  struct DsaKey { std::vector<uint8_t> p, q, gy; int cls; };   // gy: g || y, each padded to |p| bytes
  std::vector<DsaKey> dsa_keys;                  // host table; a group's domain travels with its launch
  std::map<std::string, uint32_t> dsa_lookup;
  PackerPool pool;                               // host workers of the packet-level entry points
  int rsa_t = 4;          // lanes per signature (env BFTQ_RSA_T)
  int rsa_block = 128;
};

namespace {

// Picks a free staging slot: the smallest one that is already large enough, else the largest free one
// (which then grows), else a new one.  Growing means cudaFreeHost / cudaHostAlloc / cudaMalloc — calls that
// stall the whole device — so capacities are rounded up to a power of two (at least 1 MiB): calls of varying
// size settle on a stable pool after a few batches instead of reallocating for ever.
int acquire_slot(bftq_engine* e, size_t h_bytes, size_t d_bytes, StagingSlot** out) {
  StagingSlot* s = nullptr;
  {
    std::lock_guard<std::mutex> g(e->mu);
    StagingSlot *fit = nullptr, *big = nullptr;
    for (auto* c : e->slots) {
      if (c->busy) continue;
      if (c->h_cap >= h_bytes && c->d_cap >= d_bytes) { if (!fit || c->h_cap < fit->h_cap) fit = c; }
      else if (!big || c->h_cap > big->h_cap) big = c;
    }
    s = fit ? fit : big;
    if (!s) { s = new StagingSlot(); e->slots.push_back(s); }
    s->busy = true;
  }
  auto round_up = [](size_t v) { size_t c = (size_t)1 << 20; while (c < v) c <<= 1; return c; };
  CU(cudaSetDevice(e->device));
  if (!s->stream) CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (!s->done) CU(cudaEventCreateWithFlags(&s->done, cudaEventBlockingSync | cudaEventDisableTiming));
  if (s->h_cap < h_bytes) {
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    s->h_pinned = nullptr; s->h_cap = 0;
    const size_t cap = round_up(h_bytes);
    CU(cudaHostAlloc((void**)&s->h_pinned, cap, cudaHostAllocDefault));
    s->h_cap = cap;
  }
  if (s->d_cap < d_bytes) {
    if (s->d_buf) cudaFree(s->d_buf);
    s->d_buf = nullptr; s->d_cap = 0;
    const size_t cap = round_up(d_bytes);
    CU(cudaMalloc((void**)&s->d_buf, cap));
    s->d_cap = cap;
  }
  *out = s;
  return BFTQ_OK;
}

// A call's staging plan: device buffers carved out of one slot, inputs uploaded in one go,
// outputs downloaded in one go.  Pinned caller memory is DMA'd directly, pageable memory is
// bounced through the slot's pinned mirror.
class Arena {
 public:
  explicit Arena(bftq_engine* e) : e_(e) {}
  ~Arena() { if (s_) { std::lock_guard<std::mutex> g(e_->mu); s_->busy = false; } }
  // count = elements reserved on the device, copy = elements actually copied (defaults to count)
  template <typename T> void in(T** dptr, const T* host, size_t count, size_t copy = (size_t)-1) {
    add((void**)dptr, (void*)host, count * sizeof(T), (copy == (size_t)-1 ? count : copy) * sizeof(T), true);
  }
  template <typename T> void out(T** dptr, T* host, size_t count, size_t copy = (size_t)-1) {
    add((void**)dptr, (void*)host, count * sizeof(T), (copy == (size_t)-1 ? count : copy) * sizeof(T), false);
  }
  // An input the caller composes in place: after prepare(), *hptr is the slot's pinned mirror of the
  // buffer (write `count` elements there), so the bytes cross host memory once.
  template <typename T> void stage(T** dptr, T** hptr, size_t count) {
    add((void**)dptr, nullptr, count * sizeof(T), count * sizeof(T), true);
    bufs_.back().hptr = (void**)hptr;
  }
  cudaStream_t stream() const { return s_->stream; }
  // Acquires the staging slot and resolves every device (and staged host) pointer.  upload() calls it
  // when the caller has not.
  int prepare() {
    if (s_) return BFTQ_OK;
    int rc = acquire_slot(e_, total_, total_, &s_);
    if (rc) return rc;
    for (auto& b : bufs_) {
      *b.dptr = s_->d_buf + b.off;
      if (b.hptr) { *b.hptr = s_->h_pinned + b.off; b.host = s_->h_pinned + b.off; }
    }
    return BFTQ_OK;
  }
  int upload() {
    int rc = prepare();
    if (rc) return rc;
    uint64_t h2d = 0;
    // Inputs that live in the slot's pinned mirror (staged in place or bounced) have the same layout on
    // both sides, so neighbours travel in ONE copy: a chunk of the packer costs one H2D call, not nine
    // (driver calls from many worker threads serialise on the context lock).
    size_t run_lo = (size_t)-1, run_hi = 0;
    auto flush = [&]() -> cudaError_t {
      if (run_lo == (size_t)-1) return cudaSuccess;
      cudaError_t ce = cudaMemcpyAsync(s_->d_buf + run_lo, s_->h_pinned + run_lo, run_hi - run_lo, cudaMemcpyHostToDevice, s_->stream);
      run_lo = (size_t)-1;
      return ce;
    };
    for (auto& b : bufs_) {
      if (!b.is_in || b.copy == 0) continue;
      h2d += b.copy;
      if (!b.hptr && is_pinned(b.host)) {                      // caller's pinned memory: DMA straight from it
        CU(cudaMemcpyAsync(s_->d_buf + b.off, b.host, b.copy, cudaMemcpyHostToDevice, s_->stream));
        continue;
      }
      if (!b.hptr) memcpy(s_->h_pinned + b.off, b.host, b.copy);
      if (run_lo != (size_t)-1 && b.off - run_hi > 65536) CU(flush());      // do not drag a large output region along
      if (run_lo == (size_t)-1) run_lo = b.off;
      run_hi = b.off + b.copy;
    }
    CU(flush());
    std::lock_guard<std::mutex> g(e_->mu);
    e_->stats.h2d_bytes += h2d;
    return BFTQ_OK;
  }
  int download() {
    int rc = download_async();
    if (rc) return rc;
    return finish();
  }
  // enqueue the device-to-host copies without waiting (finish() waits and un-bounces)
  int download_async() {
    uint64_t d2h = 0;
    bounce_.clear();
    for (auto& b : bufs_) {
      if (b.is_in || b.copy == 0) continue;
      if (is_pinned(b.host)) {
        CU(cudaMemcpyAsync(b.host, s_->d_buf + b.off, b.copy, cudaMemcpyDeviceToHost, s_->stream));
      } else {
        CU(cudaMemcpyAsync(s_->h_pinned + b.off, s_->d_buf + b.off, b.copy, cudaMemcpyDeviceToHost, s_->stream));
        bounce_.push_back(&b);
      }
      d2h += b.copy;
    }
    if (sleepy_) { CU(cudaEventRecord(s_->done, s_->stream)); recorded_ = true; }
    std::lock_guard<std::mutex> g(e_->mu);
    e_->stats.d2h_bytes += d2h;
    return BFTQ_OK;
  }
  // sleepy = wait on a blocking-sync event (the thread sleeps until the copy has landed) instead of spinning
  // in cudaStreamSynchronize: the packer's workers share the host's CPU quota with the threads still parsing.
  void set_sleepy(bool v) { sleepy_ = v; }
  int finish() {
    if (!s_) return BFTQ_OK;                          // nothing was enqueued (prepare() failed or was never called)
    if (sleepy_ && recorded_) CU(cudaEventSynchronize(s_->done));
    else CU(cudaStreamSynchronize(s_->stream));
    for (auto* b : bounce_) memcpy(b->host, s_->h_pinned + b->off, b->copy);
    bounce_.clear();
    return BFTQ_OK;
  }

 private:
  struct Buf { void** dptr; void* host; size_t off, bytes, copy; bool is_in; void** hptr; };
  void add(void** dptr, void* host, size_t bytes, size_t copy, bool is_in) {
    bufs_.push_back({dptr, host, total_, bytes, copy, is_in, nullptr});
    total_ += (bytes + 255) & ~(size_t)255;
  }
  static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
  }
  bftq_engine* e_;
  StagingSlot* s_ = nullptr;
  bool sleepy_ = false, recorded_ = false;
  std::vector<Buf> bufs_;
  std::vector<const Buf*> bounce_;
  size_t total_ = 0;
};

template <int T, int W, int BLOCK, int KB>
int launch_rsa(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig, const uint8_t* d_digest,
               uint32_t hash_alg, uint64_t n_items, uint32_t flags, const uint8_t* d_pre, uint8_t* d_status, cudaStream_t st) {
  auto kern = bftq::rsa_verify_kernel<T, W, BLOCK, KB>;
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
                                         n_items, flags, d_pre, d_status);
  CU(cudaGetLastError());
  return BFTQ_OK;
}

int launch_rsa_any(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig, const uint8_t* d_digest,
                   uint32_t hash_alg, uint64_t n_items, uint32_t flags, const uint8_t* d_pre, uint8_t* d_status, cudaStream_t st,
                   int kb = 256) {
  {
    std::lock_guard<std::mutex> g(e->mu);
    e->stats.launches += 1;
    e->stats.items += n_items;
  }
  const bool use32 = kb == 256 && (e->rsa_kernel == 32 || (e->rsa_kernel == 0 && e->all_2048));
  if (use32) {
    if (!e->all_2048) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "radix-2^32 kernel forced but a registered modulus is not 2048 bits");
    static const int min_blocks = [] { const char* v = getenv("BFTQ_R32_BLOCKS"); return v ? atoi(v) : 4; }();
    auto kern = min_blocks == 5 ? bftq::r32::rsa_verify_r32_kernel<128, 5> : (min_blocks == 3 ? bftq::r32::rsa_verify_r32_kernel<128, 3> : bftq::r32::rsa_verify_r32_kernel<128, 4>);
    static thread_local int occ32 = 0;
    if (!occ32) { CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, kern, 128, 0)); if (occ32 < 1) occ32 = 1; }
    const uint64_t per_block = 4 * 8;
    uint64_t grid = std::min<uint64_t>((n_items + per_block - 1) / per_block, (uint64_t)e->sm_count * occ32);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, 128, 0, st>>>(e->d_keys32, (uint32_t)e->h_keys32.size(), d_key_idx, d_sig, d_digest, hash_alg, n_items, flags,
                                         d_pre, d_status);
    CU(cudaGetLastError());
    return BFTQ_OK;
  }
  switch (kb) {
    case 128: return launch_rsa<4, 10, 128, 128>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_pre, d_status, st);
    case 192: return launch_rsa<4, 14, 128, 192>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_pre, d_status, st);
    case 384: return launch_rsa<8, 14, 128, 384>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_pre, d_status, st);
    case 512: return launch_rsa<8, 19, 128, 512>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_pre, d_status, st);
    case 256: break;
    default: return fail(BFTQ_ERR_UNSUPPORTED_KEY, "key size class not built (128/192/256/384/512 bytes are)");
  }
  if (e->rsa_t == 8) return launch_rsa<8, 10, 128, 256>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_pre, d_status, st);
  return launch_rsa<4, 19, 128, 256>(e, d_key_idx, d_sig, d_digest, hash_alg, n_items, flags, d_pre, d_status, st);
}

// ---- integer-pipe peak micro-benchmark ---------------------------------------------------------
// 16 independent accumulators per thread; the multiplicand changes every iteration (rotated through
// the accumulators' own low words) so ptxas cannot strength-reduce the products into additions —
// an earlier version with loop-invariant multiplicands was silently turned into IADD3 pairs and
// reported the ALU-pipe rate instead (profiles/int_pipe_ubench_r01.json, "imad_wide_*" rows).
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
  unsigned long long acc[16];
  uint32_t a[16];
  for (int i = 0; i < 16; i++) {
    acc[i] = (unsigned long long)(threadIdx.x + 1) * (i + seed);
    a[i] = (seed ^ 0x9e3779b9u) * (2 * i + 1) + threadIdx.x;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    const uint32_t b = __shfl_xor_sync(0xffffffffu, (uint32_t)acc[0], 1) | 1u;   // opaque, changes every iteration
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] += (unsigned long long)a[i] * b;     // IMAD.WIDE.U32 R, a, b, R
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
  if (const char* k = getenv("BFTQ_RSA_KERNEL")) {
    if (!strcmp(k, "r28")) e->rsa_kernel = 28;
    if (!strcmp(k, "r32")) e->rsa_kernel = 32;
  }
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
    if (s->done) cudaEventDestroy(s->done);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->d_buf) cudaFree(s->d_buf);
    delete s;
  }
  if (e->d_keys) cudaFree(e->d_keys);
  if (e->d_keys32) cudaFree(e->d_keys32);
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
  return bftq_register_rsa_keys_k(e, n_be, 256, exps, count, first_index);
}

int bftq_register_rsa_keys_k(bftq_engine* e, const uint8_t* n_be, uint32_t stride, const uint32_t* exps, uint32_t count,
                             uint32_t* first_index) {
  if (!e || !n_be || !exps) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (stride == 0 || stride > 512) return fail(BFTQ_ERR_INVALID_ARG, "modulus stride must be 1..512 bytes");
  std::vector<bftq::RsaKeyDev> fresh(count);
  std::vector<bftq::r32::RsaKey32> fresh32(count);
  bool fresh_all_2048 = true;
  for (uint32_t k = 0; k < count; k++) {
    UBig n;
    from_be(n, n_be + (size_t)k * stride, stride);
    const int nb = bitlen(n);
    const int kb = (nb + 7) / 8;
    const int cls = bftq::class_of(kb);
    if (!cls || !(n.w[0] & 1))
      return fail(BFTQ_ERR_UNSUPPORTED_KEY, "modulus must be odd and at most 4096 bits (key " + std::to_string(k) + ")");
    if (exps[k] == 0) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "public exponent 0");
    bftq::RsaKeyDev& kd = fresh[k];
    memset(&kd, 0, sizeof(kd));
    to_digits(n, kd.n, bftq::kMaxDigits);
    // -n^-1 mod 2^32 by Newton iteration on the low word (masked to 28 bits for the digit kernels).
    uint32_t n0 = (uint32_t)n.w[0], inv = n0;
    for (int i = 0; i < 5; i++) inv *= 2u - n0 * inv;
    kd.n0inv = (0u - inv) & bftq::kDigitMask;
    kd.e = exps[k];
    kd.nbits = (uint32_t)nb;
    kd.kbytes = (uint32_t)kb;
    // R^2 mod n for each digit layout of the class: 2^(2*28*digits) by repeated doubling from 1.
    UBig x;
    memset(&x, 0, sizeof(x));
    x.w[0] = 1;
    int exp2 = 0;
    for (int layout = 0; layout < bftq::kNumLayouts; layout++) {
      const int digits = bftq::class_digits(cls, layout == 1 && cls != 256 ? 0 : layout);
      const int target = 2 * 28 * digits;
      while (exp2 < target) { dbl_mod(x, n); exp2++; }
      if (exp2 == target) to_digits(x, kd.r2[layout], bftq::kMaxDigits);
    }
    // radix-2^32 constants (fast path, meaningful for exactly-2048-bit moduli): n, 2^4096 mod n, -n^-1 mod 2^32
    bftq::r32::RsaKey32& k32 = fresh32[k];
    memset(&k32, 0, sizeof(k32));
    for (int i = 0; i < 32; i++) { k32.n[2 * i] = (uint32_t)n.w[i]; k32.n[2 * i + 1] = (uint32_t)(n.w[i] >> 32); }
    k32.n0inv = 0u - inv;
    k32.e = exps[k];
    k32.nbits = (uint32_t)nb;
    if (nb == 2048) {
      UBig y;
      memset(&y, 0, sizeof(y));
      y.w[0] = 1;
      for (int ex = 0; ex < 4096; ex++) dbl_mod(y, n);
      for (int i = 0; i < 32; i++) { k32.r2[2 * i] = (uint32_t)y.w[i]; k32.r2[2 * i + 1] = (uint32_t)(y.w[i] >> 32); }
    }
    if (cls == 256 && nb != 2048) fresh_all_2048 = false;
  }
  std::lock_guard<std::mutex> g(e->mu);
  CU(cudaSetDevice(e->device));
  const size_t old = e->h_keys.size();
  e->h_keys.insert(e->h_keys.end(), fresh.begin(), fresh.end());
  e->h_keys32.insert(e->h_keys32.end(), fresh32.begin(), fresh32.end());
  e->all_2048 = e->all_2048 && fresh_all_2048;
  if (e->h_keys.size() > e->d_keys_cap) {
    // Kernels in flight may still read the old table: synchronise before replacing it.
    CU(cudaDeviceSynchronize());
    size_t cap = std::max<size_t>(64, e->h_keys.size() * 2);
    bftq::RsaKeyDev* nd = nullptr;
    bftq::r32::RsaKey32* nd32 = nullptr;
    CU(cudaMalloc((void**)&nd, cap * sizeof(bftq::RsaKeyDev)));
    CU(cudaMalloc((void**)&nd32, cap * sizeof(bftq::r32::RsaKey32)));
    if (e->d_keys) cudaFree(e->d_keys);
    if (e->d_keys32) cudaFree(e->d_keys32);
    e->d_keys = nd;
    e->d_keys32 = nd32;
    e->d_keys_cap = cap;
    CU(cudaMemcpy(e->d_keys, e->h_keys.data(), e->h_keys.size() * sizeof(bftq::RsaKeyDev), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_keys32, e->h_keys32.data(), e->h_keys32.size() * sizeof(bftq::r32::RsaKey32), cudaMemcpyHostToDevice));
  } else {
    CU(cudaMemcpy(e->d_keys + old, e->h_keys.data() + old, count * sizeof(bftq::RsaKeyDev), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_keys32 + old, e->h_keys32.data() + old, count * sizeof(bftq::r32::RsaKey32), cudaMemcpyHostToDevice));
  }
  if (first_index) *first_index = (uint32_t)old;
  return BFTQ_OK;
}

int bftq_rsa_verify_batch_dev(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig_be,
                              const uint8_t* d_digest, uint32_t hash_alg, uint64_t n_items, uint32_t flags,
                              uint8_t* d_status, void* cuda_stream) {
  return bftq_rsa_verify_batch_dev_k(e, 256, d_key_idx, d_sig_be, d_digest, hash_alg, n_items, flags, d_status, cuda_stream);
}

int bftq_rsa_verify_batch_dev_k(bftq_engine* e, uint32_t key_bytes, const uint32_t* d_key_idx, const uint8_t* d_sig_be,
                                const uint8_t* d_digest, uint32_t hash_alg, uint64_t n_items, uint32_t flags,
                                uint8_t* d_status, void* cuda_stream) {
  if (!bftq::class_supported((int)key_bytes)) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "key size class not built");
  if (!e || !d_key_idx || !d_sig_be || !d_digest || !d_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (bftq::host_hash_dlen(hash_alg) == 0) return fail(BFTQ_ERR_INVALID_ARG, "unknown hash algorithm id");
  if (n_items == 0) return BFTQ_OK;
  if (!e->d_keys) return fail(BFTQ_ERR_INVALID_ARG, "no keys registered");
  CU(cudaSetDevice(e->device));
  return launch_rsa_any(e, d_key_idx, d_sig_be, d_digest, hash_alg, n_items, flags, nullptr, d_status, (cudaStream_t)cuda_stream, (int)key_bytes);
}

int bftq_rsa_verify_batch(bftq_engine* e, const uint32_t* key_idx, const uint8_t* sig_be, const uint8_t* digest,
                          uint32_t hash_alg, uint64_t n_items, uint32_t flags, uint8_t* out_status) {
  return bftq_rsa_verify_batch_k(e, 256, key_idx, sig_be, digest, hash_alg, n_items, flags, out_status);
}

int bftq_rsa_verify_batch_k(bftq_engine* e, uint32_t key_bytes, const uint32_t* key_idx, const uint8_t* sig_be, const uint8_t* digest,
                            uint32_t hash_alg, uint64_t n_items, uint32_t flags, uint8_t* out_status) {
  if (!e || !key_idx || !sig_be || !digest || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (!bftq::class_supported((int)key_bytes)) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "key size class not built");
  const int dlen = bftq::host_hash_dlen(hash_alg);
  if (dlen == 0) return fail(BFTQ_ERR_INVALID_ARG, "unknown hash algorithm id");
  if (n_items == 0) return BFTQ_OK;
  if (!e->d_keys) return fail(BFTQ_ERR_INVALID_ARG, "no keys registered");
  // Large batches are cut into chunks that travel through a ring of staging slots (one stream each): the copy of
  // chunk c+1 runs under the kernel of chunk c, and kernels of neighbouring chunks run out of phase
  // (tools/e2e_experiment.py: 33.8 M/s for one caller unchunked, ~50 M/s with four 16384-item pieces in flight).
  static const uint64_t kChunk = [] { const char* v = getenv("BFTQ_HOST_CHUNK"); const long long c = v ? atoll(v) : 16384; return (uint64_t)(c > 0 ? c : 16384); }();
  constexpr int kDepth = 4;
  const uint64_t n_chunks = n_items <= kChunk + kChunk / 2 ? 1 : (n_items + kChunk - 1) / kChunk;
  const uint64_t per = (n_items + n_chunks - 1) / n_chunks;
  std::unique_ptr<Arena> ring[kDepth];
  int rc = BFTQ_OK;
  for (uint64_t c = 0; c < n_chunks && rc == BFTQ_OK; c++) {
    const uint64_t lo = c * per, cnt = std::min(per, n_items - lo);
    std::unique_ptr<Arena>& slot = ring[c % kDepth];
    if (slot) { rc = slot->finish(); slot.reset(); if (rc) break; }
    slot.reset(new Arena(e));
    Arena& a = *slot;
    uint8_t *d_sig, *d_dig, *d_st; uint32_t* d_idx;
    a.in(&d_sig, sig_be + lo * key_bytes, (size_t)cnt * key_bytes);
    a.in(&d_dig, digest + lo * dlen, (size_t)cnt * dlen);
    a.in(&d_idx, key_idx + lo, (size_t)cnt);
    a.out(&d_st, out_status + lo, (size_t)cnt);
    rc = a.upload();
    if (rc) break;
    rc = launch_rsa_any(e, d_idx, d_sig, d_dig, hash_alg, cnt, flags, nullptr, d_st, a.stream(), (int)key_bytes);
    if (rc) break;
    rc = a.download_async();
  }
  for (auto& slot : ring) if (slot) { const int r2 = slot->finish(); if (!rc) rc = r2; slot.reset(); }
  return rc;
}

// ---- K1b --------------------------------------------------------------------------------------
int bftq_ed25519_verify_batch_dev(bftq_engine* e, const uint8_t* d_pubkeys, uint32_t n_keys, const uint32_t* d_key_idx,
                                  const uint8_t* d_sig, const uint8_t* d_msg, uint64_t n_items, uint8_t* d_status,
                                  void* cuda_stream) {
  if (!e || !d_pubkeys || !d_key_idx || !d_sig || !d_msg || !d_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (n_items == 0) return BFTQ_OK;
  CU(cudaSetDevice(e->device));
  const int block = 128;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // Signatures that share keys (OpenPGP: a handful of keys, many signatures) pay the doublings once per key: window
  // tables for -A of every key and for the base point (82 KB each, stream-ordered scratch), then at most 128 table
  // additions per signature.  With few signatures per key the classic double-and-add kernel is cheaper.
  static const bool no_tables = [] { const char* v = getenv("BFTQ_ED25519_TABLES"); return v && atoi(v) == 0; }();
  const bool windowed = !no_tables && n_keys > 0 && n_keys <= 4096 && n_items >= 64ull * ((uint64_t)n_keys + 1);
  int launches = 1;
  if (windowed) {
    bftq::ed::gec* d_tab = nullptr;
    uint8_t* d_ok = nullptr;
    const size_t tab_bytes = ((size_t)n_keys + 1) * bftq::ed::kEdTableEntries * sizeof(bftq::ed::gec);
    CU(cudaMallocAsync((void**)&d_tab, tab_bytes + n_keys, st));
    d_ok = reinterpret_cast<uint8_t*>(d_tab) + tab_bytes;
    bftq::ed25519_table_kernel<<<n_keys + 1, 64, 0, st>>>(d_pubkeys, n_keys, d_tab, d_ok);
    bftq::ed25519_verify_windowed_kernel<<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(
        d_pubkeys, n_keys, d_key_idx, d_sig, d_msg, n_items, d_tab, d_ok, d_status);
    CU(cudaGetLastError());
    CU(cudaFreeAsync(d_tab, st));
    launches = 2;
  } else {
    bftq::ed25519_verify_kernel<<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(
        d_pubkeys, n_keys, d_key_idx, d_sig, d_msg, n_items, d_status);
    CU(cudaGetLastError());
  }
  std::lock_guard<std::mutex> g(e->mu);
  e->stats.launches += launches;
  e->stats.items += n_items;
  return BFTQ_OK;
}

int bftq_ed25519_verify_batch(bftq_engine* e, const uint8_t* pubkeys, uint32_t n_keys, const uint32_t* key_idx,
                              const uint8_t* sig, const uint8_t* msg, uint64_t n_items, uint8_t* out_status) {
  if (!e || !pubkeys || !key_idx || !sig || !msg || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  uint8_t *d_pk, *d_sig, *d_msg, *d_st; uint32_t* d_idx;
  a.in(&d_pk, pubkeys, (size_t)std::max<uint32_t>(n_keys, 1) * 32, (size_t)n_keys * 32);
  a.in(&d_idx, key_idx, (size_t)n_items);
  a.in(&d_sig, sig, (size_t)n_items * 64);
  a.in(&d_msg, msg, (size_t)n_items * 32);
  a.out(&d_st, out_status, (size_t)n_items);
  int rc = a.upload();
  if (rc) return rc;
  rc = bftq_ed25519_verify_batch_dev(e, d_pk, n_keys, d_idx, d_sig, d_msg, n_items, d_st, a.stream());
  if (rc) return rc;
  return a.download();
}

// ---- K1c --------------------------------------------------------------------------------------
int bftq_ecdsa_p256_verify_batch(bftq_engine* e, const uint8_t* pubkeys, uint32_t n_keys, const uint32_t* key_idx,
                                 const uint8_t* r_be, const uint8_t* s_be, const uint8_t* digest, uint32_t digest_len,
                                 uint64_t n_items, uint8_t* out_status) {
  if (!e || !pubkeys || !key_idx || !r_be || !s_be || !digest || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (digest_len == 0 || digest_len > 64) return fail(BFTQ_ERR_INVALID_ARG, "digest_len must be 1..64");
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  uint8_t *d_pk, *d_r, *d_s, *d_dg, *d_st; uint32_t* d_idx;
  a.in(&d_pk, pubkeys, (size_t)std::max<uint32_t>(n_keys, 1) * 64, (size_t)n_keys * 64);
  a.in(&d_idx, key_idx, (size_t)n_items);
  a.in(&d_r, r_be, (size_t)n_items * 32);
  a.in(&d_s, s_be, (size_t)n_items * 32);
  a.in(&d_dg, digest, (size_t)n_items * digest_len);
  a.out(&d_st, out_status, (size_t)n_items);
  int rc = a.upload();
  if (rc) return rc;
  const int block = 128;
  bftq::ecdsa_p256_verify_kernel<<<(unsigned)((n_items + block - 1) / block), block, 0, a.stream()>>>(
      d_pk, n_keys, d_idx, d_r, d_s, d_dg, digest_len, n_items, nullptr, d_st);
  CU(cudaGetLastError());
  {
    std::lock_guard<std::mutex> g(e->mu);
    e->stats.launches += 1;
    e->stats.items += n_items;
  }
  return a.download();
}

// ---- K2 ---------------------------------------------------------------------------------------
}  // extern "C"

struct bftq_quorum {
  bftq::QuorumDev dev;
  uint32_t* d_bits = nullptr;
};

namespace {
int launch_tally(bftq_engine* e, const bftq_quorum* q, const uint32_t* d_off, const uint32_t* d_idx, const uint8_t* d_status,
                 const uint64_t* d_ts, const uint32_t* d_val, uint64_t n_ops, uint32_t* d_winner, uint8_t* d_bits, cudaStream_t st) {
  const int block = 256, wpb = block / 32;
  uint64_t grid = std::min<uint64_t>((n_ops + wpb - 1) / wpb, (uint64_t)e->sm_count * 8);
  if (grid < 1) grid = 1;
  if (d_ts && d_val)
    bftq::read_tally_kernel<<<(unsigned)grid, block, 0, st>>>(q->dev, d_off, d_idx, d_status, d_ts, d_val, n_ops, d_winner, d_bits);
  else
    bftq::tally_kernel<<<(unsigned)grid, block, 0, st>>>(q->dev, d_off, d_idx, d_status, n_ops, d_bits);
  CU(cudaGetLastError());
  std::lock_guard<std::mutex> g(e->mu);
  e->stats.launches += 1;
  return BFTQ_OK;
}
}  // namespace

extern "C" {

int bftq_quorum_create(bftq_engine* e, const bftq_qc_t* qcs, uint32_t n_qc, const uint32_t* member_key_idx,
                       uint32_t n_members, bftq_quorum** out) {
  if (!e || !out || (n_qc && !qcs) || (n_members && !member_key_idx)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (n_qc > (uint32_t)bftq::kMaxQc) return fail(BFTQ_ERR_INVALID_ARG, "too many quorum cliques");
  uint32_t maxk = 0;
  for (uint32_t c = 0; c < n_qc; c++) {
    if ((uint64_t)qcs[c].member_off + qcs[c].member_cnt > n_members) return fail(BFTQ_ERR_INVALID_ARG, "clique members out of range");
    for (uint32_t m = 0; m < qcs[c].member_cnt; m++) maxk = std::max(maxk, member_key_idx[qcs[c].member_off + m]);
  }
  if (maxk > (1u << 20)) return fail(BFTQ_ERR_INVALID_ARG, "member key index too large");
  auto* q = new bftq_quorum();
  memset(&q->dev, 0, sizeof(q->dev));
  q->dev.nqc = (int32_t)n_qc;
  q->dev.nkeys_words = maxk / 32 + 1;
  std::vector<uint32_t> bits((size_t)std::max<uint32_t>(n_qc, 1) * q->dev.nkeys_words, 0u);
  for (uint32_t c = 0; c < n_qc; c++) {
    q->dev.f[c] = qcs[c].f; q->dev.min[c] = qcs[c].min; q->dev.threshold[c] = qcs[c].threshold; q->dev.suff[c] = qcs[c].suff;
    for (uint32_t m = 0; m < qcs[c].member_cnt; m++) {
      const uint32_t k = member_key_idx[qcs[c].member_off + m];
      bits[(size_t)c * q->dev.nkeys_words + (k >> 5)] |= 1u << (k & 31);
    }
  }
  cudaSetDevice(e->device);
  if (cudaMalloc((void**)&q->d_bits, bits.size() * 4) != cudaSuccess ||
      cudaMemcpy(q->d_bits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    delete q;
    return fail(BFTQ_ERR_CUDA, "quorum upload failed");
  }
  q->dev.member_bits = q->d_bits;
  *out = q;
  return BFTQ_OK;
}

void bftq_quorum_destroy(bftq_engine* e, bftq_quorum* q) {
  if (!q) return;
  if (e) cudaSetDevice(e->device);
  if (q->d_bits) { cudaDeviceSynchronize(); cudaFree(q->d_bits); }
  delete q;
}

static int tally_host(bftq_engine* e, const bftq_quorum* q, const uint32_t* op_off, const uint32_t* key_idx, const uint8_t* status,
                      const uint64_t* ts, const uint32_t* value_id, uint64_t n_ops, uint32_t* out_winner, uint8_t* out_bits) {
  if (!e || !q || !op_off || !out_bits) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (n_ops == 0) return BFTQ_OK;
  const uint64_t n_items = op_off[n_ops];
  if (n_items && (!key_idx || !status)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (ts && value_id)
    for (uint64_t i = 0; i < n_ops; i++)
      if (op_off[i + 1] - op_off[i] > 32) return fail(BFTQ_ERR_INVALID_ARG, "read tally: more than 32 responders in one operation");
  Arena a(e);
  uint32_t *d_off, *d_idx, *d_val = nullptr, *d_win = nullptr; uint8_t *d_st, *d_bits; uint64_t* d_ts = nullptr;
  a.in(&d_off, op_off, (size_t)n_ops + 1);
  a.in(&d_idx, key_idx, (size_t)std::max<uint64_t>(n_items, 1), (size_t)n_items);
  a.in(&d_st, status, (size_t)std::max<uint64_t>(n_items, 1), (size_t)n_items);
  if (ts && value_id) {
    a.in(&d_ts, ts, (size_t)std::max<uint64_t>(n_items, 1), (size_t)n_items);
    a.in(&d_val, value_id, (size_t)std::max<uint64_t>(n_items, 1), (size_t)n_items);
    a.out(&d_win, out_winner, (size_t)n_ops);
  }
  a.out(&d_bits, out_bits, (size_t)n_ops);
  int rc = a.upload();
  if (rc) return rc;
  rc = launch_tally(e, q, d_off, d_idx, d_st, d_ts, d_val, n_ops, d_win, d_bits, a.stream());
  if (rc) return rc;
  return a.download();
}

int bftq_tally_batch(bftq_engine* e, const bftq_quorum* q, const uint32_t* op_off, const uint32_t* key_idx,
                     const uint8_t* status, uint64_t n_ops, uint8_t* out_bits) {
  return tally_host(e, q, op_off, key_idx, status, nullptr, nullptr, n_ops, nullptr, out_bits);
}

int bftq_read_tally_batch(bftq_engine* e, const bftq_quorum* q, const uint32_t* op_off, const uint32_t* key_idx,
                          const uint8_t* status, const uint64_t* ts, const uint32_t* value_id, uint64_t n_ops,
                          uint32_t* out_winner, uint8_t* out_bits) {
  if (!ts || !value_id || !out_winner) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  return tally_host(e, q, op_off, key_idx, status, ts, value_id, n_ops, out_winner, out_bits);
}

int bftq_verify_tally_batch_dev(bftq_engine* e, const bftq_quorum* q, const uint32_t* d_op_off, const uint32_t* d_key_idx,
                                const uint8_t* d_sig_be, const uint8_t* d_digest, uint32_t hash_alg,
                                const uint8_t* d_pre_status, const uint64_t* d_ts, const uint32_t* d_value_id,
                                uint64_t n_ops, uint64_t n_items, uint32_t flags, uint8_t* d_status, uint8_t* d_bits,
                                uint32_t* d_winner, void* cuda_stream) {
  if (!e || !q || !d_op_off || !d_status || !d_bits) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (bftq::host_hash_dlen(hash_alg) == 0) return fail(BFTQ_ERR_INVALID_ARG, "unknown hash algorithm id");
  if (n_ops == 0) return BFTQ_OK;
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (n_items) {
    if (!e->d_keys) return fail(BFTQ_ERR_INVALID_ARG, "no keys registered");
    int rc = launch_rsa_any(e, d_key_idx, d_sig_be, d_digest, hash_alg, n_items, flags, d_pre_status, d_status, st);
    if (rc) return rc;
  }
  return launch_tally(e, q, d_op_off, d_key_idx, d_status, d_ts, d_value_id, n_ops, d_winner, d_bits, st);
}

int bftq_verify_tally_batch(bftq_engine* e, const bftq_quorum* q, const uint32_t* op_off, const uint32_t* key_idx,
                            const uint8_t* sig_be, const uint8_t* digest, uint32_t hash_alg, const uint8_t* pre_status,
                            const uint64_t* ts, const uint32_t* value_id, uint64_t n_ops, uint32_t flags,
                            uint8_t* out_status, uint8_t* out_bits, uint32_t* out_winner) {
  if (!e || !q || !op_off || !out_status || !out_bits) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  const int dlen = bftq::host_hash_dlen(hash_alg);
  if (dlen == 0) return fail(BFTQ_ERR_INVALID_ARG, "unknown hash algorithm id");
  if (n_ops == 0) return BFTQ_OK;
  const uint64_t n_items = op_off[n_ops];
  if (n_items && (!key_idx || !sig_be || !digest)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  const bool read = ts && value_id;
  if (read) {
    if (!out_winner) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
    for (uint64_t i = 0; i < n_ops; i++)
      if (op_off[i + 1] - op_off[i] > 32) return fail(BFTQ_ERR_INVALID_ARG, "read tally: more than 32 responders in one operation");
  }
  const size_t ni = (size_t)std::max<uint64_t>(n_items, 1);
  Arena a(e);
  uint32_t *d_off, *d_idx, *d_val = nullptr, *d_win = nullptr; uint8_t *d_sig, *d_dig, *d_pre = nullptr, *d_st, *d_bits; uint64_t* d_ts = nullptr;
  a.in(&d_sig, sig_be, ni * 256, (size_t)n_items * 256);
  a.in(&d_dig, digest, ni * dlen, (size_t)n_items * dlen);
  a.in(&d_off, op_off, (size_t)n_ops + 1);
  a.in(&d_idx, key_idx, ni, (size_t)n_items);
  if (pre_status) a.in(&d_pre, pre_status, ni, (size_t)n_items);
  if (read) { a.in(&d_ts, ts, ni, (size_t)n_items); a.in(&d_val, value_id, ni, (size_t)n_items); a.out(&d_win, out_winner, (size_t)n_ops); }
  a.out(&d_st, out_status, ni, (size_t)n_items);
  a.out(&d_bits, out_bits, (size_t)n_ops);
  int rc = a.upload();
  if (rc) return rc;
  rc = bftq_verify_tally_batch_dev(e, q, d_off, d_idx, d_sig, d_dig, hash_alg, d_pre, d_ts, d_val, n_ops, n_items, flags, d_st, d_bits,
                                   d_win, a.stream());
  if (rc) return rc;
  return a.download();
}

// ---- K3 ---------------------------------------------------------------------------------------
}  // extern "C"

namespace {
template <int L>
void make_lagrange_mod(const uint8_t* m_be, uint32_t mlen, bftq::LagrangeMod<L>& M) {
  memset(&M, 0, sizeof(M));
  for (uint32_t i = 0; i < mlen; i++) {
    const uint32_t bi = mlen - 1 - i;          // little-endian byte number
    M.m[bi >> 2] |= (uint32_t)m_be[i] << (8 * (bi & 3));
  }
  M.mlen = mlen;
  uint32_t inv = M.m[0];
  for (int i = 0; i < 5; i++) inv *= 2u - M.m[0] * inv;
  M.m0inv = 0u - inv;
  // R mod m and R^2 mod m by shift-and-subtract from 1
  std::vector<uint32_t> r(L + 1, 0u);
  r[0] = 1;
  auto ge_m = [&](const std::vector<uint32_t>& v) {
    if (v[L]) return true;
    for (int i = L - 1; i >= 0; i--) if (v[i] != M.m[i]) return v[i] > M.m[i];
    return true;
  };
  auto sub_m = [&](std::vector<uint32_t>& v) {
    uint64_t br = 0;
    for (int i = 0; i < L; i++) { uint64_t d = (uint64_t)v[i] - M.m[i] - br; v[i] = (uint32_t)d; br = (d >> 63) & 1; }
    v[L] -= (uint32_t)br;
  };
  while (ge_m(r)) sub_m(r);
  for (int pass = 0; pass < 2; pass++) {
    for (int b = 0; b < 32 * L; b++) {
      for (int i = L; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
      r[0] <<= 1;
      if (ge_m(r)) sub_m(r);
    }
    for (int i = 0; i < L; i++) (pass == 0 ? M.r1 : M.r2)[i] = r[i];
  }
}

int check_modulus(const uint8_t* m_be, uint32_t mlen) {
  if (mlen == 0 || mlen > 256) return fail(BFTQ_ERR_INVALID_ARG, "modulus length must be 1..256 bytes");
  if (!(m_be[mlen - 1] & 1)) return fail(BFTQ_ERR_INVALID_ARG, "modulus must be odd");
  bool gt1 = false;
  for (uint32_t i = 0; i + 1 < mlen; i++) gt1 = gt1 || m_be[i];
  if (!gt1 && m_be[mlen - 1] <= 1) return fail(BFTQ_ERR_INVALID_ARG, "modulus must be > 1");
  return BFTQ_OK;
}

// lambda / combine launcher on an existing arena stream (device pointers).
template <int L>
int launch_lagrange(bftq_engine* e, const uint8_t* m_be, uint32_t mlen, uint32_t k, const int32_t* d_x, const uint8_t* d_y, uint64_t n_items,
                    uint8_t* d_out, uint8_t* d_st, uint8_t* d_lambda, cudaStream_t st) {
  bftq::LagrangeMod<L> M;
  make_lagrange_mod<L>(m_be, mlen, M);
  const int block = 128;
  bftq::lagrange_combine_kernel<L><<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(M, k, d_x, d_y, n_items, d_out, d_st, d_lambda);
  CU(cudaGetLastError());
  { std::lock_guard<std::mutex> g(e->mu); e->stats.launches += 1; }
  return BFTQ_OK;
}
int launch_lagrange_any(bftq_engine* e, const uint8_t* m_be, uint32_t mlen, uint32_t k, const int32_t* d_x, const uint8_t* d_y,
                        uint64_t n_items, uint8_t* d_out, uint8_t* d_st, uint8_t* d_lambda, cudaStream_t st) {
  if (mlen <= 32) return launch_lagrange<8>(e, m_be, mlen, k, d_x, d_y, n_items, d_out, d_st, d_lambda, st);
  if (mlen <= 64) return launch_lagrange<16>(e, m_be, mlen, k, d_x, d_y, n_items, d_out, d_st, d_lambda, st);
  if (mlen <= 128) return launch_lagrange<32>(e, m_be, mlen, k, d_x, d_y, n_items, d_out, d_st, d_lambda, st);
  return launch_lagrange<64>(e, m_be, mlen, k, d_x, d_y, n_items, d_out, d_st, d_lambda, st);
}

// ---- K5 plumbing ---------------------------------------------------------------------------------
template <int W>
int make_moddev(const uint8_t* p_be, uint32_t plen, bftq::ModDev<W>& M) {
  if (plen != 16u * W) return fail(BFTQ_ERR_INVALID_ARG, "internal: modulus class mismatch");
  UBig n;
  from_be(n, p_be, plen);
  if (bitlen(n) != 128 * W || !(n.w[0] & 1)) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "exponentiation modulus must be odd with exactly 8*len bits");
  memset(&M, 0, sizeof(M));
  for (int i = 0; i < 2 * W; i++) { M.n[2 * i] = (uint32_t)n.w[i]; M.n[2 * i + 1] = (uint32_t)(n.w[i] >> 32); }
  uint32_t n0 = (uint32_t)n.w[0], inv = n0;
  for (int i = 0; i < 5; i++) inv *= 2u - n0 * inv;
  M.n0inv = 0u - inv;
  M.nbytes = plen;
  UBig y;
  memset(&y, 0, sizeof(y));
  y.w[0] = 1;
  for (int ex = 0; ex < 2 * 128 * W; ex++) dbl_mod(y, n);
  for (int i = 0; i < 2 * W; i++) { M.r2[2 * i] = (uint32_t)y.w[i]; M.r2[2 * i + 1] = (uint32_t)(y.w[i] >> 32); }
  return BFTQ_OK;
}
template <int W>
int launch_modexp(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* d_base, const uint8_t* d_exp, uint32_t elen, uint64_t n,
                  uint8_t* d_out, cudaStream_t st, uint32_t n_bases = 0) {
  bftq::ModDev<W> M;
  int rc = make_moddev<W>(p_be, plen, M);
  if (rc) return rc;
  const uint64_t per_block = 4 * 8;
  uint64_t grid = std::min<uint64_t>((n + per_block - 1) / per_block, (uint64_t)e->sm_count * 4);
  if (grid < 1) grid = 1;
  bftq::modexp_kernel<W, 128><<<(unsigned)grid, 128, 0, st>>>(M, d_base, d_exp, elen, n, d_out, n_bases);
  CU(cudaGetLastError());
  { std::lock_guard<std::mutex> g(e->mu); e->stats.launches += 1; }
  return BFTQ_OK;
}
template <int W>
int launch_modprod(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* d_vals, uint32_t k, uint64_t n, uint8_t* d_out, cudaStream_t st) {
  bftq::ModDev<W> M;
  int rc = make_moddev<W>(p_be, plen, M);
  if (rc) return rc;
  const uint64_t per_block = 4 * 8;
  uint64_t grid = std::min<uint64_t>((n + per_block - 1) / per_block, (uint64_t)e->sm_count * 4);
  if (grid < 1) grid = 1;
  bftq::modprod_kernel<W, 128><<<(unsigned)grid, 128, 0, st>>>(M, d_vals, k, n, d_out);
  CU(cudaGetLastError());
  { std::lock_guard<std::mutex> g(e->mu); e->stats.launches += 1; }
  return BFTQ_OK;
}
int modexp_any(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* d_base, const uint8_t* d_exp, uint32_t elen, uint64_t n,
               uint8_t* d_out, cudaStream_t st, uint32_t n_bases = 0) {
  if (plen == 128) return launch_modexp<8>(e, p_be, plen, d_base, d_exp, elen, n, d_out, st, n_bases);
  if (plen == 256) return launch_modexp<16>(e, p_be, plen, d_base, d_exp, elen, n, d_out, st, n_bases);
  return fail(BFTQ_ERR_UNSUPPORTED_KEY, "exponentiation modulus must be 128 or 256 bytes (1024 / 2048 bit)");
}
int modprod_any(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* d_vals, uint32_t k, uint64_t n, uint8_t* d_out, cudaStream_t st) {
  if (plen == 128) return launch_modprod<8>(e, p_be, plen, d_vals, k, n, d_out, st);
  if (plen == 256) return launch_modprod<16>(e, p_be, plen, d_vals, k, n, d_out, st);
  return fail(BFTQ_ERR_UNSUPPORTED_KEY, "exponentiation modulus must be 128 or 256 bytes (1024 / 2048 bit)");
}
}  // namespace

extern "C" {

int bftq_lagrange_combine_batch(bftq_engine* e, const uint8_t* m_be, uint32_t mlen, uint32_t k, const int32_t* x,
                                const uint8_t* y_be, uint64_t n_items, uint8_t* out_be, uint8_t* out_status) {
  if (!e || !m_be || !x || !y_be || !out_be || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (k == 0 || k > 255) return fail(BFTQ_ERR_INVALID_ARG, "k must be 1..255");
  int rc = check_modulus(m_be, mlen);
  if (rc) return rc;
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  int32_t* d_x; uint8_t *d_y, *d_out, *d_st;
  a.in(&d_x, x, (size_t)n_items * k);
  a.in(&d_y, y_be, (size_t)n_items * k * mlen);
  a.out(&d_out, out_be, (size_t)n_items * mlen);
  a.out(&d_st, out_status, (size_t)n_items);
  rc = a.upload();
  if (rc) return rc;
  rc = launch_lagrange_any(e, m_be, mlen, k, d_x, d_y, n_items, d_out, d_st, nullptr, a.stream());
  if (rc) return rc;
  return a.download();
}

// ---- K5 ---------------------------------------------------------------------------------------
int bftq_modexp_batch(bftq_engine* e, const uint8_t* m_be, uint32_t mlen, const uint8_t* base_be, const uint8_t* exp_be, uint32_t elen,
                      uint64_t n_items, uint8_t* out_be) {
  if (!e || !m_be || !base_be || !exp_be || !out_be || elen == 0) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  uint8_t *d_b, *d_e, *d_o;
  a.in(&d_b, base_be, (size_t)n_items * mlen);
  a.in(&d_e, exp_be, (size_t)n_items * elen);
  a.out(&d_o, out_be, (size_t)n_items * mlen);
  int rc = a.upload();
  if (rc) return rc;
  rc = modexp_any(e, m_be, mlen, d_b, d_e, elen, n_items, d_o, a.stream());
  if (rc) return rc;
  return a.download();
}

int bftq_modprod_batch(bftq_engine* e, const uint8_t* m_be, uint32_t mlen, uint32_t k, const uint8_t* vals_be, uint64_t n_items,
                       uint8_t* out_be) {
  if (!e || !m_be || !vals_be || !out_be || k == 0) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  uint8_t *d_v, *d_o;
  a.in(&d_v, vals_be, (size_t)n_items * k * mlen);
  a.out(&d_o, out_be, (size_t)n_items * mlen);
  int rc = a.upload();
  if (rc) return rc;
  rc = modprod_any(e, m_be, mlen, d_v, k, n_items, d_o, a.stream());
  if (rc) return rc;
  return a.download();
}

// prod_i Y_i^lambda_i mod p with lambda_i = Lagrange(x_i, xs, q): K3 (lambda) -> K5 modexp -> K5 product.
static int lagrange_exp_product_dev(bftq_engine* e, Arena& a, const uint8_t* p_be, uint32_t plen, const uint8_t* q_be, uint32_t qlen, uint32_t k,
                                    const int32_t* d_x, const uint8_t* d_y, uint64_t n_items, uint8_t* d_lambda, uint8_t* d_pow, uint8_t* d_dummy_out,
                                    uint8_t* d_st, uint8_t* d_prod) {
  // lambda only: feed the combine kernel zero shares (d_pow is zero-initialised scratch of sufficient size is not
  // needed: y is read but its product is discarded) — reuse d_y's first bytes as y is only multiplied in.
  int rc = launch_lagrange_any(e, q_be, qlen, k, d_x, d_lambda /* any readable k*qlen bytes per item */, n_items, d_dummy_out, d_st, d_lambda, a.stream());
  if (rc) return rc;
  rc = modexp_any(e, p_be, plen, d_y, d_lambda, qlen, n_items * k, d_pow, a.stream());
  if (rc) return rc;
  return modprod_any(e, p_be, plen, d_pow, k, n_items, d_prod, a.stream());
}

int bftq_lagrange_exp_product_batch(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* q_be, uint32_t qlen, uint32_t k,
                                    const int32_t* x, const uint8_t* y_be, uint64_t n_items, uint8_t* out_be, uint8_t* out_status) {
  if (!e || !p_be || !q_be || !x || !y_be || !out_be || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (k == 0 || k > 255) return fail(BFTQ_ERR_INVALID_ARG, "k must be 1..255");
  int rc = check_modulus(q_be, qlen);
  if (rc) return rc;
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  int32_t* d_x; uint8_t *d_y, *d_lam, *d_pow, *d_tmp, *d_st, *d_out;
  a.in(&d_x, x, (size_t)n_items * k);
  a.in(&d_y, y_be, (size_t)n_items * k * plen);
  a.out(&d_lam, (uint8_t*)nullptr, (size_t)n_items * k * qlen, 0);
  a.out(&d_pow, (uint8_t*)nullptr, (size_t)n_items * k * plen, 0);
  a.out(&d_tmp, (uint8_t*)nullptr, (size_t)n_items * qlen, 0);
  a.out(&d_st, out_status, (size_t)n_items);
  a.out(&d_out, out_be, (size_t)n_items * plen);
  rc = a.upload();
  if (rc) return rc;
  CU(cudaMemsetAsync(d_lam, 0, (size_t)n_items * k * qlen, a.stream()));
  rc = lagrange_exp_product_dev(e, a, p_be, plen, q_be, qlen, k, d_x, d_y, n_items, d_lam, d_pow, d_tmp, d_st, d_out);
  if (rc) return rc;
  return a.download();
}

int bftq_dsa_calculate_r_batch(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* q_be, uint32_t qlen, uint32_t k,
                               const int32_t* x, const uint8_t* ri_be, const uint8_t* vi_be, uint64_t n_items, uint8_t* out_r_be,
                               uint8_t* out_status) {
  if (!e || !p_be || !q_be || !x || !ri_be || !vi_be || !out_r_be || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (k == 0 || k > 255) return fail(BFTQ_ERR_INVALID_ARG, "k must be 1..255");
  if (qlen > 32) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "subgroup order longer than 256 bits");
  int rc = check_modulus(q_be, qlen);
  if (rc) return rc;
  if (n_items == 0) return BFTQ_OK;
  Arena a(e);
  int32_t* d_x; uint8_t *d_ri, *d_vi, *d_lam, *d_pow, *d_tmp, *d_st, *d_st2, *d_prod, *d_v, *d_vinv, *d_rp, *d_out;
  a.in(&d_x, x, (size_t)n_items * k);
  a.in(&d_ri, ri_be, (size_t)n_items * k * plen);
  a.in(&d_vi, vi_be, (size_t)n_items * k * qlen);
  a.out(&d_lam, (uint8_t*)nullptr, (size_t)n_items * k * qlen, 0);
  a.out(&d_pow, (uint8_t*)nullptr, (size_t)n_items * k * plen, 0);
  a.out(&d_tmp, (uint8_t*)nullptr, (size_t)n_items * qlen, 0);
  a.out(&d_prod, (uint8_t*)nullptr, (size_t)n_items * plen, 0);
  a.out(&d_v, (uint8_t*)nullptr, (size_t)n_items * qlen, 0);
  a.out(&d_vinv, (uint8_t*)nullptr, (size_t)n_items * qlen, 0);
  a.out(&d_rp, (uint8_t*)nullptr, (size_t)n_items * plen, 0);
  a.out(&d_st2, (uint8_t*)nullptr, (size_t)n_items, 0);
  a.out(&d_st, out_status, (size_t)n_items);
  a.out(&d_out, out_r_be, (size_t)n_items * qlen);
  rc = a.upload();
  if (rc) return rc;
  cudaStream_t st = a.stream();
  CU(cudaMemsetAsync(d_lam, 0, (size_t)n_items * k * qlen, st));
  // r' = prod R_i^lambda_i mod p                                    (dsa.go:41-46)
  rc = lagrange_exp_product_dev(e, a, p_be, plen, q_be, qlen, k, d_x, d_ri, n_items, d_lam, d_pow, d_tmp, d_st, d_prod);
  if (rc) return rc;
  // v = sum v_i lambda_i mod q                                      (dsa.go:47-48)
  rc = launch_lagrange_any(e, q_be, qlen, k, d_x, d_vi, n_items, d_v, d_st2, nullptr, st);
  if (rc) return rc;
  // v^-1 mod q (q prime: Fermat)                                    (dsa.go:50)
  {
    bftq::LagrangeMod<8> M;
    make_lagrange_mod<8>(q_be, qlen, M);
    const int block = 128;
    bftq::fermat_inverse_kernel<8><<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(M, d_v, n_items, d_vinv, d_st);
    CU(cudaGetLastError());
    // r = r'^(v^-1) mod p                                           (dsa.go:51)
    rc = modexp_any(e, p_be, plen, d_prod, d_vinv, qlen, n_items, d_rp, st);
    if (rc) return rc;
    // r mod q                                                       (dsa.go:52)
    bftq::mod_small_kernel<8><<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(M, d_rp, plen, n_items, d_out);
    CU(cudaGetLastError());
    std::lock_guard<std::mutex> g(e->mu);
    e->stats.launches += 2;
  }
  return a.download();
}

int bftq_ecdsa_p256_calculate_r_batch(bftq_engine* e, uint32_t k, const int32_t* x, const uint8_t* ri, const uint8_t* vi_be,
                                      uint64_t n_items, uint8_t* out_r_be, uint8_t* out_status) {
  if (!e || !x || !ri || !vi_be || !out_r_be || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (k == 0 || k > 255) return fail(BFTQ_ERR_INVALID_ARG, "k must be 1..255");
  if (n_items == 0) return BFTQ_OK;
  static const uint8_t kN[32] = {0xff, 0xff, 0xff, 0xff, 0x00, 0x00, 0x00, 0x00, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff,
                                 0xbc, 0xe6, 0xfa, 0xad, 0xa7, 0x17, 0x9e, 0x84, 0xf3, 0xb9, 0xca, 0xc2, 0xfc, 0x63, 0x25, 0x51};
  Arena a(e);
  int32_t* d_x; uint8_t *d_ri, *d_vi, *d_lam, *d_tmp, *d_st, *d_st2, *d_v, *d_vinv, *d_ok, *d_out; uint32_t* d_jac;
  a.in(&d_x, x, (size_t)n_items * k);
  a.in(&d_ri, ri, (size_t)n_items * k * 65);
  a.in(&d_vi, vi_be, (size_t)n_items * k * 32);
  a.out(&d_lam, (uint8_t*)nullptr, (size_t)n_items * k * 32, 0);
  a.out(&d_tmp, (uint8_t*)nullptr, (size_t)n_items * 32, 0);
  a.out(&d_v, (uint8_t*)nullptr, (size_t)n_items * 32, 0);
  a.out(&d_vinv, (uint8_t*)nullptr, (size_t)n_items * 32, 0);
  a.out(&d_jac, (uint32_t*)nullptr, (size_t)n_items * k * 24, 0);
  a.out(&d_ok, (uint8_t*)nullptr, (size_t)n_items * k, 0);
  a.out(&d_st2, (uint8_t*)nullptr, (size_t)n_items, 0);
  a.out(&d_st, out_status, (size_t)n_items);
  a.out(&d_out, out_r_be, (size_t)n_items * 32);
  int rc = a.upload();
  if (rc) return rc;
  cudaStream_t st = a.stream();
  CU(cudaMemsetAsync(d_lam, 0, (size_t)n_items * k * 32, st));
  rc = launch_lagrange_any(e, kN, 32, k, d_x, d_lam, n_items, d_tmp, d_st2, d_lam, st);          // lambda_i mod N
  if (rc) return rc;
  const int block = 128;
  bftq::p256_scalar_mul_kernel<<<(unsigned)((n_items * k + block - 1) / block), block, 0, st>>>(d_ri, d_lam, n_items * k, d_jac, d_ok);
  CU(cudaGetLastError());
  rc = launch_lagrange_any(e, kN, 32, k, d_x, d_vi, n_items, d_v, d_st2, nullptr, st);             // v = sum v_i lambda_i
  if (rc) return rc;
  bftq::LagrangeMod<8> M;
  make_lagrange_mod<8>(kN, 32, M);
  bftq::fermat_inverse_kernel<8><<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(M, d_v, n_items, d_vinv, nullptr);
  CU(cudaGetLastError());
  bftq::p256_sum_mul_kernel<<<(unsigned)((n_items + block - 1) / block), block, 0, st>>>(d_jac, d_ok, k, d_vinv, n_items, d_out, d_st);
  CU(cudaGetLastError());
  { std::lock_guard<std::mutex> g(e->mu); e->stats.launches += 3; }
  return a.download();
}

// ---- K1d: DSA verify ----------------------------------------------------------------------------
}  // extern "C"
namespace {
// Domain check for one DSA key: p odd with exactly 1024 / 2048 bits (K5's classes), q odd, at most 256 bits.
// Returns 0 usable, 1 dsa.Verify is false for every signature (q's bit length is not a multiple of 8), 2 not built.
int dsa_key_class(const uint8_t* p_be, uint32_t plen, const uint8_t* q_be, uint32_t qlen) {
  if (!(plen == 128 || plen == 256) || qlen == 0 || qlen > 32) return 2;
  if (!(p_be[0] & 0x80) || !(p_be[plen - 1] & 1) || !(q_be[qlen - 1] & 1)) return 2;
  if (!(q_be[0] & 0x80)) return q_be[0] == 0 ? 2 : 1;
  return 0;
}
// d_r / d_s: 32-byte right-aligned values `stride` apart; d_bases: g || y (2 x plen); scratch: d_u (2n x qlen),
// d_pow (2n x plen), d_prod (n x plen).  d_pre may be NULL.
int dsa_verify_dev(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* q_be, uint32_t qlen, const uint8_t* d_bases,
                   const uint8_t* d_r, const uint8_t* d_s, uint32_t stride, const uint8_t* d_dig, uint32_t dlen, uint64_t n,
                   const uint8_t* d_pre, uint8_t* d_u, uint8_t* d_pow, uint8_t* d_prod, uint8_t* d_st, cudaStream_t st) {
  bftq::LagrangeMod<8> M;
  make_lagrange_mod<8>(q_be, qlen, M);
  const int block = 128;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  bftq::dsa_prepare_kernel<8><<<grid, block, 0, st>>>(M, d_r, d_s, stride, d_dig, dlen, n, d_pre, d_u, d_st);
  CU(cudaGetLastError());
  int rc = modexp_any(e, p_be, plen, d_bases, d_u, qlen, 2 * n, d_pow, st, 2);
  if (rc) return rc;
  rc = modprod_any(e, p_be, plen, d_pow, 2, n, d_prod, st);
  if (rc) return rc;
  bftq::dsa_finish_kernel<8><<<grid, block, 0, st>>>(M, d_prod, plen, d_r, stride, n, d_st);
  CU(cudaGetLastError());
  std::lock_guard<std::mutex> g(e->mu);
  e->stats.launches += 2;
  e->stats.items += n;
  return BFTQ_OK;
}
}  // namespace
extern "C" {

int bftq_dsa_verify_batch(bftq_engine* e, const uint8_t* p_be, uint32_t plen, const uint8_t* q_be, uint32_t qlen, const uint8_t* g_be,
                          const uint8_t* y_be, const uint8_t* r_be, const uint8_t* s_be, const uint8_t* digest, uint32_t digest_len,
                          uint64_t n_items, uint8_t* out_status) {
  if (!e || !p_be || !q_be || !g_be || !y_be || !r_be || !s_be || !digest || !out_status) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (digest_len == 0 || digest_len > 64) return fail(BFTQ_ERR_INVALID_ARG, "digest_len must be 1..64");
  const int cls = dsa_key_class(p_be, plen, q_be, qlen);
  if (cls == 2) return fail(BFTQ_ERR_UNSUPPORTED_KEY, "DSA domain: p must be odd with exactly 1024 / 2048 bits in plen bytes, q odd of at most 256 bits");
  if (n_items == 0) return BFTQ_OK;
  if (cls == 1) { memset(out_status, BFTQ_ST_BAD_SIGNATURE, (size_t)n_items); return BFTQ_OK; }   // dsa.Verify: q.BitLen() & 7 != 0
  std::vector<uint8_t> bases(2 * (size_t)plen);
  memcpy(bases.data(), g_be, plen);
  memcpy(bases.data() + plen, y_be, plen);
  Arena a(e);
  uint8_t *d_b, *d_r, *d_s, *d_dg, *d_u, *d_pow, *d_prod, *d_st;
  a.in(&d_b, bases.data(), bases.size());
  a.in(&d_r, r_be, (size_t)n_items * 32);
  a.in(&d_s, s_be, (size_t)n_items * 32);
  a.in(&d_dg, digest, (size_t)n_items * digest_len);
  a.out(&d_u, (uint8_t*)nullptr, 2 * (size_t)n_items * qlen, 0);
  a.out(&d_pow, (uint8_t*)nullptr, 2 * (size_t)n_items * plen, 0);
  a.out(&d_prod, (uint8_t*)nullptr, (size_t)n_items * plen, 0);
  a.out(&d_st, out_status, (size_t)n_items);
  int rc = a.upload();
  if (rc) return rc;
  rc = dsa_verify_dev(e, p_be, plen, q_be, qlen, d_b, d_r, d_s, 32, d_dg, digest_len, n_items, nullptr, d_u, d_pow, d_prod, d_st, a.stream());
  if (rc) return rc;
  return a.download();
}

// ---- K4 ---------------------------------------------------------------------------------------
int bftq_pgp_digest_batch(bftq_engine* e, const uint8_t* data_blob, const uint64_t* data_off, uint32_t n_data,
                          const uint32_t* data_idx, const uint8_t* suffix_blob, const uint64_t* suffix_off,
                          uint32_t hash_alg, uint64_t n_items, uint8_t* out_digest) {
  if (!e || !data_off || !suffix_off || !out_digest) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (!bftq::digest_on_device(hash_alg)) return fail(BFTQ_ERR_INVALID_ARG, "digest algorithm not built for the device (SHA-1/224/256/384/512 are)");
  const int dlen = bftq::host_hash_dlen(hash_alg);
  if (n_items == 0) return BFTQ_OK;
  if (!data_idx && n_data < n_items) return fail(BFTQ_ERR_INVALID_ARG, "data_idx is NULL but n_data < n_items");
  if (data_idx)
    for (uint64_t i = 0; i < n_items; i++)
      if (data_idx[i] >= n_data) return fail(BFTQ_ERR_INVALID_ARG, "data_idx out of range");
  const size_t dbytes = (size_t)data_off[n_data], sbytes = (size_t)suffix_off[n_items];
  if ((dbytes && !data_blob) || (sbytes && !suffix_blob)) return fail(BFTQ_ERR_INVALID_ARG, "NULL blob");
  Arena a(e);
  uint8_t *d_data, *d_suf, *d_out; uint64_t *d_doff, *d_soff; uint32_t* d_didx = nullptr;
  a.in(&d_data, data_blob, std::max<size_t>(dbytes, 1), dbytes);
  a.in(&d_suf, suffix_blob, std::max<size_t>(sbytes, 1), sbytes);
  a.in(&d_doff, data_off, (size_t)n_data + 1);
  a.in(&d_soff, suffix_off, (size_t)n_items + 1);
  if (data_idx) a.in(&d_didx, data_idx, (size_t)n_items);
  a.out(&d_out, out_digest, (size_t)n_items * dlen);
  int rc = a.upload();
  if (rc) return rc;
  CU(bftq::launch_pgp_digest(hash_alg, d_data, d_doff, d_didx, d_suf, d_soff, n_items, d_out, nullptr, nullptr, a.stream()));
  { std::lock_guard<std::mutex> g(e->mu); e->stats.launches += 1; }
  return a.download();
}

// ---- host packer ------------------------------------------------------------------------------
}  // extern "C"

struct bftq_keyring {
  bftq_engine* e = nullptr;
  std::mutex mu;
  std::vector<bftq::pgp::Entity> secring, keyring;
};

namespace {
namespace pg = bftq::pgp;

// Enter an RSA key in the engine's table (deduplicated).  Returns -1 when the size is not built.
int32_t engine_key_index(bftq_engine* e, const pg::PubKey& k) {
  if (!e) return -1;
  if (!(k.algo == 1 || k.algo == 2 || k.algo == 3)) return -1;
  if (!bftq::class_of((int)(k.nbits + 7) / 8) || k.n_be.empty() || !(k.n_be.back() & 1) || k.e == 0) return -1;
  std::string id((const char*)k.n_be.data(), k.n_be.size());
  id.append((const char*)&k.e, 4);
  {
    std::lock_guard<std::mutex> g(e->mu);
    auto it = e->key_lookup.find(id);
    if (it != e->key_lookup.end()) return (int32_t)it->second;
  }
  uint8_t n_be[512];
  memset(n_be, 0, 512);
  memcpy(n_be + 512 - k.n_be.size(), k.n_be.data(), k.n_be.size());
  uint32_t first = 0;
  if (bftq_register_rsa_keys_k(e, n_be, 512, &k.e, 1, &first) != BFTQ_OK) return -1;
  std::lock_guard<std::mutex> g(e->mu);
  e->key_lookup[id] = first;
  return (int32_t)first;
}
int32_t dsa_key_index(bftq_engine* e, const pg::PubKey& k) {
  const auto &P = k.dsa[0], &Q = k.dsa[1], &G = k.dsa[2], &Y = k.dsa[3];
  if (P.empty() || Q.empty() || G.size() > P.size() || Y.size() > P.size()) return -1;
  const int cls = dsa_key_class(P.data(), (uint32_t)P.size(), Q.data(), (uint32_t)Q.size());
  if (cls == 2) return -1;
  std::string id;
  for (int i = 0; i < 4; i++) { id.append((const char*)k.dsa[i].data(), k.dsa[i].size()); id.push_back((char)0xff); id.push_back((char)i); }
  std::lock_guard<std::mutex> g(e->mu);
  auto it = e->dsa_lookup.find(id);
  if (it != e->dsa_lookup.end()) return (int32_t)it->second;
  bftq_engine::DsaKey dk;
  dk.p = P; dk.q = Q; dk.cls = cls;
  dk.gy.assign(2 * P.size(), 0);
  memcpy(dk.gy.data() + P.size() - G.size(), G.data(), G.size());
  memcpy(dk.gy.data() + 2 * P.size() - Y.size(), Y.data(), Y.size());
  e->dsa_keys.push_back(dk);
  e->dsa_lookup[id] = (uint32_t)e->dsa_keys.size() - 1;
  return (int32_t)e->dsa_keys.size() - 1;
}
int32_t any_key_index(bftq_engine* e, const pg::PubKey& k) {
  if (!e) return -1;
  if (k.algo == 19) return k.ec_xy.size() == 64 ? 0 : -1;       // P-256 keys travel with their tuples, no table
  if (k.algo == 17) return dsa_key_index(e, k);
  return engine_key_index(e, k);
}
void index_entity_keys(bftq_engine* e, pg::Entity& ent) {
  ent.primary.table_idx = any_key_index(e, ent.primary);
  for (auto& sk : ent.subkeys) sk.key.table_idx = any_key_index(e, sk.key);
}

struct Tuple {
  uint32_t item, call;         // item = index inside the plan (chunk-local)
  int32_t key_idx;
  uint32_t kbytes;             // RSA: key-size class of the candidate key (signature is padded to it); DSA: key table index
  uint64_t signer_id;          // primary key id of the candidate key's entity
  uint32_t data_idx;
  uint32_t suffix_pos, suffix_len;   // into Plan::suffix_blob
  uint32_t sig_pos;            // into Plan::sig_blob, sig_bytes_of(alg, kbytes) bytes
  uint16_t tag;
  uint8_t pre;                 // status decided on the host (0 = ask the GPU)
  uint8_t hash_id;
  uint8_t alg;                 // 1: RSA (K1), 19: ECDSA P-256 (K1c), 17: DSA (K1d)
};
// Bytes a tuple occupies in the signature blob.  RSA: the signature padded to the key size.
// ECDSA: r (32) || s (32) || X (32) || Y (32).  DSA: r (32) || s (32).
inline uint32_t sig_bytes_of(uint8_t alg, uint32_t kbytes) { return alg == 19 ? 128u : (alg == 17 ? 64u : kbytes); }

struct Plan {
  std::vector<Tuple> tuples;
  std::vector<uint32_t> calls_per_item;     // number of CheckDetachedSignature calls that reached a known issuer
  std::vector<uint8_t> item_failed;          // Verify mode: a structural error / unknown issuer ended the stream
  std::vector<uint8_t> suffix_blob;
  std::vector<uint8_t> sig_blob;
  std::vector<uint8_t> data_blob;            // tbs strings, plus CRLF-canonicalised copies when needed
  std::vector<uint64_t> data_off;
  void reset() {
    tuples.clear(); calls_per_item.clear(); item_failed.clear(); suffix_blob.clear(); sig_blob.clear(); data_blob.clear();
    data_off.clear(); data_off.push_back(0);
  }
};

void canonical_text(const uint8_t* d, size_t n, std::vector<uint8_t>& out) {
  for (size_t i = 0; i < n; i++) {
    if (d[i] == 0x0d && i + 1 < n && d[i + 1] == 0x0a) { out.push_back(0x0d); out.push_back(0x0a); i++; }
    else if (d[i] == 0x0a) { out.push_back(0x0d); out.push_back(0x0a); }
    else out.push_back(d[i]);
  }
}

// Parses item `i`'s signature stream against `rings`.  collective = CollectiveSignature.Verify's
// tolerant loop (crypto_pgp.go:485-500), else Signature.Verify's strict one (:319-330).
void plan_item(Plan& pl, uint32_t item, const uint8_t* tbs, size_t tbs_len, const uint8_t* sig, size_t sig_len,
               const std::vector<const std::vector<pg::Entity>*>& rings, bool collective) {
  static thread_local std::vector<uint8_t> scratch;
  static thread_local std::vector<pg::KeyRef> keys;
  const uint32_t data_plain = (uint32_t)pl.data_off.size() - 1;
  pl.data_blob.insert(pl.data_blob.end(), tbs, tbs + tbs_len);
  pl.data_off.push_back(pl.data_blob.size());
  int32_t data_text = -1;
  pg::Reader r{sig, sig_len, 0};
  pg::SigPacket sp;
  uint32_t calls = 0;
  bool failed = false;
  while (r.remaining() > 0) {
    const int rc = pg::next_known_signature(r, rings, sp, keys, scratch);
    if (rc == pg::kOk) {
      const uint32_t spos = (uint32_t)pl.suffix_blob.size(), slen = (uint32_t)sp.suffix_size();
      pl.suffix_blob.resize((size_t)spos + slen);
      sp.write_suffix(pl.suffix_blob.data() + spos);
      uint32_t didx = data_plain;
      uint8_t common_pre = 0;
      // v4 and v3 (packet.SignatureV3 -> VerifySignatureV3: same digest rule, the suffix is sig type + creation time)
      if (sp.sig_type == 0x01) {
        if (data_text < 0) {
          std::vector<uint8_t> t;
          canonical_text(tbs, tbs_len, t);
          data_text = (int32_t)pl.data_off.size() - 1;
          pl.data_blob.insert(pl.data_blob.end(), t.begin(), t.end());
          pl.data_off.push_back(pl.data_blob.size());
        }
        didx = (uint32_t)data_text;
      } else if (sp.sig_type != 0x00) common_pre = BFTQ_ST_BAD_SIGNATURE;       // hashForSignature: unsupported type
      if (!common_pre && !(sp.pk_algo == 1 || sp.pk_algo == 3 || sp.pk_algo == 17 || sp.pk_algo == 19)) common_pre = BFTQ_ST_UNSUPPORTED;
      if (!common_pre && !bftq::digest_on_device(sp.hash_id)) common_pre = BFTQ_ST_UNSUPPORTED;     // RIPEMD-160: crypto.RIPEMD160 is not linked into bftkv -> "hash function unavailable"
      for (const pg::KeyRef& kr : keys) {
        Tuple t;
        t.item = item; t.call = calls; t.signer_id = kr.entity->primary.key_id;
        t.tag = (uint16_t)((sp.hash_tag[0] << 8) | sp.hash_tag[1]);
        t.data_idx = didx; t.suffix_pos = spos; t.suffix_len = slen;
        t.pre = common_pre;
        t.hash_id = sp.hash_id;
        t.key_idx = kr.key->table_idx;
        if (!t.pre && kr.key->algo != sp.pk_algo) t.pre = BFTQ_ST_BAD_SIGNATURE;   // "different algorithms"
        if (!t.pre && t.key_idx < 0) t.pre = BFTQ_ST_UNSUPPORTED;                   // key size not built
        if (t.key_idx < 0) t.key_idx = 0;
        t.alg = 1; t.kbytes = 0;
        size_t key_len = 0;
        const bool ec = sp.pk_algo == 19 && kr.key->algo == 19, dsa = sp.pk_algo == 17 && kr.key->algo == 17;
        if (ec) t.alg = 19;
        else if (dsa) { t.alg = 17; t.kbytes = (uint32_t)t.key_idx; }
        else {
          key_len = (kr.key->nbits + 7) / 8;                                          // pub.Size()
          const int cls = bftq::class_of((int)key_len);
          t.kbytes = (uint32_t)(cls ? cls : 256);                                     // travels in its size class
          if (!cls) key_len = 256;
        }
        const uint32_t nb = sig_bytes_of(t.alg, t.kbytes);
        t.sig_pos = (uint32_t)pl.sig_blob.size();
        pl.sig_blob.resize((size_t)t.sig_pos + nb);                                   // zero-filled
        uint8_t* dst = pl.sig_blob.data() + t.sig_pos;
        if (ec || dsa) {           // ecdsa.Verify / dsa.Verify: r, s >= N resp. q (any longer than 32 bytes) fail
          if (sp.r.size() > 32 || sp.s.size() > 32) { if (!t.pre) t.pre = BFTQ_ST_BAD_SIGNATURE; }
          else {
            if (sp.r.size()) memcpy(dst + 32 - sp.r.size(), sp.r.data(), sp.r.size());
            if (sp.s.size()) memcpy(dst + 64 - sp.s.size(), sp.s.data(), sp.s.size());
          }
          if (ec && kr.key->ec_xy.size() == 64) memcpy(dst + 64, kr.key->ec_xy.data(), 64);
        } else {
          if (sp.mpi.size() <= key_len) { if (sp.mpi.size()) memcpy(dst + t.kbytes - sp.mpi.size(), sp.mpi.data(), sp.mpi.size()); }   // padToKeySize
          else if (!t.pre) t.pre = BFTQ_ST_BAD_SIGNATURE;                             // len(sig) != k
        }
        pl.tuples.push_back(t);
      }
      calls++;
    } else if (collective) {
      continue;                      // errors are ignored; the offending packet (or the rest) was consumed
    } else {
      failed = true;
      break;
    }
  }
  pl.calls_per_item.push_back(calls);
  pl.item_failed.push_back(failed ? 1 : 0);
}

// One (hash algorithm, signature algorithm, key-size class) group of a plan in flight on its own
// stream: digest (K4) -> tag check -> verify (K1 / K1c / K1d).
struct GroupRun {
  uint32_t hash_alg = 0; int alg = 0; int kb = 0;
  std::vector<uint32_t> sel;                 // tuple indices of the group, plan order
  std::unique_ptr<Arena> arena;              // null: every status was decided on the host
  std::vector<uint8_t> st;                   // statuses come back here
};

// Composes the group's flat inputs directly in the staging slot's pinned memory, uploads them and
// enqueues the kernels and the status download on the slot's stream.  Does not wait.
int group_enqueue(bftq_engine* e, const Plan& pl, GroupRun& g, std::vector<uint8_t>& status, bool sleepy) {
  const uint32_t dsa_idx = (uint32_t)g.kb;
  const uint32_t hash_alg = g.hash_alg;
  const int alg = g.alg;
  const int kb = (int)sig_bytes_of((uint8_t)alg, (uint32_t)g.kb);
  const std::vector<uint32_t>& sel = g.sel;
  const size_t nt = sel.size();
  const int dlen = bftq::host_hash_dlen(hash_alg);
  if ((alg == 1 && !e->d_keys) || !bftq::digest_on_device(hash_alg)) {   // nothing verifiable: every tuple keeps its host status
    for (size_t i = 0; i < nt; i++) { const uint8_t pre = pl.tuples[sel[i]].pre; status[sel[i]] = pre ? pre : (uint8_t)BFTQ_ST_UNSUPPORTED; }
    return BFTQ_OK;
  }
  bftq_engine::DsaKey dk;
  if (alg == 17) {
    { std::lock_guard<std::mutex> lk(e->mu); if (dsa_idx < e->dsa_keys.size()) dk = e->dsa_keys[dsa_idx]; }
    if (dk.p.empty() || dk.cls != 0) {                     // q's bit length not a multiple of 8: dsa.Verify is false
      for (size_t i = 0; i < nt; i++) {
        const uint8_t pre = pl.tuples[sel[i]].pre;
        status[sel[i]] = pre ? pre : (uint8_t)(dk.p.empty() ? BFTQ_ST_UNSUPPORTED : BFTQ_ST_BAD_SIGNATURE);
      }
      return BFTQ_OK;
    }
  }
  size_t suffix_total = 0;
  for (size_t i = 0; i < nt; i++) suffix_total += pl.tuples[sel[i]].suffix_len;
  g.arena.reset(new Arena(e));
  g.st.assign(nt, 0);
  Arena& a = *g.arena;
  a.set_sleepy(sleepy);
  uint8_t *d_data, *d_suf, *d_pre, *d_sig, *d_dig, *d_st; uint64_t *d_doff, *d_soff; uint32_t *d_didx, *d_kidx; uint16_t* d_tags;
  uint8_t *h_data, *h_suf, *h_pre, *h_sig; uint64_t *h_doff, *h_soff; uint32_t *h_didx, *h_kidx; uint16_t* h_tags;
  a.stage(&d_data, &h_data, std::max<size_t>(pl.data_blob.size(), 1));
  a.stage(&d_doff, &h_doff, pl.data_off.size());
  a.stage(&d_suf, &h_suf, std::max<size_t>(suffix_total, 1));
  a.stage(&d_soff, &h_soff, nt + 1);
  a.stage(&d_didx, &h_didx, nt);
  a.stage(&d_kidx, &h_kidx, nt);
  a.stage(&d_tags, &h_tags, nt);
  a.stage(&d_pre, &h_pre, nt);
  a.stage(&d_sig, &h_sig, nt * (size_t)kb);
  a.out(&d_dig, (uint8_t*)nullptr, nt * dlen, 0);          // device-only intermediate
  uint8_t *d_gy = nullptr, *d_u = nullptr, *d_pow = nullptr, *d_prod = nullptr;
  if (alg == 17) {
    a.in(&d_gy, dk.gy.data(), dk.gy.size());
    a.out(&d_u, (uint8_t*)nullptr, 2 * nt * dk.q.size(), 0);
    a.out(&d_pow, (uint8_t*)nullptr, 2 * nt * dk.p.size(), 0);
    a.out(&d_prod, (uint8_t*)nullptr, nt * dk.p.size(), 0);
  }
  a.out(&d_st, g.st.data(), nt);
  int rc = a.prepare();
  if (rc) return rc;
  if (!pl.data_blob.empty()) memcpy(h_data, pl.data_blob.data(), pl.data_blob.size());
  memcpy(h_doff, pl.data_off.data(), pl.data_off.size() * sizeof(uint64_t));
  // A group that is the whole plan with one suffix per tuple (the common case: one signature packet per
  // item, one candidate key) takes the plan's blobs as they are.
  bool whole = nt == pl.tuples.size() && suffix_total == pl.suffix_blob.size() && nt * (size_t)kb == pl.sig_blob.size();
  if (whole) {
    size_t pos = 0;
    for (size_t i = 0; i < nt && whole; i++) { whole = pl.tuples[i].suffix_pos == pos; pos += pl.tuples[i].suffix_len; }
  }
  if (whole) {
    if (suffix_total) memcpy(h_suf, pl.suffix_blob.data(), suffix_total);
    if (nt) memcpy(h_sig, pl.sig_blob.data(), nt * (size_t)kb);
  }
  size_t spos = 0;
  for (size_t i = 0; i < nt; i++) {
    const Tuple& t = pl.tuples[sel[i]];
    h_kidx[i] = (uint32_t)t.key_idx; h_didx[i] = t.data_idx; h_tags[i] = t.tag; h_pre[i] = t.pre;
    h_soff[i] = spos;
    if (!whole) {
      memcpy(h_sig + i * (size_t)kb, pl.sig_blob.data() + t.sig_pos, kb);
      if (t.suffix_len) memcpy(h_suf + spos, pl.suffix_blob.data() + t.suffix_pos, t.suffix_len);
    }
    spos += t.suffix_len;
  }
  h_soff[nt] = spos;
  rc = a.upload();
  if (rc) return rc;
  CU(bftq::launch_pgp_digest(hash_alg, d_data, d_doff, d_didx, d_suf, d_soff, nt, d_dig, d_tags, d_pre, a.stream()));
  { std::lock_guard<std::mutex> lk(e->mu); e->stats.launches += 1; }
  if (alg == 19) {
    const int block = 128;
    bftq::ecdsa_p256_verify_kernel<<<(unsigned)((nt + block - 1) / block), block, 0, a.stream()>>>(
        d_sig, 0, nullptr, d_sig, d_sig + 32, d_dig, (uint32_t)dlen, nt, d_pre, d_st);
    CU(cudaGetLastError());
    std::lock_guard<std::mutex> lk(e->mu);
    e->stats.launches += 1;
    e->stats.items += nt;
  } else if (alg == 17) {
    rc = dsa_verify_dev(e, dk.p.data(), (uint32_t)dk.p.size(), dk.q.data(), (uint32_t)dk.q.size(), d_gy, d_sig, d_sig + 32, 64, d_dig,
                        (uint32_t)dlen, nt, d_pre, d_u, d_pow, d_prod, d_st, a.stream());
    if (rc) return rc;
  } else {
    rc = launch_rsa_any(e, d_kidx, d_sig, d_dig, hash_alg, nt, 0, d_pre, d_st, a.stream(), kb);
    if (rc) return rc;
  }
  return a.download_async();
}

int group_finish(GroupRun& g, std::vector<uint8_t>& status) {
  if (!g.arena) return BFTQ_OK;
  int rc = g.arena->finish();
  for (size_t i = 0; i < g.sel.size(); i++) status[g.sel[i]] = g.st[i];
  g.arena.reset();
  return rc;
}

// A plan (one chunk of a batch call) on the device: its groups run on one stream each.
struct PlanRun {
  Plan pl;
  std::vector<GroupRun> groups;
  std::vector<uint8_t> status;               // per tuple, BFTQ_ST_*
  uint64_t lo = 0, hi = 0;                   // the batch items [lo, hi) this plan covers
};

void split_groups(PlanRun& pr) {
  const Plan& pl = pr.pl;
  pr.groups.clear();
  pr.status.assign(pl.tuples.size(), 0);
  for (size_t i = 0; i < pl.tuples.size(); i++) {
    const Tuple& t = pl.tuples[i];
    GroupRun* g = nullptr;
    for (auto& c : pr.groups) if (c.hash_alg == t.hash_id && c.alg == (int)t.alg && c.kb == (int)t.kbytes) { g = &c; break; }
    if (!g) { pr.groups.emplace_back(); g = &pr.groups.back(); g->hash_alg = t.hash_id; g->alg = t.alg; g->kb = (int)t.kbytes; }
    g->sel.push_back((uint32_t)i);
  }
}
int plan_enqueue(bftq_engine* e, PlanRun& pr, bool sleepy) {
  split_groups(pr);
  for (auto& g : pr.groups) { int rc = group_enqueue(e, pr.pl, g, pr.status, sleepy); if (rc) return rc; }
  return BFTQ_OK;
}
int plan_finish(PlanRun& pr) {
  int rc = BFTQ_OK;
  for (auto& g : pr.groups) { int r = group_finish(g, pr.status); if (r && !rc) rc = r; }
  return rc;
}

// Per item: did call c succeed (any candidate tuple verified) and who signed.
struct CallResult { bool ok; uint64_t signer; };
void fold_calls(const Plan& pl, const std::vector<uint8_t>& status, std::vector<std::vector<CallResult>>& out) {
  out.assign(pl.calls_per_item.size(), {});
  for (size_t i = 0; i < out.size(); i++) out[i].assign(pl.calls_per_item[i], CallResult{false, 0});
  for (size_t t = 0; t < pl.tuples.size(); t++) {
    const Tuple& tp = pl.tuples[t];
    CallResult& cr = out[tp.item][tp.call];
    if (!cr.ok && status[t] == 0) { cr.ok = true; cr.signer = tp.signer_id; }
  }
}

// Host threads a batch call may use: the CPU quota of the container (not the core count of the box),
// capped at 16; BFTQ_HOST_THREADS overrides.  Chunk = items per plan; BFTQ_PLAN_CHUNK overrides.
unsigned packer_threads() {
  if (const char* s = getenv("BFTQ_HOST_THREADS")) { const int x = atoi(s); if (x > 0) return (unsigned)std::min(x, 64); }
  static const unsigned n = [] {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
      char q[32]; unsigned long per = 0;
      if (fscanf(f, "%31s %lu", q, &per) == 2 && strcmp(q, "max") != 0 && per) hw = std::min<unsigned>(hw, (unsigned)std::max(1L, (long)((atol(q) + per / 2) / per)));
      fclose(f);
    }
    return std::min(hw, 16u);
  }();
  return n;
}
uint64_t packer_chunk() {
  if (const char* s = getenv("BFTQ_PLAN_CHUNK")) { const long x = atol(s); if (x > 0) return (uint64_t)x; }
  return 0;                 // 0: sized per call (run_batch)
}

// The packer's batch driver.  The batch is cut into chunks; worker threads (the caller plus helpers from the
// engine's pool) take chunks off a shared counter, and each keeps kDepth runs in flight: while the kernels of
// its earlier chunks run on their streams the thread prepares the next one, so host work (CPU) and digest +
// verify (GPU) overlap both across and inside threads.
//   submit(run, lo, hi, parse_ns)  prepares items [lo, hi) and enqueues their device work (no waiting)
//   retire(run)                    waits for the run's results and folds them into the caller's outputs
//                                  (disjoint item ranges, so no locking)
//   drain(run)                     error path: lets the run's in-flight copies land
template <typename RunT, typename Submit, typename Retire, typename Drain>
int run_chunks(bftq_engine* e, uint64_t n_items, unsigned max_threads, uint64_t default_chunk_cap, unsigned chunks_per_worker, Submit submit,
               Retire retire_fn, Drain drain) {
  const unsigned want = max_threads ? max_threads : packer_threads();
  // Chunk size: small enough that every worker gets several chunks (so its preparation overlaps the kernels
  // of its previous chunks and the GPU starts early), large enough to amortise the per-chunk driver calls.
  uint64_t chunk = packer_chunk();
  if (!chunk) chunk = std::min<uint64_t>(default_chunk_cap, std::max<uint64_t>(512, ((n_items / ((uint64_t)std::max(want, 4u) * chunks_per_worker) + 63) / 64) * 64));
  const uint64_t n_chunks = (n_items + chunk - 1) / chunk;
  const unsigned nthreads = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, n_chunks));
  const bool tracing = getenv("BFTQ_TRACE") != nullptr;
  const auto call_t0 = std::chrono::steady_clock::now();
  std::atomic<uint64_t> next{0};
  std::atomic<int> first_err{BFTQ_OK};
  std::mutex err_mu;
  std::string err_text;
  auto worker = [&]() {
    if (e) cudaSetDevice(e->device);
    constexpr int kDepth = 4;                    // runs a worker keeps in flight
    RunT runs[kDepth];
    uint64_t run_chunk[kDepth] = {0, 0, 0, 0};
    uint64_t head = 0, tail = 0;                 // runs[tail % kDepth .. head % kDepth) are in flight
    auto note = [&](int rc) {
      if (!rc) return;
      std::lock_guard<std::mutex> lk(err_mu);
      if (first_err.load() == BFTQ_OK) { first_err.store(rc); err_text = g_last_error; }
    };
    uint64_t parse_ns = 0, stage_ns = 0, wait_ns = 0, chunks = 0;
    std::vector<std::array<double, 6>> trace;     // BFTQ_TRACE: per chunk (index, start, parsed, enqueued, wait start, wait end) in us
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ns = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count();
    };
    auto retire = [&]() {
      RunT& pv = runs[tail % kDepth];
      const auto t0 = now();
      uint64_t w = 0;
      if (first_err.load() == BFTQ_OK) note(retire_fn(pv, w)); else drain(pv);
      wait_ns += w;
      if (tracing) for (auto& tr : trace) if ((uint64_t)tr[0] == run_chunk[tail % kDepth]) { tr[4] = ns(call_t0, t0) * 1e-3; tr[5] = tr[4] + w * 1e-3; }
      tail++;
    };
    for (;;) {
      const uint64_t c = next.fetch_add(1);
      if (c >= n_chunks || first_err.load() != BFTQ_OK) break;
      if (head - tail == kDepth) retire();
      RunT& pr = runs[head % kDepth];
      run_chunk[head % kDepth] = c;
      const auto t0 = now();
      uint64_t p_ns = 0;
      note(submit(pr, c * chunk, std::min(n_items, c * chunk + chunk), p_ns));
      const uint64_t total = ns(t0, now());
      parse_ns += p_ns; stage_ns += total > p_ns ? total - p_ns : 0; chunks++;
      if (tracing) trace.push_back({(double)c, ns(call_t0, t0) * 1e-3, (ns(call_t0, t0) + p_ns) * 1e-3, ns(call_t0, now()) * 1e-3, 0.0, 0.0});
      head++;
    }
    while (tail < head) retire();
    if (tracing) {
      std::lock_guard<std::mutex> lk(err_mu);
      for (auto& tr : trace)
        fprintf(stderr, "bftq-trace chunk %4d thread %zu parse %8.1f..%8.1f enqueued %8.1f wait %8.1f..%8.1f us\n", (int)tr[0],
                std::hash<std::thread::id>()(std::this_thread::get_id()) % 1000, tr[1], tr[2], tr[3], tr[4], tr[5]);
    }
    if (e) {
      std::lock_guard<std::mutex> lk(e->mu);
      e->stats.packer_chunks += chunks; e->stats.packer_parse_ns += parse_ns; e->stats.packer_stage_ns += stage_ns; e->stats.packer_wait_ns += wait_ns;
    }
  };
  // The caller works too; nthreads - 1 helpers come from the engine's pool (concurrent calls share its threads,
  // oldest job first).  Without an engine (parse-only diagnostics) plain threads do.
  if (nthreads <= 1) worker();
  else if (e) e->pool.run(nthreads - 1, worker, [&] { return next.load() < n_chunks; });
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t + 1 < nthreads; t++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
  }
  if (first_err.load() != BFTQ_OK) return fail(first_err.load(), err_text);
  return BFTQ_OK;
}

// Host-packer form: build(lo, hi, plan) parses the items, done(planrun) folds the statuses.  e == nullptr: parse only.
template <typename Build, typename Done>
int run_batch(bftq_engine* e, uint64_t n_items, unsigned max_threads, Build build, Done done) {
  // Waiting for a chunk: spinning in cudaStreamSynchronize measured 31 M/s against 24 M/s with a blocking-sync
  // event (tools/pgp_e2e_experiment.py) — the wake-up latency costs more than the spinning; BFTQ_BLOCKING_SYNC=1
  // selects the sleeping wait for hosts where the CPU quota is the scarcer resource.
  const bool sleepy = [] { const char* v = getenv("BFTQ_BLOCKING_SYNC"); return v && atoi(v) > 0; }();
  auto tick = [] { return std::chrono::steady_clock::now(); };
  auto since = [](std::chrono::steady_clock::time_point a) {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - a).count();
  };
  return run_chunks<PlanRun>(
      e, n_items, max_threads, 4096, 4,
      [&](PlanRun& pr, uint64_t lo, uint64_t hi, uint64_t& parse_ns) {
        pr.lo = lo; pr.hi = hi;
        pr.pl.reset();
        const auto t0 = tick();
        build(lo, hi, pr.pl);
        parse_ns = since(t0);
        if (e) return plan_enqueue(e, pr, sleepy);
        split_groups(pr);
        return (int)BFTQ_OK;
      },
      [&](PlanRun& pr, uint64_t& wait_ns) {
        const auto t0 = tick();
        const int rc = e ? plan_finish(pr) : (int)BFTQ_OK;
        wait_ns = since(t0);
        if (!rc) done(pr);
        return rc;
      },
      [&](PlanRun& pr) { for (auto& g : pr.groups) if (g.arena) { g.arena->finish(); g.arena.reset(); } });
}

int check_blobs(const void* blob, const uint64_t* off, uint64_t n) {
  if (!off) return 1;
  for (uint64_t i = 0; i < n; i++) if (off[i + 1] < off[i]) return 1;
  if (off[n] > off[0] && !blob) return 1;
  return 0;
}
}  // namespace

extern "C" {

int bftq_keyring_create(bftq_engine* e, bftq_keyring** out) {
  if (!out) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");      // e == NULL: parse-only keyring (no device, no verification)
  auto* kr = new bftq_keyring();
  kr->e = e;
  *out = kr;
  return BFTQ_OK;
}
void bftq_keyring_destroy(bftq_keyring* kr) { delete kr; }

int bftq_keyring_add(bftq_keyring* kr, const uint8_t* key_blocks, uint64_t len, int priv, uint32_t* n_entities) {
  if (!kr || (len && !key_blocks)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::vector<pg::Entity> ents;
  pg::read_entities(key_blocks, (size_t)len, ents);
  for (auto& en : ents) index_entity_keys(kr->e, en);
  std::lock_guard<std::mutex> g(kr->mu);
  auto& ring = priv ? kr->secring : kr->keyring;
  for (auto& en : ents) {                                     // replace(), crypto_pgp.go:124-140
    bool replaced = false;
    for (auto& old : ring) if (old.primary.key_id == en.primary.key_id) { old = en; replaced = true; break; }
    if (!replaced) ring.push_back(en);
  }
  if (n_entities) *n_entities = (uint32_t)ents.size();
  return BFTQ_OK;
}

int bftq_keyring_remove(bftq_keyring* kr, const uint64_t* key_ids, uint32_t n) {
  if (!kr || (n && !key_ids)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> g(kr->mu);
  std::vector<pg::Entity> keep;
  for (auto& en : kr->keyring) {
    bool drop = false;
    for (uint32_t i = 0; i < n; i++) drop = drop || key_ids[i] == en.primary.key_id;
    if (!drop) keep.push_back(en);
  }
  kr->keyring.swap(keep);
  return BFTQ_OK;
}

int bftq_keyring_ids(bftq_keyring* kr, uint64_t* out_ids, uint32_t cap, uint32_t* n) {
  if (!kr || !n) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> g(kr->mu);
  uint32_t c = 0;
  for (auto* ring : {&kr->secring, &kr->keyring})
    for (auto& en : *ring) { if (out_ids && c < cap) out_ids[c] = en.primary.key_id; c++; }
  *n = c;
  return BFTQ_OK;
}

int bftq_keyring_certifiers(bftq_keyring* kr, uint64_t key_id, uint64_t* out_ids, uint32_t cap, uint32_t* n) {
  if (!kr || !n) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> g(kr->mu);
  *n = 0;
  for (auto* ring : {&kr->keyring, &kr->secring})            // getCertById order, crypto_pgp.go:206-219
    for (auto& en : *ring)
      if (en.primary.key_id == key_id) {
        uint32_t c = 0;
        for (uint64_t id : en.certifiers) { if (out_ids && c < cap) out_ids[c] = id; c++; }
        *n = c;
        return BFTQ_OK;
      }
  return fail(BFTQ_ERR_INVALID_ARG, "key id not in keyring");
}

// ---- GPU-parsed fast path of Signature.Verify's batch form (K0, pgp_parse.cuh) -------------------------
namespace {

// EntityList.KeysByIdUsage(id, KeyFlagSign) for every key id of the keyring, flattened for the device: an id
// with exactly one usable RSA key of the 2048-bit size class is decided on the GPU, any other id that has
// candidates is left to the host packer, and an id that is not in the table has no candidates at all.
std::vector<bftq::IssuerEntry> build_issuer_table(const std::vector<const std::vector<pg::Entity>*>& rings) {
  std::vector<bftq::IssuerEntry> tab;
  std::vector<pg::KeyRef> keys;
  auto add = [&](uint64_t id) {
    for (auto& en : tab) if (en.key_id == id) return;
    pg::keys_by_id_usage(rings, id, pg::kKeyFlagSign, keys);
    if (keys.empty()) return;
    bftq::IssuerEntry en{};
    en.key_id = id; en.kind = 1;
    const pg::PubKey& k = *keys[0].key;
    const int kb = (int)((k.nbits + 7) / 8);
    if (keys.size() == 1 && (k.algo == 1 || k.algo == 2 || k.algo == 3) && k.table_idx >= 0 && bftq::class_of(kb) == 256) {
      en.kind = 0; en.key_idx = (uint32_t)k.table_idx; en.kbytes = (uint16_t)kb; en.algo = k.algo;
    }
    tab.push_back(en);
  };
  for (auto* ring : rings)
    for (const pg::Entity& e : *ring) {
      add(e.primary.key_id);
      for (const pg::Subkey& sk : e.subkeys) add(sk.key.key_id);
    }
  return tab;
}

struct FastRun {
  std::unique_ptr<Arena> arena;
  std::vector<uint8_t> st, where;
  uint64_t lo = 0, hi = 0;
};

// Copies the chunk's raw bytes into staging and enqueues K0 (parse + digest) and K1 on the slot's stream.
int fast_enqueue(bftq_engine* e, FastRun& fr, const uint8_t* tbs_blob, const uint64_t* tbs_off, const uint8_t* sig_blob,
                 const uint64_t* sig_off, const std::vector<bftq::IssuerEntry>& table) {
  const uint64_t lo = fr.lo, hi = fr.hi;
  const size_t n = (size_t)(hi - lo);
  const uint64_t t0 = tbs_off[lo], tb = tbs_off[hi] - t0, g0 = sig_off[lo], gb = sig_off[hi] - g0;
  fr.arena.reset(new Arena(e));
  fr.st.assign(n, 0); fr.where.assign(n, 0);
  Arena& a = *fr.arena;
  uint8_t *d_tbs, *d_sig, *h_tbs, *h_sig, *d_pad, *d_dig, *d_pre, *d_st, *d_where;
  uint64_t *d_toff, *d_soff, *h_toff, *h_soff;
  bftq::IssuerEntry *d_tab, *h_tab;
  uint32_t* d_kidx;
  a.stage(&d_tbs, &h_tbs, std::max<size_t>(tb, 1));
  a.stage(&d_toff, &h_toff, n + 1);
  a.stage(&d_sig, &h_sig, std::max<size_t>(gb, 1));
  a.stage(&d_soff, &h_soff, n + 1);
  a.stage(&d_tab, &h_tab, std::max<size_t>(table.size(), 1));
  a.out(&d_kidx, (uint32_t*)nullptr, n, 0);                // device-only intermediates: K0 -> K1
  a.out(&d_pad, (uint8_t*)nullptr, n * 256, 0);
  a.out(&d_dig, (uint8_t*)nullptr, n * 32, 0);
  a.out(&d_pre, (uint8_t*)nullptr, n, 0);
  a.out(&d_st, fr.st.data(), n);
  a.out(&d_where, fr.where.data(), n);
  const bool tracing = getenv("BFTQ_TRACE") != nullptr;
  const auto c0 = std::chrono::steady_clock::now();
  int rc = a.prepare();
  if (rc) return rc;
  const auto c1 = std::chrono::steady_clock::now();
  if (tb) memcpy(h_tbs, tbs_blob + t0, tb);
  if (gb) memcpy(h_sig, sig_blob + g0, gb);
  for (size_t i = 0; i <= n; i++) { h_toff[i] = tbs_off[lo + i] - t0; h_soff[i] = sig_off[lo + i] - g0; }
  if (!table.empty()) memcpy(h_tab, table.data(), table.size() * sizeof(bftq::IssuerEntry));
  const auto c2 = std::chrono::steady_clock::now();
  rc = a.upload();
  if (rc) return rc;
  const int block = 128;
  bftq::pgp_parse_digest_kernel<<<(unsigned)((n + block - 1) / block), block, 0, a.stream()>>>(
      d_tbs, d_toff, d_sig, d_soff, (uint32_t)n, d_tab, (uint32_t)table.size(), d_kidx, d_pad, d_dig, d_pre, d_where);
  CU(cudaGetLastError());
  { std::lock_guard<std::mutex> lk(e->mu); e->stats.launches += 1; }
  rc = launch_rsa_any(e, d_kidx, d_pad, d_dig, 8, n, 0, d_pre, d_st, a.stream(), 256);
  if (rc) return rc;
  rc = a.download_async();
  if (tracing) {
    const auto c3 = std::chrono::steady_clock::now();
    auto us = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) { return std::chrono::duration<double, std::micro>(y - x).count(); };
    size_t nslots; { std::lock_guard<std::mutex> lk(e->mu); nslots = e->slots.size(); }
    fprintf(stderr, "bftq-trace fast chunk @%llu: slot %.1f us, memcpy %.1f us, enqueue %.1f us, slots %zu\n", (unsigned long long)lo, us(c0, c1), us(c1, c2), us(c2, c3), nslots);
  }
  return rc;
}

bool gpu_parse_enabled() {
  const char* v = getenv("BFTQ_GPU_PARSE");
  return !(v && atoi(v) == 0);
}

}  // namespace

static int verify_batch_impl(bftq_keyring* kr, const uint8_t* tbs_blob, const uint64_t* tbs_off, const uint8_t* sig_blob,
                             const uint64_t* sig_off, const uint8_t* cert_blob, const uint64_t* cert_off, uint64_t n_items,
                             int32_t* out_err, bool parse_only = false, unsigned threads = 0, uint64_t* n_tuples = nullptr,
                             bool no_gpu_parse = false) {
  if (!kr || (!out_err && !parse_only)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (!kr->e && !parse_only) return fail(BFTQ_ERR_NO_DEVICE, "parse-only keyring: verification needs an engine (there is no CPU fallback)");
  if (check_blobs(tbs_blob, tbs_off, n_items) || check_blobs(sig_blob, sig_off, n_items) || (cert_off && check_blobs(cert_blob, cert_off, n_items)))
    return fail(BFTQ_ERR_INVALID_ARG, "bad blob offsets");
  if (n_items == 0) return BFTQ_OK;
  std::vector<pg::Entity> sec, pub;                         // snapshot: Register / Remove may run concurrently (crypto_pgp.go:142-177)
  if (!cert_off) { std::lock_guard<std::mutex> g(kr->mu); sec = kr->secring; pub = kr->keyring; }
  const std::vector<const std::vector<pg::Entity>*> shared_rings = {&sec, &pub};
  // Fast path: the packets are parsed and hashed on the GPU (K0) and only the items it flags come back to the
  // host packer below.  Needs the RSA key table on the device; VerifyWithCertificate (a keyring per item) and
  // the parse-only diagnostic always take the host packer.
  if (!cert_off && !parse_only && !no_gpu_parse && gpu_parse_enabled() && kr->e->d_keys && n_items < 0xffffffffull) {
    const std::vector<bftq::IssuerEntry> table = build_issuer_table(shared_rings);
    std::mutex fb_mu;
    std::vector<uint64_t> fallback;
    // The host's share is one memcpy per chunk, so two workers (the caller and one helper) feed the GPU; fewer,
    // larger chunks keep K1's launches efficient.  More workers bought nothing and on some boxes halved the rate
    // (profiles/pgp_e2e_experiment_r01.json: 2 callers x 2 workers 46 M/s on every box, x 8 workers 18..42 M/s).
    unsigned fast_threads = threads ? threads : std::min(packer_threads(), 2u);
    if (const char* v = getenv("BFTQ_FAST_THREADS")) { const int x = atoi(v); if (x > 0) fast_threads = (unsigned)std::min(x, 64); }
    int rc = run_chunks<FastRun>(
        kr->e, n_items, fast_threads, 4096, 2,
        [&](FastRun& fr, uint64_t lo, uint64_t hi, uint64_t& parse_ns) {
          fr.lo = lo; fr.hi = hi; parse_ns = 0;
          return fast_enqueue(kr->e, fr, tbs_blob, tbs_off, sig_blob, sig_off, table);
        },
        [&](FastRun& fr, uint64_t& wait_ns) {
          const auto t0 = std::chrono::steady_clock::now();
          const int r = fr.arena->finish();
          wait_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
          fr.arena.reset();
          if (r) return r;
          std::vector<uint64_t> mine;
          for (uint64_t i = fr.lo; i < fr.hi; i++) {
            if (fr.where[i - fr.lo] == bftq::kParseDecided) out_err[i] = fr.st[i - fr.lo] == 0 ? 0 : BFTQ_ERR_INVALID_SIGNATURE;
            else mine.push_back(i);
          }
          if (!mine.empty()) { std::lock_guard<std::mutex> lk(fb_mu); fallback.insert(fallback.end(), mine.begin(), mine.end()); }
          return (int)BFTQ_OK;
        },
        [&](FastRun& fr) { if (fr.arena) { fr.arena->finish(); fr.arena.reset(); } });
    if (rc) return rc;
    if (n_tuples) *n_tuples = n_items - fallback.size();
    if (fallback.empty()) return BFTQ_OK;
    // the flagged items, in item order, through the host packer
    std::sort(fallback.begin(), fallback.end());
    std::vector<uint8_t> tb, sb;
    std::vector<uint64_t> to{0}, so{0};
    for (uint64_t i : fallback) {
      tb.insert(tb.end(), tbs_blob + tbs_off[i], tbs_blob + tbs_off[i + 1]); to.push_back(tb.size());
      sb.insert(sb.end(), sig_blob + sig_off[i], sig_blob + sig_off[i + 1]); so.push_back(sb.size());
    }
    std::vector<int32_t> err(fallback.size(), BFTQ_ERR_INVALID_SIGNATURE);
    rc = verify_batch_impl(kr, tb.data(), to.data(), sb.data(), so.data(), nullptr, nullptr, fallback.size(), err.data(), false, threads, nullptr, true);
    if (rc) return rc;
    for (size_t j = 0; j < fallback.size(); j++) out_err[fallback[j]] = err[j];
    return BFTQ_OK;
  }
  std::atomic<uint64_t> tuples{0};
  auto build = [&](uint64_t lo, uint64_t hi, Plan& pl) {
    std::vector<pg::Entity> cert_ring;                      // VerifyWithCertificate: a one-entity ring per item
    std::vector<const std::vector<pg::Entity>*> rings;
    for (uint64_t i = lo; i < hi; i++) {
      if (cert_off) {
        cert_ring.clear();
        std::vector<pg::Entity> ents;
        pg::read_entities(cert_blob + cert_off[i], (size_t)(cert_off[i + 1] - cert_off[i]), ents);
        if (!ents.empty()) { index_entity_keys(kr->e, ents[0]); cert_ring.push_back(ents[0]); }
        rings = {&cert_ring};
      }
      // (the per-item ring dies with this iteration: a plan keeps copies, never pointers into it; an
      // empty ring makes every issuer unknown, i.e. the item fails like a missing certificate must)
      plan_item(pl, (uint32_t)(i - lo), tbs_blob + tbs_off[i], (size_t)(tbs_off[i + 1] - tbs_off[i]), sig_blob + sig_off[i],
                (size_t)(sig_off[i + 1] - sig_off[i]), cert_off ? rings : shared_rings, false);
    }
    tuples.fetch_add(pl.tuples.size());
  };
  auto done = [&](PlanRun& pr) {
    if (!out_err) return;
    const Plan& pl = pr.pl;
    const size_t n = (size_t)(pr.hi - pr.lo);
    // Verify: every call must succeed and there must be at least one ("at least we need one valid signature")
    std::vector<uint8_t> call_ok;
    std::vector<uint32_t> call_base(n + 1, 0);
    for (size_t i = 0; i < n; i++) call_base[i + 1] = call_base[i] + pl.calls_per_item[i];
    call_ok.assign(call_base[n], 0);
    for (size_t t = 0; t < pl.tuples.size(); t++)
      if (pr.status[t] == 0) call_ok[call_base[pl.tuples[t].item] + pl.tuples[t].call] = 1;
    for (size_t i = 0; i < n; i++) {
      bool ok = !pl.item_failed[i] && pl.calls_per_item[i] > 0;
      for (uint32_t c = call_base[i]; c < call_base[i + 1]; c++) ok = ok && call_ok[c];
      out_err[pr.lo + i] = ok ? 0 : BFTQ_ERR_INVALID_SIGNATURE;
    }
  };
  int rc = run_batch(parse_only ? nullptr : kr->e, n_items, threads, build, done);
  if (n_tuples) *n_tuples = tuples.load();
  return rc;
}

int bftq_signature_parse(bftq_keyring* kr, const uint8_t* sig, uint64_t sig_len, int collective, uint64_t* out_issuers,
                         uint8_t* out_hash_ids, uint32_t cap, uint32_t* n_calls, int32_t* failed) {
  if (!kr || !n_calls || !failed || (sig_len && !sig)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::vector<pg::Entity> sec, pub;
  { std::lock_guard<std::mutex> g(kr->mu); sec = kr->secring; pub = kr->keyring; }
  std::vector<const std::vector<pg::Entity>*> rings = {&sec, &pub};
  pg::Reader r{sig, (size_t)sig_len, 0};
  std::vector<uint8_t> scratch;
  std::vector<pg::KeyRef> keys;
  pg::SigPacket sp;
  uint32_t calls = 0;
  *failed = 0;
  while (r.remaining() > 0) {
    const int rc = pg::next_known_signature(r, rings, sp, keys, scratch);
    if (rc == pg::kOk) {
      if (calls < cap) { if (out_issuers) out_issuers[calls] = sp.issuer; if (out_hash_ids) out_hash_ids[calls] = sp.hash_id; }
      calls++;
    } else if (collective) {
      continue;
    } else {
      *failed = 1;
      break;
    }
  }
  *n_calls = calls;
  return BFTQ_OK;
}

int bftq_signature_verify_batch(bftq_keyring* kr, const uint8_t* tbs_blob, const uint64_t* tbs_off, const uint8_t* sig_blob,
                                const uint64_t* sig_off, uint64_t n_items, int32_t* out_err) {
  return verify_batch_impl(kr, tbs_blob, tbs_off, sig_blob, sig_off, nullptr, nullptr, n_items, out_err);
}
int bftq_signature_verify_with_cert_batch(bftq_keyring* kr, const uint8_t* tbs_blob, const uint64_t* tbs_off,
                                          const uint8_t* sig_blob, const uint64_t* sig_off, const uint8_t* cert_blob,
                                          const uint64_t* cert_off, uint64_t n_items, int32_t* out_err) {
  if (!cert_off) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  return verify_batch_impl(kr, tbs_blob, tbs_off, sig_blob, sig_off, cert_blob, cert_off, n_items, out_err);
}

int bftq_signature_plan_measure(bftq_keyring* kr, const uint8_t* tbs_blob, const uint64_t* tbs_off, const uint8_t* sig_blob,
                                const uint64_t* sig_off, uint64_t n_items, uint32_t threads, uint64_t* n_tuples, double* seconds) {
  const auto t0 = std::chrono::steady_clock::now();
  int rc = verify_batch_impl(kr, tbs_blob, tbs_off, sig_blob, sig_off, nullptr, nullptr, n_items, nullptr, true, threads, n_tuples);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

// Signers(): every parseable v4 signature packet whose issuer is a PRIMARY key id of the keyring.
static int signers_impl(bftq_keyring* kr, const uint8_t* sig, uint64_t len, std::vector<uint64_t>& ids) {
  std::vector<pg::Entity> sec, pub;
  { std::lock_guard<std::mutex> g(kr->mu); sec = kr->secring; pub = kr->keyring; }
  pg::Reader r{sig, (size_t)len, 0};
  std::vector<uint8_t> scratch;
  for (;;) {
    int tag; const uint8_t* body; size_t bl;
    const int rc = pg::read_packet(r, tag, body, bl, scratch);
    if (rc) break;                                            // EOF or framing error ends r.Next()'s loop
    if (!pg::known_tag(tag) || tag != 2) continue;
    pg::SigPacket sp;
    if (pg::parse_signature(body, bl, sp)) break;             // parse error: r.Next() returns err -> break
    if (sp.version != 4) continue;                            // *SignatureV3 is not in the type switch
    if (!sp.has_issuer) return fail(BFTQ_ERR_MALFORMED, "signature without issuer (the reference dereferences nil here)");
    bool found = false;                                       // getCertById: keyring first, then secring
    for (auto& en : pub) found = found || en.primary.key_id == sp.issuer;
    for (auto& en : sec) found = found || en.primary.key_id == sp.issuer;
    if (found) ids.push_back(sp.issuer);
  }
  return BFTQ_OK;
}

int bftq_signature_signers(bftq_keyring* kr, const uint8_t* sig, uint64_t sig_len, uint64_t* out_ids, uint32_t cap, uint32_t* n) {
  if (!kr || !n || (sig_len && !sig)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::vector<uint64_t> ids;
  int rc = signers_impl(kr, sig, sig_len, ids);
  if (rc) return rc;
  for (uint32_t i = 0; i < ids.size() && i < cap && out_ids; i++) out_ids[i] = ids[i];
  *n = (uint32_t)ids.size();
  return BFTQ_OK;
}

// Node ids -> dense indices, quorum by index, GPU tally (K2) over the verified signers.
static int sufficient_by_tally(bftq_engine* e, const bftq_qc_ids_t* qcs, uint32_t n_qc, const uint64_t* member_ids, uint32_t n_members,
                               const std::vector<std::vector<uint64_t>>& signers, std::vector<uint8_t>& bits) {
  std::map<uint64_t, uint32_t> dense;
  auto idx_of = [&](uint64_t id) { auto it = dense.find(id); if (it != dense.end()) return it->second; uint32_t v = (uint32_t)dense.size(); dense[id] = v; return v; };
  std::vector<bftq_qc_t> q(n_qc);
  std::vector<uint32_t> members(n_members);
  for (uint32_t m = 0; m < n_members; m++) members[m] = idx_of(member_ids[m]);
  for (uint32_t c = 0; c < n_qc; c++) q[c] = bftq_qc_t{qcs[c].f, qcs[c].min, qcs[c].threshold, qcs[c].suff, qcs[c].member_off, qcs[c].member_cnt};
  bftq_quorum* qh = nullptr;
  int rc = bftq_quorum_create(e, q.data(), n_qc, members.data(), n_members, &qh);
  if (rc) return rc;
  std::vector<uint32_t> off(signers.size() + 1, 0), kidx;
  for (size_t i = 0; i < signers.size(); i++) {
    for (uint64_t id : signers[i]) kidx.push_back(idx_of(id));
    off[i + 1] = (uint32_t)kidx.size();
  }
  std::vector<uint8_t> st(std::max<size_t>(kidx.size(), 1), 0);
  if (kidx.empty()) kidx.push_back(0);
  bits.assign(signers.size(), 0);
  rc = bftq_tally_batch(e, qh, off.data(), kidx.data(), st.data(), signers.size(), bits.data());
  bftq_quorum_destroy(e, qh);
  return rc;
}

int bftq_collective_verify_batch(bftq_keyring* kr, const bftq_qc_ids_t* qcs, uint32_t n_qc, const uint64_t* member_ids,
                                 uint32_t n_members, const uint8_t* tbs_blob, const uint64_t* tbs_off, const uint8_t* ss_blob,
                                 const uint64_t* ss_off, uint64_t n_items, int32_t* out_err) {
  if (!kr || !out_err || (n_qc && !qcs) || (n_members && !member_ids)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (!kr->e) return fail(BFTQ_ERR_NO_DEVICE, "parse-only keyring: verification needs an engine (there is no CPU fallback)");
  if (check_blobs(tbs_blob, tbs_off, n_items) || check_blobs(ss_blob, ss_off, n_items)) return fail(BFTQ_ERR_INVALID_ARG, "bad blob offsets");
  if (n_items == 0) return BFTQ_OK;
  std::vector<pg::Entity> sec, pub;
  { std::lock_guard<std::mutex> g(kr->mu); sec = kr->secring; pub = kr->keyring; }
  const std::vector<const std::vector<pg::Entity>*> rings = {&sec, &pub};
  std::vector<std::vector<uint64_t>> signers(n_items);
  auto build = [&](uint64_t lo, uint64_t hi, Plan& pl) {
    for (uint64_t i = lo; i < hi; i++)
      plan_item(pl, (uint32_t)(i - lo), tbs_blob + tbs_off[i], (size_t)(tbs_off[i + 1] - tbs_off[i]), ss_blob + ss_off[i],
                (size_t)(ss_off[i + 1] - ss_off[i]), rings, true);
  };
  auto done = [&](PlanRun& pr) {
    std::vector<std::vector<CallResult>> calls;
    fold_calls(pr.pl, pr.status, calls);
    for (size_t i = 0; i < calls.size(); i++)
      for (auto& c : calls[i]) if (c.ok) signers[pr.lo + i].push_back(c.signer);      // no dedupe (crypto_pgp.go:492)
  };
  int rc = run_batch(kr->e, n_items, 0, build, done);
  if (rc) return rc;
  std::vector<uint8_t> bits;
  rc = sufficient_by_tally(kr->e, qcs, n_qc, member_ids, n_members, signers, bits);
  if (rc) return rc;
  for (uint64_t i = 0; i < n_items; i++) out_err[i] = (bits[i] & BFTQ_TALLY_IS_SUFFICIENT) ? 0 : BFTQ_ERR_INSUFFICIENT_SIGS;
  return BFTQ_OK;
}

int bftq_collective_combine_sufficient(bftq_keyring* kr, const bftq_qc_ids_t* qcs, uint32_t n_qc, const uint64_t* member_ids,
                                       uint32_t n_members, const uint8_t* ss, uint64_t ss_len, int32_t* out) {
  if (!kr || !out || (ss_len && !ss)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  if (!kr->e) return fail(BFTQ_ERR_NO_DEVICE, "parse-only keyring: the sufficiency tally runs on the GPU");
  std::vector<std::vector<uint64_t>> signers(1);
  int rc = signers_impl(kr, ss, ss_len, signers[0]);
  if (rc) return rc;
  std::vector<uint8_t> bits;
  rc = sufficient_by_tally(kr->e, qcs, n_qc, member_ids, n_members, signers, bits);
  if (rc) return rc;
  *out = (bits[0] & BFTQ_TALLY_IS_SUFFICIENT) ? 1 : 0;
  return BFTQ_OK;
}

// ---- batching aggregator ------------------------------------------------------------------------
}  // extern "C"

struct bftq_aggregator {
  struct Job {
    std::vector<uint8_t> tbs, sig, cert;
    int result = 0;
    bool done = false;
  };
  bftq_keyring* kr = nullptr;
  uint32_t max_batch = 16384, max_wait_us = 200;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::vector<std::shared_ptr<Job>> queue;
  std::chrono::steady_clock::time_point first_at;
  bool stop = false;
  uint64_t n_batches = 0, n_items = 0;
  std::thread worker;

  void flush(std::vector<std::shared_ptr<Job>>& batch) {
    // split into plain and with-certificate halves, one packer call each
    for (int with_cert = 0; with_cert < 2; with_cert++) {
      std::vector<Job*> sel;
      for (auto& j : batch) if ((j->cert.empty() ? 0 : 1) == with_cert) sel.push_back(j.get());
      if (sel.empty()) continue;
      std::vector<uint8_t> tb, sb, cb;
      std::vector<uint64_t> to{0}, so{0}, co{0};
      for (Job* j : sel) {
        tb.insert(tb.end(), j->tbs.begin(), j->tbs.end()); to.push_back(tb.size());
        sb.insert(sb.end(), j->sig.begin(), j->sig.end()); so.push_back(sb.size());
        cb.insert(cb.end(), j->cert.begin(), j->cert.end()); co.push_back(cb.size());
      }
      std::vector<int32_t> err(sel.size(), BFTQ_ERR_INVALID_SIGNATURE);
      int rc = with_cert ? bftq_signature_verify_with_cert_batch(kr, tb.data(), to.data(), sb.data(), so.data(), cb.data(), co.data(), sel.size(), err.data())
                         : bftq_signature_verify_batch(kr, tb.data(), to.data(), sb.data(), so.data(), sel.size(), err.data());
      for (size_t i = 0; i < sel.size(); i++) sel[i]->result = rc ? rc : err[i];
    }
  }
  void loop() {
    std::unique_lock<std::mutex> l(mu);
    for (;;) {
      cv_work.wait(l, [&] { return stop || !queue.empty(); });
      if (stop && queue.empty()) return;
      // wait for the batch to fill up or its deadline to pass
      const auto deadline = first_at + std::chrono::microseconds(max_wait_us);
      cv_work.wait_until(l, deadline, [&] { return stop || queue.size() >= max_batch; });
      std::vector<std::shared_ptr<Job>> batch;
      batch.swap(queue);
      l.unlock();
      flush(batch);
      l.lock();
      for (auto& j : batch) j->done = true;
      n_batches += 1;
      n_items += batch.size();
      cv_done.notify_all();
    }
  }
};

extern "C" {

int bftq_aggregator_create(bftq_keyring* kr, uint32_t max_batch, uint32_t max_wait_us, bftq_aggregator** out) {
  if (!kr || !out || max_batch == 0) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  auto* a = new bftq_aggregator();
  a->kr = kr; a->max_batch = max_batch; a->max_wait_us = max_wait_us;
  a->worker = std::thread([a] { a->loop(); });
  *out = a;
  return BFTQ_OK;
}
void bftq_aggregator_destroy(bftq_aggregator* a) {
  if (!a) return;
  { std::lock_guard<std::mutex> l(a->mu); a->stop = true; }
  a->cv_work.notify_all();
  a->worker.join();
  delete a;
}
int bftq_aggregator_verify(bftq_aggregator* a, const uint8_t* tbs, uint64_t tbs_len, const uint8_t* sig, uint64_t sig_len,
                           const uint8_t* cert, uint64_t cert_len) {
  if (!a || (tbs_len && !tbs) || (sig_len && !sig) || (cert_len && !cert)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  auto j = std::make_shared<bftq_aggregator::Job>();
  j->tbs.assign(tbs, tbs + tbs_len); j->sig.assign(sig, sig + sig_len);
  if (cert_len) j->cert.assign(cert, cert + cert_len);
  std::unique_lock<std::mutex> l(a->mu);
  if (a->stop) return fail(BFTQ_ERR_INVALID_ARG, "aggregator is shutting down");
  if (a->queue.empty()) a->first_at = std::chrono::steady_clock::now();
  a->queue.push_back(j);
  a->cv_work.notify_all();
  a->cv_done.wait(l, [&] { return j->done; });
  return j->result;
}
int bftq_aggregator_stats(bftq_aggregator* a, uint64_t* n_batches, uint64_t* n_items) {
  if (!a) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> l(a->mu);
  if (n_batches) *n_batches = a->n_batches;
  if (n_items) *n_items = a->n_items;
  return BFTQ_OK;
}

// ---- quorum-descriptor builder ------------------------------------------------------------------
}  // extern "C"
struct bftq_graph { std::mutex mu; bftq::wot::Graph g; };
extern "C" {
int bftq_graph_create(bftq_graph** out) {
  if (!out) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  *out = new bftq_graph();
  return BFTQ_OK;
}
void bftq_graph_destroy(bftq_graph* g) { delete g; }
int bftq_graph_add_node(bftq_graph* g, uint64_t id, const uint64_t* signer_ids, uint32_t n_signers) {
  if (!g || (n_signers && !signer_ids)) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> l(g->mu);
  g->g.add_node(id, signer_ids, n_signers);
  return BFTQ_OK;
}
int bftq_graph_set_self(bftq_graph* g, uint64_t id) {
  if (!g) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> l(g->mu);
  g->g.set_self(id);
  return BFTQ_OK;
}
int bftq_graph_remove_node(bftq_graph* g, uint64_t id) {
  if (!g) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> l(g->mu);
  g->g.remove_node(id);
  return BFTQ_OK;
}
int bftq_graph_revoke(bftq_graph* g, uint64_t id) {
  if (!g) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> l(g->mu);
  g->g.revoke(id);
  return BFTQ_OK;
}
int bftq_graph_choose_quorum(bftq_graph* g, int rw, bftq_qc_ids_t* out_qcs, uint32_t cap_qc, uint32_t* n_qc, uint64_t* out_members,
                             uint32_t cap_members, uint32_t* n_members) {
  if (!g || !n_qc || !n_members) return fail(BFTQ_ERR_INVALID_ARG, "NULL argument");
  std::vector<bftq::wot::QC> qcs;
  {
    std::lock_guard<std::mutex> l(g->mu);
    g->g.choose_quorum(rw, qcs);
  }
  uint32_t off = 0;
  for (size_t c = 0; c < qcs.size(); c++) {
    if (out_qcs && c < cap_qc) out_qcs[c] = bftq_qc_ids_t{qcs[c].f, qcs[c].min, qcs[c].threshold, qcs[c].suff, off, (uint32_t)qcs[c].nodes.size()};
    for (uint64_t id : qcs[c].nodes) { if (out_members && off < cap_members) out_members[off] = id; off++; }
  }
  *n_qc = (uint32_t)qcs.size();
  *n_members = off;
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
