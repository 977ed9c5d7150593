// pg_runtime.cu — error state, pinned host memory, device info, PCIe probe.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "pg_common.cuh"

namespace pg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count(int dev) {
  static std::mutex mu;
  static std::unordered_map<int, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cache[dev] = n;
  return n;
}

static std::atomic<int> g_agg_reserve{0};
int agg_reserve_sms() {
  static const char* env = getenv("PG_AGG_RESERVE_SMS");   // overrides the caller's setting (experiments)
  return env ? std::max(0, atoi(env)) : g_agg_reserve.load(std::memory_order_relaxed);
}
void set_agg_reserve_sms(int n) { g_agg_reserve.store(std::max(0, n), std::memory_order_relaxed); }

void prefer_max_smem(const void* kernel) {
  static std::mutex mu;
  static std::vector<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  for (const void* k : done)
    if (k == kernel) return;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
    cudaGetLastError();   // a preference, not a requirement
  done.push_back(kernel);
}

// ------------------------------------------------------------------ live timing
struct TimingRec {
  int slot;
  cudaEvent_t a, b;
};
static std::atomic<bool> g_timing{false};
static std::mutex g_timing_mu;
static std::vector<TimingRec> g_recs;          // in launch order since the last drain
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_free_events;

bool timing_enabled() { return g_timing.load(std::memory_order_relaxed); }

void timing_begin(int slot, cudaStream_t st, int* token) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();   // a replayed graph has no host-side launch to bracket: no record for captured work
    *token = -1;
    return;
  }
  std::lock_guard<std::mutex> lk(g_timing_mu);
  TimingRec r;
  r.slot = slot;
  if (!g_free_events.empty()) {
    r.a = g_free_events.back().first;
    r.b = g_free_events.back().second;
    g_free_events.pop_back();
  } else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
    cudaGetLastError();
    *token = -1;
    return;
  }
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  *token = (int)g_recs.size() - 1;
}

void timing_end(int token, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (token >= 0 && token < (int)g_recs.size()) cudaEventRecord(g_recs[token].b, st);
}

}  // namespace pg

extern "C" {

int pg_version(void) { return 101; }

pg_status pg_timing_enable(int enabled) {
  pg::g_timing.store(enabled != 0);
  return PG_OK;
}

pg_status pg_timing_drain(int32_t* slots, float* ms, int64_t cap, int64_t* n_out) {
  PG_REQUIRE(n_out != nullptr && cap >= 0 && (cap == 0 || (slots && ms)), "pg_timing_drain: bad arguments");
  std::lock_guard<std::mutex> lk(pg::g_timing_mu);
  int64_t n = 0;
  for (const pg::TimingRec& r : pg::g_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && n < cap) {
      slots[n] = r.slot;
      ms[n] = t;
      ++n;
    }
    pg::g_free_events.emplace_back(r.a, r.b);
  }
  cudaGetLastError();
  pg::g_recs.clear();
  *n_out = n;
  return PG_OK;
}

pg_status pg_timing_drain_timeline(int32_t* slots, float* begin_ms, float* end_ms, int64_t cap, int64_t* n_out) {
  PG_REQUIRE(n_out != nullptr && cap >= 0 && (cap == 0 || (slots && begin_ms && end_ms)), "pg_timing_drain_timeline: bad arguments");
  std::lock_guard<std::mutex> lk(pg::g_timing_mu);
  int64_t n = 0;
  cudaEvent_t base = nullptr;
  for (const pg::TimingRec& r : pg::g_recs) {
    if (!base && cudaEventSynchronize(r.a) == cudaSuccess) base = r.a;
    float t0 = 0.f, t1 = 0.f;
    if (base && cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t0, base, r.a) == cudaSuccess &&
        cudaEventElapsedTime(&t1, base, r.b) == cudaSuccess && n < cap) {
      slots[n] = r.slot;
      begin_ms[n] = t0;
      end_ms[n] = t1;
      ++n;
    }
  }
  for (const pg::TimingRec& r : pg::g_recs) pg::g_free_events.emplace_back(r.a, r.b);
  cudaGetLastError();
  pg::g_recs.clear();
  *n_out = n;
  return PG_OK;
}

const char* pg_last_error(void) { return pg::g_err; }

int64_t pg_launch_count(void) { return pg::g_launches.load(); }

// Host evaluation of the dropout mask contract (pg_common.cuh drop_hash, the function the kernels inline): lets the CPU
// test-suite pin kernels and oracle to the same mask without a GPU.
pg_status pg_dropout_keep_mask(uint64_t seed_plus_step, int64_t n_rows, int32_t dim, float p, unsigned char* keep_out) {
  PG_REQUIRE(keep_out && n_rows >= 0 && dim >= 1 && p >= 0.f && p < 1.f, "pg_dropout_keep_mask: bad arguments");
  const uint32_t thr = p > 0.f ? (uint32_t)(p * 65536.0f + 0.5f) : 0u;
  const uint64_t stepkey = pg::drop_stepkey(seed_plus_step);
  for (int64_t j = 0; j < n_rows; ++j) {
    const uint64_t rk = pg::drop_rowkey(stepkey, (uint64_t)j);
    for (int32_t c = 0; c < dim; ++c) {
      const uint64_t h = pg::drop_mix(rk, pg::drop_colkey((uint32_t)(c >> 2)));
      keep_out[j * dim + c] = ((uint32_t)(h >> (16 * (c & 3))) & 0xffffu) >= thr;
    }
  }
  return PG_OK;
}

void pg_set_agg_reserve_sms(int n) { pg::set_agg_reserve_sms(n); }

pg_status pg_device_info(int dev, int* sm, size_t* total_mem, size_t* free_mem) {
  pg::DeviceGuard guard(dev);
  if (sm) *sm = pg::sm_count(dev);
  size_t f = 0, t = 0;
  PG_CUDA(cudaMemGetInfo(&f, &t));
  if (total_mem) *total_mem = t;
  if (free_mem) *free_mem = f;
  return PG_OK;
}

pg_status pg_host_alloc(void** ptr, size_t bytes) {
  PG_REQUIRE(ptr != nullptr, "pg_host_alloc: null out pointer");
  PG_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable));
  return PG_OK;
}

pg_status pg_host_free(void* ptr) {
  if (ptr) PG_CUDA(cudaFreeHost(ptr));
  return PG_OK;
}

pg_status pg_host_register(void* ptr, size_t bytes) {
  PG_REQUIRE(ptr != nullptr && bytes > 0, "pg_host_register: empty range");
  PG_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  return PG_OK;
}

pg_status pg_host_unregister(void* ptr) {
  if (ptr) PG_CUDA(cudaHostUnregister(ptr));
  return PG_OK;
}

pg_status pg_measure_h2d(int dev, size_t bytes, int iters, double* gb_per_s) {
  PG_REQUIRE(gb_per_s != nullptr && bytes > 0 && iters > 0, "pg_measure_h2d: bad arguments");
  pg::DeviceGuard guard(dev);
  void *h = nullptr, *d = nullptr;
  PG_CUDA(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  if (cudaMalloc(&d, bytes) != cudaSuccess) {
    cudaFreeHost(h);
    pg::set_error("pg_measure_h2d: cudaMalloc(%zu) failed", bytes);
    return PG_ERR_NOMEM;
  }
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0);  // warm-up
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    cudaEventRecord(a, 0);
    cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0);
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  cudaFreeHost(h);
  PG_CUDA(cudaGetLastError());
  *gb_per_s = (double)bytes / (best * 1e-3) / 1e9;
  return PG_OK;
}

}  // extern "C"
