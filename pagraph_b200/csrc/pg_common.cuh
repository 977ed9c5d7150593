// pg_common.cuh — shared helpers for the sm_100a kernels behind include/pagraph_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pagraph_b200.h"

namespace pg {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

#define PG_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      pg::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return PG_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define PG_CHECK_LAUNCH()                \
  do {                                   \
    pg::count_launch();                  \
    PG_CUDA(cudaGetLastError());         \
  } while (0)

#define PG_REQUIRE(cond, msg)                              \
  do {                                                     \
    if (!(cond)) {                                         \
      pg::set_error("%s:%d: %s", __FILE__, __LINE__, msg); \
      return PG_ERR_INVALID;                               \
    }                                                      \
  } while (0)

int sm_count(int dev);

// Live per-launch timing (pg_timing_* in the header): a TimedScope brackets the launches of one
// kernel class with a CUDA-event pair on the launching stream when timing is enabled; it costs
// nothing otherwise.
bool timing_enabled();
void timing_begin(int slot, cudaStream_t st, int* token);
void timing_end(int token, cudaStream_t st);
struct TimedScope {
  int token = -1;
  cudaStream_t st;
  TimedScope(int slot, cudaStream_t s) : st(s) {
    if (timing_enabled()) timing_begin(slot, s, &token);
  }
  ~TimedScope() {
    if (token >= 0) timing_end(token, st);
  }
};

constexpr unsigned kFullMask = 0xffffffffu;

// ------------------------------------------------------------------ Philox4x32-10 (RNG contract, oracle/pg_oracle.cpp)
struct Philox4 {
  uint32_t c[4];
};

__host__ __device__ __forceinline__ uint32_t pg_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = pg_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = pg_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// draw(v, hop, t, deg) of the contract: position in [0, deg).
__device__ __forceinline__ uint64_t draw_pos(uint32_t k0, uint32_t k1, int64_t v, uint32_t hop, uint32_t t,
                                             uint64_t deg) {
  const Philox4 r = philox4x32_10((uint32_t)(uint64_t)v, (uint32_t)((uint64_t)v >> 32), hop, t, k0, k1);
  const uint64_t x = (uint64_t)r.c[0] | ((uint64_t)r.c[1] << 32);
  return __umul64hi(x, deg);
}

// ------------------------------------------------------------------ warp / block scans (int64)
__device__ __forceinline__ int64_t warp_inclusive_scan(int64_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int64_t t = __shfl_up_sync(kFullMask, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread across the block; `total` = block sum (all threads).
// sh must hold blockDim.x/32 + 1 int64. Ends with a barrier so sh may be reused immediately.
__device__ __forceinline__ int64_t block_exclusive_scan(int64_t x, int64_t& total, int64_t* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t incl = warp_inclusive_scan(x);
  if (lane == 31) sh[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int64_t s = lane < nw ? sh[lane] : 0;
    const int64_t si = warp_inclusive_scan(s);
    if (lane < nw) sh[lane] = si - s;
    if (lane == nw - 1) sh[nw] = si;
  }
  __syncthreads();
  const int64_t r = incl - x + sh[w];
  total = sh[nw];
  __syncthreads();
  return r;
}

}  // namespace pg
