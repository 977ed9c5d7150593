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

// One shared-memory carve-out for every kernel of the library. An SM cannot hold CTAs of kernels that were configured
// with different L1 / shared-memory splits: with the default (driver-chosen) preference the sampler's small kernels ran
// under a small-shared split, and the aggregation kernel of the NEXT minibatch (192 KB of shared memory, maximum split)
// had to wait for every such CTA on an SM to drain before it could start there — and vice versa for the classifier head
// (95 KB) beside it. Measured (tools/engine_breakdown.py): the sample chain replayed beside the gather graph slowed the
// gather stage from 0.158 to 0.222 ms and the two streams ended in lock-step. Every launch site therefore asks for the
// maximum shared-memory split once per kernel. Cached per function pointer; safe under stream capture (not a stream op).
void prefer_max_smem(const void* kernel);
template <class K>
inline void prefer_max_smem_k(K kernel) { prefer_max_smem((const void*)kernel); }

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

// ------------------------------------------------------------------ aggregation from per-source row pointers (pg_aggregate.cu)
// dst[r] = scale_r * sum_{e in [indptr[r], indptr[r+1])} drop(rowptr[cols[e] - col_base][0..dim)), used by the fused
// cache-lookup + aggregation (pg_cache_aggregate): rowptr[j] points into the HBM cache table or the miss staging buffer.
// SMs left out of the fused aggregation's grid (0 = none): see launch_rows_tma_w
int agg_reserve_sms();
void set_agg_reserve_sms(int n);
struct AggRowsArgs {
  const int64_t* indptr;
  const int64_t* cols;
  int64_t col_base;
  const float* const* rowptr;  // [n_src] start of every source row (16-byte aligned for the TMA kernel); bit 0 set =
                               // "hot" row (pg_cache_set_hot): fetched with the L2 evict_last priority
  float* dst;
  int64_t dst_stride;
  int64_t n_dst;
  int64_t zero_rows_to;        // rows [n_dst, zero_rows_to) of dst are zero-filled (padding of fixed-shape buffers);
                               // a negative value -g means "up to n_dst rounded up to a multiple of g" (capped by the
                               // capacity passed as n_dst), for device-resident extents
  int dim;
  int mode;
  const float* norm;
  uint32_t drop_thr;           // drop a value when its 16-bit hash lane < drop_thr (0 = no dropout)
  float keep_scale;            // 1 / (1 - p)
  uint64_t drop_seed;
  const int64_t* drop_step;    // optional device counter added to the seed (CUDA-graph replays)
  const int64_t* lo;           // optional device-resident extents (see pg_block.d_layer_offsets): indptr/n_dst/col_base
                               // are then NodeFlow-wide base / capacity / ignored, and the kernel derives the block's
                               // own from the device (apply_extents)
  int hints;                   // 1: source rows are fetched with L2 eviction priorities (hot rows evict_last, the
                               // read-once stream evict_first); 2: hot rows evict_last, the rest default; 0: default
                               // priority for every row
};

// Device-resident block extents: lo[0..2] = NodeFlow layer offsets of the block's source layer, its destination layer
// and the layer after it. Rewrites (indptr, col_base, n_dst) in place; returns the source-layer size.
__device__ __forceinline__ int64_t apply_extents(const int64_t* lo, const int64_t*& indptr, int64_t& col_base,
                                                 int64_t& n_dst) {
  const int64_t l0 = lo[0], l1 = lo[1], l2 = lo[2];
  indptr += l1;
  col_base = l0;
  n_dst = min(n_dst, l2 - l1);
  return l1 - l0;
}
// resolves AggRowsArgs::zero_rows_to (see there) once n_dst is final; cap = the capacity n_dst held before apply_extents
__device__ __forceinline__ int64_t resolve_zero_rows(int64_t zero_rows_to, int64_t n_dst, int64_t cap) {
  if (zero_rows_to >= 0) return zero_rows_to;
  const int64_t g = -zero_rows_to;
  return min(cap, (n_dst + g - 1) / g * g);
}
pg_status launch_agg_rows(const AggRowsArgs& a, int dev, cudaStream_t st);
// tensor-core head + loss kernel of pg_dense_mma.cu (PG_ERR_INVALID = layout not eligible, nothing launched)
pg_status linear_ce_mma(const float* d_a, int64_t a_stride, const float* d_weight, const float* d_bias, const int64_t* d_labels,
                        int64_t n, int32_t in_dim, int32_t n_classes, float* d_loss, float* d_grad_a, int64_t ga_stride,
                        float* d_grad_weight, float* d_grad_bias, const int64_t* d_lo, cudaStream_t st);
pg_status block_linear_ce_mma(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_lo3, const float* d_src,
                              int64_t src_stride, int64_t cap_dst, int mode, const float* d_weight, const float* d_bias,
                              const int64_t* d_labels, int32_t in_dim, int32_t n_classes, float* d_loss, float* d_grad_src,
                              int64_t gsrc_stride, float* d_grad_weight, float* d_grad_bias, cudaStream_t st);
// tcgen05 / TMEM / TMA forward of the first NodeUpdate (pg_dense_umma.cu); PG_ERR_INVALID = not eligible, nothing launched
pg_status linear_concat_fwd_umma(const float* d_x, int64_t x_stride, const float* d_weight, const float* d_bias, int64_t n,
                                 int32_t K, int concat, float* d_out, int64_t out_stride, float* d_out_drop, int64_t od_stride,
                                 float dropout_p, uint64_t dropout_seed, const int64_t* d_step, int dev, cudaStream_t st);
pg_status linear_concat_dw_umma(const float* d_x, int64_t x_stride, const float* d_gout, int64_t g_stride, const float* d_y,
                                int64_t y_stride, int64_t n, int32_t K, int concat, float dropout_p, uint64_t dropout_seed,
                                const int64_t* d_step, float* d_gw, float* d_gb, int dev, cudaStream_t st);
// fp32-pipe dW kernel of pg_dense.cu (A/B baseline of the tensor-core kernel in pg_dense_mma.cu); outputs pre-zeroed
pg_status linear_concat_bwd_simt(const float* d_x, int64_t x_stride, const float* d_grad_out, int64_t g_stride,
                                 const float* d_out, int64_t out_stride, int64_t n, int32_t in_dim, int concat,
                                 float* d_grad_weight, float* d_grad_bias, cudaStream_t st);

// Dropout mask contract (shared with oracle.dropout_keep_mask). Element (row j, column c) of a dropped-out activation is
// decided by the (c % 4)-th 16-bit lane of
//     drop_mix(drop_rowkey(drop_stepkey(seed + step), j), drop_colkey(c / 4))
// (dropped when the lane < round(p * 65536)). Row and column keys are full splitmix64 outputs; the per-element work is
// one xor, one 64-bit multiply and one fold, so that the fused aggregation pays ~6 integer instructions per float4 in its
// inner loop (the row key is computed once per fetched row, the column keys once per thread). Hashing the step into the
// key first keeps the masks of consecutive steps unrelated.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t drop_stepkey(uint64_t seed_plus_step) { return splitmix64(seed_plus_step); }
__host__ __device__ __forceinline__ uint64_t drop_rowkey(uint64_t stepkey, uint64_t j) { return splitmix64(stepkey + j); }
__host__ __device__ __forceinline__ uint64_t drop_colkey(uint32_t g) { return splitmix64(0xD1B54A32D192ED03ull + g); }
__host__ __device__ __forceinline__ uint64_t drop_mix(uint64_t rowkey, uint64_t colkey) {
  uint64_t x = (rowkey ^ colkey) * 0x9E3779B97F4A7C15ull;
  return x ^ (x >> 32);
}
// the whole chain for one (row, 4-column group); `seed` = dropout seed + step
__host__ __device__ __forceinline__ uint64_t drop_hash(uint64_t seed, uint64_t j, uint32_t g) {
  return drop_mix(drop_rowkey(drop_stepkey(seed), j), drop_colkey(g));
}

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

// ------------------------------------------------------------------ TMA bulk copy + mbarrier helpers (sm_90+ PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
               : "memory");
}
// The same copy with an L2 eviction-priority hint (createpolicy): rows that will be read again soon are kept
// (evict_last) while the stream of read-once rows leaves first (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst)),
               "r"(src), "r"(bytes)
               : "memory");
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
