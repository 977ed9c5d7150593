// EXPERIMENTAL — round-2 draft, compiled for sm_100a in this container but NOT yet run on hardware, not part of
// libpagraph_b200.so and not reachable from the product path. See README.md in this directory.
//
// First NodeUpdate forward (PaGraph/model/gcn_nssc.py:14-24: z = x W^T + b, out = cat(z, relu z)) as an error-compensated
// TF32 product on the 5th-generation tensor cores:
//   warp 0      TMA producer: per 32-column chunk of K one [128 rows x 32] tile of x (SWIZZLE_128B) plus the matching
//               [32 x 32] chunks of W_hi and W_lo (pre-split by split_w_kernel) into a 4-stage shared-memory ring
//   warps 8-11  transform: x_lo = x - trunc_tf32(x) written next to the TMA tile (tcgen05 kind::tf32 reads 32-bit
//               containers and ignores the low 13 mantissa bits, so the tile as loaded IS x_hi), fence.proxy.async
//   warp 1      MMA issuer (one thread): per chunk 4 k-steps x 3 terms of tcgen05.mma.cta_group::1.kind::tf32
//               (M = 128, N = 32, K = 8), accumulators in TMEM (2 stages x 32 columns), tcgen05.commit frees the stage
//   warps 4-7   epilogue: tcgen05.ld 32 columns (one output row per thread), bias / relu / concat, global stores
// Per CTA and tile the tensor pipe needs 19 chunks x 12 MMAs x 16 cycles = 1.9 us against 7 us of HBM time for the
// tile's 307 KB of x: the kernel is meant to sit on the 13.5 us HBM floor where the mma.sync version takes 45 us.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr int kBlockM = 128;       // rows per tile (UMMA M)
constexpr int kN = 32;             // outputs (UMMA N)
constexpr int kChunk = 32;         // K per stage: 32 floats = 128 bytes = one swizzle row
constexpr int kStages = 4;
constexpr int kUmmaK = 8;          // K per tcgen05.mma for 32-bit operands
constexpr int kThreads = 12 * 32;
constexpr uint32_t kABytes = kBlockM * kChunk * 4;   // 16 KB
constexpr uint32_t kBBytes = kN * kChunk * 4;        // 4 KB
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // x_hi, x_lo, w_hi, w_lo = 40 KB
constexpr uint32_t kTmemCols = 64; // 2 accumulator stages x 32 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
// K-major SWIZZLE_128B operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart, tile base 1024-byte aligned):
// cute::UMMA::SmemDescriptor with version 1, layout_type 2, LBO 1, SBO 64 (both in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major; canonical value 1)
  d |= (uint64_t)64 << 32;           // stride byte offset: 1024 bytes between 8-row groups
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // arrives on the mbarrier when all prior MMAs of this thread are done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void split_w_kernel(const float* __restrict__ W, int K, int Kpad, float* __restrict__ w_hi, float* __restrict__ w_lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kN * Kpad) return;
  const int o = i / Kpad, k = i % Kpad;
  const float v = k < K ? W[(size_t)o * K + k] : 0.f;
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);   // what kind::tf32 keeps of v
  w_hi[i] = v;        // the hardware truncates on read
  w_lo[i] = v - hi;
}

__global__ void __launch_bounds__(kThreads, 1)
    linear_fwd_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_whi,
                           const __grid_constant__ CUtensorMap tm_wlo, const float* __restrict__ bias, int64_t n, int K,
                           int concat, float* __restrict__ out, int64_t out_stride) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3 * kStages + 4];   // full[s], ready[s], empty[s], tmem_full[2], tmem_empty[2]
  __shared__ uint32_t tmem_base_sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto ready = [&](int s) { return smem_u32(&bars[kStages + s]); };
  auto empty = [&](int s) { return smem_u32(&bars[2 * kStages + s]); };
  auto tfull = [&](int a) { return smem_u32(&bars[3 * kStages + a]); };
  auto tempty = [&](int a) { return smem_u32(&bars[3 * kStages + 2 + a]); };
  const int nchunks = (K + kChunk - 1) / kChunk;
  const int64_t ntiles = (n + kBlockM - 1) / kBlockM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);         // the producer's arrive.expect_tx; the TMA completes the transaction bytes
      mbar_init(ready(s), 128);      // every transform thread
      mbar_init(empty(s), 1);        // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);        // tcgen05.commit after the tile's last MMA
      mbar_init(tempty(a), 128);     // every epilogue thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                   // TMEM: allocated and later freed by the same warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty(s), ph ^ 1);                               // a fresh barrier passes the wait for parity 1
          const uint32_t st = smem_base + s * kStageBytes;
          mbar_expect_tx(full(s), kABytes + 2 * kBBytes);
          tma_load_2d(st, &tm_x, c * kChunk, (int)(tile * kBlockM), full(s));             // x_hi  (OOB rows / cols -> 0)
          tma_load_2d(st + 2 * kABytes, &tm_whi, c * kChunk, 0, full(s));                  // w_hi
          tma_load_2d(st + 2 * kABytes + kBBytes, &tm_wlo, c * kChunk, 0, full(s));        // w_lo
        }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
        const int a = tl & 1;
        mbar_wait(tempty(a), ((tl >> 1) & 1) ^ 1);                   // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + a * kN;                       // column offset of the accumulator stage
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % kStages;
          mbar_wait(ready(s), (it / kStages) & 1);                   // TMA landed and x_lo written
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_base + s * kStageBytes;
          const uint64_t a_hi = umma_desc_k_sw128(st), a_lo = umma_desc_k_sw128(st + kABytes);
          const uint64_t b_hi = umma_desc_k_sw128(st + 2 * kABytes), b_lo = umma_desc_k_sw128(st + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < kChunk / kUmmaK; ++k) {
            const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);  // 32 bytes per k-step inside the 128-byte swizzle row
            umma_tf32(d, a_lo + adv, b_hi + adv, (c | k) != 0);      // small terms first
            umma_tf32(d, a_hi + adv, b_lo + adv, 1);
            umma_tf32(d, a_hi + adv, b_hi + adv, 1);
          }
          umma_commit(empty(s));                                     // stage reusable once these MMAs have read it
        }
        umma_commit(tfull(a));                                       // accumulator complete
      }
    }
  } else if (warp >= 8) {
    // ================================================================== transform: x_lo = x - trunc_tf32(x)
    const int tt = threadIdx.x - 8 * 32;                             // 0 .. 127
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int s = it % kStages;
        mbar_wait(full(s), (it / kStages) & 1);
        const uint32_t st = smem_base + s * kStageBytes;
        // elementwise, so the swizzled placement does not matter: same offset in the lo tile
#pragma unroll
        for (int q = 0; q < (int)(kABytes / 16 / 128); ++q) {
          const uint32_t off = (uint32_t)(q * 128 + tt) * 16;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + off));
          v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + kABytes + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(ready(s));
      }
  } else if (warp >= 4) {
    // ================================================================== epilogue: one output row per thread
    const int q = warp & 3;                                          // TMEM lane quadrant this warp may read
    uint32_t tl = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
      const int a = tl & 1;
      mbar_wait(tfull(a), (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * kN;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty(a));                                        // the MMA warp may overwrite this accumulator stage
      const int64_t row = tile * kBlockM + q * 32 + lane;
      if (row < n) {
        float* orow = out + row * out_stride;
#pragma unroll
        for (int j = 0; j < kN; j += 4) {
          float4 z = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          if (bias) {
            z.x += bias[j]; z.y += bias[j + 1]; z.z += bias[j + 2]; z.w += bias[j + 3];
          }
          const float4 p = make_float4(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f), fmaxf(z.z, 0.f), fmaxf(z.w, 0.f));
          if (concat) {
            *(float4*)(orow + j) = z;
            *(float4*)(orow + kN + j) = p;
          } else {
            *(float4*)(orow + j) = p;
          }
        }
      }
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeTiledFn enc, CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_floats,
             uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};                        // innermost first
  cuuint64_t strides[1] = {row_stride_floats * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kChunk, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

// out [n, 64 | 32] = concat ? cat(z, relu z) : relu z, z = x W^T + b. d_ws: 2 * 32 * Kpad floats of workspace (Kpad = K rounded
// up to 32). Returns 0 on success, a negative number when the driver entry point / tensor maps cannot be made, else the CUDA error.
extern "C" int pgx_linear_fwd_umma(const float* d_x, int64_t x_stride, const float* d_weight, const float* d_bias, int64_t n,
                                   int32_t K, int concat, float* d_out, int64_t out_stride, float* d_ws, void* stream) {
  if (n <= 0) return 0;
  if (K % 4 != 0 || x_stride % 4 != 0 || ((uintptr_t)d_x & 15) || ((uintptr_t)d_ws & 15)) return -1;
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -2;
    enc = (EncodeTiledFn)fn;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Kpad = (K + kChunk - 1) / kChunk * kChunk;
  float* w_hi = d_ws;
  float* w_lo = d_ws + (size_t)kN * Kpad;
  split_w_kernel<<<(kN * Kpad + 255) / 256, 256, 0, st>>>(d_weight, K, Kpad, w_hi, w_lo);
  CUtensorMap tm_x, tm_whi, tm_wlo;
  if (make_map(enc, &tm_x, d_x, (uint64_t)n, (uint64_t)K, (uint64_t)x_stride, kBlockM)) return -3;
  if (make_map(enc, &tm_whi, w_hi, kN, (uint64_t)Kpad, (uint64_t)Kpad, kN)) return -4;
  if (make_map(enc, &tm_wlo, w_lo, kN, (uint64_t)Kpad, (uint64_t)Kpad, kN)) return -5;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = (size_t)kStages * kStageBytes + 1024;
  cudaFuncSetAttribute(linear_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t ntiles = (n + kBlockM - 1) / kBlockM;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  linear_fwd_umma_kernel<<<grid, kThreads, smem, st>>>(tm_x, tm_whi, tm_wlo, d_bias, n, K, concat, d_out, out_stride);
  return (int)cudaGetLastError();
}
