// pg_dense.cu — the fp32-pipe (SIMT) kernels of the dense stage: the baselines / fallbacks of pg_dense_mma.cu.
//
// Reference: PaGraph/model/gcn_nssc.py:14-24 (NodeUpdate.forward with concat=True):
//     h = Linear(h);  h = cat(h, relu(h))
// applied to the aggregated input block x [n_1, F] (F = 600, n_1 ~ 35 k, out = 32). x needs no gradient (it is an
// aggregate of constant input features), so the backward is only dW = gz^T x and db = sum gz, with
// gz = g[:, :32] + g[:, 32:] * (z > 0). cuBLAS takes 152 us for the split-K GEMM plus four more kernels (relu',
// slice-add, bias-grad reduction, split-K reduce). linear_concat_bwd_kernel makes one pass over x: tiles of x arrive in
// shared memory by TMA bulk copies (double-buffered), relu' and the concat split are folded into a per-tile gz, every CTA
// accumulates a [32, F] partial of dW (and db) in registers (packed FFMA2) and adds it to the result with float atomics:
// 81-86 us, FMA-bound. The product path is now the 3xTF32 tensor-core kernel of pg_dense_mma.cu (51 us); this one is kept
// for A/B measurements (PG_DENSE_SIMT=1, tools/micro_dense.py). The scalar head + loss kernel below is what
// pg_linear_cross_entropy falls back to when the layout rules out 16-byte accesses (in_dim % 4 != 0).
#include <algorithm>
#include <cstdlib>

#include "pg_common.cuh"

namespace {

constexpr int kOut = 32;           // output width handled here (n_hidden of the reference default)
constexpr int kBwdThreads = 256;

// Packed fp32 FMA (Blackwell FFMA2): two fused multiply-adds per instruction on a register pair.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(uint64_t& acc, uint64_t a, uint64_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// Thread t owns columns {t, t+256, t+512} x all 32 outputs (96 accumulators, held as 48 f32x2 pairs). One persistent CTA
// per SM walks a contiguous range of rows in tiles of kTile.
template <int kTile>                 // rows of x per shared-memory tile (two tiles in flight)
__global__ void __launch_bounds__(kBwdThreads, 1) linear_concat_bwd_kernel(const float* __restrict__ x, int64_t x_stride,
                                                                          const float* __restrict__ g, int64_t g_stride,
                                                                          const float* __restrict__ y, int64_t y_stride,
                                                                          int64_t n, int K, int concat, float* dW, float* db) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(16) float gz[kTile][kOut];
  __shared__ __align__(8) uint64_t bars[2];
  float* xs = (float*)smem_raw;                       // [2][kTile][K]
  const int t = threadIdx.x;
  const uint32_t row_bytes = (uint32_t)K * 4u;
  if (t == 0) {
    pg::mbar_init(pg::smem_u32(&bars[0]), 1);
    pg::mbar_init(pg::smem_u32(&bars[1]), 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(n, r_begin + rows_per_cta);
  const int ntiles = r_end > r_begin ? (int)((r_end - r_begin + kTile - 1) / kTile) : 0;
  auto issue = [&](int tile) {                        // warp 0: one bulk copy per row of the tile
    const int buf = tile & 1;
    const int64_t r0 = r_begin + (int64_t)tile * kTile;
    const int rows = (int)min((int64_t)kTile, r_end - r0);
    const uint32_t bar = pg::smem_u32(&bars[buf]);
    if (t == 0) pg::mbar_expect_tx(bar, (uint32_t)rows * row_bytes);
    __syncwarp();
    if (t < rows) pg::bulk_g2s(pg::smem_u32(xs + ((size_t)buf * kTile + t) * K), x + (r0 + t) * x_stride, row_bytes, bar);
  };
  uint64_t acc[3][kOut / 2];                          // acc[j][p] = (dW[2p][c_j], dW[2p+1][c_j]) partials
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int o = 0; o < kOut / 2; ++o) acc[j][o] = 0ull;
  float dbv = 0.f;
  // gz for one tile: each thread owns kTile*kOut/kBwdThreads entries; loaded into registers one tile ahead so that the
  // global-load latency hides under the previous tile's FMA loop
  constexpr int kPer = kTile * kOut / kBwdThreads;
  float gnext[kPer];
  auto load_gz = [&](int tile) {
    const int64_t r0 = r_begin + (int64_t)tile * kTile;
    const int rows = (int)min((int64_t)kTile, r_end - r0);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = t + q * kBwdThreads, rr = i / kOut, o = i % kOut;
      float v = 0.f;
      if (rr < rows) {
        const float* grow = g + (r0 + rr) * g_stride;
        const float* yrow = y + (r0 + rr) * y_stride;
        v = concat ? grow[o] + (yrow[kOut + o] > 0.f ? grow[kOut + o] : 0.f) : (yrow[o] > 0.f ? grow[o] : 0.f);
      }
      gnext[q] = v;
    }
  };
  if (ntiles > 0) {
    if (t < 32) issue(0);
    load_gz(0);
  }
  uint32_t phase = 0;                                 // bit b: parity to wait for on buffer b
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    const int64_t r0 = r_begin + (int64_t)tile * kTile;
    const int rows = (int)min((int64_t)kTile, r_end - r0);
    if (tile + 1 < ntiles && t < 32) issue(tile + 1);  // the other buffer was released by the barrier ending tile-1
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = t + q * kBwdThreads;
      gz[i / kOut][i % kOut] = gnext[q];
    }
    __syncthreads();
    if (tile + 1 < ntiles) load_gz(tile + 1);          // in flight during the FMA loop below
    while (!pg::mbar_try_wait(pg::smem_u32(&bars[buf]), (phase >> buf) & 1u)) {
    }
    phase ^= 1u << buf;
    if (t < kOut)
      for (int rr = 0; rr < rows; ++rr) dbv += gz[rr][t];
    const float* xt = xs + (size_t)buf * kTile * K;
#pragma unroll 2
    for (int rr = 0; rr < rows; ++rr) {
      uint64_t xv[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int c = t + j * kBwdThreads;
        const float v = c < K ? xt[(size_t)rr * K + c] : 0.f;
        xv[j] = pack2(v, v);
      }
#pragma unroll
      for (int o4 = 0; o4 < kOut / 4; ++o4) {
        const float4 gv = *(const float4*)&gz[rr][o4 * 4];
        const uint64_t g01 = pack2(gv.x, gv.y), g23 = pack2(gv.z, gv.w);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          ffma2(acc[j][o4 * 2 + 0], xv[j], g01);
          ffma2(acc[j][o4 * 2 + 1], xv[j], g23);
        }
      }
    }
    __syncthreads();                                  // gz and this x buffer may be overwritten
  }
  if (ntiles > 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = t + j * kBwdThreads;
      if (c < K)
#pragma unroll
        for (int o = 0; o < kOut / 2; ++o) {
          float lo, hi;
          unpack2(acc[j][o], lo, hi);
          atomicAdd(&dW[(size_t)(2 * o) * K + c], lo);
          atomicAdd(&dW[(size_t)(2 * o + 1) * K + c], hi);
        }
    }
    if (t < kOut && db) atomicAdd(&db[t], dbv);
  }
}

}  // namespace

namespace pg {

// The fp32-pipe (FFMA2) dW kernel, kept for A/B measurements against the tensor-core kernel of pg_dense_mma.cu
// (PG_DENSE_SIMT=1; no dropout support). Outputs must be zeroed by the caller.
pg_status linear_concat_bwd_simt(const float* d_x, int64_t x_stride, const float* d_grad_out, int64_t g_stride,
                                 const float* d_out, int64_t out_stride, int64_t n, int32_t in_dim, int concat,
                                 float* d_grad_weight, float* d_grad_bias, cudaStream_t st) {
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  const char* env_t = getenv("PG_DENSE_TILE");
  const int tile = env_t ? atoi(env_t) : 32;
  auto launch = [&](auto kern, int kt) -> pg_status {
    const size_t smem = 2 * (size_t)kt * in_dim * sizeof(float);
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int grid = (int)std::min<int64_t>(std::max<int64_t>(1, (n + kt - 1) / kt), (int64_t)pg::sm_count(dev));
    kern<<<grid, kBwdThreads, smem, st>>>(d_x, x_stride, d_grad_out, g_stride, d_out, out_stride, n, in_dim, concat,
                                          d_grad_weight, d_grad_bias);
    PG_CHECK_LAUNCH();
    return PG_OK;
  };
  return tile >= 32 ? launch(linear_concat_bwd_kernel<32>, 32) : launch(linear_concat_bwd_kernel<16>, 16);
}

}  // namespace pg

// ====================================================================== classifier head + loss, forward and backward
// Reference: the last NodeUpdate (no activation, gcn_nssc.py:48 / :14-24) followed by torch.nn.CrossEntropyLoss
// (examples/profile/pa_gcn.py:62,93-94): pred = a W^T + b, loss = mean_r(logsumexp(pred_r) - pred_r[label_r]).
// On [6000, 64] x [60, 64] that is ~12 library kernels (GEMM, log-softmax, nll, their backwards, split-K reduce, bias
// reduction) of 4-20 us each — launch latency, not work. linear_ce_kernel does the whole thing in one pass: one warp per
// row, logits in registers (2 classes per lane), W staged in shared memory in both orientations, and it emits the
// loss, d loss / d a, d loss / d W and d loss / d b (per-CTA shared-memory reduction, then one atomic per output).
namespace {

constexpr int kHeadWarps = 8;
constexpr int kHeadMaxC = 64;   // classes (2 per lane)
constexpr int kHeadMaxK = 64;   // input width (2 per lane)

__global__ void __launch_bounds__(kHeadWarps * 32) linear_ce_kernel(const float* __restrict__ a, int64_t a_stride,
                                                                   const float* __restrict__ W, const float* __restrict__ bias,
                                                                   const int64_t* __restrict__ labels, int64_t n, int K, int C,
                                                                   float inv_n, float* loss, float* grad_a, int64_t ga_stride,
                                                                   float* dW, float* db, const int64_t* __restrict__ lo) {
  extern __shared__ __align__(16) float head_smem[];  // over the 48 KB static limit, hence dynamic
  if (lo) {  // device-resident row count: n is a capacity
    n = min(n, lo[1] - lo[0]);
    inv_n = 1.0f / (float)max(n, (int64_t)1);
  }
  // w_kc[k][c] = W[c][k] (logits: lane = class; row stride 65 so that the transposing fill is conflict-free),
  // w_ck[c][k] = W[c][k] (grad_a: lane = input column), dw_sh[c][k]: the CTA's partial of dW
  float (*w_kc)[kHeadMaxC + 1] = (float (*)[kHeadMaxC + 1])head_smem;
  float (*w_ck)[kHeadMaxK] = (float (*)[kHeadMaxK])(head_smem + kHeadMaxK * (kHeadMaxC + 1));
  float (*dw_sh)[kHeadMaxK] = (float (*)[kHeadMaxK])(head_smem + kHeadMaxK * (kHeadMaxC + 1) + kHeadMaxK * kHeadMaxC);
  __shared__ float db_sh[kHeadMaxC];
  __shared__ float loss_sh;
  for (int i = threadIdx.x; i < kHeadMaxC * kHeadMaxK; i += blockDim.x) {
    const int c = i / kHeadMaxK, k = i % kHeadMaxK;
    const float v = (c < C && k < K) ? W[(size_t)c * K + k] : 0.f;
    w_ck[c][k] = v;
    w_kc[k][c] = v;
    dw_sh[c][k] = 0.f;
  }
  if (threadIdx.x < kHeadMaxC) db_sh[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) loss_sh = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float b0 = (bias && lane < C) ? bias[lane] : 0.f, b1 = (bias && lane + 32 < C) ? bias[lane + 32] : 0.f;
  float dwa[kHeadMaxC][2];
#pragma unroll
  for (int c = 0; c < kHeadMaxC; ++c) dwa[c][0] = dwa[c][1] = 0.f;
  float dba0 = 0.f, dba1 = 0.f, loss_acc = 0.f;
  const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
  for (int64_t r = (int64_t)blockIdx.x * kHeadWarps + w; r < n; r += nwarps) {
    const float* arow = a + r * a_stride;
    const float a0 = lane < K ? arow[lane] : 0.f, a1 = lane + 32 < K ? arow[lane + 32] : 0.f;
    float l0 = b0, l1 = b1;
#pragma unroll 8
    for (int k = 0; k < kHeadMaxK; ++k) {
      const float av = __shfl_sync(pg::kFullMask, k < 32 ? a0 : a1, k & 31);
      l0 += av * w_kc[k][lane];
      l1 += av * w_kc[k][lane + 32];
    }
    const bool v0 = lane < C, v1 = lane + 32 < C;
    float m = fmaxf(v0 ? l0 : -INFINITY, v1 ? l1 : -INFINITY);
#pragma unroll
    for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(pg::kFullMask, m, d));
    const float e0 = v0 ? expf(l0 - m) : 0.f, e1 = v1 ? expf(l1 - m) : 0.f;
    float s = e0 + e1;
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(pg::kFullMask, s, d);
    const int y = (int)labels[r];
    const float ly = __shfl_sync(pg::kFullMask, y < 32 ? l0 : l1, y & 31);
    loss_acc += (m + logf(s)) - ly;
    const float inv_s = 1.0f / s;
    const float g0 = v0 ? (e0 * inv_s - (lane == y ? 1.f : 0.f)) * inv_n : 0.f;
    const float g1 = v1 ? (e1 * inv_s - (lane + 32 == y ? 1.f : 0.f)) * inv_n : 0.f;
    dba0 += g0;
    dba1 += g1;
    float ga0 = 0.f, ga1 = 0.f;
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c) {
      const float gc = __shfl_sync(pg::kFullMask, c < 32 ? g0 : g1, c & 31);
      ga0 += gc * w_ck[c][lane];
      ga1 += gc * w_ck[c][lane + 32];
      dwa[c][0] += gc * a0;
      dwa[c][1] += gc * a1;
    }
    float* grow = grad_a + r * ga_stride;
    if (lane < K) grow[lane] = ga0;
    if (lane + 32 < K) grow[lane + 32] = ga1;
  }
  // CTA reduction: the warps take turns adding their register partials into shared memory (plain read-modify-write,
  // lanes on consecutive words), then one global atomic per output word (16-byte vector atomics when the layout allows)
  for (int turn = 0; turn < kHeadWarps; ++turn) {
    if (w == turn) {
#pragma unroll
      for (int c = 0; c < kHeadMaxC; ++c) {
        dw_sh[c][lane] += dwa[c][0];
        dw_sh[c][lane + 32] += dwa[c][1];
      }
      db_sh[lane] += dba0;
      db_sh[lane + 32] += dba1;
      if (lane == 0) loss_sh += loss_acc;
    }
    __syncthreads();
  }
  if ((K & 3) == 0 && ((uintptr_t)dW & 15) == 0) {
    const int kv = K >> 2;
    for (int i = threadIdx.x; i < C * kv; i += blockDim.x) {
      const int c = i / kv, k = (i % kv) << 2;
      atomicAdd((float4*)&dW[(size_t)c * K + k], *(const float4*)&dw_sh[c][k]);
    }
  } else {
    for (int i = threadIdx.x; i < C * K; i += blockDim.x) {
      const int c = i / K, k = i % K;
      atomicAdd(&dW[i], dw_sh[c][k]);
    }
  }
  if (threadIdx.x < C && db) atomicAdd(&db[threadIdx.x], db_sh[threadIdx.x]);
  if (threadIdx.x == 0) atomicAdd(loss, loss_sh * inv_n);
}

}  // namespace

extern "C" pg_status pg_linear_cross_entropy(const float* d_a, int64_t a_stride, const float* d_weight, const float* d_bias,
                                             const int64_t* d_labels, int64_t n, int32_t in_dim, int32_t n_classes,
                                             float* d_loss, float* d_grad_a, int64_t ga_stride, float* d_grad_weight,
                                             float* d_grad_bias, const int64_t* d_lo, void* stream) {
  PG_REQUIRE(d_weight && d_loss && d_grad_weight && n >= 0 && ((d_a && d_labels && d_grad_a) || n == 0),
             "pg_linear_cross_entropy: bad arguments");
  PG_REQUIRE(in_dim >= 1 && in_dim <= kHeadMaxK && n_classes >= 1 && n_classes <= kHeadMaxC,
             "pg_linear_cross_entropy: in_dim and n_classes must be <= 64");
  PG_REQUIRE(a_stride >= in_dim && ga_stride >= in_dim, "pg_linear_cross_entropy: stride smaller than in_dim");
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  pg::TimedScope timed(PG_T_HEAD, st);
  PG_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(float), st));
  PG_CUDA(cudaMemsetAsync(d_grad_weight, 0, (size_t)n_classes * in_dim * sizeof(float), st));
  if (d_grad_bias) PG_CUDA(cudaMemsetAsync(d_grad_bias, 0, (size_t)n_classes * sizeof(float), st));
  if (n == 0) return PG_OK;
  const char* env_s = getenv("PG_HEAD_SIMT");
  if (!(env_s && atoi(env_s))) {   // tensor-core kernel (pg_dense_mma.cu) whenever the layout allows 16-byte accesses
    const pg_status s = pg::linear_ce_mma(d_a, a_stride, d_weight, d_bias, d_labels, n, in_dim, n_classes, d_loss, d_grad_a,
                                          ga_stride, d_grad_weight, d_grad_bias, d_lo, st);
    if (s != PG_ERR_INVALID) return s;
  }
  // one CTA per SM: the fixed cost per CTA (staging W twice, 4 k atomics for dW) is what the kernel's time is made of
  const int grid = (int)std::min<int64_t>((n + kHeadWarps - 1) / kHeadWarps, (int64_t)pg::sm_count(dev));
  const size_t smem = ((size_t)kHeadMaxK * (kHeadMaxC + 1) + 2 * (size_t)kHeadMaxK * kHeadMaxC) * sizeof(float);
  PG_CUDA(cudaFuncSetAttribute(linear_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PG_CUDA(cudaFuncSetAttribute(linear_ce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  linear_ce_kernel<<<grid, kHeadWarps * 32, smem, st>>>(d_a, a_stride, d_weight, d_bias, d_labels, n, in_dim, n_classes,
                                                     1.0f / (float)n, d_loss, d_grad_a, ga_stride, d_grad_weight, d_grad_bias,
                                                     d_lo);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

// The last NodeFlow block + the classifier head + CrossEntropyLoss, forward and backward, as ONE kernel:
//   a = reduce(block, src);  loss = CE(a W^T + b, labels);  grad_src = reduce^T(d loss / d a);  grad_weight, grad_bias.
// Replaces pg_aggregate_fwd_dyn -> pg_linear_cross_entropy -> pg_aggregate_bwd_dyn (three latency-bound launches and
// the [n, in_dim] round trips of a and grad_a between them). Layouts that rule out 16-byte accesses are rejected (the
// caller then issues the three calls).
extern "C" pg_status pg_block_linear_cross_entropy(const int64_t* d_indptr_base, const int64_t* d_cols,
                                                   const int64_t* d_layer_offsets, const float* d_src, int64_t src_stride,
                                                   int64_t cap_dst, int64_t cap_src, int mode, const float* d_weight,
                                                   const float* d_bias, const int64_t* d_labels, int32_t in_dim,
                                                   int32_t n_classes, float* d_loss, float* d_grad_src, int64_t gsrc_stride,
                                                   float* d_grad_weight, float* d_grad_bias, void* stream) {
  PG_REQUIRE(d_indptr_base && d_cols && d_layer_offsets && d_src && d_weight && d_labels && d_loss && d_grad_src &&
                 d_grad_weight && cap_dst >= 0 && cap_src >= 0,
             "pg_block_linear_cross_entropy: bad arguments");
  PG_REQUIRE(mode == PG_AGG_SUM || mode == PG_AGG_MEAN, "pg_block_linear_cross_entropy: mode must be PG_AGG_SUM or PG_AGG_MEAN");
  PG_REQUIRE(in_dim >= 1 && in_dim <= kHeadMaxK && n_classes >= 1 && n_classes <= kHeadMaxC,
             "pg_block_linear_cross_entropy: in_dim and n_classes must be <= 64");
  PG_REQUIRE(src_stride >= in_dim && gsrc_stride >= in_dim, "pg_block_linear_cross_entropy: stride smaller than in_dim");
  const bool ok = in_dim % 4 == 0 && src_stride % 4 == 0 && gsrc_stride % 4 == 0 &&
                  (((uintptr_t)d_src | (uintptr_t)d_weight | (uintptr_t)d_grad_src | (uintptr_t)d_grad_weight) & 15) == 0;
  PG_REQUIRE(ok, "pg_block_linear_cross_entropy: in_dim and strides must be multiples of 4 with 16-byte aligned buffers");
  cudaStream_t st = (cudaStream_t)stream;
  pg::TimedScope timed(PG_T_HEAD, st);
  PG_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(float), st));
  PG_CUDA(cudaMemsetAsync(d_grad_weight, 0, (size_t)n_classes * in_dim * sizeof(float), st));
  if (d_grad_bias) PG_CUDA(cudaMemsetAsync(d_grad_bias, 0, (size_t)n_classes * sizeof(float), st));
  if (cap_src > 0) {
    if (gsrc_stride == in_dim) {
      PG_CUDA(cudaMemsetAsync(d_grad_src, 0, (size_t)cap_src * in_dim * sizeof(float), st));
    } else {
      PG_CUDA(cudaMemset2DAsync(d_grad_src, (size_t)gsrc_stride * 4, 0, (size_t)in_dim * 4, (size_t)cap_src, st));
    }
  }
  if (cap_dst == 0) return PG_OK;
  return pg::block_linear_ce_mma(d_indptr_base, d_cols, d_layer_offsets, d_src, src_stride, cap_dst, mode, d_weight, d_bias,
                                 d_labels, in_dim, n_classes, d_loss, d_grad_src, gsrc_stride, d_grad_weight, d_grad_bias, st);
}
