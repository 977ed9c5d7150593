// pg_dense_mma.cu — the first NodeUpdate of the GCN / GraphSAGE models on the tensor cores, forward and backward.
//
// Reference: PaGraph/model/gcn_nssc.py:14-24 (NodeUpdate.forward with concat=True) and :64-70 (dropout on the layer's
// output before the next block_compute):
//     z = Linear(x);  out = cat(z, relu(z));  out_drop = dropout(out)
// x [n_1, F] is the aggregated input block (F = 600, n_1 ~ 35 k), Linear is F -> 32. Both products are tall-skinny fp32
// GEMMs (1.36 GFLOP over 85 MB of x): on the fp32 pipe they are FMA-bound (cuBLAS SIMT sgemm 42 us forward; 81 us for
// dW with packed FFMA2), 3-6x over the 13 us it takes to stream x from HBM. Here they run as error-compensated TF32
// ("3xTF32": a = a_hi + a_lo, b = b_hi + b_lo, a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate), which keeps
// fp32-level accuracy (the dropped term is 2^-22 relative).
//
// Shapes are far from the 128 x N tiles tcgen05 wants (N = 32 outputs) and the hi / lo split has to pass through
// registers, so in this round the MMAs are warp-level mma.sync.m16n8k8 with operands loaded straight from global memory
// into fragment registers (ncu: ~8 cycles of tensor pipe per HMMA.1688.F32.TF32 and scheduler, i.e. a ~17 us floor for
// the 2.07 M HMMAs of one product, against a 13.5 us HBM stream; DESIGN.md lists the tcgen05 version as round-2 work):
//   * the k index of a product is a dummy index, so its order is free: a lane fetches 4 consecutive floats (one 16-byte
//     load) and feeds them to two k-steps; both operands use the same permutation. Same trick on the n index of dW
//     (a column permutation of the output, undone when the accumulators are written) and on the class index of the head.
//   * forward: W is split once per CTA into hi / lo TF32 planes in shared memory (2 x 78 KB, conflict-free stride); each
//     warp owns 32 rows of x, keeps 3 load groups of 2 x 16 columns in flight in registers, epilogue fuses bias, relu,
//     the skip-concat and (optionally) the dropout mask of the next block.
//   * backward: dW = gz^T x with gz = g[:, :32] + g[:, 32:] * (z > 0) (g first multiplied by the dropout mask, which is
//     regenerated from the hash, not stored). One persistent CTA per SM walks a contiguous range of rows; gz is built
//     (split hi / lo) in shared memory once per 256 rows by the whole CTA, each warp owns 32 columns of x / dW, partials
//     go out as 16-byte vector atomics (linear_concat_dw2_kernel; linear_concat_dw_kernel is its wider-tile predecessor,
//     still used for in_dim > 608).
//   * head: logits, softmax / loss, grad_a and dW of the classifier as three chained products (linear_ce_mma_kernel).
// Accuracy: the tensor-core accumulators truncate, so the error grows with the number of MMAs per accumulator; measured
// <= 1.5e-5 absolute on O(1) forward outputs and <= 3e-5 of the largest entry of dW at 80 k rows (tests/test_gpu_aggregate.py).
#include <algorithm>
#include <cstdlib>

#include "pg_common.cuh"

namespace {

constexpr int kOut = 32;  // output width (n_hidden of the reference default)

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// D += A (16x8, row) * B (8x8, col), TF32 inputs, fp32 accumulate.
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {  // read-once data: no L1 allocation
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

struct DropArgs {
  uint32_t thr;        // drop when the 16-bit hash lane < thr (0 = no dropout)
  float scale;         // 1 / (1 - p)
  uint64_t seed;
  const int64_t* step; // optional device counter added to the seed
};
// keep-scale factor of column `col` of a row under the drop_hash contract (pg_common.cuh): rowkey = drop_rowkey(stepkey,
// row), colkey = drop_colkey(col / 4), both hoisted by the callers
__device__ __forceinline__ float drop_factor(const DropArgs& d, uint64_t rowkey, uint64_t colkey, int col) {
  const uint64_t h = pg::drop_mix(rowkey, colkey);
  return ((uint32_t)(h >> (16 * (col & 3))) & 0xffffu) < d.thr ? 0.f : d.scale;
}

// ====================================================================================================== forward
// Template knobs (PG_FWD_VARIANT picks among the instantiations below; the default is the measured best):
//   MT     m-tiles (16 rows) per warp: 2 halves the shared-memory reads of W per MMA, 1 halves the registers
//   PAIR   16-column chunks fetched together: PAIR * 64 contiguous bytes per row and request burst (DRAM locality)
//   DEPTH  load groups (PAIR chunks each) in flight per warp, held in registers
//   WARPS  warps per CTA (one CTA per SM: the W planes take 156 KB of shared memory)
__host__ __device__ inline int fwd_wstride(int K) {  // row stride of the W planes: 16 mod 32 words -> conflict-free LDS.128
  const int k16 = (K + 15) / 16 * 16;
  return k16 + ((16 - k16 % 32) + 32) % 32;
}

template <int MT, int PAIR, int DEPTH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
    linear_concat_fwd_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ W,
                             const float* __restrict__ bias, int64_t n, int K, int concat, float* __restrict__ out,
                             int64_t out_stride, float* __restrict__ out_drop, int64_t od_stride, DropArgs drop) {
  constexpr int kThreads = WARPS * 32, kRowsPerWarp = 16 * MT, kRowsPerCta = WARPS * kRowsPerWarp;
  extern __shared__ __align__(16) uint32_t wsm[];  // [2][kOut][ws]: hi plane, lo plane
  const int ws = fwd_wstride(K);
  uint32_t* w_hi = wsm;
  uint32_t* w_lo = wsm + (size_t)kOut * ws;
  {  // W -> hi / lo planes: 16-byte loads, kPre of them in flight per thread before the first use
    const int nvec = K >> 2, total = kOut * nvec;
    constexpr int kPre = 8;
    for (int base = 0; base < total; base += kThreads * kPre) {
      float4 v[kPre];
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int idx = base + u * kThreads + (int)threadIdx.x;
        v[u] = idx < total ? __ldg((const float4*)W + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int idx = base + u * kThreads + (int)threadIdx.x;
        if (idx < total) {
          const int o = idx / nvec, k4 = idx - o * nvec;
          uint4 hi, lo;
          split_tf32(v[u].x, hi.x, lo.x);
          split_tf32(v[u].y, hi.y, lo.y);
          split_tf32(v[u].z, hi.z, lo.z);
          split_tf32(v[u].w, hi.w, lo.w);
          *(uint4*)(w_hi + o * ws + 4 * k4) = hi;
          *(uint4*)(w_lo + o * ws + 4 * k4) = lo;
        }
      }
    }
    const int pad = ws - K;  // columns [K, ws) of every row: zero (the last chunk reads up to the next multiple of 16)
    for (int i = threadIdx.x; i < kOut * pad; i += kThreads) {
      const int o = i / pad, k = K + i % pad;
      w_hi[o * ws + k] = 0u;
      w_lo[o * ws + k] = 0u;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nchunks = (K + 15) / 16;
  const int ngroups = (nchunks + PAIR - 1) / PAIR;
  const int64_t ntiles = (n + kRowsPerCta - 1) / kRowsPerCta;
  const uint64_t stepkey = drop.thr ? pg::drop_stepkey(drop.seed + (drop.step ? (uint64_t)*drop.step : 0ull)) : 0ull;
  // dropout column keys of this lane's output columns: n-tile j covers columns 8 j + 2 t, + 1 (group 2 j + t / 2) of the
  // z half and the same + 32 (group + 8) of the relu half
  uint64_t ck_z[4], ck_p[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ck_z[j] = drop.thr ? pg::drop_colkey((uint32_t)(2 * j + (t >> 1))) : 0ull;
    ck_p[j] = drop.thr ? pg::drop_colkey((uint32_t)(8 + 2 * j + (t >> 1))) : 0ull;
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = (tile * WARPS + warp) * kRowsPerWarp;
    if (r0 >= n) continue;
    // rows of this lane: r0 + g + 8 i, i = 0 .. 2 MT - 1 (m-tile i/2, half i%2)
    const float* xrow[2 * MT];
    bool rok[2 * MT];
#pragma unroll
    for (int i = 0; i < 2 * MT; ++i) {
      const int64_t r = r0 + g + 8 * i;
      rok[i] = r < n;
      xrow[i] = x + (rok[i] ? r : 0) * x_stride + 4 * t;
    }
    float acc[MT][4][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
    float4 buf[DEPTH][PAIR][2 * MT];
    auto load = [&](float4(&b)[PAIR][2 * MT], int group) {
#pragma unroll
      for (int pc = 0; pc < PAIR; ++pc) {
        const int chunk = group * PAIR + pc;
        const bool cok = chunk < nchunks && chunk * 16 + 4 * t < K;
#pragma unroll
        for (int i = 0; i < 2 * MT; ++i)
          b[pc][i] = (cok && rok[i]) ? ld_stream4(xrow[i] + chunk * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load(buf[d], d);
    for (int g0 = 0; g0 < ngroups; g0 += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int group = g0 + d;
        if (group < ngroups) {
          // A fragments of every k-step of this group: [chunk][m-tile][k-step][a0..a3]
          uint32_t ah[PAIR][MT][2][4], al[PAIR][MT][2][4];
#pragma unroll
          for (int pc = 0; pc < PAIR; ++pc)
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                split_tf32(comp(buf[d][pc][2 * m], 2 * s), ah[pc][m][s][0], al[pc][m][s][0]);          // (row g,   k = t)
                split_tf32(comp(buf[d][pc][2 * m + 1], 2 * s), ah[pc][m][s][1], al[pc][m][s][1]);      // (row g+8, k = t)
                split_tf32(comp(buf[d][pc][2 * m], 2 * s + 1), ah[pc][m][s][2], al[pc][m][s][2]);      // (row g,   k = t+4)
                split_tf32(comp(buf[d][pc][2 * m + 1], 2 * s + 1), ah[pc][m][s][3], al[pc][m][s][3]);  // (row g+8, k = t+4)
              }
          load(buf[d], group + DEPTH);
#pragma unroll
          for (int pc = 0; pc < PAIR; ++pc) {
            const int chunk = group * PAIR + pc;
            if (chunk < nchunks) {
              uint4 bh[4], bl[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int off = (j * 8 + g) * ws + chunk * 16 + 4 * t;
                bh[j] = *(const uint4*)(w_hi + off);
                bl[j] = *(const uint4*)(w_lo + off);
              }
#pragma unroll
              for (int term = 0; term < 3; ++term)  // small terms first
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                  for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const uint32_t(&a)[4] = term == 0 ? al[pc][m][s] : ah[pc][m][s];
                      const uint4& b = term == 1 ? bl[j] : bh[j];
                      mma_tf32(acc[m][j], a[0], a[1], a[2], a[3], s == 0 ? b.x : b.z, s == 0 ? b.y : b.w);
                    }
            }
          }
        }
      }
    }
    // epilogue: acc[m][j] = {(row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)} of m-tile m, n-tile j
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t r = r0 + m * 16 + h * 8 + g;
        if (r >= n) continue;
        const uint64_t rk = (out_drop && drop.thr) ? pg::drop_rowkey(stepkey, (uint64_t)r) : 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = j * 8 + 2 * t;
          float z0 = acc[m][j][2 * h], z1 = acc[m][j][2 * h + 1];
          if (bias) {
            z0 += bias[col];
            z1 += bias[col + 1];
          }
          const float p0 = fmaxf(z0, 0.f), p1 = fmaxf(z1, 0.f);
          float* orow = out + r * out_stride;
          if (concat) {
            *(float2*)(orow + col) = make_float2(z0, z1);
            *(float2*)(orow + kOut + col) = make_float2(p0, p1);
          } else {
            *(float2*)(orow + col) = make_float2(p0, p1);
          }
          if (out_drop) {
            float* drow = out_drop + r * od_stride;
            if (concat) {
              *(float2*)(drow + col) = make_float2(z0 * drop_factor(drop, rk, ck_z[j], col), z1 * drop_factor(drop, rk, ck_z[j], col + 1));
              *(float2*)(drow + kOut + col) =
                  make_float2(p0 * drop_factor(drop, rk, ck_p[j], col), p1 * drop_factor(drop, rk, ck_p[j], col + 1));
            } else {
              *(float2*)(drow + col) = make_float2(p0 * drop_factor(drop, rk, ck_z[j], col), p1 * drop_factor(drop, rk, ck_z[j], col + 1));
            }
          }
        }
      }
  }
}

// ====================================================================================================== backward (dW, db)
constexpr int kDwTile = 24;      // rows per gz tile = 3 k-steps of 8 rows, one per x register stage
constexpr int kGzStride = 40;    // floats per gz row in shared memory: conflict-free LDS.128 of the A fragments
constexpr int kDwMaxWarps = 12;  // 64 columns per warp -> in_dim <= 768

template <int kWarps, int kMaxReg>  // CTA size (the warps beyond in_dim / 64 only help staging gz); register cap
__global__ void __launch_bounds__(kWarps * 32) __maxnreg__(kMaxReg)
    linear_concat_dw_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ gout, int64_t g_stride,
                            const float* __restrict__ y, int64_t y_stride, int64_t n, int K, int concat, DropArgs drop,
                            float* dW, float* db) {
  // gz tiles, double-buffered, split into hi / lo planes; element (row, o) lives at [row][(o % 8) * 4 + o / 8] so that
  // lane (g, t) reads its four A values of row t (o = g, g+8, g+16, g+24) with one 16-byte load
  __shared__ __align__(16) uint32_t gz_hi[2][kDwTile][kGzStride];
  __shared__ __align__(16) uint32_t gz_lo[2][kDwTile][kGzStride];
  __shared__ float db_sh[kOut];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (tid < kOut) db_sh[tid] = 0.f;
  int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  rows_per_cta = (rows_per_cta + 7) / 8 * 8;  // whole k-steps; a partial last tile is zero-padded
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(n, r_begin + rows_per_cta);
  if (r_begin >= r_end) return;
  const int ntiles = (int)((r_end - r_begin + kDwTile - 1) / kDwTile);
  const uint64_t stepkey = drop.thr ? pg::drop_stepkey(drop.seed + (drop.step ? (uint64_t)*drop.step : 0ull)) : 0ull;
  // a thread stages gz[.][o] for one fixed o = tid % 32: dropout column keys of columns o and 32 + o
  const uint64_t ck_a = drop.thr ? pg::drop_colkey((uint32_t)((tid & 31) >> 2)) : 0ull;
  const uint64_t ck_b = drop.thr ? pg::drop_colkey((uint32_t)(8 + ((tid & 31) >> 2))) : 0ull;
  // this warp's columns: blocks cb = 2 warp, 2 warp + 1 of 32 columns; lane fetches floats [32 cb + 4 g, +4)
  const int col0 = warp * 64 + 4 * g;
  const bool cok[2] = {col0 < K, col0 + 32 < K};
  const bool cb_any[2] = {warp * 64 < K, warp * 64 + 32 < K};

  // ---- gz of one tile into registers (global loads), then into shared memory
  constexpr int kThreads = kWarps * 32;
  constexpr int per = (kDwTile * kOut + kThreads - 1) / kThreads;  // gz values a thread stages per tile
  // raw operands of gz, loaded one tile ahead (no arithmetic on them until store_gz: the loads stay in flight under
  // the MMAs of the current tile)
  float gr_a[per], gr_b[per], gr_y[per];
  float dbv = 0.f;
  auto load_gz = [&](int tile) {
    const int64_t r0 = r_begin + (int64_t)tile * kDwTile;
#pragma unroll
    for (int q = 0; q < per; ++q) {
      const int i = tid + q * kThreads;
      const int rr = i / kOut, o = i % kOut;
      const int64_t r = r0 + rr;
      gr_a[q] = gr_b[q] = gr_y[q] = 0.f;
      if (i < kDwTile * kOut && r < r_end) {
        const float* grow = gout + r * g_stride;
        const float* yrow = y + r * y_stride;
        if (concat) {
          gr_a[q] = __ldg(grow + o);
          gr_b[q] = __ldg(grow + kOut + o);
          gr_y[q] = __ldg(yrow + kOut + o);
        } else {
          gr_a[q] = __ldg(grow + o);
          gr_y[q] = __ldg(yrow + o);
        }
      }
    }
  };
  auto store_gz = [&](int tile, int buf) {
    const int64_t r0 = r_begin + (int64_t)tile * kDwTile;
#pragma unroll
    for (int q = 0; q < per; ++q) {
      const int i = tid + q * kThreads;
      if (i < kDwTile * kOut) {
        const int rr = i / kOut, o = i % kOut;
        const int64_t r = r0 + rr;
        float v;
        const uint64_t rk = drop.thr ? pg::drop_rowkey(stepkey, (uint64_t)r) : 0ull;
        if (concat) {
          float ga = gr_a[q], gb = gr_y[q] > 0.f ? gr_b[q] : 0.f;
          if (drop.thr) {
            ga *= drop_factor(drop, rk, ck_a, o);
            gb *= drop_factor(drop, rk, ck_b, o);   // column 32 + o: same lane (o % 4) of group 8 + o / 4
          }
          v = ga + gb;
        } else {
          v = gr_y[q] > 0.f ? gr_a[q] : 0.f;
          if (drop.thr) v *= drop_factor(drop, rk, ck_a, o);
        }
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        gz_hi[buf][rr][(o & 7) * 4 + (o >> 3)] = hi;
        gz_lo[buf][rr][(o & 7) * 4 + (o >> 3)] = lo;
        dbv += v;  // o = tid % 32 for every q (the CTA size is a multiple of 32)
      }
    }
  };

  // ---- x stages: 8 rows (one k-step) x this warp's 64 columns; lane holds rows {t, t+4} x 2 blocks; 3 stages in flight
  float4 xb[3][2][2];  // [stage buffer][block][row slot]
  const int nstages = 3 * ntiles;
  auto load_x = [&](float4(&b)[2][2], int stage) {
    const int64_t rs = r_begin + (int64_t)stage * 8;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t r = rs + t + 4 * i;
        b[c][i] = (stage < nstages && cok[c] && r < r_end) ? ld_stream4(x + r * x_stride + col0 + 32 * c)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  };
  float acc[2][2][4][4];  // [m-tile][block][n-tile j][c0..c3]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[m][c][j][q] = 0.f;

  auto compute = [&](const float4(&b)[2][2], int buf, int ks) {  // k-step ks: rows [8 ks, 8 ks + 8) of the gz tile
    const int rr = ks * 8;
    // A: (m-tile 0: a0 a1 | m-tile 1: a0 a1) from row rr + t, (a2 a3 | a2 a3) from row rr + t + 4
    const uint4 h0 = *(const uint4*)&gz_hi[buf][rr + t][g * 4], h1 = *(const uint4*)&gz_hi[buf][rr + t + 4][g * 4];
    const uint4 l0 = *(const uint4*)&gz_lo[buf][rr + t][g * 4], l1 = *(const uint4*)&gz_lo[buf][rr + t + 4][g * 4];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (!cb_any[c]) continue;  // warp-uniform
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_tf32(comp(b[c][0], j), bh[j][0], bl[j][0]);  // (k = t,   n = g) of n-tile j
        split_tf32(comp(b[c][1], j), bh[j][1], bl[j][1]);  // (k = t+4, n = g)
      }
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4& a0 = term == 0 ? l0 : h0;
          const uint4& a1 = term == 0 ? l1 : h1;
          const uint32_t b0 = term == 1 ? bl[j][0] : bh[j][0], b1 = term == 1 ? bl[j][1] : bh[j][1];
          mma_tf32(acc[0][c][j], a0.x, a0.y, a1.x, a1.y, b0, b1);
          mma_tf32(acc[1][c][j], a0.z, a0.w, a1.z, a1.w, b0, b1);
        }
    }
  };

#pragma unroll
  for (int d = 0; d < 3; ++d) load_x(xb[d], d);
  load_gz(0);
  store_gz(0, 0);
  __syncthreads();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_gz(tile + 1);  // in flight during the MMAs below
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      compute(xb[d], buf, d);
      load_x(xb[d], 3 * tile + 3 + d);
    }
    if (tile + 1 < ntiles) store_gz(tile + 1, buf ^ 1);
    __syncthreads();
  }
  // ---- write-out: acc[m][c][j] = {(o = 16m+g, nu = 2t), (o, 2t+1), (o+8, 2t), (o+8, 2t+1)}, column = 64 warp + 32 c + 4 nu + j
  // the four n-tiles j of one (m, c, q) are four consecutive columns: one 16-byte vector atomic
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = 16 * m + g + 8 * (q >> 1);
        const int col = warp * 64 + 32 * c + 4 * (2 * t + (q & 1));
        if (col < K)
          atomicAdd((float4*)&dW[(size_t)o * K + col],
                    make_float4(acc[m][c][0][q], acc[m][c][1][q], acc[m][c][2][q], acc[m][c][3][q]));
      }
  if (db) {
    atomicAdd(&db_sh[lane], dbv);
    __syncthreads();
    if (tid < kOut) atomicAdd(&db[tid], db_sh[tid]);
  }
}

// ------------------------------------------------------------------------------------------------------ dW, one warp per column block
// Same product as linear_concat_dw_kernel, restructured after its ncu profile (tensor pipe 28 % busy: 2.5 warps per
// scheduler do not hide the MMA issue latency, the 10 column warps split 3/3/2/2 over the 4 schedulers, every warp
// stalls on the gz operand loads and then on the per-tile barrier):
//   * one warp per 32-column block of x / dW: 19 warps at in_dim = 600 (5/5/5/4 per scheduler), 96 registers;
//   * gz for a whole super-tile of 256 rows (the CTA's share at config 2) is built ONCE, by all warps together, into
//     shared memory (hi / lo planes, 80 KB) while the first x stages are already in flight; after that barrier the warps
//     run their k-steps without synchronising.
constexpr int kDw2Super = 256;        // rows per gz super-tile (32 k-steps)
constexpr int kDw2MaxWarps = 19;      // in_dim <= 608
constexpr int kDw2Batch = 5;          // gz elements a thread loads together in the prologue

__global__ void __maxnreg__(96)       // 19 warps x 32 x 96 registers = 57 k of the SM's 64 k, <= 5 warps per scheduler
    linear_concat_dw2_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ gout, int64_t g_stride,
                             const float* __restrict__ y, int64_t y_stride, int64_t n, int K, int concat, DropArgs drop,
                             float* dW, float* db) {
  extern __shared__ __align__(16) uint32_t gzs[];     // [2 planes][kDw2Super][kGzStride]
  uint32_t* gz_hi = gzs;
  uint32_t* gz_lo = gzs + kDw2Super * kGzStride;
  __shared__ float db_part[kDw2MaxWarps][kOut];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  rows_per_cta = (rows_per_cta + 7) / 8 * 8;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(n, r_begin + rows_per_cta);
  if (r_begin >= r_end) return;
  const int nsuper = (int)((r_end - r_begin + kDw2Super - 1) / kDw2Super);
  const int nstages = (int)((r_end - r_begin + 7) / 8);   // k-steps of the whole CTA
  // ---- gz prologue state: this thread stages gz[rows warp + nwarps q][o = lane]
  const int o = lane;
  const uint64_t stepkey = drop.thr ? pg::drop_stepkey(drop.seed + (drop.step ? (uint64_t)*drop.step : 0ull)) : 0ull;
  const uint64_t ck_a = drop.thr ? pg::drop_colkey((uint32_t)(o >> 2)) : 0ull;
  const uint64_t ck_b = drop.thr ? pg::drop_colkey((uint32_t)(8 + (o >> 2))) : 0ull;
  float dbv = 0.f;
  auto build_gz = [&](int64_t s0) {                    // rows [s0, s0 + kDw2Super) of the CTA's range
#pragma unroll 1
    for (int rr0 = warp; rr0 < kDw2Super; rr0 += nwarps * kDw2Batch) {
      float ra[kDw2Batch], rb[kDw2Batch], ry[kDw2Batch];
#pragma unroll
      for (int q = 0; q < kDw2Batch; ++q) {            // the batch's loads are in flight before the first use
        const int rr = rr0 + q * nwarps;
        const int64_t r = s0 + rr;
        ra[q] = rb[q] = ry[q] = 0.f;
        if (rr < kDw2Super && r < r_end) {
          const float* grow = gout + r * g_stride;
          const float* yrow = y + r * y_stride;
          ra[q] = __ldg(grow + o);
          if (concat) {
            rb[q] = __ldg(grow + kOut + o);
            ry[q] = __ldg(yrow + kOut + o);
          } else {
            ry[q] = __ldg(yrow + o);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kDw2Batch; ++q) {
        const int rr = rr0 + q * nwarps;
        if (rr >= kDw2Super) continue;
        const int64_t r = s0 + rr;
        const uint64_t rk = drop.thr ? pg::drop_rowkey(stepkey, (uint64_t)r) : 0ull;
        float v;
        if (concat) {
          float ga = ra[q], gb = ry[q] > 0.f ? rb[q] : 0.f;
          if (drop.thr) {
            ga *= drop_factor(drop, rk, ck_a, o);
            gb *= drop_factor(drop, rk, ck_b, o);
          }
          v = ga + gb;
        } else {
          v = ry[q] > 0.f ? ra[q] : 0.f;
          if (drop.thr) v *= drop_factor(drop, rk, ck_a, o);
        }
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        gz_hi[rr * kGzStride + (o & 7) * 4 + (o >> 3)] = hi;
        gz_lo[rr * kGzStride + (o & 7) * 4 + (o >> 3)] = lo;
        dbv += v;
      }
    }
  };
  // ---- x: columns [32 warp, 32 warp + 32); 3 k-steps (8 rows each) in flight in registers
  const int col0 = warp * 32 + 4 * g;
  const bool cok = col0 < K;
  float4 xb[3][2];  // [ring slot][row slot]: rows t, t + 4 of the k-step
  auto load_x = [&](float4(&b)[2], int stage) {
    const int64_t rs = r_begin + (int64_t)stage * 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = rs + t + 4 * i;
      b[i] = (stage < nstages && cok && r < r_end) ? ld_stream4(x + r * x_stride + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float acc[2][4][4];  // [m-tile][n-tile j][c0..c3]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
  // k-step at rows [rr, rr + 8) of the super-tile from ring slot b; the slot is refilled with k-step `next` as soon as
  // its values have been split into the B fragments, i.e. before the MMAs: three full k-steps of lead time for the load
  auto compute = [&](float4(&b)[2], int rr, int next) {
    uint32_t bh[4][2], bl[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split_tf32(comp(b[0], j), bh[j][0], bl[j][0]);
      split_tf32(comp(b[1], j), bh[j][1], bl[j][1]);
    }
    load_x(b, next);
    const uint4 h0 = *(const uint4*)(gz_hi + (rr + t) * kGzStride + g * 4), h1 = *(const uint4*)(gz_hi + (rr + t + 4) * kGzStride + g * 4);
    const uint4 l0 = *(const uint4*)(gz_lo + (rr + t) * kGzStride + g * 4), l1 = *(const uint4*)(gz_lo + (rr + t + 4) * kGzStride + g * 4);
#pragma unroll
    for (int term = 0; term < 3; ++term)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4& a0 = term == 0 ? l0 : h0;
        const uint4& a1 = term == 0 ? l1 : h1;
        const uint32_t b0 = term == 1 ? bl[j][0] : bh[j][0], b1 = term == 1 ? bl[j][1] : bh[j][1];
        mma_tf32(acc[0][j], a0.x, a0.y, a1.x, a1.y, b0, b1);
        mma_tf32(acc[1][j], a0.z, a0.w, a1.z, a1.w, b0, b1);
      }
  };
#pragma unroll
  for (int d = 0; d < 3; ++d) load_x(xb[d], d);
  int stage = 0;                                          // next k-step to compute (its x is in ring slot stage % 3)
  for (int sp = 0; sp < nsuper; ++sp) {
    if (sp) __syncthreads();                              // every warp is done reading the previous super-tile's gz
    build_gz(r_begin + (int64_t)sp * kDw2Super);
    __syncthreads();
    const int ks_end = min(nstages, (sp + 1) * (kDw2Super / 8));
    // the ring has 3 slots and a super-tile 32 k-steps: unroll by 3 with the slot = position in the unrolled body; the
    // global k-step index stays aligned with the ring because every super-tile but the last has 32 = 3 * 10 + 2 k-steps,
    // handled by rotating the starting slot
    while (stage < ks_end) {
      const int slot = stage % 3;
      const int rr = (stage - sp * (kDw2Super / 8)) * 8;
      if (slot == 0) compute(xb[0], rr, stage + 3);
      else if (slot == 1) compute(xb[1], rr, stage + 3);
      else compute(xb[2], rr, stage + 3);
      ++stage;
    }
  }
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int oc = 16 * m + g + 8 * (q >> 1);
      const int col = warp * 32 + 4 * (2 * t + (q & 1));
      if (col < K)
        atomicAdd((float4*)&dW[(size_t)oc * K + col], make_float4(acc[m][0][q], acc[m][1][q], acc[m][2][q], acc[m][3][q]));
    }
  if (db) {
    db_part[warp][lane] = dbv;
    __syncthreads();
    if (warp == 0) {
      float sum = 0.f;
      for (int w2 = 0; w2 < nwarps; ++w2) sum += db_part[w2][lane];
      atomicAdd(&db[lane], sum);
    }
  }
}

// ====================================================================================================== classifier head + loss
// pred = a W^T + b, loss = mean_r(logsumexp(pred_r) - pred_r[label_r]) and all three gradients (reference: the last
// NodeUpdate, gcn_nssc.py:48, + CrossEntropyLoss, pa_gcn.py:62,93-94) as three chained tensor-core products per CTA of
// 64 rows (4 warps x 16 rows), 3xTF32 like the kernels above:
//   1. logits = a W^T            per warp; softmax / loss / G = d loss / d logits on the accumulator fragments
//   2. grad_a = G W              per warp; the accumulator fragment of (1) IS the A fragment of (2) under the free
//                                permutation of the summed index (class 8j+2t <-> k = t, class 8j+2t+1 <-> k = t+4)
//   3. dW    += G^T a            per CTA, G and a staged in shared memory, warp w owns classes [16 w, 16 w + 16);
//                                one 16-byte vector atomic per 4 outputs, db / loss by warp-reduced scalar atomics
// in_dim, n_classes <= 64 (zero-padded to 64), in_dim % 4 == 0.
constexpr int kHeadWarpsM = 4;
constexpr int kHeadWS = 84;   // row stride (words) of the W planes: conflict-free for both operand access patterns
constexpr int kHeadGS = 66;   // ... of the G planes (8-byte loads, 4 rows x 8 class groups per half-warp)
constexpr int kHeadAS = 72;   // ... of the staged a tile (16-byte loads, 4 rows x 2 column groups per quarter-warp)

//
// BLOCK = true fuses the NodeFlow block in front of the head and its backward into the same kernel (the last
// block_compute of gcn_nssc.py:71-74 feeding the last NodeUpdate): row r of `a` is then not loaded but REDUCED from
// the block's source rows (a = src rows, copy_src + sum / mean over indptr / cols, sequential edge order like
// agg_fwd_vec4), and grad_a is not stored but SCATTERED back over the same edges into grad_src (pre-zeroed; vector
// reductions like agg_bwd_kernel). lo3 = device-resident NodeFlow offsets {source layer, destination layer, end}.
struct HeadBlock {
  const int64_t* indptr;   // NodeFlow-wide indptr base
  const int64_t* cols;
  const int64_t* lo3;
  int mean;
  float* grad_src;
  int64_t gsrc_stride;
};

template <bool BLOCK>
__global__ void __launch_bounds__(kHeadWarpsM * 32)
    linear_ce_mma_kernel(const float* __restrict__ a, int64_t a_stride, const float* __restrict__ W,
                         const float* __restrict__ bias, const int64_t* __restrict__ labels, int64_t n, int K, int C,
                         float inv_n, float* loss, float* grad_a, int64_t ga_stride, float* dW, float* db,
                         const int64_t* __restrict__ lo, HeadBlock blk) {
  extern __shared__ __align__(16) uint32_t hsm[];
  int64_t col_base = 0;
  if (BLOCK) {
    const int64_t l0 = blk.lo3[0], l1 = blk.lo3[1], l2 = blk.lo3[2];
    blk.indptr += l1;
    col_base = l0;
    n = min(n, l2 - l1);
    inv_n = 1.0f / (float)max(n, (int64_t)1);
  } else if (lo) {  // device-resident row count: n is a capacity
    n = min(n, lo[1] - lo[0]);
    inv_n = 1.0f / (float)max(n, (int64_t)1);
  }
  uint32_t* w_hi = hsm;                                   // [64][kHeadWS]  w[class][k]
  uint32_t* w_lo = w_hi + 64 * kHeadWS;
  uint32_t* g_hi = w_lo + 64 * kHeadWS;                   // [64 rows][kHeadGS], class c at word (c % 8) * 8 + c / 8
  uint32_t* g_lo = g_hi + 64 * kHeadGS;
  float* a_s = (float*)(g_lo + 64 * kHeadGS);             // [64 rows][kHeadAS]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // ---- phase 0: W -> hi / lo planes (zero-padded to 64 x 64)
  for (int idx = tid; idx < 64 * 16; idx += kHeadWarpsM * 32) {
    const int c = idx >> 4, k4 = (idx & 15) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C && k4 < K) v = __ldg((const float4*)(W + (size_t)c * K + k4));
    uint4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    *(uint4*)(w_hi + c * kHeadWS + k4) = hi;
    *(uint4*)(w_lo + c * kHeadWS + k4) = lo;
  }
  // ---- this warp's 16 rows of a: fragments (rows g, g+8; columns 16 c + 4 t ..) and a copy in shared memory
  const int64_t r0 = ((int64_t)blockIdx.x * kHeadWarpsM + w) * 16;
  const int64_t ra = r0 + g, rb = r0 + g + 8;
  const bool va = ra < n, vb = rb < n;
  float4 xa[4], xb[4];
  int64_t ea0 = 0, ea1 = 0, eb0 = 0, eb1 = 0;   // BLOCK: edge ranges of rows ra, rb
  // BLOCK: the quad (lanes 4 g .. 4 g + 3) keeps the first 12 source indices of each of its two rows in registers —
  // lane t holds edges t, t + 4, t + 8 — for the reduction here and the scatter of the backward; longer rows (not
  // produced by the reference's fanouts) take the remaining edges from memory.
  constexpr int kHeadEdgeRegs = 3;
  int ca[kHeadEdgeRegs], cb[kHeadEdgeRegs];
  if (BLOCK) {
    if (va) { ea0 = blk.indptr[ra]; ea1 = blk.indptr[ra + 1]; }
    if (vb) { eb0 = blk.indptr[rb]; eb1 = blk.indptr[rb + 1]; }
#pragma unroll
    for (int i = 0; i < kHeadEdgeRegs; ++i) {
      const int64_t ja = ea0 + 4 * i + t, jb = eb0 + 4 * i + t;
      ca[i] = ja < ea1 ? (int)(blk.cols[ja] - col_base) : 0;
      cb[i] = jb < eb1 ? (int)(blk.cols[jb] - col_base) : 0;
    }
    // lane t holds columns 16 c + 4 t (c = 0..3) of the quad's rows; edges are summed in order (like agg_fwd_vec4)
    auto reduce_row = [&](int64_t e0, int64_t e1, const int(&cr)[kHeadEdgeRegs], float4(&x)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int deg = (int)min(e1 - e0, (int64_t)(4 * kHeadEdgeRegs));
#pragma unroll
      for (int i = 0; i < kHeadEdgeRegs; ++i) {
        float4 v[4][4];
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {   // four source rows in flight
          const int src_row = __shfl_sync(pg::kFullMask, cr[i], (lane & ~3) | tt);
          const float* p = a + (int64_t)src_row * a_stride;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col = 16 * c + 4 * t;
            v[tt][c] = (4 * i + tt < deg && col < K) ? __ldg((const float4*)(p + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int tt = 0; tt < 4; ++tt)
          if (4 * i + tt < deg) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { x[c].x += v[tt][c].x; x[c].y += v[tt][c].y; x[c].z += v[tt][c].z; x[c].w += v[tt][c].w; }
          }
      }
      for (int64_t j = e0 + 4 * kHeadEdgeRegs; j < e1; ++j) {
        const float* p = a + (blk.cols[j] - col_base) * a_stride;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = 16 * c + 4 * t;
          if (col < K) {
            const float4 v = __ldg((const float4*)(p + col));
            x[c].x += v.x; x[c].y += v.y; x[c].z += v.z; x[c].w += v.w;
          }
        }
      }
      if (blk.mean) {
        const float dg = (float)max(e1 - e0, (int64_t)1);
#pragma unroll
        for (int c = 0; c < 4; ++c) { x[c].x /= dg; x[c].y /= dg; x[c].z /= dg; x[c].w /= dg; }
      }
    };
    reduce_row(ea0, ea1, ca, xa);
    reduce_row(eb0, eb1, cb, xb);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int col = 16 * c + 4 * t;
    if (!BLOCK) {
      xa[c] = (va && col < K) ? __ldg((const float4*)(a + ra * a_stride + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      xb[c] = (vb && col < K) ? __ldg((const float4*)(a + rb * a_stride + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *(float4*)(a_s + (w * 16 + g) * kHeadAS + col) = xa[c];
    *(float4*)(a_s + (w * 16 + g + 8) * kHeadAS + col) = xb[c];
  }
  const int64_t ya = va ? labels[ra] : -1, yb = vb ? labels[rb] : -1;
  __syncthreads();
  // ---- phase 1: logits (acc[j] = n-tile j = classes 8 j .. 8 j + 7; lane holds classes 8 j + 2 t, + 1 of rows g, g+8)
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      split_tf32(comp(xa[c], 2 * s), ah[s][0], al[s][0]);
      split_tf32(comp(xb[c], 2 * s), ah[s][1], al[s][1]);
      split_tf32(comp(xa[c], 2 * s + 1), ah[s][2], al[s][2]);
      split_tf32(comp(xb[c], 2 * s + 1), ah[s][3], al[s][3]);
    }
    uint4 bh[8], bl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int off = (8 * j + g) * kHeadWS + 16 * c + 4 * t;
      bh[j] = *(const uint4*)(w_hi + off);
      bl[j] = *(const uint4*)(w_lo + off);
    }
#pragma unroll
    for (int term = 0; term < 3; ++term)
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t(&av)[4] = term == 0 ? al[s] : ah[s];
          const uint4& b = term == 1 ? bl[j] : bh[j];
          mma_tf32(acc[j], av[0], av[1], av[2], av[3], s == 0 ? b.x : b.z, s == 0 ? b.y : b.w);
        }
  }
  // ---- softmax, loss, G (overwrites acc): rows g (q = 0, 1) and g + 8 (q = 2, 3)
  float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = 8 * j + 2 * t + (q & 1);
      float z = acc[j][q] + ((bias && col < C) ? bias[col] : 0.f);
      if (col >= C) z = -INFINITY;
      acc[j][q] = z;
      if (q < 2) ma = fmaxf(ma, z);
      else mb = fmaxf(mb, z);
    }
  ma = fmaxf(ma, __shfl_xor_sync(pg::kFullMask, ma, 1));
  ma = fmaxf(ma, __shfl_xor_sync(pg::kFullMask, ma, 2));
  mb = fmaxf(mb, __shfl_xor_sync(pg::kFullMask, mb, 1));
  mb = fmaxf(mb, __shfl_xor_sync(pg::kFullMask, mb, 2));
  float sa = 0.f, sb = 0.f, la = 0.f, lb = 0.f;   // sum of exp, logit at the label
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = 8 * j + 2 * t + (q & 1);
      const float z = acc[j][q];
      const float e = col < C ? expf(z - (q < 2 ? ma : mb)) : 0.f;
      if (q < 2) {
        sa += e;
        if (col == ya) la = z;
      } else {
        sb += e;
        if (col == yb) lb = z;
      }
      acc[j][q] = e;
    }
#pragma unroll
  for (int d = 1; d <= 2; d <<= 1) {
    sa += __shfl_xor_sync(pg::kFullMask, sa, d);
    sb += __shfl_xor_sync(pg::kFullMask, sb, d);
    la += __shfl_xor_sync(pg::kFullMask, la, d);
    lb += __shfl_xor_sync(pg::kFullMask, lb, d);
  }
  float lsum = 0.f;
  if (t == 0) lsum = (va ? (ma + logf(sa)) - la : 0.f) + (vb ? (mb + logf(sb)) - lb : 0.f);
#pragma unroll
  for (int d = 16; d; d >>= 1) lsum += __shfl_xor_sync(pg::kFullMask, lsum, d);
  if (lane == 0) atomicAdd(loss, lsum * inv_n);
  const float isa = va ? inv_n / sa : 0.f, isb = vb ? inv_n / sb : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = 8 * j + 2 * t + (q & 1);
      const bool up = q < 2;
      float gv = acc[j][q] * (up ? isa : isb);
      if (col == (up ? ya : yb)) gv -= inv_n;          // ya / yb = -1 for rows beyond n: never matches
      if (col >= C || !(up ? va : vb)) gv = 0.f;
      acc[j][q] = gv;
    }
  // db: column sums over this warp's rows -> lanes g == 0 -> global
  if (db) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d0 = acc[j][0] + acc[j][2], d1 = acc[j][1] + acc[j][3];
#pragma unroll
      for (int d = 4; d <= 16; d <<= 1) {
        d0 += __shfl_xor_sync(pg::kFullMask, d0, d);
        d1 += __shfl_xor_sync(pg::kFullMask, d1, d);
      }
      if (g == 0) {
        const int col = 8 * j + 2 * t;
        if (col < C) atomicAdd(&db[col], d0);
        if (col + 1 < C) atomicAdd(&db[col + 1], d1);
      }
    }
  }
  // ---- phase 2: grad_a = G W; k-step j sums classes 8 j .. 8 j + 7 with A = the G fragment as it stands; the output
  // column of n-tile 4 h + i, n = nu, is 32 h + 4 nu + i (a lane's four n-tiles of a half are 4 consecutive columns)
  float acc2[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc2[j][q] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t ah[4], al[4];
    split_tf32(acc[j][0], ah[0], al[0]);   // (row g,   class 8j+2t)   = A[g][k = t]
    split_tf32(acc[j][2], ah[1], al[1]);   // (row g+8, class 8j+2t)   = A[g+8][t]
    split_tf32(acc[j][1], ah[2], al[2]);   // (row g,   class 8j+2t+1) = A[g][t+4]
    split_tf32(acc[j][3], ah[3], al[3]);   // (row g+8, class 8j+2t+1) = A[g+8][t+4]
    // stage G for phase 3
    {
      const int ra_s = (w * 16 + g) * kHeadGS, rb_s = (w * 16 + g + 8) * kHeadGS;
      const int p0 = (2 * t) * 8 + j, p1 = (2 * t + 1) * 8 + j;
      g_hi[ra_s + p0] = ah[0]; g_lo[ra_s + p0] = al[0];
      g_hi[rb_s + p0] = ah[1]; g_lo[rb_s + p0] = al[1];
      g_hi[ra_s + p1] = ah[2]; g_lo[ra_s + p1] = al[2];
      g_hi[rb_s + p1] = ah[3]; g_lo[rb_s + p1] = al[3];
    }
    uint4 bh0[2], bh1[2], bl0[2], bl1[2];   // [half]: W[class 8j+2t][32 h + 4 g ..], W[class 8j+2t+1][..]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o0 = (8 * j + 2 * t) * kHeadWS + 32 * h + 4 * g, o1 = o0 + kHeadWS;
      bh0[h] = *(const uint4*)(w_hi + o0);
      bh1[h] = *(const uint4*)(w_hi + o1);
      bl0[h] = *(const uint4*)(w_lo + o0);
      bl1[h] = *(const uint4*)(w_lo + o1);
    }
#pragma unroll
    for (int term = 0; term < 3; ++term)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t(&av)[4] = term == 0 ? al : ah;
          const uint4& b0 = term == 1 ? bl0[h] : bh0[h];
          const uint4& b1 = term == 1 ? bl1[h] : bh1[h];
          const uint32_t x0 = i == 0 ? b0.x : i == 1 ? b0.y : i == 2 ? b0.z : b0.w;
          const uint32_t x1 = i == 0 ? b1.x : i == 1 ? b1.y : i == 2 ? b1.z : b1.w;
          mma_tf32(acc2[4 * h + i], av[0], av[1], av[2], av[3], x0, x1);
        }
  }
  if (BLOCK) {
    // backward of the block: row r's gradient (/ degree for mean) added to every source row of r
    const float sa_ = (blk.mean && ea1 > ea0) ? (float)(ea1 - ea0) : 1.0f;   // divided, like agg_bwd_kernel
    const float sb_ = (blk.mean && eb1 > eb0) ? (float)(eb1 - eb0) : 1.0f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t e0 = half ? eb0 : ea0, e1 = half ? eb1 : ea1;
      const float sc = half ? sb_ : sa_;
      float4 gv[2][2];   // [h][qq]: this lane's 4 x 4 columns of the row's gradient
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = 2 * half + qq;
          gv[h][qq] = make_float4(acc2[4 * h][q] / sc, acc2[4 * h + 1][q] / sc, acc2[4 * h + 2][q] / sc, acc2[4 * h + 3][q] / sc);
        }
      const int deg = (int)min(e1 - e0, (int64_t)(4 * kHeadEdgeRegs));
#pragma unroll
      for (int i = 0; i < kHeadEdgeRegs; ++i)
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
          const int src_row = __shfl_sync(pg::kFullMask, half ? cb[i] : ca[i], (lane & ~3) | tt);
          if (4 * i + tt < deg) {
            float* dst = blk.grad_src + (int64_t)src_row * blk.gsrc_stride;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int qq = 0; qq < 2; ++qq) {
                const int col = 32 * h + 8 * t + 4 * qq;
                if (col < K) atomicAdd((float4*)(dst + col), gv[h][qq]);
              }
          }
        }
      for (int64_t j = e0 + 4 * kHeadEdgeRegs; j < e1; ++j) {
        float* dst = blk.grad_src + (blk.cols[j] - col_base) * blk.gsrc_stride;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {
            const int col = 32 * h + 8 * t + 4 * qq;
            if (col < K) atomicAdd((float4*)(dst + col), gv[h][qq]);
          }
      }
    }
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t r = (q < 2) ? ra : rb;
        const int col = 32 * h + 8 * t + 4 * (q & 1);
        if (r < n && col < K)
          *(float4*)(grad_a + r * ga_stride + col) =
              make_float4(acc2[4 * h][q], acc2[4 * h + 1][q], acc2[4 * h + 2][q], acc2[4 * h + 3][q]);
      }
  }
  __syncthreads();
  // ---- phase 3: dW[16 w + .][.] += G^T a over the CTA's 64 rows (8 k-steps of 8 rows)
  float acc3[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc3[j][q] = 0.f;
#pragma unroll 2
  for (int ks = 0; ks < 8; ++ks) {
    const int rt = 8 * ks + t;
    const uint2 h0 = *(const uint2*)(g_hi + rt * kHeadGS + g * 8 + 2 * w), h1 = *(const uint2*)(g_hi + (rt + 4) * kHeadGS + g * 8 + 2 * w);
    const uint2 l0 = *(const uint2*)(g_lo + rt * kHeadGS + g * 8 + 2 * w), l1 = *(const uint2*)(g_lo + (rt + 4) * kHeadGS + g * 8 + 2 * w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 v0 = *(const float4*)(a_s + rt * kHeadAS + 32 * h + 4 * g);
      const float4 v1 = *(const float4*)(a_s + (rt + 4) * kHeadAS + 32 * h + 4 * g);
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        split_tf32(comp(v0, i), bh[i][0], bl[i][0]);
        split_tf32(comp(v1, i), bh[i][1], bl[i][1]);
      }
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint2& a0 = term == 0 ? l0 : h0;
          const uint2& a1 = term == 0 ? l1 : h1;
          mma_tf32(acc3[4 * h + i], a0.x, a0.y, a1.x, a1.y, term == 1 ? bl[i][0] : bh[i][0], term == 1 ? bl[i][1] : bh[i][1]);
        }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = 16 * w + g + 8 * (q >> 1);
      const int col = 32 * h + 8 * t + 4 * (q & 1);
      if (c < C && col < K)
        atomicAdd((float4*)&dW[(size_t)c * K + col],
                  make_float4(acc3[4 * h][q], acc3[4 * h + 1][q], acc3[4 * h + 2][q], acc3[4 * h + 3][q]));
    }
}

DropArgs make_drop(float p, uint64_t seed, const int64_t* d_step) {
  DropArgs d;
  d.thr = p > 0.f ? (uint32_t)(p * 65536.0f + 0.5f) : 0u;
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  d.seed = seed;
  d.step = d_step;
  return d;
}

}  // namespace

namespace pg {

// Tensor-core head kernel; outputs (loss, grad_weight, grad_bias) must be zeroed by the caller. Returns PG_ERR_INVALID
// without launching when the layout does not allow 16-byte accesses (the caller then uses the scalar kernel).
pg_status linear_ce_mma(const float* d_a, int64_t a_stride, const float* d_weight, const float* d_bias, const int64_t* d_labels,
                        int64_t n, int32_t in_dim, int32_t n_classes, float* d_loss, float* d_grad_a, int64_t ga_stride,
                        float* d_grad_weight, float* d_grad_bias, const int64_t* d_lo, cudaStream_t st) {
  const bool ok = in_dim % 4 == 0 && in_dim <= 64 && n_classes <= 64 && a_stride % 4 == 0 && ga_stride % 4 == 0 &&
                  (((uintptr_t)d_a | (uintptr_t)d_weight | (uintptr_t)d_grad_a | (uintptr_t)d_grad_weight) & 15) == 0;
  if (!ok) return PG_ERR_INVALID;
  const size_t smem = (2 * 64 * (size_t)kHeadWS + 2 * 64 * (size_t)kHeadGS + 64 * (size_t)kHeadAS) * sizeof(uint32_t);
  PG_CUDA(cudaFuncSetAttribute(linear_ce_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PG_CUDA(cudaFuncSetAttribute(linear_ce_mma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int grid = (int)((n + 16 * kHeadWarpsM - 1) / (16 * kHeadWarpsM));
  linear_ce_mma_kernel<false><<<grid, kHeadWarpsM * 32, smem, st>>>(d_a, a_stride, d_weight, d_bias, d_labels, n, in_dim,
                                                                    n_classes, 1.0f / (float)n, d_loss, d_grad_a, ga_stride,
                                                                    d_grad_weight, d_grad_bias, d_lo, HeadBlock{});
  PG_CHECK_LAUNCH();
  return PG_OK;
}

// Block + head + loss, forward and backward, in one kernel (linear_ce_mma_kernel<true>). grad_src must be zeroed by the
// caller. PG_ERR_INVALID = layout not eligible, nothing launched.
pg_status block_linear_ce_mma(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_lo3, const float* d_src,
                              int64_t src_stride, int64_t cap_dst, int mode, const float* d_weight, const float* d_bias,
                              const int64_t* d_labels, int32_t in_dim, int32_t n_classes, float* d_loss, float* d_grad_src,
                              int64_t gsrc_stride, float* d_grad_weight, float* d_grad_bias, cudaStream_t st) {
  const bool ok = in_dim % 4 == 0 && in_dim <= 64 && n_classes <= 64 && src_stride % 4 == 0 && gsrc_stride % 4 == 0 &&
                  (((uintptr_t)d_src | (uintptr_t)d_weight | (uintptr_t)d_grad_src | (uintptr_t)d_grad_weight) & 15) == 0;
  if (!ok) return PG_ERR_INVALID;
  const size_t smem = (2 * 64 * (size_t)kHeadWS + 2 * 64 * (size_t)kHeadGS + 64 * (size_t)kHeadAS) * sizeof(uint32_t);
  PG_CUDA(cudaFuncSetAttribute(linear_ce_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PG_CUDA(cudaFuncSetAttribute(linear_ce_mma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int grid = (int)((cap_dst + 16 * kHeadWarpsM - 1) / (16 * kHeadWarpsM));
  HeadBlock blk{d_indptr_base, d_cols, d_lo3, mode == PG_AGG_MEAN ? 1 : 0, d_grad_src, gsrc_stride};
  linear_ce_mma_kernel<true><<<grid, kHeadWarpsM * 32, smem, st>>>(d_src, src_stride, d_weight, d_bias, d_labels, cap_dst,
                                                                   in_dim, n_classes, 1.0f, d_loss, nullptr, 0,
                                                                   d_grad_weight, d_grad_bias, nullptr, blk);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

}  // namespace pg

extern "C" {

pg_status pg_linear_concat_fwd(const float* d_x, int64_t x_stride, const float* d_weight, const float* d_bias, int64_t n,
                               int32_t in_dim, int32_t out_dim, int concat, float* d_out, int64_t out_stride,
                               float* d_out_drop, int64_t od_stride, float dropout_p, uint64_t dropout_seed,
                               const int64_t* d_step, void* stream) {
  PG_REQUIRE(d_weight && n >= 0 && ((d_x && d_out) || n == 0), "pg_linear_concat_fwd: bad arguments");
  PG_REQUIRE(out_dim == kOut, "pg_linear_concat_fwd: out_dim must be 32");
  PG_REQUIRE(in_dim >= 4 && in_dim % 4 == 0 && in_dim <= 768 && x_stride >= in_dim && x_stride % 4 == 0 &&
                 (uintptr_t)d_x % 16 == 0 && (uintptr_t)d_weight % 16 == 0,
             "pg_linear_concat_fwd: in_dim must be a multiple of 4 (<= 768) with 16-byte aligned rows and weight");
  const int width = concat ? 2 * kOut : kOut;
  PG_REQUIRE(out_stride >= width && out_stride % 2 == 0 && (uintptr_t)d_out % 8 == 0 &&
                 (!d_out_drop || (od_stride >= width && od_stride % 2 == 0 && (uintptr_t)d_out_drop % 8 == 0)),
             "pg_linear_concat_fwd: output rows must be 8-byte aligned");
  PG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "pg_linear_concat_fwd: dropout_p must be in [0, 1)");
  if (n == 0) return PG_OK;
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  pg::TimedScope timed(PG_T_DENSE_FWD, st);
  {  // tcgen05 path (pg_dense_umma.cu) whenever the layout allows TMA boxes; PG_FWD_UMMA=0 keeps the mma.sync kernel
    const char* env_u = getenv("PG_FWD_UMMA");
    if (!(env_u && atoi(env_u) == 0)) {
      const pg_status s = pg::linear_concat_fwd_umma(d_x, x_stride, d_weight, d_bias, n, in_dim, concat, d_out, out_stride,
                                                     d_out_drop, od_stride, dropout_p, dropout_seed, d_step, dev, st);
      if (s != PG_ERR_INVALID) return s;
    }
  }
  const size_t smem = 2 * (size_t)kOut * fwd_wstride(in_dim) * sizeof(uint32_t);
  const DropArgs drop = make_drop(d_out_drop ? dropout_p : 0.f, dropout_seed, d_step);
  auto launch = [&](auto kern, int warps, int rows_per_warp) -> pg_status {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t ntiles = (n + warps * rows_per_warp - 1) / (warps * rows_per_warp);
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)pg::sm_count(dev));
    kern<<<grid, warps * 32, smem, st>>>(d_x, x_stride, d_weight, d_bias, n, in_dim, concat, d_out, out_stride, d_out_drop,
                                         od_stride, drop);
    PG_CHECK_LAUNCH();
    return PG_OK;
  };
  const char* env_v = getenv("PG_FWD_VARIANT");
  switch (env_v ? atoi(env_v) : 0) {
    // measured at config 2 (tools/micro_dense.py, us): <2,1,5,8> 45.5, <2,2,3,8> 41.4, <1,2,4,12> 53.7, <1,4,2,12> 61.9,
    // <1,2,3,16> 41.4; cuBLAS fp32 SIMT sgemm + bias 51.7
    case 1: return launch(linear_concat_fwd_kernel<2, 1, 5, 8>, 8, 32);
    case 5: return launch(linear_concat_fwd_kernel<1, 2, 3, 16>, 16, 16);
    default: return launch(linear_concat_fwd_kernel<2, 2, 3, 8>, 8, 32);
  }
}

pg_status pg_linear_concat_bwd(const float* d_x, int64_t x_stride, const float* d_grad_out, int64_t g_stride,
                               const float* d_out, int64_t out_stride, int64_t n, int32_t in_dim, int32_t out_dim, int concat,
                               float dropout_p, uint64_t dropout_seed, const int64_t* d_step, float* d_grad_weight,
                               float* d_grad_bias, void* stream) {
  PG_REQUIRE(d_grad_weight && n >= 0 && ((d_x && d_grad_out && d_out) || n == 0), "pg_linear_concat_bwd: bad arguments");
  PG_REQUIRE(out_dim == kOut, "pg_linear_concat_bwd: out_dim must be 32");
  PG_REQUIRE(in_dim >= 4 && in_dim % 4 == 0 && in_dim <= 64 * kDwMaxWarps && x_stride >= in_dim && x_stride % 4 == 0 &&
                 (uintptr_t)d_x % 16 == 0 && (uintptr_t)d_grad_weight % 16 == 0,
             "pg_linear_concat_bwd: in_dim must be a multiple of 4 (<= 768) with 16-byte aligned rows and grad_weight");
  PG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "pg_linear_concat_bwd: dropout_p must be in [0, 1)");
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  pg::TimedScope timed(PG_T_DENSE_BWD, st);
  PG_CUDA(cudaMemsetAsync(d_grad_weight, 0, (size_t)kOut * in_dim * sizeof(float), st));
  if (d_grad_bias) PG_CUDA(cudaMemsetAsync(d_grad_bias, 0, kOut * sizeof(float), st));
  if (n == 0) return PG_OK;
  {  // tcgen05 path (pg_dense_umma.cu: x^T from tensor memory); PG_DW_UMMA=0 keeps the mma.sync kernel
    const char* env_u = getenv("PG_DW_UMMA");
    if (!(env_u && atoi(env_u) == 0)) {
      const pg_status s = pg::linear_concat_dw_umma(d_x, x_stride, d_grad_out, g_stride, d_out, out_stride, n, in_dim, concat,
                                                    dropout_p, dropout_seed, d_step, d_grad_weight, d_grad_bias, dev, st);
      if (s != PG_ERR_INVALID) return s;
    }
  }
  const char* simt = getenv("PG_DENSE_SIMT");
  if (simt && atoi(simt) && dropout_p == 0.f)
    return pg::linear_concat_bwd_simt(d_x, x_stride, d_grad_out, g_stride, d_out, out_stride, n, in_dim, concat, d_grad_weight,
                                     d_grad_bias, st);
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)pg::sm_count(dev));
  const DropArgs drop = make_drop(dropout_p, dropout_seed, d_step);
  auto launch = [&](auto kern, int warps) -> pg_status {
    pg::prefer_max_smem_k(kern);
    kern<<<grid, warps * 32, 0, st>>>(d_x, x_stride, d_grad_out, g_stride, d_out, out_stride, n, in_dim, concat, drop,
                                      d_grad_weight, d_grad_bias);
    PG_CHECK_LAUNCH();
    return PG_OK;
  };
  const char* env_d = getenv("PG_DW_VARIANT");   // 1 = the 10-warp kernel with per-tile barriers
  if (in_dim <= 32 * kDw2MaxWarps && !(env_d && atoi(env_d) == 1)) {
    const int warps = (in_dim + 31) / 32;
    const size_t smem = 2 * (size_t)kDw2Super * kGzStride * sizeof(uint32_t);
    PG_CUDA(cudaFuncSetAttribute(linear_concat_dw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(cudaFuncSetAttribute(linear_concat_dw2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    linear_concat_dw2_kernel<<<grid, warps * 32, smem, st>>>(d_x, x_stride, d_grad_out, g_stride, d_out, out_stride, n, in_dim,
                                                             concat, drop, d_grad_weight, d_grad_bias);
    PG_CHECK_LAUNCH();
    return PG_OK;
  }
  if (in_dim <= 256) return launch(linear_concat_dw_kernel<4, 255>, 4);
  if (in_dim <= 640) return launch(linear_concat_dw_kernel<10, 168>, 10);
  return launch(linear_concat_dw_kernel<kDwMaxWarps, 168>, kDwMaxWarps);
}

}  // extern "C"
