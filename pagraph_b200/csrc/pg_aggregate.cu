// pg_aggregate.cu — per-block sparse sum/mean aggregation (forward + backward).
//
// Replaces dgl==0.4.1's fused copy_src + sum/mean reducer behind nf.block_compute(...) (reference
// call sites PaGraph/model/gcn_nssc.py:71-74,94-97,139-142,159-162; graphsage_nssc.py:98-106) and,
// run over the whole graph, the server-side --preprocess fold (server/pa_server.py:45-52).
// Semantics: SURVEY.md Appendix A.5 (zero-in-degree rows -> 0; mean = sum / max(deg,1)).
//
// B200 design (HBM-bound segmented reduce, fp32 FMA-free adds — no tensor cores):
//   * a group of LANES lanes owns one destination row; each lane keeps its float4 column slices of
//     the accumulator in registers (600 floats = 150 float4 = 5 per lane at LANES=32), walks the
//     row's edges 4 at a time so up to 20 independent 16-byte loads are in flight per lane, divides
//     once and writes the row once: each source row is read once per edge, each dst row written once;
//   * no degree pass, no atomics, deterministic edge order in the forward;
//   * backward scatters grad rows with 16-byte vector atomics (red.global.add.v4.f32).
#include <algorithm>

#include "pg_common.cuh"

namespace {

constexpr int kAggThreads = 256;

struct AggArgs {
  const int64_t* indptr;
  const int64_t* cols;
  int64_t col_base;
  const float* src;
  int64_t src_stride;
  float* dst;
  int64_t dst_stride;
  int64_t n_dst;
  int dim;
  int mode;
  const float* norm;
};

// VEC-wide columns; LANES lanes per row; each lane holds CH column slices per pass.
template <int LANES, int CH>
__global__ void __launch_bounds__(kAggThreads) agg_fwd_vec4(AggArgs a) {
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const int lane = threadIdx.x & 31, sub = lane % LANES, grp = lane / LANES;
  const int64_t warp0 = (int64_t)blockIdx.x * (kAggThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAggThreads / 32);
  const int nvec = a.dim >> 2;
  for (int64_t r0 = warp0 * ROWS_PER_WARP; r0 < a.n_dst; r0 += nwarps * ROWS_PER_WARP) {
    const int64_t r = r0 + grp;
    if (r >= a.n_dst) continue;
    const int64_t s = a.indptr[r], e = a.indptr[r + 1];
    const bool mean = a.mode == PG_AGG_MEAN;
    const float deg = (float)max(e - s, (int64_t)1);
    const float nrm = a.norm ? a.norm[r] : 1.0f;
    float4* out = (float4*)(a.dst + r * a.dst_stride);
    for (int c0 = 0; c0 < nvec; c0 += LANES * CH) {
      float4 acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      int64_t j = s;
      for (; j + 4 <= e; j += 4) {
        const float4* p0 = (const float4*)(a.src + (a.cols[j] - a.col_base) * a.src_stride);
        const float4* p1 = (const float4*)(a.src + (a.cols[j + 1] - a.col_base) * a.src_stride);
        const float4* p2 = (const float4*)(a.src + (a.cols[j + 2] - a.col_base) * a.src_stride);
        const float4* p3 = (const float4*)(a.src + (a.cols[j + 3] - a.col_base) * a.src_stride);
        float4 v0[CH], v1[CH], v2[CH], v3[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int col = c0 + c * LANES + sub;
          if (col < nvec) {
            v0[c] = __ldg(p0 + col); v1[c] = __ldg(p1 + col); v2[c] = __ldg(p2 + col); v3[c] = __ldg(p3 + col);
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int col = c0 + c * LANES + sub;
          if (col < nvec) {  // sequential edge order (matches a serial sum)
            acc[c].x = (((acc[c].x + v0[c].x) + v1[c].x) + v2[c].x) + v3[c].x;
            acc[c].y = (((acc[c].y + v0[c].y) + v1[c].y) + v2[c].y) + v3[c].y;
            acc[c].z = (((acc[c].z + v0[c].z) + v1[c].z) + v2[c].z) + v3[c].z;
            acc[c].w = (((acc[c].w + v0[c].w) + v1[c].w) + v2[c].w) + v3[c].w;
          }
        }
      }
      for (; j < e; ++j) {
        const float4* p = (const float4*)(a.src + (a.cols[j] - a.col_base) * a.src_stride);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int col = c0 + c * LANES + sub;
          if (col < nvec) {
            const float4 v = __ldg(p + col);
            acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int col = c0 + c * LANES + sub;
        if (col < nvec) {
          float4 v = acc[c];
          if (mean) { v.x /= deg; v.y /= deg; v.z /= deg; v.w /= deg; }
          if (a.norm) { v.x *= nrm; v.y *= nrm; v.z *= nrm; v.w *= nrm; }
          out[col] = v;
        }
      }
    }
  }
}

// Any width / alignment: one warp per row, scalar columns.
__global__ void __launch_bounds__(kAggThreads) agg_fwd_scalar(AggArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kAggThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAggThreads / 32);
  for (int64_t r = warp0; r < a.n_dst; r += nwarps) {
    const int64_t s = a.indptr[r], e = a.indptr[r + 1];
    const float deg = (float)max(e - s, (int64_t)1);
    for (int col = lane; col < a.dim; col += 32) {
      float acc = 0.f;
      for (int64_t j = s; j < e; ++j) acc += __ldg(a.src + (a.cols[j] - a.col_base) * a.src_stride + col);
      if (a.mode == PG_AGG_MEAN) acc /= deg;
      if (a.norm) acc *= a.norm[r];
      a.dst[r * a.dst_stride + col] = acc;
    }
  }
}

struct AggBwdArgs {
  const int64_t* indptr;
  const int64_t* cols;
  int64_t col_base;
  const float* gdst;
  int64_t gdst_stride;
  float* gsrc;
  int64_t gsrc_stride;
  int64_t n_dst;
  int dim;
  int mode;
  const float* norm;
};

__global__ void __launch_bounds__(kAggThreads) agg_bwd_kernel(AggBwdArgs a, int vec4) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kAggThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAggThreads / 32);
  for (int64_t r = warp0; r < a.n_dst; r += nwarps) {
    const int64_t s = a.indptr[r], e = a.indptr[r + 1];
    if (e == s) continue;
    float scale = 1.0f;
    const float deg = (float)max(e - s, (int64_t)1);
    if (a.norm) scale *= a.norm[r];
    const float* g = a.gdst + r * a.gdst_stride;
    if (vec4) {
      const int nvec = a.dim >> 2;
      for (int col = lane; col < nvec; col += 32) {
        float4 v = __ldg((const float4*)g + col);
        if (a.mode == PG_AGG_MEAN) { v.x /= deg; v.y /= deg; v.z /= deg; v.w /= deg; }
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        for (int64_t j = s; j < e; ++j) {
          float4* d = (float4*)(a.gsrc + (a.cols[j] - a.col_base) * a.gsrc_stride) + col;
          atomicAdd(d, v);
        }
      }
    } else {
      for (int col = lane; col < a.dim; col += 32) {
        float v = __ldg(g + col);
        if (a.mode == PG_AGG_MEAN) v /= deg;
        v *= scale;
        for (int64_t j = s; j < e; ++j) atomicAdd(a.gsrc + (a.cols[j] - a.col_base) * a.gsrc_stride + col, v);
      }
    }
  }
}

template <int LANES, int CH>
void launch_fwd(const AggArgs& a, int dev, cudaStream_t st) {
  constexpr int rows_per_block = (kAggThreads / 32) * (32 / LANES);
  const int64_t need = std::max<int64_t>(1, (a.n_dst + rows_per_block - 1) / rows_per_block);
  const int grid = (int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16);
  agg_fwd_vec4<LANES, CH><<<grid, kAggThreads, 0, st>>>(a);
}

}  // namespace

extern "C" {

pg_status pg_aggregate_fwd(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base, const float* d_src,
                           int64_t src_stride, float* d_dst, int64_t dst_stride, int64_t n_dst, int32_t dim, int mode,
                           const float* d_norm, void* stream) {
  PG_REQUIRE(d_indptr && d_dst && n_dst >= 0 && dim >= 1, "pg_aggregate_fwd: bad arguments");
  PG_REQUIRE(mode == PG_AGG_SUM || mode == PG_AGG_MEAN, "pg_aggregate_fwd: mode must be PG_AGG_SUM or PG_AGG_MEAN");
  PG_REQUIRE(src_stride >= dim && dst_stride >= dim, "pg_aggregate_fwd: stride smaller than dim");
  if (n_dst == 0) return PG_OK;
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  AggArgs a{d_indptr, d_cols, col_base, d_src, src_stride, d_dst, dst_stride, n_dst, dim, mode, d_norm};
  const bool vec4 = (dim % 4 == 0) && (src_stride % 4 == 0) && (dst_stride % 4 == 0) &&
                    (((uintptr_t)d_src | (uintptr_t)d_dst) % 16 == 0);
  pg::TimedScope timed(PG_T_AGG_FWD, st);
  if (vec4) {
    const int nvec = dim / 4;
    if (nvec <= 8) launch_fwd<8, 1>(a, dev, st);
    else if (nvec <= 16) launch_fwd<16, 1>(a, dev, st);
    else if (nvec <= 32) launch_fwd<32, 1>(a, dev, st);
    else if (nvec <= 64) launch_fwd<32, 2>(a, dev, st);
    else if (nvec <= 96) launch_fwd<32, 3>(a, dev, st);
    else if (nvec <= 128) launch_fwd<32, 4>(a, dev, st);
    else launch_fwd<32, 5>(a, dev, st);  // 600 floats = 150 float4: one pass; wider rows loop in passes of 160
  } else {
    const int64_t need = std::max<int64_t>(1, (n_dst + kAggThreads / 32 - 1) / (kAggThreads / 32));
    agg_fwd_scalar<<<(int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16), kAggThreads, 0, st>>>(a);
  }
  PG_CHECK_LAUNCH();
  return PG_OK;
}

pg_status pg_aggregate_bwd(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base, const float* d_grad_dst,
                           int64_t gdst_stride, float* d_grad_src, int64_t gsrc_stride, int64_t n_dst, int64_t n_src,
                           int32_t dim, int mode, const float* d_norm, void* stream) {
  PG_REQUIRE(d_indptr && d_grad_src && n_dst >= 0 && n_src >= 0 && dim >= 1, "pg_aggregate_bwd: bad arguments");
  PG_REQUIRE(mode == PG_AGG_SUM || mode == PG_AGG_MEAN, "pg_aggregate_bwd: mode must be PG_AGG_SUM or PG_AGG_MEAN");
  PG_REQUIRE(gdst_stride >= dim && gsrc_stride >= dim, "pg_aggregate_bwd: stride smaller than dim");
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  pg::TimedScope timed(PG_T_AGG_BWD, st);
  if (n_src > 0) {
    if (gsrc_stride == dim) {
      PG_CUDA(cudaMemsetAsync(d_grad_src, 0, (size_t)n_src * dim * sizeof(float), st));
    } else {
      PG_CUDA(cudaMemset2DAsync(d_grad_src, (size_t)gsrc_stride * 4, 0, (size_t)dim * 4, (size_t)n_src, st));
    }
  }
  if (n_dst == 0 || n_src == 0) return PG_OK;
  AggBwdArgs a{d_indptr, d_cols, col_base, d_grad_dst, gdst_stride, d_grad_src, gsrc_stride, n_dst, dim, mode, d_norm};
  const int vec4 = (dim % 4 == 0) && (gdst_stride % 4 == 0) && (gsrc_stride % 4 == 0) &&
                   (((uintptr_t)d_grad_dst | (uintptr_t)d_grad_src) % 16 == 0);
  const int64_t need = std::max<int64_t>(1, (n_dst + kAggThreads / 32 - 1) / (kAggThreads / 32));
  agg_bwd_kernel<<<(int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16), kAggThreads, 0, st>>>(a, vec4);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

}  // extern "C"
