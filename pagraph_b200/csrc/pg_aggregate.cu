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
#include <cstdint>
#include <cstdlib>

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
  const int64_t* lo;        // optional device-resident extents (pg_common.cuh apply_extents); n_dst is then a capacity
  int64_t zero_rows_to;     // rows [n_dst, zero_rows_to) are zero-filled
};

// VEC-wide columns; LANES lanes per row; each lane holds CH column slices per pass.
template <int LANES, int CH>
__global__ void __launch_bounds__(kAggThreads) agg_fwd_vec4(AggArgs a) {
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const int lane = threadIdx.x & 31, sub = lane % LANES, grp = lane / LANES;
  const int64_t warp0 = (int64_t)blockIdx.x * (kAggThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAggThreads / 32);
  const int nvec = a.dim >> 2;
  if (a.lo) pg::apply_extents(a.lo, a.indptr, a.col_base, a.n_dst);
  for (int64_t r0 = a.n_dst + warp0 * ROWS_PER_WARP; r0 < a.zero_rows_to; r0 += nwarps * ROWS_PER_WARP) {
    const int64_t r = r0 + grp;
    if (r < a.zero_rows_to)
      for (int col = sub; col < nvec; col += LANES) ((float4*)(a.dst + r * a.dst_stride))[col] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
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
  if (a.lo) pg::apply_extents(a.lo, a.indptr, a.col_base, a.n_dst);
  for (int64_t r = a.n_dst + warp0; r < a.zero_rows_to; r += nwarps)
    for (int col = lane; col < a.dim; col += 32) a.dst[r * a.dst_stride + col] = 0.f;
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
  const int64_t* lo;
};

__global__ void __launch_bounds__(kAggThreads) agg_bwd_kernel(AggBwdArgs a, int vec4) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kAggThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAggThreads / 32);
  if (a.lo) pg::apply_extents(a.lo, a.indptr, a.col_base, a.n_dst);
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

// ------------------------------------------------------------------ aggregation from row pointers (fused cache lookup)
// TMA-staged: a warp owns a ring of DEPTH shared-memory buffers of GROUP rows each. For every task (one destination row,
// <= GROUP of its edges) the lanes read cols -> rowptr, lane 0 arms the buffer's mbarrier with the byte count and each
// lane pulls its source row with one cp.async.bulk (a 2400-byte row = one descriptor, no registers held while in
// flight); the warp then sums the staged rows from shared memory (conflict-free 16-byte lanes), applying the dropout
// mask on the fly, and re-arms the buffer for task t+DEPTH. Bytes in flight per SM = warps * DEPTH * GROUP * row bytes,
// independent of register pressure; every source row is read from HBM once per edge, every dst row written once.
//
// Instruction diet (r2, from the ncu source view of the r1 kernel: 316 warp instructions per edge of which 85 were the
// loads + mask + accumulate): the issuing side is the only task cursor — it leaves {count, last, row, degree} of each
// task in a shared-memory ring slot that the consuming side reads back with one LDS.128, all in 32-bit; the dropout
// column keys live in shared memory (under the 96-register cap the compiler re-derived the five splitmix64 keys in
// every round); the row width is a template constant for the 600-float rows of the reference (the per-chunk `col < nvec`
// tests become compile-time, one predicated chunk instead of five divergence regions); shared memory is addressed with
// 32-bit shared-window addresses; the mean is one reciprocal per row and a multiply, not 4 divisions per float4.
constexpr int kRowsMaxDepth = 4;
constexpr int kRowsMaxGroup = 16;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// acc += drop(v): the k-th 16-bit lane of h decides component k. The upper lanes are compared in place
// (x >> 16 >= thr  <=>  x >= thr << 16), so a float4 costs 2 shifts + 4 compares + 4 predicated FMAs.
__device__ __forceinline__ void acc_drop4(float4& acc, const float4 v, uint64_t h, uint32_t thr_hi, float scale) {
  const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
  if ((lo << 16) >= thr_hi) acc.x = fmaf(v.x, scale, acc.x);
  if (lo >= thr_hi) acc.y = fmaf(v.y, scale, acc.y);
  if ((hi << 16) >= thr_hi) acc.z = fmaf(v.z, scale, acc.z);
  if (hi >= thr_hi) acc.w = fmaf(v.w, scale, acc.w);
}

// NVEC: float4 per row when known at compile time (150 for the reference's 600-float features), 0 = a.dim / 4.
// Rows are consumed two at a time (2 * CH independent shared-memory loads in flight); 96 registers so that 16 warps
// leave room for the sampler's small CTAs on the other stream.
template <int W, int CH, bool DROP, int NVEC>
__global__ void __launch_bounds__(W * 32) __maxnreg__(NVEC >= 0 ? 96 : 80)
    agg_rows_tma_kernel(pg::AggRowsArgs a, int group, int depth) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[W * kRowsMaxDepth];
  __shared__ uint64_t row_key[W][kRowsMaxDepth][kRowsMaxGroup];  // dropout row key of every staged row
  __shared__ int4 task_meta[W][kRowsMaxDepth];                    // {rows staged, last task of its dst row, dst row, degree}
  __shared__ uint64_t col_key[DROP ? CH * 32 : 1];                // dropout column keys (one per float4 column group)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nvec = NVEC ? NVEC : (a.dim >> 2);
  const uint32_t row_bytes = (uint32_t)nvec * 16u;
  const uint32_t warp_smem = pg::smem_u32(smem) + (uint32_t)w * (uint32_t)(depth * group) * row_bytes;
  if (lane < depth) pg::mbar_init(pg::smem_u32(&bars[w * kRowsMaxDepth + lane]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __shared__ uint64_t stepkey_s;   // dropout step key (read once per task: not worth two registers)
  if (DROP) {
    for (int i = threadIdx.x; i < CH * 32; i += W * 32) col_key[i] = pg::drop_colkey((uint32_t)i);
    if (threadIdx.x == 0) stepkey_s = pg::drop_stepkey(a.drop_seed + (a.drop_step ? (uint64_t)*a.drop_step : 0ull));
    __syncthreads();
  } else {
    __syncwarp();
  }
  const int warp0 = (int)(blockIdx.x * W + w), nwarps = (int)(gridDim.x * W);
  const int64_t cap_dst = a.n_dst;
  if (a.lo) pg::apply_extents(a.lo, a.indptr, a.col_base, a.n_dst);
  a.zero_rows_to = pg::resolve_zero_rows(a.zero_rows_to, a.n_dst, cap_dst);
  const int n_dst = (int)a.n_dst;
  const uint32_t thr_hi = a.drop_thr << 16;
  const float keep_scale = a.keep_scale;
  const bool mean = a.mode == PG_AGG_MEAN;

  // ---- issuing side: the only task cursor. (ir, ipos, irem) = row being issued, absolute position of its next edge,
  // edges left; the row after it is opened one step ahead so that its indptr reads are off the critical path.
  int ir = warp0, irem = 0, ideg = 0;
  int64_t ipos = 0, nxt_s = 0, nxt_e = 0;
  auto open_next = [&](int r) {  // prefetch (s, e) of row r
    if (r < n_dst) {
      nxt_s = a.indptr[r];
      nxt_e = a.indptr[r + 1];
    }
  };
  open_next(ir);
  if (ir < n_dst) {
    ipos = nxt_s;
    ideg = irem = (int)(nxt_e - nxt_s);
    open_next(ir + nwarps);
  }
  // The task to be issued is PREPARED two steps ahead, one dependent load per step: step t reads the column ids of task
  // t + 2 (`fetch`: the cursor advances here) and turns the ids of task t + 1 into row pointers + dropout row keys
  // (`resolve`). Each of the two global loads of the cols -> rowptr chain — the longest stall of the r2a kernel in the
  // ncu source view — then has a whole task period to land instead of both sharing one, which is what lets short tasks
  // (small GROUP, i.e. a small shared-memory footprint that leaves room for the dense stage's CTAs) keep up.
  const uint64_t pol_hot = a.hints ? pg::l2_policy_evict_last() : pg::l2_policy_evict_normal();
  const uint64_t pol_cold = a.hints == 1 ? pg::l2_policy_evict_first() : pg::l2_policy_evict_normal();
  int qj = 0;                    // fetched task: this lane's source index (lane < qmeta.x)
  int3 qmeta = make_int3(0, 0, 0);   // {rows | last-task-of-its-row << 8, dst row, degree}
  bool qhave = false;
  const float* psrc = nullptr;   // resolved task: this lane's source row (lane < pmeta.x)
  uint64_t pkey = 0;             // its dropout row key
  int3 pmeta = make_int3(0, 0, 0);
  bool phave = false;
  auto fetch = [&]() {  // reads the cursor, loads the task's column ids, advances the cursor
    qhave = ir < n_dst;
    if (!qhave) return;
    const int cnt = min(group, irem);
    qmeta = make_int3(cnt | (irem <= group ? 256 : 0), ir, ideg);
    if (lane < cnt) qj = (int)(a.cols[ipos + lane] - a.col_base);
    ipos += cnt;
    irem -= cnt;
    if (irem <= 0) {  // next destination row
      ir += nwarps;
      if (ir < n_dst) {
        ipos = nxt_s;
        ideg = irem = (int)(nxt_e - nxt_s);
        open_next(ir + nwarps);
      }
    }
  };
  auto resolve = [&]() {  // fetched -> resolved
    phave = qhave;
    pmeta = qmeta;
    if (qhave && lane < (qmeta.x & 255)) {
      psrc = a.rowptr[qj];
      if (DROP) pkey = pg::drop_rowkey(stepkey_s, (uint64_t)qj);  // one hash per fetched row, lanes in parallel
    }
  };
  auto issue = [&](int buf) {  // precondition: phave
    const uint32_t bar = pg::smem_u32(&bars[w * kRowsMaxDepth + buf]);
    const int pcnt = pmeta.x & 255;
    if (DROP && lane < pcnt) row_key[w][buf][lane] = pkey;
    if (lane == 0) {
      task_meta[w][buf] = make_int4(pcnt, pmeta.x >> 8, pmeta.y, pmeta.z);
      pg::mbar_expect_tx(bar, (uint32_t)pcnt * row_bytes);
    }
    __syncwarp();
    if (lane < pcnt) {
      // bit 0 of a row pointer = hot row (kept in L2: hub vertices recur within and across minibatches); the read-once
      // rest streams through with evict_first so that it does not push the hot rows (or the output rows) out
      const uintptr_t ps = (uintptr_t)psrc;
      pg::bulk_g2s_hint(warp_smem + (uint32_t)(buf * group + lane) * row_bytes, (const void*)(ps & ~(uintptr_t)1), row_bytes, bar,
                        (ps & 1) ? pol_hot : pol_cold);
    }
    resolve();
    fetch();
  };
  fetch();
  resolve();
  fetch();
  int issued = 0;
  for (; issued < depth && phave; ++issued) issue(issued);

  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t phase_bits = 0;  // bit b = parity to wait for on buffer b
  int buf = 0;
  const uint32_t lane_off = (uint32_t)lane * 16u;
  // consumed tasks == issued tasks: the loop ends when nothing issued is left unconsumed
  for (int pending = issued; pending > 0;) {
    const uint32_t bar = pg::smem_u32(&bars[w * kRowsMaxDepth + buf]);
    while (!pg::mbar_try_wait(bar, (phase_bits >> buf) & 1u)) {
    }
    phase_bits ^= 1u << buf;
    const int4 meta = task_meta[w][buf];
    const int cnt = meta.x;
    const uint32_t base = warp_smem + (uint32_t)(buf * group) * row_bytes + lane_off;
    int k = 0;
    for (; k + 2 <= cnt; k += 2) {  // two staged rows per round
      const uint32_t r0 = base + (uint32_t)k * row_bytes, r1 = r0 + row_bytes;
      float4 v0[CH], v1[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (c * 32 + lane < nvec) {
          v0[c] = lds128(r0 + c * 512);
          v1[c] = lds128(r1 + c * 512);
        }
      uint64_t j0 = 0, j1 = 0;
      if (DROP) {
        j0 = row_key[w][buf][k];
        j1 = row_key[w][buf][k + 1];
      }
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (c * 32 + lane < nvec) {
          if (DROP) {
            const uint64_t ck = col_key[c * 32 + lane];
            acc_drop4(acc[c], v0[c], pg::drop_mix(j0, ck), thr_hi, keep_scale);
            acc_drop4(acc[c], v1[c], pg::drop_mix(j1, ck), thr_hi, keep_scale);
          } else {
            acc[c].x = (acc[c].x + v0[c].x) + v1[c].x; acc[c].y = (acc[c].y + v0[c].y) + v1[c].y;
            acc[c].z = (acc[c].z + v0[c].z) + v1[c].z; acc[c].w = (acc[c].w + v0[c].w) + v1[c].w;
          }
        }
    }
    if (k < cnt) {
      const uint32_t r0 = base + (uint32_t)k * row_bytes;
      const uint64_t j0 = DROP ? row_key[w][buf][k] : 0ull;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (c * 32 + lane < nvec) {
          const float4 v = lds128(r0 + c * 512);
          if (DROP) {
            acc_drop4(acc[c], v, pg::drop_mix(j0, col_key[c * 32 + lane]), thr_hi, keep_scale);
          } else {
            acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
          }
        }
    }
    if (meta.y) {  // row complete: scale, write once, reset
      float sc = mean ? 1.0f / (float)max(meta.w, 1) : 1.0f;
      if (a.norm) sc *= a.norm[meta.z];
      float4* out = (float4*)(a.dst + (int64_t)meta.z * a.dst_stride);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (c * 32 + lane < nvec) out[c * 32 + lane] = make_float4(acc[c].x * sc, acc[c].y * sc, acc[c].z * sc, acc[c].w * sc);
        acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    --pending;
    __syncwarp();  // every lane is done reading this buffer (and its meta / row keys)
    if (phave) {
      issue(buf);
      ++pending;
    }
    buf = (buf + 1 == depth) ? 0 : buf + 1;
  }
  // zero the padding rows of a fixed-shape destination
  for (int64_t r = a.n_dst + warp0; r < a.zero_rows_to; r += nwarps) {
    float4* out = (float4*)(a.dst + r * a.dst_stride);
    for (int col = lane; col < nvec; col += 32) out[col] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Any width / alignment: one warp per row, scalar columns, plain loads through the row pointers.
__global__ void __launch_bounds__(kAggThreads) agg_rows_ldg_kernel(pg::AggRowsArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kAggThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAggThreads / 32);
  const uint64_t seed = a.drop_thr ? a.drop_seed + (a.drop_step ? (uint64_t)*a.drop_step : 0ull) : 0ull;
  const int64_t cap_dst = a.n_dst;
  if (a.lo) pg::apply_extents(a.lo, a.indptr, a.col_base, a.n_dst);
  a.zero_rows_to = pg::resolve_zero_rows(a.zero_rows_to, a.n_dst, cap_dst);
  for (int64_t r = warp0; r < a.zero_rows_to || r < a.n_dst; r += nwarps) {
    if (r >= a.n_dst) {
      for (int col = lane; col < a.dim; col += 32) a.dst[r * a.dst_stride + col] = 0.f;
      continue;
    }
    const int64_t s = a.indptr[r], e = a.indptr[r + 1];
    const float deg = (float)max(e - s, (int64_t)1);
    for (int col = lane; col < a.dim; col += 32) {
      float acc = 0.f;
      for (int64_t q = s; q < e; ++q) {
        const int64_t j = a.cols[q] - a.col_base;
        float v = __ldg((const float*)((uintptr_t)a.rowptr[j] & ~(uintptr_t)1) + col);
        if (a.drop_thr) {
          const uint64_t h = pg::drop_hash(seed, (uint64_t)j, (uint32_t)(col >> 2));
          v = ((uint32_t)((h >> (16 * (col & 3))) & 0xffff) < a.drop_thr) ? 0.f : v * a.keep_scale;
        }
        acc += v;
      }
      if (a.mode == PG_AGG_MEAN) acc /= deg;
      if (a.norm) acc *= a.norm[r];
      a.dst[r * a.dst_stride + col] = acc;
    }
  }
}

template <int W, int CH>
pg_status launch_rows_tma_w(const pg::AggRowsArgs& a, int dev, cudaStream_t st, size_t budget, int depth_req) {
  const size_t row_bytes = (size_t)a.dim * 4;
  const int slots = (int)(budget / ((size_t)W * row_bytes));
  if (slots < 2) return PG_ERR_INVALID;  // caller tries fewer warps, then the plain-load kernel
  int depth = depth_req > 0 ? depth_req : 2;
  depth = std::max(1, std::min({depth, kRowsMaxDepth, slots}));
  const int group = std::min(kRowsMaxGroup, slots / depth);
  const size_t smem = (size_t)W * depth * group * row_bytes;
  const bool drop = a.drop_thr != 0;
  // the reference's 600-float rows get the compile-time width (CH == 5 covers 129..160 float4)
  constexpr int kRefVec = 150;
  const bool ref_width = CH == 5 && a.dim == 4 * kRefVec && !getenv("PG_AGG_GENERIC");
  auto kern = ref_width ? (drop ? agg_rows_tma_kernel<W, CH, true, (CH == 5 ? kRefVec : 0)>
                                : agg_rows_tma_kernel<W, CH, false, (CH == 5 ? kRefVec : 0)>)
                        : (drop ? agg_rows_tma_kernel<W, CH, true, 0> : agg_rows_tma_kernel<W, CH, false, 0>);
  PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int64_t rows = std::max(a.n_dst, a.zero_rows_to);   // a negative zero_rows_to is bounded by the capacity n_dst
  const int64_t need = std::max<int64_t>(1, (rows + W - 1) / W);
  // One CTA fills an SM's shared memory, so nothing that needs more than ~27 KB of it (the classifier head: 93 KB) can
  // run beside this kernel. A pipeline that wants its small latency-bound kernels to proceed while the input block of the
  // next minibatch is aggregated leaves a few SMs out of this grid (PG_AGG_RESERVE_SMS / pg_set_agg_reserve_sms).
  const int sms = std::max(1, pg::sm_count(dev) - pg::agg_reserve_sms());
  const int grid = (int)std::min<int64_t>(need, (int64_t)sms);
  kern<<<grid, W * 32, smem, st>>>(a, group, depth);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

template <int CH>
pg_status launch_rows_tma(const pg::AggRowsArgs& a, int dev, cudaStream_t st) {
  static int max_optin = 0;
  if (!max_optin) cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const char* env_b = getenv("PG_AGG_SMEM");
  const size_t budget = std::min<size_t>((size_t)max_optin - 4096, env_b ? (size_t)atoi(env_b) : 200 * 1024);
  const char* env_d = getenv("PG_AGG_DEPTH");
  const char* env_w = getenv("PG_AGG_WARPS");
  const int depth = env_d ? atoi(env_d) : 0;
  // measured at config 2 (tools/micro_fused.py, resolve + kernel, us): 16 warps x (depth 2 x group 2) 154, 12 x (2 x 3) 152,
  // 8 x (2 x 5) 150, 8 x (3 x 3) 173: after the instruction diet the kernel is latency-bound on the row fetches, so fewer
  // warps with more rows per task win — and 8 warps x 88 registers leave 2/3 of the register file to the other streams
  const int warps = env_w ? atoi(env_w) : 8;
  pg_status s = PG_ERR_INVALID;
  if (warps >= 16) s = launch_rows_tma_w<16, CH>(a, dev, st, budget, depth);
  if (s == PG_ERR_INVALID && warps >= 12) s = launch_rows_tma_w<12, CH>(a, dev, st, budget, depth);
  if (s == PG_ERR_INVALID && warps >= 8) s = launch_rows_tma_w<8, CH>(a, dev, st, budget, depth);
  if (s == PG_ERR_INVALID) s = launch_rows_tma_w<4, CH>(a, dev, st, budget, depth);
  return s;
}

template <int LANES, int CH>
void launch_fwd(const AggArgs& a, int dev, cudaStream_t st) {
  constexpr int rows_per_block = (kAggThreads / 32) * (32 / LANES);
  const int64_t need = std::max<int64_t>(1, (a.n_dst + rows_per_block - 1) / rows_per_block);
  const int grid = (int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16);
  pg::prefer_max_smem_k(agg_fwd_vec4<LANES, CH>);
  agg_fwd_vec4<LANES, CH><<<grid, kAggThreads, 0, st>>>(a);
}

}  // namespace

namespace pg {
pg_status launch_agg_rows(const AggRowsArgs& a, int dev, cudaStream_t st) {
  const int nvec = a.dim / 4;
  // the TMA kernel keeps row indices in 32 bits (a NodeFlow layer / a row block of the server-side fold)
  const bool tma_ok = (a.dim % 4 == 0) && (a.dst_stride % 4 == 0) && ((uintptr_t)a.dst % 16 == 0) && nvec <= 8 * 32 &&
                      std::max(a.n_dst, a.zero_rows_to) < (int64_t)INT32_MAX - (1 << 24) && !getenv("PG_AGG_NO_TMA");
  if (tma_ok) {
    pg_status s = PG_ERR_INVALID;
    if (nvec <= 32) s = launch_rows_tma<1>(a, dev, st);
    else if (nvec <= 64) s = launch_rows_tma<2>(a, dev, st);
    else if (nvec <= 96) s = launch_rows_tma<3>(a, dev, st);
    else if (nvec <= 128) s = launch_rows_tma<4>(a, dev, st);
    else if (nvec <= 160) s = launch_rows_tma<5>(a, dev, st);
    else s = launch_rows_tma<8>(a, dev, st);
    if (s != PG_ERR_INVALID) return s;
  }
  const int64_t rows = std::max(a.n_dst, a.zero_rows_to);
  const int64_t need = std::max<int64_t>(1, (rows + kAggThreads / 32 - 1) / (kAggThreads / 32));
  pg::prefer_max_smem_k(agg_rows_ldg_kernel);
  agg_rows_ldg_kernel<<<(int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16), kAggThreads, 0, st>>>(a);
  PG_CHECK_LAUNCH();
  return PG_OK;
}
}  // namespace pg

extern "C" {

static pg_status aggregate_fwd_impl(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base, const float* d_src,
                                    int64_t src_stride, float* d_dst, int64_t dst_stride, int64_t n_dst, int32_t dim,
                                    int mode, const float* d_norm, const int64_t* d_lo, int64_t zero_rows_to, void* stream);

pg_status pg_aggregate_fwd(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base, const float* d_src,
                           int64_t src_stride, float* d_dst, int64_t dst_stride, int64_t n_dst, int32_t dim, int mode,
                           const float* d_norm, void* stream) {
  return aggregate_fwd_impl(d_indptr, d_cols, col_base, d_src, src_stride, d_dst, dst_stride, n_dst, dim, mode, d_norm,
                            nullptr, 0, stream);
}

pg_status pg_aggregate_fwd_dyn(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_layer_offsets,
                               const float* d_src, int64_t src_stride, float* d_dst, int64_t dst_stride, int64_t cap_dst,
                               int32_t dim, int mode, const float* d_norm, void* stream) {
  PG_REQUIRE(d_layer_offsets != nullptr, "pg_aggregate_fwd_dyn: null layer offsets");
  return aggregate_fwd_impl(d_indptr_base, d_cols, 0, d_src, src_stride, d_dst, dst_stride, cap_dst, dim, mode, d_norm,
                            d_layer_offsets, cap_dst, stream);
}

static pg_status aggregate_fwd_impl(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base, const float* d_src,
                                    int64_t src_stride, float* d_dst, int64_t dst_stride, int64_t n_dst, int32_t dim,
                                    int mode, const float* d_norm, const int64_t* d_lo, int64_t zero_rows_to, void* stream) {
  PG_REQUIRE(d_indptr && d_dst && n_dst >= 0 && dim >= 1, "pg_aggregate_fwd: bad arguments");
  PG_REQUIRE(mode == PG_AGG_SUM || mode == PG_AGG_MEAN, "pg_aggregate_fwd: mode must be PG_AGG_SUM or PG_AGG_MEAN");
  PG_REQUIRE(src_stride >= dim && dst_stride >= dim, "pg_aggregate_fwd: stride smaller than dim");
  if (n_dst == 0) return PG_OK;
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  AggArgs a{d_indptr, d_cols, col_base, d_src, src_stride, d_dst, dst_stride, n_dst, dim, mode, d_norm, d_lo, zero_rows_to};
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
    pg::prefer_max_smem_k(agg_fwd_scalar);
    agg_fwd_scalar<<<(int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16), kAggThreads, 0, st>>>(a);
  }
  PG_CHECK_LAUNCH();
  return PG_OK;
}

static pg_status aggregate_bwd_impl(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base,
                                    const float* d_grad_dst, int64_t gdst_stride, float* d_grad_src, int64_t gsrc_stride,
                                    int64_t n_dst, int64_t n_src, int32_t dim, int mode, const float* d_norm,
                                    const int64_t* d_lo, void* stream);

pg_status pg_aggregate_bwd(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base, const float* d_grad_dst,
                           int64_t gdst_stride, float* d_grad_src, int64_t gsrc_stride, int64_t n_dst, int64_t n_src,
                           int32_t dim, int mode, const float* d_norm, void* stream) {
  return aggregate_bwd_impl(d_indptr, d_cols, col_base, d_grad_dst, gdst_stride, d_grad_src, gsrc_stride, n_dst, n_src, dim,
                            mode, d_norm, nullptr, stream);
}

pg_status pg_aggregate_bwd_dyn(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_layer_offsets,
                               const float* d_grad_dst, int64_t gdst_stride, float* d_grad_src, int64_t gsrc_stride,
                               int64_t cap_dst, int64_t cap_src, int32_t dim, int mode, const float* d_norm, void* stream) {
  PG_REQUIRE(d_layer_offsets != nullptr, "pg_aggregate_bwd_dyn: null layer offsets");
  return aggregate_bwd_impl(d_indptr_base, d_cols, 0, d_grad_dst, gdst_stride, d_grad_src, gsrc_stride, cap_dst, cap_src, dim,
                            mode, d_norm, d_layer_offsets, stream);
}

static pg_status aggregate_bwd_impl(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base,
                                    const float* d_grad_dst, int64_t gdst_stride, float* d_grad_src, int64_t gsrc_stride,
                                    int64_t n_dst, int64_t n_src, int32_t dim, int mode, const float* d_norm,
                                    const int64_t* d_lo, void* stream) {
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
  AggBwdArgs a{d_indptr, d_cols, col_base, d_grad_dst, gdst_stride, d_grad_src, gsrc_stride, n_dst, dim, mode, d_norm, d_lo};
  const int vec4 = (dim % 4 == 0) && (gdst_stride % 4 == 0) && (gsrc_stride % 4 == 0) &&
                   (((uintptr_t)d_grad_dst | (uintptr_t)d_grad_src) % 16 == 0);
  const int64_t need = std::max<int64_t>(1, (n_dst + kAggThreads / 32 - 1) / (kAggThreads / 32));
  pg::prefer_max_smem_k(agg_bwd_kernel);
  agg_bwd_kernel<<<(int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * 16), kAggThreads, 0, st>>>(a, vec4);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

}  // extern "C"
