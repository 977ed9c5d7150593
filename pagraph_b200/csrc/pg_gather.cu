// pg_gather.cu — GPU feature-cache lookup, hit/miss split, HBM-cache gather and host-row fetch.
//
// Replaces PaGraph/storage/storage.py (reference): fetch_data :157-204, fetch_from_cache :207-216,
// cache_fix_data :135-154, get_feat_from_server :107-132. Payload is copied bit-for-bit.
//
// B200 design (byte-moving work: HBM-bound for hits, PCIe-bound for misses):
//   * ONE split kernel classifies every NodeFlow node of every layer (flag lookup, ballot +
//     block-aggregated atomics) into a hit list (-> cache row) and a miss list (-> full-graph row):
//     no boolean-mask indexing, no host synchronisation (the reference syncs >= 3x per layer);
//   * hit rows: one warp per row, 16-byte read-only loads from the HBM cache table, all loads of a
//     row in flight before the first store, rows written straight into NodeFlow order;
//   * miss rows: the pinned host table is read directly by the GPU. Each lane of a fetch warp owns a
//     shared-memory stage + mbarrier and moves one row with TMA bulk copies
//     (cp.async.bulk global->shared, then shared->global), so a 2400-byte row is two descriptors
//     instead of 150 vector loads and hundreds of rows are in flight over PCIe;
//   * the miss kernel runs on an internal high-priority stream concurrently with the hit kernel.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pg_common.cuh"

namespace {

using pg::kFullMask;
using pg::smem_u32; using pg::mbar_init; using pg::mbar_expect_tx; using pg::mbar_try_wait;
using pg::bulk_g2s; using pg::bulk_s2g;

constexpr int kRowWarps = 8;          // warps per CTA of the row-copy kernel
constexpr int kBulkWarps = 2;         // warps per CTA of the TMA kernel
constexpr int kSplitThreads = 256;

struct RowsArgs {
  int nfields;
  const float* src[PG_MAX_FIELDS];
  int64_t src_stride[PG_MAX_FIELDS];
  float* dst[PG_MAX_FIELDS];
  int64_t dst_stride[PG_MAX_FIELDS];
  int dim[PG_MAX_FIELDS];
  int bulk_ok[PG_MAX_FIELDS];         // field may use cp.async.bulk (16-byte aligned rows)
  int smem_off[PG_MAX_FIELDS];        // byte offset of the field inside a stage
  int stage_bytes;                    // bytes per stage (bulk fields only)
  int stages;                         // lanes per warp that own a stage
  const int64_t* pos;                 // destination row of item i (null: i)
  const int64_t* row;                 // source row of item i
  const int64_t* count;               // device-resident item count (null: n)
  int64_t n;
  const int64_t* dyn_begin;           // optional device-resident range [*dyn_begin, *dyn_end) into `row` (n is then a capacity)
  const int64_t* dyn_end;
};

__device__ __forceinline__ int64_t rows_extent(RowsArgs& a) {
  if (a.dyn_begin) {
    const int64_t b = *a.dyn_begin;
    a.row += b;
    return min(*a.dyn_end - b, a.n);
  }
  return a.count ? min((int64_t)*a.count, a.n) : a.n;
}

// ------------------------------------------------------------------ split
__global__ void __launch_bounds__(kSplitThreads)
split_kernel(const int64_t* __restrict__ ids, int64_t n, const uint8_t* __restrict__ flag,
             const int64_t* __restrict__ l2c, const int64_t* __restrict__ nid_map, int64_t* hit_pos,
             int64_t* hit_row, int64_t* miss_pos, int64_t* miss_row, unsigned long long* list_counts,
             uint8_t* hit_mask, unsigned long long* user_counts, const int64_t* dyn_begin, const int64_t* dyn_end) {
  __shared__ int warp_hits[kSplitThreads / 32], warp_miss[kSplitThreads / 32];
  __shared__ unsigned long long base_hit, base_miss;
  if (dyn_begin) {  // device-resident range into ids; n is a capacity
    ids += *dyn_begin;
    n = min(n, *dyn_end - *dyn_begin);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  const int64_t ntiles = (n + kSplitThreads - 1) / kSplitThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t j = tile * kSplitThreads + threadIdx.x;
    const bool valid = j < n;
    int64_t t = 0;
    bool hit = false;
    if (valid) {
      t = ids[j];
      hit = flag[t] != 0;
      if (hit_mask) hit_mask[j] = hit;
    }
    const unsigned hb = __ballot_sync(kFullMask, valid && hit), mb = __ballot_sync(kFullMask, valid && !hit);
    if (lane == 0) {
      warp_hits[w] = __popc(hb);
      warp_miss[w] = __popc(mb);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int th = 0, tm = 0;
      for (int i = 0; i < kSplitThreads / 32; ++i) {
        const int a = warp_hits[i], b = warp_miss[i];
        warp_hits[i] = th; warp_miss[i] = tm;
        th += a; tm += b;
      }
      base_hit = th ? atomicAdd(&list_counts[0], (unsigned long long)th) : 0;
      base_miss = tm ? atomicAdd(&list_counts[1], (unsigned long long)tm) : 0;
      if (user_counts && tm) atomicAdd(&user_counts[1], (unsigned long long)tm);
    }
    __syncthreads();
    if (valid) {
      if (hit) {
        const int64_t o = (int64_t)base_hit + warp_hits[w] + __popc(hb & lt);
        hit_pos[o] = j;
        hit_row[o] = l2c[t];
      } else {
        const int64_t o = (int64_t)base_miss + warp_miss[w] + __popc(mb & lt);
        miss_pos[o] = j;
        miss_row[o] = nid_map[t];
      }
    }
    __syncthreads();
  }
  if (user_counts && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&user_counts[0], (unsigned long long)n);
}

// ------------------------------------------------------------------ peer-GPU cache tier (extends storage.py:157-204)
// The ranks of one node shard the hottest rows between them instead of each holding the same ones: with the vertices in
// one agreed caching order (position k = pos[t]), the first c_local rows are replicated in every rank's table, the next
// world * c_shard are dealt round-robin — position c_local + j lives on rank j % world at row c_local + j / world. A row
// that is not in this rank's HBM is read from the owner's HBM over NVLink (the peer tables are CUDA-IPC mappings) before
// the pinned host table over PCIe is considered.
struct PeerTier {
  int world = 0, rank = 0;
  const float* table[PG_MAX_RANKS] = {nullptr};  // every rank's cache table of the field (table[rank] = the local one)
  const int32_t* pos = nullptr;                  // [node_num] caching position of local id t, < 0 = never cached
  int64_t c_local = 0, c_shard = 0;
  unsigned long long* peer_hits = nullptr;       // optional device counter of rows resolved to a peer
};

// ------------------------------------------------------------------ resolve (fused path): node -> row pointer
// rowptr[j] = start of the feature row of NodeFlow node j: inside the HBM cache table when flag[t_j], else inside
// the miss staging buffer at a freshly allocated slot (the slot's host row goes to miss_row[] for the fetch).
__global__ void __launch_bounds__(kSplitThreads)
resolve_kernel(const int64_t* __restrict__ ids, int64_t n, const uint8_t* __restrict__ flag,
               const int64_t* __restrict__ l2c, const int64_t* __restrict__ nid_map, int is_full, const float* cache,
               int64_t cache_stride, float* stage, int64_t stage_stride, int64_t stage_rows, const float* host,
               int64_t host_stride, const float** rowptr, int64_t* miss_row, unsigned long long* list_counts,
               unsigned long long* user_counts, const int64_t* lo, PeerTier peer, const uint8_t* __restrict__ hot) {
  __shared__ int warp_miss[kSplitThreads / 32];
  __shared__ unsigned long long base_miss;
  if (lo) {  // device-resident extents: ids is the NodeFlow-wide node_mapping, n a capacity
    ids += lo[0];
    n = min(n, lo[1] - lo[0]);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  const int64_t ntiles = (n + kSplitThreads - 1) / kSplitThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t j = tile * kSplitThreads + threadIdx.x;
    const bool valid = j < n;
    int64_t t = 0;
    bool hit = true;
    if (valid) {
      t = ids[j];
      hit = is_full || flag[t] != 0;
      if (hit) {
        // bit 0 of the pointer carries the reuse hint of the row to the row-fetching kernel (rows are 16-byte aligned)
        const uintptr_t tag = hot ? (uintptr_t)(hot[t] & 1) : 0;
        rowptr[j] = (const float*)((uintptr_t)(cache + (is_full ? t : l2c[t]) * cache_stride) | tag);
      } else if (peer.world > 1) {  // another rank's shard? (this rank's own shard rows carry flag[t])
        const int64_t k = peer.pos[t] - peer.c_local;
        if (k >= 0 && k < peer.c_shard * peer.world) {
          rowptr[j] = peer.table[k % peer.world] + (peer.c_local + k / peer.world) * cache_stride;
          hit = true;
          if (peer.peer_hits) atomicAdd(peer.peer_hits, 1ull);
        }
      }
    }
    if (is_full) continue;
    const unsigned mb = __ballot_sync(kFullMask, valid && !hit);
    if (lane == 0) warp_miss[w] = __popc(mb);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tm = 0;
      for (int i = 0; i < kSplitThreads / 32; ++i) {
        const int b = warp_miss[i];
        warp_miss[i] = tm;
        tm += b;
      }
      base_miss = tm ? atomicAdd(&list_counts[1], (unsigned long long)tm) : 0;
      if (user_counts && tm) atomicAdd(&user_counts[1], (unsigned long long)tm);
    }
    __syncthreads();
    if (valid && !hit) {
      const int64_t o = (int64_t)base_miss + warp_miss[w] + __popc(mb & lt);
      const int64_t full = nid_map[t];
      if (o < stage_rows) {
        miss_row[o] = full;
        rowptr[j] = stage + o * stage_stride;
      } else {  // staging buffer exhausted: the consumer reads this row from the pinned host table directly
        rowptr[j] = host + full * host_stride;
      }
    }
    __syncthreads();
  }
  if (user_counts && !is_full && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&user_counts[0], (unsigned long long)n);
}

// ------------------------------------------------------------------ row copy with vector loads (hits; generic fallback)
template <typename V>
__device__ __forceinline__ void copy_vec(const V* __restrict__ s, V* __restrict__ d, int nvec, int lane) {
  constexpr int U = 5;  // 600 floats = 150 float4 = 5 per lane: every load issued before the first store
  for (int c = lane; c < nvec; c += 32 * U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c + 32 * u < nvec) v[u] = __ldg(s + c + 32 * u);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c + 32 * u < nvec) d[c + 32 * u] = v[u];
  }
}

__device__ __forceinline__ void copy_row(const float* src, float* dst, int dim, int lane) {
  const uintptr_t a = (uintptr_t)src | (uintptr_t)dst;
  if ((a & 15) == 0 && (dim & 3) == 0) {
    copy_vec((const float4*)src, (float4*)dst, dim >> 2, lane);
  } else if ((a & 7) == 0 && (dim & 1) == 0) {
    copy_vec((const float2*)src, (float2*)dst, dim >> 1, lane);
  } else {
    copy_vec(src, dst, dim, lane);
  }
}

__global__ void __launch_bounds__(kRowWarps * 32) rows_ldg_kernel(RowsArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * kRowWarps;
  const int64_t n = rows_extent(a);
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t r = a.row[i], p = a.pos ? a.pos[i] : i;
#pragma unroll
    for (int f = 0; f < PG_MAX_FIELDS; ++f) {
      if (f >= a.nfields) break;
      copy_row(a.src[f] + r * a.src_stride[f], a.dst[f] + p * a.dst_stride[f], a.dim[f], lane);
    }
  }
}

// ------------------------------------------------------------------ row copy with TMA bulk copies (misses / cache fill)
__global__ void __launch_bounds__(kBulkWarps * 32) rows_bulk_kernel(RowsArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[kBulkWarps * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool owner = lane < a.stages;
  const uint32_t bar = smem_u32(&bars[threadIdx.x]);
  const uint32_t stage = smem_u32(smem) + (uint32_t)((w * a.stages + (owner ? lane : 0)) * a.stage_bytes);
  if (owner) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int64_t n = rows_extent(a);
  const int64_t warp0 = (int64_t)blockIdx.x * kBulkWarps + w, nwarps = (int64_t)gridDim.x * kBulkWarps;
  uint32_t parity = 0;
  for (int64_t base = warp0 * a.stages; base < n; base += nwarps * a.stages) {
    const int64_t i = base + lane;
    if (owner && i < n) {
      const int64_t r = a.row[i], p = a.pos ? a.pos[i] : i;
      mbar_expect_tx(bar, (uint32_t)a.stage_bytes);
#pragma unroll
      for (int f = 0; f < PG_MAX_FIELDS; ++f) {
        if (f >= a.nfields) break;
        if (a.bulk_ok[f]) bulk_g2s(stage + a.smem_off[f], a.src[f] + r * a.src_stride[f], (uint32_t)a.dim[f] * 4u, bar);
      }
      // narrow / unaligned fields (e.g. norm: one float per row) ride along with plain loads
#pragma unroll
      for (int f = 0; f < PG_MAX_FIELDS; ++f) {
        if (f >= a.nfields) break;
        if (!a.bulk_ok[f]) {
          const float* s = a.src[f] + r * a.src_stride[f];
          float* d = a.dst[f] + p * a.dst_stride[f];
          for (int c = 0; c < a.dim[f]; ++c) d[c] = __ldg(s + c);
        }
      }
      while (!mbar_try_wait(bar, parity)) {
      }
      parity ^= 1;
#pragma unroll
      for (int f = 0; f < PG_MAX_FIELDS; ++f) {
        if (f >= a.nfields) break;
        if (a.bulk_ok[f]) bulk_s2g(a.dst[f] + p * a.dst_stride[f], stage + a.smem_off[f], (uint32_t)a.dim[f] * 4u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // stage may be overwritten next round
    }
    __syncwarp();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ small index kernels
__global__ void fill_setup_kernel(const int64_t* __restrict__ nids, int64_t n, const int64_t* __restrict__ nid_map,
                                  int64_t* l2c, uint8_t* flag, int64_t* row_list) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = nids[i];
    if (l2c) l2c[v] = i;
    if (flag) flag[v] = 1;
    row_list[i] = nid_map[v];
  }
}

__global__ void fill_u8_kernel(uint8_t* p, int64_t n, uint8_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

}  // namespace

// ====================================================================== handle
struct pg_cache {
  int dev = 0;
  int64_t node_num = 0;
  int nfields = 0;
  pg_field fields[PG_MAX_FIELDS];
  const float* host_dev[PG_MAX_FIELDS] = {nullptr};  // device-visible alias of the pinned host tables
  float* cache_tables[PG_MAX_FIELDS] = {nullptr};
  int64_t cached_rows = 0;
  bool is_full = false;
  uint8_t* flag = nullptr;
  int64_t* l2c = nullptr;
  const int64_t* nid_map = nullptr;
  // workspace
  int64_t ws_cap = 0;
  int64_t *hit_pos = nullptr, *hit_row = nullptr, *miss_pos = nullptr, *miss_row = nullptr;
  unsigned long long* list_counts = nullptr;  // [2] hits, misses
  cudaStream_t miss_stream = nullptr;
  cudaEvent_t ev_split = nullptr, ev_miss_done = nullptr;
  int max_smem_optin = 0;
  // fused path (pg_cache_aggregate): per-node row pointers + staging rows for misses
  const float** rowptr = nullptr;
  int64_t rowptr_cap = 0;
  float* stage = nullptr;
  int64_t stage_floats = 0;
  // buffers replaced by a larger workspace: kept until destroy, because a captured CUDA graph may still address them
  std::vector<void*> retired;
  PeerTier peer[PG_MAX_FIELDS];   // optional peer-GPU tier per field (pg_cache_set_peers)
  const uint8_t* hot = nullptr;   // optional [node_num] reuse hint (pg_cache_set_hot), caller-owned
};

static pg_status ensure_ws(pg_cache* c, int64_t n) {
  if (n <= c->ws_cap) return PG_OK;
  const int64_t cap = std::max<int64_t>(n + n / 4, 1 << 16);
  for (void* p : {(void*)c->hit_pos, (void*)c->hit_row, (void*)c->miss_pos, (void*)c->miss_row})
    if (p) c->retired.push_back(p);
  c->hit_pos = c->hit_row = c->miss_pos = c->miss_row = nullptr;
  c->ws_cap = 0;
  const size_t b = (size_t)cap * sizeof(int64_t);
  if (cudaMalloc(&c->hit_pos, b) != cudaSuccess || cudaMalloc(&c->hit_row, b) != cudaSuccess ||
      cudaMalloc(&c->miss_pos, b) != cudaSuccess || cudaMalloc(&c->miss_row, b) != cudaSuccess) {
    cudaGetLastError();
    pg::set_error("pg_cache: out of device memory for a %lld-row workspace", (long long)cap);
    return PG_ERR_NOMEM;
  }
  c->ws_cap = cap;
  return PG_OK;
}

// Launch a row-copy over `n` (or *d_count) items. use_bulk: try the TMA path.
static pg_status launch_rows(pg_cache* c, const float* const* src, const int64_t* src_stride, float* const* dst,
                             const int64_t* pos, const int64_t* row, const unsigned long long* d_count, int64_t n,
                             bool use_bulk, cudaStream_t st, int first_field = 0, int nfields = -1,
                             const int64_t* dyn_begin = nullptr, const int64_t* dyn_end = nullptr) {
  RowsArgs a;
  memset(&a, 0, sizeof(a));
  a.dyn_begin = dyn_begin; a.dyn_end = dyn_end;
  if (nfields < 0) nfields = c->nfields;
  a.nfields = nfields;
  a.pos = pos; a.row = row; a.count = (const int64_t*)d_count; a.n = n;
  int stage = 0;
  bool any_bulk = false, wide_unaligned = false;
  for (int f = 0; f < nfields; ++f) {
    const int dimf = c->fields[first_field + f].dim;
    a.src[f] = src[f]; a.src_stride[f] = src_stride[f];
    a.dst[f] = dst[f]; a.dst_stride[f] = dimf; a.dim[f] = dimf;
    const bool ok = (a.dim[f] % 4 == 0) && (a.src_stride[f] % 4 == 0) && (((uintptr_t)a.src[f] | (uintptr_t)a.dst[f]) % 16 == 0);
    a.bulk_ok[f] = ok;
    a.smem_off[f] = stage;
    if (ok) { stage += a.dim[f] * 4; any_bulk = true; }
    else if (a.dim[f] > 16) wide_unaligned = true;
  }
  const int sms = pg::sm_count(c->dev);
  if (use_bulk && any_bulk && !wide_unaligned) {
    const int budget = std::min(c->max_smem_optin - 1024, env_int("PG_BULK_SMEM", 96 * 1024));
    int stages = std::min(32, budget / (kBulkWarps * stage));
    stages = std::min(stages, env_int("PG_BULK_STAGES", 32));
    if (stages >= 1) {
      a.stage_bytes = stage;
      a.stages = stages;
      const size_t smem = (size_t)kBulkWarps * stages * stage;
      PG_CUDA(cudaFuncSetAttribute(rows_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PG_CUDA(cudaFuncSetAttribute(rows_bulk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      const int64_t need = std::max<int64_t>(1, (n + kBulkWarps * stages - 1) / (kBulkWarps * stages));
      const int grid = (int)std::min<int64_t>(need, (int64_t)sms * env_int("PG_BULK_CTAS_PER_SM", 2));
      rows_bulk_kernel<<<grid, kBulkWarps * 32, smem, st>>>(a);
      PG_CHECK_LAUNCH();
      return PG_OK;
    }
  }
  const int64_t need = std::max<int64_t>(1, (n + kRowWarps - 1) / kRowWarps);
  const int grid = (int)std::min<int64_t>(need, (int64_t)sms * 8);
  pg::prefer_max_smem_k(rows_ldg_kernel);
  rows_ldg_kernel<<<grid, kRowWarps * 32, 0, st>>>(a);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

extern "C" {

pg_status pg_cache_create(int64_t node_num, uint8_t* d_flag, int64_t* d_l2c, const int64_t* d_nid_map, int nfields,
                          const pg_field* fields, int dev, pg_cache** out) {
  PG_REQUIRE(out && d_flag && d_l2c && d_nid_map && fields && node_num >= 1, "pg_cache_create: bad arguments");
  PG_REQUIRE(nfields >= 1 && nfields <= PG_MAX_FIELDS, "pg_cache_create: nfields must be in [1, PG_MAX_FIELDS]");
  pg::DeviceGuard guard(dev);
  pg_cache* c = new pg_cache;
  c->dev = dev;
  c->node_num = node_num;
  c->nfields = nfields;
  auto fail = [&](pg_status s) { pg_cache_destroy(c); return s; };
  for (int f = 0; f < nfields; ++f) {
    c->fields[f] = fields[f];
    if (fields[f].dim < 1 || fields[f].host_stride < fields[f].dim || !fields[f].host_table) {
      pg::set_error("pg_cache_create: field %d has a bad dim/stride/table", f);
      return fail(PG_ERR_INVALID);
    }
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, (void*)fields[f].host_table, 0) != cudaSuccess) {
      cudaGetLastError();
      pg::set_error("pg_cache_create: host table of field %d is not page-locked+mapped (use pg_host_alloc / pg_host_register)", f);
      return fail(PG_ERR_INVALID);
    }
    c->host_dev[f] = (const float*)dp;
  }
  c->flag = d_flag;
  c->l2c = d_l2c;
  c->nid_map = d_nid_map;
  if (cudaMalloc(&c->list_counts, 16) != cudaSuccess) {
    cudaGetLastError();
    pg::set_error("pg_cache_create: out of device memory");
    return fail(PG_ERR_NOMEM);
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&c->miss_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_split, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_miss_done, cudaEventDisableTiming) != cudaSuccess) {
    pg::set_error("pg_cache_create: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(PG_ERR_CUDA);
  }
  cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  *out = c;
  return PG_OK;
}

void pg_cache_destroy(pg_cache* c) {
  if (!c) return;
  pg::DeviceGuard guard(c->dev);
  cudaDeviceSynchronize();
  cudaFree(c->list_counts);
  cudaFree(c->hit_pos); cudaFree(c->hit_row); cudaFree(c->miss_pos); cudaFree(c->miss_row);
  cudaFree((void*)c->rowptr); cudaFree(c->stage);
  for (void* p : c->retired) cudaFree(p);
  if (c->miss_stream) cudaStreamDestroy(c->miss_stream);
  for (cudaEvent_t e : {c->ev_split, c->ev_miss_done})
    if (e) cudaEventDestroy(e);
  cudaGetLastError();
  delete c;
}

pg_status pg_cache_fill(pg_cache* c, const int64_t* d_nids, int64_t n, int is_full, float* const* d_cache_tables,
                        int copy_rows, void* stream) {
  PG_REQUIRE(c && d_cache_tables && (d_nids || n == 0) && n >= 0, "pg_cache_fill: bad arguments");
  pg::DeviceGuard guard(c->dev);
  cudaStream_t st = (cudaStream_t)stream;
  for (int f = 0; f < c->nfields; ++f) {
    PG_REQUIRE(d_cache_tables[f] != nullptr || n == 0, "pg_cache_fill: null cache table");
    c->cache_tables[f] = d_cache_tables[f];
  }
  c->cached_rows = n;
  c->is_full = is_full != 0;
  if (n == 0) return PG_OK;
  pg_status s = ensure_ws(c, n);
  if (s != PG_OK) return s;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)pg::sm_count(c->dev) * 8);
  fill_setup_kernel<<<grid, 256, 0, st>>>(d_nids, n, c->nid_map, c->l2c, c->flag, c->miss_row);
  PG_CHECK_LAUNCH();
  if (!copy_rows) return PG_OK;
  int64_t strides[PG_MAX_FIELDS];
  for (int f = 0; f < c->nfields; ++f) strides[f] = c->fields[f].host_stride;
  return launch_rows(c, c->host_dev, strides, c->cache_tables, nullptr, c->miss_row, nullptr, n,
                     env_int("PG_MISS_MODE", 2) == 2, st);
}

pg_status pg_cache_fill_rows(pg_cache* c, const int64_t* d_full_rows, int64_t n, float* const* d_cache_tables, void* stream) {
  PG_REQUIRE(c && d_cache_tables && (d_full_rows || n == 0) && n >= 0, "pg_cache_fill_rows: bad arguments");
  pg::DeviceGuard guard(c->dev);
  for (int f = 0; f < c->nfields; ++f) {
    PG_REQUIRE(d_cache_tables[f] != nullptr || n == 0, "pg_cache_fill_rows: null cache table");
    c->cache_tables[f] = d_cache_tables[f];
  }
  c->cached_rows = n;
  c->is_full = false;
  if (n == 0) return PG_OK;
  int64_t strides[PG_MAX_FIELDS];
  for (int f = 0; f < c->nfields; ++f) strides[f] = c->fields[f].host_stride;
  return launch_rows(c, c->host_dev, strides, c->cache_tables, nullptr, d_full_rows, nullptr, n,
                     env_int("PG_MISS_MODE", 2) == 2, (cudaStream_t)stream);
}

pg_status pg_cache_fetch_host(pg_cache* c, const int64_t* d_nids, int64_t n, float* const* d_out, void* stream) {
  PG_REQUIRE(c && d_out && (d_nids || n == 0) && n >= 0, "pg_cache_fetch_host: bad arguments");
  if (n == 0) return PG_OK;
  pg::DeviceGuard guard(c->dev);
  cudaStream_t st = (cudaStream_t)stream;
  pg_status s = ensure_ws(c, n);
  if (s != PG_OK) return s;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)pg::sm_count(c->dev) * 8);
  fill_setup_kernel<<<grid, 256, 0, st>>>(d_nids, n, c->nid_map, nullptr, nullptr, c->miss_row);
  PG_CHECK_LAUNCH();
  int64_t strides[PG_MAX_FIELDS];
  for (int f = 0; f < c->nfields; ++f) strides[f] = c->fields[f].host_stride;
  return launch_rows(c, c->host_dev, strides, d_out, nullptr, c->miss_row, nullptr, n,
                     env_int("PG_MISS_MODE", 2) == 2, st);
}

static pg_status cache_fetch_impl(pg_cache* c, const int64_t* d_parent_ids, int64_t n, float* const* d_out,
                                  uint8_t* d_hit_mask, int64_t* d_counts, int mode, const int64_t* d_begin,
                                  const int64_t* d_end, int64_t* d_ws, void* stream);

pg_status pg_cache_fetch(pg_cache* c, const int64_t* d_parent_ids, int64_t n, float* const* d_out, uint8_t* d_hit_mask,
                         int64_t* d_counts, int mode, void* stream) {
  return cache_fetch_impl(c, d_parent_ids, n, d_out, d_hit_mask, d_counts, mode, nullptr, nullptr, nullptr, stream);
}

pg_status pg_cache_fetch_dyn(pg_cache* c, const int64_t* d_ids_base, const int64_t* d_begin, const int64_t* d_end,
                             int64_t cap, float* const* d_out, int64_t* d_counts, int mode, int64_t* d_ws, void* stream) {
  PG_REQUIRE(d_begin && d_end, "pg_cache_fetch_dyn: null range");
  return cache_fetch_impl(c, d_ids_base, cap, d_out, nullptr, d_counts, mode, d_begin, d_end, d_ws, stream);
}

static pg_status cache_fetch_impl(pg_cache* c, const int64_t* d_parent_ids, int64_t n, float* const* d_out,
                                  uint8_t* d_hit_mask, int64_t* d_counts, int mode, const int64_t* d_begin,
                                  const int64_t* d_end, int64_t* d_ws, void* stream) {
  PG_REQUIRE(c && d_out && (d_parent_ids || n == 0) && n >= 0, "pg_cache_fetch: bad arguments");
  PG_REQUIRE(mode >= 0 && mode <= 2, "pg_cache_fetch: mode must be 0, 1 or 2");
  if (n == 0) return PG_OK;
  pg::DeviceGuard guard(c->dev);
  cudaStream_t st = (cudaStream_t)stream;
  for (int f = 0; f < c->nfields; ++f) PG_REQUIRE(d_out[f] != nullptr, "pg_cache_fetch: null output table");
  int64_t cache_strides[PG_MAX_FIELDS], host_strides[PG_MAX_FIELDS];
  for (int f = 0; f < c->nfields; ++f) {
    cache_strides[f] = c->fields[f].dim;
    host_strides[f] = c->fields[f].host_stride;
  }
  const int sms = pg::sm_count(c->dev);
  if (c->is_full) {
    // fetch_from_cache (storage.py:207-216): every row is cached in id order, cache row == local id
    if (d_hit_mask) {
      fill_u8_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 8), 256, 0, st>>>(d_hit_mask, n, 1);
      PG_CHECK_LAUNCH();
    }
    pg::TimedScope ts(PG_T_GATHER_HIT, st);
    return launch_rows(c, c->cache_tables, cache_strides, d_out, nullptr, d_parent_ids, nullptr, n, false, st, 0, -1,
                       d_begin, d_end);
  }
  // hit / miss lists and their counters: the caller's workspace (int64[2 + 4 n]: a pipeline replaying captured graphs
  // owns one per call site) or the handle's shared one
  pg_status s = PG_OK;
  int64_t *hit_pos, *hit_row, *miss_pos, *miss_row;
  unsigned long long* list_counts;
  if (d_ws) {
    list_counts = (unsigned long long*)d_ws;
    hit_pos = d_ws + 2; hit_row = hit_pos + n; miss_pos = hit_row + n; miss_row = miss_pos + n;
  } else {
    s = ensure_ws(c, n);
    if (s != PG_OK) return s;
    list_counts = c->list_counts;
    hit_pos = c->hit_pos; hit_row = c->hit_row; miss_pos = c->miss_pos; miss_row = c->miss_row;
  }
  {
    pg::TimedScope ts(PG_T_SPLIT, st);
    PG_CUDA(cudaMemsetAsync(list_counts, 0, 16, st));
    const int grid = (int)std::min<int64_t>((n + kSplitThreads - 1) / kSplitThreads, (int64_t)sms * 8);
    pg::prefer_max_smem_k(split_kernel);
    split_kernel<<<grid, kSplitThreads, 0, st>>>(d_parent_ids, n, c->flag, c->l2c, c->nid_map, hit_pos, hit_row,
                                                 miss_pos, miss_row, list_counts, d_hit_mask,
                                                 (unsigned long long*)d_counts, d_begin, d_end);
    PG_CHECK_LAUNCH();
  }
  // misses first, on the high-priority side stream, so PCIe is busy while the hit rows stream from HBM
  if (mode == 0) mode = env_int("PG_MISS_MODE", 2);
  PG_CUDA(cudaEventRecord(c->ev_split, st));
  PG_CUDA(cudaStreamWaitEvent(c->miss_stream, c->ev_split, 0));
  {
    pg::TimedScope ts(PG_T_GATHER_MISS, c->miss_stream);
    s = launch_rows(c, c->host_dev, host_strides, d_out, miss_pos, miss_row, list_counts + 1, n, mode == 2,
                    c->miss_stream);
  }
  if (s != PG_OK) return s;
  PG_CUDA(cudaEventRecord(c->ev_miss_done, c->miss_stream));
  if (c->cached_rows > 0) {
    pg::TimedScope ts(PG_T_GATHER_HIT, st);
    s = launch_rows(c, c->cache_tables, cache_strides, d_out, hit_pos, hit_row, list_counts, n, false, st);
    if (s != PG_OK) return s;
  }
  PG_CUDA(cudaStreamWaitEvent(st, c->ev_miss_done, 0));
  return PG_OK;
}

static pg_status ensure_fused_ws(pg_cache* c, int64_t n_src, int dim, bool need_stage) {
  if (n_src > c->rowptr_cap) {
    if (c->rowptr) c->retired.push_back((void*)c->rowptr);
    c->rowptr = nullptr;
    c->rowptr_cap = 0;
    const int64_t cap = std::max<int64_t>(n_src + n_src / 4, 1 << 16);
    if (cudaMalloc((void**)&c->rowptr, (size_t)cap * sizeof(float*)) != cudaSuccess) {
      cudaGetLastError();
      pg::set_error("pg_cache_aggregate: out of device memory for %lld row pointers", (long long)cap);
      return PG_ERR_NOMEM;
    }
    c->rowptr_cap = cap;
  }
  if (need_stage) {
    const int64_t need = std::max<int64_t>(n_src, 1) * dim;
    if (need > c->stage_floats) {
      if (c->stage) c->retired.push_back(c->stage);
      c->stage = nullptr;
      c->stage_floats = 0;
      const int64_t cap = need + need / 4;
      if (cudaMalloc(&c->stage, (size_t)cap * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        pg::set_error("pg_cache_aggregate: out of device memory for a %lld-float miss staging buffer", (long long)cap);
        return PG_ERR_NOMEM;
      }
      c->stage_floats = cap;
    }
  }
  return PG_OK;
}

pg_status pg_cache_resolve(pg_cache* c, int field, const pg_block* blk, const float** d_rowptr, float* d_stage,
                           int64_t stage_rows, int64_t* d_counts, int64_t* d_ws, void* stream) {
  PG_REQUIRE(c && blk && d_rowptr, "pg_cache_resolve: bad arguments");
  PG_REQUIRE(field >= 0 && field < c->nfields, "pg_cache_resolve: no such field");
  const int64_t n_src = blk->n_src;
  PG_REQUIRE(n_src >= 0 && (blk->parent_ids || n_src == 0), "pg_cache_resolve: bad block");
  PG_REQUIRE(c->is_full || stage_rows == 0 || d_stage, "pg_cache_resolve: null staging buffer");
  if (n_src == 0) return PG_OK;
  pg::DeviceGuard guard(c->dev);
  cudaStream_t st = (cudaStream_t)stream;
  const int dim = c->fields[field].dim;
  const bool full = c->is_full;
  // miss list + its counters: the caller's workspace (a pipeline that replays captured graphs owns one per slot, so an
  // eager fetch on another stream can neither move nor race it), else the handle's shared one
  unsigned long long* list_counts = c->list_counts;
  int64_t* miss_row = nullptr;
  if (d_ws) {
    list_counts = (unsigned long long*)d_ws;
    miss_row = d_ws + 2;
  } else if (!full) {
    pg_status s = ensure_ws(c, n_src);
    if (s != PG_OK) return s;
    miss_row = c->miss_row;
  }
  {
    pg::TimedScope timed(PG_T_SPLIT, st);
    if (!full) PG_CUDA(cudaMemsetAsync(list_counts, 0, 16, st));
    const int grid = (int)std::min<int64_t>((n_src + kSplitThreads - 1) / kSplitThreads, (int64_t)pg::sm_count(c->dev) * 8);
    pg::prefer_max_smem_k(resolve_kernel);
    resolve_kernel<<<grid, kSplitThreads, 0, st>>>(blk->parent_ids, n_src, c->flag, c->l2c, c->nid_map, full ? 1 : 0,
                                                  c->cache_tables[field], dim, d_stage, dim, stage_rows, c->host_dev[field],
                                                  c->fields[field].host_stride, d_rowptr, miss_row, list_counts,
                                                  (unsigned long long*)d_counts, blk->d_layer_offsets, c->peer[field], c->hot);
    PG_CHECK_LAUNCH();
  }
  if (!full && stage_rows > 0) {  // missed rows: pinned host table -> staging rows, slot order (TMA bulk copies over PCIe)
    pg::TimedScope timed(PG_T_GATHER_MISS, st);
    const float* src[1] = {c->host_dev[field]};
    const int64_t stride[1] = {c->fields[field].host_stride};
    float* dst[1] = {d_stage};
    pg_status s = launch_rows(c, src, stride, dst, nullptr, miss_row, list_counts + 1, std::min(n_src, stage_rows),
                              env_int("PG_MISS_MODE", 2) == 2, st, field, 1);
    if (s != PG_OK) return s;
  }
  return PG_OK;
}

pg_status pg_cache_set_hot(pg_cache* c, const uint8_t* d_hot) {
  PG_REQUIRE(c, "pg_cache_set_hot: null cache");
  c->hot = d_hot;
  return PG_OK;
}

pg_status pg_cache_set_peers(pg_cache* c, int field, int world, int rank, const float* const* d_peer_tables,
                             const int32_t* d_pos, int64_t c_local, int64_t c_shard, int64_t* d_peer_hits) {
  PG_REQUIRE(c && field >= 0 && field < c->nfields, "pg_cache_set_peers: no such field");
  if (world <= 1) {  // tier off
    c->peer[field] = PeerTier();
    return PG_OK;
  }
  PG_REQUIRE(world <= PG_MAX_RANKS && rank >= 0 && rank < world && d_peer_tables && d_pos && c_local >= 0 && c_shard >= 0,
             "pg_cache_set_peers: bad arguments");
  PG_REQUIRE(!c->is_full, "pg_cache_set_peers: the local cache already holds every row");
  PeerTier p;
  p.world = world; p.rank = rank; p.pos = d_pos; p.c_local = c_local; p.c_shard = c_shard;
  p.peer_hits = (unsigned long long*)d_peer_hits;
  for (int r = 0; r < world; ++r) {
    PG_REQUIRE(d_peer_tables[r] != nullptr, "pg_cache_set_peers: null peer table");
    p.table[r] = d_peer_tables[r];
  }
  c->peer[field] = p;
  return PG_OK;
}

pg_status pg_aggregate_rows(const float* const* d_rowptr, const pg_block* blk, int32_t dim, float* d_dst,
                            int64_t dst_stride, int mode, const float* d_norm, float dropout_p, uint64_t dropout_seed,
                            const int64_t* d_step, int64_t zero_rows_to, void* stream) {
  PG_REQUIRE(d_rowptr && blk && d_dst && dim >= 1, "pg_aggregate_rows: bad arguments");
  PG_REQUIRE(mode == PG_AGG_SUM || mode == PG_AGG_MEAN, "pg_aggregate_rows: mode must be PG_AGG_SUM or PG_AGG_MEAN");
  PG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "pg_aggregate_rows: dropout_p must be in [0, 1)");
  PG_REQUIRE(blk->n_dst >= 0 && dst_stride >= dim && (blk->indptr || blk->n_dst == 0), "pg_aggregate_rows: bad sizes");
  PG_REQUIRE(zero_rows_to >= 0 || blk->d_layer_offsets, "pg_aggregate_rows: a rounding zero_rows_to needs device extents");
  if (std::max(blk->n_dst, zero_rows_to) == 0) return PG_OK;
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  pg::AggRowsArgs a;
  a.indptr = blk->indptr; a.cols = blk->cols; a.col_base = blk->col_base; a.rowptr = d_rowptr;
  a.dst = d_dst; a.dst_stride = dst_stride; a.n_dst = blk->n_dst; a.zero_rows_to = zero_rows_to;
  a.dim = dim; a.mode = mode; a.norm = d_norm;
  a.drop_thr = (uint32_t)(dropout_p * 65536.0f + 0.5f);
  a.keep_scale = 1.0f / (1.0f - dropout_p);
  a.drop_seed = dropout_seed; a.drop_step = d_step;
  a.lo = blk->d_layer_offsets;
  a.hints = env_int("PG_AGG_L2HINT", 1);   // 1: hot evict_last / rest evict_first, 2: hot evict_last / rest default, 0: off
  pg::TimedScope timed(PG_T_FUSED, st);
  return pg::launch_agg_rows(a, dev, st);
}

pg_status pg_cache_aggregate(pg_cache* c, int field, const pg_block* blk, float* d_dst, int64_t dst_stride, int mode,
                             const float* d_norm, float dropout_p, uint64_t dropout_seed, const int64_t* d_step,
                             int64_t zero_rows_to, int64_t* d_counts, void* stream) {
  PG_REQUIRE(c && blk && d_dst, "pg_cache_aggregate: bad arguments");
  PG_REQUIRE(field >= 0 && field < c->nfields, "pg_cache_aggregate: no such field");
  PG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "pg_cache_aggregate: dropout_p must be in [0, 1)");
  PG_REQUIRE(mode == PG_AGG_SUM || mode == PG_AGG_MEAN, "pg_cache_aggregate: mode must be PG_AGG_SUM or PG_AGG_MEAN");
  const int dim = c->fields[field].dim;
  PG_REQUIRE(blk->n_src >= 0 && blk->n_dst >= 0 && dst_stride >= dim, "pg_cache_aggregate: bad sizes");
  pg::DeviceGuard guard(c->dev);
  // workspaces grow outside of any stream capture: size them with one eager call first
  pg_status s = ensure_fused_ws(c, blk->n_src, dim, !c->is_full);
  if (s != PG_OK) return s;
  s = pg_cache_resolve(c, field, blk, c->rowptr, c->stage, c->is_full ? 0 : blk->n_src, d_counts, nullptr, stream);
  if (s != PG_OK) return s;
  return pg_aggregate_rows(c->rowptr, blk, dim, d_dst, dst_stride, mode, d_norm, dropout_p, dropout_seed, d_step,
                           zero_rows_to, stream);
}

}  // extern "C"
