// pg_sample.cu — k-hop neighbour sampling + NodeFlow construction on the GPU.
//
// Replaces dgl==0.4.1 SampleSubgraph / GetUniformSample / ConstructNodeFlow (C++/OpenMP, called from
// dgl.contrib.sampling.NeighborSampler; reference call site examples/profile/pa_gcn.py:71-76).
// Semantics follow SURVEY.md Appendix A.3/A.4 and are bit-identical to oracle/pg_oracle.cpp.
//
// B200 design (integer, HBM/latency-bound work — no tensor cores):
//   * the CSR stays resident in HBM; one warp expands one frontier vertex (coalesced take-all rows,
//     warp-parallel "set until k distinct" draws with match/ballot for the sampled rows);
//   * per-layer dedup + "sort by parent id" is a V-bit bitmap (fits L2: 1.25 MB at 10 M vertices):
//     atomicOr marks, a popcount prefix-sum ranks — no sort, no hash table, deterministic;
//   * every count stays on the device (grids are fixed, kernels read sizes from HBM), so a
//     minibatch is sampled without a single host synchronisation; sizes go back through `meta`.
#include <algorithm>
#include <climits>
#include <type_traits>
#include <vector>

#include "pg_common.cuh"

namespace {

using pg::kFullMask;

constexpr int kSeedThreads = 256;    // seed kernel CTA: 256 x 64 registers fits beside the aggregation kernel's CTA
constexpr int kSeedItems = 8;        // frontier vertices per thread per round of its row-offset scan
constexpr int kBitsThreads = 256;    // bitmap -> layer kernel: 4 words per thread
constexpr int kBitsTile = kBitsThreads * 4;
constexpr int kPickWarps = 8;        // warps per CTA in the pick kernel
constexpr int kSmemPicks = 64;       // accepted-position slots per warp kept in shared memory
constexpr int64_t kEmptyKey = -1;

// Device-resident counters of one pg_sample call.
struct Counts {
  int64_t n_layer[PG_MAX_HOPS + 1];  // sampling order: [0] = seeds
  int64_t e_hop[PG_MAX_HOPS + 1];    // [h] = edges sampled when expanding layer h-1
  int64_t overflow;
};

// Work-distribution state of the per-hop bitmap kernels (zeroed by the seed kernel of every call).
struct Control {
  unsigned tile_ctr1[PG_MAX_HOPS + 1];   // next tile of phase 1 (tile sums)
  unsigned tiles_done[PG_MAX_HOPS + 1];  // tiles whose sums are published
  unsigned tile_ctr2[PG_MAX_HOPS + 1];   // next tile of phase 2 (emit)
  unsigned ready[PG_MAX_HOPS + 1];       // tile prefixes are final
  unsigned front_ctr1[PG_MAX_HOPS + 1];  // the same four for front_kernel
  unsigned front_done[PG_MAX_HOPS + 1];
  unsigned front_ctr2[PG_MAX_HOPS + 1];
  unsigned front_ready[PG_MAX_HOPS + 1];
  unsigned seed_ticket;                  // CTAs of the seed kernel that have finished their slice
  unsigned seed_dup;                     // some seed occurred twice
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// ------------------------------------------------------------------ kernel 1 of a call: the seed layer
// Every CTA: its slice of (a) zeroing the per-hop bitmaps of this call (one contiguous range — the only work proportional
// to V, written at HBM/L2 speed) and (b) the duplicate check of the seeds: atomicOr into a V-bit scratch bitmap returns
// the old bit, so a batch without duplicates — every training batch of a duplicate-free train set — is recognised in one
// parallel pass and copied through as the seed layer.
// The CTA that finishes last (ticket) runs the serial tail: for a batch WITH duplicates the exact ordered dedup (first
// occurrence wins, Appendix A.3: open-addressing table keyed by vertex holding the minimum position, then an ordered
// compaction); the scratch bits are cleared again through the seed list (work proportional to the batch); the row
// offsets of hop 1 (exclusive scan of min(in_degree, fanout) over the seed layer).
struct SeedArgs {
  const int64_t* seeds;
  int64_t n;
  const int64_t* indptr;
  int64_t fanout0;
  uint32_t* seedbits;
  int64_t* hash_keys;
  int* hash_minpos;
  uint64_t hash_mask;
  int64_t* layer0;
  int64_t* row_off1;
  Counts* counts;
  Control* ctl;
  uint4* zero_base;
  int64_t zero_vec;   // uint4 count
};

// row_off[i] = exclusive prefix of min(in_degree(front[i]), fanout), i < n; row_off[n] = total. One CTA; kSeedItems
// consecutive vertices per thread so that a round's indptr reads are all in flight together. Returns the total.
__device__ __forceinline__ int64_t cta_row_offsets(const int64_t* __restrict__ indptr, const int64_t* front, int64_t n,
                                                   int64_t fanout, int64_t* row_off, int64_t* sh) {
  const int tid = threadIdx.x;
  int64_t running = 0;
  for (int64_t c0 = 0; c0 < n; c0 += (int64_t)kSeedThreads * kSeedItems) {
    const int64_t i0 = c0 + (int64_t)tid * kSeedItems;
    int64_t v[kSeedItems], c[kSeedItems];
#pragma unroll
    for (int k = 0; k < kSeedItems; ++k) v[k] = (i0 + k < n) ? __ldcg(front + i0 + k) : -1;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kSeedItems; ++k) {
      c[k] = v[k] >= 0 ? min(indptr[v[k] + 1] - indptr[v[k]], fanout) : 0;
      s += c[k];
    }
    int64_t total;
    int64_t excl = pg::block_exclusive_scan(s, total, sh) + running;
#pragma unroll
    for (int k = 0; k < kSeedItems; ++k) {
      if (i0 + k < n) row_off[i0 + k] = excl;
      excl += c[k];
    }
    running += total;
  }
  if (tid == 0) row_off[n] = running;
  return running;
}

__global__ void __launch_bounds__(kSeedThreads) seed_kernel(SeedArgs a) {
  __shared__ int64_t sh[kSeedThreads / 32 + 1];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int64_t gtid = (int64_t)blockIdx.x * kSeedThreads + tid, gth = (int64_t)gridDim.x * kSeedThreads;
  for (int64_t i = gtid; i < a.zero_vec; i += gth) a.zero_base[i] = make_uint4(0, 0, 0, 0);
  // duplicate check + optimistic copy (all CTAs, one seed per thread and round)
  int dup = 0;
  for (int64_t i = gtid; i < a.n; i += gth) {
    const int64_t v = a.seeds[i];
    const uint32_t bit = 1u << (v & 31);
    dup |= (atomicOr(&a.seedbits[v >> 5], bit) & bit) != 0;
    a.layer0[i] = v;
  }
  if (__syncthreads_or(dup) && tid == 0) atomicOr(&a.ctl->seed_dup, 1u);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&a.ctl->seed_ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  // ---- serial tail (one CTA)
  __threadfence();
  const bool any_dup = atomicAdd(&a.ctl->seed_dup, 0u) != 0;
  for (int64_t i = tid; i < a.n; i += kSeedThreads) a.seedbits[a.seeds[i] >> 5] = 0;   // scratch bitmap back to all-zero
  for (int i = tid; i < (int)(sizeof(Counts) / 8); i += kSeedThreads) ((int64_t*)a.counts)[i] = 0;
  for (int i = tid; i < (int)(sizeof(Control) / 4); i += kSeedThreads) ((unsigned*)a.ctl)[i] = 0;   // incl. ticket / dup
  int64_t n0 = a.n;
  if (any_dup) {
    for (uint64_t s = tid; s <= a.hash_mask; s += kSeedThreads) {
      a.hash_keys[s] = kEmptyKey;
      a.hash_minpos[s] = INT_MAX;
    }
    __syncthreads();
    for (int64_t i = tid; i < a.n; i += kSeedThreads) {
      const int64_t v = a.seeds[i];
      uint64_t slot = mix64((uint64_t)v) & a.hash_mask;
      while (true) {
        const int64_t old = (int64_t)atomicCAS((unsigned long long*)&a.hash_keys[slot], (unsigned long long)kEmptyKey,
                                               (unsigned long long)v);
        if (old == kEmptyKey || old == v) {
          atomicMin(&a.hash_minpos[slot], (int)i);
          break;
        }
        slot = (slot + 1) & a.hash_mask;
      }
    }
    __syncthreads();
    n0 = 0;
    for (int64_t c0 = 0; c0 < a.n; c0 += kSeedThreads) {   // ordered compaction of the first occurrences
      const int64_t i = c0 + tid;
      int64_t keep = 0, v = 0;
      if (i < a.n) {
        v = a.seeds[i];
        uint64_t slot = mix64((uint64_t)v) & a.hash_mask;
        while (__ldcg(&a.hash_keys[slot]) != v) slot = (slot + 1) & a.hash_mask;
        keep = __ldcg(&a.hash_minpos[slot]) == (int)i;
      }
      int64_t total;
      const int64_t excl = pg::block_exclusive_scan(keep, total, sh);
      if (keep) a.layer0[n0 + excl] = v;
      n0 += total;
    }
  }
  __syncthreads();
  const int64_t e1 = cta_row_offsets(a.indptr, a.layer0, n0, a.fanout0, a.row_off1, sh);
  if (tid == 0) {
    a.counts->n_layer[0] = n0;
    a.counts->e_hop[1] = e1;
  }
}

// ------------------------------------------------------------------ hops >= 2: row offsets of the frontier
// row_off[h][i] = exclusive prefix of min(in_degree(layer[h-1][i]), fanout): one kernel, tiles of the frontier handed
// out by an atomic counter (work proportional to the frontier and balanced by frontier index, whatever the vertex ids),
// per-tile sums -> the CTA publishing the last one turns them into tile prefixes and raises `ready` -> every CTA
// re-derives its tiles' offsets. Same hand-off as bits_kernel below; no co-residency requirement.
constexpr int kFrontItems = 4;
constexpr int kFrontTile = kSeedThreads * kFrontItems;

struct FrontArgs {
  const int64_t* indptr;
  const int64_t* front;      // layer[h-1]
  const int64_t* n_front;    // device count
  int64_t cap_front;
  int64_t fanout;
  int64_t* row_off;          // [cap_front + 1]
  int64_t* tile_b;           // [tiles]
  int64_t* e_hop;            // &counts->e_hop[h]
  Control* ctl;
  int h;
};

__global__ void __launch_bounds__(kSeedThreads) front_kernel(FrontArgs a) {
  __shared__ int64_t sh[kSeedThreads / 32 + 1];
  __shared__ unsigned s_tile;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int64_t n = min(*a.n_front, a.cap_front);
  const int64_t ntiles = max((n + kFrontTile - 1) / kFrontTile, (int64_t)1);
  for (int phase = 0; phase < 2; ++phase) {
    unsigned* ctr = phase == 0 ? &a.ctl->front_ctr1[a.h] : &a.ctl->front_ctr2[a.h];
    while (true) {
      if (tid == 0) s_tile = atomicAdd(ctr, 1u);
      __syncthreads();
      const int64_t tile = s_tile;
      if (tile >= ntiles) break;
      const int64_t i0 = tile * kFrontTile + (int64_t)tid * kFrontItems;
      int64_t v[kFrontItems], c[kFrontItems];
#pragma unroll
      for (int k = 0; k < kFrontItems; ++k) v[k] = (i0 + k < n) ? __ldcg(a.front + i0 + k) : -1;
      int64_t s = 0;
#pragma unroll
      for (int k = 0; k < kFrontItems; ++k) {
        c[k] = v[k] >= 0 ? min(a.indptr[v[k] + 1] - a.indptr[v[k]], a.fanout) : 0;
        s += c[k];
      }
      int64_t total;
      int64_t excl = pg::block_exclusive_scan(s, total, sh);
      if (phase == 0) {
        if (tid == 0) {
          a.tile_b[tile] = total;
          __threadfence();
          s_last = atomicAdd(&a.ctl->front_done[a.h], 1u) == (unsigned)(ntiles - 1);
        }
        __syncthreads();
        if (s_last) {
          __threadfence();
          const int64_t chunk = (ntiles + kSeedThreads - 1) / kSeedThreads;
          const int64_t lo = min((int64_t)tid * chunk, ntiles), hi = min(lo + chunk, ntiles);
          int64_t lb = 0;
          for (int64_t i = lo; i < hi; ++i) lb += __ldcg(a.tile_b + i);
          int64_t tot;
          int64_t run = pg::block_exclusive_scan(lb, tot, sh);
          for (int64_t i = lo; i < hi; ++i) {
            const int64_t x = __ldcg(a.tile_b + i);
            a.tile_b[i] = run;
            run += x;
          }
          if (tid == 0) {
            *a.e_hop = tot;
            a.row_off[n] = tot;
          }
          __threadfence();
          __syncthreads();
          if (tid == 0) atomicExch(&a.ctl->front_ready[a.h], 1u);
        }
      } else {
        excl += __ldcg(a.tile_b + tile);
#pragma unroll
        for (int k = 0; k < kFrontItems; ++k) {
          if (i0 + k < n) a.row_off[i0 + k] = excl;
          excl += c[k];
        }
      }
      __syncthreads();
    }
    if (phase == 0) {
      if (tid == 0) {
        while (atomicAdd(&a.ctl->front_ready[a.h], 0u) == 0u) __nanosleep(64);
        __threadfence();
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ per-hop: pick neighbours (one warp per frontier vertex)
struct PickArgs {
  const int64_t* indptr;
  const int64_t* indices;
  const int64_t* eids;       // may be null
  const int64_t* front;
  const int64_t* n_front;
  int64_t cap_front;
  const int64_t* row_off;
  int64_t fanout;
  uint32_t hop;
  uint32_t k0, k1;
  const uint32_t* key_ptr;   // optional device-resident (k0, k1) — CUDA-graph replays re-key without re-capturing
  int64_t* nb_src;           // [cap_edges]
  int64_t* nb_eid;           // [cap_edges]
  int64_t cap_edges;
  uint32_t* bitmap;
  uint32_t* scratch;         // per-warp accepted lists when m > kSmemPicks
  int64_t scratch_stride;
};

__device__ __forceinline__ void emit_edge(const PickArgs& a, int64_t out, int64_t s, int64_t p) {
  const int64_t u = a.indices[s + p];
  atomicOr(&a.bitmap[u >> 5], 1u << (u & 31));
  if (out < a.cap_edges) {
    a.nb_src[out] = u;
    a.nb_eid[out] = a.eids ? a.eids[s + p] : s + p;
  }
}

__global__ void __launch_bounds__(kPickWarps * 32) pick_kernel(PickArgs a) {
  __shared__ uint32_t sh_acc[kPickWarps][kSmemPicks];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kPickWarps + w, nwarps = (int64_t)gridDim.x * kPickWarps;
  const int64_t n = min(*a.n_front, a.cap_front);
  const unsigned lt_mask = (1u << lane) - 1;
  if (a.key_ptr) {
    a.k0 = a.key_ptr[0];
    a.k1 = a.key_ptr[1];
  }
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t v = a.front[i];
    const int64_t s = a.indptr[v], deg = a.indptr[v + 1] - s, k = a.fanout;
    const int64_t off = a.row_off[i];
    if (deg <= k) {  // take every in-neighbour, row order
      for (int64_t p = lane; p < deg; p += 32) emit_edge(a, off + p, s, p);
      continue;
    }
    const bool complement = deg <= 2 * k;
    const int64_t m = complement ? deg - k : k;
    uint32_t* acc = (m <= kSmemPicks) ? sh_acc[w] : a.scratch + warp0 * a.scratch_stride;
    // draws t = 0,1,2,... inserted until m distinct positions (32 draws per round, order preserved)
    int64_t cnt = 0;
    for (uint32_t t0 = 0; cnt < m; t0 += 32) {
      const uint32_t p = (uint32_t)pg::draw_pos(a.k0, a.k1, v, a.hop, t0 + lane, (uint64_t)deg);
      bool dup = false;
      for (int64_t j = 0; j < cnt; ++j) dup |= (acc[j] == p);
      const unsigned peers = __match_any_sync(kFullMask, p);
      const bool is_new = !dup && ((__ffs(peers) - 1) == lane);
      const unsigned new_mask = __ballot_sync(kFullMask, is_new);
      const int64_t slot = cnt + __popc(new_mask & lt_mask);
      if (is_new && slot < m) acc[slot] = p;
      cnt = min(m, cnt + (int64_t)__popc(new_mask));
      __syncwarp();
    }
    if (!complement) {
      // ascending order by rank counting (positions are distinct)
      for (int64_t j = lane; j < m; j += 32) {
        const uint32_t p = acc[j];
        int64_t r = 0;
        for (int64_t q = 0; q < m; ++q) r += (acc[q] < p);
        emit_edge(a, off + r, s, (int64_t)p);
      }
    } else {
      // acc holds the excluded positions; keep the complement, ascending
      int64_t written = 0;
      for (int64_t base = 0; base < deg; base += 32) {
        const int64_t p = base + lane;
        bool keep = p < deg;
        if (keep)
          for (int64_t q = 0; q < m; ++q) keep &= (acc[q] != (uint32_t)p);
        const unsigned km = __ballot_sync(kFullMask, keep);
        if (keep) emit_edge(a, off + written + __popc(km & lt_mask), s, p);
        written += __popc(km);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ per-hop: bitmap -> sorted unique layer + ranks
// One kernel per hop. The vertices picked by hop h are the set bits of bitmap[h]; ascending bit order IS the layer's
// order (sorted by parent id, Appendix A.4), so "dedup + sort" is a popcount prefix sum:
//   phase 1  tiles of kBitsTile words, handed out by an atomic counter: per-tile vertex counts; the CTA that publishes
//            the last tile turns the sums into exclusive tile prefixes and raises `ready`;
//   phase 2  tiles handed out again: word_prefix[w] (rank of the word's first vertex) and layer[h][rank] = vertex.
// CTAs that find no tile left just wait for `ready`; nothing a running CTA waits for depends on a CTA that is not
// running yet, so the kernel needs no co-residency guarantee (and no cooperative launch).
struct BitsArgs {
  const uint32_t* bitmap;     // padded to a multiple of 4 words, padding zero
  int64_t nwords;
  uint32_t* word_prefix;
  int64_t* layer;             // [cap]
  int64_t cap;
  int64_t* tile_a;            // [ntiles] vertex counts -> exclusive prefixes
  int64_t* n_layer;           // &counts->n_layer[h]
  Control* ctl;
  int h;
};

__global__ void __launch_bounds__(kBitsThreads) bits_kernel(BitsArgs a) {
  __shared__ int64_t sh[kBitsThreads / 32 + 1];
  __shared__ unsigned s_tile;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int64_t ntiles = (a.nwords + kBitsTile - 1) / kBitsTile;
  for (int phase = 0; phase < 2; ++phase) {
    unsigned* ctr = phase == 0 ? &a.ctl->tile_ctr1[a.h] : &a.ctl->tile_ctr2[a.h];
    while (true) {
      if (tid == 0) s_tile = atomicAdd(ctr, 1u);
      __syncthreads();
      const int64_t tile = s_tile;
      if (tile >= ntiles) break;
      const int64_t w0 = tile * kBitsTile + (int64_t)tid * 4;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (w0 < a.nwords) q = __ldcg((const uint4*)(a.bitmap + w0));
      const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
      const int64_t na = __popc(wd[0]) + __popc(wd[1]) + __popc(wd[2]) + __popc(wd[3]);
      int64_t ta;
      int64_t ea = pg::block_exclusive_scan(na, ta, sh);
      if (phase == 0) {
        if (tid == 0) {
          a.tile_a[tile] = ta;
          __threadfence();
          s_last = atomicAdd(&a.ctl->tiles_done[a.h], 1u) == (unsigned)(ntiles - 1);
        }
        __syncthreads();
        if (s_last) {  // every tile sum is published: exclusive prefixes over the tiles, total, release
          __threadfence();
          const int64_t chunk = (ntiles + kBitsThreads - 1) / kBitsThreads;
          const int64_t lo = min((int64_t)tid * chunk, ntiles), hi = min(lo + chunk, ntiles);
          int64_t la = 0;
          for (int64_t i = lo; i < hi; ++i) la += __ldcg(a.tile_a + i);
          int64_t tot;
          int64_t run = pg::block_exclusive_scan(la, tot, sh);
          for (int64_t i = lo; i < hi; ++i) {
            const int64_t x = __ldcg(a.tile_a + i);
            a.tile_a[i] = run;
            run += x;
          }
          if (tid == 0) *a.n_layer = tot;
          __threadfence();
          __syncthreads();
          if (tid == 0) atomicExch(&a.ctl->ready[a.h], 1u);
        }
      } else {
        ea += __ldcg(a.tile_a + tile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (w0 + k < a.nwords) a.word_prefix[w0 + k] = (uint32_t)ea;
          uint32_t bits = wd[k];
          while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (ea < a.cap) a.layer[ea] = (w0 + k) * 32 + b;
            ++ea;
          }
        }
      }
      __syncthreads();
    }
    if (phase == 0) {
      if (tid == 0) {
        while (atomicAdd(&a.ctl->ready[a.h], 0u) == 0u) __nanosleep(64);
        __threadfence();
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ assemble the NodeFlow arrays (Appendix A.4)
struct AssembleArgs {
  int L;
  const Counts* counts;
  const int64_t* layer[PG_MAX_HOPS + 1];
  const int64_t* nb_src[PG_MAX_HOPS + 1];
  const int64_t* nb_eid[PG_MAX_HOPS + 1];
  const int64_t* row_off[PG_MAX_HOPS + 1];
  const uint32_t* bitmap[PG_MAX_HOPS + 1];       // [hop]: the vertices hop picked
  const uint32_t* word_prefix[PG_MAX_HOPS + 1];  // [hop]: vertices of that layer below each bitmap word
  int64_t cap_layer[PG_MAX_HOPS + 1];
  int64_t cap_nodes, cap_edges;
  pg_nodeflow_buffers out;
  const int64_t* labels;   // optional: label table indexed by parent id
  int64_t* seed_labels;    // optional: [n_layer[0]] labels of the seed layer, in its (deduplicated) order
};

__global__ void assemble_kernel(AssembleArgs a) {
  __shared__ int64_t lay_off[PG_MAX_HOPS + 2], flow_off[PG_MAX_HOPS + 1];
  __shared__ bool overflow;
  const int L = a.L;
  if (threadIdx.x == 0) {
    bool ovf = false;
    lay_off[0] = 0;
    for (int j = 0; j <= L; ++j) {  // NodeFlow layer j = sampling layer L-j
      const int64_t nl = a.counts->n_layer[L - j];
      ovf |= nl > a.cap_layer[L - j];
      lay_off[j + 1] = lay_off[j] + nl;
    }
    flow_off[0] = 0;
    for (int j = 1; j <= L; ++j) flow_off[j] = flow_off[j - 1] + a.counts->e_hop[L - j + 1];
    ovf |= lay_off[L + 1] > a.cap_nodes || flow_off[L] > a.cap_edges;
    overflow = ovf;
    if (blockIdx.x == 0) {
      int64_t* m = a.out.meta;
      m[0] = ovf ? PG_ERR_OVERFLOW : PG_OK;
      m[1] = lay_off[L + 1];
      m[2] = flow_off[L];
      m[3] = L + 1;
      for (int j = 0; j <= L + 1; ++j) m[4 + j] = lay_off[j];
      for (int j = 0; j <= L; ++j) m[4 + L + 2 + j] = flow_off[j];
    }
  }
  __syncthreads();
  if (overflow) return;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int j = 0; j <= L; ++j) {
    const int h = L - j;  // sampling layer
    const int64_t base = lay_off[j], nl = lay_off[j + 1] - base;
    const int64_t* lay = a.layer[h];
    for (int64_t i = tid; i < nl; i += nth) a.out.node_mapping[base + i] = lay[i];
    if (j == L && a.seed_labels)
      for (int64_t i = tid; i < nl; i += nth) a.seed_labels[i] = a.labels[lay[i]];
    if (j == 0) {
      for (int64_t i = tid; i <= nl; i += nth) a.out.indptr[i] = 0;
    } else {
      const int hop = h + 1;  // expansion of sampling layer h produced NodeFlow block j-1
      const int64_t eb = flow_off[j - 1], ne = flow_off[j] - eb, col_base = lay_off[j - 1];
      const int64_t* ro = a.row_off[hop];
      for (int64_t i = tid; i < nl; i += nth) a.out.indptr[base + i + 1] = eb + ro[i + 1];
      const int64_t* src = a.nb_src[hop];
      const int64_t* eid = a.nb_eid[hop];
      const uint32_t* __restrict__ bm = a.bitmap[hop];
      const uint32_t* __restrict__ wp = a.word_prefix[hop];
      for (int64_t e = tid; e < ne; e += nth) {
        // NodeFlow id of the source = its rank in the (sorted, unique) layer the hop produced
        const int64_t u = src[e], w = u >> 5;
        a.out.indices[eb + e] = col_base + (int64_t)wp[w] + __popc(bm[w] & ((1u << (u & 31)) - 1));
        a.out.edge_mapping[eb + e] = eid[e];
      }
    }
  }
}

__global__ void degree_kernel(const int64_t* __restrict__ indptr, int64_t n, int64_t* out, unsigned long long* max_out) {
  unsigned long long mx = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = indptr[i + 1] - indptr[i];
    if (out) out[i] = d;
    mx = max(mx, (unsigned long long)d);
  }
  if (max_out) {
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(kFullMask, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_out, mx);
  }
}

__global__ void column_count_kernel(const int64_t* __restrict__ indices, int64_t nnz, int64_t* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    atomicAdd((unsigned long long*)&out[indices[e]], 1ull);
}

int grid_for(int64_t items, int per_block, int dev, int waves = 8) {
  const int64_t need = std::max<int64_t>(1, (items + per_block - 1) / per_block);
  return (int)std::min<int64_t>(need, (int64_t)pg::sm_count(dev) * waves);
}

}  // namespace

// ====================================================================== handles
struct pg_graph {
  int dev = 0;
  int64_t num_nodes = 0, num_edges = 0, max_in_degree = 0;
  const int64_t* indptr = nullptr;
  const int64_t* indices = nullptr;
  const int64_t* eids = nullptr;
  bool owned = false;
};

struct pg_sampler {
  pg_graph* g = nullptr;
  int L = 0;
  int64_t fanouts[PG_MAX_HOPS] = {0};
  uint64_t seed = 0;
  int64_t max_seeds = 0, cap_nodes = 0, cap_edges = 0;
  int64_t nwords = 0;
  int64_t nwords_pad = 0;   // nwords rounded up to a multiple of 4 (uint4 loads / stores)
  // workspace (device)
  uint32_t* seedbits = nullptr;      // V-bit scratch of the seed dedup; all-zero between calls
  uint32_t* bitmaps = nullptr;       // [L][nwords_pad]: vertices picked by hop h+1 (zeroed by every call's seed kernel)
  uint32_t* word_prefix = nullptr;   // [L][nwords_pad]
  int64_t* tile_a = nullptr;
  int64_t* tile_b = nullptr;
  Control* ctl = nullptr;
  Counts* counts = nullptr;
  int64_t* hash_keys = nullptr;
  int* hash_minpos = nullptr;
  uint64_t hash_mask = 0;
  int64_t* layer[PG_MAX_HOPS + 1] = {nullptr};
  int64_t cap_layer[PG_MAX_HOPS + 1] = {0};
  int64_t* nb_src[PG_MAX_HOPS + 1] = {nullptr};
  int64_t* nb_eid[PG_MAX_HOPS + 1] = {nullptr};
  int64_t* row_off[PG_MAX_HOPS + 1] = {nullptr};
  uint32_t* scratch = nullptr;
  int64_t scratch_stride = 0;
  int pick_grid = 0;
  std::vector<void*> allocs;
};

// Host copy of the minibatch-key derivation (oracle/pg_oracle.cpp minibatch_key).
static void minibatch_key(uint64_t seed, int64_t epoch, int64_t batch, uint32_t* k0, uint32_t* k1) {
  const pg::Philox4 r = pg::philox4x32_10((uint32_t)(uint64_t)epoch, (uint32_t)((uint64_t)epoch >> 32),
                                          (uint32_t)(uint64_t)batch, (uint32_t)((uint64_t)batch >> 32),
                                          (uint32_t)seed, (uint32_t)(seed >> 32));
  *k0 = r.c[0];
  *k1 = r.c[1];
}

static pg_status graph_finish(pg_graph* g) {
  unsigned long long* d_max = nullptr;
  PG_CUDA(cudaMalloc(&d_max, sizeof(*d_max)));
  PG_CUDA(cudaMemset(d_max, 0, sizeof(*d_max)));
  if (g->num_nodes > 0) {
    degree_kernel<<<grid_for(g->num_nodes, 256, g->dev), 256>>>(g->indptr, g->num_nodes, nullptr, d_max);
    PG_CHECK_LAUNCH();
  }
  unsigned long long mx = 0;
  PG_CUDA(cudaMemcpy(&mx, d_max, sizeof(mx), cudaMemcpyDeviceToHost));
  cudaFree(d_max);
  g->max_in_degree = (int64_t)mx;
  PG_REQUIRE(mx < (1ull << 32), "in-degree >= 2^32 is not supported");
  return PG_OK;
}

extern "C" {

pg_status pg_graph_create(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t num_nodes,
                          int64_t num_edges, int dev, pg_graph** out) {
  PG_REQUIRE(out && indptr && (indices || num_edges == 0) && num_nodes >= 0 && num_edges >= 0,
             "pg_graph_create: bad arguments");
  PG_REQUIRE(indptr[0] == 0 && indptr[num_nodes] == num_edges, "pg_graph_create: indptr does not span [0, num_edges]");
  pg::DeviceGuard guard(dev);
  pg_graph* g = new pg_graph;
  g->dev = dev;
  g->num_nodes = num_nodes;
  g->num_edges = num_edges;
  g->owned = true;
  int64_t *d_ip = nullptr, *d_ix = nullptr, *d_e = nullptr;
  const size_t eb = (size_t)std::max<int64_t>(num_edges, 1) * sizeof(int64_t);
  if (cudaMalloc(&d_ip, (size_t)(num_nodes + 1) * sizeof(int64_t)) != cudaSuccess ||
      cudaMalloc(&d_ix, eb) != cudaSuccess || (eids && cudaMalloc(&d_e, eb) != cudaSuccess)) {
    cudaFree(d_ip); cudaFree(d_ix); cudaFree(d_e);
    delete g;
    pg::set_error("pg_graph_create: out of device memory (%lld nodes, %lld edges)", (long long)num_nodes,
                  (long long)num_edges);
    cudaGetLastError();
    return PG_ERR_NOMEM;
  }
  g->indptr = d_ip; g->indices = d_ix; g->eids = d_e;
  pg_status st = PG_OK;
  auto fail = [&](pg_status s) { pg_graph_destroy(g); return s; };
  if (cudaMemcpy(d_ip, indptr, (size_t)(num_nodes + 1) * sizeof(int64_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      (num_edges && cudaMemcpy(d_ix, indices, (size_t)num_edges * sizeof(int64_t), cudaMemcpyHostToDevice) != cudaSuccess) ||
      (num_edges && eids && cudaMemcpy(d_e, eids, (size_t)num_edges * sizeof(int64_t), cudaMemcpyHostToDevice) != cudaSuccess)) {
    pg::set_error("pg_graph_create: host->device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(PG_ERR_CUDA);
  }
  if ((st = graph_finish(g)) != PG_OK) return fail(st);
  *out = g;
  return PG_OK;
}

pg_status pg_graph_create_device(const int64_t* d_indptr, const int64_t* d_indices, const int64_t* d_eids,
                                 int64_t num_nodes, int64_t num_edges, int dev, pg_graph** out) {
  PG_REQUIRE(out && d_indptr && (d_indices || num_edges == 0) && num_nodes >= 0 && num_edges >= 0,
             "pg_graph_create_device: bad arguments");
  pg::DeviceGuard guard(dev);
  pg_graph* g = new pg_graph;
  g->dev = dev;
  g->num_nodes = num_nodes;
  g->num_edges = num_edges;
  g->indptr = d_indptr; g->indices = d_indices; g->eids = d_eids;
  g->owned = false;
  pg_status st = graph_finish(g);
  if (st != PG_OK) { delete g; return st; }
  *out = g;
  return PG_OK;
}

void pg_graph_destroy(pg_graph* g) {
  if (!g) return;
  if (g->owned) {
    pg::DeviceGuard guard(g->dev);
    cudaFree((void*)g->indptr);
    cudaFree((void*)g->indices);
    cudaFree((void*)g->eids);
  }
  delete g;
}

pg_status pg_graph_degrees(pg_graph* g, int in_edges, int64_t* d_out, void* stream) {
  PG_REQUIRE(g && d_out, "pg_graph_degrees: bad arguments");
  pg::DeviceGuard guard(g->dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (g->num_nodes == 0) return PG_OK;
  if (in_edges) {
    degree_kernel<<<grid_for(g->num_nodes, 256, g->dev), 256, 0, st>>>(g->indptr, g->num_nodes, d_out, nullptr);
    PG_CHECK_LAUNCH();
  } else {
    PG_CUDA(cudaMemsetAsync(d_out, 0, (size_t)g->num_nodes * sizeof(int64_t), st));
    if (g->num_edges) {
      column_count_kernel<<<grid_for(g->num_edges, 256, g->dev), 256, 0, st>>>(g->indices, g->num_edges, d_out);
      PG_CHECK_LAUNCH();
    }
  }
  return PG_OK;
}

pg_status pg_sampler_create(pg_graph* g, int num_hops, const int64_t* fanouts, uint64_t seed, int64_t max_seeds,
                            int64_t cap_nodes, int64_t cap_edges, pg_sampler** out) {
  PG_REQUIRE(g && out && fanouts, "pg_sampler_create: bad arguments");
  PG_REQUIRE(num_hops >= 1 && num_hops <= PG_MAX_HOPS, "pg_sampler_create: num_hops must be in [1, PG_MAX_HOPS]");
  PG_REQUIRE(max_seeds >= 1 && max_seeds < INT_MAX && cap_nodes >= max_seeds && cap_edges >= 1,
             "pg_sampler_create: bad capacities");
  PG_REQUIRE(g->num_nodes < (1ll << 32), "pg_sampler_create: more than 2^32 vertices per partition are not supported");
  for (int h = 0; h < num_hops; ++h) PG_REQUIRE(fanouts[h] >= 1, "pg_sampler_create: fanout must be >= 1");
  pg::DeviceGuard guard(g->dev);
  pg_sampler* s = new pg_sampler;
  s->g = g;
  s->L = num_hops;
  s->seed = seed;
  s->max_seeds = max_seeds;
  s->cap_nodes = cap_nodes;
  s->cap_edges = cap_edges;
  s->nwords = std::max<int64_t>(1, (g->num_nodes + 31) / 32);
  s->nwords_pad = (s->nwords + 3) / 4 * 4;
  int64_t max_m = 0;
  for (int h = 0; h < num_hops; ++h) {
    s->fanouts[h] = fanouts[h];
    if (fanouts[h] < g->max_in_degree) max_m = std::max(max_m, std::min(fanouts[h], g->max_in_degree - fanouts[h]));
  }
  s->pick_grid = pg::sm_count(g->dev) * 8;
  bool ok = true;
  auto alloc = [&](auto** p, size_t count) {
    using T = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
    if (!ok) return;
    void* q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) { ok = false; cudaGetLastError(); return; }
    s->allocs.push_back(q);
    *p = (T*)q;
  };
  alloc(&s->seedbits, (size_t)s->nwords_pad);
  alloc(&s->bitmaps, (size_t)s->nwords_pad * num_hops);
  alloc(&s->word_prefix, (size_t)s->nwords_pad * num_hops);
  const int64_t ntiles = (s->nwords + kBitsTile - 1) / kBitsTile;
  alloc(&s->tile_a, (size_t)ntiles);
  alloc(&s->tile_b, (size_t)((cap_nodes + kFrontTile - 1) / kFrontTile + 1));
  alloc(&s->ctl, 1);
  alloc(&s->counts, 1);
  uint64_t tsize = 64;
  while (tsize < (uint64_t)max_seeds * 2) tsize <<= 1;
  s->hash_mask = tsize - 1;
  alloc(&s->hash_keys, tsize);
  alloc(&s->hash_minpos, tsize);
  for (int h = 0; h <= num_hops; ++h) {
    s->cap_layer[h] = (h == 0) ? max_seeds : cap_nodes;
    alloc(&s->layer[h], (size_t)s->cap_layer[h]);
    if (h >= 1) {
      alloc(&s->nb_src[h], (size_t)cap_edges);
      alloc(&s->nb_eid[h], (size_t)cap_edges);
      alloc(&s->row_off[h], (size_t)s->cap_layer[h - 1] + 1);
    }
  }
  if (max_m > kSmemPicks) {
    s->scratch_stride = max_m;
    alloc(&s->scratch, (size_t)s->pick_grid * kPickWarps * (size_t)max_m);
  }
  if (!ok) {
    pg_sampler_destroy(s);
    pg::set_error("pg_sampler_create: out of device memory (cap_nodes=%lld cap_edges=%lld)", (long long)cap_nodes,
                  (long long)cap_edges);
    return PG_ERR_NOMEM;
  }
  PG_CUDA(cudaMemset(s->seedbits, 0, (size_t)s->nwords_pad * sizeof(uint32_t)));
  PG_CUDA(cudaMemset(s->ctl, 0, sizeof(Control)));   // every call's seed kernel leaves it zeroed for the next one
  *out = s;
  return PG_OK;
}

void pg_sampler_destroy(pg_sampler* s) {
  if (!s) return;
  pg::DeviceGuard guard(s->g->dev);
  for (void* p : s->allocs) cudaFree(p);
  delete s;
}

static pg_status sample_impl(pg_sampler* s, const int64_t* d_seeds, int64_t n_seeds, uint32_t k0, uint32_t k1,
                             const uint32_t* d_key, const pg_nodeflow_buffers* out, int64_t* h_meta, const int64_t* d_labels,
                             int64_t* d_seed_labels, void* stream);

void pg_minibatch_key(uint64_t seed, int64_t epoch, int64_t batch, uint32_t* key) {
  minibatch_key(seed, epoch, batch, &key[0], &key[1]);
}

pg_status pg_sample(pg_sampler* s, const int64_t* d_seeds, int64_t n_seeds, int64_t epoch, int64_t batch,
                    const pg_nodeflow_buffers* out, int64_t* h_meta, void* stream) {
  PG_REQUIRE(s != nullptr, "pg_sample: null sampler");
  uint32_t k0, k1;
  minibatch_key(s->seed, epoch, batch, &k0, &k1);
  return sample_impl(s, d_seeds, n_seeds, k0, k1, nullptr, out, h_meta, nullptr, nullptr, stream);
}

pg_status pg_sample_keyed(pg_sampler* s, const int64_t* d_seeds, int64_t n_seeds, const uint32_t* d_key,
                          const pg_nodeflow_buffers* out, int64_t* h_meta, const int64_t* d_labels, int64_t* d_seed_labels,
                          void* stream) {
  PG_REQUIRE(s != nullptr && d_key != nullptr, "pg_sample_keyed: null sampler or key");
  PG_REQUIRE((d_labels == nullptr) == (d_seed_labels == nullptr), "pg_sample_keyed: labels in and out go together");
  return sample_impl(s, d_seeds, n_seeds, 0, 0, d_key, out, h_meta, d_labels, d_seed_labels, stream);
}

static pg_status sample_impl(pg_sampler* s, const int64_t* d_seeds, int64_t n_seeds, uint32_t k0, uint32_t k1,
                             const uint32_t* d_key, const pg_nodeflow_buffers* out, int64_t* h_meta, const int64_t* d_labels,
                             int64_t* d_seed_labels, void* stream) {
  PG_REQUIRE(s && out && out->node_mapping && out->indptr && out->indices && out->edge_mapping && out->meta,
             "pg_sample: null output buffer");
  PG_REQUIRE(n_seeds >= 0 && n_seeds <= s->max_seeds, "pg_sample: n_seeds exceeds max_seeds");
  PG_REQUIRE(d_seeds || n_seeds == 0, "pg_sample: null seeds");
  pg_graph* g = s->g;
  pg::DeviceGuard guard(g->dev);
  cudaStream_t st = (cudaStream_t)stream;
  const int dev = g->dev;

  pg::TimedScope timed(PG_T_SAMPLE, st);
  // measurement aid (tools/engine_breakdown.py interference runs): launch only the first PG_SAMPLE_KERNELS kernels of the
  // chain — the outputs are then incomplete, never set outside a timing experiment
  const char* dbg_env = getenv("PG_SAMPLE_KERNELS");
  const int dbg_limit = dbg_env ? atoi(dbg_env) : 1 << 30;
  int dbg_n = 0;
  // kernels of one call (no memset nodes, no host synchronisation):
  //   seed_kernel | hop 1: pick_kernel, bits_kernel | hop h >= 2: front_kernel, pick_kernel, bits_kernel | assemble_kernel
  {
    SeedArgs sa{d_seeds, n_seeds, g->indptr, s->fanouts[0], s->seedbits, s->hash_keys, s->hash_minpos, s->hash_mask,
                s->layer[0], s->row_off[1], s->counts, s->ctl, (uint4*)s->bitmaps, s->nwords_pad * s->L / 4};
    const int64_t want = std::max((sa.zero_vec + kSeedThreads * 4 - 1) / (kSeedThreads * 4), (n_seeds + kSeedThreads - 1) / kSeedThreads);
    const int grid_s = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)pg::sm_count(dev) * 2));
    pg::prefer_max_smem_k(seed_kernel);
    if (dbg_n++ < dbg_limit) seed_kernel<<<grid_s, kSeedThreads, 0, st>>>(sa);
    PG_CHECK_LAUNCH();
  }
  const int64_t ntiles = (s->nwords + kBitsTile - 1) / kBitsTile;
  for (int h = 1; h <= s->L; ++h) {
    const int64_t cap_front = s->cap_layer[h - 1];
    uint32_t* bitmap = s->bitmaps + (size_t)(h - 1) * s->nwords_pad;
    if (h >= 2) {
      FrontArgs fa{g->indptr, s->layer[h - 1], &s->counts->n_layer[h - 1], cap_front, s->fanouts[h - 1], s->row_off[h],
                   s->tile_b, &s->counts->e_hop[h], s->ctl, h};
      const int grid_f = (int)std::min<int64_t>((cap_front + kFrontTile - 1) / kFrontTile, (int64_t)pg::sm_count(dev) * 2);
      pg::prefer_max_smem_k(front_kernel);
      if (dbg_n++ < dbg_limit) front_kernel<<<std::max(grid_f, 1), kSeedThreads, 0, st>>>(fa);
      PG_CHECK_LAUNCH();
    }
    PickArgs pa{g->indptr, g->indices, g->eids, s->layer[h - 1], &s->counts->n_layer[h - 1], cap_front, s->row_off[h],
                s->fanouts[h - 1], (uint32_t)h, k0, k1, d_key, s->nb_src[h], s->nb_eid[h], s->cap_edges, bitmap,
                s->scratch, s->scratch_stride};
    const int grid_p = (int)std::min<int64_t>(s->pick_grid, std::max<int64_t>(1, (cap_front + kPickWarps - 1) / kPickWarps));
    pg::prefer_max_smem_k(pick_kernel);
    if (dbg_n++ < dbg_limit) pick_kernel<<<grid_p, kPickWarps * 32, 0, st>>>(pa);
    PG_CHECK_LAUNCH();
    BitsArgs ba{bitmap, s->nwords, s->word_prefix + (size_t)(h - 1) * s->nwords_pad, s->layer[h], s->cap_layer[h],
                s->tile_a, &s->counts->n_layer[h], s->ctl, h};
    const int grid_b = (int)std::min<int64_t>(ntiles, (int64_t)pg::sm_count(dev) * 4);
    pg::prefer_max_smem_k(bits_kernel);
    if (dbg_n++ < dbg_limit) bits_kernel<<<grid_b, kBitsThreads, 0, st>>>(ba);
    PG_CHECK_LAUNCH();
  }
  // ---- assemble
  AssembleArgs aa;
  aa.L = s->L;
  aa.counts = s->counts;
  for (int h = 0; h <= PG_MAX_HOPS; ++h) {
    aa.layer[h] = s->layer[h];
    aa.nb_src[h] = s->nb_src[h];
    aa.nb_eid[h] = s->nb_eid[h];
    aa.row_off[h] = s->row_off[h];
    aa.bitmap[h] = (h >= 1 && h <= s->L) ? s->bitmaps + (size_t)(h - 1) * s->nwords_pad : nullptr;
    aa.word_prefix[h] = (h >= 1 && h <= s->L) ? s->word_prefix + (size_t)(h - 1) * s->nwords_pad : nullptr;
    aa.cap_layer[h] = s->cap_layer[h];
  }
  aa.cap_nodes = s->cap_nodes;
  aa.cap_edges = s->cap_edges;
  aa.out = *out;
  aa.labels = d_labels;
  aa.seed_labels = d_seed_labels;
  pg::prefer_max_smem_k(assemble_kernel);
  if (dbg_n++ < dbg_limit) assemble_kernel<<<grid_for(s->cap_nodes + s->cap_edges, 256, dev, 4), 256, 0, st>>>(aa);
  PG_CHECK_LAUNCH();
  if (h_meta) PG_CUDA(cudaMemcpyAsync(h_meta, out->meta, PG_META_LEN * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  return PG_OK;
}

}  // extern "C"
