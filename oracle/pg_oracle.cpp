// oracle/pg_oracle.cpp — CPU restatement of the PaGraph hot path. TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (pagraph_b200/) never does.
//
// PARITY STATUS
//   * sampling / NodeFlow construction: the arithmetic lives in the third-party dependency
//     dgl==0.4.1 (reference README.md:14), which is NOT present in /root/reference and is not
//     installable here. This file restates its published algorithm (src/graph/sampler.cc:
//     SampleSubgraph / GetUniformSample / ConstructNodeFlow — SURVEY.md Appendix A.3/A.4),
//     anchored on the reference's call sites (examples/profile/pa_gcn.py:71-76,
//     PaGraph/partition/utils.py:11-30, examples/eval.py:20-25). The reference holds no golden
//     vectors for it => "parity unpinned" for sampled ids; RNG-free (full-fanout) cases are pinned
//     by construction because the NodeFlow is then a pure function of (graph, seeds).
//   * DGL draws from an unseeded thread-local std::default_random_engine, so the reference is
//     not reproducible with itself; oracle and GPU share the counter-based generator below.
//     pgo_sample_stdlib draws the same minibatch from DGL's own generator (libstdc++ minstd_rand0 +
//     uniform_int_distribution, one sequential stream, discovery-order expansion) given a seed: it
//     cannot be compared draw by draw with the counter-based mode, but it shares GetUniformSample's
//     structure and ConstructNodeFlow with it and must agree exactly wherever no draw is made.
//   * gather (PaGraph/storage/storage.py:157-216) is pinned by golden vectors produced by the
//     real reference module (tests/golden/make_golden.py).
//   * aggregation (dgl block_compute copy_src + sum/mean, call sites PaGraph/model/gcn_nssc.py:71-74)
//     is restated in float64, sequential edge order.
//
// RNG contract (shared with pagraph_b200/csrc/pg_common.cuh: philox4x32_10, draw_pos)
//   mbkey    = philox4x32_10(ctr=(epoch_lo, epoch_hi, batch_lo, batch_hi), key=(seed_lo, seed_hi))[0..1]
//   draw(v, hop, t, deg) = mulhi64(w0 | w1<<32, deg),  (w0..w3) = philox4x32_10(ctr=(v_lo, v_hi, hop, t), key=mbkey)
//   (range reduction bias < deg * 2^-64). hop = 1 for the expansion of the seeds.
//   GetUniformSample(deg, k): deg<=k -> all; deg>2k -> draw t=0,1,.. into a set until k distinct, sort;
//   k<deg<=2k -> same with deg-k members = excluded positions, keep the complement ascending.

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Philox {
  static inline void round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  static inline void run(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint32_t k[2] = {k0, k1};
    for (int r = 0; r < 10; ++r) {
      round(c, k);
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
  }
};

inline void minibatch_key(uint64_t seed, int64_t epoch, int64_t batch, uint32_t* k0, uint32_t* k1) {
  uint32_t c[4] = {(uint32_t)(uint64_t)epoch, (uint32_t)((uint64_t)epoch >> 32),
                   (uint32_t)(uint64_t)batch, (uint32_t)((uint64_t)batch >> 32)};
  Philox::run(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  *k0 = c[0];
  *k1 = c[1];
}

inline uint64_t draw_pos(uint32_t k0, uint32_t k1, int64_t v, uint32_t hop, uint32_t t, uint64_t deg) {
  uint32_t c[4] = {(uint32_t)(uint64_t)v, (uint32_t)((uint64_t)v >> 32), hop, t};
  Philox::run(c, k0, k1);
  const uint64_t r = (uint64_t)c[0] | ((uint64_t)c[1] << 32);
  return (uint64_t)(((unsigned __int128)r * deg) >> 64);
}

// SURVEY.md Appendix A.3 GetUniformSample: returns the chosen positions in [0,deg), ascending.
void uniform_positions(uint32_t k0, uint32_t k1, int64_t v, uint32_t hop, int64_t deg, int64_t k,
                       std::vector<int64_t>* pos) {
  pos->clear();
  if (deg <= k) {
    for (int64_t p = 0; p < deg; ++p) pos->push_back(p);
    return;
  }
  const bool complement = deg <= 2 * k;
  const int64_t m = complement ? deg - k : k;
  std::unordered_set<int64_t> chosen;
  std::vector<int64_t> order;
  for (uint32_t t = 0; (int64_t)chosen.size() < m; ++t) {
    const int64_t p = (int64_t)draw_pos(k0, k1, v, hop, t, (uint64_t)deg);
    if (chosen.insert(p).second) order.push_back(p);
  }
  if (!complement) {
    std::sort(order.begin(), order.end());
    *pos = order;
  } else {
    for (int64_t p = 0; p < deg; ++p)
      if (!chosen.count(p)) pos->push_back(p);
  }
}

struct NodeFlow {
  int64_t num_layers = 0;
  std::vector<int64_t> node_mapping, layer_offsets, indptr, indices, edge_mapping, flow_offsets;
};

// SURVEY.md Appendix A.4 ConstructNodeFlow. layer[h]: vertices of sampling layer h (0 = seeds, in seed order; h >= 1 ascending);
// nb_src / nb_eid / nb_off[h]: sampled sources, edge ids and row offsets of hop h, rows in the NodeFlow order of layer[h-1].
NodeFlow* construct_nodeflow(const std::vector<std::vector<int64_t>>& layer, const std::vector<std::vector<int64_t>>& nb_src,
                             const std::vector<std::vector<int64_t>>& nb_eid, const std::vector<std::vector<int64_t>>& nb_off,
                             int L) {
  NodeFlow* nf = new NodeFlow;
  nf->num_layers = L + 1;
  nf->layer_offsets.push_back(0);
  for (int j = 0; j <= L; ++j) {
    const std::vector<int64_t>& lay = layer[L - j];
    nf->node_mapping.insert(nf->node_mapping.end(), lay.begin(), lay.end());
    nf->layer_offsets.push_back((int64_t)nf->node_mapping.size());
  }
  nf->indptr.assign(nf->layer_offsets[1] + 1, 0);  // layer-0 rows are empty
  nf->flow_offsets.push_back(0);
  for (int j = 1; j <= L; ++j) {
    const int h = L - j + 1;                       // hop whose expansion feeds NodeFlow layer j
    const std::vector<int64_t>& srcs = layer[h];   // NodeFlow layer j-1 (sorted)
    const int64_t col_base = nf->layer_offsets[j - 1];
    const int64_t e_base = (int64_t)nf->indices.size();
    for (size_t e = 0; e < nb_src[h].size(); ++e) {
      const int64_t rank = std::lower_bound(srcs.begin(), srcs.end(), nb_src[h][e]) - srcs.begin();
      nf->indices.push_back(col_base + rank);
      nf->edge_mapping.push_back(nb_eid[h][e]);
    }
    for (size_t r = 1; r < nb_off[h].size(); ++r) nf->indptr.push_back(e_base + nb_off[h][r]);
    nf->flow_offsets.push_back((int64_t)nf->indices.size());
  }
  return nf;
}

// SURVEY.md Appendix A.3 SampleSubgraph (counter-based generator of the RNG contract above).
NodeFlow* sample_one(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t V,
                     const int64_t* seeds, int64_t n_seeds, int num_hops, const int64_t* fanouts,
                     uint64_t seed, int64_t epoch, int64_t batch) {
  uint32_t k0, k1;
  minibatch_key(seed, epoch, batch, &k0, &k1);
  const int L = num_hops;
  std::vector<std::vector<int64_t>> layer(L + 1);               // sampling order: 0 = seeds
  std::vector<std::vector<int64_t>> nb_src(L + 1), nb_eid(L + 1);  // per hop h (1..L), concatenated
  std::vector<std::vector<int64_t>> nb_off(L + 1);              // per hop row offsets over layer[h-1]
  {
    std::unordered_set<int64_t> seen;
    for (int64_t i = 0; i < n_seeds; ++i)
      if (seen.insert(seeds[i]).second) layer[0].push_back(seeds[i]);
  }
  std::vector<int64_t> pos;
  for (int h = 1; h <= L; ++h) {
    // The previous layer is expanded in its NodeFlow order: seed order for h==1, ascending
    // parent id otherwise. (DGL expands in discovery order and sorts afterwards; since draws are
    // keyed by the vertex id, not by its position, both orders give the same edges per vertex.)
    const std::vector<int64_t>& front = layer[h - 1];
    std::unordered_set<int64_t> seen;
    nb_off[h].push_back(0);
    for (int64_t v : front) {
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      uniform_positions(k0, k1, v, (uint32_t)h, deg, fanouts[h - 1], &pos);
      for (int64_t p : pos) {
        nb_src[h].push_back(indices[s + p]);
        nb_eid[h].push_back(eids ? eids[s + p] : s + p);
        seen.insert(indices[s + p]);
      }
      nb_off[h].push_back((int64_t)nb_src[h].size());
    }
    layer[h].assign(seen.begin(), seen.end());
    std::sort(layer[h].begin(), layer[h].end());
  }
  (void)V;
  return construct_nodeflow(layer, nb_src, nb_eid, nb_off, L);
}

// DGL's own random stream: ONE std::default_random_engine — libstdc++'s minstd_rand0 —
// consumed sequentially through std::uniform_int_distribution<size_t>(0, deg - 1), the previous layer expanded in
// DISCOVERY order (the order in which its vertices were first seen), layers sorted only when the NodeFlow is built.
// DGL seeds the engine from std::random_device unless dgl.random.seed is called, which PaGraph never does, so the
// reference is not reproducible with itself; with an explicit seed this mode reproduces what a single sampling thread
// of DGL 0.4.1 built against libstdc++ would draw. It shares GetUniformSample's structure and ConstructNodeFlow with
// the counter-based mode above and exists to cross-check them (tests/test_oracle.py): RNG-free cases must agree
// exactly, random cases must satisfy the same invariants and the same inclusion frequencies.
NodeFlow* sample_one_stdlib(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t V,
                            const int64_t* seeds, int64_t n_seeds, int num_hops, const int64_t* fanouts, uint64_t seed) {
  std::default_random_engine engine((std::default_random_engine::result_type)seed);
  auto rand_int = [&](int64_t upper) {                        // RandInt(upper): uniform in [0, upper)
    return (int64_t)std::uniform_int_distribution<size_t>(0, (size_t)upper - 1)(engine);
  };
  const int L = num_hops;
  std::vector<std::vector<int64_t>> disc(L + 1), layer(L + 1);
  std::vector<std::vector<int64_t>> nb_src(L + 1), nb_eid(L + 1), nb_off(L + 1);
  {
    std::unordered_set<int64_t> seen;
    for (int64_t i = 0; i < n_seeds; ++i)
      if (seen.insert(seeds[i]).second) disc[0].push_back(seeds[i]);
    layer[0] = disc[0];
  }
  struct Rec { int64_t v, start, cnt; };
  for (int h = 1; h <= L; ++h) {
    std::unordered_set<int64_t> seen;
    std::vector<Rec> recs;
    std::vector<int64_t> tsrc, teid, pos;
    const int64_t k = fanouts[h - 1];
    for (int64_t v : disc[h - 1]) {
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      pos.clear();
      if (deg <= k) {
        for (int64_t p = 0; p < deg; ++p) pos.push_back(p);
      } else {
        const bool complement = deg <= 2 * k;
        const int64_t m = complement ? deg - k : k;
        std::unordered_set<int64_t> chosen;
        while ((int64_t)chosen.size() < m) chosen.insert(rand_int(deg));
        if (!complement) {
          pos.assign(chosen.begin(), chosen.end());
          std::sort(pos.begin(), pos.end());
        } else {
          for (int64_t p = 0; p < deg; ++p)
            if (!chosen.count(p)) pos.push_back(p);
        }
      }
      recs.push_back({v, (int64_t)tsrc.size(), (int64_t)pos.size()});
      for (int64_t p : pos) {
        const int64_t u = indices[s + p];
        tsrc.push_back(u);
        teid.push_back(eids ? eids[s + p] : s + p);
        if (seen.insert(u).second) disc[h].push_back(u);
      }
    }
    // rows of the block follow the NodeFlow order of the expanded layer: seed order for the seeds, ascending id otherwise
    if (h > 1) std::sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) { return x.v < y.v; });
    nb_off[h].push_back(0);
    for (const Rec& r : recs) {
      nb_src[h].insert(nb_src[h].end(), tsrc.begin() + r.start, tsrc.begin() + r.start + r.cnt);
      nb_eid[h].insert(nb_eid[h].end(), teid.begin() + r.start, teid.begin() + r.start + r.cnt);
      nb_off[h].push_back((int64_t)nb_src[h].size());
    }
    layer[h] = disc[h];
    std::sort(layer[h].begin(), layer[h].end());
  }
  (void)V;
  return construct_nodeflow(layer, nb_src, nb_eid, nb_off, L);
}

}  // namespace

// dgl block_compute(copy_src, sum|mean) restated: accumulation in edge
// order, zero-in-degree rows -> 0, mean divides by max(deg,1). mode: 0 = sum, 1 = mean.
// indptr has n_dst+1 entries (absolute edge offsets), cols[e]-col_base indexes src rows.
// Acc = double for the parity checker; Acc = float is DGL's own CPU arithmetic (timing arm).
template <typename Acc>
static void aggregate_impl(const int64_t* indptr, const int64_t* cols, int64_t col_base, const float* src,
                           int64_t n_dst, int64_t dim, int mode, float* dst, int threads) {
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    std::vector<Acc> acc((size_t)dim);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
    for (int64_t r = 0; r < n_dst; ++r) {
      std::fill(acc.begin(), acc.end(), (Acc)0);
      const int64_t s = indptr[r], e = indptr[r + 1];
      for (int64_t j = s; j < e; ++j) {
        const float* row = src + (cols[j] - col_base) * dim;
        for (int64_t d = 0; d < dim; ++d) acc[(size_t)d] += (Acc)row[d];
      }
      const Acc scale = (mode == 1) ? (Acc)1 / (Acc)std::max<int64_t>(e - s, 1) : (Acc)1;
      for (int64_t d = 0; d < dim; ++d) dst[r * dim + d] = (float)(acc[(size_t)d] * scale);
    }
  }
  (void)threads;
}

extern "C" {

void pgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  Philox::run(c, key[0], key[1]);
  std::memcpy(out, c, sizeof(c));
}

uint64_t pgo_draw(uint64_t seed, int64_t epoch, int64_t batch, int64_t v, uint32_t hop, uint32_t t,
                  uint64_t deg) {
  uint32_t k0, k1;
  minibatch_key(seed, epoch, batch, &k0, &k1);
  return draw_pos(k0, k1, v, hop, t, deg);
}

void* pgo_sample(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t V,
                 const int64_t* seeds, int64_t n_seeds, int num_hops, const int64_t* fanouts,
                 uint64_t seed, int64_t epoch, int64_t batch) {
  return sample_one(indptr, indices, eids, V, seeds, n_seeds, num_hops, fanouts, seed, epoch, batch);
}

// The same minibatch drawn from DGL's own generator (std::default_random_engine, sequential; see sample_one_stdlib).
void* pgo_sample_stdlib(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t V,
                        const int64_t* seeds, int64_t n_seeds, int num_hops, const int64_t* fanouts, uint64_t seed) {
  return sample_one_stdlib(indptr, indices, eids, V, seeds, n_seeds, num_hops, fanouts, seed);
}

// sizes: [num_layers, total_nodes, total_edges]
void pgo_nf_sizes(void* h, int64_t* sizes) {
  NodeFlow* nf = (NodeFlow*)h;
  sizes[0] = nf->num_layers;
  sizes[1] = (int64_t)nf->node_mapping.size();
  sizes[2] = (int64_t)nf->indices.size();
}

void pgo_nf_copy(void* h, int64_t* node_mapping, int64_t* layer_offsets, int64_t* indptr,
                 int64_t* indices, int64_t* edge_mapping, int64_t* flow_offsets) {
  NodeFlow* nf = (NodeFlow*)h;
  auto cp = [](int64_t* d, const std::vector<int64_t>& s) {
    if (!s.empty()) std::memcpy(d, s.data(), s.size() * sizeof(int64_t));
  };
  cp(node_mapping, nf->node_mapping);
  cp(layer_offsets, nf->layer_offsets);
  cp(indptr, nf->indptr);
  cp(indices, nf->indices);
  cp(edge_mapping, nf->edge_mapping);
  cp(flow_offsets, nf->flow_offsets);
}

void pgo_nf_free(void* h) { delete (NodeFlow*)h; }

// CPU-baseline helper: sample `n_batches` consecutive minibatches the way DGL runs them —
// several batches in flight, one thread each (OpenMP over batches). Returns total nodes+edges
// (so the work cannot be optimised away); per-batch sizes go to out_nodes/out_edges if non-null.
int64_t pgo_sample_batches(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t V,
                           const int64_t* seeds, int64_t n_seeds, int64_t batch_size, int64_t first_batch,
                           int64_t n_batches, int num_hops, const int64_t* fanouts, uint64_t seed,
                           int64_t epoch, int threads, int64_t* out_nodes, int64_t* out_edges) {
  int64_t total = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : total)
#endif
  for (int64_t b = 0; b < n_batches; ++b) {
    const int64_t lo = (first_batch + b) * batch_size;
    if (lo >= n_seeds) continue;
    const int64_t n = std::min(batch_size, n_seeds - lo);
    NodeFlow* nf = sample_one(indptr, indices, eids, V, seeds + lo, n, num_hops, fanouts, seed, epoch,
                              first_batch + b);
    if (out_nodes) out_nodes[b] = (int64_t)nf->node_mapping.size();
    if (out_edges) out_edges[b] = (int64_t)nf->indices.size();
    total += (int64_t)nf->node_mapping.size() + (int64_t)nf->indices.size();
    delete nf;
  }
  (void)threads;
  return total;
}

// Same parallelism model, but the NodeFlows are kept: handles[b] (pgo_nf_* accessors, pgo_nf_free).
void pgo_sample_many(const int64_t* indptr, const int64_t* indices, const int64_t* eids, int64_t V,
                     const int64_t* seeds, int64_t n_seeds, int64_t batch_size, int64_t first_batch,
                     int64_t n_batches, int num_hops, const int64_t* fanouts, uint64_t seed, int64_t epoch,
                     int threads, void** handles) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
#endif
  for (int64_t b = 0; b < n_batches; ++b) {
    const int64_t lo = (first_batch + b) * batch_size;
    const int64_t n = lo < n_seeds ? std::min(batch_size, n_seeds - lo) : 0;
    handles[b] = sample_one(indptr, indices, eids, V, seeds + (n ? lo : 0), n, num_hops, fanouts, seed, epoch,
                            first_batch + b);
  }
  (void)threads;
}

// storage.py:173-204 restated for one field: out[j] = flag[t_j] ? cache[l2c[t_j]] : host[nid_map[t_j]].
// Returns the number of misses. hit_mask (optional) receives flag[t_j].
int64_t pgo_fetch(const int64_t* tnid, int64_t n, const uint8_t* flag, const int64_t* l2c,
                  const int64_t* nid_map, const float* cache, int64_t cache_stride, const float* host,
                  int64_t host_stride, int64_t dim, float* out, uint8_t* hit_mask, int threads) {
  int64_t miss = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : miss)
#endif
  for (int64_t j = 0; j < n; ++j) {
    const int64_t t = tnid[j];
    const bool hit = flag[t] != 0;
    const float* src = hit ? cache + l2c[t] * cache_stride : host + nid_map[t] * host_stride;
    std::memcpy(out + j * dim, src, (size_t)dim * sizeof(float));
    if (hit_mask) hit_mask[j] = hit;
    miss += !hit;
  }
  (void)threads;
  return miss;
}

void pgo_aggregate(const int64_t* indptr, const int64_t* cols, int64_t col_base, const float* src,
                   int64_t n_dst, int64_t dim, int mode, float* dst, int threads) {
  aggregate_impl<double>(indptr, cols, col_base, src, n_dst, dim, mode, dst, threads);
}

void pgo_aggregate_f32(const int64_t* indptr, const int64_t* cols, int64_t col_base, const float* src,
                       int64_t n_dst, int64_t dim, int mode, float* dst, int threads) {
  aggregate_impl<float>(indptr, cols, col_base, src, n_dst, dim, mode, dst, threads);
}

// Backward of the above: grad_src[u] += grad_dst[v] * scale(v) over block edges (float64 accumulate).
void pgo_aggregate_bwd(const int64_t* indptr, const int64_t* cols, int64_t col_base, const float* grad_dst,
                       int64_t n_dst, int64_t n_src, int64_t dim, int mode, float* grad_src) {
  // Column-parallel: every thread owns a slice of the feature columns and walks the edges in the same sequential order,
  // so the float64 sums do not depend on the thread count.
  std::vector<double> acc((size_t)(n_src * dim), 0.0);
#pragma omp parallel
  {
#ifdef _OPENMP
    const int64_t nt = omp_get_num_threads(), tid = omp_get_thread_num();
#else
    const int64_t nt = 1, tid = 0;
#endif
    const int64_t d0 = dim * tid / nt, d1 = dim * (tid + 1) / nt;
    if (d0 < d1)
      for (int64_t r = 0; r < n_dst; ++r) {
        const int64_t s = indptr[r], e = indptr[r + 1];
        const double scale = (mode == 1) ? 1.0 / (double)std::max<int64_t>(e - s, 1) : 1.0;
        for (int64_t j = s; j < e; ++j) {
          double* a = acc.data() + (cols[j] - col_base) * dim;
          for (int64_t d = d0; d < d1; ++d) a[d] += (double)grad_dst[r * dim + d] * scale;
        }
      }
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_src * dim; ++i) grad_src[i] = (float)acc[(size_t)i];
}

int pgo_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
