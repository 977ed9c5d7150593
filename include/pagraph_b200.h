/* pagraph_b200.h — C-ABI of the B200-native PaGraph hot path (libpagraph_b200.so).
 *
 * The reference (zhiqi-0/PaGraph) has no FFI: its boundary is the Python API the trainer calls.
 * Each entry point below replaces the work behind one reference interface (file:line cited,
 * relative to the reference tree); pagraph_b200/*.py mirrors those Python interfaces and is the
 * only caller (ctypes). INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types. Device pointers are CUDA device pointers of
 *     the device the handle was created on; "host" pointers must be page-locked + mapped
 *     (pg_host_alloc / pg_host_register) when a kernel reads them.
 *   - every function returns pg_status (0 = ok); pg_last_error() returns a thread-local message.
 *   - nothing throws across the ABI; nothing synchronises the device unless stated ("SYNC").
 *   - the caller owns every output buffer; work is enqueued on the caller's stream
 *     (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - handles are not thread-safe: one per process/GPU, like the reference's process model
 *     (examples/profile/pa_gcn.py:157 — mp.spawn, one trainer per GPU).
 *   - ids are int64 (the reference's width: dgl_id_t / torch.LongTensor).
 */
#ifndef PAGRAPH_B200_H_
#define PAGRAPH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int pg_status;
enum {
  PG_OK = 0,
  PG_ERR_INVALID = 1,   /* bad argument */
  PG_ERR_CUDA = 2,      /* a CUDA runtime call failed; see pg_last_error() */
  PG_ERR_OVERFLOW = 3,  /* output capacity too small (reported through meta, see pg_sample) */
  PG_ERR_NOMEM = 4
};

typedef struct pg_graph pg_graph;     /* in-CSR adjacency resident in HBM */
typedef struct pg_sampler pg_sampler; /* sampling workspace bound to one graph */
typedef struct pg_cache pg_cache;     /* feature-cache lookup state (flag / l2c / nid_map / tables) */
typedef struct pg_peer_group pg_peer_group; /* NVLink peer memory of the one-node gradient all-reduce */

#define PG_MAX_FIELDS 4
#define PG_MAX_HOPS 8
#define PG_MAX_RANKS 8            /* GPUs of one node in a peer group */
#define PG_IPC_HANDLE_BYTES 64    /* sizeof(cudaIpcMemHandle_t) */

/* ---------------------------------------------------------------- runtime */
int pg_version(void);
const char* pg_last_error(void);
pg_status pg_device_info(int dev, int* sm_count, size_t* total_mem, size_t* free_mem); /* SYNC */

/* Page-locked, device-mapped host memory: the B200 replacement for the pageable POSIX-shm tensors
 * DGL's graph store hands out (server/pa_server.py:53-54, PaGraph/storage/storage.py:128). */
pg_status pg_host_alloc(void** ptr, size_t bytes);
pg_status pg_host_free(void* ptr);
pg_status pg_host_register(void* ptr, size_t bytes); /* pin + map an existing mapping (e.g. /dev/shm) */
pg_status pg_host_unregister(void* ptr);

/* ---------------------------------------------------------------- graph (replaces DGLGraph(adj, readonly=True),
 * examples/profile/pa_gcn.py:36: the structure the sampler walks). In-CSR: row v lists the sources
 * of edges u->v in increasing edge-id order; `eids` may be NULL (edge id = CSR position).
 * Arrays are HOST pointers and are copied to the device. SYNC. */
pg_status pg_graph_create(const int64_t* indptr, const int64_t* indices, const int64_t* eids,
                          int64_t num_nodes, int64_t num_edges, int dev, pg_graph** out);
/* Same, but the arrays already live on device `dev` and are borrowed (caller keeps them alive). */
pg_status pg_graph_create_device(const int64_t* d_indptr, const int64_t* d_indices, const int64_t* d_eids,
                                 int64_t num_nodes, int64_t num_edges, int dev, pg_graph** out);
void pg_graph_destroy(pg_graph* g);
/* deg[v] = in-degree (row length) if in_edges else out-degree (column count). d_out: device int64[num_nodes]. */
pg_status pg_graph_degrees(pg_graph* g, int in_edges, int64_t* d_out, void* stream);

/* ---------------------------------------------------------------- sampler (replaces dgl.contrib.sampling.NeighborSampler's
 * C++ SampleSubgraph + ConstructNodeFlow; call site examples/profile/pa_gcn.py:71-76).
 * fanouts[h] = expand factor of hop h+1 (index 0 expands the seeds); the reference passes one
 * scalar for all hops. max_seeds / cap_nodes / cap_edges size the workspace and the outputs. */
pg_status pg_sampler_create(pg_graph* g, int num_hops, const int64_t* fanouts, uint64_t seed,
                            int64_t max_seeds, int64_t cap_nodes, int64_t cap_edges, pg_sampler** out);
void pg_sampler_destroy(pg_sampler* s);

/* Device output buffers of one NodeFlow (layer 0 = inputs, layer L = seeds). */
typedef struct {
  int64_t* node_mapping;  /* [cap_nodes]   parent id of every NodeFlow node, layer by layer          */
  int64_t* indptr;        /* [cap_nodes+1] CSR over all NodeFlow nodes (layer-0 rows empty)          */
  int64_t* indices;       /* [cap_edges]   NodeFlow id of each edge's source (previous layer)        */
  int64_t* edge_mapping;  /* [cap_edges]   parent edge ids                                           */
  int64_t* meta;          /* [PG_META_LEN] see below                                                 */
} pg_nodeflow_buffers;

/* meta layout (int64): [0] status (0 ok, PG_ERR_OVERFLOW if a capacity was exceeded — then [1],[2]
 * hold the capacities that would have sufficed so far), [1] total nodes, [2] total edges,
 * [3] num_layers, [4 .. 4+num_layers] layer_offsets, then [.. +num_layers-1 +1] flow_offsets. */
#define PG_META_LEN (4 + (PG_MAX_HOPS + 2) + (PG_MAX_HOPS + 1))

/* Sample one minibatch. d_seeds: device int64[n_seeds] (duplicates allowed; first occurrence kept,
 * order preserved). (epoch, batch) key the counter-based RNG (oracle/pg_oracle.cpp header).
 * If h_meta (pinned host, PG_META_LEN int64) is non-NULL the meta block is also copied there on
 * `stream`; the caller synchronises on the stream/event before reading it. */
pg_status pg_sample(pg_sampler* s, const int64_t* d_seeds, int64_t n_seeds, int64_t epoch, int64_t batch,
                    const pg_nodeflow_buffers* out, int64_t* h_meta, void* stream);
/* Same with the minibatch key read from device memory (uint32[2], as produced by pg_minibatch_key): the launch
 * sequence then depends on nothing but pointers, so one captured CUDA graph samples every minibatch. */
void pg_minibatch_key(uint64_t seed, int64_t epoch, int64_t batch, uint32_t* key);
/* d_labels / d_seed_labels (optional, both or neither): `label = labels[nf.layer_parent_nid(-1)]` of the trainer
 * (examples/profile/pa_gcn.py:89-90) done in the same pass — d_seed_labels[i] = d_labels[parent id of seed-layer row i],
 * i.e. in the order of the DEDUPLICATED seed layer, which is the row order of the model's output. */
pg_status pg_sample_keyed(pg_sampler* s, const int64_t* d_seeds, int64_t n_seeds, const uint32_t* d_key,
                          const pg_nodeflow_buffers* out, int64_t* h_meta, const int64_t* d_labels, int64_t* d_seed_labels,
                          void* stream);

/* ---------------------------------------------------------------- feature cache (replaces PaGraph/storage/storage.py) */
typedef struct {
  int32_t dim;               /* floats per row                                                       */
  int64_t host_stride;       /* floats between consecutive rows of the host table                    */
  const float* host_table;   /* [V_full, dim] pinned+mapped host rows, indexed by FULL-graph id      */
} pg_field;

/* GraphCacheServer.__init__ (storage.py:23-56). The lookup state is caller-owned DEVICE memory the
 * handle borrows (the Python attributes gpu_flag / localid2cacheid / nid_map, storage.py:34-51):
 * d_flag uint8[node_num] (0 = row lives on the host), d_l2c int64[node_num] (local id -> cache row),
 * d_nid_map int64[node_num] (sub-graph id -> full-graph id). */
pg_status pg_cache_create(int64_t node_num, uint8_t* d_flag, int64_t* d_l2c, const int64_t* d_nid_map,
                          int nfields, const pg_field* fields, int dev, pg_cache** out);
void pg_cache_destroy(pg_cache* c);

/* cache_fix_data (storage.py:135-154): l2c[nids[i]] = i, flag[nids[i]] = 1; d_cache_tables[f] are
 * caller-allocated [n, dim] device buffers that the handle keeps borrowing afterwards. If
 * copy_rows != 0 the rows host_table[f][nid_map[nids[i]]] are pulled into d_cache_tables[f][i] by the
 * GPU (auto_cache, storage.py:94-95,103-104); with copy_rows == 0 the caller has filled them.
 * d_nids: device int64[n]. */
pg_status pg_cache_fill(pg_cache* c, const int64_t* d_nids, int64_t n, int is_full,
                        float* const* d_cache_tables, int copy_rows, void* stream);

/* get_feat_from_server(to_gpu=True) (storage.py:107-132): d_out[f][i] = host_table[f][nid_map[d_nids[i]]]. */
pg_status pg_cache_fetch_host(pg_cache* c, const int64_t* d_nids, int64_t n, float* const* d_out,
                              void* stream);

/* fetch_data (storage.py:157-204) for n consecutive NodeFlow nodes (all layers at once):
 *   d_out[f][j] = flag[t_j] ? cache[f][l2c[t_j]] : host_table[f][nid_map[t_j]],  t_j = d_parent_ids[j].
 * d_hit_mask (optional, uint8[n]) receives flag[t_j]; d_counts (int64[2], optional) is
 * INCREMENTED by (tries, misses) — storage.py:219-221. When the cache is full (is_full) this is
 * fetch_from_cache (storage.py:207-216).
 * mode: 0 = auto, 1 = plain vector loads for misses, 2 = TMA bulk copies for misses (needs 16-byte
 * aligned rows). */
pg_status pg_cache_fetch(pg_cache* c, const int64_t* d_parent_ids, int64_t n, float* const* d_out,
                         uint8_t* d_hit_mask, int64_t* d_counts, int mode, void* stream);
/* Same for the id range [*d_begin, *d_end) of d_ids_base, the range living on the device (e.g. two entries of a
 * NodeFlow's meta block): nothing about the minibatch's size is needed on the host, so the call can be captured in a
 * CUDA graph and replayed. `cap` bounds the row count (output capacity, grid sizing). d_ws: optional caller-owned
 * workspace int64[2 + 4 * cap] (hit / miss lists) — one per captured call site, so that nothing a replayed graph
 * addresses is shared with calls on other streams; NULL = the handle's shared workspace. */
pg_status pg_cache_fetch_dyn(pg_cache* c, const int64_t* d_ids_base, const int64_t* d_begin, const int64_t* d_end,
                             int64_t cap, float* const* d_out, int64_t* d_counts, int mode, int64_t* d_ws, void* stream);
/* ---------------------------------------------------------------- aggregation (replaces nf.block_compute(i, fn.copy_src,
 * fn.sum|fn.mean, ...), PaGraph/model/gcn_nssc.py:71-74, graphsage_nssc.py:98-106; and, run over
 * the full graph with mode=PG_AGG_SUM + norm, the server-side --preprocess fold, server/pa_server.py:45-52).
 *   dst[r] = scale_r * sum_{e in [indptr[r], indptr[r+1])} src[cols[e] - col_base]
 * scale_r = 1 (SUM), 1/max(deg_r,1) (MEAN); if d_norm != NULL the row is additionally multiplied by
 * d_norm[r] (NodeUpdate test=True, gcn_nssc.py:16-17). Strides in floats. */
enum { PG_AGG_SUM = 0, PG_AGG_MEAN = 1 };
pg_status pg_aggregate_fwd(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base,
                           const float* d_src, int64_t src_stride, float* d_dst, int64_t dst_stride,
                           int64_t n_dst, int32_t dim, int mode, const float* d_norm, void* stream);
/* grad_src[cols[e]-col_base] += scale_r * grad_dst[r]; d_grad_src ([n_src, dim]) is zeroed first. */
pg_status pg_aggregate_bwd(const int64_t* d_indptr, const int64_t* d_cols, int64_t col_base,
                           const float* d_grad_dst, int64_t gdst_stride, float* d_grad_src,
                           int64_t gsrc_stride, int64_t n_dst, int64_t n_src, int32_t dim, int mode,
                           const float* d_norm, void* stream);
/* Device-resident extents (see pg_block.d_layer_offsets): d_indptr_base is the NodeFlow-wide indptr, cap_dst / cap_src
 * are capacities; rows [n_dst, cap_dst) of d_dst are zero-filled, all cap_src rows of d_grad_src are zeroed first. */
pg_status pg_aggregate_fwd_dyn(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_layer_offsets,
                               const float* d_src, int64_t src_stride, float* d_dst, int64_t dst_stride, int64_t cap_dst,
                               int32_t dim, int mode, const float* d_norm, void* stream);
pg_status pg_aggregate_bwd_dyn(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_layer_offsets,
                               const float* d_grad_dst, int64_t gdst_stride, float* d_grad_src, int64_t gsrc_stride,
                               int64_t cap_dst, int64_t cap_src, int32_t dim, int mode, const float* d_norm, void* stream);

/* ---------------------------------------------------------------- fused cache lookup + aggregation
 * The GCN / GraphSAGE models consume the input layer's features only through the first block's aggregation
 * (PaGraph/model/gcn_nssc.py:64-74: activation = features; h = dropout(activation); block_compute(0, copy_src, mean)).
 * pg_cache_aggregate computes that block straight from the feature cache, without materialising the gathered
 * [n_src, dim] rows that fetch_data (storage.py:186-200) would write and block_compute would read back:
 *   dst[r] = scale_r * sum_{e in row r} drop(row(t_e)),   t_e = parent_ids[cols[e] - col_base],
 *   row(t) = flag[t] ? cache[field][l2c[t]] : host_table[field][nid_map[t]]
 * (missed rows are first pulled over PCIe into a staging buffer, one TMA bulk copy per row). drop() is inverted
 * dropout with probability dropout_p, mask keyed by (dropout_seed + *d_step, source node, column) so that every
 * edge of a source node sees the same dropped row, as dropout-then-aggregate does; dropout_p = 0 disables it.
 * d_norm / mode as pg_aggregate_fwd. Rows [n_dst, zero_rows_to) of d_dst are zero-filled (fixed-shape buffers); with
 * device-resident extents a negative zero_rows_to = -g zero-fills up to n_dst rounded up to a multiple of g.
 * d_counts (optional int64[2]) is incremented by (n_src, misses) like pg_cache_fetch. */
typedef struct {
  const int64_t* parent_ids; /* [n_src] parent (local) ids of the block's source layer               */
  const int64_t* indptr;     /* [n_dst+1] absolute offsets into cols                                  */
  const int64_t* cols;       /* NodeFlow ids of edge sources; cols[e] - col_base indexes parent_ids   */
  int64_t col_base, n_src, n_dst;
  /* Optional device-resident extents (CUDA-graph replays, sizes unknown to the host): int64[3] = the NodeFlow layer
   * offsets of the source layer, the destination layer and the layer after it (&meta[4 + block]). When non-NULL,
   * parent_ids / indptr are the NodeFlow-wide node_mapping / indptr arrays, col_base is ignored and n_src / n_dst are
   * capacities (grid sizing, padding); the kernels read the real extents on the device. */
  const int64_t* d_layer_offsets;
} pg_block;
/* The two stages of pg_cache_aggregate, callable separately so that stage 1 (with its PCIe transfer of the missed rows)
 * can run ahead on another stream while the previous minibatch computes. d_rowptr: caller-owned float*[n_src];
 * d_stage: caller-owned [stage_rows, dim]. Misses beyond stage_rows are not staged: their row pointer addresses the
 * pinned host table directly (slower, still correct). d_ws: optional caller-owned workspace int64[2 + stage_rows] (miss
 * counters + the host rows of the staged misses); a pipeline that replays captured graphs passes one per ring slot so
 * that calls on other streams (an eager fetch_data, an evaluation NodeFlow) can neither move nor race it. NULL = the
 * handle's shared workspace (grown on demand; replaced buffers stay allocated until pg_cache_destroy). */
pg_status pg_cache_resolve(pg_cache* c, int field, const pg_block* blk, const float** d_rowptr, float* d_stage,
                           int64_t stage_rows, int64_t* d_counts, int64_t* d_ws, void* stream);
/* Peer-GPU cache tier over NVLink (optional; not in the reference — extends the lookup of storage.py:157-204): the ranks of
 * one node shard the hottest rows between them. With the vertices in one agreed caching order (d_pos[t] = position of local
 * id t, < 0 = never cached), positions [0, c_local) are replicated in every rank's table and position c_local + j lives on
 * rank j % world at row c_local + j / world (j < world * c_shard). d_peer_tables[r]: rank r's cache table of `field`
 * ([c_local + c_shard, dim] fp32, a CUDA-IPC mapping for r != rank). pg_cache_resolve then points a row that is not in the
 * local table at the owner's HBM before falling back to the pinned host row; flag / l2c keep describing the LOCAL table, so
 * pg_cache_fetch is unchanged (it reads non-local rows from the host). d_peer_hits: optional device counter. world <= 1
 * switches the tier off. */
/* Fill the cache tables straight from full-graph row ids (no local-id bookkeeping): table row i <- host_table[d_full_rows[i]]
 * for every field, and install the tables (cached_rows = n). flag / l2c are left to the caller — a peer-tier owner also
 * holds rows of vertices its own partition never names. */
pg_status pg_cache_fill_rows(pg_cache* c, const int64_t* d_full_rows, int64_t n, float* const* d_cache_tables, void* stream);
/* Peer-visible device buffers for those tables: cudaMalloc + CUDA-IPC handle (PG_IPC_HANDLE_BYTES), opened by the other
 * ranks of the node with lazy peer access. */
pg_status pg_peer_alloc(size_t bytes, int dev, void** d_out, unsigned char* handle_out);
pg_status pg_peer_open(const unsigned char* handle, int dev, void** d_out);
pg_status pg_peer_close(void* d_ptr, int dev);
pg_status pg_peer_free(void* d_ptr, int dev);
pg_status pg_cache_set_peers(pg_cache* c, int field, int world, int rank, const float* const* d_peer_tables,
                             const int32_t* d_pos, int64_t c_local, int64_t c_shard, int64_t* d_peer_hits);
/* Reuse hint for the fused path (optional; no reference counterpart — the reference's gather has no notion of L2): d_hot is a
 * caller-owned uint8[node_num], non-zero for the local ids whose rows are worth keeping in the L2 cache — the highest
 * out-degree vertices recur as sources within a minibatch and from one minibatch to the next (auto_cache ranks by the same
 * out-degree, storage.py:98-101). pg_cache_resolve then sets bit 0 of the row pointers of those rows (rows are 16-byte
 * aligned); pg_aggregate_rows fetches a tagged row with the L2 evict_last priority and the others with evict_first, and
 * masks the bit off. Results are unchanged; NULL switches the hint off. */
pg_status pg_cache_set_hot(pg_cache* c, const uint8_t* d_hot);
/* Scheduling knob of the fused path (no reference counterpart): the row-fetching kernel of pg_aggregate_rows / pg_cache_aggregate
 * runs one CTA per SM and fills the SM's shared memory, so kernels that need more than ~27 KB of it cannot run beside it. A
 * pipeline that overlaps the input aggregation of the next minibatch with the small latency-bound kernels of the current one
 * leaves `n` SMs out of that grid for launches made after this call (process-wide; 0 = use every SM, the default). */
void pg_set_agg_reserve_sms(int n);
pg_status pg_aggregate_rows(const float* const* d_rowptr, const pg_block* blk, int32_t dim, float* d_dst,
                            int64_t dst_stride, int mode, const float* d_norm, float dropout_p, uint64_t dropout_seed,
                            const int64_t* d_step, int64_t zero_rows_to, void* stream);
pg_status pg_cache_aggregate(pg_cache* c, int field, const pg_block* blk, float* d_dst, int64_t dst_stride, int mode,
                             const float* d_norm, float dropout_p, uint64_t dropout_seed, const int64_t* d_step,
                             int64_t zero_rows_to, int64_t* d_counts, void* stream);

/* ---------------------------------------------------------------- the first NodeUpdate (PaGraph/model/gcn_nssc.py:14-24, :64-70)
 * out = concat ? cat(z, relu(z)) : relu(z),  z = x W^T + b, x [n, in_dim] = the aggregated input block (needs no
 * gradient), W [32, in_dim] contiguous, out [n, 64 | 32]. Both directions run on the tensor cores as error-compensated
 * TF32 (3xTF32, fp32 accumulate: fp32-level accuracy). in_dim % 4 == 0, <= 768; out_dim == 32.
 *   pg_linear_concat_fwd: writes out and, when d_out_drop != NULL, out_drop = dropout(out) (the `h = self.dropout(h)`
 *     of gcn_nssc.py:66-67 ahead of the next block_compute) under the mask contract of pg_cache_aggregate: element
 *     (row, col) is dropped when the (col % 4)-th 16-bit lane of hash(dropout_seed + *d_step, row, col/4) is
 *     below round(p * 65536); kept values are scaled by 1/(1-p). d_step: optional int64 on the device.
 *   pg_linear_concat_bwd: grad_weight [32, in_dim] = gz^T x and grad_bias [32] = sum_r gz, with gz (the gradient of z)
 *     recovered from grad_out and out: relu' and the concat split are folded in; with dropout_p > 0 grad_out is the
 *     gradient of out_drop and the mask is regenerated from the same (seed, *d_step). Both outputs are overwritten. */
pg_status pg_linear_concat_fwd(const float* d_x, int64_t x_stride, const float* d_weight, const float* d_bias, int64_t n,
                               int32_t in_dim, int32_t out_dim, int concat, float* d_out, int64_t out_stride,
                               float* d_out_drop, int64_t od_stride, float dropout_p, uint64_t dropout_seed,
                               const int64_t* d_step, void* stream);
pg_status pg_linear_concat_bwd(const float* d_x, int64_t x_stride, const float* d_grad_out, int64_t g_stride,
                               const float* d_out, int64_t out_stride, int64_t n, int32_t in_dim, int32_t out_dim, int concat,
                               float dropout_p, uint64_t dropout_seed, const int64_t* d_step, float* d_grad_weight,
                               float* d_grad_bias, void* stream);

/* Classifier head + loss in one pass (the last NodeUpdate, gcn_nssc.py:48, followed by torch.nn.CrossEntropyLoss,
 * examples/profile/pa_gcn.py:62,93-94): pred = a W^T + b, loss = mean_r(logsumexp(pred_r) - pred_r[label_r]); also emits
 * d loss/d a [n, in_dim], d loss/d W [n_classes, in_dim] and d loss/d b [n_classes] (all overwritten). in_dim, n_classes <= 64;
 * labels in [0, n_classes). d_lo (optional, device): int64[2]; the row count is then min(n, d_lo[1] - d_lo[0]) read on the
 * device — the seed layer's extent &meta[4 + L] of a sampled NodeFlow, whose size the host does not know when a captured
 * graph is replayed (duplicate seeds are dropped by the sampler). */
pg_status pg_linear_cross_entropy(const float* d_a, int64_t a_stride, const float* d_weight, const float* d_bias,
                                  const int64_t* d_labels, int64_t n, int32_t in_dim, int32_t n_classes, float* d_loss,
                                  float* d_grad_a, int64_t ga_stride, float* d_grad_weight, float* d_grad_bias,
                                  const int64_t* d_lo, void* stream);

/* The last NodeFlow block, the classifier head and the loss as ONE kernel, forward and backward (the last
 * nf.block_compute(i, fn.copy_src, fn.mean) feeding the last NodeUpdate, gcn_nssc.py:71-74 + :48, then CrossEntropyLoss,
 * pa_gcn.py:62,93-94, and their backward):
 *   a = reduce over the block of d_src rows;  loss = CE(a W^T + b, labels);
 *   d_grad_src [cap_src, in_dim] = d loss / d src (overwritten), d_grad_weight, d_grad_bias (overwritten).
 * Same results as pg_aggregate_fwd_dyn -> pg_linear_cross_entropy -> pg_aggregate_bwd_dyn without the two latency-bound
 * launches and the [n, in_dim] round trips of a and its gradient. The block is given like the _dyn calls: NodeFlow-wide
 * indptr base, cols, d_layer_offsets = &meta[4 + block] (device: source-layer, destination-layer and end offsets);
 * cap_dst / cap_src = row capacities. labels are indexed by destination row. in_dim (% 4 == 0), n_classes <= 64,
 * 16-byte aligned rows. */
pg_status pg_block_linear_cross_entropy(const int64_t* d_indptr_base, const int64_t* d_cols, const int64_t* d_layer_offsets,
                                        const float* d_src, int64_t src_stride, int64_t cap_dst, int64_t cap_src, int mode,
                                        const float* d_weight, const float* d_bias, const int64_t* d_labels, int32_t in_dim,
                                        int32_t n_classes, float* d_loss, float* d_grad_src, int64_t gsrc_stride,
                                        float* d_grad_weight, float* d_grad_bias, void* stream);

/* ---------------------------------------------------------------- gradient all-reduce fused with the optimizer step
 * Replaces DistributedDataParallel's all-reduce of the flat gradient followed by Adam (examples/profile/pa_gcn.py:65,96-97)
 * with ONE kernel over NVLink peer memory: every CTA pushes its slice of the gradient into every peer's receive area
 * (CUDA-IPC mapped), raises a per-slice flag, waits for the peers' slices, sums them in rank order, divides by the world
 * size and applies Adam to its slice in place. No NCCL call, no host synchronisation, capturable in a CUDA graph.
 *   pg_peer_group_create: allocates this rank's receive area for n floats and returns its CUDA IPC handle (exchange the
 *     handles of all ranks by any means, e.g. an all_gather); pg_peer_group_connect maps the peers' areas.
 *   pg_allreduce_adam: d_grad is replaced by the rank-averaged gradient; d_param / d_exp_avg / d_exp_avg_sq are updated in
 *     place with torch.optim.Adam's formulas (no amsgrad). d_step: the optimizer's step count for THIS step (float, on
 *     the device, already incremented); d_step_id: a step number >= 1 on the device that is equal on all ranks and grows
 *     by one per call (it is the flag value peers wait for). world == 1 degenerates to the optimizer step.
 *   pg_allreduce_adam_next: the same step, but *d_step and *d_step_id hold the counts BEFORE it: the kernel works with
 *     count + 1 and leaves the incremented counts behind (the last CTA to finish writes them), which takes the two
 *     one-element increment kernels of `optimizer.step()` off the critical path of a captured training step. */
pg_status pg_peer_group_create(int world, int rank, int64_t n, int dev, pg_peer_group** out, unsigned char* handle_out);
pg_status pg_peer_group_connect(pg_peer_group* g, const unsigned char* handles);
void pg_peer_group_destroy(pg_peer_group* g);
pg_status pg_allreduce_adam(pg_peer_group* g, float* d_param, float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                            const float* d_step, const int64_t* d_step_id, float lr, float beta1, float beta2, float eps,
                            float weight_decay, void* stream);
pg_status pg_allreduce_adam_next(pg_peer_group* g, float* d_param, float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                                 float* d_step, int64_t* d_step_id, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, void* stream);

/* ---------------------------------------------------------------- offline partitioner (host code, host pointers)
 * The streaming "dg" assignment of PaGraph/partition/dg.py:59-103, same assignments bit for bit (see pg_partition.cu).
 * indptr / indices: in-neighbour lists (CSC of the row=src, col=dst adjacency). belongs_out: int8[V], partition of every
 * train vertex, -1 elsewhere. member_out: uint8[P*V], member_out[p*V + v] = 1 iff v is in partition p's vertex set (train
 * vertices plus their `hops`-hop in-neighbour redundancy). 2 <= P <= 127. */
pg_status pg_partition_dg(const int64_t* indptr, const int64_t* indices, int64_t V, const int64_t* train,
                          int64_t n_train, int P, int hops, int8_t* belongs_out, uint8_t* member_out);

/* ---------------------------------------------------------------- measurement helpers */
/* Pinned H2D copy bandwidth probe (the PCIe roofline denominator). SYNC. */
pg_status pg_measure_h2d(int dev, size_t bytes, int iters, double* gb_per_s);
/* Live kernel timing for the roofline numbers: while enabled, every kernel class below brackets its
 * launches with a CUDA-event pair on the launching stream. pg_timing_drain (SYNC) returns the
 * records made since the previous drain, in launch order: slots[i] = class, ms[i] = duration. */
enum {
  PG_T_SAMPLE = 0,      /* all kernels of one pg_sample call            */
  PG_T_SPLIT = 1,       /* hit/miss split of one pg_cache_fetch         */
  PG_T_GATHER_HIT = 2,  /* HBM-cache row gather                         */
  PG_T_GATHER_MISS = 3, /* pinned-host row fetch (runs on a side stream) */
  PG_T_AGG_FWD = 4,
  PG_T_AGG_BWD = 5,
  PG_T_FUSED = 6,       /* fused cache-lookup + aggregation (pg_cache_aggregate) */
  PG_T_DENSE_FWD = 7,   /* pg_linear_concat_fwd                          */
  PG_T_DENSE_BWD = 8,   /* pg_linear_concat_bwd (memsets + kernel)       */
  PG_T_HEAD = 9,        /* pg_linear_cross_entropy (memsets + kernel)    */
  PG_T_OPT = 10         /* pg_allreduce_adam                             */
};
pg_status pg_timing_enable(int enabled);
pg_status pg_timing_drain(int32_t* slots, float* ms, int64_t cap, int64_t* n_out);
/* Same drain as a timeline: begin_ms[i] / end_ms[i] = when record i's bracket opened / closed on its stream, relative to
 * the first record of the drain (records of different streams are comparable). SYNC. */
pg_status pg_timing_drain_timeline(int32_t* slots, float* begin_ms, float* end_ms, int64_t cap, int64_t* n_out);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
int64_t pg_launch_count(void);

/* The dropout mask contract of pg_cache_aggregate / pg_aggregate_rows / pg_linear_concat_fwd / _bwd, evaluated on the host
 * (no GPU work): keep_out[j * dim + c] = 1 iff element (row j, column c) is kept for dropout probability p under
 * seed_plus_step = dropout_seed + *d_step. Element (j, c) is decided by the (c % 4)-th 16-bit lane of
 * mix(rowkey ^ colkey), rowkey = splitmix64(splitmix64(seed_plus_step) + j), colkey = splitmix64(0xD1B54A32D192ED03 + c / 4),
 * mix(x) = (x * 0x9E3779B97F4A7C15) ^ ((x * 0x9E3779B97F4A7C15) >> 32); dropped when the lane < round(p * 65536). */
pg_status pg_dropout_keep_mask(uint64_t seed_plus_step, int64_t n_rows, int32_t dim, float p, unsigned char* keep_out);

#ifdef __cplusplus
}
#endif
#endif /* PAGRAPH_B200_H_ */
