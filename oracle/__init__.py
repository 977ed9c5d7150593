"""CPU oracle for the PaGraph hot path — TEST INFRASTRUCTURE, never imported by pagraph_b200/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

Parity status (see pg_oracle.cpp header and DESIGN.md):
  * gather / cache (PaGraph/storage/storage.py:59-227): PINNED by tests/golden/storage_*.npz,
    produced by executing the real reference module (tests/golden/make_golden.py).
  * dg partition scoring (PaGraph/partition/dg.py:59-103): PINNED by tests/golden/dg_*.npz.
  * sampling / NodeFlow / aggregation: arithmetic lives in dgl==0.4.1 (absent) — PARITY UNPINNED;
    restated from its published algorithm.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpg_oracle.so")
_SRC = os.path.join(_HERE, "pg_oracle.cpp")
_lib = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """g++ -O3 -fopenmp the C++ restatement into oracle/libpg_oracle.so."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    cmd = ["g++", "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()                                   # no-op when the library is newer than its source
        _lib = ctypes.CDLL(_SO)
        _lib.pgo_sample.restype = ctypes.c_void_p
        _lib.pgo_sample_stdlib.restype = ctypes.c_void_p
        _lib.pgo_draw.restype = ctypes.c_uint64
        _lib.pgo_sample_batches.restype = ctypes.c_int64
        _lib.pgo_fetch.restype = ctypes.c_int64
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (ctypes.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (ctypes.c_uint32 * 4)()
    lib().pgo_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def draw(seed, epoch, batch, v, hop, t, deg):
    return int(lib().pgo_draw(ctypes.c_uint64(seed), ctypes.c_int64(epoch), ctypes.c_int64(batch),
                              ctypes.c_int64(v), ctypes.c_uint32(hop), ctypes.c_uint32(t),
                              ctypes.c_uint64(deg)))


class OracleNodeFlow:
    """Plain-numpy NodeFlow: layer 0 = inputs ... layer L = seeds."""

    def __init__(self, node_mapping, layer_offsets, indptr, indices, edge_mapping, flow_offsets):
        self.node_mapping = node_mapping
        self.layer_offsets = layer_offsets
        self.indptr = indptr
        self.indices = indices
        self.edge_mapping = edge_mapping
        self.flow_offsets = flow_offsets

    @property
    def num_layers(self):
        return len(self.layer_offsets) - 1

    @property
    def num_blocks(self):
        return self.num_layers - 1

    def layer_parent_nid(self, i):
        i = i % self.num_layers
        return self.node_mapping[self.layer_offsets[i]:self.layer_offsets[i + 1]]

    def block(self, i):
        """(indptr over layer i+1 rows [absolute edge offsets], cols (NodeFlow ids), col_base)."""
        lo, hi = self.layer_offsets[i + 1], self.layer_offsets[i + 2]
        return self.indptr[lo:hi + 1], self.indices, int(self.layer_offsets[i])


def sample(indptr, indices, eids, seeds, fanouts, seed=0, epoch=0, batch=0):
    """DGL-0.4.1 NeighborSampler semantics for one minibatch over an in-CSR (Appendix A.3/A.4)."""
    L = lib()
    indptr, indices, seeds = _c64(indptr), _c64(indices), _c64(seeds)
    eids = _c64(eids) if eids is not None else None
    fan = _c64(fanouts)
    h = L.pgo_sample(_p(indptr, _i64p), _p(indices, _i64p), _p(eids, _i64p),
                     ctypes.c_int64(len(indptr) - 1), _p(seeds, _i64p), ctypes.c_int64(len(seeds)),
                     ctypes.c_int(len(fan)), _p(fan, _i64p), ctypes.c_uint64(seed),
                     ctypes.c_int64(epoch), ctypes.c_int64(batch))
    h = ctypes.c_void_p(h)
    sizes = np.zeros(3, dtype=np.int64)
    L.pgo_nf_sizes(h, _p(sizes, _i64p))
    nl, nn, ne = (int(x) for x in sizes)
    out = [np.zeros(nn, np.int64), np.zeros(nl + 1, np.int64), np.zeros(nn + 1, np.int64),
           np.zeros(ne, np.int64), np.zeros(ne, np.int64), np.zeros(nl, np.int64)]
    L.pgo_nf_copy(h, *[_p(a, _i64p) for a in out])
    L.pgo_nf_free(h)
    return OracleNodeFlow(*out)


def sample_stdlib(indptr, indices, eids, seeds, fanouts, seed=1):
    """The same minibatch drawn the way a single DGL-0.4.1 sampling thread draws it: one std::default_random_engine
    (libstdc++ minstd_rand0) seeded with `seed`, consumed sequentially, layers expanded in discovery order (SURVEY §8c,
    Appendix A.3). Not comparable draw by draw with `sample` (a different generator); used to cross-check its structure."""
    L = lib()
    indptr, indices, seeds = _c64(indptr), _c64(indices), _c64(seeds)
    eids = _c64(eids) if eids is not None else None
    fan = _c64(fanouts)
    h = L.pgo_sample_stdlib(_p(indptr, _i64p), _p(indices, _i64p), _p(eids, _i64p), ctypes.c_int64(len(indptr) - 1),
                            _p(seeds, _i64p), ctypes.c_int64(len(seeds)), ctypes.c_int(len(fan)), _p(fan, _i64p),
                            ctypes.c_uint64(seed))
    return _nf_from_handle(h)


def sample_batches(indptr, indices, eids, seeds, batch_size, first_batch, n_batches, fanouts,
                   seed=0, epoch=0, threads=1):
    """Timed CPU-baseline leg: OpenMP over batches, one thread per batch (DGL's model)."""
    L = lib()
    fan = _c64(fanouts)
    nodes = np.zeros(n_batches, np.int64)
    edges = np.zeros(n_batches, np.int64)
    L.pgo_sample_batches(_p(indptr, _i64p), _p(indices, _i64p), _p(eids, _i64p),
                         ctypes.c_int64(len(indptr) - 1), _p(seeds, _i64p), ctypes.c_int64(len(seeds)),
                         ctypes.c_int64(batch_size), ctypes.c_int64(first_batch),
                         ctypes.c_int64(n_batches), ctypes.c_int(len(fan)), _p(fan, _i64p),
                         ctypes.c_uint64(seed), ctypes.c_int64(epoch), ctypes.c_int(threads),
                         _p(nodes, _i64p), _p(edges, _i64p))
    return nodes, edges


def _nf_from_handle(h):
    L = lib()
    h = ctypes.c_void_p(h)
    sizes = np.zeros(3, dtype=np.int64)
    L.pgo_nf_sizes(h, _p(sizes, _i64p))
    nl, nn, ne = (int(x) for x in sizes)
    out = [np.zeros(nn, np.int64), np.zeros(nl + 1, np.int64), np.zeros(nn + 1, np.int64),
           np.zeros(ne, np.int64), np.zeros(ne, np.int64), np.zeros(nl, np.int64)]
    L.pgo_nf_copy(h, *[_p(a, _i64p) for a in out])
    L.pgo_nf_free(h)
    return OracleNodeFlow(*out)


def sample_many(indptr, indices, eids, seeds, batch_size, first_batch, n_batches, fanouts, seed=0,
                epoch=0, threads=1):
    """`n_batches` consecutive minibatches sampled DGL's way (OpenMP over batches, one thread per
    batch); returns the OracleNodeFlows. Inputs must be contiguous int64 (no copies are made)."""
    L = lib()
    fan = _c64(fanouts)
    handles = (ctypes.c_void_p * n_batches)()
    L.pgo_sample_many(_p(indptr, _i64p), _p(indices, _i64p), _p(eids, _i64p),
                      ctypes.c_int64(len(indptr) - 1), _p(seeds, _i64p), ctypes.c_int64(len(seeds)),
                      ctypes.c_int64(batch_size), ctypes.c_int64(first_batch), ctypes.c_int64(n_batches),
                      ctypes.c_int(len(fan)), _p(fan, _i64p), ctypes.c_uint64(seed), ctypes.c_int64(epoch),
                      ctypes.c_int(threads), handles)
    return [_nf_from_handle(h) for h in handles]


def fetch_c(tnid, flag, l2c, nid_map, cache, host, threads=1):
    """C restatement of storage.py:173-204 for one field (used for timing and large cases)."""
    tnid = _c64(tnid)
    dim = host.shape[1]
    out = np.empty((len(tnid), dim), np.float32)
    mask = np.empty(len(tnid), np.uint8)
    flag8 = np.ascontiguousarray(flag).view(np.uint8)
    if cache is None or cache.shape[0] == 0:
        cache = np.zeros((1, dim), np.float32)
    assert host.strides[1] == 4 and cache.strides[1] == 4
    miss = lib().pgo_fetch(_p(tnid, _i64p), ctypes.c_int64(len(tnid)), _p(flag8, _u8p),
                           _p(_c64(l2c), _i64p), _p(_c64(nid_map), _i64p), _p(cache, _f32p),
                           ctypes.c_int64(cache.strides[0] // 4), _p(host, _f32p),
                           ctypes.c_int64(host.strides[0] // 4), ctypes.c_int64(dim), _p(out, _f32p),
                           _p(mask, _u8p), ctypes.c_int(threads))
    return out, mask.astype(bool), int(miss)


def aggregate(indptr, cols, col_base, src, mode, threads=1, f32=False):
    """copy_src+sum/mean over one block (Appendix A.5), float64 accumulation (the checker) or, with
    f32=True, float32 like DGL's CPU kernel (the timing arm). mode: 'sum' | 'mean'."""
    indptr, cols = _c64(indptr), _c64(cols)
    src = np.ascontiguousarray(src, np.float32)
    n_dst, dim = len(indptr) - 1, src.shape[1]
    dst = np.zeros((n_dst, dim), np.float32)
    fn = lib().pgo_aggregate_f32 if f32 else lib().pgo_aggregate
    fn(_p(indptr, _i64p), _p(cols, _i64p), ctypes.c_int64(col_base), _p(src, _f32p),
                        ctypes.c_int64(n_dst), ctypes.c_int64(dim),
                        ctypes.c_int({"sum": 0, "mean": 1}[mode]), _p(dst, _f32p), ctypes.c_int(threads))
    return dst


def aggregate_bwd(indptr, cols, col_base, grad_dst, n_src, mode):
    indptr, cols = _c64(indptr), _c64(cols)
    grad_dst = np.ascontiguousarray(grad_dst, np.float32)
    n_dst, dim = grad_dst.shape
    out = np.zeros((n_src, dim), np.float32)
    lib().pgo_aggregate_bwd(_p(indptr, _i64p), _p(cols, _i64p), ctypes.c_int64(col_base),
                            _p(grad_dst, _f32p), ctypes.c_int64(n_dst), ctypes.c_int64(n_src),
                            ctypes.c_int64(dim), ctypes.c_int({"sum": 0, "mean": 1}[mode]), _p(out, _f32p))
    return out


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def dropout_keep_mask(seed, n_src, dim, p):
    """Mask contract of the fused dropout (pagraph_b200/csrc/pg_common.cuh drop_hash): element (j, col) is KEPT iff
    the (col % 4)-th 16-bit lane of  mix(rowkey ^ colkey)  is >= round(p * 65536), with
    rowkey = splitmix64(splitmix64(seed) + j), colkey = splitmix64(0xD1B54A32D192ED03 + col // 4),
    mix(x) = (x * 0x9E3779B97F4A7C15) ^ ((x * 0x9E3779B97F4A7C15) >> 32). `seed` = dropout seed + step.
    Returns bool [n_src, dim]."""
    groups = (dim + 3) // 4
    thr = np.uint64(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))
    with np.errstate(over="ignore"):
        stepkey = _splitmix64(np.array([seed & (2 ** 64 - 1)], dtype=np.uint64))[0]
        rk = _splitmix64(stepkey + np.arange(n_src, dtype=np.uint64))[:, None]
        ck = _splitmix64(np.uint64(0xD1B54A32D192ED03) + np.arange(groups, dtype=np.uint64))[None, :]
        x = (rk ^ ck) * np.uint64(0x9E3779B97F4A7C15)
        x = x ^ (x >> np.uint64(32))
    lanes = np.stack([(x >> np.uint64(16 * k)) & np.uint64(0xFFFF) for k in range(4)], axis=-1)   # [n, groups, 4]
    keep = (lanes >= thr).reshape(n_src, groups * 4)[:, :dim]
    return keep


def max_threads():
    return int(lib().pgo_max_threads())


# ----------------------------------------------------------------------------------------------
# numpy restatement of PaGraph/storage/storage.py (the cache object), used by the CPU test-suite.
# ----------------------------------------------------------------------------------------------
class OracleCache:
    """GraphCacheServer restated with numpy (storage.py:18-227); no GPU, no torch."""

    def __init__(self, host_tables, node_num, nid_map):
        self.host = host_tables                       # {name: [V, dim] float32}, full-graph ids
        self.node_num = int(node_num)
        self.nid_map = np.asarray(nid_map, np.int64)  # storage.py:34
        self.gpu_flag = np.zeros(self.node_num, bool)  # storage.py:38
        self.localid2cacheid = np.zeros(self.node_num, np.int64)  # storage.py:50
        self.gpu_fix_cache = {}
        self.dims = {}
        self.total_dim = 0
        self.cached_num = 0
        self.capability = self.node_num
        self.full_cached = False
        self.try_num = 0
        self.miss_num = 0

    def init_field(self, names):  # storage.py:59-67
        self.total_dim = 0
        for n in names:
            self.dims[n] = self.host[n].shape[1]
            self.total_dim += self.dims[n]

    def get_feat_from_server(self, nids, names):  # storage.py:107-132
        full = self.nid_map[nids]
        return {n: self.host[n][full] for n in names}

    def cache_fix_data(self, nids, data, is_full=False):  # storage.py:135-154
        rows = len(nids)
        self.localid2cacheid[nids] = np.arange(rows)
        self.cached_num = rows
        for n in data:
            assert data[n].shape[0] == rows
            self.dims[n] = data[n].shape[1]
            self.gpu_fix_cache[n] = np.array(data[n], np.float32)
        self.gpu_flag[nids] = True
        self.full_cached = is_full

    def auto_cache(self, out_degrees, names, capability):  # storage.py:70-104 with capacity given
        self.capability = int(capability)
        if self.capability >= self.node_num:
            nids = np.arange(self.node_num)
            self.cache_fix_data(nids, self.get_feat_from_server(nids, names), is_full=True)
        else:
            # reference: torch.argsort(descending=True) (tie order unspecified, storage.py:101);
            # the contract here is the stable order (-out_degree, node id).
            order = np.argsort(-np.asarray(out_degrees, np.int64), kind="stable")
            nids = order[:self.capability]
            self.cache_fix_data(nids, self.get_feat_from_server(nids, names), is_full=False)

    def fetch_layer(self, tnid):  # storage.py:176-204 body for one layer
        tnid = np.asarray(tnid, np.int64)
        if self.full_cached:  # storage.py:207-216
            return ({n: self.gpu_fix_cache[n][tnid] for n in self.gpu_fix_cache},
                    np.ones(len(tnid), bool))
        mask = self.gpu_flag[tnid]
        frame = {n: np.empty((len(tnid), self.dims[n]), np.float32) for n in self.dims}
        in_gpu, in_cpu = tnid[mask], tnid[~mask]
        if len(in_gpu):
            cid = self.localid2cacheid[in_gpu]
            for n in self.dims:
                frame[n][mask] = self.gpu_fix_cache[n][cid]
        if len(in_cpu):
            rows = self.get_feat_from_server(in_cpu, list(self.dims))
            for n in self.dims:
                frame[n][~mask] = rows[n]
        self.try_num += len(tnid)
        self.miss_num += len(in_cpu)
        return frame, mask

    def fetch_data(self, nf):
        return [self.fetch_layer(nf.layer_parent_nid(i)) for i in range(nf.num_layers)]
