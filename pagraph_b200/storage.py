"""GraphCacheServer — drop-in for PaGraph/storage/storage.py:18-227 on the B200 path.

Same constructor, methods, attributes and error behaviour as the reference class; the work behind
them is the CUDA library (pg_cache_*): one split kernel + one HBM-cache gather + one TMA fetch of
missed rows straight from the pinned host table, for all NodeFlow layers in one call, with no host
synchronisation (the reference: ~10 torch kernels and >= 3 syncs per layer, CPU gather of misses).
"""
import ctypes
import os
import sys

import torch

from . import _lib, profiling
from .nodeflow import Frame, FrameRef

_REGISTERED = {}   # data_ptr -> nbytes of host tables this process pinned with pg_host_register


def _ensure_device_visible(t):
    """Make a CPU feature table readable by the GPU (page-locked + mapped)."""
    if t.is_pinned():
        return
    key = t.data_ptr()
    nbytes = t.untyped_storage().nbytes() - (t.data_ptr() - t.untyped_storage().data_ptr())
    if key in _REGISTERED and _REGISTERED[key] >= nbytes:
        return
    _lib.check(_lib.lib().pg_host_register(ctypes.c_void_p(key), nbytes), "pg_host_register")
    _REGISTERED[key] = nbytes


class _DevicePtr:
    """raw device allocation exposed through __cuda_array_interface__ so that torch can view it"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _tensor_from_ptr(ptr, shape, dev):
    with torch.cuda.device(dev):
        return torch.as_tensor(_DevicePtr(ptr, shape), device=dev)


class LazyCacheRows:
    """Stand-in for the not-yet-gathered rows of one field of one NodeFlow layer (`lazy_input`).

    `NodeFlow.block_compute` / the models reduce it straight from the cache with the fused kernel
    (pg_cache_aggregate); anything else that needs the tensor calls `materialize()` (also triggered by
    `Frame.__getitem__` through `NodeFlow` layer views), which runs the ordinary gather for that layer.
    """

    def __init__(self, cacher, name, parent_ids):
        self.cacher, self.name, self.parent_ids = cacher, name, parent_ids
        self.shape = (parent_ids.numel(), cacher.dims[name])
        self.dropout_p = 0.0
        self._rows = None

    def with_dropout(self, p):
        """The same rows with (inverted) dropout of probability p folded into the fused aggregation."""
        if self._rows is not None:
            return torch.nn.functional.dropout(self._rows, p, True) if p > 0 else self._rows
        out = LazyCacheRows(self.cacher, self.name, self.parent_ids)
        out.dropout_p = 1.0 - (1.0 - self.dropout_p) * (1.0 - p)
        return out

    def materialize(self):
        if self._rows is None:
            self._rows = self.cacher._gather(self.parent_ids, [self.name])[0]
            if self.dropout_p > 0:
                self._rows = torch.nn.functional.dropout(self._rows, self.dropout_p, True)
        return self._rows


class GraphCacheServer:
    """
    Manage graph features: fetch the feature tensors of a NodeFlow from the GPU cache or from the
    host feature store (reference: PaGraph/storage/storage.py).
    """

    def __init__(self, graph, node_num, nid_map, gpuid):
        """
        graph:    feature store client exposing graph._node_frame._frame[name].data -> CPU tensor
                  [V, dim] indexed by full-graph id (storage.py:128)
        node_num: number of nodes of the local (sub-)graph
        nid_map:  LongTensor[node_num], local id -> full-graph id
        """
        self.graph = graph
        self.gpuid = gpuid
        self.node_num = node_num
        dev = torch.device("cuda", gpuid)
        self._dev = dev
        self.nid_map = nid_map.clone().detach().to(dev, torch.int64).contiguous()
        self.nid_map.requires_grad_(False)
        self.gpu_flag = torch.zeros(self.node_num, dtype=torch.bool, device=dev)
        self.cached_num = 0
        self.capability = node_num
        self.full_cached = False
        self.dims = {}
        self.total_dim = 0
        self.gpu_fix_cache = dict()
        self.localid2cacheid = torch.zeros(node_num, dtype=torch.int64, device=dev)
        self.log = False
        self._counts = torch.zeros(2, dtype=torch.int64, device=dev)   # (tries, misses) on device
        self._try_base = 0
        self._miss_base = 0
        self._handle = None
        self._field_names = []
        self._host_tables = {}
        self.fetch_mode = 0            # 0 auto, 1 plain loads, 2 TMA bulk (pg_cache_fetch `mode`)
        self.last_hit_mask = None      # bool[N] of the most recent fetch_data when keep_hit_mask
        self.keep_hit_mask = False
        # lazy_input (extension): fetch_data leaves the INPUT layer (layer 0) ungathered as LazyCacheRows so
        # the first block's aggregation can read the cache directly; every other layer is gathered as usual.
        self.lazy_input = False
        self._drop_seed = None         # base seed of the fused dropout masks (drawn from torch's RNG on first use)
        self._drop_calls = 0

    # ---- logging counters (storage.py:54-56,219-227); kept on the device, read lazily
    @property
    def try_num(self):
        return self._try_base + int(self._counts[0].item())

    @try_num.setter
    def try_num(self, v):
        self._try_base = v - int(self._counts[0].item())

    @property
    def miss_num(self):
        return self._miss_base + int(self._counts[1].item())

    @miss_num.setter
    def miss_num(self, v):
        self._miss_base = v - int(self._counts[1].item())

    def log_miss_rate(self, miss_num, total_num):
        self._try_base += total_num
        self._miss_base += miss_num

    def get_miss_rate(self):
        tries, misses = self.try_num, self.miss_num
        miss_rate = float(misses) / tries   # ZeroDivisionError when nothing was logged, as the reference
        self._counts.zero_()
        self._try_base = 0
        self._miss_base = 0
        return miss_rate

    # ---- handle
    def _table(self, name):
        return self.graph._node_frame._frame[name].data

    def _make_handle(self, embed_names):
        if self._handle is not None:
            _lib.lib().pg_cache_destroy(self._handle)
            self._handle = None
        if len(embed_names) > _lib.PG_MAX_FIELDS:
            raise ValueError("at most %d feature fields are supported" % _lib.PG_MAX_FIELDS)
        fields = (_lib.pg_field * len(embed_names))()
        self._host_tables = {}
        for i, name in enumerate(embed_names):
            t = self._table(name)
            if t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
                raise TypeError("feature table %r must be a CPU float32 [V, dim] tensor with unit inner stride" % name)
            _ensure_device_visible(t)
            self._host_tables[name] = t
            fields[i].dim = t.shape[1]
            fields[i].host_stride = t.stride(0)
            fields[i].host_table = t.data_ptr()
        h = ctypes.c_void_p()
        flag_u8 = self.gpu_flag.view(torch.uint8)
        _lib.check(_lib.lib().pg_cache_create(self.node_num, _lib.ptr(flag_u8), _lib.ptr(self.localid2cacheid),
                                              _lib.ptr(self.nid_map), len(embed_names), fields, self.gpuid,
                                              ctypes.byref(h)), "pg_cache_create")
        self._handle = h
        self._field_names = list(embed_names)

    def _out_ptrs(self, tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    # ---- reference API
    def init_field(self, embed_names):
        self._make_handle(embed_names)
        self.total_dim = 0
        for name in embed_names:
            self.dims[name] = self._table(name).size(1)
            self.total_dim += self.dims[name]
        print('total dims: {}'.format(self.total_dim), file=sys.stderr)

    def auto_cache(self, dgl_g, embed_names, capability=None):
        """
        Cache node features on the GPU (storage.py:70-104): capacity from the free HBM, then the
        top-`capability` out-degree nodes of the local graph — order (-out_degree, node id).
        `capability=` (extension) overrides the memory-derived capacity.
        """
        if self._field_names != list(embed_names):
            self._make_handle(embed_names)
            self.total_dim = sum(self._table(n).size(1) for n in embed_names)
            for n in embed_names:
                self.dims[n] = self._table(n).size(1)
        peak_allocated_mem = torch.cuda.max_memory_allocated(device=self.gpuid)
        peak_cached_mem = torch.cuda.max_memory_reserved(device=self.gpuid)
        total_mem = torch.cuda.get_device_properties(self.gpuid).total_memory
        available = total_mem - peak_allocated_mem - peak_cached_mem - 1024 * 1024 * 1024
        self.capability = int(available / (self.total_dim * 4)) if capability is None else int(capability)
        print('Cache Memory: {:.2f}G. Capability: {}'.format(available / 1024 / 1024 / 1024, self.capability), file=sys.stderr)
        if self.capability >= self.node_num:
            print('cache the full graph...', file=sys.stderr)
            nids = torch.arange(self.node_num, device=self._dev)
            self._fill(nids, is_full=True)
            self._mark_hot(dgl_g, None)
        else:
            print('cache the part of graph... caching percentage: {:.4f}'.format(self.capability / self.node_num), file=sys.stderr)
            out_degrees = dgl_g.out_degrees().to(self._dev)
            sort_nid = torch.sort(out_degrees, descending=True, stable=True).indices
            self._fill(sort_nid[:max(self.capability, 0)].contiguous(), is_full=False)
            self._mark_hot(dgl_g, sort_nid[:max(self.capability, 0)])

    def _mark_hot(self, dgl_g, cached_by_rank):
        """L2 reuse hint for the fused lookup + aggregation (pg_cache_set_hot; no reference counterpart): the cached rows
        with the highest out-degree — the ranking auto_cache itself uses, storage.py:98-101 — recur as sources within a
        minibatch and across minibatches, so about PG_CACHE_HOT_MB (default 40) of them are fetched with the L2
        evict_last priority while the read-once rows stream through. Results do not depend on it."""
        budget = float(os.environ.get("PG_CACHE_HOT_MB", "40")) * 1e6
        k = int(min(budget // max(self.total_dim * 4, 1), self.node_num if cached_by_rank is None else len(cached_by_rank)))
        if getattr(self, "_hot", None) is None:              # one buffer per cacher: captured graphs keep its address
            self._hot = torch.zeros(self.node_num, dtype=torch.uint8, device=self._dev)
        else:
            self._hot.zero_()
        if k > 0:
            if cached_by_rank is None:
                top = torch.topk(dgl_g.out_degrees().to(self._dev), k, sorted=False).indices
            else:
                top = cached_by_rank[:k]
            self._hot[top] = 1
        _lib.check(_lib.lib().pg_cache_set_hot(self._handle, _lib.ptr(self._hot)), "pg_cache_set_hot")

    # ---- peer-GPU cache tier (extension): the ranks of one node shard the hot rows over NVLink
    def auto_cache_peers(self, dgl_g, embed_names, capability=None, group=None, local_rows=None):
        """auto_cache for `world` ranks that pool their HBM (collective: every rank of `group` calls it).

        Each rank still holds `capability` rows (same memory as auto_cache), but not the same ones: with all vertices of the
        job ranked once by out-degree (full-graph ids, max over the ranks' partition graphs; ties by id), the first
        `local_rows` are replicated on every rank and the next world * (capability - local_rows) are dealt round-robin. A row
        another rank owns is then read from that rank's HBM over NVLink by the fused lookup (pg_cache_resolve /
        GCNTrainEngine), before the pinned host table over PCIe is considered. `local_rows` defaults to the largest value
        for which the pooled caches still cover every vertex (0 if they cannot). gpu_flag / localid2cacheid keep describing
        the LOCAL table, so fetch_data stays correct (it reads rows it does not hold from the host)."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self._field_names != list(embed_names):
            self._make_handle(embed_names)
            self.total_dim = sum(self._table(n).size(1) for n in embed_names)
            for n in embed_names:
                self.dims[n] = self._table(n).size(1)
        if capability is None:
            total_mem = torch.cuda.get_device_properties(self.gpuid).total_memory
            available = (total_mem - torch.cuda.max_memory_allocated(self.gpuid) - torch.cuda.max_memory_reserved(self.gpuid)
                         - 1024 * 1024 * 1024)
            capability = int(available / (self.total_dim * 4))
        cap_t = torch.tensor([int(capability), int(self.nid_map.max().item()) + 1], dtype=torch.int64, device=self._dev)
        if world > 1:
            lo = cap_t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(cap_t, op=dist.ReduceOp.MAX, group=group)
            cap_t[0] = lo[0]
        C, v_full = int(cap_t[0].item()), int(cap_t[1].item())
        if world == 1 or C >= self.node_num:
            return self.auto_cache(dgl_g, embed_names, capability=C)
        rank = dist.get_rank(group)
        self.capability = C
        # one caching order for the whole job, in full-graph ids
        deg_full = torch.zeros(v_full, dtype=torch.int64, device=self._dev)
        deg_full[self.nid_map] = dgl_g.out_degrees().to(self._dev)
        dist.all_reduce(deg_full, op=dist.ReduceOp.MAX, group=group)
        order_full = torch.sort(deg_full, descending=True, stable=True).indices
        del deg_full
        if local_rows is None:
            local_rows = max(0, (world * C - v_full) // (world - 1))
        c_local = int(min(max(local_rows, 0), C))
        c_shard = C - c_local
        mine = torch.cat((order_full[:c_local], order_full[c_local + rank:c_local + world * c_shard:world]))   # full ids, row order
        n_rows = mine.numel()
        # peer-visible tables (cudaMalloc + CUDA IPC), wrapped as torch tensors for the public gpu_fix_cache attribute
        L = _lib.lib()
        self._peer_alloc, self._peer_open = [], []
        ptrs, handles = [], []
        for name in self._field_names:
            p, h = ctypes.c_void_p(), (ctypes.c_ubyte * _lib.PG_IPC_HANDLE_BYTES)()
            _lib.check(L.pg_peer_alloc(max(C, 1) * self.dims[name] * 4, self.gpuid, ctypes.byref(p), h), "pg_peer_alloc")
            self._peer_alloc.append(p)
            ptrs.append(p)
            handles.append(bytes(h))
            self.gpu_fix_cache[name] = _tensor_from_ptr(p.value, (C, self.dims[name]), self._dev)
        with torch.cuda.device(self._dev):
            _lib.check(L.pg_cache_fill_rows(self._handle, _lib.ptr(mine), n_rows, (ctypes.c_void_p * len(ptrs))(*[p.value for p in ptrs]),
                                            _lib.stream_ptr()), "pg_cache_fill_rows")
        # local bookkeeping (storage.py:145,153) for the rows of MY table that my partition names
        full2local = torch.full((v_full,), -1, dtype=torch.int64, device=self._dev)
        full2local[self.nid_map] = torch.arange(self.node_num, device=self._dev)
        loc = full2local[mine]
        has = loc >= 0
        self.localid2cacheid[loc[has]] = torch.arange(n_rows, device=self._dev)[has]
        self.gpu_flag[loc[has]] = True
        self.cached_num = int(has.sum().item())
        self.full_cached = False
        pos_full = torch.empty(v_full, dtype=torch.int32, device=self._dev)
        pos_full[order_full] = torch.arange(v_full, dtype=torch.int32, device=self._dev)
        self._peer_pos = pos_full[self.nid_map].contiguous()
        del pos_full, full2local, order_full
        self._peer_hits = torch.zeros(1, dtype=torch.int64, device=self._dev)
        torch.cuda.synchronize(self._dev)
        # exchange the IPC handles and map the peers' tables
        allh = [None] * world
        dist.all_gather_object(allh, handles, group=group)
        for fi, name in enumerate(self._field_names):
            tabs = (ctypes.c_void_p * world)()
            for r in range(world):
                if r == rank:
                    tabs[r] = ptrs[fi].value
                else:
                    q = ctypes.c_void_p()
                    _lib.check(L.pg_peer_open(allh[r][fi], self.gpuid, ctypes.byref(q)), "pg_peer_open")
                    self._peer_open.append(q)
                    tabs[r] = q.value
            _lib.check(L.pg_cache_set_peers(self._handle, fi, world, rank, tabs, _lib.ptr(self._peer_pos), c_local, c_shard,
                                            _lib.ptr(self._peer_hits)), "pg_cache_set_peers")
        dist.barrier(group=group)       # every table is filled and mapped before anyone reads a peer's rows
        self.peer_tier = dict(world=world, rank=rank, local_rows=c_local, shard_rows=c_shard, covered=c_local + world * c_shard,
                              vertices=v_full)
        print('peer cache tier: {} replicated + {} sharded rows per rank over {} ranks ({:.1f}% of {} vertices pooled)'.format(
            c_local, c_shard, world, 100.0 * min(c_local + world * c_shard, v_full) / v_full, v_full), file=sys.stderr)

    def peer_hits(self, reset=True):
        """rows the fused lookup resolved to another rank's HBM since the last reset (0 without a peer tier)"""
        if getattr(self, "_peer_hits", None) is None:
            return 0
        n = int(self._peer_hits.item())
        if reset:
            self._peer_hits.zero_()
        return n

    def _release_peers(self):
        L = _lib.lib()
        for q in getattr(self, "_peer_open", []):
            L.pg_peer_close(q, self.gpuid)
        for p in getattr(self, "_peer_alloc", []):
            L.pg_peer_free(p, self.gpuid)
        self._peer_open, self._peer_alloc = [], []

    def _fill(self, nids, is_full):
        """cache_fix_data with the rows pulled from the pinned host table by the GPU itself."""
        rows = nids.size(0)
        tables = [torch.empty((rows, self.dims[n]), dtype=torch.float32, device=self._dev) for n in self._field_names]
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().pg_cache_fill(self._handle, _lib.ptr(nids), rows, int(is_full),
                                                self._out_ptrs(tables), 1, _lib.stream_ptr()), "pg_cache_fill")
        for n, t in zip(self._field_names, tables):
            self.gpu_fix_cache[n] = t
        self._keep_nids = nids
        self.cached_num = rows
        self.full_cached = is_full

    def get_feat_from_server(self, nids, embed_names, to_gpu=False):
        """
        Fetch features of local ids `nids` from the host store (storage.py:107-132).
        to_gpu=False returns CPU tensors (host gather, as the reference); to_gpu=True lets the GPU
        read the rows directly from pinned host memory.
        """
        if to_gpu and list(embed_names) == self._field_names:
            nids = nids.to(self._dev).contiguous()
            outs = [torch.empty((nids.numel(), self.dims[n]), dtype=torch.float32, device=self._dev)
                    for n in embed_names]
            with torch.cuda.device(self._dev):
                _lib.check(_lib.lib().pg_cache_fetch_host(self._handle, _lib.ptr(nids), nids.numel(),
                                                          self._out_ptrs(outs), _lib.stream_ptr()),
                           "pg_cache_fetch_host")
            return dict(zip(embed_names, outs))
        nids_in_full = self.nid_map[nids.to(self._dev)].cpu()
        frame = {name: self._table(name)[nids_in_full] for name in embed_names}
        if to_gpu:
            frame = {k: v.to(self._dev, non_blocking=True) for k, v in frame.items()}
        return frame

    def cache_fix_data(self, nids, data, is_full=False):
        """
        Install caller-provided rows as the GPU cache (storage.py:135-154).
        nids: local ids (on the GPU); data: {'field name': tensor [len(nids), dim]}
        """
        rows = nids.size(0)
        nids = nids.to(self._dev).contiguous()
        for name in data:
            data_rows = data[name].size(0)
            assert (rows == data_rows)
            self.dims[name] = data[name].size(1)
            self.gpu_fix_cache[name] = data[name].to(self._dev, torch.float32).contiguous()
        if self._handle is None or self._field_names != list(data):
            self._make_handle(list(data))
        tables = [self.gpu_fix_cache[n] for n in self._field_names]
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().pg_cache_fill(self._handle, _lib.ptr(nids), rows, int(is_full),
                                                self._out_ptrs(tables), 0, _lib.stream_ptr()), "pg_cache_fill")
        self._keep_nids = nids
        self.cached_num = rows
        self.full_cached = is_full

    def fetch_data(self, nodeflow):
        """
        Fill nodeflow._node_frames[i] for every layer from the GPU cache / host store
        (storage.py:157-204; fully cached -> fetch_from_cache). One library call for all layers.
        """
        if self._handle is None:
            raise RuntimeError("fetch_data before init_field")
        with profiling.range('cache-idxload'):
            offsets = nodeflow._layer_offsets
            ids = nodeflow._node_mapping.tousertensor()
            if not ids.is_cuda:
                ids = ids.to(self._dev)
        n = offsets[nodeflow.num_layers]
        first = 0
        if self.lazy_input and nodeflow.num_layers > 1:
            first = 1
            layer0 = ids[offsets[0]:offsets[1]]
            nodeflow._node_frames[0] = FrameRef(Frame({name: LazyCacheRows(self, name, layer0)
                                                       for name in self._field_names}))
        lo0 = offsets[first]
        with profiling.range('cache-fetch'):
            outs = self._gather(ids[lo0:n], self._field_names)
        for i in range(first, nodeflow.num_layers):
            lo, hi = offsets[i] - lo0, offsets[i + 1] - lo0
            frame = {name: out[lo:hi] for name, out in zip(self._field_names, outs)}
            nodeflow._node_frames[i] = FrameRef(Frame(frame))

    def next_dropout_seed(self):
        """A fresh 64-bit seed per fused-dropout call, derived from torch's global RNG state."""
        if self._drop_seed is None:
            self._drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._drop_calls += 1
        return (self._drop_seed + self._drop_calls * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)

    def _gather(self, ids, names, outs=None):
        """rows of `names` (all fields, or a subset -> all fields are gathered and the subset returned) for local
        ids `ids` (CUDA int64), through pg_cache_fetch. `outs`: optional preallocated [>= n, dim] buffers per field."""
        n = ids.numel()
        if outs is None:
            outs = [torch.empty((n, self.dims[name]), dtype=torch.float32, device=self._dev)
                    for name in self._field_names]
        else:
            outs = [o[:n] for o in outs]
        mask = None
        if self.keep_hit_mask:
            mask = torch.empty(n, dtype=torch.bool, device=self._dev)
        counts = self._counts if (self.log and not self.full_cached) else None
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().pg_cache_fetch(self._handle, _lib.ptr(ids), n, self._out_ptrs(outs),
                                                 None if mask is None else ctypes.c_void_p(mask.data_ptr()),
                                                 _lib.ptr(counts), self.fetch_mode, _lib.stream_ptr()),
                       "pg_cache_fetch")
        self.last_hit_mask = mask
        if list(names) == self._field_names:
            return outs
        return [outs[self._field_names.index(nm)] for nm in names]

    def fetch_from_cache(self, nodeflow):
        """Fully-cached fast path (storage.py:207-216)."""
        if not self.full_cached:
            raise RuntimeError("fetch_from_cache requires a fully cached graph")
        self.fetch_data(nodeflow)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().pg_cache_destroy(self._handle)
                self._handle = None
            self._release_peers()
        except Exception:
            pass
