"""GPU NeighborSampler with the signature of `dgl.contrib.sampling.NeighborSampler` (dgl 0.4.1), as the
reference calls it: examples/profile/pa_gcn.py:71-76, PaGraph/partition/utils.py:11-18,
examples/eval.py:20-25. Semantics: SURVEY.md Appendix A.2-A.4; sampled ids are bit-identical to
oracle.sample() under the shared counter-based RNG contract.

Differences that are extensions, not changes: `expand_factor` may also be a per-hop list (index 0
expands the seeds — BASELINE.json's "fanout 25/10"), and `seed=` keys the RNG (the reference's RNG
is unseeded). `num_workers` is accepted and ignored: one GPU replaces the OpenMP worker pool.
`prefetch=True` keeps the next minibatches in flight on a side stream (DGL: a prefetch thread).
"""
import collections
import ctypes

import numpy as np
import torch

from . import _lib
from .nodeflow import NodeFlow

_PREFETCH_DEPTH = 2
_SOFT_CAP = 1 << 28   # start below this many int64 entries per buffer; grow on overflow


class NeighborSampler:
    def __init__(self, g, batch_size, expand_factor=None, num_hops=1, neighbor_type='in',
                 transition_prob=None, seed_nodes=None, shuffle=False, num_workers=1, prefetch=False,
                 add_self_loop=False, seed=0, device=None, device_seeds=True, reuse_buffers=False):
        if not getattr(g, "is_readonly", False):
            raise ValueError("NeighborSampler requires a read-only graph")
        if neighbor_type != 'in':
            raise NotImplementedError("only neighbor_type='in' is on the PaGraph hot path")
        if transition_prob is not None:
            raise NotImplementedError("transition_prob (non-uniform sampling) is not used by the reference")
        if add_self_loop:
            raise NotImplementedError("add_self_loop=True is not used by the reference")
        if num_hops < 1 or num_hops > _lib.PG_MAX_HOPS:
            raise ValueError("num_hops must be in [1, %d]" % _lib.PG_MAX_HOPS)
        self.g = g
        self._dev = torch.cuda.current_device() if device is None else torch.device(device).index
        self._batch_size = int(batch_size)
        V = g.number_of_nodes()
        if expand_factor is None:
            expand_factor = V
        if isinstance(expand_factor, (list, tuple)):
            if len(expand_factor) != num_hops:
                raise ValueError("per-hop expand_factor needs num_hops entries")
            self._fanouts = [int(f) for f in expand_factor]
        else:
            self._fanouts = [int(expand_factor)] * num_hops
        self._num_hops = num_hops
        self._seed = int(seed)
        self._prefetch = bool(prefetch)
        self._epoch = 0
        if seed_nodes is None:
            seed_nodes = torch.arange(V, dtype=torch.int64)
        seed_nodes = torch.as_tensor(np.asarray(seed_nodes) if not torch.is_tensor(seed_nodes) else seed_nodes,
                                     dtype=torch.int64).cpu()
        if shuffle:  # once, at construction (Appendix A.2)
            seed_nodes = seed_nodes[torch.randperm(len(seed_nodes))]
        self._seeds_cpu = seed_nodes.contiguous()
        # device_seeds=True keeps the whole (shuffled) seed list in HBM; False leaves it in pinned host
        # memory and copies each minibatch's slice H2D on the sampling stream (what bench.py's e2e times).
        self._device_seeds = bool(device_seeds)
        # reuse_buffers=True: NodeFlow arrays live in a ring of preallocated buffers (no allocation and no
        # cross-stream allocator bookkeeping per minibatch). A NodeFlow is then valid until `ring` more minibatches
        # have been requested — the training loop's use (one NodeFlow alive at a time); keep the default for
        # callers that hold on to NodeFlows.
        self._reuse = bool(reuse_buffers)
        self._ring, self._ring_i = [], 0
        with torch.cuda.device(self._dev):
            if self._device_seeds:
                self._seeds_dev = self._seeds_cpu.cuda(self._dev)
            else:
                self._seeds_cpu = self._seeds_cpu.pin_memory()
                self._seeds_dev = None
            self._stream = torch.cuda.Stream(device=self._dev)
        self._num_batches = (len(seed_nodes) + self._batch_size - 1) // self._batch_size
        self._graph_handle = g.handle(self._dev)
        self._handle = None
        self._max_seeds = max(1, min(self._batch_size, len(seed_nodes)))
        self._cap_nodes, self._cap_edges = self._initial_caps()
        self._create_handle()
        self._metas = [torch.empty(_lib.PG_META_LEN, dtype=torch.int64).pin_memory()
                       for _ in range(_PREFETCH_DEPTH + 1)]
        self._meta_i = 0

    # ---- capacity management
    def _initial_caps(self):
        V, E = self.g.number_of_nodes(), self.g.number_of_edges()
        n, nodes, edges = self._max_seeds, self._max_seeds, 0
        for f in self._fanouts:
            e = min(n * min(f, V), E)     # a layer holds each vertex once => at most every in-edge
            n = min(e, V)
            nodes += n
            edges += e
        return max(min(nodes, _SOFT_CAP), self._max_seeds), max(min(edges, _SOFT_CAP), 1)

    def _create_handle(self):
        if self._handle is not None:
            _lib.lib().pg_sampler_destroy(self._handle)
            self._handle = None
        h = ctypes.c_void_p()
        fan = (ctypes.c_int64 * self._num_hops)(*self._fanouts)
        _lib.check(_lib.lib().pg_sampler_create(self._graph_handle, self._num_hops, fan, self._seed,
                                                self._max_seeds, self._cap_nodes, self._cap_edges,
                                                ctypes.byref(h)), "pg_sampler_create")
        self._handle = h

    def _alloc_bufs(self, dev):
        return dict(node_mapping=torch.empty(self._cap_nodes, dtype=torch.int64, device=dev),
                    indptr=torch.empty(self._cap_nodes + 1, dtype=torch.int64, device=dev),
                    indices=torch.empty(self._cap_edges, dtype=torch.int64, device=dev),
                    edge_mapping=torch.empty(self._cap_edges, dtype=torch.int64, device=dev),
                    meta=torch.empty(_lib.PG_META_LEN, dtype=torch.int64, device=dev))

    # ---- one minibatch
    def _issue(self, epoch, k):
        lo = k * self._batch_size
        n = min(self._batch_size, len(self._seeds_cpu) - lo)
        dev = "cuda:%d" % self._dev
        if self._reuse:
            # the slot being overwritten belonged to a minibatch whose compute is already enqueued on the
            # caller's stream: order the sampling stream after it
            ev_free = torch.cuda.Event()
            ev_free.record(torch.cuda.current_stream(self._dev))
            self._stream.wait_event(ev_free)
        with torch.cuda.device(self._dev), torch.cuda.stream(self._stream):
            if self._reuse:
                if not self._ring:
                    self._ring = [self._alloc_bufs(dev) for _ in range(_PREFETCH_DEPTH + 2)]
                bufs = dict(self._ring[self._ring_i])
                self._ring_i = (self._ring_i + 1) % len(self._ring)
            else:
                bufs = self._alloc_bufs(dev)
            h_meta = self._metas[self._meta_i]
            self._meta_i = (self._meta_i + 1) % len(self._metas)
            c = _lib.pg_nodeflow_buffers(*[_lib.ptr(bufs[k_]) for k_ in
                                           ("node_mapping", "indptr", "indices", "edge_mapping", "meta")])
            if self._device_seeds:
                seeds_ptr = ctypes.c_void_p(self._seeds_dev.data_ptr() + lo * 8)
            else:
                bufs["seeds"] = self._seeds_cpu[lo:lo + n].to(dev, non_blocking=True)
                seeds_ptr = _lib.ptr(bufs["seeds"])
            _lib.check(_lib.lib().pg_sample(self._handle, seeds_ptr, n, epoch, k, ctypes.byref(c),
                                            _lib.ptr(h_meta), _lib.stream_ptr(self._stream)), "pg_sample")
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return dict(k=k, epoch=epoch, lo=lo, n=n, bufs=bufs, h_meta=h_meta, ev=ev)

    def _finish(self, job):
        job["ev"].synchronize()
        meta = job["h_meta"].numpy()
        if meta[0] == _lib.PG_ERR_OVERFLOW:
            # grow and redo this batch (only the full-fanout closure calls are expected to get here)
            self._stream.synchronize()
            self._cap_nodes = max(int(meta[1] * 1.25) + 16, self._cap_nodes * 2)
            self._cap_edges = max(int(meta[2] * 1.25) + 16, self._cap_edges * 2)
            self._ring, self._ring_i = [], 0
            self._create_handle()
            return self._finish(self._issue(job["epoch"], job["k"]))
        if meta[0] != _lib.PG_OK:
            raise _lib.PGError("pg_sample reported status %d" % meta[0])
        L1 = int(meta[3])
        layer_offsets = meta[4:4 + L1 + 1].tolist()
        flow_offsets = meta[4 + L1 + 1:4 + L1 + 1 + L1].tolist()
        b = job["bufs"]
        cur = torch.cuda.current_stream(self._dev)
        cur.wait_event(job["ev"])
        if not self._reuse:
            for t in b.values():
                t.record_stream(cur)
        return NodeFlow(b["node_mapping"], b["indptr"], b["indices"], b["edge_mapping"], layer_offsets,
                        flow_offsets, seeds_cpu=self._seeds_cpu[job["lo"]:job["lo"] + job["n"]], parent=self.g)

    def sample_batch(self, k, epoch=0):
        """Sample minibatch k of `epoch` synchronously (tests / tools)."""
        return self._finish(self._issue(epoch, k))

    # ---- iteration (one pass over the seeds = one epoch; same seed order every epoch)
    def __len__(self):
        return self._num_batches

    def __iter__(self):
        epoch = self._epoch
        self._epoch += 1
        return self.batches(0, self._num_batches, epoch)

    def batches(self, start, count, epoch=0):
        """Minibatches [start, start+count) of `epoch`, wrapping around the seed list (batch index
        modulo the batches per epoch, epoch advanced on every wrap). With prefetch the next
        minibatches are already being sampled on the side stream while the caller trains on this one;
        exactly `count` minibatches are sampled."""
        depth = _PREFETCH_DEPTH if self._prefetch else 1
        pending = collections.deque()
        nxt, end = start, start + count
        nb = self._num_batches

        def issue(i):
            return self._issue(epoch + i // nb, i % nb)

        while nxt < end and len(pending) < depth:
            pending.append(issue(nxt))
            nxt += 1
        while pending:
            job = pending.popleft()
            nf = self._finish(job)
            if nxt < end:
                pending.append(issue(nxt))
                nxt += 1
            yield nf

    def __del__(self):
        try:
            if self._handle is not None:
                torch.cuda.synchronize(self._dev)
                _lib.lib().pg_sampler_destroy(self._handle)
        except Exception:
            pass
