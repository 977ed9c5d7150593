"""GCNTrainEngine — the training loop of examples/profile/pa_gcn.py:86-97 as a three-stream, CUDA-graph pipeline.

The reference loop is `for nf in sampler: cacher.fetch_data(nf); label = ...; pred = model(nf); loss; backward;
step`. Issued op by op it is launch-bound on a B200 (≈ 40 kernels, ≈ 0.65 ms of GPU work per minibatch at config 2).
Here the same work is three captured graphs per minibatch, replayed on three streams that run one minibatch apart:

  sample graph  (stream A, per ring slot)   pg_sample_keyed [-> label gather]: a chain of small latency-bound kernels
  gather graph  (stream B, per ring slot)   pg_cache_resolve (row pointers of
                                            the input layer + PCIe staging of its missed rows) -> pg_aggregate_rows (fused
                                            cache lookup + dropout + block-0 aggregation; it has no trainable input, so it
                                            need not wait for the previous optimizer step): the HBM / PCIe-bound part
  compute graph (main stream, per slot x bucket)  first NodeUpdate + dropout (pg_linear_concat_fwd, tensor cores) ->
                                            pg_aggregate_fwd_dyn -> head + loss forward/backward (pg_linear_cross_entropy)
                                            -> pg_aggregate_bwd_dyn -> dW / db (pg_linear_concat_bwd) -> gradient
                                            all-reduce + Adam (pg_allreduce_adam): six kernels writing the gradients
                                            straight into the flat bucket when the model is the standard 2-block GCN
                                            with n_hidden = 32 (_compute_body_fused); any other shape runs the same
                                            stage through torch autograd over the ops of pagraph_b200.ops

While minibatch k trains, minibatch k+1 is gathered and aggregated and minibatch k+2 is sampled.
Nothing about a minibatch's size is needed on the host to launch it: every kernel reads the NodeFlow extents from
the device (`meta`), the host only picks the padded-shape bucket of the dense layers from the pinned copy of `meta`
that the sample graph leaves behind two minibatches ahead. The PCIe transfer of minibatch k+1's missed rows runs under
minibatch k's compute. Semantics (what is sampled, fetched, aggregated, and the model math) are those of the eager
classes in this package; tests/test_gpu_engine.py checks the two paths against each other.
"""
import contextlib
import ctypes
import gc
import os
import sys
import time

import numpy as np
import torch

from . import _lib, profiling
from .nodeflow import NodeBatch
from .ops import _MODES, LinearConcat, LinearCrossEntropy, linear_concat_backward, linear_concat_forward
from .parallel import PeerAdam

_RING = 4          # ring slots: compute k | gather k+1 | sampling k+2 .. k+_AHEAD
_RESERVE_SMS = int(os.environ.get("PG_ENGINE_RESERVE_SMS", "24"))   # SMs the input aggregation leaves to the dense stage's small kernels
_AHEAD = 1 + max(1, int(os.environ.get("PG_ENGINE_SAMPLERS", "2")))   # sampling runs this many minibatches ahead of compute
_BUCKET = 4096     # padded-shape granularity of the dense layers


class _BlockAggregateDyn(torch.autograd.Function):
    """copy_src + mean/sum over NodeFlow block `block` with device-resident extents (fixed-shape buffers)."""

    @staticmethod
    def forward(ctx, src, indptr, cols, meta, block, cap_dst, mode):
        ctx.args = (indptr, cols, meta, block, cap_dst, mode, src.shape[0])
        dim = src.shape[1]
        out = torch.empty((cap_dst, dim), dtype=torch.float32, device=src.device)
        lo = ctypes.c_void_p(meta.data_ptr() + 8 * (4 + block))
        _lib.check(_lib.lib().pg_aggregate_fwd_dyn(_lib.ptr(indptr), _lib.ptr(cols), lo, _lib.ptr(src), src.stride(0),
                                                   _lib.ptr(out), out.stride(0), cap_dst, dim, _MODES[mode], None,
                                                   _lib.stream_ptr()), "pg_aggregate_fwd_dyn")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        indptr, cols, meta, block, cap_dst, mode, cap_src = ctx.args
        grad_out = grad_out.contiguous()
        dim = grad_out.shape[1]
        grad_src = torch.empty((cap_src, dim), dtype=torch.float32, device=grad_out.device)
        lo = ctypes.c_void_p(meta.data_ptr() + 8 * (4 + block))
        _lib.check(_lib.lib().pg_aggregate_bwd_dyn(_lib.ptr(indptr), _lib.ptr(cols), lo, _lib.ptr(grad_out),
                                                   grad_out.stride(0), _lib.ptr(grad_src), grad_src.stride(0), cap_dst,
                                                   cap_src, dim, _MODES[mode], None, _lib.stream_ptr()),
                   "pg_aggregate_bwd_dyn")
        return grad_src, None, None, None, None, None, None


class _Slot:
    pass


class _GraphSeq:
    """CUDA graphs replayed in order on one stream"""

    def __init__(self, parts):
        self.parts = list(parts)

    def replay(self):
        for g in self.parts:
            g.replay()


@contextlib.contextmanager
def _capture_guard():
    """No cyclic garbage collection while a stream capture is open: a collected GraphCacheServer / sampler / engine
    would run its destructor (cudaFree, cudaDeviceSynchronize) on the capturing thread and invalidate the capture."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class GCNTrainEngine:
    _first_dense_layer = 1       # lowest NodeFlow layer whose rows enter a dense (padded-shape) kernel

    def __init__(self, g, cacher, model, optimizer, train_nid, labels, batch_size, fanouts, sync=None, seed=0,
                 shuffle=True, host_inputs=False, stage_rows=131072, use_graphs=True, loss_fcn=None):
        """
        g:          pagraph_b200.DGLGraph of the partition; cacher: GraphCacheServer (init_field done; auto_cache may
                    come later — graphs are re-captured when the cache state changes)
        model:      pagraph_b200.model.gcn_nssc.GCNSampling (preprocess=False) on the GPU
        optimizer:  torch optimizer built with capturable=True when use_graphs (e.g. Adam)
        train_nid:  int64 seed vertices (local ids); labels: int64 tensor indexed by local id (CUDA, or CPU when
                    host_inputs — then seeds and labels of every minibatch are copied from pinned host memory)
        sync:       pagraph_b200.parallel.FlatGradAllReduce or None
        """
        self.g, self.cacher, self.model, self.opt, self.sync = g, cacher, model, optimizer, sync
        self.dev = cacher._dev
        self.batch = int(batch_size)
        self.fanouts = [int(f) for f in fanouts]
        self.L = len(self.fanouts)
        self._check_model(model)
        self.seed = int(seed)
        self.host_inputs = bool(host_inputs)
        self.use_graphs = bool(use_graphs)
        self.loss_fcn = loss_fcn or torch.nn.CrossEntropyLoss()
        self.field = "features"
        self.fi = cacher._field_names.index(self.field)
        self.F = cacher.dims[self.field]
        seeds = torch.as_tensor(train_nid, dtype=torch.int64).cpu()
        # The sampler's seed layer drops duplicate seeds (first occurrence wins) and the model's output rows follow the
        # seed LAYER, so labels are gathered in that order: on the device by the sampler itself (pg_sample_keyed's
        # d_seed_labels), on the host (host_inputs) by the same first-occurrence rule when the seed list has duplicates —
        # the reference's own partition files do (isolated train vertices all map to sub-graph id 0, utils.py:47-51).
        self.has_dups = bool(torch.unique(seeds).numel() != seeds.numel())
        if shuffle:                                   # once, like NeighborSampler
            seeds = seeds[torch.randperm(len(seeds))]
        self.num_batches = (len(seeds) + self.batch - 1) // self.batch
        self.n_seeds = len(seeds)
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            # the load stage is a chain of small latency-bound kernels: high priority lets them slip in between the
            # compute stage's full-GPU kernels instead of queueing behind them
            # stream(s) A: sampling. Sampling a minibatch is a dependent chain of small kernels whose time is memory
            # LATENCY (it doubles while the HBM-bound stages run beside it), and minibatches are sampled independently
            # of each other (draws are keyed by (epoch, batch, vertex)): with two sampler states on two streams the
            # chains of consecutive minibatches overlap, so the stage's throughput doubles at the same latency.
            self.n_samplers = _AHEAD - 1
            self.sides = [torch.cuda.Stream(device=self.dev, priority=-1) for _ in range(self.n_samplers)]
            self.side = self.sides[0]
            self.gather = torch.cuda.Stream(device=self.dev)                # stream B: fetch / resolve / aggregate
            if self.host_inputs:
                self.seeds_host = seeds.pin_memory()
                self.labels_host = labels.cpu()
                # without duplicate seeds the seed layer IS the batch, in order: its labels are a slice of labels[seeds]
                self.labels_by_seed_host = None if self.has_dups else self.labels_host[seeds].pin_memory()
            else:
                self.seeds_dev = seeds.to(self.dev)
                self.labels_dev = labels.to(self.dev)
            # capacities: worst case of the sampler (every frontier vertex contributes `fanout` new vertices)
            V, E = g.number_of_nodes(), g.number_of_edges()
            n, self.cap_layer = self.batch, [self.batch]          # sampling order: [0] = seeds
            cap_edges = 0
            for f in self.fanouts:
                e = min(n * min(f, V), E)
                n = min(e, V)
                self.cap_layer.append(n)
                cap_edges += e
            self.cap_nodes, self.cap_edges = sum(self.cap_layer), max(cap_edges, 1)
            self.cap_n0 = self.cap_layer[-1]                          # NodeFlow layer 0 = last sampled layer
            self.cap_rest = sum(self.cap_layer[:-1])                  # NodeFlow layers 1..L
            self._stage_rows_req = int(stage_rows)
            self.stage_rows = 0 if cacher.full_cached else int(min(stage_rows, self.cap_n0))
            fan = (ctypes.c_int64 * self.L)(*self.fanouts)
            self.samplers = []                                 # one sampler state (bitmaps, frontier lists) per stream A
            for _ in range(self.n_samplers):
                h = ctypes.c_void_p()
                _lib.check(L.pg_sampler_create(g.handle(self.dev.index), self.L, fan, self.seed, self.batch, self.cap_nodes,
                                               self.cap_edges, ctypes.byref(h)), "pg_sampler_create")
                self.samplers.append(h)
            self.sampler = self.samplers[0]
            self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)   # optimizer steps taken
            self.load_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)   # minibatches loaded: keys the fused dropout mask
            self.drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self.drop_seed_hidden = int(torch.randint(0, 2 ** 62, (1,)).item())
            self.slots = [self._make_slot() for _ in range(_RING)]
            for j, sl in enumerate(self.slots):               # _RING is a multiple of the sampler count: a slot keeps its
                sl.sampler, sl.side = self.samplers[j % self.n_samplers], self.sides[j % self.n_samplers]   # sampler
        self.fused_opt = None
        if sync is not None and os.environ.get("PG_ENGINE_FUSED_OPT", "1") != "0" and PeerAdam.supported(sync, optimizer):
            try:                                            # PeerAdam agrees on the outcome across ranks before it returns
                self.fused_opt = PeerAdam(sync, optimizer)
            except Exception as e:                          # e.g. CUDA IPC unavailable: NCCL all-reduce + optimizer.step()
                print("GCNTrainEngine: fused all-reduce + Adam unavailable (%s); using all_reduce + optimizer.step()" % e,
                      file=sys.stderr)
        self.serialize = False       # True: every stage runs alone (host sync after each) — per-kernel timing passes
        self._dense_ok = False
        self._block_head = os.environ.get("PG_ENGINE_BLOCK_HEAD", "0") != "0"   # block + head + loss as one kernel
        self._split = False          # compute stage issued in two halves around the next minibatch's input aggregation
        self.dense = None            # buffers of the fused dense stage (_compute_body_fused), made on first use
        self.pool = None
        self.next_issue = 0          # global minibatch index of the next stage A to issue
        self.next_gather = 0         # ... of the next stage B
        self.next_compute = 0
        self.launches = 0            # kernels launched / replayed by this engine (bench.py's gpu_launches)
        self.trace = None            # set to [] to record a per-stage GPU timeline (timing events around every stage)
        self.host_prof = None        # set to {} to accumulate the host's blocked time in steps() (tools/engine_breakdown.py)
        self.size_log = None         # set to [] to record (layer offsets, block offsets) of every computed minibatch
        self._cache_state = None
        self._warm = False

    # ------------------------------------------------------------------ model-specific hooks (overridden by the siblings below)
    def _check_model(self, model):
        if getattr(model, "preprocess", False):
            raise NotImplementedError("GCNTrainEngine drives the non-preprocess model (input block fused from the cache); "
                                      "use GCNPreprocessTrainEngine")
        if len(model.layers) != self.L:
            raise ValueError("model has %d blocks but %d hops are sampled" % (len(model.layers), self.L))

    def _cap_nf(self, l):
        """worst-case size of NodeFlow layer l (0 = inputs .. L = seeds)"""
        return self.cap_layer[self.L - l]

    def _capb(self, l):
        """padded (bucketed) row capacity of the dense buffers of NodeFlow layer l"""
        return self.batch if l == self.L else -(-self._cap_nf(l) // _BUCKET) * _BUCKET

    # ------------------------------------------------------------------ buffers
    def _make_slot(self):
        s, dev = _Slot(), self.dev
        i64 = dict(dtype=torch.int64, device=dev)
        s.seeds_key = torch.zeros(self.batch + 1, **i64)                   # [seeds..., key(2 x uint32)]
        s.stage_host = torch.zeros(self.batch + 1, dtype=torch.int64).pin_memory()
        s.labels = torch.zeros(self.batch, **i64)
        s.labels_host = torch.zeros(self.batch, dtype=torch.int64).pin_memory()
        s.nf = dict(node_mapping=torch.zeros(self.cap_nodes, **i64), indptr=torch.zeros(self.cap_nodes + 1, **i64),
                    indices=torch.zeros(self.cap_edges, **i64), edge_mapping=torch.zeros(self.cap_edges, **i64),
                    meta=torch.zeros(_lib.PG_META_LEN, **i64))
        s.h_meta = torch.zeros(_lib.PG_META_LEN, dtype=torch.int64).pin_memory()
        s.h_meta_np = s.h_meta.numpy()
        self._make_slot_buffers(s)
        s.loss = torch.zeros((), dtype=torch.float32, device=dev)
        s.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()   # read_loss: every step's loss lands here
        s.loss_ready = torch.cuda.Event()
        s.sampled, s.loaded, s.done, s.fwd_done = (torch.cuda.Event() for _ in range(4))
        s.sample_graph, s.sample_kernels = None, 0
        s.gather_graph, s.gather_kernels = None, 0
        s.compute_graphs = {}
        s.n_valid = self.batch
        return s

    def _make_slot_buffers(self, s):
        """model-specific device buffers of one ring slot (row pointers, miss staging, aggregates)"""
        dev = self.dev
        s.rowptr = torch.zeros(self.cap_n0, dtype=torch.int64, device=dev)
        self._make_stage(s, self.stage_rows)
        s.agg = torch.empty((self._capb(1), self.F), dtype=torch.float32, device=dev)   # block-0 aggregate, padded rows zeroed

    def _make_stage(self, s, stage_rows):
        """miss staging rows + miss list of the slot's pg_cache_resolve call(s); re-made when the cache state changes"""
        s.stage = torch.empty((max(stage_rows, 1), self.F), dtype=torch.float32, device=self.dev)
        s.ws = torch.zeros(2 + max(stage_rows, 1), dtype=torch.int64, device=self.dev)   # pg_cache_resolve d_ws

    def _meta_ptr(self, s, idx):
        return ctypes.c_void_p(s.nf["meta"].data_ptr() + 8 * idx)

    # ------------------------------------------------------------------ load stage (side stream)
    def _sample_body(self, s, n_seeds):
        """stage A: sample the minibatch (sizes stay on the device; meta also goes to pinned host for the bucket)."""
        L = _lib.lib()
        nfb = _lib.pg_nodeflow_buffers(*[_lib.ptr(s.nf[k]) for k in
                                         ("node_mapping", "indptr", "indices", "edge_mapping", "meta")])
        key = ctypes.c_void_p(s.seeds_key.data_ptr() + 8 * self.batch)
        lab_in, lab_out = (None, None) if self.host_inputs else (self.labels_dev, s.labels)
        _lib.check(L.pg_sample_keyed(s.sampler, _lib.ptr(s.seeds_key), n_seeds, key, ctypes.byref(nfb), _lib.ptr(s.h_meta),
                                     _lib.ptr(lab_in), _lib.ptr(lab_out), _lib.stream_ptr()), "pg_sample_keyed")

    _GATHER_PARTS = 2    # _gather_body(s, part): 0 = resolve (+ miss fetch), 1 = aggregate; None = both

    def _gather_body(self, s, part=None):
        """stage B: resolve + stage the input layer (part 0), aggregate block 0 (part 1); all sized on the device. The
        frames of layers 1..L (which the reference's fetch_data also fills, storage.py:173-187) have no reader in the
        training step — GCN / GraphSAGE consume the input layer only (gcn_nssc.py:64) — so they are not gathered here.
        The two parts are separate graphs: the lookup (and, with a partial cache, the PCIe fetch of the missed rows) only
        needs the sampled minibatch and runs as soon as it exists; the aggregation is what the split schedule orders
        behind the previous minibatch's NodeUpdate forward."""
        L, c = _lib.lib(), self.cacher
        st = _lib.stream_ptr()
        counts = c._counts if (c.log and not c.full_cached) else None
        blk = _lib.pg_block(_lib.ptr(s.nf["node_mapping"]), _lib.ptr(s.nf["indptr"]), _lib.ptr(s.nf["indices"]), 0,
                            self.cap_n0, s.agg.shape[0], self._meta_ptr(s, 4))
        if part in (None, 0):
            _lib.check(L.pg_cache_resolve(c._handle, self.fi, ctypes.byref(blk), _lib.ptr(s.rowptr), _lib.ptr(s.stage),
                                          self.stage_rows, _lib.ptr(counts), _lib.ptr(s.ws), st), "pg_cache_resolve")
        if part == 0:
            return
        m = self.model
        p = m.dropout.p if (m.dropout is not None and m.training) else 0.0
        # The row-fetching kernel fills every SM it runs on (one CTA, ~200 KB of shared memory), so the classifier head of
        # the PREVIOUS minibatch (93 KB per CTA) could only start when it ends. With the fused dense stage the grid leaves
        # _RESERVE_SMS SMs free: the 64-wide aggregation, the head and its backward run there beside it (measured at
        # config 2, minibatches/s: 0 SMs 3287, 16 3586, 24 3633, 32 3619, 48 3422 — profiles/r2e_*).
        L.pg_set_agg_reserve_sms(_RESERVE_SMS if self._dense_ok else 0)
        try:
            _lib.check(L.pg_aggregate_rows(_lib.ptr(s.rowptr), ctypes.byref(blk), self.F, _lib.ptr(s.agg), s.agg.stride(0),
                                           _MODES["mean"], None, float(p), self.drop_seed, _lib.ptr(self.load_counter),
                                           -_BUCKET, st), "pg_aggregate_rows")
        finally:
            L.pg_set_agg_reserve_sms(0)
        self.load_counter.add_(1)

    def _issue_sample(self, k):
        """Enqueue stage A of global minibatch k."""
        s = self.slots[k % _RING]
        epoch, b = divmod(k, self.num_batches)
        lo = b * self.batch
        n = min(self.batch, self.n_seeds - lo)
        key = (ctypes.c_uint32 * 2)()
        _lib.lib().pg_minibatch_key(self.seed, epoch, b, key)
        keyword = key[0] | (key[1] << 32)
        s.n_valid, s.k = n, k
        side = s.side
        side.wait_event(s.done)                         # the slot's previous minibatch has been consumed
        with torch.cuda.stream(side):
            if self.host_inputs:                             # this minibatch's inputs: pinned host -> device
                s.stage_host[self.batch] = keyword if keyword < 2 ** 63 else keyword - 2 ** 64
                if self.labels_by_seed_host is not None:     # seeds and labels straight from the pinned epoch arrays
                    s.seeds_key[:n].copy_(self.seeds_host[lo:lo + n], non_blocking=True)
                    s.seeds_key[self.batch:].copy_(s.stage_host[self.batch:], non_blocking=True)
                    s.labels[:n].copy_(self.labels_by_seed_host[lo:lo + n], non_blocking=True)
                else:                                        # seed-layer order = first occurrences, in order
                    s.stage_host[:n] = self.seeds_host[lo:lo + n]
                    s.seeds_key.copy_(s.stage_host, non_blocking=True)
                    b = self.seeds_host[lo:lo + n].numpy()
                    first = np.sort(np.unique(b, return_index=True)[1])
                    batch_seeds = torch.from_numpy(b[first])
                    torch.index_select(self.labels_host, 0, batch_seeds, out=s.labels_host[:len(batch_seeds)])
                    s.labels.copy_(s.labels_host, non_blocking=True)
            else:
                s.seeds_key[:n].copy_(self.seeds_dev[lo:lo + n], non_blocking=True)
                s.stage_host[0] = keyword if keyword < 2 ** 63 else keyword - 2 ** 64
                s.seeds_key[self.batch:].copy_(s.stage_host[:1], non_blocking=True)
            self._mark("sample+", k, side)
            if self.use_graphs and n == self.batch:
                if s.sample_graph is None:
                    s.sample_graph, s.sample_kernels = self._capture(side, lambda: self._sample_body(s, self.batch))
                s.sample_graph.replay()
                self.launches += s.sample_kernels
            else:
                l0 = _lib.launch_count()
                self._sample_body(s, n)
                self.launches += _lib.launch_count() - l0
            s.sampled.record(side)
            self._mark("sample-", k, side)
        if self.serialize:
            torch.cuda.synchronize(self.dev)

    def _mark(self, stage, k, stream):
        """stage timeline (tools/engine_breakdown.py): a timing event on `stream`, kept with (stage, minibatch)"""
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            self.trace.append((stage, k, e))

    def _issue_gather(self, k, after=None):
        """Enqueue stage B of global minibatch k (after its stage A, and after the event `after` if given)."""
        s = self.slots[k % _RING]
        self.gather.wait_event(s.sampled)
        parts = tuple(range(self._GATHER_PARTS)) if self._GATHER_PARTS > 1 else (None,)
        with torch.cuda.stream(self.gather):
            self._mark("gather+", k, self.gather)
            if self.use_graphs and s.gather_graph is None:
                caught = [self._capture(self.gather, lambda p=p: self._gather_body(s, p) if p is not None else self._gather_body(s))
                          for p in parts]
                s.gather_graph, s.gather_kernels = _GraphSeq([g for g, _ in caught]), sum(n for _, n in caught)
            l0 = _lib.launch_count()
            for i, p in enumerate(parts):
                if after is not None and i == len(parts) - 1:        # only the last part (the aggregation) is ordered behind `after`
                    self.gather.wait_event(after)
                if self.use_graphs:
                    s.gather_graph.parts[i].replay()
                elif p is None:
                    self._gather_body(s)
                else:
                    self._gather_body(s, p)
            self.launches += s.gather_kernels if self.use_graphs else _lib.launch_count() - l0
            s.loaded.record(self.gather)
            self._mark("gather-", k, self.gather)
        if self.serialize:
            torch.cuda.synchronize(self.dev)

    def _capture(self, stream, body):
        body()                                               # eager once: sizes every workspace outside the capture
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with _capture_guard(), torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
            body()
        return g, _lib.launch_count() - l0

    # ------------------------------------------------------------------ compute stage (main stream)
    def _compute_body(self, s, caps, n_valid, part=None):
        """caps[j]: padded row count of NodeFlow layer j (j = 1..L; caps[L] = batch). `part` (0 / 1) selects one half of
        the fused dense stage when the step is scheduled around the input aggregation (_split)."""
        if self._dense_ok:
            return self._compute_body_fused(s, caps, n_valid, part)
        L, m = _lib.lib(), self.model
        nf = s.nf
        agg = s.agg[:caps[1]]                                 # aggregated by the load stage; rows >= n_1 are zero
        h = m.layers[0](NodeBatch({"h": agg}))["activation"]
        loss = None
        for i in range(1, self.L):
            if m.dropout is not None:
                h = m.dropout(h)
            h = _BlockAggregateDyn.apply(h, nf["indptr"], nf["indices"], nf["meta"], i, caps[i + 1], "mean")
            layer = m.layers[i]
            if i == self.L - 1 and self._head_fusable(layer, h):
                # last NodeUpdate (plain linear) + CrossEntropyLoss, forward and backward, in one kernel
                loss = LinearCrossEntropy.apply(h[:n_valid], layer.linear.weight, layer.linear.bias, s.labels[:n_valid])
            else:
                h = layer(NodeBatch({"h": h}))["activation"]
        if loss is None:
            loss = self.loss_fcn(h[:n_valid], s.labels[:n_valid])
        self._backward_and_step(s, loss)

    def _backward_and_step(self, s, loss):
        """zero_grad -> backward -> gradient all-reduce -> optimizer step (pa_gcn.py:94-97), then the loss into the slot"""
        if self.sync is not None:
            self.sync.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=False)
        loss.backward()
        if self.fused_opt is not None:                       # gradient all-reduce + Adam: one kernel over NVLink peer memory
            self.fused_opt.step(self.step_counter, advance=True)   # ... which also advances both step counters
        else:
            if self.sync is not None:
                self.sync()
            self.opt.step()
            self.step_counter.add_(1)
        s.loss.copy_(loss.detach())

    # ---- fused dense stage: the standard 2-block GCN (NodeUpdate(F -> 32, relu, concat) -> NodeUpdate(64 -> classes) ->
    # CrossEntropyLoss, gcn_nssc.py:43-48 with n_layers = 1) as six of our kernels writing the gradients straight into
    # the parameters' .grad (the flat all-reduce bucket): tensor-core NodeUpdate forward (+ bias, relu, concat, dropout),
    # block-1 aggregation, head + loss forward/backward, aggregation backward, tensor-core dW/db (+ dropout', relu',
    # concat split), all-reduce + Adam. No autograd graph, no elementwise library kernels.
    def _fused_spec(self):
        """What the fused dense stage works on: the first linear (F -> 32) with its activation flags, the classifier head,
        the NodeFlow layer whose rows are its input x (held in s.agg), the block aggregated at width 64 behind it and
        whether dropout follows the first layer. None = this model shape has no fused stage."""
        m = self.model
        if self.L != 2 or len(m.layers) != 2:
            return None
        l0 = m.layers[0]
        return dict(lin=l0.linear, act=l0.activation, concat=l0.concat and not l0.test, head=m.layers[1], xl=1, blk=1,
                    drop_hidden=True)

    def _dense_fusable(self):
        m = self.model
        spec = self._fused_spec() if os.environ.get("PG_ENGINE_FUSED_DENSE", "1") != "0" else None
        if spec is None:
            return False
        relu = spec["act"] is torch.relu or spec["act"] is torch.nn.functional.relu
        w0, w1 = spec["lin"].weight, spec["head"].linear.weight
        agg = self.slots[0].agg
        ok = (relu and spec["concat"] and w0.shape == (32, self.F) and w1.shape[1] == 64
              and LinearConcat.supported(agg, w0) and w0.is_contiguous() and w1.is_contiguous()
              and (w0.grad is None or w0.grad.data_ptr() % 16 == 0)
              and all(p.requires_grad for p in m.parameters())
              and self._head_fusable(spec["head"], torch.empty((1, 64), dtype=torch.float32, device=self.dev)))
        return bool(ok)

    def _dense_buffers(self):
        if self.dense is None:
            d, dev = _Slot(), self.dev
            cap1 = self.slots[0].agg.shape[0]
            f32 = dict(dtype=torch.float32, device=dev)
            d.out = torch.zeros((cap1, 64), **f32)          # cat(z, relu z) of the first layer (pre-dropout)
            d.hd = torch.zeros((cap1, 64), **f32)           # dropout(out): source rows of the 64-wide block
            d.ghd = torch.zeros((cap1, 64), **f32)          # d loss / d hd
            d.a2 = torch.zeros((self.batch, 64), **f32)     # 64-wide aggregate: input of the head
            d.ga2 = torch.zeros((self.batch, 64), **f32)
            for p in self.model.parameters():               # no sync object: plain .grad tensors the kernels overwrite
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            self.dense = d
        return self.dense

    def _compute_body_fused(self, s, caps, n_valid, part=None):
        """part: None = the whole stage; 0 = the first NodeUpdate forward only; 1 = everything after it"""
        L, m, d = _lib.lib(), self.model, self._dense_buffers()
        spec = self._fused_spec()
        nf, st = s.nf, _lib.stream_ptr()
        n1, blk = caps[spec["xl"]], spec["blk"]
        p = float(m.dropout.p) if (spec["drop_hidden"] and m.dropout is not None and m.training) else 0.0
        x, out, hd, ghd = s.agg[:n1], d.out[:n1], d.hd[:n1], d.ghd[:n1]
        w0, b0, w1, b1 = spec["lin"].weight, spec["lin"].bias, spec["head"].linear.weight, spec["head"].linear.bias
        if part in (None, 0):
            linear_concat_forward(x, w0, b0, True, out=out, out_drop=hd, dropout_p=p, seed=self.drop_seed_hidden,
                                  step=self.step_counter)
        if part == 0:
            return
        h = hd if p > 0 else out
        lo = self._meta_ptr(s, 4 + blk)
        C = w1.shape[0]
        if self._block_head:
            # the 64-wide block, the classifier head, the loss and their backward: one kernel
            _lib.check(L.pg_block_linear_cross_entropy(_lib.ptr(nf["indptr"]), _lib.ptr(nf["indices"]), lo, _lib.ptr(h), 64,
                                                       caps[blk + 1], n1, _MODES["mean"], _lib.ptr(w1), _lib.ptr(b1),
                                                       _lib.ptr(s.labels), 64, C, _lib.ptr(s.loss), _lib.ptr(ghd), 64,
                                                       _lib.ptr(w1.grad), _lib.ptr(b1.grad if b1 is not None else None), st),
                       "pg_block_linear_cross_entropy")
        else:
            _lib.check(L.pg_aggregate_fwd_dyn(_lib.ptr(nf["indptr"]), _lib.ptr(nf["indices"]), lo, _lib.ptr(h), 64,
                                              _lib.ptr(d.a2), 64, caps[blk + 1], 64, _MODES["mean"], None, st), "pg_aggregate_fwd_dyn")
            _lib.check(L.pg_linear_cross_entropy(_lib.ptr(d.a2), 64, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(s.labels), n_valid, 64, C,
                                                 _lib.ptr(s.loss), _lib.ptr(d.ga2), 64, _lib.ptr(w1.grad),
                                                 _lib.ptr(b1.grad if b1 is not None else None), self._meta_ptr(s, 4 + self.L), st),
                       "pg_linear_cross_entropy")
            _lib.check(L.pg_aggregate_bwd_dyn(_lib.ptr(nf["indptr"]), _lib.ptr(nf["indices"]), lo, _lib.ptr(d.ga2), 64,
                                              _lib.ptr(ghd), 64, caps[blk + 1], n1, 64, _MODES["mean"], None, st), "pg_aggregate_bwd_dyn")
        linear_concat_backward(x, ghd, out, True, w0.grad, b0.grad if b0 is not None else None, p, self.drop_seed_hidden,
                               self.step_counter)
        if self.fused_opt is not None:                       # gradient all-reduce + Adam: one kernel over NVLink peer memory
            self.fused_opt.step(self.step_counter, advance=True)   # ... which also advances both step counters
        else:
            if self.sync is not None:
                self.sync()
            self.opt.step()
            self.step_counter.add_(1)

    def _head_fusable(self, layer, h):
        lf = self.loss_fcn
        return (isinstance(lf, torch.nn.CrossEntropyLoss) and lf.reduction == "mean" and lf.weight is None
                and lf.label_smoothing == 0.0 and layer.activation is None and not layer.concat and not layer.test
                and LinearCrossEntropy.supported(h, layer.linear.weight))

    def _caps_for(self, s):
        meta = s.h_meta_np
        lay = [int(meta[4 + j + 1] - meta[4 + j]) for j in range(self.L + 1)]       # NodeFlow layer sizes
        caps = [0] * (self.L + 1)
        for j in range(self._first_dense_layer, self.L):
            caps[j] = min(-(-max(lay[j], 1) // _BUCKET) * _BUCKET, self._capb(j))
        caps[self.L] = self.batch
        return tuple(caps), lay

    def _capture_compute(self, s, caps):
        """[graph, ...] of the compute stage (one graph, or the two halves of the split schedule) and its kernel count"""
        graphs, l0 = [], _lib.launch_count()
        for part in ((0, 1) if self._split else (None,)):
            g = torch.cuda.CUDAGraph()
            with _capture_guard(), torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                self._compute_body(s, caps, self.batch, part)
            if self.pool is None:
                self.pool = g.pool()
            graphs.append(g)
        return graphs, _lib.launch_count() - l0

    # ------------------------------------------------------------------ public loop
    def _check_cache_state(self):
        st = (self.cacher.full_cached, self.cacher.cached_num, self.model.training)
        if st != self._cache_state:                          # auto_cache ran (or train/eval flipped): graphs are stale
            torch.cuda.synchronize(self.dev)
            self._cache_state = st
            new_stage = 0 if self.cacher.full_cached else int(min(self._stage_rows_req, self.cap_n0))
            for s in self.slots:
                s.sample_graph, s.gather_graph, s.compute_graphs = None, None, {}
                if new_stage != self.stage_rows:
                    self._make_stage(s, new_stage)
            self.stage_rows = new_stage
            self.pool = None

    def steps(self, count, read_loss=False):
        """Run `count` training minibatches (continuing from the previous call, wrapping over epochs). Returns the
        last loss: a CUDA scalar, or a float when read_loss — then EVERY step's loss is copied to pinned host memory
        behind its compute stage and read by the host (kept in `self.losses`). The host reads step k's loss after it
        has enqueued step k + 1, so the read never drains the pipeline; the call returns once the last one is in."""
        self._check_cache_state()
        self._dense_ok = self._dense_fusable()               # decided (and buffers made) outside any stream capture
        if self._dense_ok:
            self._dense_buffers()
        split = self._dense_ok and not self.serialize and os.environ.get("PG_ENGINE_SPLIT", "1") != "0"
        if split != self._split:                             # the compute graphs were captured for the other schedule
            torch.cuda.synchronize(self.dev)
            for sl in self.slots:
                sl.compute_graphs = {}
            self._split = split
        main = torch.cuda.current_stream(self.dev)
        end = self.next_compute + count
        self.next_gather = max(self.next_gather, self.next_compute)
        self.next_issue = max(self.next_issue, self.next_gather)
        while self.next_issue < min(end, self.next_compute + _AHEAD):   # prologue: A runs _AHEAD ahead, B 1 ahead
            self._issue_sample(self.next_issue)
            self.next_issue += 1
        while self.next_gather < min(end, self.next_compute + 1):
            self._issue_gather(self.next_gather)
            self.next_gather += 1
        loss, pending = None, None
        if read_loss:
            self.losses = []
        while self.next_compute < end:
            k = self.next_compute
            s = self.slots[k % _RING]
            if self.host_prof is not None:                   # host time blocked on the GPU vs. spent issuing work
                t0 = time.perf_counter()
                s.sampled.synchronize()
                self.host_prof["wait_s"] = self.host_prof.get("wait_s", 0.0) + time.perf_counter() - t0
                self.host_prof["steps"] = self.host_prof.get("steps", 0) + 1
            else:
                s.sampled.synchronize()                      # sampling is >= 2 minibatches ahead: normally no wait
            if s.h_meta_np[0] != _lib.PG_OK:
                raise _lib.PGError("sampler reported status %d for minibatch %d" % (int(s.h_meta_np[0]), k))
            main.wait_event(s.loaded)
            caps, lay = self._caps_for(s)
            if self.size_log is not None:
                self.size_log.append(self.layer_sizes(k % _RING))
            n_rows = lay[self.L]                             # seed-layer rows (< n_valid when duplicates were dropped)
            # the fused dense stage reads the row count on the device, so its captured graph serves any batch whose SEED
            # count is full; the autograd body slices on the host and needs the full row count as well
            full = s.n_valid == self.batch and (self._dense_ok or n_rows == self.batch)
            self._mark("compute+", k, main)
            parts = (0, 1) if self._split else (None,)
            graphs = None
            if self.use_graphs and full and self._warm:
                entry = s.compute_graphs.get(caps)
                if entry is None:
                    entry = s.compute_graphs[caps] = self._capture_compute(s, caps)
                graphs = entry[0]
                self.launches += entry[1]
            l0 = _lib.launch_count()
            for i, part in enumerate(parts):
                with profiling.range('gpu-compute'):
                    if graphs is not None:
                        graphs[i].replay()
                    else:
                        self._compute_body(s, caps, n_rows, part)
                if part == 0:
                    # Split schedule: the input aggregation of the NEXT minibatch (every SM's shared memory, HBM-bound)
                    # and the two tensor-core kernels of this one (NodeUpdate forward, dW: one CTA per SM each) cannot
                    # share an SM, so left to the hardware they interleave badly. Ordered explicitly, the aggregation
                    # starts when the forward ends and runs beside the small latency-bound kernels between forward
                    # and dW (64-wide aggregation, head + loss, its backward); dW follows when it drains.
                    s.fwd_done.record(main)
                    if self.next_gather < end:
                        with profiling.range('gpu-load'):
                            self._issue_gather(self.next_gather, after=s.fwd_done)
                        self.next_gather += 1
            if graphs is None:
                self.launches += _lib.launch_count() - l0
                self._warm = True                                # optimizer state exists after the first eager step
            s.done.record(main)
            if read_loss:                                    # D2H copy of the step's result, behind its compute stage
                s.loss_host.copy_(s.loss, non_blocking=True)
                s.loss_ready.record(main)
            self._mark("compute-", k, main)
            if self.serialize:
                torch.cuda.synchronize(self.dev)
            self.next_compute += 1
            with profiling.range('gpu-load'):                # the NEXT minibatches' sampling / cache stages
                if self.next_issue < end:
                    self._issue_sample(self.next_issue)
                    self.next_issue += 1
                if self.next_gather < min(end, self.next_compute + 1):
                    self._issue_gather(self.next_gather)
                    self.next_gather += 1
            loss = s.loss
            if read_loss:
                if pending is not None:                      # the previous step's loss: its successor is already enqueued
                    pending.loss_ready.synchronize()
                    self.losses.append(float(pending.loss_host))
                pending = s
        if pending is not None:
            pending.loss_ready.synchronize()
            self.losses.append(float(pending.loss_host))
            loss = self.losses[-1]
        return loss

    def layer_sizes(self, k_slot):
        """(NodeFlow layer offsets, block offsets) of the minibatch last loaded into ring slot k_slot (host copy)."""
        m = self.slots[k_slot].h_meta_np
        L1 = int(m[3])
        return [int(x) for x in m[4:4 + L1 + 1]], [int(x) for x in m[4 + L1 + 1:4 + L1 + 1 + L1]]

    def close(self):
        torch.cuda.synchronize(self.dev)
        for s in self.slots:
            s.sample_graph, s.gather_graph, s.compute_graphs = None, None, {}
        for h in getattr(self, "samplers", []):
            _lib.lib().pg_sampler_destroy(h)
        self.samplers, self.sampler = [], None
        if self.fused_opt is not None:
            self.fused_opt.close()
            self.fused_opt = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GCNPreprocessTrainEngine(GCNTrainEngine):
    """The `--preprocess` GCN (PaGraph/model/gcn_nssc.py:80-100: the server has already folded the first aggregation into
    the features, server/pa_server.py:45-52) on the same three-stream CUDA-graph pipeline. num_hops = n_layers, so the
    default model is ONE block:
        h = dropout(features[layer 0]);  h = cat(z, relu z), z = linear(h);  mean over block 0 at width 64;  head + loss.
    Gather stage: pg_cache_resolve of the input layer + the fused row kernel over an IDENTITY block (one edge per row), i.e.
    a cache gather with the dropout mask applied on the fly, written once in padded shape. Compute stage: the same fused
    dense kernels as GCNTrainEngine (tcgen05 forward, dW, head + loss, all-reduce + Adam) with the 64-wide aggregation
    behind the first linear instead of a 600-wide one in front of it; other shapes run through autograd."""
    _first_dense_layer = 0

    def _check_model(self, model):
        if not getattr(model, "preprocess", False):
            raise ValueError("GCNPreprocessTrainEngine drives GCNSampling(preprocess=True)")
        if len(model.layers) != self.L:
            raise ValueError("model has %d blocks but %d hops are sampled" % (len(model.layers), self.L))

    def _make_slot_buffers(self, s):
        dev = self.dev
        s.rowptr = torch.zeros(self.cap_n0, dtype=torch.int64, device=dev)
        self._make_stage(s, self.stage_rows)
        s.agg = torch.empty((self._capb(0), self.F), dtype=torch.float32, device=dev)   # dropout(features[layer 0]), padded
        s.idlo = torch.zeros(3, dtype=torch.int64, device=dev)                          # extents of the identity block
        if not hasattr(self, "_ident"):
            self._ident = torch.arange(self.cap_n0 + 1, dtype=torch.int64, device=dev)  # indptr and cols of the identity block

    _GATHER_PARTS = 1

    def _gather_body(self, s):
        L, c = _lib.lib(), self.cacher
        st = _lib.stream_ptr()
        counts = c._counts if (c.log and not c.full_cached) else None
        blk = _lib.pg_block(_lib.ptr(s.nf["node_mapping"]), _lib.ptr(s.nf["indptr"]), _lib.ptr(s.nf["indices"]), 0,
                            self.cap_n0, s.agg.shape[0], self._meta_ptr(s, 4))
        _lib.check(L.pg_cache_resolve(c._handle, self.fi, ctypes.byref(blk), _lib.ptr(s.rowptr), _lib.ptr(s.stage),
                                      self.stage_rows, _lib.ptr(counts), _lib.ptr(s.ws), st), "pg_cache_resolve")
        s.idlo[2:3].copy_(s.nf["meta"][5:6], non_blocking=True)      # [0, 0, n_0]: row r of the block has the one edge r
        ident = _lib.pg_block(None, _lib.ptr(self._ident), _lib.ptr(self._ident), 0, self.cap_n0, s.agg.shape[0], _lib.ptr(s.idlo))
        m = self.model
        p = m.dropout.p if (m.dropout is not None and m.training) else 0.0
        _lib.check(L.pg_aggregate_rows(_lib.ptr(s.rowptr), ctypes.byref(ident), self.F, _lib.ptr(s.agg), s.agg.stride(0),
                                       _MODES["sum"], None, float(p), self.drop_seed, _lib.ptr(self.load_counter), -_BUCKET,
                                       st), "pg_aggregate_rows")
        self.load_counter.add_(1)

    def _fused_spec(self):
        m = self.model
        if self.L != 1 or len(m.layers) != 1 or m.n_layers != 1:
            return None
        return dict(lin=m.linear, act=m.activation, concat=True, head=m.layers[0], xl=0, blk=0, drop_hidden=False)

    def _compute_body(self, s, caps, n_valid, part=None):
        if self._dense_ok:
            return self._compute_body_fused(s, caps, n_valid, part)
        m, nf = self.model, s.nf
        h = m.linear(s.agg[:caps[0]])                        # the dropout is already in s.agg
        h = torch.cat((h, m.activation(h)), dim=1) if m.n_layers == 1 else m.activation(h)
        for i, layer in enumerate(m.layers):
            h = _BlockAggregateDyn.apply(h, nf["indptr"], nf["indices"], nf["meta"], i, caps[i + 1], "mean")
            h = layer(NodeBatch({"h": h}))["activation"]
        self._backward_and_step(s, self.loss_fcn(h[:n_valid], s.labels[:n_valid]))


class SageTrainEngine(GCNTrainEngine):
    """GraphSAGE with mean (or 'gcn' = sum) aggregation — PaGraph/model/graphsage_nssc.py:74-134, trained by
    examples/profile/pa_gs.py:75-100 — on the three-stream CUDA-graph pipeline. With L hops the first NodeUpdate is applied
    to every block (graphsage_nssc.py:96-97), so a minibatch needs, at feature width F,
        neigh_{i+1} = reduce over block i of dropout(features[layer i])      for i = 0 .. L-1   (L aggregations), and
        features[layer l] themselves (the fc_self input)                     for l = 1 .. L.
    Gather stage (HBM / PCIe bound, ours): per block one pg_cache_resolve + the fused cache-lookup + dropout + aggregation
    kernel — the F-wide source rows are never materialised — and per layer l >= 1 one pg_cache_fetch_dyn. Compute stage:
    the model's own NodeUpdate modules (fc_self + fc_neigh, 16 hidden units: library GEMMs over [n_l, F] x [F, 16]) with
    the 2 * n_hidden-wide aggregations on pg_aggregate_fwd_dyn / bwd_dyn, captured per padded-shape bucket."""

    def _check_model(self, model):
        if getattr(model, "preprocess", False):
            raise NotImplementedError("SageTrainEngine drives GraphSageSampling(preprocess=False)")
        if model.aggregator_type not in ("mean", "gcn"):
            raise KeyError("aggregator %r is not on the rebuilt hot path" % model.aggregator_type)
        if len(model.layers) != self.L:
            raise ValueError("model has %d NodeUpdate layers but %d hops are sampled" % (len(model.layers), self.L))
        self.mode = "mean" if model.aggregator_type == "mean" else "sum"

    def _make_slot_buffers(self, s):
        dev, F = self.dev, self.F
        i64 = dict(dtype=torch.int64, device=dev)
        s.rowptrs = [torch.zeros(self._cap_nf(i), **i64) for i in range(self.L)]
        self._make_stage(s, self.stage_rows)
        s.neigh = {i + 1: torch.zeros((self._capb(i + 1), F), dtype=torch.float32, device=dev) for i in range(self.L)}
        s.feat = {l: torch.zeros((self._capb(l), F), dtype=torch.float32, device=dev) for l in range(1, self.L + 1)}
        s.fetch_ws = {l: torch.zeros(2 + 4 * self._cap_nf(l), **i64) for l in range(1, self.L + 1)}
        s.agg = s.neigh[1]                                   # what the base class sizes its (unused) dense buffers by

    def _make_stage(self, s, stage_rows):
        rows = [min(stage_rows, self._cap_nf(i)) for i in range(self.L)]
        s.stages = [torch.empty((max(r, 1), self.F), dtype=torch.float32, device=self.dev) for r in rows]
        s.stage_rows = rows
        s.wss = [torch.zeros(2 + max(r, 1), dtype=torch.int64, device=self.dev) for r in rows]

    def _fused_spec(self):
        return None

    _GATHER_PARTS = 1

    def _gather_body(self, s):
        L, c = _lib.lib(), self.cacher
        st = _lib.stream_ptr()
        counts = c._counts if (c.log and not c.full_cached) else None
        m = self.model
        p = float(m.dropout.p) if m.training else 0.0
        nm, ip, ix = (_lib.ptr(s.nf[k]) for k in ("node_mapping", "indptr", "indices"))
        for i in range(self.L):                              # neigh_{i+1}: block i reduced straight from the cache
            dst = s.neigh[i + 1]
            blk = _lib.pg_block(nm, ip, ix, 0, self._cap_nf(i), dst.shape[0], self._meta_ptr(s, 4 + i))
            _lib.check(L.pg_cache_resolve(c._handle, self.fi, ctypes.byref(blk), _lib.ptr(s.rowptrs[i]), _lib.ptr(s.stages[i]),
                                          s.stage_rows[i], _lib.ptr(counts), _lib.ptr(s.wss[i]), st), "pg_cache_resolve")
            _lib.check(L.pg_aggregate_rows(_lib.ptr(s.rowptrs[i]), ctypes.byref(blk), self.F, _lib.ptr(dst), dst.stride(0),
                                           _MODES[self.mode], None, p, (self.drop_seed + 0x632BE59BD9B4E019 * i) & (2 ** 64 - 1),
                                           _lib.ptr(self.load_counter), -_BUCKET, st), "pg_aggregate_rows")
        for l in range(1, self.L + 1):                       # fc_self inputs: the layers' own rows
            outs = (ctypes.c_void_p * 1)(s.feat[l].data_ptr())
            _lib.check(L.pg_cache_fetch_dyn(c._handle, nm, self._meta_ptr(s, 4 + l), self._meta_ptr(s, 4 + l + 1),
                                            self._cap_nf(l) if l < self.L else self.batch, outs, _lib.ptr(counts), 0,
                                            _lib.ptr(s.fetch_ws[l]), st), "pg_cache_fetch_dyn")
        self.load_counter.add_(1)

    def _compute_body(self, s, caps, n_valid, part=None):
        m, nf = self.model, s.nf
        h = {l: s.feat[l][:caps[l]] for l in range(1, self.L + 1)}
        layer0 = m.layers[0]
        h = {i + 1: layer0(NodeBatch({"h": h[i + 1], "neigh": s.neigh[i + 1][:caps[i + 1]]}))["activation"]
             for i in range(self.L)}
        for lid in range(1, len(m.layers)):
            layer, new = m.layers[lid], {}
            for i in range(lid, self.L):
                hd = m.dropout(h[i])
                neigh = _BlockAggregateDyn.apply(hd, nf["indptr"], nf["indices"], nf["meta"], i, caps[i + 1], self.mode)
                new[i + 1] = layer(NodeBatch({"h": h[i + 1], "neigh": neigh}))["activation"]
            h = new
        self._backward_and_step(s, self.loss_fcn(h[self.L][:n_valid], s.labels[:n_valid]))


def make_train_engine(g, cacher, model, optimizer, train_nid, labels, batch_size, fanouts, **kw):
    """The pipeline engine for `model`: GCNSampling -> GCNTrainEngine / GCNPreprocessTrainEngine, GraphSageSampling ->
    SageTrainEngine."""
    from .model.graphsage_nssc import GraphSageSampling
    if isinstance(model, GraphSageSampling):
        cls = SageTrainEngine
    elif getattr(model, "preprocess", False):
        cls = GCNPreprocessTrainEngine
    else:
        cls = GCNTrainEngine
    return cls(g, cacher, model, optimizer, train_nid, labels, batch_size, fanouts, **kw)
