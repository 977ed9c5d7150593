#!/usr/bin/env python
"""bench.py — minibatches/s and feature-gather GB/s of the PaGraph hot path on B200.

One "step" = one training minibatch of examples/profile/pa_gcn.py:86-97 on BASELINE.json configs[1]
(R-MAT 10 M vertices / 100 M edges, feat 600, 2-layer GCN, fanout 25/10, batch 6000):
    sample (pg_sample_keyed, 7 kernels) -> input layer straight from the cache (pg_cache_resolve + the fused
    lookup + dropout + block-0 mean kernel) -> first NodeUpdate on tcgen05 (pg_linear_concat_fwd) -> block-1 mean
    (pg_aggregate_fwd_dyn) -> classifier head + CrossEntropyLoss forward/backward (pg_linear_cross_entropy) ->
    pg_aggregate_bwd_dyn -> dW/db on tcgen05 (pg_linear_concat_bwd) -> gradient all-reduce + Adam in one kernel
    (pg_allreduce_adam).  --path engine (default) replays this as three CUDA graphs on three streams; --path eager is
    the reference's op sequence (fetch_data of every layer, torch dropout, block_compute, autograd).
Nothing is skipped or cached between steps; every step samples a new minibatch.

    python bench.py [--gpus N] [--steps K] [--warmup W]        # our arm (torchrun for N>1)
    python bench.py --impl reference ...                        # CPU restatement of the reference path

Cache modes ("cache=20 % HBM" is ambiguous, both are measured every run):
    hbm20   capacity = 20 % of the HBM bytes (= what BASELINE.json says literally); the 24 GB table
            fits, so auto_cache takes the reference's full_cached branch (storage.py:90-95).  HEADLINE.
    vtx20   capacity = 20 % of the partition's vertices (the ratio the authors study,
            examples/opt_cache_hit.py:58): real hit/miss split, misses pulled over PCIe by TMA.
            Reported under "variants".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

ID_BYTES = 8


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                   help="BASELINE.json configs[N-1]; 2 (the one the metric is quoted on) is the default and the headline")
    p.add_argument("--scale", type=float, default=1.0, help="shrink a config's graph (vertices and edges) by this factor; "
                                                            "recorded in config.workload — a scaled run is not the config")
    p.add_argument("--vnum", type=int, default=None)
    p.add_argument("--nnz", type=int, default=None)
    p.add_argument("--feat-size", type=int, default=None)
    p.add_argument("--n-classes", type=int, default=None)
    p.add_argument("--n-hidden", type=int, default=None)
    p.add_argument("--batch-size", type=int, default=6000)
    p.add_argument("--fanout", default=None, help="per hop, index 0 expands the seeds")
    p.add_argument("--dropout", type=float, default=0.2)
    p.add_argument("--lr", type=float, default=3e-2)
    p.add_argument("--modes", default=None, help="cache modes to run; the first is the headline (default hbmNN,vtxNN with NN "
                                                 "the config's cache percentage)")
    p.add_argument("--path", default="engine", choices=["engine", "fused", "eager"],
                   help="engine: GCNTrainEngine — two-stream CUDA-graph pipeline over the fused kernels (default); "
                        "fused: the eager Python loop with the input layer aggregated straight from the cache "
                        "(pg_cache_aggregate, dropout folded in); eager: fetch_data gathers every layer, then torch dropout + "
                        "pg_aggregate_fwd (the reference's op sequence)")
    p.add_argument("--kernel-steps", type=int, default=40, help="steps of the instrumented per-kernel timing pass (engine)")
    p.add_argument("--gather-batches", type=int, default=50, help="minibatches of the gather-only measurement")
    p.add_argument("--cpu-batches", type=int, default=32, help="minibatches of the cpu_baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-parity-gate", action="store_true", help="skip the oracle check of one minibatch before timing")
    p.add_argument("--seed", type=int, default=1)
    return apply_config(p.parse_args())


# BASELINE.json `configs`, made concrete. model: gcn = examples/profile/pa_gcn.py, gcn-pre = the same with
# --preprocess (num_hops = n_layers, features folded by the server), sage = examples/profile/pa_gs.py (n_hidden 16).
# partition: hash = hash.py over the train ids (config 1/2: the 2-hop closure of any share is the whole graph, so every
# rank walks the full graph); dg = dg.py assignment + get_sub_graph closure per rank (in-process, on the GPU).
CONFIGS = {
    1: dict(tag="Reddit-shaped", vnum=232_965, nnz=114_615_892, feat=602, classes=41, model="gcn", hidden=32, fanout="2,2",
            partition="hash", parts=1, rmat=(0.25, 0.25, 0.25), cache_frac=0.2),
    2: dict(tag="R-MAT", vnum=10_000_000, nnz=100_000_000, feat=600, classes=60, model="gcn", hidden=32, fanout="25,10",
            partition="hash", parts=0, rmat=(0.45, 0.22, 0.22), cache_frac=0.2),
    3: dict(tag="R-MAT", vnum=10_000_000, nnz=100_000_000, feat=600, classes=60, model="sage", hidden=16, fanout="25,10",
            partition="dg", parts=4, rmat=(0.45, 0.22, 0.22), cache_frac=0.2),
    4: dict(tag="R-MAT", vnum=50_000_000, nnz=500_000_000, feat=600, classes=60, model="gcn-pre", hidden=32, fanout="25",
            partition="dg", parts=8, rmat=(0.45, 0.22, 0.22), cache_frac=0.4),
    5: dict(tag="papers100M-shaped", vnum=111_000_000, nnz=1_600_000_000, feat=128, classes=172, model="sage", hidden=16,
            fanout="15,10,5", partition="dg", parts=8, rmat=(0.45, 0.22, 0.22), cache_frac=0.2, n_layers=2),
}


def apply_config(args):
    c = CONFIGS[args.config]
    args.model, args.partition, args.parts, args.rmat = c["model"], c["partition"], c["parts"], c["rmat"]
    args.cache_frac, args.n_layers, args.tag = c["cache_frac"], c.get("n_layers", 1), c["tag"]
    if args.vnum is None:
        args.vnum = max(1000, int(c["vnum"] * args.scale))
    if args.nnz is None:
        args.nnz = max(2000, int(c["nnz"] * args.scale) // 2 * 2)
    args.feat_size = c["feat"] if args.feat_size is None else args.feat_size
    args.n_classes = c["classes"] if args.n_classes is None else args.n_classes
    args.n_hidden = c["hidden"] if args.n_hidden is None else args.n_hidden
    args.fanout = c["fanout"] if args.fanout is None else args.fanout
    args.fields = ["features"] if args.model == "sage" else ["features", "norm"]
    pct = int(round(args.cache_frac * 100))
    if args.modes is None:
        args.modes = "hbm%d,vtx%d" % (pct, pct)
    return args


# ------------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Polls SM clock / power / throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, dev, period=0.004):
        super().__init__(daemon=True)
        self.period, self.samples, self._stop_ev, self.ok = period, [], threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(dev)
            self.nv = pynvml
            try:
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.device(dev).index or 0)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: the clocks object says so instead of inventing numbers
            self.err = str(e)

    def run(self):
        nv = self.nv
        while not self._stop_ev.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), mhz, reasons, watts))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()

    def sample_now(self):
        """one sample taken by the caller (the driver's default run times ~6 ms: the polling thread may miss it)"""
        if not self.ok:
            return
        try:
            nv = self.nv
            mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.samples.append((time.perf_counter(), mhz, reasons, nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
        except Exception:
            pass

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        s = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples
        bits = 0
        for x in s:
            bits |= x[2]
        return {"sm_mhz": float(np.median([x[1] for x in s])) if s else None, "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(v for k, v in self.REASONS.items() if bits & k),
                "power_w_max": max([x[3] for x in s]) if s else None, "samples": len(s)}


def sampling_bytes(lo, bo):
    """B_S of SURVEY.md §8d for one NodeFlow (layer offsets lo, block offsets bo)."""
    n = [lo[i + 1] - lo[i] for i in range(len(lo) - 1)]
    e = [bo[i + 1] - bo[i] for i in range(len(bo) - 1)]
    b = sum(16 * n[i + 1] + 16 * e[i] for i in range(len(e)))
    return b + 8 * sum(n) + 16 * sum(e) + 8 * sum(x + 1 for x in n)


def agg_bytes(n_src, n_dst, e, d):
    """B_A of SURVEY.md §8d: every source row read once, every dst row written once, CSR once."""
    return 4 * d * (n_src + n_dst) + 8 * e + 8 * (n_dst + 1)


# ------------------------------------------------------------------------------------ workload
class Workload:
    """BASELINE.json configs[args.config - 1] made concrete: synthetic R-MAT graph, U[0,1) features,
    norm = 1/in_degree, random labels, 65 % train split. Partition per rank:
      hash  hash.py split of the train ids; at these densities the 2-hop in-neighbour closure of any 1/N share of the train
            vertices is the whole graph, so every rank's partition graph is the full graph and nid_map is the identity;
      dg    dg.py assignment (pg_partition_dg, rank 0, host) into `parts` partitions + the get_sub_graph closure of THIS
            rank's partition on the GPU: the rank walks its own relabelled sub-graph and nid_map = sub -> full id.
    Features live in ONE pinned host table indexed by full id (the reference's shared-memory store)."""

    def __init__(self, args, rank, world, dev):
        from pagraph_b200 import DGLGraph, data
        from pagraph_b200 import graph_store as gs
        from pagraph_b200.parallel import hash_split
        self.args, self.rank, self.world, self.dev = args, rank, world, dev
        V, Fdim = args.vnum, args.feat_size
        self.setup = {}
        t0 = time.time()
        a_, b_, c_ = args.rmat
        self.indptr, self.indices = data.rmat_in_csr_cuda(V, args.nnz, seed=args.seed, device=dev, a=a_, b=b_, c=c_)
        self.g = DGLGraph.from_in_csr(self.indptr, self.indices)
        self.setup["graph_s"] = round(time.time() - t0, 2)
        t0 = time.time()
        name = "bench%d" % os.getpid() if world == 1 else "bench_w%s" % os.environ.get("MASTER_PORT", "0")
        self.server = None
        norm = None
        if world == 1:
            self.store = gs.LocalGraphStore(name=name)
            feat = self.store.alloc_field("features", V, Fdim)
            norm = self.store.alloc_field("norm", V, 1) if "norm" in args.fields else None
        elif rank == 0:
            self.server = gs.create_graph_store_server(None, name, "shared_mem", world)
            feat = self.server.alloc_field("features", V, Fdim)
            norm = self.server.alloc_field("norm", V, 1) if "norm" in args.fields else None
        if rank == 0:
            gen = torch.Generator(device=dev)
            gen.manual_seed(2)
            chunk = 1 << 18
            deg = (self.indptr[1:] - self.indptr[:-1]).float()
            if args.model == "gcn-pre":
                # server-side --preprocess fold (server/pa_server.py:45-52): features' = (1/in_degree) * sum over in-edges,
                # through pg_aggregate_fwd in row blocks over the whole graph
                from pagraph_b200 import ops
                tp = time.time()
                raw = torch.empty((V, Fdim), dtype=torch.float32, device=dev)
                for lo in range(0, V, chunk):
                    hi = min(V, lo + chunk)
                    raw[lo:hi] = torch.rand((hi - lo, Fdim), device=dev, generator=gen)
                nrm = 1.0 / deg
                for lo in range(0, V, 1 << 20):
                    hi = min(V, lo + (1 << 20))
                    feat[lo:hi].copy_(ops.aggregate_forward(self.indptr[lo:hi + 1], self.indices, 0, raw, hi - lo, "sum",
                                                            norm=nrm[lo:hi]))
                del raw
                torch.cuda.synchronize()
                self.setup["preprocess_fold_s"] = round(time.time() - tp, 2)
            else:
                for lo in range(0, V, chunk):
                    hi = min(V, lo + chunk)
                    feat[lo:hi].copy_(torch.rand((hi - lo, Fdim), device=dev, generator=gen))
            if norm is not None:
                norm.copy_((1.0 / deg).unsqueeze(1))          # inf where in-degree is 0 (pa_server.py:43)
            torch.cuda.synchronize()
            if self.server is not None:
                self.server.commit()
        if world > 1:
            dist.barrier()
            self.store = gs.SharedMemoryStoreClient(name, expect_fields=list(args.fields))
        self.setup["store_s"] = round(time.time() - t0, 2)
        gen = torch.Generator(device=dev)
        gen.manual_seed(3)
        self.labels_dev = torch.randint(0, args.n_classes, (V,), device=dev, generator=gen)
        gen.manual_seed(4)
        perm = torch.randperm(V, device=dev, generator=gen)
        train = torch.sort(perm[:int(V * 0.65)]).values.cpu().numpy()
        del perm
        self.fanouts = [int(x) for x in args.fanout.split(",")]
        self.R = 4 * sum(Fdim if f == "features" else 1 for f in args.fields)
        self.nid_map = None
        if args.partition == "hash":
            parts = args.parts or world
            self.train_nid = np.sort(hash_split(train, parts, seed=5)[rank % parts])
            self.setup["partition"] = {"kind": "hash", "parts": parts, "vertices": int(V), "edges": int(args.nnz),
                                       "train": int(len(self.train_nid))}
        else:
            self._partition_dg(train)
        self.labels_cpu = self.labels_dev.cpu()
        self.V_p = self.g.number_of_nodes()

    def _partition_dg(self, train):
        """dg assignment on rank 0 (host code, scoring over 1-hop in-neighbourhoods: the reference's 2-hop scoring is
        quadratic in hub degrees), broadcast, then this rank's closure on its own GPU."""
        import ctypes
        from pagraph_b200 import DGLGraph, _lib
        from pagraph_b200.partition.utils import get_sub_graph_device
        args, dev, V = self.args, self.dev, self.args.vnum
        P = args.parts
        belongs = torch.empty(V, dtype=torch.int8)
        t0 = time.time()
        if self.rank == 0:
            indptr, indices = self.host_graph()
            member = np.empty((P, V), dtype=np.uint8)
            tr = np.ascontiguousarray(train, np.int64)
            b = belongs.numpy()
            _lib.check(_lib.lib().pg_partition_dg(indptr.ctypes.data, indices.ctypes.data, V, tr.ctypes.data, len(tr), P, 1,
                                                  b.ctypes.data, member.ctypes.data), "pg_partition_dg")
            del member
        t_dg = time.time() - t0
        if self.world > 1:
            bd = belongs.to(dev)
            dist.broadcast(bd, src=0)
            belongs = bd.cpu()
        mine = np.nonzero(belongs.numpy() == (self.rank % P))[0].astype(np.int64)
        t0 = time.time()
        hops = len(self.fanouts)
        ip, ix, sub2full, subtrain = get_sub_graph_device(self.g, mine, hops)
        torch.cuda.synchronize()
        t_cl = time.time() - t0
        self.indptr, self.indices, self._host_graph = ip, ix, None
        self.g = DGLGraph.from_in_csr(ip, ix)                # the rank's own relabelled partition graph
        self.nid_map = sub2full
        self.labels_dev = self.labels_dev[sub2full]          # labels by sub-graph id (pa_gcn.py:38-41)
        self.train_nid = np.sort(subtrain.cpu().numpy())
        torch.cuda.empty_cache()
        self.setup["partition"] = {"kind": "dg", "parts": P, "dg_assign_s": round(t_dg, 2), "closure_s": round(t_cl, 2),
                                   "vertices": int(ip.numel() - 1), "edges": int(ix.numel()), "train": int(len(self.train_nid)),
                                   "scoring_hops": 1, "closure_hops": hops}

    def host_graph(self):
        if getattr(self, "_host_graph", None) is None:
            self._host_graph = (self.indptr.cpu().numpy(), self.indices.cpu().numpy())
        return self._host_graph

    def host_rows(self, name, local_ids):
        """rows of field `name` for local (partition) ids, from the host table (the parity gate's reference)"""
        ids = torch.as_tensor(local_ids, dtype=torch.int64)
        if self.nid_map is not None:
            ids = self.nid_map.cpu()[ids]
        return self.store.ndata[name][ids]

    def close(self):
        if self.world > 1:
            self.store.destroy()
            dist.barrier()
            if self.server is not None:
                self.server.destroy()


class Trainer:
    """The trainer of examples/profile/pa_gcn.py:27-113 (pa_gs.py for GraphSAGE) on the rebuilt path."""

    def __init__(self, wl, mode, host_inputs):
        from pagraph_b200.model.gcn_nssc import GCNSampling
        from pagraph_b200.model.graphsage_nssc import GraphSageSampling
        from pagraph_b200.parallel import FlatGradAllReduce
        from pagraph_b200.sampling import NeighborSampler
        from pagraph_b200.storage import GraphCacheServer
        a = wl.args
        self.wl, self.mode, self.host_inputs = wl, mode, host_inputs
        dev = wl.dev
        V = wl.V_p
        nid_map = wl.nid_map if wl.nid_map is not None else torch.arange(V, dtype=torch.int64)
        self.cacher = GraphCacheServer(wl.store, V, nid_map, dev.index)
        self.cacher.init_field(list(a.fields))
        self.cacher.log = True
        self.cacher.lazy_input = (a.path == "fused")
        torch.manual_seed(wl.rank)                                            # pa_gcn.py:23
        if a.model == "sage":
            self.model = GraphSageSampling(a.feat_size, a.n_hidden, a.n_classes, a.n_layers, F.relu, a.dropout, 'mean').cuda(dev)
        else:
            self.model = GCNSampling(a.feat_size, a.n_hidden, a.n_classes, a.n_layers, F.relu, a.dropout,
                                     a.model == "gcn-pre").cuda(dev)
        self.sync = FlatGradAllReduce(self.model)
        self.opt = torch.optim.Adam(self.sync.flat_parameters(), lr=a.lr, weight_decay=0, capturable=(a.path == "engine"),
                                    fused=True)
        self.loss_fcn = torch.nn.CrossEntropyLoss()
        self.engine = None
        if a.path == "engine":
            from pagraph_b200.engine import make_train_engine
            self.engine = make_train_engine(wl.g, self.cacher, self.model, self.opt, wl.train_nid,
                                            wl.labels_cpu if host_inputs else wl.labels_dev, a.batch_size, wl.fanouts,
                                            sync=self.sync, seed=a.seed, shuffle=True, host_inputs=host_inputs)
        elif a.model != "gcn":
            raise SystemExit("--path %s drives the GCN model only; configs 3-5 need --path engine" % a.path)
        self.sampler = NeighborSampler(wl.g, a.batch_size, wl.fanouts, neighbor_type='in', shuffle=True,
                                       num_workers=16, num_hops=len(wl.fanouts),
                                       seed_nodes=torch.from_numpy(wl.train_nid), prefetch=True, seed=a.seed,
                                       device_seeds=not host_inputs, reuse_buffers=True)
        self.label_stage = [torch.empty(a.batch_size, dtype=torch.int64).pin_memory() for _ in range(4)]
        self.next_batch = 0
        self.sizes = []          # (layer_offsets, block_offsets) of every timed step
        self.fetch_events = []
        # step 1 cold, then fill the cache (pa_gcn.py:99-100)
        self.run(1, record=False)
        total = torch.cuda.get_device_properties(dev).total_memory
        # hbmNN: capacity = NN % of the HBM bytes (BASELINE.json's literal wording); vtxNN: NN % of the partition's vertices
        frac = a.cache_frac
        cap = int(frac * total / (4 * self.cacher.total_dim)) if mode.startswith("hbm") else int(V * frac)
        if mode.startswith("peer"):          # peerNN: vtxNN's capacity per rank, pooled over NVLink
            self.cacher.auto_cache_peers(wl.g, list(a.fields), capability=cap)
        else:
            self.cacher.auto_cache(wl.g, list(a.fields), capability=cap)
        self.cacher.get_miss_rate()
        self.cacher.peer_hits()

    def run(self, count, record, read_loss=False):
        """`count` consecutive training steps. Returns the last loss (host float if read_loss)."""
        wl, dev = self.wl, self.wl.dev
        loss = None
        if self.engine is not None:
            self.engine.size_log = self.sizes if record else None
            return self.engine.steps(count, read_loss=read_loss)
        for nf in self.sampler.batches(self.next_batch, count):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self.cacher.fetch_data(nf)
            if record:
                e1.record()
                self.fetch_events.append((e0, e1))
                self.sizes.append((nf._layer_offsets, nf._block_offsets))
            if self.host_inputs:                     # pa_gcn.py:89-91, through a pinned staging buffer
                batch_nids = nf.layer_parent_nid(-1)
                self._stage_i = (getattr(self, "_stage_i", 0) + 1) % len(self.label_stage)
                st = self.label_stage[self._stage_i][:len(batch_nids)]
                torch.index_select(wl.labels_cpu, 0, batch_nids, out=st)
                label = st.to(dev, non_blocking=True)
            else:
                label = wl.labels_dev[nf.layer_parent_nid_dev(-1)]
            pred = self.model(nf)
            loss = self.loss_fcn(pred, label)
            self.sync.zero_grad()
            loss.backward()
            self.sync()
            self.opt.step()
            if read_loss:
                loss = loss.item()                   # D2H read of the step's result
        self.next_batch += count
        return loss


def timed_region(tr, steps, read_loss, clock, world):
    """barrier + sync | K steps | sync + barrier; CUDA events on the compute stream, max over ranks."""
    from pagraph_b200 import _lib
    tr.sizes, tr.fetch_events = [], []
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tr.cacher.get_miss_rate() if tr.cacher.try_num else None
    tr.cacher.peer_hits()
    _lib.timing_drain()
    _lib.timing_enable(True)
    launches0 = _lib.launch_count()
    eng_launches0 = tr.engine.launches if tr.engine is not None else 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    ev0.record()
    loss = tr.run(steps, record=True, read_loss=read_loss)
    ev1.record()
    if clock is not None:
        clock.sample_now()                   # the GPU is still draining the enqueued steps: a sample under load
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    _lib.timing_enable(False)
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    if tr.engine is not None:
        launches = tr.engine.launches - eng_launches0
    if world > 1:
        t = torch.tensor([ms, (w1 - w0) * 1e3], device=tr.wl.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall_ms = t.tolist()
        lt = torch.tensor([launches], device=tr.wl.dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    else:
        wall_ms = (w1 - w0) * 1e3
    recs = _lib.timing_drain()
    tries, misses = tr.cacher.try_num, tr.cacher.miss_num
    tr.cacher.get_miss_rate() if tries else None
    return dict(ms=ms, wall_ms=wall_ms, launches=launches, recs=recs, tries=tries, misses=misses, peer_hits=tr.cacher.peer_hits(),
                loss=float(loss), clocks=clock.summary(w0, w1) if clock else None,
                fetch_ms=sum(a.elapsed_time(b) for a, b in tr.fetch_events))


def kernel_report(tr, reg, hbm_peak, pcie_peak):
    """Per-kernel-class live timing (pg_timing_*: CUDA events on the launching stream, inside the timed region)
    against the algorithmic bytes of SURVEY.md §8d."""
    from pagraph_b200 import _lib
    wl = tr.wl
    a = wl.args
    R, Fdim, H2 = wl.R, a.feat_size, 2 * a.n_hidden
    L = len(wl.fanouts)
    by = {}
    for slot, ms in reg["recs"]:
        by.setdefault(slot, []).append(ms)
    steps = len(tr.sizes)
    N = sum(lo[-1] for lo, _ in tr.sizes)
    M = reg["misses"]
    fused = _lib.T_FUSED in by
    if not by:
        return {}, N, M
    out = {}

    def n_(lo, l):
        return lo[l + 1] - lo[l]

    def e_(bo, i):
        return bo[i + 1] - bo[i]

    def add(name, slot_ms, nbytes, peak, peak_name):
        if not slot_ms:
            return
        t = sum(slot_ms)
        out[name] = {"launches": len(slot_ms), "avg_ms": t / len(slot_ms), "alg_bytes_per_launch": nbytes / len(slot_ms),
                     "achieved_gbs": nbytes / t / 1e6, "peak_gbs": peak, "frac": nbytes / t / 1e6 / peak, "bound": peak_name}

    add("sample(all kernels of one pg_sample)", by.get(_lib.T_SAMPLE), sum(sampling_bytes(lo, bo) for lo, bo in tr.sizes),
        hbm_peak, "hbm")
    # rows the step looked up in the cache: gcn / gcn-pre the input layer, sage layers 0..L-1 (sources) + 1..L (self rows)
    if a.model == "sage":
        looked = sum(lo[L] + (lo[-1] - lo[1]) for lo, _ in tr.sizes)
    elif tr.engine is not None or fused:
        looked = sum(n_(lo, 0) for lo, _ in tr.sizes)
    else:
        looked = N
    add("split_kernel/resolve_kernel", by.get(_lib.T_SPLIT), 25 * looked, hbm_peak, "hbm")
    if a.model == "sage":
        rows_g = sum(lo[-1] - lo[1] for lo, _ in tr.sizes)
    else:
        rows_g = 0 if (tr.engine is not None) else (N - sum(n_(lo, 0) for lo, _ in tr.sizes) if fused else N)
    miss_share = M / max(looked, 1)
    add("gather_hit(rows_ldg_kernel)", by.get(_lib.T_GATHER_HIT), (2 * R + ID_BYTES) * rows_g * (1 - miss_share), hbm_peak, "hbm")
    add("gather_miss(rows_bulk_kernel)", by.get(_lib.T_GATHER_MISS), 4 * Fdim * M, pcie_peak, "pcie")
    fw = by.get(_lib.T_AGG_FWD, [])
    if a.model == "sage":       # L fused launches per step at width F; the 2*hidden-wide blocks behind them
        bF = sum(sum(agg_bytes(n_(lo, i), n_(lo, i + 1), e_(bo, i), Fdim) + 8 * n_(lo, i) for i in range(L)) for lo, bo in tr.sizes)
        add("cache_aggregate_blocks0..%d(agg_rows_tma_kernel,D=%d)" % (L - 1, Fdim), by.get(_lib.T_FUSED), bF, hbm_peak, "hbm")
        bh = sum(sum(agg_bytes(n_(lo, i), n_(lo, i + 1), e_(bo, i), H2) for i in range(1, L)) for lo, bo in tr.sizes)
        add("agg_fwd_hidden_blocks(D=%d)" % H2, fw, bh, hbm_peak, "hbm")
        add("agg_bwd_hidden_blocks(D=%d)" % H2, by.get(_lib.T_AGG_BWD), bh, hbm_peak, "hbm")
    elif a.model == "gcn-pre":  # the fused launch is the cache gather (+ dropout) of the input layer; one 64-wide block
        bg = sum((2 * 4 * Fdim + 8 + 16) * n_(lo, 0) for lo, _ in tr.sizes)
        add("cache_gather_dropout_layer0(agg_rows_tma_kernel,D=%d)" % Fdim, by.get(_lib.T_FUSED), bg, hbm_peak, "hbm")
        b1 = sum(agg_bytes(n_(lo, 0), n_(lo, 1), e_(bo, 0), H2) for lo, bo in tr.sizes)
        add("agg_fwd_block0(D=%d)" % H2, fw, b1, hbm_peak, "hbm")
        add("agg_bwd_block0(D=%d)" % H2, by.get(_lib.T_AGG_BWD), b1, hbm_peak, "hbm")
    else:
        b0 = sum(agg_bytes(n_(lo, 0), n_(lo, 1), e_(bo, 0), Fdim) for lo, bo in tr.sizes)
        b1 = sum(agg_bytes(n_(lo, 1), n_(lo, 2), e_(bo, 1), H2) for lo, bo in tr.sizes)
        n0 = sum(n_(lo, 0) for lo, _ in tr.sizes)
        if fused:
            add("cache_aggregate_block0(agg_rows_tma_kernel,D=%d)" % Fdim, by.get(_lib.T_FUSED), b0 + 8 * n0, hbm_peak, "hbm")
            add("agg_fwd_block1(D=%d)" % H2, fw, b1, hbm_peak, "hbm")
        elif len(fw) == 2 * steps:
            add("agg_fwd_block0(D=%d)" % Fdim, fw[0::2], b0, hbm_peak, "hbm")
            add("agg_fwd_block1(D=%d)" % H2, fw[1::2], b1, hbm_peak, "hbm")
        add("agg_bwd_block1(D=%d)" % H2, by.get(_lib.T_AGG_BWD), b1, hbm_peak, "hbm")
    # dense stage (the engines' fused path): x [n_x, F] streamed once per direction; out / out_drop / grad rows are H2 wide
    xl = 0 if a.model == "gcn-pre" else 1
    n1 = sum(n_(lo, xl) for lo, _ in tr.sizes)
    nb = sum(n_(lo, L) for lo, _ in tr.sizes)
    drop = 1 if (a.dropout > 0 and a.model == "gcn") else 0
    wbytes = 4 * steps * (Fdim + 1) * a.n_hidden
    add("node_update_fwd(linear_concat_fwd_umma_kernel,tcgen05 3xTF32)", by.get(_lib.T_DENSE_FWD),
        4 * n1 * (Fdim + H2 * (1 + drop)) + wbytes, hbm_peak, "hbm")
    add("node_update_bwd(linear_concat_dw_umma_kernel,tcgen05 3xTF32)", by.get(_lib.T_DENSE_BWD),
        4 * n1 * (Fdim + 2 * H2) + wbytes, hbm_peak, "hbm")
    head_bytes = 4 * nb * 2 * H2 + 8 * nb + 8 * steps * (H2 + 1) * a.n_classes
    if tr.engine is not None and tr.engine._dense_ok and tr.engine._block_head:
        # the 64-wide block, the head, the loss and their backward in one kernel: block read + gradient scatter + head
        bl = sum(agg_bytes(n_(lo, L - 1), n_(lo, L), e_(bo, L - 1), H2) for lo, bo in tr.sizes)
        add("block%d+head+loss fwd/bwd(linear_ce_mma_kernel<BLOCK>)" % (L - 1), by.get(_lib.T_HEAD), 2 * bl + head_bytes,
            hbm_peak, "hbm")
    else:
        add("head+loss(linear_ce_mma_kernel)", by.get(_lib.T_HEAD), head_bytes, hbm_peak, "hbm")
    nparam = sum(p.numel() for p in tr.model.parameters())
    add("allreduce+adam(allreduce_adam_kernel)", by.get(_lib.T_OPT), 4 * 7 * nparam * steps, hbm_peak, "hbm")
    return out, N, M


def gather_only(tr, n_batches, hbm_peak, pcie_peak):
    """Feature-gather GB/s with the hit/miss split: pg_cache_fetch of EVERY NodeFlow layer (what the reference's
    fetch_data does, storage.py:157-204), on fresh minibatches, no training work in between."""
    from pagraph_b200 import _lib
    wl, c = tr.wl, tr.cacher
    R = wl.R
    torch.cuda.synchronize()
    if c.try_num:
        c.get_miss_rate()
    cap = tr.sampler._cap_nodes                          # output buffers allocated once (the kernels are what is measured)
    bufs = [torch.empty((cap, c.dims[name]), dtype=torch.float32, device=wl.dev) for name in c._field_names]
    for nf in tr.sampler.batches(tr.next_batch, 3):      # untimed warm-up
        c._gather(nf._node_mapping.tousertensor(), c._field_names, outs=bufs)
    tr.next_batch += 3
    torch.cuda.synchronize()
    if c.try_num:
        c.get_miss_rate()
    _lib.timing_drain()
    _lib.timing_enable(True)
    evs, N = [], 0
    for nf in tr.sampler.batches(tr.next_batch, n_batches):
        ids = nf._node_mapping.tousertensor()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        c._gather(ids, c._field_names, outs=bufs)
        b.record()
        evs.append((a, b))
        N += ids.numel()
    tr.next_batch += n_batches
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    by = {}
    for slot, ms in _lib.timing_drain():
        by.setdefault(slot, []).append(ms)
    M = c.miss_num if not c.full_cached else 0
    if c.try_num:
        c.get_miss_rate()
    t = sum(a.elapsed_time(b) for a, b in evs)
    t_hit, t_miss = sum(by.get(_lib.T_GATHER_HIT, [0])), sum(by.get(_lib.T_GATHER_MISS, [0]))
    H = N - M
    out = {"batches": n_batches, "rows_per_batch": N / n_batches, "hit_rate": H / max(N, 1),
           "gather_gbs": R * N / t / 1e6, "ms_per_batch": t / n_batches,
           "hit": {"kernel": "rows_ldg_kernel", "payload_gbs": R * H / max(t_hit, 1e-9) / 1e6, "avg_ms": t_hit / n_batches,
                   "alg_bytes_per_launch": (2 * R * H + ID_BYTES * H) / n_batches,
                   "achieved_gbs": (2 * R * H + ID_BYTES * H) / max(t_hit, 1e-9) / 1e6, "peak_gbs": hbm_peak,
                   "frac": (2 * R * H + ID_BYTES * H) / max(t_hit, 1e-9) / 1e6 / hbm_peak, "bound": "hbm"}}
    if M:
        out["miss"] = {"kernel": "rows_bulk_kernel", "payload_gbs": R * M / t_miss / 1e6, "avg_ms": t_miss / n_batches,
                       "alg_bytes_per_launch": R * M / n_batches, "achieved_gbs": R * M / t_miss / 1e6,
                       "peak_gbs": pcie_peak, "frac": R * M / t_miss / 1e6 / pcie_peak, "bound": "pcie"}
    return out


_T0 = time.time()


def note(msg):
    """progress on stderr (rank 0): where the time of a run goes, and where a run that hangs stopped"""
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench %7.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


def run_mode(wl, mode, args, world, hbm_peak, pcie_peak):
    """value (device-resident inputs) and e2e (host inputs + loss read back) of one cache mode."""
    res = {}
    for host_inputs in (False, True):
        note("mode %s, host inputs %s: trainer" % (mode, host_inputs))
        tr = Trainer(wl, mode, host_inputs)
        gate = None
        if not host_inputs and not args.no_parity_gate:
            # rank 0 checks one minibatch of ITS partition against the oracle; every rank learns the outcome
            err = ""
            if wl.rank == 0:
                try:
                    gate = parity_gate(wl, tr)
                except SystemExit as ex:
                    err = str(ex)
            if world > 1:
                flag = torch.tensor([1 if err else 0], device=wl.dev)
                dist.all_reduce(flag)
                if flag.item() and not err:
                    err = "parity gate failed on rank 0"
            if err:
                raise SystemExit(err)
        note("warm-up")
        tr.run(args.warmup, record=False, read_loss=host_inputs)
        note("timed region")
        clock = None
        if wl.rank == 0:
            clock = ClockSampler(wl.dev)
            if clock.ok:
                clock.start()
        prof = os.environ.get("PG_BENCH_CUDA_PROFILER") == "1" and not host_inputs
        if prof:                                  # ncu --profile-from-start off: capture the timed region only
            torch.cuda.profiler.start()
        reg = timed_region(tr, args.steps, host_inputs, clock, world)
        if prof:
            torch.cuda.profiler.stop()
        note("timed region done: %.4f ms/step" % (reg["ms"] / args.steps))
        replicas = replica_check(tr, world)
        if not replicas:
            raise SystemExit("replica check: ranks hold different parameters after %d steps" % args.steps)
        if clock is not None and clock.ok:
            clock.stop()
        if tr.engine is not None:
            # per-kernel CUDA-event timing needs host-side launches: same pipeline, same kernels, graphs off. Two passes:
            # stages serialised (every kernel alone on the GPU: the roofline numbers, comparable with ncu) and the
            # three-stream pipeline as it runs (avg_ms_in_pipeline: what contention between the stages adds)
            steps_total, misses_total, sizes_total = len(tr.sizes), reg["misses"], tr.sizes
            ksteps = min(args.kernel_steps, args.steps)
            note("per-kernel passes")
            tr.engine.use_graphs = False
            kreg = timed_region(tr, ksteps, host_inputs, None, world)
            kern_pipe, _, _ = kernel_report(tr, kreg, hbm_peak, pcie_peak)
            tr.engine.serialize = True
            areg = timed_region(tr, ksteps, host_inputs, None, world)
            kern, _, _ = kernel_report(tr, areg, hbm_peak, pcie_peak)
            # The engine leaves some SMs out of the input aggregation's grid (the step gets faster, that kernel slower):
            # one more serialised pass with the full grid, so that both figures of the dominant kernel are on record.
            import pagraph_b200.engine as _eng
            if getattr(tr.engine, "_dense_ok", False) and _eng._RESERVE_SMS > 0 and not host_inputs:
                keep, _eng._RESERVE_SMS = _eng._RESERVE_SMS, 0
                try:
                    freg = timed_region(tr, ksteps, host_inputs, None, world)
                    kfull, _, _ = kernel_report(tr, freg, hbm_peak, pcie_peak)
                finally:
                    _eng._RESERVE_SMS = keep
                for name, k in kern.items():
                    if name.startswith("cache_aggregate") and name in kfull:
                        k["avg_ms_all_sms"], k["frac_all_sms"] = kfull[name]["avg_ms"], kfull[name]["frac"]
                        k["sms_left_out"] = keep
            tr.engine.serialize = False
            tr.engine.use_graphs = True
            for name, k in kern.items():
                if name in kern_pipe:
                    k["avg_ms_in_pipeline"] = kern_pipe[name]["avg_ms"]
            tr.sizes = sizes_total
            N, M = sum(lo[-1] for lo, _ in tr.sizes), misses_total
        else:
            kern, N, M = kernel_report(tr, reg, hbm_peak, pcie_peak)
        note("gather-only pass")
        gather = gather_only(tr, args.gather_batches, hbm_peak, pcie_peak) if not host_inputs else None
        steps = args.steps
        mbps = world * steps / (reg["ms"] * 1e-3)
        r = dict(minibatches_per_s=mbps, ms_per_step=reg["ms"] / steps, wall_ms_per_step=reg["wall_ms"] / steps,
                 rows_per_step=N / steps, miss_rows_per_step=M / steps,
                 hit_rate=(1.0 - reg["misses"] / reg["tries"]) if reg["tries"] else 1.0,   # over the rows the step looked up
                 fetch_ms_per_step=reg["fetch_ms"] / steps, gather=gather,
                 launches=reg["launches"], loss=reg["loss"], clocks=reg["clocks"], kernels=kern,
                 full_cached=tr.cacher.full_cached, cached_rows=tr.cacher.cached_num, parity_gate=gate,
                 peer_tier=getattr(tr.cacher, "peer_tier", None), peer_rows_per_step=reg["peer_hits"] / steps,
                 nvlink_bytes_per_step=4 * args.feat_size * reg["peer_hits"] / steps,
                 pcie_bytes_per_step=4 * args.feat_size * reg["misses"] / steps,
                 replicas_identical=replicas,
                 layer_sizes=[int(np.mean([lo[i + 1] - lo[i] for lo, _ in tr.sizes])) for i in range(len(wl.fanouts) + 1)],
                 block_edges=[int(np.mean([bo[i + 1] - bo[i] for _, bo in tr.sizes])) for i in range(len(wl.fanouts))])
        if host_inputs:
            b = args.batch_size * ID_BYTES
            r["h2d_bytes_per_step"] = int(2 * b + 4 * args.feat_size * M / steps)   # seeds + labels + missed input rows over PCIe
            r["d2h_bytes_per_step"] = int(4 + 8 * 23)                   # loss + NodeFlow meta block
        res["e2e" if host_inputs else "value"] = r
        del tr
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------ CPU arm
class CpuAgg(torch.autograd.Function):
    """block_compute(copy_src, mean) on the CPU through the oracle (for the CPU arm's model)."""

    @staticmethod
    def forward(ctx, src, block, threads):
        import oracle
        ip, cols, base = block
        ctx.block, ctx.n_src = block, src.shape[0]
        return torch.from_numpy(oracle.aggregate(ip, cols, base, src.detach().numpy(), "mean", threads, f32=True))

    @staticmethod
    def backward(ctx, g):
        import oracle
        ip, cols, base = ctx.block
        return torch.from_numpy(oracle.aggregate_bwd(ip, cols, base, g.contiguous().numpy(), ctx.n_src, "mean")), None, None


def _cpu_model(args):
    """torch-CPU modules of the config's model (parameter shapes of PaGraph/model/gcn_nssc.py / graphsage_nssc.py)"""
    Fd, H, C, nl = args.feat_size, args.n_hidden, args.n_classes, args.n_layers
    torch.manual_seed(0)
    if args.model == "sage":
        dims = [(Fd, H)] + [(H, H)] * (nl - 1) + [(2 * H, C)]
        mods = [(torch.nn.Linear(i, o), torch.nn.Linear(i, o)) for i, o in dims]          # (fc_self, fc_neigh) per NodeUpdate
        params = [p for pair in mods for m in pair for p in m.parameters()]
    else:
        mods = [torch.nn.Linear(Fd, H)] + [torch.nn.Linear(H, H) for _ in range(nl - 1)] + [torch.nn.Linear(2 * H, C)]
        params = [p for m in mods for p in m.parameters()]
    return mods, params


def _cpu_forward(args, mods, nf, frames, threads):
    """forward of the config's model on one OracleNodeFlow with the oracle's CPU aggregation (fp32, `threads` threads)"""
    L, nl = nf.num_layers - 1, args.n_layers

    def act(z, last_hidden):
        return torch.cat((z, F.relu(z)), 1) if last_hidden else F.relu(z)
    feats = [torch.from_numpy(fr["features"]) for fr in frames]
    if args.model == "gcn":
        h = feats[0]
        for i, lin in enumerate(mods):
            h = lin(CpuAgg.apply(h, nf.block(i), threads))
            if i < len(mods) - 1:
                h = act(h, i == nl - 1)
        return h
    if args.model == "gcn-pre":
        h = act(mods[0](feats[0]), nl == 1)
        for i, lin in enumerate(mods[1:]):
            h = lin(CpuAgg.apply(h, nf.block(i), threads))
            if i < len(mods) - 2:
                h = act(h, i == nl - 2)
        return h
    h = {l: feats[l] for l in range(L + 1)}                 # sage: NodeUpdate `lid` is applied to every remaining block
    for lid, (fc_self, fc_neigh) in enumerate(mods):
        new = {}
        for i in range(lid, L):
            z = fc_self(h[i + 1]) + fc_neigh(CpuAgg.apply(h[i], nf.block(i), threads))
            new[i + 1] = z if lid == len(mods) - 1 else act(z, lid == nl - 1)
        h = new
    return h[L]


def cpu_path(args, indptr, indices, tables, seeds, n_batches, threads, first_batch=0):
    """The reference path restated on the CPU (oracle/): DGL-style sampling (OpenMP over batches, one
    thread each), storage.py's host gather of every layer's rows for every field (all rows come from
    the host table — the reference's miss path, storage.py:117-129 — and no H2D copy is charged),
    fp32 mean aggregation and the model's forward/backward/Adam on torch-CPU. Returns seconds."""
    import oracle
    fan = [int(x) for x in args.fanout.split(",")]
    V = len(indptr) - 1
    mods, params = _cpu_model(args)
    opt = torch.optim.Adam(params, lr=args.lr)
    labels = torch.randint(0, args.n_classes, (V,))
    flag = np.zeros(V, np.uint8)
    ident = np.zeros(1, np.int64)
    t_total = 0.0
    done = 0
    stage = {"sample": 0.0, "gather": 0.0, "aggregate+model": 0.0}
    while done < n_batches:
        nb = min(threads, n_batches - done)
        t0 = time.perf_counter()
        nfs = oracle.sample_many(indptr, indices, None, seeds, args.batch_size, first_batch + done, nb, fan,
                                 seed=args.seed, threads=threads)
        t1 = time.perf_counter()
        stage["sample"] += t1 - t0
        for nf in nfs:
            t1 = time.perf_counter()
            frames = []
            for i in range(nf.num_layers):
                ids = nf.layer_parent_nid(i)
                fr = {}
                for name, tab in tables.items():
                    fr[name] = _cpu_fetch(oracle, ids, flag, ident, tab, threads)
                frames.append(fr)
            t2 = time.perf_counter()
            # dropout mask generation is left out of the CPU arm: torch-CPU bernoulli is single-threaded
            # (0.6 s per minibatch here) and would dominate; leaving it out favours the CPU arm.
            pred = _cpu_forward(args, mods, nf, frames, threads)
            loss = F.cross_entropy(pred, labels[torch.from_numpy(nf.layer_parent_nid(-1))])
            opt.zero_grad()
            loss.backward()
            opt.step()
            t3 = time.perf_counter()
            stage["gather"] += t2 - t1
            stage["aggregate+model"] += t3 - t2
        done += nb
        t_total += time.perf_counter() - t0
    return t_total, stage


def _cpu_fetch(oracle, ids, flag, ident, tab, threads):
    """storage.py:126-128 host gather of one field with `threads` OpenMP threads (nid_map = identity)."""
    import ctypes
    L = oracle.lib()
    n, dim = len(ids), tab.shape[1]
    out = np.empty((n, dim), np.float32)
    i64p, u8p, f32p = (ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float))
    ids = np.ascontiguousarray(ids, np.int64)
    tp = ctypes.cast(ctypes.c_void_p(tab.data_ptr()), f32p)
    # flag is all-zero => every row is read from host[nid_map[t]]; nid_map is passed as the ids themselves
    L.pgo_fetch(ids.ctypes.data_as(i64p), ctypes.c_int64(n), flag.ctypes.data_as(u8p),
                ident.ctypes.data_as(i64p), _identity(len(flag)).ctypes.data_as(i64p), tp, ctypes.c_int64(0), tp,
                ctypes.c_int64(tab.stride(0)), ctypes.c_int64(dim), out.ctypes.data_as(f32p), None,
                ctypes.c_int(threads))
    return out


_IDENT = {}


def _identity(n):
    if n not in _IDENT:
        _IDENT[n] = np.arange(n, dtype=np.int64)
    return _IDENT[n]


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(args):
    model = {"gcn": "GCN-2L", "gcn-pre": "GCN-2L --preprocess", "sage": "GraphSAGE-%dL" % (args.n_layers + 1)}[args.model]
    part = "hash" if args.partition == "hash" else "dg/%d" % args.parts
    scale = "" if args.scale == 1.0 else " scale=%g" % args.scale
    return ("cfg%d %s %.3gM vtx/%.3gM edges f%d, %s h%d c%d, fanout %s, batch %d, %s%s"
            % (args.config, args.tag, args.vnum / 1e6, args.nnz / 1e6, args.feat_size, model, args.n_hidden, args.n_classes,
               args.fanout.replace(",", "/"), args.batch_size, part, scale))



# ------------------------------------------------------------------------------------ parity gate / replica check
def parity_gate(wl, tr, batch_idx=3):
    """One minibatch at the benchmarked size through the CUDA path against the CPU oracle, before any timing
    (reference loop examples/profile/pa_gcn.py:86-97): pg_sample vs oracle.sample bit-exact (ids, CSR, edge ids),
    the cache fetch of EVERY layer bit-exact against the host feature table, the fused block-0 aggregation within
    1e-5 relative of the float64 oracle. The oracle is the checker here, never the measured path. Raises on mismatch."""
    import oracle
    from pagraph_b200 import ops
    a = wl.args
    indptr, indices = wl.host_graph()
    sm = tr.sampler
    nf = sm.sample_batch(batch_idx, epoch=0)
    seeds = sm._seeds_cpu[batch_idx * a.batch_size:(batch_idx + 1) * a.batch_size].numpy()
    ref = oracle.sample(indptr, indices, None, seeds, wl.fanouts, seed=a.seed, epoch=0, batch=batch_idx)
    ok = nf._layer_offsets == ref.layer_offsets.tolist()
    ok = ok and np.array_equal(nf._node_mapping.tousertensor().cpu().numpy(), ref.node_mapping)
    ok = ok and np.array_equal(nf._indptr.cpu().numpy(), ref.indptr)
    ok = ok and np.array_equal(nf._indices.cpu().numpy(), ref.indices)
    ok = ok and np.array_equal(nf._edge_mapping.tousertensor().cpu().numpy(), ref.edge_mapping)
    if not ok:
        raise SystemExit("parity gate: sampled NodeFlow differs from the oracle (config size, batch %d)" % batch_idx)
    c = tr.cacher
    ids = nf._node_mapping.tousertensor()
    outs = c._gather(ids, c._field_names)
    ids_cpu = torch.from_numpy(ref.node_mapping)
    for name, got in zip(c._field_names, outs):
        want = wl.host_rows(name, ids_cpu)
        if not torch.equal(got.cpu().view(torch.int32), want.view(torch.int32)):
            raise SystemExit("parity gate: fetched rows of %r differ from the host table" % name)
    bi, bc, bb, n_dst, n_src = nf.block_csr(0)
    got = ops.cache_aggregate(c, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean").cpu().numpy()
    ip, cols, base = ref.block(0)
    feats0 = wl.host_rows("features", torch.from_numpy(ref.layer_parent_nid(0))).numpy()
    want = oracle.aggregate(ip, cols, base, feats0, "mean", threads=host_threads())
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
    if not err <= 1e-5:
        raise SystemExit("parity gate: block-0 aggregate off by %.3g relative (> 1e-5)" % err)
    return {"status": "ok", "nodes": int(ref.layer_offsets[-1]), "edges": int(len(ref.indices)), "agg_max_rel_err": err}


def replica_check(tr, world):
    """After the timed steps every rank must hold bit-identical parameters (the gradient all-reduce is the only
    exchange, pa_gcn.py:65,96): all-gather the flat bucket's checksum and first/last words."""
    flat = tr.sync.flat_param.detach()
    sig = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[0].double(), flat[-1].double()])
    bits = flat.view(torch.int32).to(torch.int64)
    sig = torch.cat([sig, torch.stack([(bits * (torch.arange(bits.numel(), device=bits.device) % 8191 + 1)).sum().double()])])
    if world == 1:
        return True
    allsig = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(allsig, sig)
    return all(torch.equal(allsig[0], x) for x in allsig[1:])


# ------------------------------------------------------------------------------------ the driver-facing JSON line
def _r(x, nd=5):
    """round to nd significant digits (keeps the line short)"""
    if x is None or isinstance(x, (bool, int, str)):
        return x
    x = float(x)
    if x == 0 or x != x or x in (float("inf"), float("-inf")):
        return x if x == x and abs(x) != float("inf") else None
    from math import floor, log10
    return round(x, nd - 1 - int(floor(log10(abs(x)))))


def dominant_kernel(kernels):
    """The HBM-bound kernel class with the largest serialised time per launch: the roofline record's subject. Only
    single-kernel classes qualify (`sample` is a chain of seven), and the all-reduce does not — its duration is the
    wait for the slowest rank of the step, not bytes moved."""
    hbm_k = {k: x for k, x in kernels.items() if x["bound"] == "hbm" and "sample" not in k and "allreduce" not in k}
    return max(hbm_k, key=lambda k: hbm_k[k]["avg_ms"])


def compact_line(d):
    """The LAST stdout line: < 1150 bytes (the driver keeps a 1500-byte tail), every string <= 120 chars, one JSON object the driver parses. `d` is the full
    detail record (written to gpurun_out/bench_detail_n{N}.json); only the contract keys and the headline split go here."""
    rf, cpu, e, ck = d["roofline"], d.get("cpu_baseline"), d["e2e"], d.get("clocks") or {}
    line = {
        "metric": "minibatches/s", "value": _r(d["value"], 6), "unit": "minibatches/s", "n_gpus": d["n_gpus"],
        "steps": d["steps"], "warmup": d["warmup"], "ms_per_step": _r(d["ms_per_step"]), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": d["config"]["workload"][:120], "cache_mode": d["config"]["cache_mode"],
                   "path": d["config"]["path"], "l2": "inputs>L2 (feature table, new batch/step)"},
        "e2e": {"value": _r(e["value"], 6), "unit": "minibatches/s", "h2d_bytes_per_step": e["h2d_bytes_per_step"],
                "d2h_bytes_per_step": e["d2h_bytes_per_step"]},
        "gpu_launches": d["gpu_launches"],
        "roofline": {"bound": rf["bound"], "kernel": rf["kernel"].split("(")[-1].split(",")[0].rstrip(")")[:40], "achieved": _r(rf["achieved"]), "peak": _r(rf["peak"]),
                     "unit": "GB/s", "frac": _r(rf["frac"], 4), "traffic": rf.get("traffic"),
                     "alg_bytes_per_launch": _r(rf["alg_bytes_per_launch"], 6), "avg_ms": _r(rf["avg_ms"], 4),
                     **({"frac_all_sms": _r(rf["frac_all_sms"], 4), "sms_left_out": rf.get("sms_left_out")}
                        if rf.get("frac_all_sms") else {})},
        "cpu_baseline": None if not cpu else {"value": _r(cpu["value"], 4), "unit": "minibatches/s", "cores": cpu["cores"],
                                              "kind": cpu["kind"], "sample": cpu["sample"][:64]},
        "clocks": {"sm_mhz": ck.get("sm_mhz"), "sm_max_mhz": ck.get("sm_max_mhz"), "reasons": ck.get("reasons", [])},
        "gather_gbs": _r(d.get("gather_gbs"), 4), "hit_rate": _r(d.get("hit_rate"), 4),
        "parity_gate": d.get("parity_gate"), "replicas_identical": d.get("replicas_identical"),
    }
    vkey = next((k for k in d if k.startswith("vtx") and isinstance(d[k], dict)), None)
    if vkey:
        line[vkey] = {k: _r(d[vkey].get(k), 4) for k in ("value", "e2e", "gather_gbs", "hit_rate", "miss_frac_pcie")}
    pkey = next((k for k in d if k.startswith("peer") and isinstance(d[k], dict)), None)
    if pkey:
        line[pkey] = {k: _r(d[pkey].get(k), 4) for k in ("value", "e2e", "nvlink_mb_per_step", "pcie_mb_per_step")}
    out = json.dumps(line, separators=(",", ":"))
    # never let the line grow past what the driver keeps: shorten the free-text strings first, then drop the optional tail
    for shrink in (lambda: line["cpu_baseline"] and line["cpu_baseline"].update(sample=line["cpu_baseline"]["sample"][:48]),
                   lambda: line["roofline"].update(kernel=line["roofline"]["kernel"][:32]),
                   lambda: line["config"].update(l2="inputs>L2"),
                   lambda: line.pop(pkey, None), lambda: line.pop(vkey, None), lambda: line.pop("gather_gbs", None)):
        if len(out) < 1150:
            break
        shrink()
        out = json.dumps(line, separators=(",", ":"))
    return out


def write_detail(detail, n, config=2):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = "bench_detail_n%d.json" % n if config == 2 else "bench_detail_cfg%d_n%d.json" % (config, n)
        path = os.path.join(ROOT, "gpurun_out", name)
        with open(path, "w") as f:
            json.dump(detail, f, indent=1)
        return os.path.relpath(path, ROOT)
    except Exception as ex:   # read-only checkout: the detail goes to stderr instead
        print(json.dumps(detail), file=sys.stderr)
        return "stderr (%s)" % type(ex).__name__


# ------------------------------------------------------------------------------------ reference arm
def main_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port; the reference itself is pure Python over
    dgl==0.4.1, which is not installable here) on this box's host cores. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pagraph_b200 import data
    threads = host_threads()
    if torch.cuda.is_available():            # data generation only (setup, untimed) — same graph as our arm
        dev = torch.device("cuda", 0)
        a_, b_, c_ = args.rmat
        ip, ix = data.rmat_in_csr_cuda(args.vnum, args.nnz, seed=args.seed, device=dev, a=a_, b=b_, c=c_)
        indptr, indices = ip.cpu().numpy(), ix.cpu().numpy()
        feat = torch.empty((args.vnum, args.feat_size), dtype=torch.float32)
        gen = torch.Generator(device=dev)
        gen.manual_seed(2)
        chunk = 1 << 18
        for lo in range(0, args.vnum, chunk):
            hi = min(args.vnum, lo + chunk)
            feat[lo:hi].copy_(torch.rand((hi - lo, args.feat_size), device=dev, generator=gen))
        norm = (1.0 / (ip[1:] - ip[:-1]).float()).unsqueeze(1).cpu()
        gen.manual_seed(4)
        perm = torch.randperm(args.vnum, device=dev, generator=gen)
        train = torch.sort(perm[:int(args.vnum * 0.65)]).values.cpu().numpy()
        del ip, ix, perm
        torch.cuda.empty_cache()
    else:
        adj = data.rmat_adj(args.vnum, args.nnz, seed=args.seed).tocsc()
        indptr, indices = adj.indptr.astype(np.int64), adj.indices.astype(np.int64)
        feat = torch.rand((args.vnum, args.feat_size))
        norm = torch.from_numpy(1.0 / np.maximum(np.diff(indptr), 1).astype(np.float32)).unsqueeze(1)
        train = np.sort(np.random.default_rng(4).permutation(args.vnum)[:int(args.vnum * 0.65)])
    from pagraph_b200.parallel import hash_split
    # the CPU arm walks the full graph with a 1/parts hash share of the train ids (for the dg configs too: the partition
    # only decides which seeds a rank owns, and a dg partition's closure is nearly the whole graph at these densities)
    seeds = np.sort(hash_split(train, max(args.parts or args.gpus, 1), seed=5)[0])
    torch.manual_seed(0)
    seeds = np.ascontiguousarray(seeds[torch.randperm(len(seeds)).numpy()])
    tables = {"features": feat, "norm": norm}
    tables = {f: tables[f] for f in args.fields}
    torch.set_num_threads(threads)
    cpu_path(args, indptr, indices, tables, seeds, args.warmup, threads, first_batch=0)
    t, stage = cpu_path(args, indptr, indices, tables, seeds, args.steps, threads, first_batch=args.warmup)
    v = args.steps / t
    line = {"impl": "reference", "metric": "minibatches/s", "value": v,
            "unit": "minibatches/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "cache_mode": "none (CPU arm: every row from the host table)",
                       "path": "oracle port, 1 process (also at n_gpus > 1)"},
            "cpu_baseline": {"value": v, "unit": "minibatches/s", "cores": threads, "kind": "port",
                             "sample": "%d full minibatches; sample %d thr over batches, gather+aggregate %d thr"
                                       % (args.steps, threads, threads),
                             "stage_s": {k: round(x, 3) for k, x in stage.items()}},
            "e2e": {"value": v, "unit": "minibatches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line, separators=(",", ":")), flush=True)


# ------------------------------------------------------------------------------------ our arm
def main_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from pagraph_b200 import _lib
    import ctypes
    hbm_peak, hbm_src = measured_peaks()
    bw = ctypes.c_double()
    _lib.check(_lib.lib().pg_measure_h2d(dev.index, 1 << 30, 5, ctypes.byref(bw)), "pg_measure_h2d")
    pcie_peak = bw.value
    t_setup = time.time()
    watchdog = int(os.environ.get("PG_BENCH_WATCHDOG", "0"))
    if watchdog > 0:                       # a hung run leaves the Python stacks of every thread on stderr
        import faulthandler
        faulthandler.dump_traceback_later(watchdog, repeat=True, file=sys.stderr)
    note("workload")
    wl = Workload(args, rank, world, dev)
    modes = args.modes.split(",")
    results = {}
    for m in modes:
        results[m] = run_mode(wl, m, args, world, hbm_peak, pcie_peak)
    t_setup = time.time() - t_setup

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        note("cpu baseline")
        threads = host_threads()
        indptr, indices = wl.host_graph()
        tables = {f: wl.store.ndata[f] for f in args.fields}
        if wl.nid_map is not None:       # the CPU arm walks the same partition graph; its rows come through nid_map
            nm = wl.nid_map.cpu()
            tables = {f: t[nm] for f, t in tables.items()}
        torch.manual_seed(0)
        seeds = np.ascontiguousarray(wl.train_nid[torch.randperm(len(wl.train_nid)).numpy()])
        torch.set_num_threads(threads)
        t, stage = cpu_path(args, indptr, indices, tables, seeds, args.cpu_batches, threads)
        cpu = {"value": args.cpu_batches / t, "unit": "minibatches/s", "cores": threads, "kind": "port",
               "sample": "%d full minibatches, oracle port, %d threads" % (args.cpu_batches, threads),
               "stage_s": {k: round(v, 3) for k, v in stage.items()}}
    note("done")
    wl.close()
    if rank != 0:
        return
    head = results[modes[0]]
    v, e = head["value"], head["e2e"]
    # dominant HBM-bound kernel of the headline mode
    top = dominant_kernel(v["kernels"])
    tk = v["kernels"][top]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(top.split("(")[0])
            if isinstance(traffic, dict):
                traffic = traffic.get("bytes_per_launch")
    except Exception:
        pass
    detail = {
        "metric": "minibatches/sec (sample + cache fetch + train step) + feature-gather GB/s",
        "value": v["minibatches_per_s"], "unit": "minibatches/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "cache_mode": modes[0], "path": args.path,
                   "cache": "hbmNN = capacity NN% of the HBM bytes (BASELINE.json's wording; >= the feature table -> "
                            "full_cached, headline); vtxNN = top-NN%-out-degree vertices of the partition cached (real "
                            "hit/miss split)",
                   "setup": wl.setup,
                   "dropout": args.dropout, "optimizer": "Adam lr %g" % args.lr,
                   "l2": "inputs larger than L2 (24 GB feature table, new random minibatch every step); no flush",
                   "kernel_timing": ("pg_timing_* CUDA-event pairs on the launching stream; with --path engine the timed region "
                                     "replays CUDA graphs (no host-side launch to bracket), so the per-kernel numbers come from "
                                     "instrumented un-graphed passes of the same pipeline over %d further minibatches each: "
                                     "avg_ms / achieved / roofline = stages serialised, every kernel alone on the GPU (as ncu "
                                     "sees it); avg_ms_in_pipeline = the three streams overlapped as in the timed region"
                                     % min(args.kernel_steps, args.steps)) if args.path == "engine" else
                                    "pg_timing_* CUDA-event pairs on the launching stream inside the timed region",
                   "parallelism": "dp%d (one partition per GPU; flat-bucket gradient all-reduce fused with Adam over NVLink peer memory)" % world},
        "gather_gbs": v["gather"]["gather_gbs"], "hit_rate": v["hit_rate"],
        "gather": {m: results[m]["value"]["gather"] for m in modes},
        "e2e": {"value": e["minibatches_per_s"], "unit": "minibatches/s", "h2d_bytes_per_step": e["h2d_bytes_per_step"],
                "d2h_bytes_per_step": e["d2h_bytes_per_step"], "ms_per_step": e["ms_per_step"],
                "wall_ms_per_step": e["wall_ms_per_step"]},
        "gpu_launches": v["launches"],
        "clocks": v["clocks"],
        "roofline": {"bound": "hbm", "kernel": top, "achieved": tk["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                     "frac": tk["frac"], "traffic": traffic, "peak_source": hbm_src,
                     "alg_bytes_per_launch": tk["alg_bytes_per_launch"], "avg_ms": tk["avg_ms"],
                     "frac_all_sms": tk.get("frac_all_sms"), "sms_left_out": tk.get("sms_left_out")},
        "pcie": {"peak_gbs_measured_h2d": pcie_peak},
        "cpu_baseline": cpu,
        "parity_gate": (v.get("parity_gate") or {}).get("status"),
        "replicas_identical": bool(v["replicas_identical"] and e["replicas_identical"]),
        "kernels": v["kernels"],
        "variants": {m: results[m] for m in modes},
        "setup_s": round(t_setup, 1),
    }
    vmode = next((m for m in modes[1:] if m.startswith("vtx")), None)
    if vmode:
        x = results[vmode]
        g = x["value"]["gather"] or {}
        detail[vmode] = {"value": x["value"]["minibatches_per_s"], "e2e": x["e2e"]["minibatches_per_s"],
                           "gather_gbs": g.get("gather_gbs"), "hit_rate": g.get("hit_rate"),
                           "miss_frac_pcie": (g.get("miss") or {}).get("frac")}
    pmode = next((m for m in modes[1:] if m.startswith("peer")), None)
    if pmode:
        x = results[pmode]
        detail[pmode] = {"value": x["value"]["minibatches_per_s"], "e2e": x["e2e"]["minibatches_per_s"],
                         "nvlink_mb_per_step": x["value"]["nvlink_bytes_per_step"] / 1e6,
                         "pcie_mb_per_step": x["value"]["pcie_bytes_per_step"] / 1e6, "hit_rate": x["value"]["hit_rate"]}
    detail["detail_file"] = write_detail(detail, world, args.config)
    sys.stdout.flush()
    print(compact_line(detail), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
    if dist.is_initialized():
        dist.destroy_process_group()
