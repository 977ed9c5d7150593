#!/usr/bin/env python
"""Where does a GCNTrainEngine step go? Times, at bench.py scale (hbm20), N back-to-back replays of (a) the compute graph
alone, (b) the load graph alone, (c) both pipelined (the real step). CUDA events, one GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    mode = sys.argv[2] if len(sys.argv) > 2 else "hbm20"
    sys.argv = [sys.argv[0]]
    args = bench.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = bench.Workload(args, 0, 1, dev)
    tr = bench.Trainer(wl, mode, False)
    eng = tr.engine
    tr.run(12, record=False)
    torch.cuda.synchronize()
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a, b = ev(), ev()
    a.record()
    tr.run(n, record=False)
    b.record()
    torch.cuda.synchronize()
    out["pipelined_step_ms"] = a.elapsed_time(b) / n
    s = eng.slots[0]
    caps = next(iter(s.compute_graphs))
    g = s.compute_graphs[caps][0]
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    out["compute_graph_only_ms"] = a.elapsed_time(b) / n
    with torch.cuda.stream(eng.side):
        a, b = ev(), ev()
        a.record(eng.side)
        for _ in range(n):
            s.load_graph.replay()
        b.record(eng.side)
    torch.cuda.synchronize()
    out["load_graph_only_ms"] = a.elapsed_time(b) / n
    # the engine's own loop with the load stage stubbed out (slot data left as is): host overhead + compute chain
    real_issue = eng._issue_load

    def fake_issue(k):
        sl = eng.slots[k % 3]
        sl.n_valid, sl.k = eng.batch, k
        eng.side.wait_event(sl.done)
        sl.loaded.record(eng.side)
    eng._issue_load = fake_issue
    tr.run(10, record=False)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    tr.run(n, record=False)
    b.record()
    torch.cuda.synchronize()
    out["engine_loop_without_load_ms"] = a.elapsed_time(b) / n
    eng._issue_load = real_issue
    # and with the compute graph stubbed out: host overhead + load chain
    real_compute = {}
    for sl in eng.slots:
        real_compute[id(sl)] = dict(sl.compute_graphs)

    class _Nop:
        def replay(self):
            pass
    for sl in eng.slots:
        sl.compute_graphs = {k: (_Nop(), v[1]) for k, v in sl.compute_graphs.items()}
    tr.run(10, record=False)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    tr.run(n, record=False)
    b.record()
    torch.cuda.synchronize()
    out["engine_loop_without_compute_ms"] = a.elapsed_time(b) / n
    for sl in eng.slots:
        sl.compute_graphs = real_compute[id(sl)]
    # which part of the load stage is visible in the step? re-capture the load graphs with only one part each
    import ctypes
    from pagraph_b200 import _lib

    def recapture(part):
        def body(sl, n_seeds):
            L, c = _lib.lib(), eng.cacher
            st = _lib.stream_ptr()
            if part in ("sample", "all"):
                nfb = _lib.pg_nodeflow_buffers(*[_lib.ptr(sl.nf[k]) for k in ("node_mapping", "indptr", "indices", "edge_mapping", "meta")])
                key = ctypes.c_void_p(sl.seeds_key.data_ptr() + 8 * eng.batch)
                _lib.check(L.pg_sample_keyed(eng.sampler, _lib.ptr(sl.seeds_key), n_seeds, key, ctypes.byref(nfb), _lib.ptr(sl.h_meta), st), "s")
            if part in ("fetch", "all"):
                outs = (ctypes.c_void_p * len(sl.rest))(*[t.data_ptr() for t in sl.rest])
                _lib.check(L.pg_cache_fetch_dyn(c._handle, _lib.ptr(sl.nf["node_mapping"]), eng._meta_ptr(sl, 5), eng._meta_ptr(sl, 4 + eng.L + 1),
                                                eng.cap_rest, outs, None, 0, st), "f")
                blk = _lib.pg_block(_lib.ptr(sl.nf["node_mapping"]), _lib.ptr(sl.nf["indptr"]), _lib.ptr(sl.nf["indices"]), 0, eng.cap_n0,
                                    eng.cap_layer[-2], eng._meta_ptr(sl, 4))
                _lib.check(L.pg_cache_resolve(c._handle, eng.fi, ctypes.byref(blk), _lib.ptr(sl.rowptr), _lib.ptr(sl.stage), eng.stage_rows, None, st), "r")
        torch.cuda.synchronize()
        for sl in eng.slots:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.stream(eng.side):
                with torch.cuda.graph(g2, stream=eng.side, capture_error_mode="thread_local"):
                    body(sl, eng.batch)
            sl.load_graph = g2
    for part in ("sample", "fetch", "all"):
        recapture(part)
        tr.run(10, record=False)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        tr.run(n, record=False)
        b.record()
        torch.cuda.synchronize()
        out["pipelined_with_load=%s_ms" % part] = a.elapsed_time(b) / n
    import time
    t0 = time.perf_counter()
    tr.run(n, record=False)
    out["host_issue_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / n
    torch.cuda.synchronize()
    out["caps"] = list(caps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
