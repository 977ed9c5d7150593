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
    import time
    t0 = time.perf_counter()
    tr.run(n, record=False)
    out["host_issue_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / n
    torch.cuda.synchronize()
    out["caps"] = list(caps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
