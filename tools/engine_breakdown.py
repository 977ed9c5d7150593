#!/usr/bin/env python
"""Where does a GCNTrainEngine step go? At bench.py scale, times N back-to-back replays of each stage's graph alone
(sample / gather+aggregate / compute) and the real three-stream pipelined step. CUDA events, one GPU.
    python tools/engine_breakdown.py [N] [hbm20|vtx20]"""
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

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn, stream=None):
        st = stream or torch.cuda.current_stream()
        with torch.cuda.stream(st):
            a, b = ev(), ev()
            a.record(st)
            for _ in range(n):
                fn()
            b.record(st)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    a, b = ev(), ev()
    a.record()
    tr.run(n, record=False)
    b.record()
    torch.cuda.synchronize()
    out["pipelined_step_ms"] = a.elapsed_time(b) / n
    s = eng.slots[0]
    caps = next(iter(s.compute_graphs))
    out["compute_graph_only_ms"] = timed(s.compute_graphs[caps][0].replay)
    out["sample_graph_only_ms"] = timed(s.sample_graph.replay, eng.side)
    out["gather_graph_only_ms"] = timed(s.gather_graph.replay, eng.gather)
    out["caps"] = list(caps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
