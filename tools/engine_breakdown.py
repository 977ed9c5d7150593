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

    import time
    if os.environ.get("PG_BD_INTER"):      # only gather alone vs gather beside the sample graph
        g0 = eng.slots[0].gather_graph.replay
        s1 = eng.slots[1].sample_graph.replay
        out["gather_alone_ms"] = timed(g0, eng.gather)
        out["sample_alone_ms"] = timed(s1, eng.slots[1].side)
        for limit in (1, 2, 3, 4, 5, 6, 7):      # the sample chain cut after its first `limit` kernels
            os.environ["PG_SAMPLE_KERNELS"] = str(limit)
            sl = eng.slots[1]
            with torch.cuda.stream(sl.side):
                gr, _ = eng._capture(sl.side, lambda: eng._sample_body(sl, eng.batch))
            e = [ev() for _ in range(4)]
            e[0].record(eng.gather)
            e[1].record(sl.side)
            for _ in range(n):
                with torch.cuda.stream(eng.gather):
                    g0()
                with torch.cuda.stream(sl.side):
                    gr.replay()
            e[2].record(eng.gather)
            e[3].record(sl.side)
            torch.cuda.synchronize()
            alone = timed(gr.replay, sl.side)
            out["first_%d_kernels" % limit] = dict(gather=round(e[0].elapsed_time(e[2]) / n, 4),
                                                    sample=round(e[1].elapsed_time(e[3]) / n, 4), sample_alone=round(alone, 4))
        os.environ.pop("PG_SAMPLE_KERNELS")
        e = [ev() for _ in range(4)]
        e[0].record(eng.gather)
        e[1].record(eng.slots[1].side)
        for _ in range(n):
            with torch.cuda.stream(eng.gather):
                g0()
            with torch.cuda.stream(eng.slots[1].side):
                s1()
        e[2].record(eng.gather)
        e[3].record(eng.slots[1].side)
        torch.cuda.synchronize()
        out["gather|sample"] = [round(e[0].elapsed_time(e[2]) / n, 4), round(e[1].elapsed_time(e[3]) / n, 4)]
        print(json.dumps(out))
        return
    a, b = ev(), ev()
    eng.host_prof = {}
    t0 = time.perf_counter()
    a.record()
    tr.run(n, record=False)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    out["pipelined_step_ms"] = a.elapsed_time(b) / n
    out["host_issue_ms_per_step"] = ((t1 - t0) - eng.host_prof.get("wait_s", 0.0)) / n * 1e3   # host time not blocked on the GPU
    out["host_blocked_ms_per_step"] = eng.host_prof.get("wait_s", 0.0) / n * 1e3
    out["samplers"] = eng.n_samplers
    eng.host_prof = None
    # GPU timeline of 12 consecutive steps: begin / end of every stage relative to the first event (ms)
    eng.trace = []
    tr.run(12, record=False)
    torch.cuda.synchronize()
    base = min(eng.trace, key=lambda t: -t[2].elapsed_time(eng.trace[0][2]))[2]
    tl = {}
    for stage, k, e in eng.trace:
        tl.setdefault(k, {})[stage] = round(base.elapsed_time(e), 4)
    out["timeline_ms"] = [dict(k=k, **v) for k, v in sorted(tl.items())]
    eng.trace = None
    # per-kernel-class timeline of 6 pipelined steps issued eagerly (no graph replay: pg_timing_* brackets every class)
    from pagraph_b200 import _lib
    names = ["sample", "split", "hit", "miss", "agg_fwd", "agg_bwd", "fused_agg", "dense_fwd", "dense_bwd", "head", "opt"]
    eng.use_graphs = False
    tr.run(6, record=False)
    torch.cuda.synchronize()
    _lib.timing_enable(True)
    tr.run(6, record=False)
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    out["kernel_timeline_ms"] = [[names[sl], round(a_, 4), round(b_, 4)] for sl, a_, b_ in _lib.timing_drain_timeline()]
    eng.use_graphs = True
    s = eng.slots[0]
    caps = next(iter(s.compute_graphs))
    out["compute_graph_only_ms"] = timed(lambda: [g.replay() for g in s.compute_graphs[caps][0]])
    out["split"] = bool(eng._split)
    out["sample_graph_only_ms"] = timed(s.sample_graph.replay, eng.side)
    out["gather_graph_only_ms"] = timed(s.gather_graph.replay, eng.gather)
    out["caps"] = list(caps)
    # interference: the gather graph of slot 0 replayed n times while other stages' graphs (of OTHER slots, so nothing it
    # reads is rewritten) are replayed beside it; per-replay time of each stream
    def together(pairs):
        evs = []
        for st, _ in pairs:
            e = ev()
            e.record(st)
            evs.append([e, None])
        for i in range(n):
            for st, fn in pairs:
                with torch.cuda.stream(st):
                    fn()
        for j, (st, _) in enumerate(pairs):
            e = ev()
            e.record(st)
            evs[j][1] = e
        torch.cuda.synchronize()
        return [round(a_.elapsed_time(b_) / n, 4) for a_, b_ in evs]

    main = torch.cuda.current_stream()
    g0 = eng.slots[0].gather_graph.replay
    s1 = eng.slots[1].sample_graph.replay
    sl2 = eng.slots[2]
    c2 = sl2.compute_graphs[next(iter(sl2.compute_graphs))][0]
    inter = {}
    inter["gather|sample"] = together([(eng.gather, g0), (eng.slots[1].side, s1)])
    if eng.n_samplers > 1:
        s2 = eng.slots[2].sample_graph.replay
        inter["gather|sample|sample"] = together([(eng.gather, g0), (eng.slots[1].side, s1), (eng.slots[2].side, s2)])
    inter["gather|compute"] = together([(eng.gather, g0), (main, lambda: [g.replay() for g in c2])])
    inter["compute|sample"] = together([(main, lambda: [g.replay() for g in c2]), (eng.slots[1].side, s1)])
    if len(c2) > 1:
        inter["gather|compute_rest"] = together([(eng.gather, g0), (main, c2[1].replay)])
    out["interference_ms"] = inter
    print(json.dumps(out))


if __name__ == "__main__":
    main()
