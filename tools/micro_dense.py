#!/usr/bin/env python
"""Micro-benchmark of the dense-stage kernels at bench.py scale (config 2: n_1 = 36 864 padded rows, F = 600, hidden 32,
6000 seeds, 60 classes) on synthetic inputs — no graph, no cache: seconds to set up, so it is also the cheap target for
`ncu --set full`. Inputs rotate over buffers larger than L2. CUDA events around CUDA-graph replays, median.
    python tools/micro_dense.py [--iters 30] [--fwd-variants 0,1,5] [--only fwd|bwd|head]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def med(fn, iters, reps=6):
    """median GPU time of one call, in us. The calls are replayed from a CUDA graph (`reps` per replay, inputs rotating):
    issued one by one from Python, a 25 us kernel is followed by an idle GPU and the event pair times the host."""
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for r in range(reps):
            fn(r)
    g.replay()
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3 / reps   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--n", type=int, default=36864)
    ap.add_argument("--F", type=int, default=600)
    ap.add_argument("--fwd-variants", default="0,u", help="mma.sync template variants 0/1/5, u = the tcgen05 kernel")
    ap.add_argument("--only", default="")
    ap.add_argument("--p", type=float, default=0.2)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    from pagraph_b200 import _lib
    from pagraph_b200.ops import linear_concat_backward, linear_concat_forward
    torch.manual_seed(0)
    n, F, nb, C = a.n, a.F, 6000, 60
    xs = [torch.randn(n, F, device="cuda") for _ in range(3)]            # 3 x 88 MB > L2
    lin = torch.nn.Linear(F, 32).cuda()
    W, b = lin.weight.detach(), lin.bias.detach()
    out = torch.empty(n, 64, device="cuda")
    od = torch.empty(n, 64, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    res = {"n": n, "F": F, "hbm_floor_us": 4 * n * F / 6551.7e3}
    ref = None
    if a.only in ("", "fwd"):
        for v in a.fwd_variants.split(","):
            os.environ["PG_FWD_UMMA"] = "1" if v == "u" else "0"
            os.environ["PG_FWD_VARIANT"] = "0" if v == "u" else v
            t = med(lambda i: linear_concat_forward(xs[i % 3], W, b, True, out=out, out_drop=od, dropout_p=a.p, seed=5, step=step),
                    a.iters)
            linear_concat_forward(xs[0], W, b, True, out=out, out_drop=od, dropout_p=a.p, seed=5, step=step)
            if ref is None:
                ref = out.clone()
                z = torch.nn.functional.linear(xs[0].double(), W.double(), b.double())
                res["fwd_max_err_vs_fp64"] = float((out[:, :32].double() - z).abs().max())
            zz = torch.nn.functional.linear(xs[0].double(), W.double(), b.double())
            res["fwd_variant_%s_max_err_vs_fp64" % v] = float((out[:, :32].double() - zz).abs().max())
            keep = od != 0
            res["fwd_variant_%s_drop_ok" % v] = bool(torch.equal(od[keep], (out * (1.0 / (1.0 - a.p)))[keep])) if a.p > 0 else None
            res["fwd_variant_%s_us" % v] = round(t, 2)
            res["fwd_variant_%s_maxdiff" % v] = float((out - ref).abs().max())
        os.environ.pop("PG_FWD_VARIANT", None)
        os.environ.pop("PG_FWD_UMMA", None)
        tl = med(lambda i: torch.nn.functional.linear(xs[i % 3], W, b), a.iters)
        res["cublas_fp32_linear_us"] = round(tl, 2)
    if a.only in ("", "bwd"):
        g = torch.randn(n, 64, device="cuda")
        gw = torch.empty(32, F, device="cuda")
        gb = torch.empty(32, device="cuda")
        linear_concat_forward(xs[0], W, b, True, out=out)
        gz_ref = None
        for name, env in (("umma_3xtf32", {"PG_DW_UMMA": "1"}), ("mma_3xtf32", {"PG_DW_UMMA": "0"}),
                          ("simt_ffma2", {"PG_DW_UMMA": "0", "PG_DENSE_SIMT": "1"})):
            os.environ.update(env)
            p = 0.0 if name == "simt_ffma2" else a.p
            t = med(lambda i: linear_concat_backward(xs[i % 3], g, out, True, gw, gb, p, 5, step), a.iters)
            res["bwd_%s_us" % name] = round(t, 2)
            linear_concat_forward(xs[0], W, b, True, out=out, out_drop=od, dropout_p=p, seed=5, step=step)
            linear_concat_backward(xs[0], g, out, True, gw, gb, p, 5, step)
            keep = (od != 0) if p > 0 else torch.ones_like(od, dtype=torch.bool)
            gm = g.double() * keep * (1.0 / (1.0 - p))
            gz64 = gm[:, :32] + gm[:, 32:] * (out[:, 32:] > 0)
            res["bwd_%s_dw_rel_err" % name] = float((gw.double() - gz64.t() @ xs[0].double()).abs().max() / (gz64.t() @ xs[0].double()).abs().max())
            res["bwd_%s_db_rel_err" % name] = float((gb.double() - gz64.sum(0)).abs().max() / gz64.sum(0).abs().max())
            for k in env:
                os.environ.pop(k, None)
        linear_concat_backward(xs[0], g, out, True, gw, gb, 0.0, 5, step)
        pos = out[:, 32:] > 0
        gz = g[:, :32].double() + g[:, 32:].double() * pos
        res["bwd_max_err_vs_fp64"] = float((gw.double() - gz.t() @ xs[0].double()).abs().max())
    if a.only in ("", "head"):
        L = _lib.lib()
        a2 = torch.randn(nb, 64, device="cuda")
        head = torch.nn.Linear(64, C).cuda()
        y = torch.randint(0, C, (nb,), device="cuda")
        loss = torch.zeros((), device="cuda")
        ga = torch.empty(nb, 64, device="cuda")
        gw1 = torch.empty(C, 64, device="cuda")
        gb1 = torch.empty(C, device="cuda")

        def run_head(i):
            _lib.check(L.pg_linear_cross_entropy(_lib.ptr(a2), 64, _lib.ptr(head.weight), _lib.ptr(head.bias), _lib.ptr(y), nb,
                                                 64, C, _lib.ptr(loss), _lib.ptr(ga), 64, _lib.ptr(gw1), _lib.ptr(gb1),
                                                 None, _lib.stream_ptr()), "pg_linear_cross_entropy")
        res["head_us"] = round(med(run_head, a.iters), 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
