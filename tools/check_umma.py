#!/usr/bin/env python
"""Correctness sweep of the tcgen05 NodeUpdate kernels (forward ring depths, backward) against float64 at row counts that
give 1 .. several tiles per CTA. Run under `timeout`: python tools/check_umma.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from pagraph_b200.ops import linear_concat_backward, linear_concat_forward
    torch.manual_seed(0)
    res = {}
    K = 600
    lin = torch.nn.Linear(K, 32).cuda()
    W, b = lin.weight.detach(), lin.bias.detach()
    for n in (1000, 18944, 19000, 36864, 40000, 110000):
        x = torch.randn(n, K, device="cuda")
        z = torch.nn.functional.linear(x.double(), W.double(), b.double())
        ref = torch.cat((z, torch.relu(z)), 1)
        for tag, env in (("umma", {}), ("umma_again", {}), ("mma", {"PG_FWD_UMMA": "0"})):
            os.environ.update(env)
            errs = []
            for rep in range(4):
                out = torch.full((n, 64), 7.0, device="cuda")
                linear_concat_forward(x, W, b, True, out=out)
                torch.cuda.synchronize()
                errs.append(float((out.double() - ref).abs().max()))
            res["fwd_n%d_%s" % (n, tag)] = ["%.2e" % e for e in errs]
            for k in env:
                os.environ.pop(k)
        # backward
        g = torch.randn(n, 64, device="cuda")
        out = torch.empty(n, 64, device="cuda")
        linear_concat_forward(x, W, b, True, out=out)
        gz = g[:, :32].double() + g[:, 32:].double() * (out[:, 32:] > 0)
        gw_ref, gb_ref = gz.t() @ x.double(), gz.sum(0)
        for tag, env in (("umma", {"PG_DW_UMMA": "1"}), ("mma", {"PG_DW_UMMA": "0"})):
            os.environ.update(env)
            errs = []
            for rep in range(2):
                gw = torch.empty(32, K, device="cuda")
                gb = torch.empty(32, device="cuda")
                linear_concat_backward(x, g, out, True, gw, gb, 0.0, 0, None)
                torch.cuda.synchronize()
                errs.append("%.2e/%.2e" % (float((gw.double() - gw_ref).abs().max() / gw_ref.abs().max()),
                                           float((gb.double() - gb_ref).abs().max() / gb_ref.abs().max())))
            res["bwd_n%d_%s" % (n, tag)] = errs
            for k in env:
                os.environ.pop(k)
    print(json.dumps(res, indent=0))


if __name__ == "__main__":
    main()
