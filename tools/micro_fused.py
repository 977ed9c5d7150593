#!/usr/bin/env python
"""Micro-benchmark at bench.py scale: fused cache-aggregate (pg_cache_aggregate) vs fetch + dropout + aggregate for
block 0, for both cache modes, sweeping the TMA staging parameters. CUDA-event timed over distinct minibatches.
    python tools/micro_fused.py [--iters 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def timeit(fn, nfs):
    evs = []
    for nf in nfs:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(nf)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", default="2:200,3:200,2:120,4:200,1:200", help="depth:smem_kb[:warps],...")
    ap.add_argument("--modes", default="hbm20,vtx20")
    ap.add_argument("--quick", action="store_true", help="only the fused_only sweep")
    ap.add_argument("--hot-sweep", default="", help="hbm20 only: L2 reuse-hint budgets in MB (PG_CACHE_HOT_MB), e.g. 0,20,40,80; "
                                                    "'off' = no eviction priorities at all (PG_AGG_L2HINT=0)")
    a = ap.parse_args()
    sys.argv = [sys.argv[0]]
    args = bench.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from pagraph_b200 import ops
    from pagraph_b200.sampling import NeighborSampler
    from pagraph_b200.storage import GraphCacheServer
    wl = bench.Workload(args, 0, 1, dev)
    sampler = NeighborSampler(wl.g, args.batch_size, wl.fanouts, num_hops=2, seed_nodes=torch.from_numpy(wl.train_nid),
                              shuffle=True, seed=1)
    nfs = [sampler.sample_batch(k) for k in range(a.iters)]
    out = {}
    for mode, cap in (("hbm20", args.vnum), ("vtx20", args.vnum // 5)):
        if mode not in a.modes.split(","):
            continue
        cs = GraphCacheServer(wl.store, args.vnum, torch.arange(args.vnum), 0)
        cs.init_field(["features", "norm"])
        cs.auto_cache(wl.g, ["features", "norm"], capability=cap)

        def unfused(nf):
            cs.lazy_input = False
            cs.fetch_data(nf)
            h = torch.nn.functional.dropout(nf.layers[0].data["features"], 0.2, True)
            bi, bc, bb, n_dst, n_src = nf.block_csr(0)
            return ops.aggregate_forward(bi, bc, bb, h, n_dst, "mean")

        def fused(nf, p=0.2):
            cs.lazy_input = True
            cs.fetch_data(nf)                      # layers 1..L eagerly
            bi, bc, bb, n_dst, n_src = nf.block_csr(0)
            return ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean",
                                       dropout_p=p, seed=7)

        def fused_only(nf, p=0.2):
            bi, bc, bb, n_dst, n_src = nf.block_csr(0)
            return ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean",
                                       dropout_p=p, seed=7)

        res = {}
        if a.hot_sweep and mode == "hbm20":
            for hb in a.hot_sweep.split(","):
                # "off" = no priorities; "N" = N MB hot (evict_last), rest evict_first; "Nn" = rest default priority
                os.environ["PG_AGG_L2HINT"] = "0" if hb == "off" else ("2" if hb.endswith("n") else "1")
                os.environ["PG_CACHE_HOT_MB"] = "0" if hb == "off" else hb.rstrip("n")
                cs._mark_hot(wl.g, None)
                timeit(fused_only, nfs[:3])
                res["fused_only_hot%s_ms" % hb] = [timeit(fused_only, nfs), timeit(fused_only, nfs)]
            os.environ.pop("PG_AGG_L2HINT"), os.environ.pop("PG_CACHE_HOT_MB")
            cs._mark_hot(wl.g, None)
        if not a.quick:
            timeit(unfused, nfs[:3])
            res["unfused_fetch+dropout+agg_ms"] = timeit(unfused, nfs)
        for sw in a.sweep.split(","):
            depth, kb, *rest = sw.split(":")
            os.environ["PG_AGG_DEPTH"], os.environ["PG_AGG_SMEM"] = depth, str(int(kb) * 1024)
            if rest:
                os.environ["PG_AGG_WARPS"] = rest[0]
            tag = "depth%s_smem%sk%s" % (depth, kb, "_w" + rest[0] if rest else "")
            timeit(fused_only, nfs[:3])
            res["fused_only_%s_ms" % tag] = timeit(fused_only, nfs)
            res["fused_only_nodrop_%s_ms" % tag] = timeit(lambda nf: fused_only(nf, 0.0), nfs)
        os.environ.pop("PG_AGG_DEPTH"), os.environ.pop("PG_AGG_SMEM"), os.environ.pop("PG_AGG_WARPS", None)
        if a.quick:
            lo, bo = nfs[0]._layer_offsets, nfs[0]._block_offsets
            res["n0,n1,E0"] = [lo[1] - lo[0], lo[2] - lo[1], bo[1] - bo[0]]
            out[mode] = res
            continue
        res["fused+fetch_other_layers_ms"] = timeit(fused, nfs)
        os.environ["PG_AGG_NO_TMA"] = "1"
        timeit(fused_only, nfs[:2])
        res["fused_only_ldg_fallback_ms"] = timeit(fused_only, nfs)
        os.environ.pop("PG_AGG_NO_TMA")
        lo, bo = nfs[0]._layer_offsets, nfs[0]._block_offsets
        res["n0,n1,E0"] = [lo[1] - lo[0], lo[2] - lo[1], bo[1] - bo[0]]
        res["read_MB(E0 rows)+write_MB"] = (bo[1] - bo[0]) * 2400 / 1e6 + (lo[2] - lo[1]) * 2400 / 1e6
        out[mode] = res
        del cs
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
