"""Peer-GPU cache tier: two processes / two GPUs pool their feature caches over NVLink. Rows fetched through
the tier are bit-identical to the host table, the fused block-0 aggregate matches the float64 oracle, peer hits replace
most PCIe misses, and an engine trained on it matches the eager loop. Skipped with fewer than 2 GPUs."""
import numpy as np
import pytest

from conftest import random_in_csr

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    import oracle
    from pagraph_b200 import DGLGraph, ops
    from pagraph_b200 import graph_store as gs
    from pagraph_b200.sampling import NeighborSampler
    from pagraph_b200.storage import GraphCacheServer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    V, F = 6000, 600
    rng = np.random.default_rng(7)
    indptr, indices, eids, _ = random_in_csr(V, 90000, 7, hub=True)
    feats = rng.random((V, F), dtype=np.float32)
    store = gs.LocalGraphStore(name="peer%d" % rank)
    store.ndata["features"] = torch.from_numpy(feats)
    g = DGLGraph.from_in_csr(indptr, indices, eids)
    cs = GraphCacheServer(store, V, torch.arange(V), rank)
    cs.init_field(["features"])
    cs.log = True
    cap = 1500                                   # 25 % per rank: two ranks pool 50 % of the vertices
    cs.auto_cache_peers(g, ["features"], capability=cap)
    assert cs.peer_tier["world"] == 2 and cs.peer_tier["local_rows"] == 0 and cs.peer_tier["shard_rows"] == cap
    flag = cs.gpu_flag.cpu().numpy()
    assert flag.sum() == cap                     # same memory as auto_cache, but a different half of the top 3000 on each rank
    order = np.argsort(-np.bincount(indices, minlength=V), kind="stable")
    np.testing.assert_array_equal(np.sort(np.nonzero(flag)[0]), np.sort(order[rank:2 * cap:2]))
    seeds = rng.choice(V, 300, replace=False).astype(np.int64)
    sampler = NeighborSampler(g, 300, [8, 6], num_hops=2, seed_nodes=torch.from_numpy(seeds), seed=3)
    nf = sampler.sample_batch(0)
    ref = oracle.sample(indptr, indices, eids, seeds, [8, 6], seed=3)
    ids0 = ref.layer_parent_nid(0)
    # fetch_data semantics unchanged: every row bit-exact, non-local rows come from the host
    cs.fetch_data(nf)
    assert np.array_equal(nf.layers[0].data["features"].cpu().numpy(), feats[ids0])
    cs.get_miss_rate()
    # fused path: local rows + peer rows over NVLink + host rows
    bi, bc, bb, n_dst, n_src = nf.block_csr(0)
    got = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean").cpu().numpy()
    ip, cols, base = ref.block(0)
    want = oracle.aggregate(ip, cols, base, feats[ids0], "mean")
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-30)
    pooled = np.zeros(V, bool)
    pooled[order[:2 * cap]] = True
    peer_expected = int((pooled[ids0] & ~flag[ids0]).sum())
    assert cs.peer_hits() == peer_expected and peer_expected > 0
    assert cs.try_num == len(ids0) and cs.miss_num == int((~pooled[ids0]).sum())      # only unpooled rows go to PCIe
    out[rank] = (peer_expected, int((~pooled[ids0]).sum()), len(ids0))
    dist.barrier()
    del cs
    dist.destroy_process_group()


def test_peer_cache_tier_two_ranks():
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for peer, miss, n in res.values():
        assert peer > 0.1 * n and miss < 0.6 * n
