"""Fused cache lookup + dropout + aggregation (pg_cache_aggregate, GraphCacheServer.lazy_input) vs the oracle:
the block-0 result must equal fetch-then-aggregate (<= 1e-5 relative; tolerance because the fp32 summation
order differs from the float64 oracle), for fully cached, partially cached and cold caches, with and without
the dropout mask, through the C-ABI and through the model."""
import numpy as np
import pytest

import oracle
from conftest import random_in_csr

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _setup(dims, V=3000, nnz=40000, cap=None, seeds=200, fanouts=(7, 5), seed=3):
    import torch
    from pagraph_b200 import DGLGraph
    from pagraph_b200.graph_store import LocalGraphStore
    from pagraph_b200.sampling import NeighborSampler
    from pagraph_b200.storage import GraphCacheServer
    rng = np.random.default_rng(seed)
    indptr, indices, eids, _ = random_in_csr(V, nnz, seed)
    V_full = V + 500
    nid_map = np.sort(rng.choice(V_full, V, replace=False)).astype(np.int64)
    host = {n: rng.random((V_full, d), dtype=np.float32) for n, d in dims.items()}
    store = LocalGraphStore(name="fused")
    for k, v in host.items():
        store.ndata[k] = torch.from_numpy(v)
    g = DGLGraph.from_in_csr(indptr, indices, eids)
    cs = GraphCacheServer(store, V, torch.from_numpy(nid_map), 0)
    cs.init_field(list(dims))
    if cap is not None:
        cs.auto_cache(g, list(dims), capability=cap)
    seed_ids = rng.choice(V, seeds, replace=False).astype(np.int64)
    sampler = NeighborSampler(g, seeds, list(fanouts), num_hops=len(fanouts), seed_nodes=torch.from_numpy(seed_ids), seed=9)
    nf = sampler.sample_batch(0)
    ref = oracle.sample(indptr, indices, eids, seed_ids, list(fanouts), seed=9)
    return cs, nf, ref, host, nid_map, store


def _close(got, want):
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-30)


@pytest.mark.parametrize("cap", [None, 600, 10 ** 9])          # cold (all rows from the host), partial, full_cached
@pytest.mark.parametrize("dims", [{"features": 600, "norm": 1}, {"features": 602}, {"features": 128}, {"features": 64},
                                  {"features": 1100}])
@pytest.mark.parametrize("mode", ["mean", "sum"])
def test_fused_matches_fetch_then_aggregate(dims, cap, mode):
    from pagraph_b200 import ops
    cs, nf, ref, host, nid_map, _ = _setup(dims, cap=cap)
    ip, cols, base = ref.block(0)
    src = host["features"][nid_map[ref.layer_parent_nid(0)]]
    want = oracle.aggregate(ip, cols, base, src, mode)
    bi, bc, bb, n_dst, n_src = nf.block_csr(0)
    cs.log = True
    got = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, mode)
    _close(got.cpu().numpy(), want)
    if not cs.full_cached:
        flag = cs.gpu_flag.cpu().numpy()
        assert cs.try_num == n_src and cs.miss_num == int((~flag[ref.layer_parent_nid(0)]).sum())


@pytest.mark.parametrize("cap", [600, 10 ** 9])
@pytest.mark.parametrize("hot_mb,hint", [("0.12", "1"), ("0", "1"), ("40", "0")])
def test_fused_l2_reuse_hint_does_not_change_results(cap, hot_mb, hint, monkeypatch):
    """pg_cache_set_hot tags the row pointers of the top-out-degree rows (bit 0) and the row-fetch kernels mask the tag
    off: a budget of 50 rows mixes tagged and untagged pointers in one block; results equal the oracle either way."""
    from pagraph_b200 import ops
    monkeypatch.setenv("PG_CACHE_HOT_MB", hot_mb)
    monkeypatch.setenv("PG_AGG_L2HINT", hint)
    dims = {"features": 600, "norm": 1}
    cs, nf, ref, host, nid_map, _ = _setup(dims, cap=cap)
    n_hot = int(cs._hot.sum().item())
    assert n_hot == int(float(hot_mb) * 1e6 // (601 * 4)) or n_hot == min(cs.node_num, cs.cached_num if not cs.full_cached else cs.node_num)
    ip, cols, base = ref.block(0)
    src = host["features"][nid_map[ref.layer_parent_nid(0)]]
    bi, bc, bb, n_dst, n_src = nf.block_csr(0)
    for mode in ("mean", "sum"):
        got = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, mode)
        _close(got.cpu().numpy(), oracle.aggregate(ip, cols, base, src, mode))
    got = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean", dropout_p=0.2, seed=4)
    keep = oracle.dropout_keep_mask(4, len(src), 600, 0.2)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(0.2))
    _close(got.cpu().numpy(), oracle.aggregate(ip, cols, base, np.where(keep, src * scale, np.float32(0)), "mean"))


@pytest.mark.parametrize("dim,p", [(600, 0.2), (602, 0.5), (64, 0.2)])
def test_fused_dropout_mask_contract(dim, p):
    """Same mask for every edge of a source node, keyed by (seed, node, column): equals dropout-then-aggregate."""
    import torch
    from pagraph_b200 import ops
    cs, nf, ref, host, nid_map, _ = _setup({"features": dim}, cap=700)
    ip, cols, base = ref.block(0)
    src = host["features"][nid_map[ref.layer_parent_nid(0)]]
    seed = 0xDEADBEEF12345
    keep = oracle.dropout_keep_mask(seed, src.shape[0], dim, p)
    assert abs(keep.mean() - (1 - p)) < 0.01
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    want = oracle.aggregate(ip, cols, base, np.where(keep, src * scale, np.float32(0)), "mean")
    bi, bc, bb, n_dst, n_src = nf.block_csr(0)
    got = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean",
                              dropout_p=p, seed=seed)
    _close(got.cpu().numpy(), want)
    step = torch.tensor([41], dtype=torch.int64, device="cuda")       # device-resident seed offset (graph replays)
    got2 = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "mean",
                               dropout_p=p, seed=seed - 41, step=step)
    np.testing.assert_array_equal(got2.cpu().numpy(), got.cpu().numpy())


def test_fused_norm_padding_and_errors():
    import ctypes
    import torch
    from pagraph_b200 import _lib, ops
    cs, nf, ref, host, nid_map, _ = _setup({"features": 600, "norm": 1}, cap=500)
    ip, cols, base = ref.block(0)
    src = host["features"][nid_map[ref.layer_parent_nid(0)]]
    bi, bc, bb, n_dst, n_src = nf.block_csr(0)
    norm = torch.rand(n_dst, device="cuda")
    out = torch.full((n_dst + 37, 604), -1.0, device="cuda")[:, :600]
    got = ops.cache_aggregate(cs, "features", nf.layer_parent_nid_dev(0), bi, bc, bb, n_src, n_dst, "sum", norm=norm,
                              out=out, zero_rows_to=n_dst + 37)
    want = oracle.aggregate(ip, cols, base, src, "sum") * norm.cpu().numpy()[:, None]
    _close(got[:n_dst].cpu().numpy(), want)
    assert (got[n_dst:] == 0).all()
    blk = _lib.pg_block(_lib.ptr(nf.layer_parent_nid_dev(0)), _lib.ptr(bi), _lib.ptr(bc), bb, n_src, n_dst)
    L = _lib.lib()
    assert L.pg_cache_aggregate(cs._handle, 5, ctypes.byref(blk), _lib.ptr(out), 604, 0, None, 0.0, 0, None, 0, None,
                                None) == _lib.PG_ERR_INVALID
    assert L.pg_cache_aggregate(cs._handle, 0, ctypes.byref(blk), _lib.ptr(out), 604, 0, None, 1.0, 0, None, 0, None,
                                None) == _lib.PG_ERR_INVALID


@pytest.mark.parametrize("cap", [500, 10 ** 9])
def test_gcn_forward_lazy_input_equals_eager(cap):
    """GCNSampling over a lazy input layer (fused kernel) == the eager fetch_data path; other layers identical."""
    import torch
    from pagraph_b200.model.gcn_nssc import GCNSampling
    from pagraph_b200.storage import LazyCacheRows
    cs, nf, ref, host, nid_map, _ = _setup({"features": 600, "norm": 1}, cap=cap)
    torch.manual_seed(0)
    model = GCNSampling(600, 32, 10, 1, torch.relu, 0.0).cuda()
    cs.fetch_data(nf)
    eager = model(nf).detach().cpu().numpy()
    feats1 = nf.layers[1].data["features"].cpu().numpy()
    cs.lazy_input = True
    cs.fetch_data(nf)
    assert isinstance(nf.layers[0].data["features"], LazyCacheRows)
    np.testing.assert_array_equal(nf.layers[1].data["features"].cpu().numpy(), feats1)
    np.testing.assert_array_equal(nf.layers[0].data["features"].materialize().cpu().numpy(),
                                  host["features"][nid_map[ref.layer_parent_nid(0)]])
    cs.fetch_data(nf)
    pred = model(nf)
    pred.square().mean().backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), eager, rtol=1e-4, atol=1e-6)
    # training-mode dropout goes through the fused mask: finite, different from eval, same shape
    model_d = GCNSampling(600, 32, 10, 1, torch.relu, 0.5).cuda()
    model_d.load_state_dict(model.state_dict())
    cs.fetch_data(nf)
    a = model_d(nf)
    model_d.eval()
    cs.fetch_data(nf)
    b = model_d(nf)
    assert torch.isfinite(a).all() and not torch.allclose(a, b)
    np.testing.assert_allclose(b.detach().cpu().numpy(), eager, rtol=1e-4, atol=1e-6)


def test_graphsage_forward_lazy_input_equals_eager():
    import torch
    from pagraph_b200.model.graphsage_nssc import GraphSageSampling
    cs, nf, ref, host, nid_map, _ = _setup({"features": 128}, cap=800)
    torch.manual_seed(0)
    model = GraphSageSampling(128, 16, 10, 1, torch.relu, 0.0, "mean").cuda()
    cs.fetch_data(nf)
    eager = model(nf).detach().cpu().numpy()
    cs.lazy_input = True
    cs.fetch_data(nf)
    np.testing.assert_allclose(model(nf).detach().cpu().numpy(), eager, rtol=1e-4, atol=1e-6)
