"""GPU sampler (pg_sample through the C-ABI) vs the CPU oracle: bit-exact NodeFlow arrays."""
import ctypes

import numpy as np
import pytest

import oracle
from conftest import random_in_csr

pytestmark = pytest.mark.gpu


def _gpu_sample(indptr, indices, eids, seeds, fanouts, seed=0, epoch=0, batch=0, cap_nodes=None, cap_edges=None):
    """Direct C-ABI call sequence: pg_graph_create -> pg_sampler_create -> pg_sample."""
    import torch
    from pagraph_b200 import _lib
    L = _lib.lib()
    V, E = len(indptr) - 1, len(indices)
    gh = ctypes.c_void_p()
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices = np.ascontiguousarray(indices, np.int64)
    e = None if eids is None else np.ascontiguousarray(eids, np.int64)
    _lib.check(L.pg_graph_create(indptr.ctypes.data, indices.ctypes.data, None if e is None else e.ctypes.data,
                                 V, E, 0, ctypes.byref(gh)), "pg_graph_create")
    n = len(seeds)
    if cap_nodes is None:
        cap_nodes, cap_edges, fr = max(n, 1), 1, max(n, 1)
        for f in fanouts:
            ed = min(fr * min(f, V), max(E, 1))
            fr = min(ed, V)
            cap_nodes += fr
            cap_edges += ed
    sh = ctypes.c_void_p()
    fan = (ctypes.c_int64 * len(fanouts))(*fanouts)
    _lib.check(L.pg_sampler_create(gh, len(fanouts), fan, seed, max(n, 1), cap_nodes, cap_edges, ctypes.byref(sh)),
               "pg_sampler_create")
    dev = "cuda:0"
    bufs = [torch.full((cap_nodes,), -7, dtype=torch.int64, device=dev),
            torch.full((cap_nodes + 1,), -7, dtype=torch.int64, device=dev),
            torch.full((cap_edges,), -7, dtype=torch.int64, device=dev),
            torch.full((cap_edges,), -7, dtype=torch.int64, device=dev),
            torch.zeros(_lib.PG_META_LEN, dtype=torch.int64, device=dev)]
    c = _lib.pg_nodeflow_buffers(*[_lib.ptr(b) for b in bufs])
    d_seeds = torch.from_numpy(np.ascontiguousarray(seeds, np.int64)).to(dev)
    h_meta = torch.zeros(_lib.PG_META_LEN, dtype=torch.int64).pin_memory()
    _lib.check(L.pg_sample(sh, _lib.ptr(d_seeds) if n else None, n, epoch, batch, ctypes.byref(c), _lib.ptr(h_meta),
                           _lib.stream_ptr()), "pg_sample")
    torch.cuda.synchronize()
    meta = h_meta.numpy().copy()
    assert np.array_equal(meta, bufs[4].cpu().numpy())
    L.pg_sampler_destroy(sh)
    L.pg_graph_destroy(gh)
    return meta, [b.cpu().numpy() for b in bufs[:4]]


def _assert_same(meta, arrs, ref):
    from pagraph_b200 import _lib
    assert meta[0] == _lib.PG_OK
    L1 = int(meta[3])
    assert L1 == ref.num_layers
    n, e = int(meta[1]), int(meta[2])
    np.testing.assert_array_equal(meta[4:4 + L1 + 1], ref.layer_offsets)
    np.testing.assert_array_equal(meta[4 + L1 + 1:4 + L1 + 1 + L1], ref.flow_offsets)
    np.testing.assert_array_equal(arrs[0][:n], ref.node_mapping)
    np.testing.assert_array_equal(arrs[1][:n + 1], ref.indptr)
    np.testing.assert_array_equal(arrs[2][:e], ref.indices)
    np.testing.assert_array_equal(arrs[3][:e], ref.edge_mapping)


@pytest.mark.parametrize("fanouts", [[2, 2], [5, 3], [3], [4, 2, 3], [25, 10], [1, 1, 1, 1]])
@pytest.mark.parametrize("with_eids", [True, False])
def test_sample_matches_oracle(fanouts, with_eids):
    indptr, indices, eids, _ = random_in_csr(3000, 60000, seed=1, with_eids=with_eids)
    seeds = np.random.default_rng(2).choice(3000, 257, replace=False)
    for batch in (0, 3):
        ref = oracle.sample(indptr, indices, eids, seeds, fanouts, seed=9, epoch=1, batch=batch)
        meta, arrs = _gpu_sample(indptr, indices, eids, seeds, fanouts, seed=9, epoch=1, batch=batch)
        _assert_same(meta, arrs, ref)


@pytest.mark.parametrize("fanouts", [[3, 3], [40, 7], [100, 90], [70]])
def test_sample_hub_graph_all_branches(fanouts):
    """Geometric in-degrees up to several hundred: take-all, direct (deg > 2k, incl. m > 64 -> global
    scratch) and complement (k < deg <= 2k) branches of GetUniformSample all occur."""
    indptr, indices, eids, _ = random_in_csr(1500, 90000, seed=5, hub=True)
    deg = np.diff(indptr)
    k = fanouts[0]
    assert (deg <= k).any() and (deg > 2 * k).any() and ((deg > k) & (deg <= 2 * k)).any()
    seeds = np.arange(0, 1500, 3)
    ref = oracle.sample(indptr, indices, eids, seeds, fanouts, seed=4, epoch=2, batch=11)
    meta, arrs = _gpu_sample(indptr, indices, eids, seeds, fanouts, seed=4, epoch=2, batch=11)
    _assert_same(meta, arrs, ref)


def test_sample_duplicate_seeds_keep_first_occurrence():
    indptr, indices, eids, _ = random_in_csr(400, 4000, seed=3)
    seeds = np.array([5, 9, 5, 7, 9, 9, 1, 399, 0, 399])
    ref = oracle.sample(indptr, indices, eids, seeds, [3, 3])
    meta, arrs = _gpu_sample(indptr, indices, eids, seeds, [3, 3])
    _assert_same(meta, arrs, ref)
    np.testing.assert_array_equal(ref.layer_parent_nid(-1), [5, 9, 7, 1, 399, 0])
    rng = np.random.default_rng(0)
    seeds = rng.integers(0, 400, 3000)           # heavy duplication, > one scan tile
    ref = oracle.sample(indptr, indices, eids, seeds, [2, 2])
    meta, arrs = _gpu_sample(indptr, indices, eids, seeds, [2, 2])
    _assert_same(meta, arrs, ref)


def test_sample_empty_seed_batch_and_isolated_vertices():
    indptr, indices, eids, _ = random_in_csr(100, 300, seed=3)
    ref = oracle.sample(indptr, indices, eids, np.zeros(0, np.int64), [3, 3])
    meta, arrs = _gpu_sample(indptr, indices, eids, np.zeros(0, np.int64), [3, 3])
    _assert_same(meta, arrs, ref)
    iso = np.where(np.diff(indptr) == 0)[0]
    assert len(iso) > 0
    ref = oracle.sample(indptr, indices, eids, iso, [3, 3])
    meta, arrs = _gpu_sample(indptr, indices, eids, iso, [3, 3])
    _assert_same(meta, arrs, ref)
    assert meta[2] == 0


def test_sample_full_fanout_is_rng_free_closure():
    """fanout >= max degree — the reference's own deterministic usage (partition/utils.py:11-18)."""
    indptr, indices, eids, coo = random_in_csr(800, 9000, seed=4)
    seeds = np.array([3, 17, 42, 99, 150, 799])
    a = _gpu_sample(indptr, indices, eids, seeds, [800, 800], seed=0, epoch=0, batch=0)
    b = _gpu_sample(indptr, indices, eids, seeds, [800, 800], seed=5, epoch=2, batch=7)
    ref = oracle.sample(indptr, indices, eids, seeds, [800, 800])
    _assert_same(a[0], a[1], ref)
    _assert_same(b[0], b[1], ref)
    csc = coo.tocsc()
    hop1 = np.unique(np.concatenate([csc.indices[csc.indptr[v]:csc.indptr[v + 1]] for v in seeds]))
    n = int(a[0][1])
    lo = a[0][4:8]
    np.testing.assert_array_equal(a[1][0][lo[1]:lo[2]], hop1)
    assert n == lo[3]


def test_sample_overflow_is_reported_not_written():
    from pagraph_b200 import _lib
    indptr, indices, eids, _ = random_in_csr(1000, 20000, seed=7)
    seeds = np.arange(200)
    ref = oracle.sample(indptr, indices, eids, seeds, [10, 10])
    meta, arrs = _gpu_sample(indptr, indices, eids, seeds, [10, 10], cap_nodes=300, cap_edges=500)
    assert meta[0] == _lib.PG_ERR_OVERFLOW
    assert (arrs[0] == -7).all() and (arrs[2] == -7).all()      # outputs untouched
    # and the Python sampler regrows transparently
    import torch
    from pagraph_b200 import DGLGraph
    from pagraph_b200.sampling import NeighborSampler
    g = DGLGraph.from_in_csr(indptr, indices, eids)
    s = NeighborSampler(g, 200, [10, 10], num_hops=2, seed_nodes=torch.from_numpy(seeds))
    s._cap_nodes, s._cap_edges = 300, 500
    s._create_handle()
    nf = s.sample_batch(0)
    np.testing.assert_array_equal(nf._node_mapping.tousertensor().cpu().numpy(), ref.node_mapping)
    np.testing.assert_array_equal(nf._indices.cpu().numpy(), ref.indices)


def test_python_sampler_iterates_batches_like_the_oracle():
    """NeighborSampler (the drop-in for dgl.contrib.sampling.NeighborSampler) over 2 epochs, with
    prefetch: batch k of epoch e == oracle.sample(seeds[k*B:(k+1)*B], epoch=e, batch=k)."""
    import torch
    from pagraph_b200 import DGLGraph
    from pagraph_b200.sampling import NeighborSampler
    indptr, indices, eids, coo = random_in_csr(2000, 30000, seed=8)
    g = DGLGraph(coo, readonly=True)
    np.testing.assert_array_equal(g.indptr, indptr)
    np.testing.assert_array_equal(g.indices, indices)
    np.testing.assert_array_equal(g.eids, eids)
    train = np.random.default_rng(1).choice(2000, 700, replace=False)
    torch.manual_seed(3)
    s = NeighborSampler(g, 128, 4, num_hops=2, neighbor_type='in', shuffle=True, num_workers=16,
                        seed_nodes=torch.from_numpy(train), prefetch=True, seed=21)
    order = s._seeds_cpu.numpy()
    assert sorted(order.tolist()) == sorted(train.tolist()) and not np.array_equal(order, train)
    assert len(s) == 6
    for epoch in range(2):
        count = 0
        for k, nf in enumerate(s):
            ref = oracle.sample(indptr, indices, eids, order[k * 128:(k + 1) * 128], [4, 4], seed=21, epoch=epoch,
                                batch=k)
            assert nf._layer_offsets == ref.layer_offsets.tolist()
            assert nf._block_offsets == ref.flow_offsets.tolist()
            np.testing.assert_array_equal(nf._node_mapping.tousertensor().cpu().numpy(), ref.node_mapping)
            np.testing.assert_array_equal(nf._indptr.cpu().numpy(), ref.indptr)
            np.testing.assert_array_equal(nf._indices.cpu().numpy(), ref.indices)
            np.testing.assert_array_equal(nf._edge_mapping.tousertensor().cpu().numpy(), ref.edge_mapping)
            np.testing.assert_array_equal(nf.layer_parent_nid(-1).numpy(), ref.layer_parent_nid(-1))
            np.testing.assert_array_equal(nf.layer_parent_nid(0).numpy(), ref.layer_parent_nid(0))
            src, dst, eid = nf.block_edges(1)
            np.testing.assert_array_equal(src.numpy(), ref.indices[ref.flow_offsets[1]:ref.flow_offsets[2]])
            np.testing.assert_array_equal(nf.map_to_parent_nid(src).numpy(), ref.node_mapping[src.numpy()])
            count += 1
        assert count == 6


def test_graph_degrees():
    import torch
    from pagraph_b200 import DGLGraph
    indptr, indices, eids, coo = random_in_csr(500, 5000, seed=9)
    dev_g = DGLGraph.from_in_csr(torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda())
    np.testing.assert_array_equal(dev_g.out_degrees().numpy(), np.bincount(coo.row, minlength=500))
    np.testing.assert_array_equal(dev_g.in_degrees().numpy(), np.diff(indptr))
    host_g = DGLGraph(coo)
    np.testing.assert_array_equal(host_g.out_degrees().numpy(), np.bincount(coo.row, minlength=500))


def test_sample_offsets_beyond_2_31():
    """64-bit paths: CSR offsets, edge ids and one in-degree above 2^31 (configs 4/5 of BASELINE.json have nnz > 2^31).
    Vertex 0 owns a 2^31 + 7 entry row of zeros (a 17 GB `indices` array on the device, calloc'ed — untouched — on the
    host); every other row lives behind it, so each offset the sampler touches needs more than 32 bits, and expanding
    vertex 0 draws positions in [0, 2^31 + 7)."""
    import torch
    from pagraph_b200 import _lib
    free, _ = torch.cuda.mem_get_info()
    big = (1 << 31) + 7
    if free < (big + (1 << 20)) * 8 + (4 << 30):
        pytest.skip("needs ~21 GB of free device memory")
    rng = np.random.default_rng(11)
    V = 2000
    deg = rng.integers(0, 30, V)
    deg[0] = big
    deg[1:21] = 3                                                   # short rows (taken whole) that will name vertex 0
    indptr = np.zeros(V + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    nnz = int(indptr[-1])
    tail = rng.integers(0, V, nnz - big).astype(np.int64)          # neighbours of vertices 1..V-1 (vertex 0 among them)
    tail[indptr[1:21] - big] = 0
    try:
        indices = np.zeros(nnz, np.int64)                             # lazily mapped zero pages on the host
    except MemoryError:
        pytest.skip("host cannot map a 17 GB array")
    indices[big:] = tail
    d_indices = torch.zeros(nnz, dtype=torch.int64, device="cuda")
    d_indices[big:] = torch.from_numpy(tail).cuda()
    d_indptr = torch.from_numpy(indptr).cuda()
    L = _lib.lib()
    gh, sh = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.pg_graph_create_device(_lib.ptr(d_indptr), _lib.ptr(d_indices), None, V, nnz, 0, ctypes.byref(gh)),
               "pg_graph_create_device")
    seeds = np.concatenate([np.arange(1, 21), rng.choice(np.arange(21, V), 280, replace=False)]).astype(np.int64)
    fanouts = [5, 4]
    cap_nodes, cap_edges = 300 + 1500 + 6000, 1500 + 6000
    fan = (ctypes.c_int64 * 2)(*fanouts)
    _lib.check(L.pg_sampler_create(gh, 2, fan, 13, 300, cap_nodes, cap_edges, ctypes.byref(sh)), "pg_sampler_create")
    bufs = [torch.zeros(cap_nodes, dtype=torch.int64, device="cuda"), torch.zeros(cap_nodes + 1, dtype=torch.int64, device="cuda"),
            torch.zeros(cap_edges, dtype=torch.int64, device="cuda"), torch.zeros(cap_edges, dtype=torch.int64, device="cuda"),
            torch.zeros(_lib.PG_META_LEN, dtype=torch.int64, device="cuda")]
    c = _lib.pg_nodeflow_buffers(*[_lib.ptr(b) for b in bufs])
    d_seeds = torch.from_numpy(seeds).cuda()
    try:
        _lib.check(L.pg_sample(sh, _lib.ptr(d_seeds), len(seeds), 0, 2, ctypes.byref(c), None, _lib.stream_ptr()), "pg_sample")
        torch.cuda.synchronize()
        meta = bufs[4].cpu().numpy()
        ref = oracle.sample(indptr, indices, None, seeds, fanouts, seed=13, epoch=0, batch=2)
        _assert_same(meta, [b.cpu().numpy() for b in bufs[:4]], ref)
        assert ref.edge_mapping.max() > (1 << 31)                     # parent edge ids = CSR positions beyond 2^31
        assert 0 in ref.layer_parent_nid(1) and (ref.edge_mapping < big).any()   # vertex 0 was expanded by random draws
    finally:
        L.pg_sampler_destroy(sh)
        L.pg_graph_destroy(gh)
