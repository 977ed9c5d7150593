"""GPU cache lookup / hit-miss split / gather (pg_cache_* through the C-ABI and through the
GraphCacheServer drop-in) vs golden vectors from the real reference storage.py and vs the oracle.
Payload is copied bit-for-bit: every comparison is exact."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

STORAGE_CASES = ["storage_gcn_partial", "storage_ties_odd", "storage_full", "storage_sage4"]


class _FakeNF:
    """The members GraphCacheServer.fetch_data touches (storage.py:171-173,202)."""

    def __init__(self, node_mapping, layer_offsets, device=None):
        import torch
        from pagraph_b200.nodeflow import _Index
        t = torch.from_numpy(np.ascontiguousarray(node_mapping, np.int64))
        self._node_mapping = _Index(t if device is None else t.to(device))
        self._layer_offsets = [int(x) for x in layer_offsets]
        self.num_layers = len(layer_offsets) - 1
        self._node_frames = [None] * self.num_layers

    def layer_parent_nid(self, i):
        i %= self.num_layers
        return self._node_mapping.tousertensor()[self._layer_offsets[i]:self._layer_offsets[i + 1]].cpu()


class _LocalG:
    def __init__(self, out_deg):
        import torch
        self._d = torch.from_numpy(np.asarray(out_deg, np.int64))

    def out_degrees(self):
        return self._d


def _store(fields):
    import torch
    from pagraph_b200.graph_store import LocalGraphStore
    s = LocalGraphStore(name="t")
    for k, v in fields.items():
        s.ndata[k] = torch.from_numpy(v)
    return s


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("case", STORAGE_CASES)
def test_cache_server_replays_reference_golden(case, mode):
    """Same call sequence the golden generator ran against the real reference class."""
    import torch
    from pagraph_b200.storage import GraphCacheServer
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    names = [str(x) for x in g["field_names"]]
    store = _store({n: g["host_" + n] for n in names})
    cs = GraphCacheServer(store, len(g["nid_map"]), torch.from_numpy(g["nid_map"]), 0)
    cs.fetch_mode = mode
    cs.init_field(names)
    assert cs.total_dim == int(g["total_dim"])
    cs.log = True
    nf = _FakeNF(g["node_mapping"], g["layer_offsets"])          # CPU ids, like dgl's NodeFlow
    cs.fetch_data(nf)
    for i in range(nf.num_layers):
        for n in names:
            got = nf._node_frames[i][n]
            assert got.is_cuda and got.dtype == torch.float32
            np.testing.assert_array_equal(got.cpu().numpy(), g["cold_l%d_%s" % (i, n)])
    assert cs.get_miss_rate() == float(g["cold_miss_rate"])
    cs.auto_cache(_LocalG(g["out_deg"]), names, capability=int(g["capability"]))
    assert cs.full_cached == bool(g["full_cached"]) and cs.cached_num == int(g["cached_num"])
    flag = cs.gpu_flag.cpu().numpy()
    np.testing.assert_array_equal(flag, g["gpu_flag"])           # the cache hit SET
    l2c = cs.localid2cacheid.cpu().numpy()
    for n in names:                                              # same rows cached
        ref = g["cache_" + n][g["l2c_on_cached"][g["gpu_flag"]]]
        np.testing.assert_array_equal(cs.gpu_fix_cache[n].cpu().numpy()[l2c[flag]], ref)
    nf = _FakeNF(g["node_mapping"], g["layer_offsets"], device="cuda:0")
    cs.keep_hit_mask = True
    cs.fetch_data(nf)
    for i in range(nf.num_layers):
        for n in names:
            np.testing.assert_array_equal(nf._node_frames[i][n].cpu().numpy(), g["warm_l%d_%s" % (i, n)])
    np.testing.assert_array_equal(cs.last_hit_mask.cpu().numpy(), g["gpu_flag"][g["node_mapping"]])
    if not cs.full_cached:
        assert cs.miss_num == int(g["warm_miss_num"]) and cs.try_num == int(g["warm_try_num"])
        assert cs.get_miss_rate() == float(g["warm_miss_rate"])
        with pytest.raises(ZeroDivisionError):
            cs.get_miss_rate()
    else:
        cs.fetch_from_cache(nf)
        for i in range(nf.num_layers):
            for n in names:
                np.testing.assert_array_equal(nf._node_frames[i][n].cpu().numpy(), g["warm_l%d_%s" % (i, n)])


@pytest.mark.parametrize("dims", [{"features": 600, "norm": 1}, {"features": 602, "norm": 1}, {"features": 600},
                                  {"features": 128}, {"features": 600, "neigh": 600}, {"features": 7},
                                  {"features": 2052}])
@pytest.mark.parametrize("mode", [1, 2])
def test_fetch_matches_oracle_random(dims, mode):
    import torch
    from pagraph_b200.storage import GraphCacheServer
    rng = np.random.default_rng(42)
    V_full, V_sub, N = 5000, 3500, 20000
    host = {n: rng.random((V_full, d), dtype=np.float32) for n, d in dims.items()}
    nid_map = np.sort(rng.choice(V_full, V_sub, replace=False)).astype(np.int64)
    out_deg = rng.integers(0, 50, V_sub)
    names = list(dims)
    oc = oracle.OracleCache(host, V_sub, nid_map)
    oc.init_field(names)
    oc.auto_cache(out_deg, names, 700)
    cs = GraphCacheServer(_store(host), V_sub, torch.from_numpy(nid_map), 0)
    cs.fetch_mode = mode
    cs.init_field(names)
    cs.auto_cache(_LocalG(out_deg), names, capability=700)
    np.testing.assert_array_equal(cs.gpu_flag.cpu().numpy(), oc.gpu_flag)
    np.testing.assert_array_equal(cs.localid2cacheid.cpu().numpy()[oc.gpu_flag], oc.localid2cacheid[oc.gpu_flag])
    ids = rng.integers(0, V_sub, N)                          # duplicates allowed
    offs = [0, 12000, 17000, N]
    nf = _FakeNF(ids, offs, device="cuda:0")
    cs.log = True
    cs.keep_hit_mask = True
    cs.fetch_data(nf)
    for n in names:
        want, mask, miss = oracle.fetch_c(ids, oc.gpu_flag, oc.localid2cacheid, oc.nid_map, oc.gpu_fix_cache[n],
                                          host[n])
        got = np.concatenate([nf._node_frames[i][n].cpu().numpy() for i in range(3)])
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(cs.last_hit_mask.cpu().numpy(), mask)
        assert cs.miss_num == miss and cs.try_num == N
    # get_feat_from_server: GPU pull from pinned host rows and the CPU path agree with the oracle
    q = torch.from_numpy(rng.integers(0, V_sub, 333)).cuda()
    a = cs.get_feat_from_server(q, names, to_gpu=True)
    b = cs.get_feat_from_server(q, names)
    for n in names:
        want = host[n][nid_map[q.cpu().numpy()]]
        np.testing.assert_array_equal(a[n].cpu().numpy(), want)
        np.testing.assert_array_equal(b[n].numpy(), want)
        assert not b[n].is_cuda


def test_cache_fix_data_installs_caller_rows():
    """cache_fix_data(nids, data) (storage.py:135-154) with caller-supplied rows."""
    import torch
    from pagraph_b200.storage import GraphCacheServer
    rng = np.random.default_rng(1)
    host = {"features": rng.random((300, 16), dtype=np.float32)}
    cs = GraphCacheServer(_store(host), 300, torch.arange(300), 0)
    cs.init_field(["features"])
    nids = torch.from_numpy(rng.choice(300, 40, replace=False)).cuda()
    fake_rows = torch.full((40, 16), 5.0).cuda() + torch.arange(40).cuda()[:, None]
    cs.cache_fix_data(nids, {"features": fake_rows})
    with pytest.raises(AssertionError):
        cs.cache_fix_data(nids, {"features": fake_rows[:39]})
    ids = np.arange(300)
    nf = _FakeNF(ids, [0, 300], device="cuda:0")
    cs.fetch_data(nf)
    got = nf._node_frames[0]["features"].cpu().numpy()
    want = host["features"].copy()
    want[nids.cpu().numpy()] = fake_rows.cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_raw_abi_fetch_with_row_strides_and_inf():
    """Direct C-ABI: padded host row stride (602 -> 604), inf/nan payload preserved bit-for-bit."""
    import torch
    from pagraph_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    V, dim, stride = 1000, 602, 604
    p = ctypes.c_void_p()
    _lib.check(L.pg_host_alloc(ctypes.byref(p), V * stride * 4), "pg_host_alloc")
    host = np.ctypeslib.as_array((ctypes.c_float * (V * stride)).from_address(p.value)).reshape(V, stride)
    host[:] = rng.random((V, stride), dtype=np.float32)
    host[5, 3] = np.inf
    host[6, 0] = np.nan
    host.view(np.uint32)[7, 1] = 0x7fc12345                      # NaN payload must survive
    flag = torch.zeros(V, dtype=torch.uint8, device="cuda")
    l2c = torch.zeros(V, dtype=torch.int64, device="cuda")
    nid_map = torch.from_numpy(rng.permutation(V)).cuda()
    f = (_lib.pg_field * 1)()
    f[0].dim, f[0].host_stride, f[0].host_table = dim, stride, p.value
    h = ctypes.c_void_p()
    _lib.check(L.pg_cache_create(V, _lib.ptr(flag), _lib.ptr(l2c), _lib.ptr(nid_map), 1, f, 0, ctypes.byref(h)),
               "pg_cache_create")
    cached = torch.from_numpy(rng.choice(V, 300, replace=False)).cuda()
    table = torch.empty((300, dim), dtype=torch.float32, device="cuda")
    tabs = (ctypes.c_void_p * 1)(table.data_ptr())
    _lib.check(L.pg_cache_fill(h, _lib.ptr(cached), 300, 0, tabs, 1, None), "pg_cache_fill")
    ids = torch.from_numpy(rng.integers(0, V, 4097)).cuda()
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    mask = torch.zeros(4097, dtype=torch.uint8, device="cuda")
    for mode in (1, 2, 0):
        out = torch.zeros((4097, dim), dtype=torch.float32, device="cuda")
        outs = (ctypes.c_void_p * 1)(out.data_ptr())
        _lib.check(L.pg_cache_fetch(h, _lib.ptr(ids), 4097, outs, _lib.ptr(mask), _lib.ptr(counts), mode, None),
                   "pg_cache_fetch")
        torch.cuda.synchronize()
        want = host[nid_map.cpu().numpy()[ids.cpu().numpy()], :dim]
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    fl = flag.cpu().numpy().astype(bool)
    assert fl.sum() == 300
    np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), fl[ids.cpu().numpy()])
    assert counts.cpu().tolist() == [3 * 4097, 3 * int((~fl[ids.cpu().numpy()]).sum())]
    # live timing records (pg_timing_*): split, miss fetch and hit gather of one call, in launch order
    _lib.timing_drain()
    _lib.timing_enable(True)
    _lib.check(L.pg_cache_fetch(h, _lib.ptr(ids), 4097, outs, None, None, 0, None), "pg_cache_fetch")
    _lib.timing_enable(False)
    recs = _lib.timing_drain()
    assert [r[0] for r in recs] == [_lib.T_SPLIT, _lib.T_GATHER_MISS, _lib.T_GATHER_HIT]
    assert all(0 < r[1] < 100 for r in recs)
    assert _lib.timing_drain() == []
    # error behaviour: bad mode, pageable host table
    assert L.pg_cache_fetch(h, _lib.ptr(ids), 4097, outs, None, None, 9, None) == _lib.PG_ERR_INVALID
    assert b"mode" in L.pg_last_error()
    L.pg_cache_destroy(h)
    pageable = np.zeros((10, 4), np.float32)
    f[0].dim, f[0].host_stride, f[0].host_table = 4, 4, pageable.ctypes.data
    assert L.pg_cache_create(10, _lib.ptr(flag), _lib.ptr(l2c), _lib.ptr(nid_map), 1, f, 0,
                             ctypes.byref(h)) == _lib.PG_ERR_INVALID
    _lib.check(L.pg_host_free(p), "pg_host_free")


def test_h2d_probe():
    from pagraph_b200 import _lib
    bw = ctypes.c_double()
    _lib.check(_lib.lib().pg_measure_h2d(0, 256 << 20, 3, ctypes.byref(bw)), "pg_measure_h2d")
    assert 5.0 < bw.value < 200.0
