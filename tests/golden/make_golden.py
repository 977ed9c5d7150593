"""Generate tests/golden/*.npz by EXECUTING the real reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

What runs:
  * PaGraph/storage/storage.py::GraphCacheServer (init_field, cache_fix_data, auto_cache,
    fetch_data, fetch_from_cache, get_miss_rate) — unmodified source, imported behind a ~40-line
    stub `dgl` (dgl==0.4.1 is not installable here) and with torch's `.cuda()` family mapped to CPU
    (no GPU in this container). The module only moves tensors with those calls; every index /
    mask / copy statement executes as written.
  * PaGraph/partition/dg.py::dg — pure numpy/scipy, unmodified.
Nothing from /root/reference is copied into the repo; only inputs and outputs are stored.

numpy dispatch: dg.py breaks score ties through `np.argsort(score)[-2:]` (dg.py:31), whose tie order is whatever
numpy's default sort does. The numpy the reference targets (2019, introsort -> insertion sort below 16 elements) is
stable for P <= 16; numpy >= 1.25 on an AVX-512 host dispatches to a vectorised sorting network that is not, and then
the reference's own assignments depend on the CPU it runs on. This script therefore re-executes itself with the SIMD
sort disabled (NPY_DISABLE_CPU_FEATURES), so that the fixtures hold the era-faithful, CPU-independent behaviour.
"""
import contextlib
import importlib.util
import os
import sys
import types

_NO_SIMD = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2"
if os.environ.get("NPY_DISABLE_CPU_FEATURES") != _NO_SIMD:
    os.environ["NPY_DISABLE_CPU_FEATURES"] = _NO_SIMD
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np
import scipy.sparse as spsp
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- stub dgl + CPU "cuda"
def install_stubs():
    dgl = types.ModuleType("dgl")

    class DGLGraph:  # only constructed by scripts, not by the functions we call
        def __init__(self, *a, **k):
            pass

    class Frame(dict):
        def __init__(self, data=None):
            super().__init__(data or {})

    class FrameRef:
        def __init__(self, frame):
            self._frame = frame

        def __getitem__(self, k):
            return self._frame[k]

        def keys(self):
            return self._frame.keys()

    frame_mod = types.ModuleType("dgl.frame")
    frame_mod.Frame, frame_mod.FrameRef = Frame, FrameRef
    utils_mod = types.ModuleType("dgl.utils")
    dgl.DGLGraph, dgl.frame, dgl.utils = DGLGraph, frame_mod, utils_mod
    sys.modules.update({"dgl": dgl, "dgl.frame": frame_mod, "dgl.utils": utils_mod})
    if "numba" not in sys.modules:
        try:
            import numba  # noqa: F401  (storage.py:12 imports it, never uses it)
        except Exception:
            sys.modules["numba"] = types.ModuleType("numba")

    # CPU stand-ins for the device-placement calls storage.py makes
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.LongTensor = torch.LongTensor
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.device = lambda *_a, **_k: contextlib.nullcontext()
    torch.cuda.max_memory_allocated = lambda device=None: 0
    torch.cuda.max_memory_cached = lambda device=None: 0


def load_ref(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------- fakes for the objects storage.py touches
class _Col:
    def __init__(self, t):
        self.data = t


class FakeStoreGraph:
    """graph._node_frame._frame[name].data -> CPU tensor [V, dim] (storage.py:128,131)."""

    def __init__(self, fields):
        self._node_frame = types.SimpleNamespace(_frame={k: _Col(v) for k, v in fields.items()})


class _Idx:
    def __init__(self, t):
        self.t = t

    def tousertensor(self):
        return self.t


class FakeNodeFlow:
    def __init__(self, node_mapping, layer_offsets):
        self._node_mapping = _Idx(torch.from_numpy(node_mapping))
        self._layer_offsets = [int(x) for x in layer_offsets]
        self.num_layers = len(layer_offsets) - 1
        self._node_frames = [None] * self.num_layers

    def layer_parent_nid(self, i):
        i = i % self.num_layers
        return self._node_mapping.t[self._layer_offsets[i]:self._layer_offsets[i + 1]]


class FakeLocalGraph:
    def __init__(self, out_deg):
        self._d = torch.from_numpy(out_deg)

    def out_degrees(self):
        return self._d


def storage_case(storage, seed, V_full, V_sub, dims, layer_sizes, cap, out_deg_distinct=True):
    rng = np.random.default_rng(seed)
    fields = {n: rng.random((V_full, d), dtype=np.float32) for n, d in dims.items()}
    if "norm" in fields:
        fields["norm"][rng.integers(0, V_full, 3)] = np.inf  # 1/in_deg==inf rows (pa_server.py:43)
    nid_map = np.sort(rng.choice(V_full, V_sub, replace=False)).astype(np.int64)
    if out_deg_distinct:
        out_deg = rng.permutation(V_sub).astype(np.int64)          # no ties -> hit set well defined
    else:
        out_deg = rng.integers(0, 6, V_sub).astype(np.int64)
        # make the cut fall between two different degrees so the tie order cannot change the set
        order = np.argsort(-out_deg, kind="stable")
        while cap < V_sub and cap > 0 and out_deg[order[cap - 1]] == out_deg[order[cap]]:
            cap += 1
    node_mapping = np.concatenate(
        [np.sort(rng.choice(V_sub, n, replace=False)) if i < len(layer_sizes) - 1
         else rng.choice(V_sub, n, replace=False) for i, n in enumerate(layer_sizes)]).astype(np.int64)
    layer_offsets = np.concatenate([[0], np.cumsum(layer_sizes)]).astype(np.int64)

    g = FakeStoreGraph({k: torch.from_numpy(v) for k, v in fields.items()})
    cs = storage.GraphCacheServer(g, V_sub, torch.from_numpy(nid_map), 0)
    cs.init_field(list(dims))
    cs.log = True
    out = {"nid_map": nid_map, "out_deg": out_deg, "node_mapping": node_mapping,
           "layer_offsets": layer_offsets, "cap": np.int64(cap), "field_names": np.array(list(dims)),
           "total_dim": np.int64(cs.total_dim)}
    for n, v in fields.items():
        out["host_" + n] = v

    # (1) before any caching: every row is a miss (first training step, pa_gcn.py:88 before :100)
    nf = FakeNodeFlow(node_mapping, layer_offsets)
    cs.fetch_data(nf)
    for i in range(nf.num_layers):
        for n in dims:
            out["cold_l%d_%s" % (i, n)] = nf._node_frames[i][n].numpy().copy()
    out["cold_miss_rate"] = np.float64(cs.get_miss_rate())

    # (2) auto_cache with a capacity we control: total_memory is the only knob storage.py:78-84 reads
    total = cap * cs.total_dim * 4 + 1024 ** 3 + 2
    torch.cuda.get_device_properties = lambda _d: types.SimpleNamespace(total_memory=total)
    cs.auto_cache(FakeLocalGraph(out_deg), list(dims))
    out["capability"] = np.int64(cs.capability)
    out["cached_num"] = np.int64(cs.cached_num)
    out["full_cached"] = np.bool_(cs.full_cached)
    out["gpu_flag"] = cs.gpu_flag.numpy().copy()
    out["l2c_on_cached"] = np.where(out["gpu_flag"], cs.localid2cacheid.numpy(), -1)
    for n in dims:
        out["cache_" + n] = cs.gpu_fix_cache[n].numpy().copy()

    nf = FakeNodeFlow(node_mapping, layer_offsets)
    cs.fetch_data(nf)
    for i in range(nf.num_layers):
        for n in dims:
            out["warm_l%d_%s" % (i, n)] = nf._node_frames[i][n].numpy().copy()
    if not cs.full_cached:
        out["warm_miss_num"] = np.int64(cs.miss_num)
        out["warm_try_num"] = np.int64(cs.try_num)
        out["warm_miss_rate"] = np.float64(cs.get_miss_rate())
    return out


def dg_case(dgmod, seed, V, nnz, P, hops):
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, V, nnz), rng.integers(0, V, nnz)
    keep = src != dst
    adj = spsp.coo_matrix((np.ones(keep.sum(), np.int64), (src[keep], dst[keep])), shape=(V, V))
    adj.sum_duplicates()
    train = np.sort(rng.choice(V, int(V * 0.65), replace=False)).astype(np.int64)
    sub_v, sub_trainv = dgmod.dg(P, adj, train, hops)
    out = {"row": adj.row.astype(np.int64), "col": adj.col.astype(np.int64), "V": np.int64(V),
           "train": train, "P": np.int64(P), "hops": np.int64(hops)}
    for p in range(P):
        out["sub_v_%d" % p] = sub_v[p].astype(np.int64)
        out["sub_trainv_%d" % p] = sub_trainv[p].astype(np.int64)
    return out


def main():
    install_stubs()
    storage = load_ref("PaGraph/storage/storage.py", "ref_storage")
    cases = {
        # GCN fields, partial cache, 3 NodeFlow layers (2-hop)
        "storage_gcn_partial": dict(seed=11, V_full=400, V_sub=300, dims={"features": 24, "norm": 1},
                                    layer_sizes=[120, 40, 12], cap=60),
        # tie-heavy degrees, odd feature width (Reddit-like 602 -> here 10: 8B-aligned rows only)
        "storage_ties_odd": dict(seed=12, V_full=257, V_sub=257, dims={"features": 10, "norm": 1},
                                 layer_sizes=[90, 31, 7], cap=50, out_deg_distinct=False),
        # capacity >= node_num -> full_cached -> fetch_from_cache path (storage.py:207-216)
        "storage_full": dict(seed=13, V_full=128, V_sub=96, dims={"features": 16}, layer_sizes=[50, 9],
                             cap=96),
        # GraphSAGE --preprocess fields (pa_gs.py:45), 4 layers, one empty-miss layer likely
        "storage_sage4": dict(seed=14, V_full=500, V_sub=350, dims={"features": 8, "neigh": 8},
                              layer_sizes=[200, 80, 20, 4], cap=300),
    }
    for name, kw in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **storage_case(storage, **kw))
        print("wrote", name)

    sys.path.insert(0, os.path.join(REF, "PaGraph/partition"))  # dg.py uses script-relative imports
    pkg = types.ModuleType("PaGraph")
    pkg.data = types.ModuleType("PaGraph.data")
    sys.modules.update({"PaGraph": pkg, "PaGraph.data": pkg.data,
                        "networkx": sys.modules.get("networkx") or types.ModuleType("networkx")})
    # partition/utils.py imports dgl at module scope and uses dgl.contrib only inside get_sub_graph
    dgmod = load_ref("PaGraph/partition/dg.py", "ref_dg")
    for name, kw in {"dg_p2_h1": dict(seed=21, V=200, nnz=1200, P=2, hops=1),
                     "dg_p4_h2": dict(seed=22, V=300, nnz=900, P=4, hops=2)}.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **dg_case(dgmod, **kw))
        print("wrote", name)


if __name__ == "__main__":
    main()
