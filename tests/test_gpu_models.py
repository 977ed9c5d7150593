"""Model mirrors (pagraph_b200/model), the server-side --preprocess fold, the partition closure and the drop-in entry
scripts, on the GPU: model outputs against a float64 numpy restatement of PaGraph/model/gcn_nssc.py /
graphsage_nssc.py evaluated on the ORACLE's NodeFlow; closure against a numpy BFS; entries end to end on a tiny dataset."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as spsp

import oracle
from conftest import ROOT, random_in_csr

pytestmark = pytest.mark.gpu


def _setup(F=40, fanouts=(5, 4), V=1500, nnz=20000, seeds=120, seed=2, fields=("features", "norm")):
    import torch
    from pagraph_b200 import DGLGraph
    from pagraph_b200.graph_store import LocalGraphStore
    from pagraph_b200.sampling import NeighborSampler
    from pagraph_b200.storage import GraphCacheServer
    rng = np.random.default_rng(seed)
    indptr, indices, eids, _ = random_in_csr(V, nnz, seed)
    feats = rng.standard_normal((V, F)).astype(np.float32)
    norm = (1.0 / np.maximum(np.diff(indptr), 1)).astype(np.float32)[:, None]
    host = {"features": feats, "norm": norm, "neigh": rng.standard_normal((V, F)).astype(np.float32)}
    store = LocalGraphStore(name="models")
    for f in fields:
        store.ndata[f] = torch.from_numpy(host[f])
    g = DGLGraph.from_in_csr(indptr, indices, eids)
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(list(fields))
    cs.auto_cache(g, list(fields), capability=V // 3)
    sd = rng.choice(V, seeds, replace=False).astype(np.int64)
    nf = NeighborSampler(g, seeds, list(fanouts), num_hops=len(fanouts), seed_nodes=torch.from_numpy(sd), seed=4).sample_batch(0)
    ref = oracle.sample(indptr, indices, eids, sd, list(fanouts), seed=4)
    cs.fetch_data(nf)
    return nf, ref, host, store


def _agg(ref, i, h, mode):
    ip, cols, base = ref.block(i)
    out = np.zeros((len(ip) - 1, h.shape[1]))
    for r in range(len(ip) - 1):
        rows = h[cols[ip[r]:ip[r + 1]] - base]
        if len(rows):
            out[r] = rows.sum(0) / (len(rows) if mode == "mean" else 1)
    return out


def _lin(m, x):
    return x @ m.weight.detach().cpu().double().numpy().T + m.bias.detach().cpu().double().numpy()


def _relu(x):
    return np.maximum(x, 0)


@pytest.mark.parametrize("n_layers", [1, 2])
def test_gcn_sampling_and_infer_match_numpy(n_layers):
    """gcn_nssc.py:60-77 (train, mean) and :103-164 (infer: sum then x norm) on a 2- / 3-hop NodeFlow."""
    import torch
    from pagraph_b200.model.gcn_nssc import GCNInfer, GCNSampling
    fan = (5, 4) if n_layers == 1 else (4, 3, 3)
    nf, ref, host, _ = _setup(fanouts=fan)
    torch.manual_seed(1)
    model = GCNSampling(40, 32, 6, n_layers, torch.relu, 0.0).cuda()
    got = model(nf).detach().cpu().numpy()
    h = host["features"][ref.layer_parent_nid(0)].astype(np.float64)
    for i, layer in enumerate(model.layers):
        h = _lin(layer.linear, _agg(ref, i, h, "mean"))
        if layer.concat:
            h = np.concatenate((h, _relu(h)), 1)
        elif layer.activation is not None:
            h = _relu(h)
    np.testing.assert_allclose(got, h, rtol=1e-4, atol=1e-5)
    infer = GCNInfer(40, 32, 6, n_layers, torch.relu).cuda()
    infer.load_state_dict(model.state_dict())
    nf2, ref2, host2, _ = _setup(fanouts=fan)
    got = infer(nf2).detach().cpu().numpy()
    h = host["features"][ref.layer_parent_nid(0)].astype(np.float64)
    for i, layer in enumerate(infer.layers):
        h = _agg(ref, i, h, "sum") * host["norm"][ref.layer_parent_nid(i + 1)].astype(np.float64)
        h = _lin(layer.linear, h)
        if layer.concat:
            h = np.concatenate((h, _relu(h)), 1)
        elif layer.activation is not None:
            h = _relu(h)
    np.testing.assert_allclose(got, h, rtol=1e-4, atol=1e-5)


def test_gcn_preprocess_matches_numpy():
    """gcn_nssc.py:80-100: input linear on layer 0 (+skip concat when n_layers == 1), then the block layers."""
    import torch
    from pagraph_b200.model.gcn_nssc import GCNSampling
    nf, ref, host, _ = _setup(fanouts=(5,))
    torch.manual_seed(2)
    model = GCNSampling(40, 32, 6, 1, torch.relu, 0.0, preprocess=True).cuda()
    got = model(nf).detach().cpu().numpy()
    h = _lin(model.linear, host["features"][ref.layer_parent_nid(0)].astype(np.float64))
    h = np.concatenate((h, _relu(h)), 1)
    h = _lin(model.layers[0].linear, _agg(ref, 0, h, "mean"))
    np.testing.assert_allclose(got, h, rtol=1e-4, atol=1e-5)


def test_graphsage_matches_numpy():
    """graphsage_nssc.py:74-134, mean aggregator, n_layers = 1 over a 2-hop NodeFlow (3 aggregations)."""
    import torch
    from pagraph_b200.model.graphsage_nssc import GraphSageSampling
    nf, ref, host, _ = _setup(fanouts=(5, 4), fields=("features",))
    torch.manual_seed(3)
    model = GraphSageSampling(40, 16, 6, 1, torch.relu, 0.0, "mean").cuda()
    got = model(nf).detach().cpu().numpy()
    L = ref.num_layers
    h = [host["features"][ref.layer_parent_nid(i)].astype(np.float64) for i in range(L)]
    for lid, layer in enumerate(model.layers):
        act = {}
        for i in range(lid, L - 1):
            z = _lin(layer.fc_self, h[i + 1]) + _lin(layer.fc_neigh, _agg(ref, i, h[i], "mean"))
            act[i + 1] = np.concatenate((z, _relu(z)), 1) if layer.concat else (_relu(z) if layer.activation else z)
        for i in range(lid + 1, L):
            h[i] = act[i]
    np.testing.assert_allclose(got, h[L - 1], rtol=1e-4, atol=1e-5)


def test_server_preprocess_fold_matches_scipy():
    """server/pa_server.py:45-52: features' = norm * (A^T features), here through pg_aggregate_fwd in row blocks."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "server"))
    import pa_server
    from pagraph_b200 import DGLGraph
    rng = np.random.default_rng(0)
    V, nnz, F = 3000, 30000, 24
    src, dst = rng.integers(0, V, nnz), rng.integers(0, V, nnz)
    coo = spsp.coo_matrix((np.ones(nnz), (src, dst)), shape=(V, V))
    feats = rng.standard_normal((V, F)).astype(np.float32)
    g = DGLGraph(coo, readonly=True)
    with np.errstate(divide="ignore"):
        norm = (1.0 / np.asarray(coo.sum(0)).ravel()).astype(np.float32)[:, None]     # 1 / in_degree, inf at 0
    got = pa_server.preprocess_features(g, torch.from_numpy(feats), torch.from_numpy(norm), rows_per_block=700).numpy()
    with np.errstate(invalid="ignore"):
        want = (coo.T.tocsr() @ feats.astype(np.float64)) * norm.astype(np.float64)       # 0 * inf = nan, as the reference
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("hops", [1, 2])
def test_get_sub_graph_is_the_in_neighbour_closure(hops):
    """PaGraph/partition/utils.py:9-52 against a numpy BFS over in-edges."""
    from pagraph_b200 import DGLGraph
    from pagraph_b200.partition import get_sub_graph
    rng = np.random.default_rng(hops)
    V, nnz = 900, 4000
    src, dst = rng.integers(0, V, nnz), rng.integers(0, V, nnz)
    coo = spsp.coo_matrix((np.ones(nnz, np.int64), (src, dst)), shape=(V, V))
    g = DGLGraph(coo, readonly=True)
    train = np.sort(rng.choice(V, 60, replace=False)).astype(np.int64)
    subadj, sub2full, subtrain = get_sub_graph(g, train, hops)
    csc = coo.tocsc()
    frontier, edges = set(train.tolist()), set()
    for _ in range(hops):
        nxt = set()
        for v in frontier:
            for u in csc.indices[csc.indptr[v]:csc.indptr[v + 1]]:
                edges.add((int(u), int(v)))
                nxt.add(int(u))
        frontier = nxt
    verts = np.array(sorted({x for e in edges for x in e}))
    np.testing.assert_array_equal(sub2full, verts)
    got_edges = set(zip(sub2full[subadj.tocoo().row].tolist(), sub2full[subadj.tocoo().col].tolist()))
    assert got_edges == edges
    assert subadj.dtype == np.uint8 and (subadj.data == 1).all()
    present = train[np.isin(train, verts)]                  # train vertices that have or feed an edge of the closure
    assert set(present.tolist()) <= set(sub2full[subtrain].tolist())
    # the in-process variant (no host round trip): the in-CSR DGLGraph(subadj) would build, sub2full, train ids
    from pagraph_b200.partition.utils import get_sub_graph_device
    ip, ix, s2f, st = get_sub_graph_device(g, train, hops)
    ref_g = DGLGraph(subadj, readonly=True)
    np.testing.assert_array_equal(ip.cpu().numpy(), ref_g.indptr)
    np.testing.assert_array_equal(ix.cpu().numpy(), ref_g.indices)
    np.testing.assert_array_equal(s2f.cpu().numpy(), sub2full)
    np.testing.assert_array_equal(st.cpu().numpy(), subtrain)


def test_entry_scripts_end_to_end(tmp_path):
    """hash partition -> pa_server.py -> pa_gcn.py (1 GPU, 2 epochs) on a tiny dataset with the reference's file layout."""
    from pagraph_b200 import data
    ds = str(tmp_path / "tiny")
    V = 3000
    adj = data.rmat_adj(V, 30000, seed=1)
    data.write_dataset(ds, adj, data.random_feature(V, 600), data.random_label(V, 60), data.split_dataset(V))
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-m", "pagraph_b200.partition.hash", "--dataset", ds, "--partition", "1", "--num-hops", "2",
                    "--seed", "0"], check=True, env=env, cwd=ROOT, timeout=300)
    for f in ("subadj_0.npz", "sub_trainid_0.npy", "sub_train2fullid_0.npy", "sub_label_0.npy"):
        assert os.path.exists(os.path.join(ds, "1naive", f))
    server = subprocess.Popen([sys.executable, os.path.join(ROOT, "server", "pa_server.py"), "--dataset", ds, "--num-workers", "1"],
                              env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        for engine, port in (("eager", "29533"), ("graph", "29534")):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "profile", "pa_gcn.py"), "--dataset", ds,
                                  "--gpu", "0", "--n-epochs", "2", "--batch-size", "500", "--num-neighbors", "5,3",
                                  "--engine", engine, "--keep-store" if engine == "eager" else "--seed=0"],
                                 env=dict(env, MASTER_PORT=port), cwd=ROOT, capture_output=True, text=True, timeout=600)
            assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
            assert "Total Time" in out.stdout and "Epoch average time" in out.stdout
        server.wait(timeout=60)            # leaves once its single worker has signalled completion
        assert server.returncode == 0
    finally:
        if server.poll() is None:
            server.kill()


def test_graphsage_entry_script_end_to_end(tmp_path):
    """hash partition -> pa_server.py --model graphsage -> pa_gs.py (1 GPU, 2 epochs), the reference's GraphSAGE entry
    (examples/profile/pa_gs.py) on the drop-in classes."""
    from pagraph_b200 import data
    ds = str(tmp_path / "tiny_gs")
    V = 3000
    adj = data.rmat_adj(V, 30000, seed=2)
    data.write_dataset(ds, adj, data.random_feature(V, 600), data.random_label(V, 60), data.split_dataset(V))
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-m", "pagraph_b200.partition.hash", "--dataset", ds, "--partition", "1", "--num-hops", "2",
                    "--seed", "0"], check=True, env=env, cwd=ROOT, timeout=300)
    server = subprocess.Popen([sys.executable, os.path.join(ROOT, "server", "pa_server.py"), "--dataset", ds, "--num-workers", "1",
                               "--model", "graphsage"], env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True)
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "profile", "pa_gs.py"), "--dataset", ds, "--gpu", "0",
                              "--n-epochs", "2", "--batch-size", "500", "--num-neighbors", "5,3"],
                             env=dict(env, MASTER_PORT="29535"), cwd=ROOT, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "Total Time" in out.stdout and "Epoch average time" in out.stdout
        server.wait(timeout=60)
        assert server.returncode == 0
    finally:
        if server.poll() is None:
            server.kill()


def test_training_learns_and_eval_entry_reports_accuracy(tmp_path):
    """f5: the pipeline LEARNS. A homophilous 8-class synthetic graph (features carry a noisy class signal, 80 % of the
    edges stay inside a class) is trained with pa_gcn.py (engine path, per-epoch checkpoints), then examples/eval.py
    (reference examples/eval.py:13-46: full-neighbourhood NodeFlow, GCNInfer sum x norm) reports test accuracy far above
    the 12.5 % chance level; the eager loop reaches the same."""
    import scipy.sparse as spsp
    from pagraph_b200 import data
    rng = np.random.default_rng(0)
    V, E, C, Fdim = 4000, 40000, 8, 600
    cls = rng.integers(0, C, V)
    src = rng.integers(0, V, E)
    same = rng.random(E) < 0.8
    by_class = [np.nonzero(cls == c)[0] for c in range(C)]
    dst = np.where(same, [by_class[cls[s]][rng.integers(len(by_class[cls[s]]))] for s in src], rng.integers(0, V, E))
    keep = src != dst
    key = np.unique(src[keep] * V + dst[keep])
    src, dst = key // V, key % V
    adj = spsp.coo_matrix((np.ones(2 * len(src), np.float32), (np.concatenate([src, dst]), np.concatenate([dst, src]))), shape=(V, V))
    adj.sum_duplicates()
    adj.data[:] = 1
    feat = (0.5 * rng.random((V, Fdim))).astype(np.float32)
    blk = Fdim // C
    for c in range(C):
        feat[cls == c, c * blk:(c + 1) * blk] += 0.25
    ds = str(tmp_path / "learn")
    data.write_dataset(ds, adj.tocoo(), feat, cls.astype(np.int64), data.split_dataset(V))
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-m", "pagraph_b200.partition.hash", "--dataset", ds, "--partition", "1", "--num-hops", "2",
                    "--seed", "0"], check=True, env=env, cwd=ROOT, timeout=300)
    server = subprocess.Popen([sys.executable, os.path.join(ROOT, "server", "pa_server.py"), "--dataset", ds, "--num-workers", "1"],
                              env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        accs = {}
        for engine, port in (("graph", "29541"), ("eager", "29542")):
            ck = str(tmp_path / ("ckpt_" + engine))
            out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "profile", "pa_gcn.py"), "--dataset", ds,
                                  "--gpu", "0", "--n-epochs", "6", "--batch-size", "256", "--num-neighbors", "10,10",
                                  "--n-classes", str(C), "--lr", "0.01", "--engine", engine, "--ckpt", ck, "--keep-store"],
                                 env=dict(env, MASTER_PORT=port), cwd=ROOT, capture_output=True, text=True, timeout=900)
            assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
            assert os.path.exists(os.path.join(ck, "gcn-nssc_5"))
            ev = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "eval.py"), "--dataset", ds, "--gpu", "0",
                                 "--feat-size", str(Fdim), "--start", "0", "--end", "6", "--interval", "5", "--ckpt", ck,
                                 "--keep-store"] if engine == "graph" else
                                [sys.executable, os.path.join(ROOT, "examples", "eval.py"), "--dataset", ds, "--gpu", "0",
                                 "--feat-size", str(Fdim), "--start", "0", "--end", "6", "--interval", "5", "--ckpt", ck],
                                env=dict(env, PG_DEBUG_HANG="150"), cwd=ROOT, capture_output=True, text=True, timeout=300)
            assert ev.returncode == 0, ev.stdout[-2000:] + ev.stderr[-3000:]
            lines = [ln for ln in ev.stdout.splitlines() if "Test Accuracy" in ln]
            assert len(lines) == 2, ev.stdout
            accs[engine] = [float(ln.split()[-1]) for ln in lines]
        for engine, (first, last) in accs.items():
            assert last > 0.8, (engine, accs)             # chance = 0.125
            assert last >= first - 0.02, (engine, accs)
        server.wait(timeout=60)
        assert server.returncode == 0
    finally:
        if server.poll() is None:
            server.kill()
