"""GCNTrainEngine (two-stream CUDA-graph pipeline, device-resident sizes) vs the eager public-API loop of
examples/profile/pa_gcn.py: same minibatches, same model math => same losses and parameters (tolerance only for
fp32 summation order / atomics in the backward scatter)."""
import numpy as np
import pytest

from conftest import random_in_csr

pytestmark = pytest.mark.gpu


def _world(V=4000, nnz=60000, F=600, classes=7, seed=5):
    import torch
    from pagraph_b200 import DGLGraph
    from pagraph_b200.graph_store import LocalGraphStore
    rng = np.random.default_rng(seed)
    indptr, indices, eids, _ = random_in_csr(V, nnz, seed)
    store = LocalGraphStore(name="engine")
    store.ndata["features"] = torch.from_numpy(rng.random((V, F), dtype=np.float32))
    store.ndata["norm"] = torch.from_numpy((1.0 / np.maximum(np.diff(indptr), 1)).astype(np.float32)[:, None])
    g = DGLGraph.from_in_csr(indptr, indices, eids)
    labels = torch.from_numpy(rng.integers(0, classes, V))
    train = rng.choice(V, 1100, replace=False).astype(np.int64)
    return g, store, labels, train, V, F, classes


def _model(F, classes, dropout, n_hidden=16):
    import torch
    from pagraph_b200.model.gcn_nssc import GCNSampling
    torch.manual_seed(0)
    return GCNSampling(F, n_hidden, classes, 1, torch.relu, dropout).cuda()


def _eager_losses(g, store, labels, train, V, F, classes, cap, batch, fanouts, steps, n_hidden=16):
    import torch
    from pagraph_b200.sampling import NeighborSampler
    from pagraph_b200.storage import GraphCacheServer
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(["features", "norm"])
    model = _model(F, classes, 0.0, n_hidden)
    opt = torch.optim.Adam(model.parameters(), lr=3e-2)
    sampler = NeighborSampler(g, batch, fanouts, num_hops=len(fanouts), seed_nodes=torch.from_numpy(train), seed=11)
    lab = labels.cuda()
    losses = []
    for k, nf in enumerate(sampler.batches(0, steps)):
        cs.fetch_data(nf)
        pred = model(nf)
        loss = torch.nn.functional.cross_entropy(pred, lab[nf.layer_parent_nid_dev(-1)])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
        if k == 0 and cap is not None:
            cs.auto_cache(g, ["features", "norm"], capability=cap)
    return losses, [p.detach().cpu().numpy() for p in model.parameters()]


@pytest.mark.parametrize("host_inputs", [False, True])
@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("cap", [None, 900, 10 ** 9])
def test_engine_matches_eager_loop(cap, use_graphs, host_inputs):
    import torch
    from pagraph_b200.engine import GCNTrainEngine
    from pagraph_b200.storage import GraphCacheServer
    g, store, labels, train, V, F, classes = _world()
    batch, fanouts, steps = 256, [6, 4], 7                     # 1100 seeds / 256: minibatch 4 is partial (76 seeds), then wraps
    want, want_params = _eager_losses(g, store, labels, train, V, F, classes, cap, batch, fanouts, steps)
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(["features", "norm"])
    cs.log = True
    model = _model(F, classes, 0.0)
    opt = torch.optim.Adam(model.parameters(), lr=3e-2, capturable=use_graphs)
    eng = GCNTrainEngine(g, cs, model, opt, train, labels, batch, fanouts, seed=11, shuffle=False,
                         host_inputs=host_inputs, use_graphs=use_graphs, stage_rows=300)   # tiny staging: overflow path too
    got = [eng.steps(1, read_loss=True)]
    if cap is not None:
        cs.auto_cache(g, ["features", "norm"], capability=cap)
    got += [eng.steps(1, read_loss=True) for _ in range(2)]
    last = eng.steps(steps - 3)                                # pipelined, no read-back
    got_params = [p.detach().cpu().numpy() for p in model.parameters()]
    np.testing.assert_allclose(got, want[:3], rtol=2e-4)
    np.testing.assert_allclose(float(last), want[-1], rtol=2e-4)
    for a, b in zip(got_params, want_params):
        np.testing.assert_allclose(a, b, rtol=1e-2, atol=1e-3)   # 7 Adam steps at lr 3e-2 amplify fp32 summation-order noise
    assert eng.launches > 0
    if not cs.full_cached:
        assert cs.try_num > 0 and 0 < cs.miss_num <= cs.try_num
    eng.close()


@pytest.mark.parametrize("flat_bucket", [False, True])
@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("cap", [900, 10 ** 9])
def test_engine_fused_dense_stage_matches_eager_loop(cap, use_graphs, flat_bucket):
    """n_hidden = 32 (the reference default): the engine's dense stage is the six-kernel fused path (tensor-core NodeUpdate
    forward / dW, head + loss, gradients written straight into .grad / the flat bucket) — same losses and parameters as
    the eager autograd loop."""
    import torch
    from pagraph_b200.engine import GCNTrainEngine
    from pagraph_b200.parallel import FlatGradAllReduce
    from pagraph_b200.storage import GraphCacheServer
    g, store, labels, train, V, F, classes = _world()
    batch, fanouts, steps = 256, [6, 4], 7
    want, want_params = _eager_losses(g, store, labels, train, V, F, classes, cap, batch, fanouts, steps, n_hidden=32)
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(["features", "norm"])
    model = _model(F, classes, 0.0, 32)
    sync = FlatGradAllReduce(model) if flat_bucket else None
    opt = torch.optim.Adam(sync.flat_parameters() if flat_bucket else model.parameters(), lr=3e-2, capturable=use_graphs)
    eng = GCNTrainEngine(g, cs, model, opt, train, labels, batch, fanouts, sync=sync, seed=11, shuffle=False,
                         use_graphs=use_graphs, stage_rows=300)
    got = [eng.steps(1, read_loss=True)]
    assert eng._dense_ok and (eng.fused_opt is not None) == flat_bucket
    cs.auto_cache(g, ["features", "norm"], capability=cap)
    got += [eng.steps(1, read_loss=True) for _ in range(2)]
    last = eng.steps(steps - 3)
    got_params = [p.detach().cpu().numpy() for p in model.parameters()]
    np.testing.assert_allclose(got, want[:3], rtol=2e-4)
    np.testing.assert_allclose(float(last), want[-1], rtol=2e-4)
    for a, b in zip(got_params, want_params):
        np.testing.assert_allclose(a, b, rtol=1e-2, atol=1e-3)
    eng.close()


def test_engine_fused_dense_stage_dropout():
    """Dropout inside the fused dense stage (hash mask keyed by the device step counter, regenerated in the backward):
    finite, re-keyed every replay, and the loss goes down on a learnable target."""
    import torch
    from pagraph_b200.engine import GCNTrainEngine
    from pagraph_b200.storage import GraphCacheServer
    g, store, labels, train, V, F, classes = _world(classes=3)
    labels = (store.ndata["features"][:, 0] > 0.5).long()            # two of the three classes occur
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(["features", "norm"])
    cs.auto_cache(g, ["features", "norm"], capability=V)
    model = _model(F, 3, 0.3, 32)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True)
    eng = GCNTrainEngine(g, cs, model, opt, train[:1024], labels, 256, [6, 4], seed=3)
    losses = [eng.steps(1, read_loss=True) for _ in range(24)]
    assert eng._dense_ok
    assert all(np.isfinite(losses))
    assert len(set(losses)) == len(losses)
    assert int(eng.step_counter.item()) == 24
    assert np.mean(losses[-4:]) < np.mean(losses[:4])                # 3 classes, one never used: the loss must drop from ln 3
    eng.close()


def test_engine_dropout_trains_and_graph_replays_rekey():
    """With dropout the fused mask changes every step (device step counter) and the loss goes down."""
    import torch
    from pagraph_b200.engine import GCNTrainEngine
    from pagraph_b200.storage import GraphCacheServer
    g, store, labels, train, V, F, classes = _world(classes=3)
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(["features", "norm"])
    cs.auto_cache(g, ["features", "norm"], capability=V)
    model = _model(F, 3, 0.3)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True)
    eng = GCNTrainEngine(g, cs, model, opt, train[:1024], labels, 256, [6, 4], seed=3)
    losses = [eng.steps(1, read_loss=True) for _ in range(24)]
    assert all(np.isfinite(losses))
    assert len(set(losses)) == len(losses)
    assert int(eng.step_counter.item()) == 24
    eng.close()


@pytest.mark.parametrize("advance", [False, True])
def test_peer_adam_single_rank_matches_torch_adam(advance):
    """pg_allreduce_adam with world == 1 is torch.optim.Adam (capturable math) on the flat bucket; pg_allreduce_adam_next
    (advance) is the same step with both step counters incremented by the kernel itself."""
    import torch
    from pagraph_b200.parallel import FlatGradAllReduce, PeerAdam
    torch.manual_seed(0)
    make = lambda: torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.ReLU(), torch.nn.Linear(19, 5)).cuda()
    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    sync = FlatGradAllReduce(a)
    opt_a = torch.optim.Adam(sync.flat_parameters(), lr=3e-2, weight_decay=1e-3, capturable=True)
    assert PeerAdam.supported(sync, opt_a)
    fused = PeerAdam(sync, opt_a)
    opt_b = torch.optim.Adam(b.parameters(), lr=3e-2, weight_decay=1e-3)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for it in range(6):
        x = torch.randn(64, 37, device="cuda")
        for m in (a, b):
            m.zero_grad(set_to_none=False) if m is b else sync.zero_grad()
            m(x).square().mean().backward()
        if advance:
            fused.step(step, advance=True)
        else:
            step.add_(1)
            fused.step(step)
        opt_b.step()
        assert int(step.item()) == it + 1
    for pa, pb in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(pa, pb, rtol=1e-4, atol=1e-6)
    assert float(opt_a.state[sync.flat_parameters()[0]]["step"]) == 6.0
    fused.close()


def _two_rank_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    from pagraph_b200.parallel import FlatGradAllReduce, PeerAdam
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.ReLU(), torch.nn.Linear(19, 5)).cuda()
    sync = FlatGradAllReduce(model)
    opt = torch.optim.Adam(sync.flat_parameters(), lr=1e-2, capturable=True)
    fused = PeerAdam(sync, opt)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(world * 32, 37, device="cuda", generator=g)
    graph, static_x = None, x[rank * 32:(rank + 1) * 32].clone()
    for it in range(40):                       # eager for 3 steps, then the same step as a replayed CUDA graph
        def body():
            sync.zero_grad()
            model(static_x).square().mean().backward()
            fused.step(step, advance=True)         # the counter-advancing launch, as the engine issues it
        if it < 3:
            body()
        else:
            if graph is None:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    body()
            graph.replay()
    torch.cuda.synchronize()
    out[rank] = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    dist.barrier()
    fused.close()
    dist.destroy_process_group()


def test_peer_adam_two_ranks_equals_global_batch():
    """Two processes / two GPUs: the NVLink one-shot all-reduce + Adam keeps replicas identical and equals single-process
    training on the concatenated batch (mean of the per-rank mean-gradients). Skipped with fewer than 2 GPUs."""
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
        mp.spawn(_two_rank_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    np.testing.assert_array_equal(res[0], res[1])
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.ReLU(), torch.nn.Linear(19, 5)).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(64, 37, device="cuda", generator=g)
    for _ in range(40):
        opt.zero_grad()
        (0.5 * (model(x[:32]).square().mean() + model(x[32:]).square().mean())).backward()
        opt.step()
    want = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    np.testing.assert_allclose(res[0], want, rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("host_inputs", [False, True])
@pytest.mark.parametrize("n_hidden", [16, 32])
def test_engine_with_duplicate_train_ids_matches_eager_loop(n_hidden, host_inputs):
    """The reference's partition files contain duplicate train ids (isolated train vertices all become sub-graph id 0,
    PaGraph/partition/utils.py:47-51). The sampler drops duplicate seeds, so the model's rows follow the seed LAYER: the
    engine must pair them with the labels of that layer (not of the raw seed positions) and average over its row count —
    exactly what the eager loop does with labels[nf.layer_parent_nid(-1)]."""
    import torch
    from pagraph_b200.engine import GCNTrainEngine
    from pagraph_b200.storage import GraphCacheServer
    g, store, labels, train, V, F, classes = _world()
    rng = np.random.default_rng(3)
    train = np.concatenate([train[:600], np.full(90, train[0]), train[5:25], train[600:900]])   # many repeats, every batch
    rng.shuffle(train)
    batch, fanouts, steps = 256, [6, 4], 6
    want, want_params = _eager_losses(g, store, labels, train, V, F, classes, 900, batch, fanouts, steps, n_hidden=n_hidden)
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(["features", "norm"])
    model = _model(F, classes, 0.0, n_hidden)
    opt = torch.optim.Adam(model.parameters(), lr=3e-2, capturable=True)
    eng = GCNTrainEngine(g, cs, model, opt, train, labels, batch, fanouts, seed=11, shuffle=False, host_inputs=host_inputs)
    assert eng.has_dups
    got = [eng.steps(1, read_loss=True)]
    cs.auto_cache(g, ["features", "norm"], capability=900)
    got += [eng.steps(1, read_loss=True) for _ in range(2)]
    last = eng.steps(steps - 3)
    np.testing.assert_allclose(got, want[:3], rtol=2e-4)
    np.testing.assert_allclose(float(last), want[-1], rtol=2e-4)
    for a, b in zip([p.detach().cpu().numpy() for p in model.parameters()], want_params):
        np.testing.assert_allclose(a, b, rtol=1e-2, atol=1e-3)
    eng.close()


def _eager_loop(model, g, store, labels, train, V, fields, cap, batch, fanouts, steps):
    """the reference's op-by-op loop (examples/profile/pa_gcn.py / pa_gs.py:86-97) on the drop-in classes"""
    import torch
    from pagraph_b200.sampling import NeighborSampler
    from pagraph_b200.storage import GraphCacheServer
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(fields)
    opt = torch.optim.Adam(model.parameters(), lr=3e-2)
    sampler = NeighborSampler(g, batch, fanouts, num_hops=len(fanouts), seed_nodes=torch.from_numpy(train), seed=11)
    lab = labels.cuda()
    losses = []
    for k, nf in enumerate(sampler.batches(0, steps)):
        cs.fetch_data(nf)
        loss = torch.nn.functional.cross_entropy(model(nf), lab[nf.layer_parent_nid_dev(-1)])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
        if k == 0 and cap is not None:
            cs.auto_cache(g, fields, capability=cap)
    return losses, [p.detach().cpu().numpy() for p in model.parameters()]


def _run_engine(cls, model, g, store, labels, train, V, fields, cap, batch, fanouts, steps, use_graphs, **kw):
    import torch
    from pagraph_b200.storage import GraphCacheServer
    cs = GraphCacheServer(store, V, torch.arange(V), 0)
    cs.init_field(fields)
    opt = torch.optim.Adam(model.parameters(), lr=3e-2, capturable=use_graphs)
    eng = cls(g, cs, model, opt, train, labels, batch, fanouts, seed=11, shuffle=False, use_graphs=use_graphs, stage_rows=300, **kw)
    got = [eng.steps(1, read_loss=True)]
    if cap is not None:
        cs.auto_cache(g, fields, capability=cap)
    got += [eng.steps(1, read_loss=True) for _ in range(2)]
    last = eng.steps(steps - 3)
    params = [p.detach().cpu().numpy() for p in model.parameters()]
    return got, float(last), params, eng


@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("cap", [None, 900, 10 ** 9])
def test_sage_engine_matches_eager_loop(cap, use_graphs):
    """GraphSAGE-mean (PaGraph/model/graphsage_nssc.py:74-134, trainer examples/profile/pa_gs.py): three aggregations per
    step, two of them 600 wide and fused with the cache lookup in the engine — same losses and parameters as the eager loop."""
    import torch
    from pagraph_b200.engine import SageTrainEngine, make_train_engine
    from pagraph_b200.model.graphsage_nssc import GraphSageSampling
    g, store, labels, train, V, F, classes = _world()
    batch, fanouts, steps = 256, [6, 4], 7

    def model():
        torch.manual_seed(0)
        return GraphSageSampling(F, 16, classes, 1, torch.relu, 0.0, 'mean').cuda()
    want, want_params = _eager_loop(model(), g, store, labels, train, V, ["features"], cap, batch, fanouts, steps)
    m = model()
    got, last, params, eng = _run_engine(SageTrainEngine, m, g, store, labels, train, V, ["features"], cap, batch, fanouts, steps,
                                         use_graphs)
    assert isinstance(make_train_engine(g, eng.cacher, m, eng.opt, train, labels, batch, fanouts, use_graphs=False), SageTrainEngine)
    np.testing.assert_allclose(got, want[:3], rtol=2e-4)
    np.testing.assert_allclose(last, want[-1], rtol=2e-4)
    for a, b in zip(params, want_params):
        np.testing.assert_allclose(a, b, rtol=1e-2, atol=1e-3)
    eng.close()


@pytest.mark.parametrize("n_hidden", [16, 32])          # 32 = the fused dense stage, 16 = the autograd body
@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("cap", [None, 900, 10 ** 9])
def test_gcn_preprocess_engine_matches_eager_loop(cap, use_graphs, n_hidden):
    """GCN --preprocess (PaGraph/model/gcn_nssc.py:80-100; num_hops = n_layers = 1): linear on the dropped-out input rows,
    one 64-wide block aggregation, head — engine (cache gather fused with the dropout mask, fused dense kernels) vs eager."""
    import torch
    from pagraph_b200.engine import GCNPreprocessTrainEngine
    from pagraph_b200.model.gcn_nssc import GCNSampling
    g, store, labels, train, V, F, classes = _world()
    batch, fanouts, steps = 256, [6], 7

    def model():
        torch.manual_seed(0)
        return GCNSampling(F, n_hidden, classes, 1, torch.relu, 0.0, True).cuda()
    fields = ["features", "norm"]
    want, want_params = _eager_loop(model(), g, store, labels, train, V, fields, cap, batch, fanouts, steps)
    got, last, params, eng = _run_engine(GCNPreprocessTrainEngine, model(), g, store, labels, train, V, fields, cap, batch,
                                         fanouts, steps, use_graphs)
    assert eng._dense_ok == (n_hidden == 32)
    np.testing.assert_allclose(got, want[:3], rtol=2e-4)
    # seven Adam steps at lr 3e-2 on un-normalised 600-wide inputs: the loss is O(10) and still moving fast, so the last
    # value carries the amplified difference between the 3xTF32 kernels and cuBLAS fp32
    np.testing.assert_allclose(last, want[-1], rtol=5e-3)
    for a, b in zip(params, want_params):
        # Adam moves every weight by ~lr per step whatever the gradient's size: a weight whose gradient is noise-level
        # can differ by a few lr-sized steps between two fp32 summation orders
        assert np.mean(np.abs(a - b) > 1e-3 + 1e-2 * np.abs(b)) < 1e-3
        np.testing.assert_allclose(a, b, rtol=1e-2, atol=3e-2)
    eng.close()
