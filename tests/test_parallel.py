"""CPU tests of the data-parallel plumbing (pagraph_b200/parallel.py): flat-bucket gradient all-reduce over gloo
(world_size 2), hash split of train ids, per-rank batch-count equalisation."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pagraph_b200.parallel import FlatGradAllReduce, equalised_num_batches, hash_split


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(12, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))   # different init per rank
    sync = FlatGradAllReduce(model)                       # broadcasts rank 0's weights
    opt = torch.optim.Adam(sync.flat_parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(64, 12, generator=g), torch.randint(0, 3, (64,), generator=g)
    lo, hi = rank * 32, (rank + 1) * 32                   # each rank: its own half of the global batch
    for _ in range(3):
        loss = torch.nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi])
        sync.zero_grad()
        loss.backward()
        sync()
        opt.step()
    nb = equalised_num_batches(10 + rank)                 # uneven counts -> min
    out[rank] = (torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy(), nb)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process_global_batch():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    np.testing.assert_array_equal(res[0][0], res[1][0])   # replicas stay identical
    assert res[0][1] == res[1][1] == 10
    # single process over the whole batch with rank 0's initial weights = mean of the per-rank gradients
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(12, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(64, 12, generator=g), torch.randint(0, 3, (64,), generator=g)
    for _ in range(3):
        loss = torch.nn.functional.cross_entropy(model(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    want = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    np.testing.assert_allclose(res[0][0], want, rtol=1e-4, atol=1e-6)


def test_flat_parameters_alias_module_parameters():
    model = _make_model()
    sync = FlatGradAllReduce(model)
    flat = sync.flat_parameters()[0]
    assert flat.numel() == sum(p.numel() for p in model.parameters())
    model(torch.randn(4, 12)).sum().backward()
    assert flat.grad.abs().sum() > 0
    before = [p.detach().clone() for p in model.parameters()]
    torch.optim.SGD([flat], lr=0.1).step()
    assert all(not torch.equal(a, p) for a, p in zip(before, model.parameters()))
    sync.zero_grad()
    assert all((p.grad == 0).all() for p in model.parameters())


def test_hash_split_is_an_equal_chunk_partition():
    ids = np.arange(1003, dtype=np.int64) * 3
    parts = hash_split(ids, 4, seed=1)
    assert [len(p) for p in parts] == [250, 250, 250, 253]          # last chunk takes the remainder (hash.py:44-51)
    assert np.array_equal(np.sort(np.concatenate(parts)), ids)
    again = hash_split(ids, 4, seed=1)
    assert all(np.array_equal(a, b) for a, b in zip(parts, again))
    assert equalised_num_batches(17) == 17                           # no process group: identity
