"""GCN trainer entry — drop-in for the reference examples/profile/pa_gcn.py (same flags :120-150, same
per-GPU process model :157, same loop :86-113) on the B200-native hot path.

    python server/pa_server.py --dataset D --num-workers N [--preprocess]      # feature store
    python examples/profile/pa_gcn.py --dataset D --gpu 0,1,..  [--preprocess]  # one trainer per GPU

Differences from the reference, all below its API: the sampler runs on the GPU (pg_sample), the cache
object fetches every layer with one split + gather + TMA miss-fetch (pg_cache_fetch), block_compute is
the sm_100a aggregation kernel, and the gradient all-reduce is one flat NCCL bucket instead of DDP.
`--num-neighbors` also accepts a per-hop list "25,10" (index 0 expands the seeds).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from pagraph_b200 import DGLGraph  # noqa: E402
from pagraph_b200 import data, graph_store, profiling, storage  # noqa: E402
from pagraph_b200.model.gcn_nssc import GCNSampling  # noqa: E402
from pagraph_b200.parallel import FlatGradAllReduce, equalised_num_batches  # noqa: E402
from pagraph_b200.sampling import NeighborSampler  # noqa: E402


def init_process(rank, world_size, backend):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ.setdefault('MASTER_PORT', '29501')
    dist.init_process_group(backend, rank=rank, world_size=world_size)
    torch.cuda.set_device(rank)
    torch.manual_seed(rank)
    print('rank [{}] process successfully launches'.format(rank))


def trainer(rank, world_size, args, backend='nccl'):
    init_process(rank, world_size, backend)

    # load data
    dataname = os.path.basename(args.dataset.rstrip('/'))
    remote_g = graph_store.create_graph_from_store(dataname, "shared_mem")

    adj, t2fid = data.get_sub_train_graph(args.dataset, rank, world_size)
    g = DGLGraph(adj, readonly=True)
    n_classes = args.n_classes
    train_nid = data.get_sub_train_nid(args.dataset, rank, world_size)
    sub_labels = data.get_sub_train_labels(args.dataset, rank, world_size)
    labels = np.zeros(np.max(train_nid) + 1, dtype=np.int64)
    labels[train_nid] = sub_labels

    t2fid = torch.LongTensor(t2fid)
    labels = torch.LongTensor(labels).cuda(rank)
    embed_names = ['features', 'norm']
    cacher = storage.GraphCacheServer(remote_g, adj.shape[0], t2fid, rank)
    cacher.init_field(embed_names)
    cacher.log = False

    # prepare model
    num_hops = args.n_layers if args.preprocess else args.n_layers + 1
    model = GCNSampling(args.feat_size, args.n_hidden, n_classes, args.n_layers, F.relu, args.dropout,
                        args.preprocess)
    loss_fcn = torch.nn.CrossEntropyLoss()
    model.cuda(rank)
    sync = FlatGradAllReduce(model)                       # stands where DistributedDataParallel stands
    optimizer = torch.optim.Adam(sync.flat_parameters(), lr=args.lr, weight_decay=args.weight_decay)

    fanout = [int(x) for x in str(args.num_neighbors).split(',')]
    if args.engine == 'graph':
        return train_with_engine(rank, args, g, cacher, model, sync, labels, train_nid, fanout, num_hops, embed_names, remote_g)
    fanout = fanout[0] if len(fanout) == 1 else fanout
    sampler = NeighborSampler(g, args.batch_size, fanout, neighbor_type='in', shuffle=True,
                              num_workers=args.num_workers, num_hops=num_hops, seed_nodes=train_nid,
                              prefetch=True, seed=args.seed)
    steps_per_epoch = equalised_num_batches(len(sampler))  # uneven counts would hang the all-reduce

    epoch_dur = []
    tic = time.time()
    for epoch in range(args.n_epochs):
        model.train()
        epoch_start_time = time.time()
        step = 0
        for nf in sampler.batches(0, steps_per_epoch, epoch):
            with profiling.range('gpu-load'):                        # pa_gcn.py:87-91
                cacher.fetch_data(nf)
                label = labels[nf.layer_parent_nid_dev(-1)]
            with profiling.range('gpu-compute'):                     # pa_gcn.py:92-97
                pred = model(nf)
                loss = loss_fcn(pred, label)
                sync.zero_grad()
                loss.backward()
                sync()
                optimizer.step()
            step += 1
            if epoch == 0 and step == 1:
                cacher.auto_cache(g, embed_names)
            if rank == 0 and step % 20 == 0:
                print('epoch [{}] step [{}]. Loss: {:.4f}'.format(epoch + 1, step, loss.item()))
        torch.cuda.synchronize()
        if rank == 0:
            epoch_dur.append(time.time() - epoch_start_time)
            print('Epoch average time: {:.4f}'.format(np.mean(np.array(epoch_dur[2:])) if len(epoch_dur) > 2
                                                      else epoch_dur[-1]))
        if cacher.log:
            print('Epoch average miss rate: {:.4f}'.format(cacher.get_miss_rate()))
        save_checkpoint(args, model, epoch, rank)
    toc = time.time()
    print('Total Time: {:.4f}s'.format(toc - tic))
    if not args.keep_store:
        remote_g.destroy()
    dist.destroy_process_group()


def save_checkpoint(args, model, epoch, rank, arch='gcn-nssc'):
    """`--ckpt DIR` (extension): rank 0 saves the parameters after every epoch as DIR/{arch}_{epoch}, the file name
    examples/eval.py loads (the reference's eval.py:28-32 expects them; nothing in its tree writes them)."""
    if args.ckpt and rank == 0:
        os.makedirs(args.ckpt, exist_ok=True)
        torch.save({k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                   os.path.join(args.ckpt, '{}_{}'.format(arch, epoch)))


def train_with_engine(rank, args, g, cacher, model, sync, labels, train_nid, fanout, num_hops, embed_names, remote_g,
                      arch='gcn-nssc'):
    """Same loop on pagraph_b200.engine.GCNTrainEngine: the minibatch work is two CUDA-graph replays (load + compute)."""
    from pagraph_b200.engine import make_train_engine
    from pagraph_b200.parallel import equalised_num_batches
    optimizer = torch.optim.Adam(sync.flat_parameters(), lr=args.lr, weight_decay=args.weight_decay, capturable=True, fused=True)
    fanouts = fanout * num_hops if len(fanout) == 1 else fanout
    engine = make_train_engine(g, cacher, model, optimizer, train_nid, labels, args.batch_size, fanouts, sync=sync,
                               seed=args.seed, shuffle=True)       # GCN / GCN --preprocess / GraphSAGE pipelines
    steps_per_epoch = equalised_num_batches(engine.num_batches)
    epoch_dur = []
    tic = time.time()
    model.train()
    for epoch in range(args.n_epochs):
        epoch_start_time = time.time()
        engine.next_compute = engine.next_issue = epoch * engine.num_batches     # every epoch starts at its first minibatch
        done = 0
        while done < steps_per_epoch:
            n = 1 if (epoch == 0 and done == 0) else min(20 - done % 20, steps_per_epoch - done)
            loss = engine.steps(n)
            done += n
            if epoch == 0 and done == 1:
                cacher.auto_cache(g, embed_names)
            if rank == 0 and done % 20 == 0:
                print('epoch [{}] step [{}]. Loss: {:.4f}'.format(epoch + 1, done, loss.item()))
        torch.cuda.synchronize()
        if rank == 0:
            epoch_dur.append(time.time() - epoch_start_time)
            print('Epoch average time: {:.4f}'.format(np.mean(np.array(epoch_dur[2:])) if len(epoch_dur) > 2
                                                      else epoch_dur[-1]))
        save_checkpoint(args, model, epoch, rank, arch)
    print('Total Time: {:.4f}s'.format(time.time() - tic))
    engine.close()
    if not args.keep_store:
        remote_g.destroy()
    dist.destroy_process_group()


def make_parser():
    parser = argparse.ArgumentParser(description='GCN')
    parser.add_argument("--gpu", type=str, default='0', help="gpu ids. such as 0 or 0,1,2")
    parser.add_argument("--dataset", type=str, default=None, help="path to the dataset folder")
    # model arch
    parser.add_argument("--feat-size", type=int, default=600, help='input feature size')
    parser.add_argument("--n-classes", type=int, default=60)
    parser.add_argument("--dropout", type=float, default=0.2, help="dropout probability")
    parser.add_argument("--n-hidden", type=int, default=32, help="number of hidden gcn units")
    parser.add_argument("--n-layers", type=int, default=1, help="number of hidden gcn layers")
    parser.add_argument("--preprocess", dest='preprocess', action='store_true')
    parser.set_defaults(preprocess=False)
    # training hyper-params
    parser.add_argument("--lr", type=float, default=3e-2, help="learning rate")
    parser.add_argument("--n-epochs", type=int, default=10, help="number of training epochs")
    parser.add_argument("--batch-size", type=int, default=6000, help="batch size")
    parser.add_argument("--weight-decay", type=float, default=0, help="Weight for L2 loss")
    # sampling hyper-params
    parser.add_argument("--num-neighbors", type=str, default='2', help="neighbors sampled per hop (int or a,b,..)")
    parser.add_argument("--num-workers", type=int, default=16)
    parser.add_argument("--remote-sample", dest='remote_sample', action='store_true')
    parser.set_defaults(remote_sample=False)
    parser.add_argument("--seed", type=int, default=0, help="sampling RNG seed (the reference's is unseeded)")
    parser.add_argument("--ckpt", type=str, default=None, help="directory for per-epoch checkpoints (read by examples/eval.py)")
    parser.add_argument("--keep-store", action='store_true', help="do not tell the feature server this trainer is done")
    parser.add_argument("--engine", default="graph", choices=["graph", "eager"],
                        help="graph: GCNTrainEngine (CUDA-graph pipeline, default); eager: the reference's op-by-op loop")
    return parser


if __name__ == '__main__':
    args = make_parser().parse_args()
    if args.remote_sample:
        raise SystemExit("--remote-sample (server-side CPU sampling, parallel/dataloader.py) has no role when the "
                         "sampler runs on the trainer's GPU; see DESIGN.md 'out of scope'")
    os.environ['CUDA_VISIBLE_DEVICES'] = args.gpu
    gpu_num = len(args.gpu.split(','))
    mp.spawn(trainer, args=(gpu_num, args), nprocs=gpu_num, join=True)
