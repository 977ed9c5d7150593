"""GraphSAGE trainer entry — drop-in for the reference examples/profile/pa_gs.py (same flags :124-154, same per-GPU
process model, same loop :90-117) on the B200-native hot path.

    python server/pa_server.py --dataset D --num-workers N --model graphsage [--preprocess]
    python examples/profile/pa_gs.py --dataset D --gpu 0,1,..  [--preprocess]

The loop is the reference's op-by-op loop over this package's drop-in classes (GPU sampler, cache object, sm_100a
aggregation behind block_compute, flat-bucket gradient all-reduce). The CUDA-graph engine (pa_gcn.py --engine graph)
drives the GCN model only; GraphSAGE's per-layer self + neighbour update runs through autograd here.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from pa_gcn import init_process, make_parser, save_checkpoint, train_with_engine  # noqa: E402  (shared with the GCN entry)
from pagraph_b200 import DGLGraph  # noqa: E402
from pagraph_b200 import data, graph_store, storage  # noqa: E402
from pagraph_b200.model.graphsage_nssc import GraphSageSampling  # noqa: E402
from pagraph_b200.parallel import FlatGradAllReduce, equalised_num_batches  # noqa: E402
from pagraph_b200.sampling import NeighborSampler  # noqa: E402


def trainer(rank, world_size, args, backend='nccl'):
    init_process(rank, world_size, backend)

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
    embed_names = ['features', 'neigh'] if args.preprocess else ['features']       # pa_gs.py:46-49
    cacher = storage.GraphCacheServer(remote_g, adj.shape[0], t2fid, rank)
    cacher.init_field(embed_names)
    cacher.log = False

    num_hops = args.n_layers if args.preprocess else args.n_layers + 1
    model = GraphSageSampling(args.feat_size, args.n_hidden, n_classes, args.n_layers, F.relu, args.dropout, 'mean',
                              args.preprocess)
    loss_fcn = torch.nn.CrossEntropyLoss()
    model.cuda(rank)
    sync = FlatGradAllReduce(model)                       # stands where DistributedDataParallel stands
    optimizer = torch.optim.Adam(sync.flat_parameters(), lr=args.lr, weight_decay=args.weight_decay)

    fanout = [int(x) for x in str(args.num_neighbors).split(',')]
    if args.engine == 'graph' and not args.preprocess:       # SageTrainEngine: the same loop as a CUDA-graph pipeline
        return train_with_engine(rank, args, g, cacher, model, sync, labels, train_nid, fanout, num_hops, embed_names,
                                 remote_g, arch='gs-nssc')
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
            cacher.fetch_data(nf)
            label = labels[nf.layer_parent_nid_dev(-1)]
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
        save_checkpoint(args, model, epoch, rank, 'gs-nssc')
    print('Total Time: {:.4f}s'.format(time.time() - tic))
    if not args.keep_store:
        remote_g.destroy()
    dist.destroy_process_group()


if __name__ == '__main__':
    parser = make_parser()
    parser.description = 'GraphSAGE'
    parser.set_defaults(n_hidden=16)                        # pa_gs.py:134
    args = parser.parse_args()
    if args.remote_sample:
        raise SystemExit("--remote-sample (server-side CPU sampling, parallel/dataloader.py) has no role when the "
                         "sampler runs on the trainer's GPU; see DESIGN.md 'out of scope'")
    os.environ['CUDA_VISIBLE_DEVICES'] = args.gpu
    gpu_num = len(args.gpu.split(','))
    mp.spawn(trainer, args=(gpu_num, args), nprocs=gpu_num, join=True)
