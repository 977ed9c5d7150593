#!/usr/bin/env python
"""Full-neighbourhood inference + test accuracy — counterpart of the reference's examples/eval.py:13-46 on the B200 path.

Same flags and output line (`[ckpt]: Test Accuracy x`). One NodeFlow holding the complete `num_hops` in-neighbourhood of
the test vertices is built by the GPU sampler (expand_factor = number of nodes: no random draws, eval.py:20-25), its
frames come from the feature server (`nf.copy_from_parent`, eval.py:26), and every checkpoint `<ckpt>/<arch>_<epoch>`
(written by `pa_gcn.py --ckpt DIR` / `pa_gs.py --ckpt DIR`) is evaluated with GCNInfer (sum x norm, gcn_nssc.py:103-164) or
GraphSageSampling in eval mode. The reference samples through the graph store's shared graph; here the structure is
read from the dataset directory (the store only carries the feature tables).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pagraph_b200 import DGLGraph  # noqa: E402
from pagraph_b200 import data, graph_store  # noqa: E402
from pagraph_b200.sampling import NeighborSampler  # noqa: E402


def gnneval(args, infer_model, graph, labels, rank, test_nid):
    """Evaluation on a single GPU; returns [(ckpt, accuracy)]."""
    ctx = torch.device("cuda", rank)
    num_hops = args.n_layers if args.preprocess else args.n_layers + 1
    nf = None
    for nf in NeighborSampler(graph, len(test_nid), graph.number_of_nodes(), neighbor_type='in', num_workers=16,
                              num_hops=num_hops, seed_nodes=test_nid):
        pass
    results = []
    batch_nids = nf.layer_parent_nid(-1)
    batch_labels = labels[batch_nids].cuda(rank)
    for ckpt in range(args.start, args.end, args.interval):
        path = os.path.join(args.ckpt, args.arch + '_' + str(ckpt))
        if not os.path.exists(path):
            continue
        state = torch.load(path, map_location="cpu")
        if not isinstance(state, dict):                      # a pickled module, as the reference saves it
            state = state.state_dict()
        with torch.no_grad():
            for infer_param, param in zip(infer_model.parameters(), state.values()):
                infer_param.data.copy_(param.data)
        infer_model.cuda(ctx)
        infer_model.eval()
        with torch.no_grad():
            nf.copy_from_parent(ctx=ctx)
            pred = infer_model(nf)
            num_acc = (pred.argmax(dim=1) == batch_labels).sum().cpu().item()
        acc = num_acc / len(test_nid)
        results.append((ckpt, acc))
        print("[{}]: Test Accuracy {:.4f}".format(ckpt, acc))
    return results


def main(args):
    dataname = os.path.basename(args.dataset.rstrip('/'))
    store = graph_store.create_graph_from_store(dataname, "shared_mem")
    graph = DGLGraph(data.get_struct(args.dataset), readonly=True)
    names = ['features', 'norm'] if args.arch == 'gcn-nssc' else ['features']
    graph.ndata = {name: store._node_frame._frame[name].data for name in names}
    labels = data.get_labels(args.dataset)
    n_classes = len(np.unique(labels))
    train_mask, val_mask, test_mask = data.get_masks(args.dataset)
    test_nid = np.nonzero(test_mask)[0].astype(np.int64)
    labels = torch.LongTensor(labels)
    if args.arch == 'gcn-nssc':
        from pagraph_b200.model.gcn_nssc import GCNInfer
        infer_model = GCNInfer(args.feat_size, args.n_hidden, n_classes, args.n_layers, F.relu, args.preprocess)
    elif args.arch == 'gs-nssc':
        from pagraph_b200.model.graphsage_nssc import GraphSageSampling
        infer_model = GraphSageSampling(args.feat_size, 16, n_classes, args.n_layers, F.relu, 0, 'mean', args.preprocess)
    else:
        print('Unknown arch')
        sys.exit(-1)
    res = gnneval(args, infer_model, graph, labels, 0, test_nid)
    if not args.keep_store:
        store.destroy()
    return res


def make_parser():
    parser = argparse.ArgumentParser(description='GCNInfer')
    parser.add_argument("--gpu", type=int, default=None, help="gpu id. such as 0 or 1 or 2")
    parser.add_argument("--dataset", type=str, default=None, help="path to the dataset folder")
    parser.add_argument("--arch", type=str, default='gcn-nssc', help='model arch')
    parser.add_argument("--feat-size", type=int, default=602, help='input feature size')
    parser.add_argument("--n-hidden", type=int, default=32, help="hidden units (the reference hard-codes 32)")
    parser.add_argument("--n-layers", type=int, default=1, help="number of hidden gcn layers")
    parser.add_argument("--start", type=int, default=0, help="eval epoch start")
    parser.add_argument("--interval", type=int, default=5, help="eval epoch interval")
    parser.add_argument("--end", type=int, default=60, help='eval epoch end (not include)')
    parser.add_argument("--ckpt", type=str, default='checkpoint', help="checkpoint dir")
    parser.add_argument("--preprocess", dest='preprocess', action='store_true')
    parser.set_defaults(preprocess=False)
    parser.add_argument("--keep-store", action='store_true', help="do not tell the feature server this client is done")
    return parser


if __name__ == '__main__':
    if os.environ.get("PG_DEBUG_HANG"):        # dump every thread's stack and exit if the run takes longer than this many seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PG_DEBUG_HANG"]), exit=True)
    a = make_parser().parse_args()
    if a.gpu is not None:
        os.environ['CUDA_VISIBLE_DEVICES'] = str(a.gpu)
    main(a)
