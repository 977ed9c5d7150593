"""Feature-store server — drop-in for the reference server/pa_server.py:15-110 (same flags).

Loads adj.npz/feat.npy, computes norm = 1/in_degree (pa_server.py:43), optionally folds one hop into the
features (`--preprocess`, pa_server.py:45-52: X' = diag(norm) * A^T X — here the sm_100a aggregation
kernel run over the full in-CSR in row blocks instead of a CPU update_all), publishes `features` / `norm`
(gcn) or `features` (/`neigh`) (graphsage) as shared-memory fields the trainers' GPUs read directly, and
blocks until every trainer has left. `--sample` (server-side sampling for trainers) is out of scope: the
sampler runs on each trainer's GPU.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pagraph_b200 import DGLGraph  # noqa: E402
from pagraph_b200 import data, graph_store  # noqa: E402


def preprocess_features(graph, features, norm, device="cuda:0", rows_per_block=1 << 20):
    """features'[v] = norm[v] * sum_{u->v} features[u] over the whole graph (pa_server.py:45-52)."""
    from pagraph_b200 import ops
    dev = torch.device(device)
    indptr = torch.from_numpy(graph.indptr).to(dev)
    indices = torch.from_numpy(graph.indices).to(dev)
    src = torch.as_tensor(features, dtype=torch.float32).to(dev)
    nrm = torch.as_tensor(norm, dtype=torch.float32).reshape(-1).to(dev)
    V = graph.number_of_nodes()
    out = torch.empty_like(src, device="cpu")
    for lo in range(0, V, rows_per_block):
        hi = min(V, lo + rows_per_block)
        blk = ops.aggregate_forward(indptr[lo:hi + 1], indices, 0, src, hi - lo, "sum", norm=nrm[lo:hi])
        out[lo:hi].copy_(blk)
    return out


def main(args):
    if args.sample:        # refused before anything is published
        raise SystemExit("--sample: server-side sampling is replaced by the GPU sampler in each trainer")
    coo_adj, feat = data.get_graph_data(args.dataset)
    graph = DGLGraph(coo_adj, readonly=True)
    features = torch.as_tensor(np.asarray(feat), dtype=torch.float32)
    graph_name = os.path.basename(args.dataset.rstrip('/'))
    vnum, enum, feat_size = graph.number_of_nodes(), graph.number_of_edges(), features.shape[1]
    print('=' * 30)
    print("Graph Name: {}\nNodes Num: {}\tEdges Num: {}\nFeature Size: {}".format(graph_name, vnum, enum, feat_size))
    print('=' * 30)

    g = graph_store.create_graph_store_server(graph, graph_name, 'shared_mem', args.num_workers, False, edge_dir='in')
    if args.model == 'gcn':
        norm = 1. / graph.in_degrees().float().unsqueeze(1)
        if args.preprocess:
            print('Preprocessing features...')
            features = preprocess_features(graph, features, norm)
        g.ndata['norm'] = norm
        g.ndata['features'] = features
    elif args.model == 'graphsage':
        if args.preprocess:
            print('preprocessing: warning: jusy copy')
            g.ndata['neigh'] = features
        g.ndata['features'] = features
    print('start running graph server on dataset: {}'.format(graph_name))
    g.run()


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='GraphServer')
    parser.add_argument("--dataset", type=str, default=None, help="dataset folder path")
    parser.add_argument("--num-workers", type=int, default=1, help="the number of workers")
    parser.add_argument("--model", type=str, default="gcn", help="model type for preprocessing")
    parser.add_argument("--sample", dest='sample', action='store_true')
    parser.set_defaults(sample=False)
    parser.add_argument("--num-neighbors", type=int, default=2)
    parser.add_argument("--gnn-layers", type=int, default=2)
    parser.add_argument("--batch-size", type=int, default=6000)
    parser.add_argument("--n-epochs", type=int, default=10)
    parser.add_argument("--one2all", dest='one2all', action='store_true')
    parser.set_defaults(one2all=False)
    parser.add_argument("--preprocess", dest='preprocess', action='store_true')
    parser.set_defaults(preprocess=False)
    main(parser.parse_args())
