"""dg partitioner — drop-in for the reference PaGraph/partition/dg.py (same `dg(partition_num, adj, train_nids, hops)`
function, same CLI flags and `{P}naive/` outputs). The streaming assignment loop runs in native code
(pg_partition_dg, pagraph_b200/csrc/pg_partition.cu) and reproduces the reference's assignments exactly
(tests/golden/dg_*.npz were produced by the real dg.py); the per-partition closure is get_sub_graph on the GPU.

    python -m pagraph_b200.partition.dg --dataset D --partition P --num-hops H

The reference script defines `--num-hops` but reads `args.num_hop` (dg.py:113,129,141,153) and cannot run as shipped;
both spellings are accepted here. `--ordering` (optional degree re-ordering, partition/ordering.py) is out of scope.
"""
import argparse
import os

import numpy as np
import scipy.sparse as spsp

from .. import _lib, data
from ..graph import DGLGraph
from .hash import save_partition
from .utils import get_sub_graph


def dg(partition_num, adj, train_nids, hops):
    """(sub_v, sub_trainv): per partition, the vertex set with redundancy and the train vertices (dg.py:59-103)."""
    csc = adj.tocsc()
    csc.sum_duplicates()
    csc.sort_indices()
    vnum = adj.shape[0]
    indptr = np.ascontiguousarray(csc.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(csc.indices, dtype=np.int64)
    train = np.ascontiguousarray(train_nids, dtype=np.int64)
    belongs = np.empty(vnum, dtype=np.int8)
    member = np.empty((partition_num, vnum), dtype=np.uint8)
    print('total vertices: {} | train vertices: {}'.format(vnum, train.shape[0]))
    _lib.check(_lib.lib().pg_partition_dg(indptr.ctypes.data, indices.ctypes.data, vnum, train.ctypes.data, len(train),
                                          partition_num, hops, belongs.ctypes.data, member.ctypes.data), "pg_partition_dg")
    sub_v, sub_trainv = [], []
    for pid in range(partition_num):
        sub_trainv.append(np.where(belongs == pid)[0])
        sub_v.append(np.where(member[pid] != 0)[0])
        print('vertex# with self-reliance: ', len(sub_v[-1]))
        print('vertex# w/o  self-reliance: ', len(sub_trainv[-1]))
    return sub_v, sub_trainv


def main(args):
    adj = spsp.load_npz(os.path.join(args.dataset, 'adj.npz'))
    train_mask, _, _ = data.get_masks(args.dataset)
    train_nids = np.nonzero(train_mask)[0].astype(np.int64)
    labels = data.get_labels(args.dataset)
    _, p_trainv = dg(args.partition, adj, train_nids, args.num_hops)
    partition_dataset = os.path.join(args.dataset, '{}naive'.format(args.partition))
    os.makedirs(partition_dataset, exist_ok=True)
    dgl_g = DGLGraph(adj, readonly=True)
    for pid, ptrainv in enumerate(p_trainv):
        print('generating subgraph# {}...'.format(pid))
        subadj, sub2fullid, subtrainid = get_sub_graph(dgl_g, ptrainv, args.num_hops)
        sublabel = labels[sub2fullid[subtrainid]]
        save_partition(partition_dataset, pid, subadj, subtrainid, sub2fullid, sublabel)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Partition')
    parser.add_argument("--dataset", type=str, default=None, help="dataset dir")
    parser.add_argument("--partition", type=int, default=2, help="num of partitions")
    parser.add_argument("--num-hops", "--num-hop", dest="num_hops", type=int, default=1,
                        help="num of hop neighbors required for a batch")
    parser.add_argument("--ordering", dest='ordering', action='store_true')
    parser.set_defaults(ordering=False)
    a = parser.parse_args()
    if a.ordering:
        raise SystemExit("--ordering (partition/ordering.py) is not part of the rebuilt path")
    main(a)
