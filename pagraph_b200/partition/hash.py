"""Hash partitioner — drop-in for the reference PaGraph/partition/hash.py:14-70 (same flags, same
`{P}naive/` outputs): shuffle the train ids, cut them into P equal chunks, expand every chunk to its
`--num-hops` in-neighbour closure. `--seed` makes the shuffle reproducible (the reference's is not).

    python -m pagraph_b200.partition.hash --dataset D --partition P --num-hops H
"""
import argparse
import os

import numpy as np
import scipy.sparse as spsp

from .. import data
from ..graph import DGLGraph
from ..parallel import hash_split
from .utils import get_sub_graph


def save_partition(partition_dataset, pid, subadj, subtrainid, sub2fullid, sublabel):
    spsp.save_npz(os.path.join(partition_dataset, 'subadj_{}.npz'.format(pid)), subadj)
    np.save(os.path.join(partition_dataset, 'sub_trainid_{}.npy'.format(pid)), subtrainid)
    np.save(os.path.join(partition_dataset, 'sub_train2fullid_{}.npy'.format(pid)), sub2fullid)
    np.save(os.path.join(partition_dataset, 'sub_label_{}.npy'.format(pid)), sublabel)


def main(args):
    adj = spsp.load_npz(os.path.join(args.dataset, 'adj.npz'))
    dgl_g = DGLGraph(adj, readonly=True)
    train_mask, _, _ = data.get_masks(args.dataset)
    train_nid = np.nonzero(train_mask)[0].astype(np.int64)
    labels = data.get_labels(args.dataset)
    partition_dataset = os.path.join(args.dataset, '{}naive'.format(args.partition))
    os.makedirs(partition_dataset, exist_ok=True)
    for pid, part_nid in enumerate(hash_split(train_nid, args.partition, seed=args.seed)):
        subadj, sub2fullid, subtrainid = get_sub_graph(dgl_g, part_nid, args.num_hops)
        sublabel = labels[sub2fullid[subtrainid]]
        save_partition(partition_dataset, pid, subadj, subtrainid, sub2fullid, sublabel)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description='Hash')
    parser.add_argument("--dataset", type=str, default=None, help="path to the dataset folder")
    parser.add_argument("--num-hops", type=int, default=1, help="num hops for the extended graph")
    parser.add_argument("--partition", type=int, default=2, help="partition number")
    parser.add_argument("--seed", type=int, default=None, help="shuffle seed (reference: unseeded)")
    main(parser.parse_args())
