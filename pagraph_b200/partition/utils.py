"""get_sub_graph — the k-hop in-neighbour closure of a train-vertex set, as a relabelled sub-graph.

Same outputs as the reference PaGraph/partition/utils.py:9-52 (CSR sub-adjacency with unit uint8
weights, sorted sub->full id map, train ids in sub-graph space). The reference obtains the closure with
ONE full-fanout NeighborSampler batch (fanout = |V|, hence deterministic) and unions the block edges;
here that batch is sampled by the GPU sampler (pg_sample) and the union / relabel stay on the GPU.
"""
import numpy as np
import scipy.sparse as spsp
import torch

from ..sampling import NeighborSampler


def get_sub_graph(dgl_g, train_nid, num_hops):
    train_nid = np.asarray(train_nid, dtype=np.int64)
    nfs = []
    for nf in NeighborSampler(dgl_g, len(train_nid), dgl_g.number_of_nodes(), neighbor_type='in', shuffle=False,
                              num_workers=16, num_hops=num_hops, seed_nodes=train_nid, prefetch=False):
        nfs.append(nf)
    assert (len(nfs) == 1)
    nf = nfs[0]
    dev = nf.device
    node_map = nf._node_mapping.tousertensor()
    full_src, full_dst = [], []
    for i in range(nf.num_blocks):
        lo, hi = nf._layer_offsets[i + 1], nf._layer_offsets[i + 2]
        eb, ee = nf._block_offsets[i], nf._block_offsets[i + 1]
        deg = nf._indptr[lo + 1:hi + 1] - nf._indptr[lo:hi]
        dst = torch.repeat_interleave(torch.arange(lo, hi, device=dev), deg, output_size=ee - eb)
        full_src.append(node_map[nf._indices[eb:ee]])
        full_dst.append(node_map[dst])
    full_srcs, full_dsts = torch.cat(full_src), torch.cat(full_dst)
    # mappings (utils.py:32-37)
    sub2full = torch.unique(torch.cat((full_srcs, full_dsts)))
    full2sub = torch.zeros(int(sub2full.max().item()) + 1, dtype=torch.int64, device=dev)
    full2sub[sub2full] = torch.arange(sub2full.numel(), device=dev)
    sub_srcs, sub_dsts = full2sub[full_srcs], full2sub[full_dsts]
    vnum = sub2full.numel()
    # CSR with duplicate edges merged and unit weights (utils.py:40-44)
    key = torch.unique(sub_srcs * vnum + sub_dsts)
    rows, cols = (key // vnum).cpu().numpy(), (key % vnum).cpu().numpy()
    indptr = np.zeros(vnum + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=vnum), out=indptr[1:])
    idx_dtype = np.int32 if max(vnum, len(cols)) < 2 ** 31 else np.int64
    csr_adj = spsp.csr_matrix((np.ones(len(cols), dtype=np.uint8), cols.astype(idx_dtype), indptr.astype(idx_dtype)),
                              shape=(vnum, vnum))
    print('vertex#: {} edge#: {}'.format(vnum, len(cols)))
    sub2full_np = sub2full.cpu().numpy()
    full2sub_np = full2sub.cpu().numpy()
    # train nid (utils.py:46-51, including the clamp of out-of-range ids)
    tnid = nf.layer_parent_nid(-1).numpy()
    valid_t_max = np.max(sub2full_np)
    valid_t_min = np.min(tnid)
    tnid = np.where(tnid <= valid_t_max, tnid, valid_t_min)
    subtrainid = full2sub_np[np.unique(tnid)]
    return csr_adj, sub2full_np, subtrainid
