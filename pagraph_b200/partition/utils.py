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


def _closure_edges(dgl_g, train_nid, num_hops):
    """ONE full-fanout NeighborSampler batch over all of `train_nid` (utils.py:11-18): the union of its block edges is the
    `num_hops`-hop in-neighbour closure. Returns (full_srcs, full_dsts, seed parent ids) on the GPU."""
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
    return torch.cat(full_src), torch.cat(full_dst), nf.layer_parent_nid_dev(-1)


def _relabel(full_srcs, full_dsts, tnid):
    """utils.py:32-51 on the GPU: sorted sub->full map, unique (src, dst) pairs in sub-graph ids (CSR order), train ids."""
    dev = full_srcs.device
    sub2full = torch.unique(torch.cat((full_srcs, full_dsts)))
    full2sub = torch.zeros(int(sub2full.max().item()) + 1, dtype=torch.int64, device=dev)
    full2sub[sub2full] = torch.arange(sub2full.numel(), device=dev)
    vnum = sub2full.numel()
    key = torch.unique(full2sub[full_srcs] * vnum + full2sub[full_dsts])       # duplicate edges merged, row-major order
    # train nid (utils.py:46-51, including the clamp of out-of-range ids and the id-0 fate of isolated train vertices)
    valid_t_max, valid_t_min = sub2full.max(), tnid.min()
    tnid = torch.where(tnid <= valid_t_max, tnid, valid_t_min)
    subtrainid = full2sub[torch.unique(tnid)]
    return sub2full, key, vnum, subtrainid


def get_sub_graph_device(dgl_g, train_nid, num_hops):
    """get_sub_graph without leaving the GPU, for a trainer that partitions in-process: (in_indptr, in_indices,
    sub2full, subtrainid) as CUDA tensors — the in-CSR (row v = sources of u -> v, ascending) a GPU sampler walks, i.e.
    what `DGLGraph(subadj)` builds from the `subadj_{r}.npz` this partition would be saved as."""
    full_srcs, full_dsts, tnid = _closure_edges(dgl_g, train_nid, num_hops)
    sub2full, key, vnum, subtrainid = _relabel(full_srcs, full_dsts, tnid)
    del full_srcs, full_dsts
    src, dst = key // vnum, key % vnum
    del key
    order = torch.sort(dst * vnum + src).values                                 # by destination, then ascending source
    in_indices = order % vnum
    counts = torch.bincount(order // vnum, minlength=vnum)
    in_indptr = torch.zeros(vnum + 1, dtype=torch.int64, device=src.device)
    torch.cumsum(counts, 0, out=in_indptr[1:])
    return in_indptr, in_indices.contiguous(), sub2full, subtrainid


def get_sub_graph(dgl_g, train_nid, num_hops):
    full_srcs, full_dsts, tnid = _closure_edges(dgl_g, train_nid, num_hops)
    sub2full, key, vnum, subtrainid = _relabel(full_srcs, full_dsts, tnid)
    # CSR with duplicate edges merged and unit weights (utils.py:40-44)
    rows, cols = (key // vnum).cpu().numpy(), (key % vnum).cpu().numpy()
    indptr = np.zeros(vnum + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=vnum), out=indptr[1:])
    idx_dtype = np.int32 if max(vnum, len(cols)) < 2 ** 31 else np.int64
    csr_adj = spsp.csr_matrix((np.ones(len(cols), dtype=np.uint8), cols.astype(idx_dtype), indptr.astype(idx_dtype)),
                              shape=(vnum, vnum))
    print('vertex#: {} edge#: {}'.format(vnum, len(cols)))
    return csr_adj, sub2full.cpu().numpy(), subtrainid.cpu().numpy()
