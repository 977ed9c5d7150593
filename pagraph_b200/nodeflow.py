"""NodeFlow: the layered minibatch container (stands where `dgl.NodeFlow` stands in the reference).

Members mirror what the reference touches: `_node_mapping.tousertensor()`,
`_layer_offsets`, `num_layers`, `_node_frames[i] = ...` (PaGraph/storage/storage.py:171-173,202,216),
`layer_parent_nid` (storage.py:211, examples/profile/pa_gcn.py:89), `layers[i].data` and
`block_compute` (PaGraph/model/gcn_nssc.py:64-74), `num_blocks` / `block_edges` / `map_to_parent_nid`
(PaGraph/partition/utils.py:25-28), `layer_nid` (examples/count_vnum.py:19), `copy_from_parent`
(examples/profile/dgl_gcn.py:83). Layer 0 = inputs ... layer L = seeds (Appendix A.4).

All arrays live in HBM; only the (L+2)+(L+1) offsets are host integers.
"""
import torch

from . import function as fn
from . import ops


class Frame(dict):
    """{field name: tensor [n, dim]} — stands for dgl.frame.Frame."""

    def __init__(self, data=None):
        super().__init__(data or {})


class FrameRef:
    """Mutable mapping over a Frame — stands for dgl.frame.FrameRef."""

    def __init__(self, frame=None):
        self._frame = frame if frame is not None else Frame()

    def __getitem__(self, k):
        return self._frame[k]

    def __setitem__(self, k, v):
        self._frame[k] = v

    def __contains__(self, k):
        return k in self._frame

    def __iter__(self):
        return iter(self._frame)

    def __len__(self):
        return len(self._frame)

    def pop(self, k, *default):
        return self._frame.pop(k, *default)

    def keys(self):
        return self._frame.keys()

    def items(self):
        return self._frame.items()

    def update(self, other):
        self._frame.update(other)


class _Index:
    def __init__(self, t):
        self._t = t

    def tousertensor(self):
        return self._t


class _LayerView:
    def __init__(self, nf, i):
        self._nf, self._i = nf, i

    @property
    def data(self):
        return self._nf._frame_of(self._i)


class _Layers:
    def __init__(self, nf):
        self._nf = nf

    def __getitem__(self, i):
        return _LayerView(self._nf, i % self._nf.num_layers)

    def __len__(self):
        return self._nf.num_layers


class NodeBatch:
    """What apply_node_func receives: `.data` is the destination layer's frame."""

    def __init__(self, data):
        self.data = data


class NodeFlow:
    def __init__(self, node_mapping, indptr, indices, edge_mapping, layer_offsets, flow_offsets,
                 seeds_cpu=None, parent=None):
        self._layer_offsets = [int(x) for x in layer_offsets]
        self._block_offsets = [int(x) for x in flow_offsets]
        n, e = self._layer_offsets[-1], self._block_offsets[-1]
        self._node_mapping = _Index(node_mapping[:n])
        self._edge_mapping = _Index(edge_mapping[:e])
        self._indptr = indptr[:n + 1]
        self._indices = indices[:e]
        self._node_frames = [FrameRef(Frame()) for _ in range(self.num_layers)]
        self._seeds_cpu = seeds_cpu
        self._parent = parent
        self._parent_nid_cpu = None
        self.layers = _Layers(self)

    # ---- structure
    @property
    def num_layers(self):
        return len(self._layer_offsets) - 1

    @property
    def num_blocks(self):
        return self.num_layers - 1

    @property
    def device(self):
        return self._indptr.device

    def number_of_nodes(self):
        return self._layer_offsets[-1]

    def number_of_edges(self):
        return self._block_offsets[-1]

    def layer_size(self, i):
        i %= self.num_layers
        return self._layer_offsets[i + 1] - self._layer_offsets[i]

    def block_size(self, i):
        return self._block_offsets[i + 1] - self._block_offsets[i]

    def layer_nid(self, i):
        i %= self.num_layers
        return torch.arange(self._layer_offsets[i], self._layer_offsets[i + 1], dtype=torch.int64)

    def layer_parent_nid_dev(self, i):
        """Parent ids of layer i as a CUDA tensor view (no copy, no sync)."""
        i %= self.num_layers
        return self._node_mapping._t[self._layer_offsets[i]:self._layer_offsets[i + 1]]

    def layer_parent_nid(self, i):
        """Parent ids of layer i as a CPU tensor (what dgl returns). The seed layer is served from
        the host copy of the seed batch when no duplicate was dropped; other layers cost one D2H."""
        i %= self.num_layers
        if i == self.num_layers - 1 and self._seeds_cpu is not None and len(self._seeds_cpu) == self.layer_size(i):
            return self._seeds_cpu
        if self._parent_nid_cpu is None:
            self._parent_nid_cpu = self._node_mapping._t.cpu()
        return self._parent_nid_cpu[self._layer_offsets[i]:self._layer_offsets[i + 1]]

    def map_to_parent_nid(self, nfids):
        if not torch.is_tensor(nfids):
            nfids = torch.as_tensor(nfids, dtype=torch.int64)
        out = self._node_mapping._t[nfids.to(self.device)]
        return out if nfids.is_cuda else out.cpu()

    def map_to_parent_eid(self, efids):
        if not torch.is_tensor(efids):
            efids = torch.as_tensor(efids, dtype=torch.int64)
        out = self._edge_mapping._t[efids.to(self.device)]
        return out if efids.is_cuda else out.cpu()

    def block_csr(self, i):
        """(indptr over the rows of layer i+1 [absolute edge offsets], cols (NodeFlow ids), col_base,
        n_dst, n_src) — the arguments of pg_aggregate_*."""
        lo, hi = self._layer_offsets[i + 1], self._layer_offsets[i + 2]
        return (self._indptr[lo:hi + 1], self._indices, self._layer_offsets[i], hi - lo,
                self._layer_offsets[i + 1] - self._layer_offsets[i])

    def block_edges(self, i, remap_local=False):
        """(src, dst, eid) NodeFlow ids of block i as CPU tensors (dgl semantics)."""
        lo, hi = self._layer_offsets[i + 1], self._layer_offsets[i + 2]
        eb, ee = self._block_offsets[i], self._block_offsets[i + 1]
        deg = self._indptr[lo + 1:hi + 1] - self._indptr[lo:hi]
        dst = torch.repeat_interleave(torch.arange(lo, hi, device=self.device), deg, output_size=ee - eb)
        src = self._indices[eb:ee]
        eid = torch.arange(eb, ee, device=self.device)
        if remap_local:
            src = src - self._layer_offsets[i]
            dst = dst - lo
        return src.cpu(), dst.cpu(), eid.cpu()

    # ---- frames
    def _frame_of(self, i):
        fr = self._node_frames[i]
        if fr is None:
            fr = self._node_frames[i] = FrameRef(Frame())
        elif isinstance(fr, dict) and not isinstance(fr, FrameRef):
            fr = self._node_frames[i] = FrameRef(fr if isinstance(fr, Frame) else Frame(fr))
        return fr

    def copy_from_parent(self, ctx=None):
        """No-cache baseline (examples/profile/dgl_gcn.py:83): frame_i = parent.ndata[parent ids]."""
        if self._parent is None or not self._parent.ndata:
            raise RuntimeError("copy_from_parent: the parent graph has no node data")
        dev = self.device if ctx is None else torch.device(ctx)
        ids_cpu = None
        for i in range(self.num_layers):
            frame = Frame()
            for name, table in self._parent.ndata.items():
                if table.is_cuda:
                    frame[name] = table[self.layer_parent_nid_dev(i)].to(dev)
                else:
                    if ids_cpu is None:
                        ids_cpu = self._node_mapping._t.cpu()
                    rows = table[ids_cpu[self._layer_offsets[i]:self._layer_offsets[i + 1]]]
                    frame[name] = rows.to(dev, non_blocking=True)
            self._node_frames[i] = FrameRef(frame)

    # ---- compute
    def block_compute(self, block_id, message_func, reduce_func, apply_node_func=None):
        """dst.data[out] = reduce over sampled in-edges of src.data[field]; then apply_node_func."""
        if not isinstance(message_func, fn.CopySrc) or not isinstance(reduce_func, fn.Reduce):
            raise NotImplementedError("block_compute supports the builtin copy_src + sum/mean pair "
                                      "(the only one the reference models use)")
        if reduce_func.msg != message_func.out:
            raise KeyError("reduce reads message field %r but copy_src writes %r" % (reduce_func.msg, message_func.out))
        indptr, cols, col_base, n_dst, n_src = self.block_csr(block_id)
        h = self._frame_of(block_id)[message_func.src]
        if hasattr(h, "materialize"):        # LazyCacheRows (GraphCacheServer.lazy_input): reduce straight from the cache
            c = h.cacher
            out = ops.cache_aggregate(c, h.name, h.parent_ids, indptr, cols, col_base, n_src, n_dst, reduce_func.mode,
                                      dropout_p=h.dropout_p, seed=c.next_dropout_seed() if h.dropout_p > 0 else 0)
        else:
            out = ops.BlockAggregate.apply(h, indptr, cols, col_base, n_dst, reduce_func.mode)
        dst_frame = self._frame_of(block_id + 1)
        dst_frame[reduce_func.out] = out
        if apply_node_func is not None:
            ret = apply_node_func(NodeBatch(dst_frame))
            if ret:
                dst_frame.update(ret)
