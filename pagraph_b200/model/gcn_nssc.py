"""GCN with neighbour sampling + skip-concat on the last hidden layer — same architecture, parameter
names and forward semantics as the reference PaGraph/model/gcn_nssc.py (NodeUpdate :6-24,
GCNSampling :27-100, GCNInfer :103-164), so state_dicts interchange. The only difference is under
`nf.block_compute`: the copy_src + mean/sum reducer is the sm_100a aggregation kernel."""
import torch
import torch.nn as nn

from .. import function as fn


class NodeUpdate(nn.Module):
    def __init__(self, in_feats, out_feats, activation=None, test=False, concat=False):
        super().__init__()
        self.linear = nn.Linear(in_feats, out_feats)
        self.activation = activation
        self.concat = concat
        self.test = test

    def forward(self, node):
        h = node.data['h']
        if self.test:                       # inference: sum-aggregate then scale by 1/in_degree
            h = h * node.data['norm']
        if self.concat and self.activation is torch.relu or self.concat and self.activation is torch.nn.functional.relu:
            from ..ops import LinearConcat   # fused linear + bias + cat(h, relu(h)) when the input carries no gradient
            if LinearConcat.supported(h, self.linear.weight):
                return {'activation': LinearConcat.apply(h, self.linear.weight, self.linear.bias, True)}
        h = self.linear(h)
        if self.concat:                     # skip connection
            h = torch.cat((h, self.activation(h)), dim=1)
        elif self.activation:
            h = self.activation(h)
        return {'activation': h}


def _build_layers(owner, in_feats, n_hidden, n_classes, n_layers, activation, preprocess, test):
    owner.layers = nn.ModuleList()
    if preprocess:
        owner.linear = nn.Linear(in_feats, n_hidden)
        owner.activation = activation
    else:
        owner.layers.append(NodeUpdate(in_feats, n_hidden, activation, test=test, concat=(n_layers == 1)))
    for i in range(1, n_layers):
        owner.layers.append(NodeUpdate(n_hidden, n_hidden, activation, test=test, concat=(i == n_layers - 1)))
    owner.layers.append(NodeUpdate(2 * n_hidden, n_classes, test=test))


def apply_dropout(dropout, h):
    """dropout(h); rows still living in the feature cache (LazyCacheRows) get the mask folded into the fused
    aggregation instead of being materialised."""
    if hasattr(h, "with_dropout"):
        return h.with_dropout(dropout.p if dropout.training else 0.0)
    return dropout(h)


class _GCNBase(nn.Module):
    _reduce = staticmethod(fn.mean)

    def _input_transform(self, nf, dropout):
        h = nf.layers[0].data['features']
        if hasattr(h, "materialize"):
            h = h.materialize()
        if dropout is not None:
            h = dropout(h)
        h = self.linear(h)
        if self.n_layers == 1:
            return torch.cat((h, self.activation(h)), dim=1)
        return self.activation(h)

    def _run(self, nf, dropout):
        if self.preprocess:
            h = self._input_transform(nf, dropout)
            for i, layer in enumerate(self.layers):
                nf.layers[i].data['h'] = h
                nf.block_compute(i, fn.copy_src(src='h', out='m'), self._reduce(msg='m', out='h'), layer)
                h = nf.layers[i + 1].data.pop('activation')
            return h
        nf.layers[0].data['activation'] = nf.layers[0].data['features']
        for i, layer in enumerate(self.layers):
            h = nf.layers[i].data.pop('activation')
            if dropout is not None:
                h = apply_dropout(dropout, h)
            nf.layers[i].data['h'] = h
            nf.block_compute(i, fn.copy_src(src='h', out='m'), self._reduce(msg='m', out='h'), layer)
        return nf.layers[-1].data.pop('activation')


class GCNSampling(_GCNBase):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout, preprocess=False):
        super().__init__()
        self.preprocess = preprocess
        self.n_layers = n_layers
        self.dropout = nn.Dropout(p=dropout) if dropout != 0 else None
        _build_layers(self, in_feats, n_hidden, n_classes, n_layers, activation, preprocess, test=False)

    def forward(self, nf):
        return self._run(nf, self.dropout)


class GCNInfer(_GCNBase):
    _reduce = staticmethod(fn.sum)

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, preprocess=False):
        super().__init__()
        self.preprocess = preprocess
        self.n_layers = n_layers
        _build_layers(self, in_feats, n_hidden, n_classes, n_layers, activation, preprocess, test=True)

    def forward(self, nf):
        return self._run(nf, None)
