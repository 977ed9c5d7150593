"""GraphSAGE with neighbour sampling — architecture, parameter names and forward semantics of the
reference PaGraph/model/graphsage_nssc.py (NodeUpdate :6-30, GraphSageSampling :33-134). Aggregators
on the rebuilt path: 'mean' (what examples/profile/pa_gs.py trains) and 'gcn' (sum)."""
import torch
import torch.nn as nn

from .. import function as fn
from .gcn_nssc import apply_dropout


class NodeUpdate(nn.Module):
    def __init__(self, in_feats, out_feats, activation=None, concat=False):
        super().__init__()
        self.fc_neigh = nn.Linear(in_feats, out_feats)
        self.fc_self = nn.Linear(in_feats, out_feats)
        self.activation = activation
        self.concat = concat
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)

    def forward(self, node):
        h = self.fc_self(node.data['h']) + self.fc_neigh(node.data['neigh'])
        if self.concat:
            h = torch.cat((h, self.activation(h)), dim=1)
        elif self.activation:
            h = self.activation(h)
        return {'activation': h}


class GraphSageSampling(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation=None, dropout=0.,
                 aggregator_type='pool', preprocess=False):
        super().__init__()
        if aggregator_type not in ('mean', 'gcn'):
            raise KeyError('Aggregator type {} is not on the rebuilt hot path (mean | gcn).'.format(aggregator_type))
        self.preprocess = preprocess
        self.n_layers = n_layers
        self.dropout = nn.Dropout(dropout)
        self.activation = activation
        self.aggregator_type = aggregator_type
        self.layers = nn.ModuleList()
        self.reducer = nn.ModuleList()   # kept for state_dict compatibility (lstm aggregator in the reference)
        if preprocess:
            self.fc_self = nn.Linear(in_feats, n_hidden)
            self.fc_neigh = nn.Linear(in_feats, n_hidden)
        else:
            self.layers.append(NodeUpdate(in_feats, n_hidden, activation, concat=(n_layers == 1)))
        for i in range(1, n_layers):
            self.layers.append(NodeUpdate(n_hidden, n_hidden, activation, concat=(i == n_layers - 1)))
        self.layers.append(NodeUpdate(2 * n_hidden, n_classes))

    def forward(self, nf):
        reduce_fn = fn.mean if self.aggregator_type == 'mean' else fn.sum
        if self.preprocess:
            for i in range(nf.num_layers):
                h = nf.layers[i].data.pop('features')
                h = self.dropout(h.materialize() if hasattr(h, "materialize") else h)
                neigh = nf.layers[i].data.pop('neigh')
                h = self.fc_self(h) + self.fc_neigh(neigh.materialize() if hasattr(neigh, "materialize") else neigh)
                if self.n_layers == 1:
                    h = torch.cat((h, self.activation(h)), dim=1)
                else:
                    h = self.activation(h)
                nf.layers[i].data['h'] = h
        else:
            for lid in range(nf.num_layers):
                nf.layers[lid].data['h'] = nf.layers[lid].data.pop('features')
        # layer `lid` is applied to every remaining block, so each NodeFlow layer keeps a current 'h'
        for lid, layer in enumerate(self.layers):
            for i in range(lid, nf.num_layers - 1):
                nf.layers[i].data['h'] = apply_dropout(self.dropout, nf.layers[i].data.pop('h'))
                nf.block_compute(i, fn.copy_src(src='h', out='m'), reduce_fn('m', 'neigh'), layer)
            for i in range(lid + 1, nf.num_layers):
                nf.layers[i].data['h'] = nf.layers[i].data.pop('activation')
        return nf.layers[nf.num_layers - 1].data.pop('h')
