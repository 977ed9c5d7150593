"""dgl.utils.toindex as PaGraph/storage/storage.py's commented-out call sites use it: an index wrapper."""
import torch


class Index:
    def __init__(self, data):
        self._t = torch.as_tensor(data, dtype=torch.int64)

    def tousertensor(self, ctx=None):
        return self._t

    def __len__(self):
        return self._t.numel()


def toindex(data):
    return data if isinstance(data, Index) else Index(data)
