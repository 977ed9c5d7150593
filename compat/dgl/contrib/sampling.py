from pagraph_b200.sampling import NeighborSampler  # noqa: F401


def _remote(name):
    class _Remote:
        def __init__(self, *a, **k):
            raise NotImplementedError("dgl.contrib.sampling.%s (sampling over sockets, PaGraph/parallel/dataloader.py) is out "
                                      "of scope: NeighborSampler runs on the trainer's GPU" % name)
    _Remote.__name__ = name
    return _Remote


SamplerPool, SamplerSender, SamplerReceiver = _remote("SamplerPool"), _remote("SamplerSender"), _remote("SamplerReceiver")
