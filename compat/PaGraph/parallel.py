from pagraph_b200.parallel import FlatGradAllReduce, PeerAdam, equalised_num_batches, hash_split  # noqa: F401


class SampleLoader:
    """PaGraph/parallel/dataloader.py:19-100 receives NodeFlows sampled by the server's CPU cores over sockets so that
    sampling does not compete with the trainer. With the sampler on the trainer's own GPU there is nothing to receive."""

    def __init__(self, *a, **k):
        raise NotImplementedError("remote sampling (--remote-sample) is out of scope: use dgl.contrib.sampling.NeighborSampler "
                                  "(GPU) or pagraph_b200.engine")


class SampleDeliver(SampleLoader):
    pass
