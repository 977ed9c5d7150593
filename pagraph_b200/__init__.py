"""pagraph_b200 — B200-native minibatch construction + aggregation behind the PaGraph API.

Hot path (hand-written sm_100a CUDA in csrc/, C-ABI in include/pagraph_b200.h):
  sampling.NeighborSampler  -> pg_sample            (k-hop sampling + NodeFlow construction)
  storage.GraphCacheServer  -> pg_cache_fetch/fill  (cache hit/miss split, HBM gather, TMA miss fetch)
  nodeflow.NodeFlow.block_compute -> pg_aggregate_* (segmented mean/sum aggregation)
There is no CPU fallback: the CUDA library must load (see _lib.lib()).
"""
from . import function  # noqa: F401
from .graph import DGLGraph  # noqa: F401
from .nodeflow import NodeFlow, Frame, FrameRef  # noqa: F401

__version__ = "0.1.0"
