"""compat/: the reference's import lines (examples/profile/pa_gcn.py:10-16, pa_gs.py, server/pa_server.py,
PaGraph/model/gcn_nssc.py:1-5) resolve to the B200 classes without editing the importing file."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import dgl
from dgl import DGLGraph
import dgl.function as fn
from dgl.contrib.sampling import NeighborSampler, SamplerPool
from dgl.frame import Frame, FrameRef
import dgl.utils
from PaGraph.model.gcn_nssc import GCNSampling, GCNInfer
from PaGraph.model.graphsage_nssc import GraphSageSampling
import PaGraph.data as data
import PaGraph.storage as storage
from PaGraph.parallel import SampleLoader, SampleDeliver
import PaGraph.partition
import pagraph_b200, pagraph_b200.storage, pagraph_b200.sampling, pagraph_b200.graph_store, pagraph_b200.data
import pagraph_b200.model.gcn_nssc as m
assert DGLGraph is pagraph_b200.DGLGraph and dgl.DGLGraph is DGLGraph
assert storage.GraphCacheServer is pagraph_b200.storage.GraphCacheServer
assert NeighborSampler is pagraph_b200.sampling.NeighborSampler and dgl.contrib.sampling.NeighborSampler is NeighborSampler
assert dgl.contrib.graph_store.create_graph_from_store is pagraph_b200.graph_store.create_graph_from_store
assert GCNSampling is m.GCNSampling and GCNInfer is m.GCNInfer
assert data.get_sub_train_graph is pagraph_b200.data.get_sub_train_graph
assert fn.copy_src(src='h', out='m').src == 'h' and fn.mean(msg='m', out='h').mode == 'mean' and fn.sum('m', 'h').mode == 'sum'
assert Frame is pagraph_b200.Frame and FrameRef is pagraph_b200.FrameRef
assert len(dgl.utils.toindex([1, 2, 3])) == 3
for cls in (SampleLoader, SampleDeliver, SamplerPool):
    try:
        cls(None, 0)
    except NotImplementedError:
        pass
    else:
        raise AssertionError(cls)
assert hasattr(PaGraph.partition, "dg") and hasattr(PaGraph.partition, "hash")
print("compat ok")
'''


def test_reference_import_lines_resolve():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "compat")]))
    out = subprocess.run([sys.executable, "-c", SNIPPET], env=env, cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "compat ok" in out.stdout
