from .gcn_nssc import GCNSampling, GCNInfer
from .graphsage_nssc import GraphSageSampling
