"""`dgl` name shim: the handful of dgl==0.4.1 names the PaGraph hot path imports, bound to pagraph_b200 (see ../README.md)."""
from pagraph_b200 import DGLGraph, NodeFlow  # noqa: F401
from pagraph_b200 import function  # noqa: F401

from . import contrib, frame, utils  # noqa: F401,E402

__version__ = "0.4.1+pagraph_b200"
