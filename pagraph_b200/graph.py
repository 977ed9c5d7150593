"""Read-only graph object standing where `dgl.DGLGraph(adj, readonly=True)` stands in the reference
(examples/profile/pa_gcn.py:36, server/pa_server.py:18, PaGraph/partition/hash.py:26).

Holds the in-CSR (row v = sources of edges u->v, increasing edge id — SURVEY.md Appendix A.1) on the
host and, lazily, in HBM behind a pg_graph handle for the GPU sampler.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _in_csr_from_scipy(adj):
    """scipy matrix with row=src, col=dst -> (indptr, indices, eids) of the in-CSR.

    Edge ids follow DGL 0.4.1: COO input -> position in COO order; CSR input -> position in CSR
    (row-major) order. Rows of the in-CSR list sources in increasing edge id (stable sort by dst).
    """
    import scipy.sparse as spsp
    n = adj.shape[0]
    if spsp.isspmatrix_coo(adj):
        src = np.asarray(adj.row, dtype=np.int64)
        dst = np.asarray(adj.col, dtype=np.int64)
    else:
        csr = adj.tocsr() if not spsp.isspmatrix_csr(adj) else adj
        src = np.repeat(np.arange(n, dtype=np.int64), np.diff(csr.indptr))
        dst = np.asarray(csr.indices, dtype=np.int64)
    order = np.argsort(dst, kind="stable").astype(np.int64)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
    return indptr, src[order], order


class DGLGraph:
    """Minimal read-only graph with the members the PaGraph hot path touches."""

    def __init__(self, graph_data=None, readonly=True, multigraph=None, **_ignored):
        if not readonly:
            raise NotImplementedError("pagraph_b200 graphs are read-only (the reference only builds readonly graphs)")
        self._handles = {}       # device index -> pg_graph*
        self._dev_arrays = {}    # device index -> tensors kept alive for borrowed handles
        self.ndata = {}
        if isinstance(graph_data, DGLGraph):
            self.indptr, self.indices, self.eids = graph_data.indptr, graph_data.indices, graph_data.eids
        elif graph_data is not None:
            self.indptr, self.indices, self.eids = _in_csr_from_scipy(graph_data)
        else:
            self.indptr, self.indices, self.eids = np.zeros(1, np.int64), np.zeros(0, np.int64), None
        self._out_deg = None

    # ---- constructors for already-built CSR
    @classmethod
    def from_in_csr(cls, indptr, indices, eids=None):
        """Host numpy arrays, or CUDA torch tensors (borrowed — no copy, no host mirror)."""
        g = cls()
        if isinstance(indptr, torch.Tensor) and indptr.is_cuda:
            dev = indptr.device.index
            keep = (indptr.contiguous(), indices.contiguous(), None if eids is None else eids.contiguous())
            g._dev_arrays[dev] = keep
            g.indptr = g.indices = g.eids = None
            g._n, g._m = indptr.numel() - 1, indices.numel()
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().pg_graph_create_device(_lib.ptr(keep[0]), _lib.ptr(keep[1]), _lib.ptr(keep[2]),
                                                         g._n, g._m, dev, ctypes.byref(h)), "pg_graph_create_device")
            g._handles[dev] = h
        else:
            g.indptr = np.ascontiguousarray(indptr, np.int64)
            g.indices = np.ascontiguousarray(indices, np.int64)
            g.eids = None if eids is None else np.ascontiguousarray(eids, np.int64)
        return g

    # ---- DGLGraph surface
    @property
    def is_readonly(self):
        return True

    def number_of_nodes(self):
        return self._n if self.indptr is None else len(self.indptr) - 1

    def number_of_edges(self):
        return self._m if self.indptr is None else len(self.indices)

    def in_degrees(self):
        if self.indptr is None:
            ip = self._dev_arrays[next(iter(self._dev_arrays))][0]
            return (ip[1:] - ip[:-1]).cpu()
        return torch.from_numpy(np.diff(self.indptr))

    def out_degrees(self):
        """int64 CPU tensor, as dgl returns (used by auto_cache, storage.py:100)."""
        if self._out_deg is None:
            if self.indptr is None:
                dev = next(iter(self._handles))
                out = torch.empty(self.number_of_nodes(), dtype=torch.int64, device="cuda:%d" % dev)
                with torch.cuda.device(dev):
                    _lib.check(_lib.lib().pg_graph_degrees(self._handles[dev], 0, _lib.ptr(out), _lib.stream_ptr()),
                               "pg_graph_degrees")
                self._out_deg = out.cpu()
            else:
                self._out_deg = torch.from_numpy(np.bincount(self.indices, minlength=self.number_of_nodes())
                                                 .astype(np.int64))
        return self._out_deg

    # ---- device handle
    def handle(self, dev):
        """pg_graph* on device `dev` (uploads the CSR on first use)."""
        if dev not in self._handles:
            if self.indptr is None:
                raise _lib.PGError("graph was built from tensors of another device")
            h = ctypes.c_void_p()
            e = self.eids
            _lib.check(_lib.lib().pg_graph_create(self.indptr.ctypes.data, self.indices.ctypes.data,
                                                  None if e is None else e.ctypes.data,
                                                  self.number_of_nodes(), self.number_of_edges(), dev,
                                                  ctypes.byref(h)), "pg_graph_create")
            self._handles[dev] = h
        return self._handles[dev]

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().pg_graph_destroy(h)
        except Exception:
            pass
