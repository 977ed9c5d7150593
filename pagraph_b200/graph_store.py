"""Host feature store standing where `dgl.contrib.graph_store` stands in the reference
(server/pa_server.py:33-36,53-54,78; examples/profile/pa_gcn.py:33; PaGraph/storage/storage.py:128).

Contract kept: the server publishes `{name: [V, dim] float32}` node fields under a graph name; each
trainer attaches zero-copy and reads `g._node_frame._frame[name].data` (CPU tensor indexed by
full-graph id). B200 change: the shared segments are page-locked and device-mapped
(`pg_host_register` over the /dev/shm mapping, or `pg_host_alloc` in-process) so the GPU gather
kernel reads rows directly over PCIe; row strides are padded to 16 bytes so TMA bulk copies apply
(e.g. Reddit's 602 floats -> stride 604). DGL's XML-RPC control plane is out of scope (SURVEY.md §2
#7): rendezvous is a metadata file next to the segments.
"""
import ctypes
import json
import mmap
import os
import time

import numpy as np
import torch

from . import _lib

_SHM_DIR = "/dev/shm"


def _padded_stride(dim):
    return dim if dim < 4 else (dim + 3) // 4 * 4


class _Column:
    def __init__(self, data):
        self.data = data


class _NodeFrame:
    def __init__(self):
        self._frame = {}


class _NData:
    """`g.ndata[name] = tensor` publishes; `g.ndata[name]` reads back."""

    def __init__(self, store):
        self._s = store

    def __setitem__(self, name, value):
        self._s._publish(name, value)

    def __getitem__(self, name):
        return self._s._node_frame._frame[name].data

    def __contains__(self, name):
        return name in self._s._node_frame._frame

    def keys(self):
        return self._s._node_frame._frame.keys()

    def items(self):
        return [(k, v.data) for k, v in self._s._node_frame._frame.items()]


class LocalGraphStore:
    """In-process store: fields live in pinned, device-mapped host memory (pg_host_alloc)."""

    def __init__(self, graph=None, name="local"):
        self.graph, self.name = graph, name
        self._node_frame = _NodeFrame()
        self._allocs = []
        self.ndata = _NData(self)

    def alloc_field(self, name, num_rows, dim):
        """Allocate an uninitialised pinned [num_rows, dim] float32 field and return its tensor."""
        stride = _padded_stride(dim)
        nbytes = max(num_rows * stride * 4, 4)
        p = ctypes.c_void_p()
        _lib.check(_lib.lib().pg_host_alloc(ctypes.byref(p), nbytes), "pg_host_alloc")
        self._allocs.append(p)
        buf = (ctypes.c_float * (num_rows * stride)).from_address(p.value)
        t = torch.frombuffer(buf, dtype=torch.float32, count=num_rows * stride).view(num_rows, stride)[:, :dim]
        self._node_frame._frame[name] = _Column(t)
        return t

    def _publish(self, name, value):
        value = torch.as_tensor(value, dtype=torch.float32)
        if value.dim() == 1:
            value = value.unsqueeze(1)
        t = self.alloc_field(name, value.shape[0], value.shape[1])
        t.copy_(value)

    def run(self):
        return

    def close(self):
        self._node_frame._frame.clear()
        for p in self._allocs:
            try:
                _lib.lib().pg_host_free(p)
            except Exception:
                pass
        self._allocs = []

    def __del__(self):
        self.close()


def _seg_path(graph_name, field, gen=""):
    """One segment per (field, server generation): a restarted server never rewrites a segment that a client of the
    previous run may still have mapped."""
    return os.path.join(_SHM_DIR, "pagraph_%s__%s%s.f32" % (graph_name, field, "." + gen if gen else ""))


def _purge(graph_name):
    """Remove every file a previous (possibly crashed) server of this name left under /dev/shm."""
    for f in os.listdir(_SHM_DIR):
        if f.startswith("pagraph_%s." % graph_name) or f.startswith("pagraph_%s__" % graph_name):
            try:
                os.unlink(os.path.join(_SHM_DIR, f))
            except FileNotFoundError:
                pass


def _meta_path(graph_name):
    return os.path.join(_SHM_DIR, "pagraph_%s.meta.json" % graph_name)


class SharedMemoryStoreServer:
    """Publishes fields as files in /dev/shm (one POSIX-shm segment per field)."""

    def __init__(self, graph, graph_name, num_workers=1):
        self.graph, self.name, self.num_workers = graph, graph_name, num_workers
        self._node_frame = _NodeFrame()
        self.gen = "%x%x" % (os.getpid(), int(time.time() * 1e3))
        self._meta = {"fields": {}, "num_workers": num_workers, "gen": self.gen}
        self._maps = []
        self._pending = {}
        self.ndata = _NData(self)
        _purge(graph_name)          # stale metadata / segments / .done markers of an earlier run, before any alloc

    def alloc_field(self, name, rows, dim):
        """Create the shared segment of field `name` and return its [rows, dim] tensor to be filled in
        place (no staging copy of a 24 GB table); `commit()` then makes it visible to clients."""
        stride = _padded_stride(dim)
        path = _seg_path(self.name, name, self.gen)
        arr = np.memmap(path, mode="w+", dtype=np.float32, shape=(rows, stride))
        self._maps.append(arr)
        t = torch.from_numpy(arr)[:, :dim]
        self._node_frame._frame[name] = _Column(t)
        self._pending[name] = {"rows": rows, "dim": dim, "stride": stride}
        return t

    def commit(self):
        """Publish the metadata of every allocated field (clients poll for it)."""
        self._meta["fields"].update(self._pending)
        self._pending = {}
        with open(_meta_path(self.name) + ".tmp", "w") as f:
            json.dump(self._meta, f)
        os.replace(_meta_path(self.name) + ".tmp", _meta_path(self.name))

    def _publish(self, name, value):
        value = torch.as_tensor(value, dtype=torch.float32)
        if value.dim() == 1:
            value = value.unsqueeze(1)
        t = self.alloc_field(name, value.shape[0], value.shape[1])
        t.copy_(value)
        self.commit()

    def run(self, poll_s=0.5, timeout_s=None):
        """Block until `num_workers` clients have signalled completion (DGL: until all disconnect)."""
        t0 = time.time()
        while True:
            done = [f for f in os.listdir(_SHM_DIR)
                    if f.startswith("pagraph_%s." % self.name) and f.endswith(".done")]
            if len(done) >= self.num_workers:
                break
            if timeout_s is not None and time.time() - t0 > timeout_s:
                break
            time.sleep(poll_s)
        self.destroy()

    def destroy(self):
        self._node_frame._frame.clear()
        self._maps = []
        _purge(self.name)


class _LazyFrame(dict):
    """Field table of a client: a field the server publishes after the client attached is picked up on first use."""

    def __init__(self, client):
        super().__init__()
        self._client = client

    def __missing__(self, name):
        self._client._attach(expect_fields=[name])
        return dict.__getitem__(self, name)


class SharedMemoryStoreClient:
    """Attaches to a server's segments; pins + maps them for the GPU."""

    def __init__(self, graph_name, wait_s=600.0, expect_fields=None):
        self.name = graph_name
        self._wait_s = wait_s
        self._node_frame = _NodeFrame()
        self._node_frame._frame = _LazyFrame(self)
        self._maps = []
        self._registered = []
        self.ndata = _NData(self)
        self._attach(expect_fields)

    def _attach(self, expect_fields=None):
        graph_name = self.name
        t0 = time.time()
        while True:
            try:
                with open(_meta_path(graph_name)) as f:
                    meta = json.load(f)
                if expect_fields is None or all(k in meta["fields"] for k in expect_fields):
                    break
            except (FileNotFoundError, json.JSONDecodeError):
                pass
            if time.time() - t0 > self._wait_s:
                raise TimeoutError("graph store %r (fields %s) did not appear under %s" % (graph_name, expect_fields, _SHM_DIR))
            time.sleep(0.2)
        gen = meta.get("gen", "")
        if getattr(self, "_gen", None) not in (None, gen):
            raise RuntimeError("graph store %r was restarted (generation %s -> %s) under a live client" % (graph_name, self._gen, gen))
        self._gen = gen
        for name, m in meta["fields"].items():
            if dict.__contains__(self._node_frame._frame, name):
                continue
            try:
                fd = os.open(_seg_path(graph_name, name, gen), os.O_RDWR)
            except FileNotFoundError:      # metadata of a server that has just been replaced: wait for the new one
                self._gen = None
                time.sleep(0.2)
                return self._attach(expect_fields)
            try:
                mm = mmap.mmap(fd, m["rows"] * m["stride"] * 4, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
            finally:
                os.close(fd)
            self._maps.append(mm)
            t = torch.frombuffer(mm, dtype=torch.float32, count=m["rows"] * m["stride"]).view(m["rows"], m["stride"])
            self._node_frame._frame[name] = _Column(t[:, :m["dim"]])

    def _publish(self, name, value):
        raise RuntimeError("graph store clients are read-only")

    def destroy(self):
        """Signal the server that this worker is done."""
        open(os.path.join(_SHM_DIR, "pagraph_%s.%d.done" % (self.name, os.getpid())), "w").close()


def create_graph_store_server(graph_data, graph_name, store_type="shared_mem", num_workers=1,
                              multigraph=False, edge_dir='in', port=8000):
    """Signature of dgl.contrib.graph_store.create_graph_store_server (server/pa_server.py:33-36)."""
    if store_type != "shared_mem":
        raise ValueError("only the 'shared_mem' store exists")
    return SharedMemoryStoreServer(graph_data, graph_name, num_workers)


def create_graph_from_store(graph_name, store_type="shared_mem", port=8000):
    """Signature of dgl.contrib.graph_store.create_graph_from_store (examples/profile/pa_gcn.py:33)."""
    if store_type != "shared_mem":
        raise ValueError("only the 'shared_mem' store exists")
    return SharedMemoryStoreClient(graph_name)
