"""Named profiler ranges with the reference's labels.

The reference brackets its loop with `torch.autograd.profiler.record_function('gpu-load' | 'gpu-compute')`
(examples/profile/pa_gcn.py:81-92,112) and the cache with 'cache-idxload' / 'cache-index' / 'cache-allocate' /
'cache-gpu' / 'cache-cpu' / 'cache-asign' (PaGraph/storage/storage.py:170-201). Here the same names are emitted as
NVTX ranges (visible to nsys / ncu --nvtx) and as autograd-profiler ranges, behind an env flag so that the default
path pays nothing:

    PG_PROFILE=1 python examples/profile/pa_gcn.py ...      # ranges on
    PG_PROFILE=1 ncu --nvtx --nvtx-include "gpu-load/" ...  # one range's kernels

One library call replaces the reference's index / allocate / gpu / cpu / assign steps, so the cache emits
'cache-idxload' and one 'cache-fetch' range (pg_cache_fetch: split + HBM gather + TMA miss fetch).
"""
import contextlib
import os

_ON = os.environ.get("PG_PROFILE", "0") not in ("", "0")


def enabled():
    return _ON


def enable(on=True):
    global _ON
    _ON = bool(on)


@contextlib.contextmanager
def _range(name):
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        with torch.autograd.profiler.record_function(name):
            yield
    finally:
        torch.cuda.nvtx.range_pop()


_NULL = contextlib.nullcontext()


def range(name):  # noqa: A001 - mirrors nvtx.range
    """Context manager: an NVTX + record_function range called `name` when profiling is enabled, else a no-op."""
    return _range(name) if _ON else _NULL
