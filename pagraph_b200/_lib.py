"""ctypes binding of libpagraph_b200.so (include/pagraph_b200.h). No CPU fallback: if the CUDA
library is missing or a call fails, this raises."""
import ctypes
import os

from . import build as _build

PG_MAX_FIELDS = 4
PG_MAX_HOPS = 8
PG_MAX_RANKS = 8
PG_IPC_HANDLE_BYTES = 64
PG_META_LEN = 4 + (PG_MAX_HOPS + 2) + (PG_MAX_HOPS + 1)
PG_OK, PG_ERR_INVALID, PG_ERR_CUDA, PG_ERR_OVERFLOW, PG_ERR_NOMEM = 0, 1, 2, 3, 4
PG_AGG_SUM, PG_AGG_MEAN = 0, 1

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p


class PGError(RuntimeError):
    pass


class pg_field(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("host_stride", ctypes.c_int64), ("host_table", c_vp)]


class pg_nodeflow_buffers(ctypes.Structure):
    _fields_ = [("node_mapping", c_vp), ("indptr", c_vp), ("indices", c_vp), ("edge_mapping", c_vp),
                ("meta", c_vp)]


class pg_block(ctypes.Structure):
    _fields_ = [("parent_ids", c_vp), ("indptr", c_vp), ("cols", c_vp), ("col_base", ctypes.c_int64),
                ("n_src", ctypes.c_int64), ("n_dst", ctypes.c_int64), ("d_layer_offsets", c_vp)]


# name -> (restype, argtypes); every symbol declared in include/pagraph_b200.h
SIGNATURES = {
    "pg_version": (ctypes.c_int, []),
    "pg_last_error": (ctypes.c_char_p, []),
    "pg_device_info": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "pg_host_alloc": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "pg_host_free": (ctypes.c_int, [c_vp]),
    "pg_host_register": (ctypes.c_int, [c_vp, ctypes.c_size_t]),
    "pg_host_unregister": (ctypes.c_int, [c_vp]),
    "pg_graph_create": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                       ctypes.POINTER(c_vp)]),
    "pg_graph_create_device": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                              ctypes.POINTER(c_vp)]),
    "pg_graph_destroy": (None, [c_vp]),
    "pg_graph_degrees": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp]),
    "pg_sampler_create": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64p, ctypes.c_uint64, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "pg_sampler_destroy": (None, [c_vp]),
    "pg_sample": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                 ctypes.POINTER(pg_nodeflow_buffers), c_vp, c_vp]),
    "pg_cache_create": (ctypes.c_int, [ctypes.c_int64, c_vp, c_vp, c_vp, ctypes.c_int,
                                       ctypes.POINTER(pg_field), ctypes.c_int, ctypes.POINTER(c_vp)]),
    "pg_cache_destroy": (None, [c_vp]),
    "pg_cache_fill": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(c_vp),
                                     ctypes.c_int, c_vp]),
    "pg_cache_fetch_host": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.POINTER(c_vp), c_vp]),
    "pg_cache_fetch": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.POINTER(c_vp), c_vp, c_vp,
                                      ctypes.c_int, c_vp]),
    "pg_cache_resolve": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(pg_block), c_vp, c_vp, ctypes.c_int64, c_vp,
                                        c_vp, c_vp]),
    "pg_cache_fill_rows": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.POINTER(c_vp), c_vp]),
    "pg_peer_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_vp), c_vp]),
    "pg_peer_open": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "pg_peer_close": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "pg_peer_free": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "pg_cache_set_peers": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp), c_vp,
                                          ctypes.c_int64, ctypes.c_int64, c_vp]),
    "pg_cache_set_hot": (ctypes.c_int, [c_vp, c_vp]),
    "pg_set_agg_reserve_sms": (None, [ctypes.c_int]),
    "pg_aggregate_rows": (ctypes.c_int, [c_vp, ctypes.POINTER(pg_block), ctypes.c_int32, c_vp, ctypes.c_int64, ctypes.c_int,
                                         c_vp, ctypes.c_float, ctypes.c_uint64, c_vp, ctypes.c_int64, c_vp]),
    "pg_cache_fetch_dyn": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.POINTER(c_vp), c_vp,
                                          ctypes.c_int, c_vp, c_vp]),
    "pg_minibatch_key": (None, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint32)]),
    "pg_sample_keyed": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.POINTER(pg_nodeflow_buffers), c_vp,
                                       c_vp, c_vp, c_vp]),
    "pg_aggregate_fwd_dyn": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int32, ctypes.c_int, c_vp, c_vp]),
    "pg_aggregate_bwd_dyn": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int64, ctypes.c_int32, ctypes.c_int, c_vp, c_vp]),
    "pg_cache_aggregate": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(pg_block), c_vp, ctypes.c_int64,
                                          ctypes.c_int, c_vp, ctypes.c_float, ctypes.c_uint64, c_vp, ctypes.c_int64,
                                          c_vp, c_vp]),
    "pg_aggregate_fwd": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int32, ctypes.c_int, c_vp, c_vp]),
    "pg_aggregate_bwd": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, c_vp, c_vp]),
    "pg_linear_concat_fwd": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, c_vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_int, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_float,
                                            ctypes.c_uint64, c_vp, c_vp]),
    "pg_linear_concat_bwd": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int32, ctypes.c_int32, ctypes.c_int, ctypes.c_float, ctypes.c_uint64,
                                            c_vp, c_vp, c_vp, c_vp]),
    "pg_linear_cross_entropy": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int32,
                                               ctypes.c_int32, c_vp, c_vp, ctypes.c_int64, c_vp, c_vp, c_vp, c_vp]),
    "pg_block_linear_cross_entropy": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                     ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_int32, ctypes.c_int32, c_vp,
                                                     c_vp, ctypes.c_int64, c_vp, c_vp, c_vp]),
    "pg_peer_group_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(c_vp),
                                            c_vp]),
    "pg_peer_group_connect": (ctypes.c_int, [c_vp, c_vp]),
    "pg_peer_group_destroy": (None, [c_vp]),
    "pg_allreduce_adam": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp]),
    "pg_allreduce_adam_next": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_float, ctypes.c_float,
                                              ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp]),
    "pg_partition_dg": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       c_vp, c_vp]),
    "pg_measure_h2d": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_double)]),
    "pg_timing_enable": (ctypes.c_int, [ctypes.c_int]),
    "pg_timing_drain": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float),
                                       ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "pg_timing_drain_timeline": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float),
                                                ctypes.POINTER(ctypes.c_float), ctypes.c_int64,
                                                ctypes.POINTER(ctypes.c_int64)]),
    "pg_launch_count": (ctypes.c_int64, []),
    "pg_dropout_keep_mask": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, c_vp]),
}

_LIB = None


def lib_path():
    return _build.LIB_PATH


def lib():
    """Load the CUDA library (building it on first use if nvcc is available). Raises if absent."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            try:
                _build.build()
            except Exception as e:  # no nvcc on this box and no prebuilt library: hard error
                raise PGError("libpagraph_b200.so is missing and could not be built (%s); "
                              "run `python -m pagraph_b200.build`" % e)
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status, what=""):
    if status != PG_OK:
        msg = lib().pg_last_error()
        raise PGError("%s failed (status %d): %s" % (what or "pagraph_b200 call", status,
                                                     msg.decode() if msg else ""))


def ptr(t):
    """Device/host address of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


T_SAMPLE, T_SPLIT, T_GATHER_HIT, T_GATHER_MISS, T_AGG_FWD, T_AGG_BWD, T_FUSED, T_DENSE_FWD, T_DENSE_BWD, T_HEAD, T_OPT = range(11)


def timing_enable(on):
    check(lib().pg_timing_enable(int(on)), "pg_timing_enable")


def timing_drain(cap=1 << 16):
    """[(slot, ms)] of every timed launch since the previous drain, in launch order. Synchronises."""
    slots = (ctypes.c_int32 * cap)()
    ms = (ctypes.c_float * cap)()
    n = ctypes.c_int64()
    check(lib().pg_timing_drain(slots, ms, cap, ctypes.byref(n)), "pg_timing_drain")
    return [(slots[i], ms[i]) for i in range(n.value)]


def timing_drain_timeline(cap=1 << 16):
    """[(slot, begin_ms, end_ms)] of every timed launch since the previous drain, relative to the first. Synchronises."""
    slots = (ctypes.c_int32 * cap)()
    t0 = (ctypes.c_float * cap)()
    t1 = (ctypes.c_float * cap)()
    n = ctypes.c_int64()
    check(lib().pg_timing_drain_timeline(slots, t0, t1, cap, ctypes.byref(n)), "pg_timing_drain_timeline")
    return [(slots[i], t0[i], t1[i]) for i in range(n.value)]


def launch_count():
    return int(lib().pg_launch_count())
