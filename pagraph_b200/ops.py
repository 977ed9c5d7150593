"""Autograd wrapper of the CUDA block aggregation (pg_aggregate_fwd / pg_aggregate_bwd)."""
import torch

from . import _lib

_MODES = {"sum": _lib.PG_AGG_SUM, "mean": _lib.PG_AGG_MEAN}


def aggregate_forward(indptr, cols, col_base, src, n_dst, mode, norm=None, out=None):
    """dst[r] = reduce_{e in row r} src[cols[e] - col_base]; fp32 CUDA only.

    indptr: int64 CUDA tensor with >= n_dst + 1 entries of absolute offsets into `cols`."""
    if not (src.is_cuda and src.dtype == torch.float32):
        raise _lib.PGError("aggregate: expected a float32 CUDA tensor (no CPU path)")
    if src.stride(-1) != 1:
        src = src.contiguous()
    dim = src.shape[1]
    if out is None:
        out = torch.empty((n_dst, dim), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().pg_aggregate_fwd(_lib.ptr(indptr), _lib.ptr(cols), col_base, _lib.ptr(src),
                                               src.stride(0), _lib.ptr(out), out.stride(0), n_dst, dim,
                                               _MODES[mode], _lib.ptr(norm), _lib.stream_ptr()),
                   "pg_aggregate_fwd")
    return out


def aggregate_backward(indptr, cols, col_base, grad_dst, n_src, mode, norm=None):
    if grad_dst.stride(-1) != 1:
        grad_dst = grad_dst.contiguous()
    n_dst, dim = grad_dst.shape
    grad_src = torch.empty((n_src, dim), dtype=torch.float32, device=grad_dst.device)
    with torch.cuda.device(grad_dst.device):
        _lib.check(_lib.lib().pg_aggregate_bwd(_lib.ptr(indptr), _lib.ptr(cols), col_base, _lib.ptr(grad_dst),
                                               grad_dst.stride(0), _lib.ptr(grad_src), grad_src.stride(0), n_dst,
                                               n_src, dim, _MODES[mode], _lib.ptr(norm), _lib.stream_ptr()),
                   "pg_aggregate_bwd")
    return grad_src


class BlockAggregate(torch.autograd.Function):
    """copy_src + sum/mean over one NodeFlow block."""

    @staticmethod
    def forward(ctx, src, indptr, cols, col_base, n_dst, mode):
        ctx.block = (indptr, cols, col_base, mode, src.shape[0])
        return aggregate_forward(indptr, cols, col_base, src, n_dst, mode)

    @staticmethod
    def backward(ctx, grad_out):
        indptr, cols, col_base, mode, n_src = ctx.block
        grad_src = None
        if ctx.needs_input_grad[0]:
            grad_src = aggregate_backward(indptr, cols, col_base, grad_out, n_src, mode)
        return grad_src, None, None, None, None, None


def cache_aggregate(cacher, field, parent_ids, indptr, cols, col_base, n_src, n_dst, mode, norm=None, dropout_p=0.0,
                    seed=0, step=None, out=None, zero_rows_to=0):
    """Fused cache lookup + dropout + block aggregation (pg_cache_aggregate): dst rows are reduced straight
    from the feature cache / host table of `cacher` (a GraphCacheServer), without materialising the
    gathered source rows. `field` is the field name; `step` an optional int64 CUDA scalar added to the seed."""
    import ctypes
    fi = cacher._field_names.index(field)
    dim = cacher.dims[field]
    dev = cacher._dev
    if out is None:
        out = torch.empty((max(n_dst, zero_rows_to), dim), dtype=torch.float32, device=dev)
    blk = _lib.pg_block(_lib.ptr(parent_ids), _lib.ptr(indptr), _lib.ptr(cols), col_base, n_src, n_dst)
    counts = cacher._counts if (cacher.log and not cacher.full_cached) else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pg_cache_aggregate(cacher._handle, fi, ctypes.byref(blk), _lib.ptr(out), out.stride(0),
                                                 _MODES[mode], _lib.ptr(norm), float(dropout_p), int(seed) & (2 ** 64 - 1),
                                                 _lib.ptr(step), zero_rows_to, _lib.ptr(counts), _lib.stream_ptr()),
                   "pg_cache_aggregate")
    return out


def linear_concat_forward(x, weight, bias, concat, out=None, out_drop=None, dropout_p=0.0, seed=0, step=None):
    """out = cat(z, relu(z)) | relu(z), z = x W^T + b (pg_linear_concat_fwd, 3xTF32 tensor-core product); with
    dropout_p > 0 also out_drop = dropout(out) under the hash-mask contract. Returns (out, out_drop | None)."""
    n, K = x.shape
    width = 64 if concat else 32
    if out is None:
        out = torch.empty((n, width), dtype=torch.float32, device=x.device)
    if dropout_p > 0 and out_drop is None:
        out_drop = torch.empty((n, width), dtype=torch.float32, device=x.device)
    if dropout_p <= 0:
        out_drop = None
    w = weight.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pg_linear_concat_fwd(_lib.ptr(x), x.stride(0), _lib.ptr(w), _lib.ptr(bias), n, K, 32, int(concat),
                                                   _lib.ptr(out), out.stride(0), _lib.ptr(out_drop),
                                                   out_drop.stride(0) if out_drop is not None else 0, float(dropout_p),
                                                   int(seed) & (2 ** 64 - 1), _lib.ptr(step), _lib.stream_ptr()),
                   "pg_linear_concat_fwd")
    return out, out_drop


def linear_concat_backward(x, grad_out, out, concat, gw, gb, dropout_p=0.0, seed=0, step=None):
    """dW [32, K] -> gw, db [32] -> gb (both overwritten) of linear_concat_forward; grad_out is the gradient of out_drop
    when dropout_p > 0 (the mask is regenerated from (seed, step)), else of out."""
    n, K = x.shape
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pg_linear_concat_bwd(_lib.ptr(x), x.stride(0), _lib.ptr(grad_out), grad_out.stride(0),
                                                   _lib.ptr(out), out.stride(0), n, K, 32, int(concat), float(dropout_p),
                                                   int(seed) & (2 ** 64 - 1), _lib.ptr(step), _lib.ptr(gw), _lib.ptr(gb),
                                                   _lib.stream_ptr()), "pg_linear_concat_bwd")


class LinearConcat(torch.autograd.Function):
    """cat(z, relu(z)) (concat=True) or relu(z), z = x W^T + b, for an input x that needs no gradient (the aggregated
    input block), optionally followed by dropout (hash mask keyed by (seed, *step, row, column); `step` an int64 CUDA
    scalar that must not change between forward and backward). Same math as NodeUpdate.forward
    (PaGraph/model/gcn_nssc.py:14-24) + the dropout of :66-67. Forward and backward are the tensor-core kernels of
    pg_dense_mma.cu (3xTF32: fp32-level accuracy); bias, relu, the concat split, dropout, dW and db are fused into them."""

    @staticmethod
    def supported(x, weight):
        return (x.is_cuda and x.dtype == torch.float32 and not x.requires_grad and weight.shape[0] == 32
                and x.shape[1] % 4 == 0 and x.shape[1] <= 768 and x.stride(1) == 1 and x.stride(0) % 4 == 0
                and x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0)

    @staticmethod
    def forward(ctx, x, weight, bias, concat, dropout_p=0.0, seed=0, step=None):
        out, out_drop = linear_concat_forward(x, weight, bias, concat, dropout_p=dropout_p, seed=seed, step=step)
        ctx.save_for_backward(x, out)
        ctx.concat, ctx.has_bias, ctx.drop = concat, bias is not None, (float(dropout_p), seed, step)
        return out_drop if out_drop is not None else out

    @staticmethod
    def backward(ctx, grad_out):
        x, out = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gw = torch.empty((32, x.shape[1]), dtype=torch.float32, device=x.device)
        gb = torch.empty(32, dtype=torch.float32, device=x.device)
        p, seed, step = ctx.drop
        linear_concat_backward(x, grad_out, out, ctx.concat, gw, gb, p, seed, step)
        return None, gw, (gb if ctx.has_bias else None), None, None, None, None


class LinearCrossEntropy(torch.autograd.Function):
    """mean cross-entropy of (x W^T + b) against integer labels — the last NodeUpdate (no activation) followed by
    torch.nn.CrossEntropyLoss — forward and backward in one kernel (pg_linear_cross_entropy)."""

    @staticmethod
    def supported(x, weight):
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] <= 64 and weight.shape[0] <= 64
                and x.stride(1) == 1)

    @staticmethod
    def forward(ctx, x, weight, bias, labels):
        n, K = x.shape
        C = weight.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        gx = torch.empty((n, K), dtype=torch.float32, device=x.device)
        gw = torch.empty((C, K), dtype=torch.float32, device=x.device)
        gb = torch.empty(C, dtype=torch.float32, device=x.device)
        w = weight.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().pg_linear_cross_entropy(_lib.ptr(x), x.stride(0), _lib.ptr(w), _lib.ptr(bias),
                                                          _lib.ptr(labels), n, K, C, _lib.ptr(loss), _lib.ptr(gx),
                                                          gx.stride(0), _lib.ptr(gw), _lib.ptr(gb), None, _lib.stream_ptr()),
                       "pg_linear_cross_entropy")
        ctx.save_for_backward(gx, gw, gb)
        ctx.has_bias = bias is not None
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gx, gw, gb = ctx.saved_tensors
        return gx * g, gw * g, (gb * g if ctx.has_bias else None), None
