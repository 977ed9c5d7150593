"""GPU block aggregation (pg_aggregate_fwd/bwd through the C-ABI) vs the float64 oracle.
Tolerance (BASELINE.json north_star): 1e-5 relative, fp32."""
import numpy as np
import pytest

import oracle
from conftest import random_in_csr

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(got, want, scale_ref=None):
    # relative to the magnitude of the row sums (cancellation makes per-element rtol meaningless)
    denom = np.maximum(np.abs(want), 1e-30) if scale_ref is None else scale_ref
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / denom
    assert err.max() <= RTOL, err.max()


def _nodeflow(V=4000, nnz=80000, seeds=600, fanouts=(25, 10), hub=False):
    indptr, indices, eids, _ = random_in_csr(V, nnz, seed=6, hub=hub)
    s = np.random.default_rng(0).choice(V, seeds, replace=False)
    return oracle.sample(indptr, indices, eids, s, list(fanouts), seed=3)


@pytest.mark.parametrize("dim", [600, 64, 32, 128, 16, 4, 13, 602, 1, 2400])
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_forward_matches_oracle(dim, mode):
    import torch
    from pagraph_b200 import ops
    nf = _nodeflow()
    rng = np.random.default_rng(dim)
    for i in range(nf.num_blocks):
        ip, cols, base = nf.block(i)
        n_src = len(nf.layer_parent_nid(i))
        x = rng.random((n_src, dim), dtype=np.float32)          # U[0,1) like the reference features
        want = oracle.aggregate(ip, cols, base, x, mode)
        got = ops.aggregate_forward(torch.from_numpy(ip).cuda(), torch.from_numpy(cols).cuda(), base,
                                    torch.from_numpy(x).cuda(), len(ip) - 1, mode)
        _close(got.cpu().numpy(), want)
        # signed inputs: error relative to sum |x| over the row (the fp32 accumulation bound)
        xs = rng.standard_normal((n_src, dim)).astype(np.float32)
        want = oracle.aggregate(ip, cols, base, xs, mode)
        bound = oracle.aggregate(ip, cols, base, np.abs(xs), mode)
        got = ops.aggregate_forward(torch.from_numpy(ip).cuda(), torch.from_numpy(cols).cuda(), base,
                                    torch.from_numpy(xs).cuda(), len(ip) - 1, mode)
        _close(got.cpu().numpy(), want, np.maximum(bound.astype(np.float64), 1e-30))


@pytest.mark.parametrize("dim", [600, 64, 13])
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_backward_matches_oracle(dim, mode):
    import torch
    from pagraph_b200 import ops
    nf = _nodeflow(hub=True, V=1500, nnz=60000, seeds=300, fanouts=(10, 5))
    rng = np.random.default_rng(dim + 1)
    for i in range(nf.num_blocks):
        ip, cols, base = nf.block(i)
        n_src, n_dst = len(nf.layer_parent_nid(i)), len(ip) - 1
        gd = rng.random((n_dst, dim), dtype=np.float32)
        want = oracle.aggregate_bwd(ip, cols, base, gd, n_src, mode)
        got = ops.aggregate_backward(torch.from_numpy(ip).cuda(), torch.from_numpy(cols).cuda(), base,
                                     torch.from_numpy(gd).cuda(), n_src, mode)
        assert got.shape == (n_src, dim)
        _close(got.cpu().numpy(), want, np.maximum(np.abs(want.astype(np.float64)), 1e-30))


def test_zero_degree_rows_norm_and_strides():
    import ctypes
    import torch
    from pagraph_b200 import _lib
    indptr = torch.tensor([0, 0, 2, 2, 5], dtype=torch.int64).cuda()
    cols = torch.tensor([10, 11, 12, 10, 10], dtype=torch.int64).cuda()
    src_full = torch.arange(3 * 12, dtype=torch.float32).reshape(3, 12).cuda()
    src = src_full[:, :8]                                             # stride 12, dim 8
    dst = torch.full((4, 16), -1.0, device="cuda")
    norm = torch.tensor([2.0, 0.5, 3.0, float("inf")], device="cuda")
    L = _lib.lib()
    _lib.check(L.pg_aggregate_fwd(_lib.ptr(indptr), _lib.ptr(cols), 10, _lib.ptr(src), 12, _lib.ptr(dst), 16, 4, 8,
                                  _lib.PG_AGG_SUM, _lib.ptr(norm), None), "fwd")
    torch.cuda.synchronize()
    s = src.cpu().numpy()
    want = np.stack([np.zeros(8), (s[0] + s[1]) * 0.5, np.zeros(8), (s[2] + s[0] + s[0]) * np.inf])
    np.testing.assert_array_equal(dst[:, :8].cpu().numpy(), want.astype(np.float32))
    assert (dst[:, 8:] == -1).all()                                   # padding untouched
    _lib.check(L.pg_aggregate_fwd(_lib.ptr(indptr), _lib.ptr(cols), 10, _lib.ptr(src), 12, _lib.ptr(dst), 16, 4, 8,
                                  _lib.PG_AGG_MEAN, None, None), "fwd")
    want = np.stack([np.zeros(8), (s[0] + s[1]) / 2, np.zeros(8), (s[2] + s[0] + s[0]) / 3])
    np.testing.assert_allclose(dst[:, :8].cpu().numpy(), want, rtol=1e-6)
    assert L.pg_aggregate_fwd(_lib.ptr(indptr), _lib.ptr(cols), 10, _lib.ptr(src), 4, _lib.ptr(dst), 16, 4, 8,
                              _lib.PG_AGG_SUM, None, None) == _lib.PG_ERR_INVALID
    assert L.pg_aggregate_fwd(_lib.ptr(indptr), _lib.ptr(cols), 10, _lib.ptr(src), 12, _lib.ptr(dst), 16, 4, 8,
                              7, None, None) == _lib.PG_ERR_INVALID
    assert ctypes.string_at(L.pg_last_error()) != b""


def test_autograd_function_matches_torch_sparse():
    import torch
    from pagraph_b200 import ops
    nf = _nodeflow(V=1000, nnz=20000, seeds=100, fanouts=(6, 4))
    ip, cols, base = nf.block(1)
    n_src, n_dst = len(nf.layer_parent_nid(1)), len(ip) - 1
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((n_src, 64), dtype=np.float32)).cuda().requires_grad_(True)
    A = torch.sparse_csr_tensor(torch.from_numpy(ip - ip[0]), torch.from_numpy(cols[ip[0]:ip[-1]] - base),
                                torch.ones(int(ip[-1] - ip[0]), dtype=torch.float64), size=(n_dst, n_src)).cuda()
    deg = torch.from_numpy(np.maximum(np.diff(ip), 1)).cuda()[:, None]
    w = torch.from_numpy(rng.random((n_dst, 64), dtype=np.float32)).cuda()
    out = ops.BlockAggregate.apply(x, torch.from_numpy(ip).cuda(), torch.from_numpy(cols).cuda(), base, n_dst, "mean")
    (out * w).sum().backward()
    x64 = x.detach().double().requires_grad_(True)
    ref = (A @ x64) / deg
    (ref * w.double()).sum().backward()
    torch.testing.assert_close(out.double(), ref, rtol=RTOL, atol=0)
    torch.testing.assert_close(x.grad.double(), x64.grad, rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("n,K", [(1000, 600), (37, 128), (4099, 64), (1, 600), (0, 600), (300, 768), (513, 4), (2500, 260)])
@pytest.mark.parametrize("concat", [True, False])
def test_linear_concat_matches_torch(n, K, concat):
    """Fused first NodeUpdate (linear + bias + cat(z, relu(z))) forward and its dW / db against torch autograd (float64).
    The kernels are 3xTF32 tensor-core products: the tolerance is fp32-level, not TF32-level."""
    import torch
    from pagraph_b200.ops import LinearConcat
    torch.manual_seed(n + K)
    x = torch.randn(n, K, device="cuda")
    lin = torch.nn.Linear(K, 32).cuda()
    gout = torch.randn(n, 64 if concat else 32, device="cuda")
    assert LinearConcat.supported(x, lin.weight)
    out = LinearConcat.apply(x, lin.weight, lin.bias, concat)
    out.backward(gout)
    torch.cuda.synchronize()
    gw, gb = lin.weight.grad.clone(), lin.bias.grad.clone()
    lin.zero_grad()
    z = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
    ref = torch.cat((z, torch.relu(z)), 1) if concat else torch.relu(z)
    torch.testing.assert_close(out.double(), ref, rtol=1e-5, atol=1e-5)
    if n:                                   # fp32-level: far below one TF32 ulp (2^-11) of the operands
        assert (out.double() - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max().item())
    w64 = lin.weight.detach().double().requires_grad_(True)
    b64 = lin.bias.detach().double().requires_grad_(True)
    z = torch.nn.functional.linear(x.double(), w64, b64)
    # relu' is taken from the fp32 forward (z > 0 as the kernel saw it), like autograd does with the saved output
    pos = (out[:, 32:] > 0) if concat else (out > 0)
    gz = (gout[:, :32].double() + gout[:, 32:].double() * pos) if concat else gout.double() * pos
    z.backward(gz)
    torch.testing.assert_close(gw.double(), w64.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gb.double(), b64.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("n,K", [(40000, 600), (80003, 64)])
def test_linear_concat_large_row_counts(n, K):
    """More rows than one pass of the persistent kernels covers: the forward loops over row tiles (n > 148 x 256), the dW
    kernel walks several 256-row gz super-tiles per CTA with its x register ring running across them. Norm-wise
    tolerance: the tensor-core accumulators truncate, so the error grows with the number of MMAs per accumulator."""
    import torch
    from pagraph_b200.ops import LinearConcat
    torch.manual_seed(7)
    x = torch.randn(n, K, device="cuda")
    lin = torch.nn.Linear(K, 32).cuda()
    gout = torch.randn(n, 64, device="cuda")
    out = LinearConcat.apply(x, lin.weight, lin.bias, True)
    out.backward(gout)
    z = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
    ref = torch.cat((z, torch.relu(z)), 1)
    assert (out.double() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    pos = out[:, 32:] > 0
    gz = gout[:, :32].double() + gout[:, 32:].double() * pos
    gw_ref, gb_ref = gz.t() @ x.double(), gz.sum(0)
    assert (lin.weight.grad.double() - gw_ref).abs().max().item() < 3e-5 * gw_ref.abs().max().item()
    assert (lin.bias.grad.double() - gb_ref).abs().max().item() < 1e-5 * max(gb_ref.abs().max().item(), n ** 0.5)


@pytest.mark.parametrize("n,K,p", [(1000, 600, 0.2), (77, 64, 0.5), (4100, 600, 0.2)])
@pytest.mark.parametrize("concat", [True, False])
def test_linear_concat_fused_dropout(n, K, p, concat):
    """Dropout folded into the NodeUpdate kernels: out_drop = out * keep / (1 - p) with the oracle's hash mask (keyed by
    seed + device step), and the backward regenerates the same mask."""
    import torch
    import oracle
    from pagraph_b200.ops import LinearConcat, linear_concat_forward
    torch.manual_seed(n + K)
    width = 64 if concat else 32
    x = torch.randn(n, K, device="cuda")
    lin = torch.nn.Linear(K, 32).cuda()
    gout = torch.randn(n, width, device="cuda")
    seed = 0x1234567890ABCDEF
    step = torch.tensor([7], dtype=torch.int64, device="cuda")
    out, _ = linear_concat_forward(x, lin.weight, lin.bias, concat)
    od = LinearConcat.apply(x, lin.weight, lin.bias, concat, p, seed - 7, step)
    keep = torch.from_numpy(oracle.dropout_keep_mask(seed, n, width, p)).cuda()
    assert abs(keep.float().mean().item() - (1 - p)) < 0.02
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    want = torch.where(keep, out * float(scale), torch.zeros_like(out))
    torch.testing.assert_close(od, want, rtol=0, atol=0)
    od.backward(gout)
    gw, gb = lin.weight.grad.clone(), lin.bias.grad.clone()
    g = gout.double() * keep * float(scale)
    pos = (out[:, 32:] > 0) if concat else (out > 0)
    gz = (g[:, :32] + g[:, 32:] * pos) if concat else g * pos
    torch.testing.assert_close(gw.double(), gz.t() @ x.double(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gb.double(), gz.sum(0), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("n,K,C", [(6000, 64, 60), (77, 64, 41), (1, 32, 7), (513, 50, 64), (0, 64, 60)])
def test_linear_cross_entropy_matches_torch(n, K, C):
    """Classifier head + CrossEntropyLoss in one kernel: loss and all three gradients against torch (float64)."""
    import torch
    from pagraph_b200.ops import LinearCrossEntropy
    torch.manual_seed(n + K + C)
    x = torch.randn(n, K, device="cuda", requires_grad=True)
    lin = torch.nn.Linear(K, C).cuda()
    y = torch.randint(0, C, (n,), device="cuda")
    assert LinearCrossEntropy.supported(x, lin.weight)
    loss = LinearCrossEntropy.apply(x, lin.weight, lin.bias, y)
    (loss * 3.0).backward()
    got = [t.clone() for t in (x.grad, lin.weight.grad, lin.bias.grad)]
    if n == 0:
        assert all((t == 0).all() for t in got)
        return
    x64 = x.detach().double().requires_grad_(True)
    w64 = lin.weight.detach().double().requires_grad_(True)
    b64 = lin.bias.detach().double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(torch.nn.functional.linear(x64, w64, b64), y)
    (ref * 3.0).backward()
    torch.testing.assert_close(loss.double(), ref, rtol=1e-5, atol=1e-6)
    for a, b in zip(got, (x64.grad, w64.grad, b64.grad)):
        torch.testing.assert_close(a.double(), b, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["mean", "sum"])
@pytest.mark.parametrize("n_dst,n_src,K,C", [(6000, 35000, 64, 60), (130, 400, 64, 41), (64, 64, 32, 7), (1, 5, 64, 60)])
def test_block_linear_cross_entropy_matches_torch(n_dst, n_src, K, C, mode):
    """pg_block_linear_cross_entropy (block + head + loss, forward and backward, one kernel) through the raw C-ABI against
    float64 torch: a = reduce(block, src); CE(a W^T + b); gradients wrt src, W, b. The block sits inside a larger
    NodeFlow (non-zero layer offsets, read from the device), with zero-degree rows and row capacities above the sizes."""
    import ctypes

    import torch
    from pagraph_b200 import _lib
    rng = np.random.default_rng(n_dst + n_src + K + C)
    pre = 37                                                     # a layer in front of the source layer
    deg = rng.integers(0, 11, n_dst)
    if n_dst > 1:
        deg[rng.integers(0, n_dst, max(1, n_dst // 10))] = 0
    else:
        deg[:] = 3
    indptr_blk = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    cols_blk = rng.integers(0, n_src, int(indptr_blk[-1])).astype(np.int64) + pre      # NodeFlow-wide source ids
    # NodeFlow-wide indptr: rows [0, pre + n_src) have no in-edges here, rows of the destination layer follow
    indptr = np.concatenate([np.zeros(pre + n_src, np.int64), indptr_blk])
    lo = torch.tensor([pre, pre + n_src, pre + n_src + n_dst], dtype=torch.int64, device="cuda")
    cap_dst, cap_src = n_dst + 50, n_src + 100
    src = torch.randn(cap_src, K, device="cuda")
    lin = torch.nn.Linear(K, C).cuda()
    y = torch.randint(0, C, (cap_dst,), device="cuda")
    d_indptr, d_cols = torch.from_numpy(indptr).cuda(), torch.from_numpy(cols_blk).cuda()
    loss = torch.full((1,), 7.0, device="cuda")
    gsrc = torch.full((cap_src, K), 3.0, device="cuda")
    gw, gb = torch.full((C, K), 5.0, device="cuda"), torch.full((C,), 5.0, device="cuda")
    L = _lib.lib()
    w = lin.weight.detach().contiguous()
    _lib.check(L.pg_block_linear_cross_entropy(_lib.ptr(d_indptr), _lib.ptr(d_cols), _lib.ptr(lo), _lib.ptr(src), K, cap_dst,
                                               cap_src, 1 if mode == "mean" else 0, _lib.ptr(w), _lib.ptr(lin.bias.detach()),
                                               _lib.ptr(y), K, C, _lib.ptr(loss), _lib.ptr(gsrc), K, _lib.ptr(gw), _lib.ptr(gb),
                                               _lib.stream_ptr()), "pg_block_linear_cross_entropy")
    torch.cuda.synchronize()
    src64 = src.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    rows = torch.from_numpy(np.repeat(np.arange(n_dst), deg)).cuda()
    a = torch.zeros(n_dst, K, dtype=torch.float64, device="cuda").index_add(0, rows, src64[d_cols - pre])
    if mode == "mean":
        a = a / torch.from_numpy(np.maximum(deg, 1)).cuda().double()[:, None]
    ref = torch.nn.functional.cross_entropy(torch.nn.functional.linear(a, w64, b64), y[:n_dst])
    ref.backward()
    torch.testing.assert_close(loss[0].double(), ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gsrc.double(), src64.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gw.double(), w64.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(gb.double(), b64.grad, rtol=1e-4, atol=1e-6)
