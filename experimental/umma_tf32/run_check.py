#!/usr/bin/env python
"""Build and check experimental/umma_tf32/linear_fwd_umma.cu on a B200 (run under `timeout`). Not part of the test-suite."""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import torch  # noqa: E402


def build():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpgx_umma.so")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                           "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", so,
                           os.path.join(HERE, "linear_fwd_umma.cu")])
    return ctypes.CDLL(so)


def main():
    L = build()
    L.pgx_linear_fwd_umma.restype = ctypes.c_int
    L.pgx_linear_fwd_umma.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int32, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                      ctypes.c_void_p]
    torch.manual_seed(0)
    for n, K in [(128, 32), (128, 64), (300, 600), (36864, 600), (40000, 128)]:
        x = torch.randn(n, K, device="cuda")
        lin = torch.nn.Linear(K, 32).cuda()
        out = torch.zeros(n, 64, device="cuda")
        ws = torch.zeros(2 * 32 * ((K + 31) // 32 * 32), device="cuda")
        st = L.pgx_linear_fwd_umma(x.data_ptr(), x.stride(0), lin.weight.data_ptr(), lin.bias.data_ptr(), n, K, 1,
                                   out.data_ptr(), out.stride(0), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        z = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
        ref = torch.cat((z, torch.relu(z)), 1)
        err = (out.double() - ref).abs().max().item()
        print("n=%d K=%d status=%d max_err=%.3e %s" % (n, K, st, err, "ok" if st == 0 and err < 2e-5 * max(1, ref.abs().max().item()) else "MISMATCH"))
    # timing at config-2 shape against the product kernel
    from pagraph_b200.ops import linear_concat_forward
    n, K = 36864, 600
    xs = [torch.randn(n, K, device="cuda") for _ in range(3)]
    lin = torch.nn.Linear(K, 32).cuda()
    out = torch.zeros(n, 64, device="cuda")
    ws = torch.zeros(2 * 32 * 608, device="cuda")

    def t(fn):
        for i in range(3):
            fn(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(30):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 30 * 1e3
    s = torch.cuda.current_stream().cuda_stream
    print("umma    %.1f us" % t(lambda i: L.pgx_linear_fwd_umma(xs[i % 3].data_ptr(), K, lin.weight.data_ptr(), lin.bias.data_ptr(), n, K, 1,
                                                                out.data_ptr(), 64, ws.data_ptr(), s)))
    print("mma.sync %.1f us" % t(lambda i: linear_concat_forward(xs[i % 3], lin.weight.detach(), lin.bias.detach(), True, out=out)))


if __name__ == "__main__":
    main()
