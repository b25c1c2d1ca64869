"""Does programmatic dependent launch help inside a torch-captured CUDA graph? Chain of dependent kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import _lib, ops  # noqa: E402

a = torch.randn(1500, 768, device="cuda")
b = torch.empty_like(a)
w = torch.randn(768, 768, device="cuda").half()
xb = torch.randn(1500, 768, device="cuda").half()
o = torch.empty(1500, 768, device="cuda")


def chain_ew(n=200):
    for i in range(n // 2):
        ops.ew(a, out_f32=b)
        ops.ew(b, out_f32=a)


def chain_gemm(n=100):
    for i in range(n):
        ops.gemm(xb, w, out_f32=o, out_ld=768)


for name, fn, n in (("ew x200", chain_ew, 200), ("gemm x100", chain_gemm, 100)):
    for pdl in (0, 1):
        _lib.lib().vb_set_pdl(pdl)
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} pdl={pdl}: {e0.elapsed_time(e1) / 10 / n * 1e3:.2f} us per kernel")
