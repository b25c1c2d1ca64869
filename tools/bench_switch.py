"""Cost of switching between different kernels inside a graph (instruction-cache / launch effects)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import ops  # noqa: E402

REP = 24


def timeit(fn, iters=10):
    for i in range(REP):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / REP * 1e3


dt = torch.float32
M, K = 1500, 768
a = torch.randn(M, K, device="cuda").to(dt)
w1 = torch.randn(2304, K, device="cuda").to(dt)
w2 = torch.randn(768, K, device="cuda").to(dt)
o1 = torch.empty(M, 2304, device="cuda")
o2 = torch.empty(M, 768, device="cuda")
A = lambda bn=0: ops.gemm(a, w1, out_f32=o1, out_ld=2304, bn=bn)
Bq = lambda bn=0: ops.gemm(a, w2, out_f32=o2, out_ld=768, bn=bn)
tA, tB = timeit(lambda i: A()), timeit(lambda i: Bq())
tAB = timeit(lambda i: (A(), Bq()))
print(f"auto tiles : A {tA:.1f}  B {tB:.1f}  A+B alternating {tAB:.1f} (sum {tA + tB:.1f}, switch cost {(tAB - tA - tB) / 2:.1f} us per launch)")
tA, tB = timeit(lambda i: A(128)), timeit(lambda i: Bq(128))
tAB = timeit(lambda i: (A(128), Bq(128)))
print(f"same kernel (bn128 both): A {tA:.1f}  B {tB:.1f}  alternating {tAB:.1f} (sum {tA + tB:.1f}, switch cost {(tAB - tA - tB) / 2:.1f})")
