"""Micro-benchmark of vb_gemm on the hot shapes, per N-tile. Launches are captured into a CUDA
graph (20 per replay) so the numbers are device time, not Python/ctypes overhead."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import ops  # noqa: E402

REP = 20


def timeit(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / REP * 1e3  # us per launch


def plain(M, N, K, bns=(32, 64, 128)):
    a = torch.randn(M, K, device="cuda").half()
    b = torch.randn(N, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda")
    res = []
    for bn in bns:
        us = timeit(lambda: ops.gemm(a, b, out_f32=out, out_ld=N, bn=bn))
        res.append(f"bn{bn}: {us:6.1f}us {2 * M * N * K / us / 1e6:7.1f}TF")
    print(f"plain M={M} N={N} K={K}: " + "  ".join(res), flush=True)


def conv(T, Cin, Cout, k, bns=(32, 64, 128), B=1):
    x = torch.randn(B, T, Cin, device="cuda").half()
    kpt = (Cin + 63) // 64
    w = torch.randn(Cout, k * kpt * 64, device="cuda").half()
    out = torch.empty(B, T, Cout, device="cuda", dtype=torch.float16)
    res = []
    for bn in bns:
        if bn > max(Cout, 32):
            continue
        us = timeit(lambda: ops.gemm(x, w, M=T, G1=B, kpt=kpt, dil=1, pad=k // 2, a_c2=(1, 0), out_fp16=out,
                                     outb_ld=Cout, outb_g=(T * Cout, 0), bn=bn), iters=5)
        res.append(f"bn{bn}: {us:7.1f}us {2 * T * Cout * Cin * k * B / us / 1e6:7.1f}TF")
    byt = B * T * (Cin + Cout) * 2
    print(f"conv T={T} Cin={Cin} Cout={Cout} k={k}: " + "  ".join(res) + f"   (in+out {byt / 1e6:.0f} MB)", flush=True)


if __name__ == "__main__":
    for M in (750, 1500, 6000):
        plain(M, 768, 768)
        plain(M, 2304, 768)
        plain(M, 1024, 768)
        plain(M, 768, 512)
    plain(1500, 32, 768, bns=(32,))
    plain(8192, 8192, 8192, bns=(128,))
    conv(750, 1536, 1536, 3)
    conv(1500, 384, 384, 3)
    for (T, C) in ((12000, 256), (60000, 128), (240000, 64), (480000, 32)):
        for k in (3, 11):
            conv(T, C, C, k)
