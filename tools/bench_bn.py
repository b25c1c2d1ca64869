"""N-tile sweep (1-CTA persistent kernel) on the DiT shapes, device time via graph replay."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_gemm import timeit  # noqa: E402
from versband_b200 import ops  # noqa: E402


def run(M, N, K, dt, G=1, swiglu=False):
    a = torch.randn(G, M, K, device="cuda").to(dt)
    b = torch.randn(G, N, K, device="cuda").to(dt)
    out = torch.empty(G, M, N, device="cuda")
    res = []
    for bn in (32, 64, 96, 128, 192, 256):
        if bn > N and bn > 32:
            continue
        us = timeit(lambda: ops.gemm(a, b, M=M, G1=G, a_c2=(1, 0), b_c2=(1, 0, 0), out_f32=out, out_ld=N,
                                     out_g=(M * N, 0), bn=bn))
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn) * G
        res.append(f"bn{bn}:{us:5.1f}({tiles})")
    name = "tf32" if dt == torch.float32 else "fp16"
    print(f"{name} G={G} M={M} N={N} K={K}: " + "  ".join(res), flush=True)


for dt in (torch.float32, torch.float16):
    run(1500, 2304, 768, dt)
    run(1500, 768, 768, dt)
    run(1500, 1024, 768, dt)
    run(4096, 1024, 768, dt)
    run(4096, 768, 512, dt)
    run(1500, 1024, 192, dt, G=4)
    run(1500, 192, 512, dt, G=4)
