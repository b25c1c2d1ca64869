"""1-CTA vs 2-CTA (cta_group::2) GEMM on the DiT / VAE shapes (device time via graph replay)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_gemm import timeit  # noqa: E402
from versband_b200 import _lib, ops  # noqa: E402


def run(M, N, K, dt, G=1):
    a = torch.randn(G, M, K, device="cuda").to(dt)
    b = torch.randn(G, N, K, device="cuda").to(dt)
    out = torch.empty(G, M, N, device="cuda")
    res = []
    for mode in (0, 2):
        _lib.lib().vb_set_gemm2(mode)
        us = timeit(lambda: ops.gemm(a, b, M=M, G1=G, a_c2=(1, 0), b_c2=(1, 0, 0), out_f32=out, out_ld=N, out_g=(M * N, 0)))
        res.append(us)
    _lib.lib().vb_set_gemm2(0)
    name = "tf32" if dt == torch.float32 else "fp16"
    print(f"{name} G={G} M={M} N={N} K={K}: 1-CTA {res[0]:6.1f} us   2-CTA {res[1]:6.1f} us   ({2 * G * M * N * K / res[1] / 1e6:6.0f} TF)", flush=True)


for dt in (torch.float32, torch.float16):
    run(1500, 2304, 768, dt)
    run(1500, 768, 768, dt)
    run(1500, 1024, 768, dt)
    run(1500, 1024, 192, dt, G=4)
    run(1500, 192, 512, dt, G=4)
    run(3000, 2304, 768, dt)
    run(6000, 2304, 768, dt)
run(750, 1536, 4608, torch.float16)
run(1500, 768, 2304, torch.float16)
run(8192, 8192, 8192, torch.float16)
