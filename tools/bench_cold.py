"""Warm vs cold-weight GEMM timing: rotate through enough distinct weight copies to defeat the L2,
optionally interleave a different kernel (norm_mod) to defeat the instruction cache."""
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


M, N, K = 1500, 2304, 768
for dt in (torch.float32, torch.float16):
    a = torch.randn(M, K, device="cuda").to(dt)
    ws = [torch.randn(N, K, device="cuda").to(dt) for _ in range(REP)]   # 24 x 7 MB (fp32) > L2
    out = torch.empty(M, N, device="cuda")
    xn = torch.randn(M, K, device="cuda")
    wn = torch.ones(K, device="cuda")
    xo = torch.empty(M, K, device="cuda", dtype=dt)
    t_warm = timeit(lambda i: ops.gemm(a, ws[0], out_f32=out, out_ld=N))
    t_cold = timeit(lambda i: ops.gemm(a, ws[i], out_f32=out, out_ld=N))

    def mixed(i):
        ops.norm_mod(xn, kind=0, w=wn, eps=1e-5, out_fp16=xo)
        ops.gemm(xo, ws[i], out_f32=out, out_ld=N)
    t_norm = timeit(lambda i: ops.norm_mod(xn, kind=0, w=wn, eps=1e-5, out_fp16=xo))
    t_mixed = timeit(mixed)
    print(f"{'tf32' if dt == torch.float32 else 'fp16'} wqkv: warm {t_warm:.1f} us, cold weights {t_cold:.1f} us, "
          f"norm alone {t_norm:.1f} us, norm+gemm(cold) {t_mixed:.1f} us (gemm part {t_mixed - t_norm:.1f})")
