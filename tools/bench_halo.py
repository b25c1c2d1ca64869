"""Per-tap implicit GEMM vs haloed tile kernel on the last two HiFi-GAN stages (device time via graph replay)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_gemm import timeit  # noqa: E402
from versband_b200 import _lib, ops  # noqa: E402

for (T, C) in ((240000, 64), (480000, 32)):
    for k, dil in ((3, 1), (7, 3), (11, 5), (11, 1)):
        x = torch.randn(1, T, C, device="cuda").half()
        w = torch.randn(C, k * 64, device="cuda").half()
        res = torch.randn(1, T, C, device="cuda")
        out = torch.empty(1, T, C, device="cuda")
        nxt = torch.empty(1, T, C, device="cuda", dtype=torch.float16)
        pad = (k * dil - dil) // 2
        r = []
        for mode in (0, 1):
            _lib.lib().vb_set_conv_halo(mode)
            us1 = timeit(lambda: ops.gemm(x, w, M=T, G1=1, kpt=1, dil=dil, pad=pad, a_c2=(1, 0), out_fp16=nxt, outb_ld=C,
                                          outb_g=(T * C, 0), act_out=ops.ACT_LRELU, act_out_slope=0.1), iters=5)
            us2 = timeit(lambda: ops.gemm(x, w, M=T, G1=1, kpt=1, dil=dil, pad=pad, a_c2=(1, 0), residual=res, out_f32=out,
                                          out_ld=C, out_g=(T * C, 0), out_fp16=nxt, outb_ld=C, outb_g=(T * C, 0),
                                          act_out=ops.ACT_LRELU, act_out_slope=0.1), iters=5)
            r.append((us1, us2))
        _lib.lib().vb_set_conv_halo(0)
        b1 = T * C * 4 / 1e6
        b2 = T * C * (2 + 4 + 4 + 2) / 1e6
        print(f"T={T} C={C} k={k} d={dil}: conv1 (fp16 in/out, {b1:.0f} MB) per-tap {r[0][0]:6.1f} -> halo {r[1][0]:6.1f} us "
              f"({b1 / r[1][0] * 1e-3:.2f} TB/s);  conv2 (+fp32 residual/out, {b2:.0f} MB) {r[0][1]:6.1f} -> {r[1][1]:6.1f} us "
              f"({b2 / r[1][1] * 1e-3:.2f} TB/s)", flush=True)
