"""Experiment: does a row-shifted UMMA A descriptor (start + s*128 B, base_offset = s%8) read rows s.. of a
SWIZZLE_128B K-major tile correctly? (prerequisite for haloed conv tiles)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
M, N, K = 1024, 64, 128
a = torch.randn(M, K, device="cuda").half()
b = torch.randn(N, K, device="cuda").half()
ref = a.float() @ b.float().t()
for variant in ("base_offset", "no_base_offset"):
    for s in (0, 1, 3, 5, 8, 11, 16, 40):
        _lib.lib().vb_set_debug_shift(s | ((variant == 'no_base_offset') << 16))
        _lib.lib().vb_set_gemm2(0)
        out = torch.zeros(M, N, device="cuda")
        ops.gemm(a, b, out_f32=out, out_ld=N, bn=64)
        torch.cuda.synchronize()
        ok_rows = 0
        bad = 0
        for t0 in range(0, M, 128):
            good = out[t0:t0 + 128 - s] - ref[t0:t0 + 128 - s]
            e = good.abs().max().item() / ref.abs().max().item()
            bad = max(bad, e)
        print(f"{variant:15s} shift {s:2d}: max rel err over rows that stay inside the tile = {bad:.2e}")
_lib.lib().vb_set_debug_shift(0)
_lib.lib().vb_set_gemm2(1)
