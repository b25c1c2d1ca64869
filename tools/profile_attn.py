"""A few vb_attn_fused launches at the DiT shape (2B=2, H=8, T=750, hd=96, L=80) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import ops  # noqa: E402

B, H, T, hd, L = 2, 8, 750, 96, 80
Tp = 752
H16 = torch.float16
q = torch.randn(B, H, T, hd, device="cuda").to(H16)
k = torch.randn(B, H, T, hd, device="cuda").to(H16)
vt = torch.randn(B, H, hd, Tp, device="cuda").to(H16)
kt = torch.randn(B, H, L, hd, device="cuda").to(H16)
vtt = torch.randn(B, H, hd, 80, device="cuda").to(H16)
gate = torch.randn(H, device="cuda")
out = torch.zeros(B * T, H * hd, device="cuda", dtype=H16)
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
for _ in range(4):
    if mode == "both":
        ops.attn_fused(q, (hd, T * hd, H * T * hd), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, k_self=k, vt_self=vt, Tp=Tp,
                       k_text=kt, vt_text=vtt, L=L, Lp=80, gate=gate, out_fp16=out)
    else:
        ops.attn_fused(q, (hd, T * hd, H * T * hd), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, k_text=kt, vt_text=vtt, L=L,
                       Lp=80, out_fp16=out)
torch.cuda.synchronize()

from versband_b200 import _lib
import ctypes
tr = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.lib().vb_attn_set_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attn_fused(q, (hd, T * hd, H * T * hd), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, k_self=k, vt_self=vt, Tp=Tp,
               k_text=kt, vt_text=vtt, L=L, Lp=80, gate=gate, out_fp16=out)
torch.cuda.synchronize()
t = tr.cpu().view(8, 8)
t0 = int(t[0, 0])
print("tile: wait_S  passA  passB+arrive  wait_O  accum   (cycles); start offsets")
for j in range(7):
    r = [int(v) for v in t[j]]
    print(j, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], " start", r[0] - t0)
