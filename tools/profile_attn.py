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
tr = torch.zeros(128, dtype=torch.int64, device="cuda")
_lib.lib().vb_attn_set_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attn_fused(q, (hd, T * hd, H * T * hd), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, k_self=k, vt_self=vt, Tp=Tp,
               k_text=kt, vt_text=vtt, L=L, Lp=80, gate=gate, out_fp16=out)
torch.cuda.synchronize()
_lib.lib().vb_attn_set_trace(None)
t = tr.cpu()
t0 = int(t[0])
print("softmax thread (row 0, group 0): start | wait_S | max+barrier | exp | wait_Pbuf | store+arrive   (cycles)")
for j in range(7):
    r = [int(v) for v in t[j * 8: j * 8 + 6]]
    print(f" tile {j}: @{r[0] - t0:6d} | {r[1] - r[0]:5d} | {r[2] - r[1]:5d} | {r[3] - r[2]:5d} | {r[4] - r[3]:5d} | {r[5] - r[4]:5d}")
print("MMA thread: start | wait_P | issue MMA2")
for j in range(7):
    r = [int(v) for v in t[64 + 4 * j: 64 + 4 * j + 3]]
    print(f" tile {j}: @{r[0] - t0:6d} | {r[1] - r[0]:5d} | {r[2] - r[1]:5d}")
# device time per launch from a graph of 20
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        ops.attn_fused(q, (hd, T * hd, H * T * hd), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, k_self=k, vt_self=vt, Tp=Tp,
                       k_text=kt, vt_text=vtt, L=L, Lp=80, gate=gate, out_fp16=out)
g.replay()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"attn_fused self+text: {a.elapsed_time(b) / 100 * 1e3:.2f} us per launch (graph of 20)")
