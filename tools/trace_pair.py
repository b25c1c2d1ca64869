"""Per-tile timeline of vb_resblock_pair (CTA 0): MMA thread, one epilogue-1 thread, one epilogue-2 thread."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from versband_b200 import ops, _lib
from versband_b200.dit import pack_conv_weight
C, k, d, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.randn(1, T, C, device="cuda", generator=g).half()
w1 = pack_conv_weight(torch.randn(C, C, k, device="cuda", generator=g) / (C * k) ** 0.5, dtype=torch.float16)
w2 = pack_conv_weight(torch.randn(C, C, k, device="cuda", generator=g) / (C * k) ** 0.5, dtype=torch.float16)
b1 = torch.zeros(C, device="cuda"); b2 = torch.zeros(C, device="cuda")
out = torch.zeros_like(y)
for _ in range(3):
    ops.resblock_pair(y, out, w1, w2, b1, b2, B=1, T=T, C=C, k=k, dil=d, slope=0.1)
tr = torch.zeros(12 * 16, dtype=torch.int64, device="cuda")
_lib.lib().vb_resblock_set_trace(ctypes.c_void_p(tr.data_ptr()))
ops.resblock_pair(y, out, w1, w2, b1, b2, B=1, T=T, C=C, k=k, dil=d, slope=0.1)
torch.cuda.synchronize()
_lib.lib().vb_resblock_set_trace(None)
t = tr.cpu().view(12, 16)
t0 = int(t[0, 0])
print(f"C={C} k={k} d={d}: per tile [cycles]  MMA1: start waitY issue | MMA2: start waitT1 issue | epi1: start waitA1 work | epi2: start waitA2 work")
for i in range(10):
    r = [int(v) for v in t[i]]
    print(f" tile {i}: MMA1 @{r[0]-t0:6d} {r[1]-r[0]:5d} {r[2]-r[1]:5d} | MMA2 @{r[3]-t0:6d} {r[4]-r[3]:5d} {r[5]-r[4]:5d} | "
          f"epi1 @{r[6]-t0:6d} {r[7]-r[6]:5d} {r[8]-r[7]:5d} | epi2 @{r[9]-t0:6d} {r[10]-r[9]:5d} {r[11]-r[10]:5d}")
