"""A few vb_resblock_pair launches at full vocoder size, for `ncu --set full` captures: argv = C k dil."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import ops  # noqa: E402
from versband_b200.dit import pack_conv_weight  # noqa: E402

C, k, d = (int(v) for v in sys.argv[1:4])
T = {32: 480000, 64: 240000}[C]
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.randn(1, T, C, device="cuda", generator=g).half()
w1 = pack_conv_weight(torch.randn(C, C, k, device="cuda", generator=g) / (C * k) ** 0.5, dtype=torch.float16)
w2 = pack_conv_weight(torch.randn(C, C, k, device="cuda", generator=g) / (C * k) ** 0.5, dtype=torch.float16)
b1 = torch.zeros(C, device="cuda")
b2 = torch.zeros(C, device="cuda")
out = torch.zeros_like(y)
for _ in range(5):
    ops.resblock_pair(y, out, w1, w2, b1, b2, B=1, T=T, C=C, k=k, dil=d, slope=0.1)
torch.cuda.synchronize()
