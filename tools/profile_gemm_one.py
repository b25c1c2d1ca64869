"""A few vb_gemm launches of chosen shapes, for `ncu --set full` captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versband_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "voc32"
half = which.endswith("_h")  # fp16 operands (the default precision mode of the engines)
which = which[:-2] if half else which
opd = (lambda t: t.half()) if half else (lambda t: t)
if which.startswith("voc"):
    C = int(which[3:])
    T = {32: 480000, 64: 240000, 128: 60000, 256: 12000}[C]
    k = 3
    x = torch.randn(1, T, C, device="cuda").half()
    w = torch.randn(C, k * 64 * ((C + 63) // 64), device="cuda").half()
    out = torch.empty(1, T, C, device="cuda", dtype=torch.float16)
    fn = lambda: ops.gemm(x, w, M=T, G1=1, kpt=(C + 63) // 64, dil=1, pad=1, a_c2=(1, 0), out_fp16=out, outb_ld=C,
                          outb_g=(T * C, 0))
elif which == "vae":
    T, C, k = 750, 1536, 3
    x = torch.randn(1, T, C, device="cuda").half()
    w = torch.randn(C, k * C, device="cuda").half()
    out = torch.empty(1, T, C, device="cuda")
    fn = lambda: ops.gemm(x, w, M=T, G1=1, kpt=C // 64, dil=1, pad=1, a_c2=(1, 0), out_f32=out, out_ld=C,
                          out_g=(T * C, 0))
elif which == "mq":  # the DiT mq projection in tf32 mode: bias, fp32 operand-slot output
    M, N, K = 1500, 768, 768
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda")
    fn = lambda: ops.gemm(a, b, bias=bias, out_fp16=o, outb_ld=N)
elif which == "wo":  # the DiT attention output projection in tf32 mode: gate, residual, fp32 out; 1-CTA bn96
    M, N, K = 1500, 768, 768
    a = opd(torch.randn(M, K, device="cuda"))
    b = opd(torch.randn(N, K, device="cuda"))
    gate = torch.randn(2, N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    o = torch.empty(M, N, device="cuda")
    fn = lambda: ops.gemm(a, b, gate=gate, gate_rows=750, gate_ld=N, residual=res, out_f32=o, out_ld=N, bn=96)
elif which == "w13":  # swiglu, fp32 operand-slot output, bn128
    M, N, K = 3072, 1024, 768
    a = opd(torch.randn(M, K, device="cuda"))
    b = opd(torch.randn(N, K, device="cuda"))
    o = torch.empty(M, N // 2, device="cuda", dtype=torch.float16 if half else torch.float32)
    fn = lambda: ops.gemm(a, b, swiglu=True, out_fp16=o, outb_ld=N // 2, bn=128)
else:
    M, N, K = 1500, 2304 if "qkv" in which else 768, 768
    dt = torch.float32 if "tf32" in which else torch.float16
    a = torch.randn(M, K, device="cuda").to(dt)
    b = torch.randn(N, K, device="cuda").to(dt)
    o = torch.empty(M, N, device="cuda")
    fn = lambda: ops.gemm(a, b, out_f32=o, out_ld=N)
for _ in range(5):
    fn()
torch.cuda.synchronize()
