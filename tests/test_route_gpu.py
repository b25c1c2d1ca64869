"""The caption cross-attention folded into its gate (VB_EPI_ROUTE + vb_route_prepare + vb_router's
cap_partial input) against a plain fp32 PyTorch restatement of vocal2music_moe.py:117-131:
MultiheadAttention(q = tokens, k = v = caption) -> out_proj -> caption gating Linear."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from versband_b200 import ops  # noqa: E402
from versband_b200.dit import tf32_round  # noqa: E402


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("B,T,L", [(2, 750, 80), (1, 130, 96), (3, 64, 17)])
def test_route_gemm_matches_attention_gate(B, T, L, dtype):
    torch.manual_seed(1)
    dev = "cuda"
    H, hd, E, LK = 8, 96, 4, 96
    D = H * hd
    x = torch.randn(B, T, D, device=dev)
    kv = torch.randn(B * L, 2 * D, device=dev) * 0.5          # fp32 k|v projection of the caption
    wq = torch.randn(D, D, device=dev) / math.sqrt(D)
    bq = torch.randn(D, device=dev) * 0.1
    capg_w = torch.randn(E, D, device=dev) / math.sqrt(D)     # W_gate W_o
    sc = 1.0 / math.sqrt(hd)
    rnd = (lambda t: t.to(torch.float16).float()) if dtype == torch.float16 else tf32_round
    xo = rnd(x)

    # ---- reference: plain attention -> gate, fp32, on the operand-rounded inputs
    q = (xo @ wq.t() + bq).view(B, T, H, hd).permute(0, 2, 1, 3)
    k = kv[:, :D].view(B, L, H, hd).permute(0, 2, 1, 3)
    v = kv[:, D:].view(B, L, H, hd).permute(0, 2, 1, 3)
    p = torch.softmax((q @ k.transpose(-1, -2)) * sc, -1)
    ctx = (p @ v).permute(0, 2, 1, 3).reshape(B, T, D)
    ref = ctx @ capg_w.t()                                     # [B, T, E]

    # ---- ours
    kpad = hd if dtype == torch.float32 else 128
    wqT = torch.zeros(H, D, kpad, device=dev)
    wqT[:, :, :hd] = (wq * sc).view(H, hd, D).permute(0, 2, 1)
    wqT = rnd(wqT).to(dtype)
    kvo = rnd(kv).to(dtype)
    wk = torch.zeros(B, H * LK, D, device=dev, dtype=dtype)
    ops.gemm(kvo[:, :D].view(B, L, H, hd).permute(0, 2, 1, 3), wqT, M=L, N=D, G1=H, G2=B, a_c2=(1, 0), a_c3=(0, 1),
             b_c2=(1, 0, 0), out_fp16=wk, outb_ld=D, outb_g=(LK * D, H * LK * D))
    bs = torch.empty(B, H, LK, device=dev)
    vg = torch.empty(B, H, LK, E, device=dev)
    ops.route_prepare(kv, B, L, H, hd, (bq * sc).contiguous(), capg_w, bs, vg)
    assert (bs[:, :, L:] == -1e30).all() and (vg[:, :, L:] == 0).all()
    part = torch.full((B * T, H, E), 7.0, device=dev)
    ops.gemm(xo.to(dtype), wk, M=T, G1=B, a_c2=(1, 0), b_c2=(1, 0, 0), mode=ops.EPI_ROUTE, bias=bs, bias_g1=H * LK,
             route_vg=vg, route_vg_g1=H * LK * E, out_f32=part, out_ld=H * E)
    got = part.sum(1).view(B, T, E)
    err = float((got - ref).abs().max() / ref.abs().max())
    tol = 3e-3 if dtype == torch.float32 else 2e-2   # WK is itself a rounded product (tf32 / fp16 operands)
    assert err < tol, f"max rel-to-peak err {err:.3e}"
