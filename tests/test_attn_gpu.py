"""vb_attn_fused (tcgen05 flash attention) vs a plain PyTorch fp32 reference of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from versband_b200 import ops  # noqa: E402

H16 = torch.float16


def ref_attn(q, k, v, scale):
    s = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()) * scale
    return torch.einsum("bhts,bhsd->bhtd", torch.softmax(s, -1), v.float())


def make_vt(v, pad):
    B, H, S, hd = v.shape
    vt = torch.zeros(B, H, hd, pad, device=v.device, dtype=H16)
    vt[..., :S] = v.transpose(2, 3)
    return vt


@pytest.mark.parametrize("T", [750, 128, 100, 300])
@pytest.mark.parametrize("mode", ["self", "text", "both"])
@pytest.mark.parametrize("hd", [96, 64])
def test_attn_fused(T, mode, hd):
    torch.manual_seed(T + hd)
    B, H, L = 2, 4, 80
    Tp, Lp = (T + 7) // 8 * 8, 80
    scale = 1 / math.sqrt(hd)
    q = (torch.randn(B, H, T, hd, device="cuda") * 1.5).to(H16)
    k = (torch.randn(B, H, T, hd, device="cuda") * 1.5).to(H16)
    v = torch.randn(B, H, T, hd, device="cuda").to(H16)
    kt = (torch.randn(B, H, L, hd, device="cuda") * 1.5).to(H16)
    vtx = torch.randn(B, H, L, hd, device="cuda").to(H16)
    gate = torch.randn(H, device="cuda")
    out = torch.zeros(B * T, H * hd, device="cuda", dtype=H16)
    outf = torch.zeros(B * T, H * hd, device="cuda")
    kw = dict(B=B, T=T, H=H, hd=hd, scale=scale, out_fp16=out, out_f32=outf)
    ref = torch.zeros(B, H, T, hd, device="cuda")
    if mode in ("self", "both"):
        kw.update(k_self=k, vt_self=make_vt(v, Tp), Tp=Tp)
        ref = ref + ref_attn(q, k, v, scale)
    if mode in ("text", "both"):
        kw.update(k_text=kt, vt_text=make_vt(vtx, Lp), L=L, Lp=Lp)
        g = gate.tanh().view(1, H, 1, 1) if mode == "both" else 1.0
        if mode == "both":
            kw.update(gate=gate)
        ref = ref + g * ref_attn(q, kt, vtx, scale)
    ops.attn_fused(q, (hd, T * hd, H * T * hd), **kw)
    torch.cuda.synchronize()
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, H * hd)
    err = float((outf - ref).abs().max() / ref.abs().max())
    assert err < 8e-3, err
    assert float((out.float() - outf).abs().max()) <= float(outf.abs().max()) * 8e-3


def test_attn_fused_token_major_q():
    """q in [N, H*hd] layout (the MoE cross-attention's projection output)."""
    torch.manual_seed(0)
    B, H, T, hd, L = 2, 8, 750, 96, 80
    D = H * hd
    qf = (torch.randn(B * T, D, device="cuda")).to(H16)
    kt = torch.randn(B, H, L, hd, device="cuda").to(H16)
    vtx = torch.randn(B, H, L, hd, device="cuda").to(H16)
    out = torch.zeros(B * T, D, device="cuda")
    ops.attn_fused(qf, (D, hd, T * D), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, k_text=kt, vt_text=make_vt(vtx, 80),
                   L=L, Lp=80, out_f32=out)
    q4 = qf.view(B, T, H, hd).permute(0, 2, 1, 3)
    ref = ref_attn(q4, kt, vtx, hd ** -0.5).permute(0, 2, 1, 3).reshape(B * T, D)
    assert float((out - ref).abs().max() / ref.abs().max()) < 8e-3
