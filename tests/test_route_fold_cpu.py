"""Algebra behind VB_EPI_ROUTE (include/versband.h), checked on CPU in fp64 against the modules the reference
uses (vocal2music_moe.py:117-131,141): nn.MultiheadAttention(q = tokens, k = v = caption) followed by the caption
gating Linear equals, per head h,  softmax(x (s K_h W_q,h)^T + s K_h b_q,h) . (V_h (W_gate W_o)_h^T)  summed over
heads, plus W_gate b_o + b_gate. The GPU path builds exactly these operands (DiTEngine.pack / prepare)."""
import math

import torch


def test_folded_gate_equals_mha_then_linear():
    torch.manual_seed(0)
    B, T, L, D, H, E = 2, 37, 11, 48, 4, 4
    hd = D // H
    mha = torch.nn.MultiheadAttention(D, H, batch_first=True).double()
    gate = torch.nn.Linear(D, E).double()
    x = torch.randn(B, T, D, dtype=torch.float64)
    cap = torch.randn(B, L, D, dtype=torch.float64)
    with torch.no_grad():
        ref = gate(mha(x, cap, cap)[0])                                   # [B, T, E]

        ipw, ipb = mha.in_proj_weight, mha.in_proj_bias
        wq, bq = ipw[:D], ipb[:D]
        kv = cap.reshape(B * L, D) @ ipw[D:].t() + ipb[D:]                # [B*L, 2D]  (k | v)
        K = kv[:, :D].view(B, L, H, hd).permute(0, 2, 1, 3)               # [B, H, L, hd]
        V = kv[:, D:].view(B, L, H, hd).permute(0, 2, 1, 3)
        s = 1.0 / math.sqrt(hd)
        capg_w = gate.weight @ mha.out_proj.weight                        # [E, D]
        capg_b = gate.weight @ mha.out_proj.bias + gate.bias
        WK = torch.einsum("bhjd,hdi->bhji", K, (wq * s).view(H, hd, D))   # [B, H, L, D]
        bias_s = torch.einsum("bhjd,hd->bhj", K, (bq * s).view(H, hd))    # [B, H, L]
        VG = torch.einsum("bhjd,ehd->bhje", V, capg_w.view(E, H, hd))     # [B, H, L, E]
        scores = torch.einsum("bti,bhji->bhtj", x, WK) + bias_s[:, :, None, :]
        part = torch.einsum("bhtj,bhje->bthe", torch.softmax(scores, -1), VG)   # per-head partial logits
        got = part.sum(2) + capg_b
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-12)


def test_split_sum_gemm_identity():
    """(y_c + y_a) W^T == [y_c | y_a] [W | W]^T — how the frequency experts read both MoE group outputs
    (DiTEngine.forward: conv addressing with two taps N rows apart, weights repeated along K)."""
    torch.manual_seed(1)
    yc, ya = torch.randn(50, 24, dtype=torch.float64), torch.randn(50, 24, dtype=torch.float64)
    w = torch.randn(16, 24, dtype=torch.float64)
    assert torch.allclose((yc + ya) @ w.t(), torch.cat([yc, ya], 1) @ torch.cat([w, w], 1).t(), rtol=1e-12, atol=1e-12)
