"""ORACLE — test infrastructure only, never the product path.

CPU fp32 restatement (plain PyTorch tensor ops, functional, state_dict in / tensors out) of the
AccompBand inference hot path of AaronZ345/VersBand.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import this file, and only as the
checker or the timed CPU baseline.

Pinning: the reference ships no tests or golden vectors.  This restatement
is pinned against the *live reference modules* imported from /root/reference in the build
container (oracle/make_golden.py → tests/golden/*.pt; tests/test_oracle_golden.py), so
it is "pinned to reference outputs generated here", not to reference-owned fixtures.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Arithmetic is the reference's: fp32 everywhere, dense evaluation of all 12 experts.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# small pieces (ldm/modules/diffusionmodules/flag_large_dit_moe.py)
# --------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """flag_large_dit_moe.py:52-77 — x * rsqrt(mean(x^2) + eps) in fp32, then * weight."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * weight


def modulate(x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """flag_large_dit_moe.py:80-81."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def timestep_embedding(t: Tensor, dim: int = 256, max_period: float = 10000.0) -> Tensor:
    """flag_large_dit_moe.py:109-127 — cat[cos, sin] of t * exp(-ln(P) i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None].to(t.device)
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def t_embedder(sd: SD, t: Tensor) -> Tensor:
    """flag_large_dit_moe.py:129-133 — Linear(256→D) SiLU Linear(D→D)."""
    h = F.linear(timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    return F.linear(F.silu(h), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])


def c_embedder(sd: SD, caption: Tensor) -> Tensor:
    """flag_large_dit_moe.py:149-160 — Linear, exact GELU, Linear, LayerNorm(eps 1e-5)."""
    h = F.linear(caption, sd["c_embedder.mlp.0.weight"], sd["c_embedder.mlp.0.bias"])
    h = F.linear(F.gelu(h), sd["c_embedder.mlp.2.weight"], sd["c_embedder.mlp.2.bias"])
    return F.layer_norm(h, (h.shape[-1],), sd["c_embedder.mlp.3.weight"], sd["c_embedder.mlp.3.bias"], 1e-5)


def rope_angles(head_dim: int, length: int, theta: float = 10000.0) -> Tensor:
    """vocal2music_moe.py:464-475 — angle[pos, i] = pos / theta^(2i/head_dim)."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    return torch.outer(torch.arange(length, dtype=torch.float), freqs).float()


def apply_rope(x: Tensor, ang: Tensor) -> Tensor:
    """flag_large_dit_moe.py:237-269 — adjacent pairs rotated as complex numbers. x: [B,T,H,hd]."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    cis = torch.polar(torch.ones_like(ang), ang)[None, :, None, :]
    return torch.view_as_real(xc * cis).flatten(3).type_as(x)


def sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """F.scaled_dot_product_attention with an all-True mask (flag_large_dit_moe.py:381-386):
    softmax(q k^T / sqrt(d)) v.  q,k,v: [B,H,T,d]."""
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    return torch.matmul(torch.softmax(s, dim=-1), v)


def attention(sd: SD, pre: str, x: Tensor, y: Tensor, ang: Tensor, n_heads: int) -> Tensor:
    """Attention.forward, flag_large_dit_moe.py:323-406 (fp32 SDPA branch, qk_norm=None)."""
    B, T, D = x.shape
    hd = D // n_heads
    q = F.linear(x, sd[pre + "wq.weight"]).view(B, T, n_heads, hd)
    k = F.linear(x, sd[pre + "wk.weight"]).view(B, T, n_heads, hd)
    v = F.linear(x, sd[pre + "wv.weight"]).view(B, T, n_heads, hd)
    q, k = apply_rope(q, ang), apply_rope(k, ang)
    out = sdpa(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    yk = F.linear(y, sd[pre + "wk_y.weight"]).view(B, -1, n_heads, hd)
    yv = F.linear(y, sd[pre + "wv_y.weight"]).view(B, -1, n_heads, hd)
    out_y = sdpa(q.permute(0, 2, 1, 3), yk.permute(0, 2, 1, 3), yv.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    out = out + out_y * sd[pre + "gate"].tanh().view(1, 1, -1, 1)
    return F.linear(out.flatten(-2), sd[pre + "wo.weight"])


def swiglu_ffn(sd: SD, pre: str, x: Tensor) -> Tensor:
    """FeedForward.forward, flag_large_dit_moe.py:480-485 — w2(silu(w1 x) * w3 x)."""
    return F.linear(F.silu(F.linear(x, sd[pre + "w1.weight"])) * F.linear(x, sd[pre + "w3.weight"]),
                    sd[pre + "w2.weight"])


# --------------------------------------------------------------------------------------
# Band-MoE (ldm/modules/diffusionmodules/vocal2music_moe.py:36-185)
# --------------------------------------------------------------------------------------
def mha_cross(sd: SD, pre: str, x: Tensor, kv: Tensor, n_heads: int = 8) -> Tensor:
    """nn.MultiheadAttention(dim, 8, batch_first=True)(x, kv, kv)[0] — vocal2music_moe.py:79,119."""
    D = x.shape[-1]
    w, b = sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"]
    B, T, _ = x.shape
    L = kv.shape[1]
    hd = D // n_heads
    q = F.linear(x, w[:D], b[:D]).view(B, T, n_heads, hd).transpose(1, 2)
    k = F.linear(kv, w[D:2 * D], b[D:2 * D]).view(B, L, n_heads, hd).transpose(1, 2)
    v = F.linear(kv, w[2 * D:], b[2 * D:]).view(B, L, n_heads, hd).transpose(1, 2)
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, T, D)
    return F.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def gumbel_softmax(logits: Tensor, temperature: float, hard: bool, expo: Optional[Tensor]):
    """MoE.gumbel_softmax, vocal2music_moe.py:81-93.  `expo` is the Exp(1) draw the reference
    obtains from `torch.empty_like(logits).exponential_()`; None → draw it here (same call)."""
    if expo is None:
        expo = torch.empty_like(logits).exponential_()
    g = (logits + (-expo.log())) / temperature
    y_soft = g.softmax(dim=-1)
    if not hard:
        return y_soft, None
    index = y_soft.max(dim=-1, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft, index.squeeze(-1)


def load_balancing_loss(probs: Sequence[Tensor], masks: Sequence[Tensor], num_experts: int) -> Tensor:
    """vocal2music_moe.py:95-115."""
    all_p = torch.cat(list(probs), dim=1)
    all_m = torch.cat([m.repeat(1, num_experts) for m in masks], dim=1)
    usage = (all_p * all_m).sum(dim=0) / (all_m.sum() + 1e-10)
    return torch.mean(usage * torch.log(usage + 1e-10))


def moe_forward(sd: SD, pre: str, x: Tensor, time: Tensor, caption: Tensor, acoustic: Tensor,
                num_experts: int, temperature: float, training: bool,
                expo: Optional[Tuple[Tensor, Tensor, Tensor]] = None):
    """MoE.forward, vocal2music_moe.py:117-185 (dense evaluation, as the reference does).
    Returns (z, lb_loss, aux) with aux = router indices / probabilities for parity tests."""
    cap_feat = mha_cross(sd, pre + "cross_attention.", x, caption)
    D = acoustic.shape[-1]
    cap_flat = cap_feat.reshape(-1, D)
    shape = x.shape
    xf = x.reshape(-1, shape[-1])
    ac_flat = acoustic.reshape(-1, D)

    hl = F.linear(time, sd[pre + "high_level_gating_network.weight"], sd[pre + "high_level_gating_network.bias"])
    hl = hl.repeat_interleave(shape[1], dim=0)
    e0, e1, e2 = expo if expo is not None else (None, None, None)
    hl_p, _ = gumbel_softmax(hl, 1.0, False, e0)
    cap_mask = hl_p[:, 0].float().unsqueeze(1)
    ac_mask = hl_p[:, 1].float().unsqueeze(1)

    cap_logits = F.linear(cap_flat, sd[pre + "caption_gating_network.weight"], sd[pre + "caption_gating_network.bias"])
    ac_logits = F.linear(ac_flat, sd[pre + "acoustic_gating_network.weight"], sd[pre + "acoustic_gating_network.bias"])
    hard = not training
    cap_p, cap_idx = gumbel_softmax(cap_logits, temperature, hard, e1)
    ac_p, ac_idx = gumbel_softmax(ac_logits, temperature, hard, e2)

    z_c = torch.zeros_like(xf)
    z_a = torch.zeros_like(xf)
    for i in range(num_experts):
        z_c += swiglu_ffn(sd, f"{pre}caption_experts.{i}.", xf) * cap_p[:, i].unsqueeze(1).expand_as(xf) * cap_mask
    for i in range(num_experts):
        z_a += swiglu_ffn(sd, f"{pre}acoustic_experts.{i}.", xf) * ac_p[:, i].unsqueeze(1).expand_as(xf) * ac_mask
    y = (z_c + z_a).view(*shape)

    z = torch.zeros_like(y)
    band = shape[-1] // num_experts
    for i in range(num_experts):
        region = torch.zeros_like(z)
        region[:, :, band * i: band * (i + 1)] = 1.0
        z[:, :, band * i: band * (i + 1)] = swiglu_ffn(sd, f"{pre}freq_experts.{i}.", y * region)[:, :, band * i: band * (i + 1)]

    lb = load_balancing_loss([cap_p, ac_p], [cap_mask, ac_mask], num_experts)
    aux = {"cap_idx": cap_idx, "ac_idx": ac_idx, "cap_logits": cap_logits, "ac_logits": ac_logits,
           "cap_mask": cap_mask, "ac_mask": ac_mask, "cap_feat": cap_feat, "y": y}
    return z, lb, aux


# --------------------------------------------------------------------------------------
# DiT (vocal2music_moe.py:187-434)
# --------------------------------------------------------------------------------------
def transformer_block(sd: SD, pre: str, x: Tensor, caption: Tensor, ang: Tensor, adaln_input: Tensor,
                      time: Tensor, acoustic: Tensor, n_heads: int, num_experts: int, temperature: float,
                      training: bool, expo=None, norm_eps: float = 1e-5):
    """TransformerBlock.forward, vocal2music_moe.py:215-266 (adaLN branch)."""
    mod = F.linear(F.silu(adaln_input), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    y = rms_norm(caption, sd[pre + "attention_y_norm.weight"], norm_eps)
    h = x + gate_msa.unsqueeze(1) * attention(
        sd, pre + "attention.", modulate(rms_norm(x, sd[pre + "attention_norm.weight"], norm_eps), shift_msa, scale_msa),
        y, ang, n_heads)
    ffn_in = modulate(rms_norm(h, sd[pre + "ffn_norm.weight"], norm_eps), shift_mlp, scale_mlp)
    out, lb, aux = moe_forward(sd, pre + "feed_forward.", ffn_in, time, caption, acoustic, num_experts,
                               temperature, training, expo)
    aux["h"] = h
    aux["ffn_in"] = ffn_in
    return h + gate_mlp.unsqueeze(1) * out, lb, aux


def stem_branch(sd: SD, emb_key: str, proj_pre: str, idx: Tensor) -> Tensor:
    """Embedding → Conv1d k5 → LeakyReLU(0.01) → AvgPool1d(2) — vocal2music_moe.py:336-350,388-391."""
    e = F.embedding(idx.squeeze(1), sd[emb_key]).transpose(1, 2)
    h = F.conv1d(e, sd[proj_pre + "0.weight"], sd[proj_pre + "0.bias"], padding=sd[proj_pre + "0.weight"].shape[-1] // 2)
    return F.avg_pool1d(F.leaky_relu(h, 0.01), 2).transpose(1, 2)


def dit_conditioning(sd: SD, context: dict) -> dict:
    """Step-invariant part of TxtFlagLargeDiT.forward (vocal2music_moe.py:384-393,407-413)."""
    midi = stem_branch(sd, "midi_embedding.weight", "midi_proj.", context["c_concat"]["midi"])
    beats = stem_branch(sd, "beats_embedding.weight", "beats_proj.", context["c_concat"]["beats"])
    acoustic = midi + beats
    acoustic = F.conv1d(acoustic.transpose(1, 2), sd["final_proj.weight"], sd["final_proj.bias"]).transpose(1, 2)
    caption = c_embedder(sd, context["c_crossattn"])
    pool = caption.sum(dim=1) / caption.shape[1]
    cap_emb = F.linear(F.layer_norm(pool, (pool.shape[-1],), sd["cap_embedder.0.weight"], sd["cap_embedder.0.bias"], 1e-5),
                       sd["cap_embedder.1.weight"], sd["cap_embedder.1.bias"])
    return {"acoustic": acoustic, "caption": caption, "cap_emb": cap_emb}


def dit_forward(sd: SD, x: Tensor, t: Tensor, context: dict, *, depth: int, n_heads: int,
                num_experts: int = 4, temperature: float = 2.0, training: bool = False,
                loss_w: float = 1.0, expo: Optional[List[Tuple[Tensor, Tensor, Tensor]]] = None,
                return_aux: bool = False):
    """TxtFlagLargeDiT.forward, vocal2music_moe.py:375-434.

    x [B,C,T] fp32, t [B] int64, context = {'c_crossattn': [B,L,ori_dim], 'c_concat': {'midi','beats'}}.
    `expo[k]` = the three Exp(1) draws of block k in reference order (group gate [N,2], caption
    gate [N,E], acoustic gate [N,E]); None → drawn from torch's default generator exactly where
    the reference draws them.  `loss_w` is the reference's running scale *before* this call's
    decay (vocal2music_moe.py:427-429)."""
    cond = dit_conditioning(sd, context)
    acoustic, caption = cond["acoustic"], cond["caption"]
    h = F.conv1d(x, sd["proj_in.weight"], sd["proj_in.bias"], padding=sd["proj_in.weight"].shape[-1] // 2).transpose(1, 2)
    if abs(h.shape[1] - acoustic.shape[1]) <= 2:  # vocal2music_moe.py:397-401
        if h.shape[1] > acoustic.shape[1]:
            acoustic = torch.cat([acoustic, acoustic[:, -1:, :].repeat(1, h.shape[1] - acoustic.shape[1], 1)], dim=1)
        else:
            acoustic = acoustic[:, : h.shape[1], :]
    t_emb = t_embedder(sd, t)
    h = acoustic + h
    adaln_input = t_emb + cond["cap_emb"]
    hd = h.shape[-1] // n_heads
    ang = rope_angles(hd, h.shape[1]).to(h.device)
    loss = 0
    auxes = []
    for k in range(depth):
        h, lb, aux = transformer_block(sd, f"blocks.{k}.", h, caption, ang, adaln_input, t_emb, acoustic,
                                       n_heads, num_experts, temperature, training,
                                       None if expo is None else expo[k])
        loss = loss + lb
        auxes.append(aux)
    loss = loss / depth
    if loss_w > 0.01:
        loss_w = loss_w * 0.9999
    loss = loss * loss_w
    # FinalLayer, vocal2music_moe.py:268-291
    mod = F.linear(F.silu(adaln_input), sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    o = modulate(F.layer_norm(h, (h.shape[-1],), None, None, 1e-6), shift, scale)
    o = F.linear(o, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"]).transpose(1, 2)
    if return_aux:
        return o, loss, auxes
    return o, loss


# --------------------------------------------------------------------------------------
# Flow sampler (ldm/models/diffusion/cfm1_audio.py:145-162, cfm1_audio_sampler.py:87-116,
# torchdyn NeuralODE(solver='euler') — third party, unpinned, restated: SURVEY.md §8c)
# --------------------------------------------------------------------------------------
def euler_schedule(n_points: int = 25) -> List[Tuple[int, float]]:
    """(t_long, dt) per Euler step for t_span = linspace(0, 1, n_points) in fp32.
    t_long = trunc(1000 * t) as in cfm1_audio.py:156; dt re-derived from t_span every step."""
    t_span = torch.linspace(0, 1, n_points)
    t = t_span[0]
    dt = t_span[1] - t
    out = []
    for k in range(1, n_points):
        out.append((int(torch.tensor([t * 1000]).long().item()), float(dt)))
        t = t + dt
        if k < n_points - 1:
            dt = t_span[k + 1] - t
    return out


def sample_cfg(sd: SD, x0: Tensor, cond: dict, uncond: dict, scale: float, *, depth: int, n_heads: int,
               num_experts: int = 4, temperature: float = 2.0, n_points: int = 25,
               expo: Optional[list] = None, return_traj: bool = False):
    """Wrapper_cfg + fixed-step Euler: x ← x + dt * (e_u + s (e_c − e_u)), cond pass first."""
    x = x0
    traj = [x0]
    for step, (t_long, dt) in enumerate(euler_schedule(n_points)):
        t = torch.full((x.shape[0],), t_long, dtype=torch.long, device=x.device)
        e_c, _ = dit_forward(sd, x, t, cond, depth=depth, n_heads=n_heads, num_experts=num_experts,
                             temperature=temperature, expo=None if expo is None else expo[step][0])
        e_u, _ = dit_forward(sd, x, t, uncond, depth=depth, n_heads=n_heads, num_experts=num_experts,
                             temperature=temperature, expo=None if expo is None else expo[step][1])
        e = e_u + scale * (e_c - e_u)
        x = x + dt * e
        if return_traj:
            traj.append(x)
    return (x, traj) if return_traj else x


# --------------------------------------------------------------------------------------
# VAE (ldm/models/autoencoder1d.py)
# --------------------------------------------------------------------------------------
def _gn(sd: SD, pre: str, x: Tensor) -> Tensor:
    """Normalize(): GroupNorm(32, eps=1e-6, affine) — autoencoder1d.py:165-166."""
    return F.group_norm(x, 32, sd[pre + "weight"], sd[pre + "bias"], 1e-6)


def _swish(x: Tensor) -> Tensor:
    """autoencoder1d.py:168-170."""
    return x * torch.sigmoid(x)


def _conv(sd: SD, pre: str, x: Tensor, stride: int = 1, padding: Optional[int] = None) -> Tensor:
    w = sd[pre + "weight"]
    if padding is None:
        padding = w.shape[-1] // 2
    return F.conv1d(x, w, sd.get(pre + "bias"), stride=stride, padding=padding)


def resnet_block(sd: SD, pre: str, x: Tensor) -> Tensor:
    """ResnetBlock1D.forward, autoencoder1d.py:211-231 (temb None, dropout 0)."""
    h = _conv(sd, pre + "conv1.", _swish(_gn(sd, pre + "norm1.", x)))
    h = _conv(sd, pre + "conv2.", _swish(_gn(sd, pre + "norm2.", h)))
    if (pre + "nin_shortcut.weight") in sd:
        x = _conv(sd, pre + "nin_shortcut.", x)
    elif (pre + "conv_shortcut.weight") in sd:
        x = _conv(sd, pre + "conv_shortcut.", x)
    return x + h


def attn_block(sd: SD, pre: str, x: Tensor) -> Tensor:
    """AttnBlock1D.forward, autoencoder1d.py:253-274 — note the scale is C^-0.5 (channel count)."""
    h = _gn(sd, pre + "norm.", x)
    q, k, v = _conv(sd, pre + "q.", h), _conv(sd, pre + "k.", h), _conv(sd, pre + "v.", h)
    c = q.shape[1]
    w = torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** (-0.5))
    w = torch.softmax(w, dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1))
    return x + _conv(sd, pre + "proj_out.", h)


def vae_decode(sd: SD, z: Tensor, ddconfig: dict, prefix: str = "") -> Tensor:
    """AutoencoderKL.decode → Decoder1D.forward, autoencoder1d.py:55-58,480-513."""
    p = prefix
    n_levels = len(ddconfig["ch_mult"])
    n_res = ddconfig["num_res_blocks"]
    up_levels = [i + 1 for i in ddconfig.get("down_layers", [])]  # autoencoder1d.py:423
    h = _conv(sd, p + "post_quant_conv.", z)
    d = p + "decoder."
    h = _conv(sd, d + "conv_in.", h)
    h = resnet_block(sd, d + "mid.block_1.", h)
    h = attn_block(sd, d + "mid.attn_1.", h)
    h = resnet_block(sd, d + "mid.block_2.", h)
    for lvl in reversed(range(n_levels)):
        for blk in range(n_res + 1):
            h = resnet_block(sd, f"{d}up.{lvl}.block.{blk}.", h)
            if f"{d}up.{lvl}.attn.{blk}.norm.weight" in sd:
                h = attn_block(sd, f"{d}up.{lvl}.attn.{blk}.", h)
        if lvl in up_levels:  # Upsample1D, autoencoder1d.py:287-291
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"{d}up.{lvl}.upsample.conv.", h)
    h = _conv(sd, d + "conv_out.", _swish(_gn(sd, d + "norm_out.", h)))
    return h


def vae_encode_moments(sd: SD, x: Tensor, ddconfig: dict, prefix: str = "") -> Tensor:
    """AutoencoderKL.encode up to `moments` — Encoder1D.forward + quant_conv, autoencoder1d.py:49-53,384-409."""
    p = prefix
    n_levels = len(ddconfig["ch_mult"])
    n_res = ddconfig["num_res_blocks"]
    down_levels = list(ddconfig.get("down_layers", []))
    e = p + "encoder."
    h = _conv(sd, e + "conv_in.", x)
    for lvl in range(n_levels):
        for blk in range(n_res):
            h = resnet_block(sd, f"{e}down.{lvl}.block.{blk}.", h)
            if f"{e}down.{lvl}.attn.{blk}.norm.weight" in sd:
                h = attn_block(sd, f"{e}down.{lvl}.attn.{blk}.", h)
        if lvl in down_levels:  # Downsample1D, autoencoder1d.py:306-313: pad (0,1), k3 s2 p0
            h = _conv(sd, f"{e}down.{lvl}.downsample.conv.", F.pad(h, (0, 1)), stride=2, padding=0)
    h = resnet_block(sd, e + "mid.block_1.", h)
    h = attn_block(sd, e + "mid.attn_1.", h)
    h = resnet_block(sd, e + "mid.block_2.", h)
    h = _conv(sd, e + "conv_out.", _swish(_gn(sd, e + "norm_out.", h)))
    return _conv(sd, p + "quant_conv.", h)


def posterior_mode(moments: Tensor) -> Tensor:
    """DiagonalGaussianDistribution.mode(), distributions.py:7-8,42-43."""
    mean, _ = torch.chunk(moments, 2, dim=1)
    return mean


# --------------------------------------------------------------------------------------
# HiFi-GAN (vocoder/hifigan/modules/hifigan.py)
# --------------------------------------------------------------------------------------
def wn_weight(sd: SD, pre: str) -> Tensor:
    """old-style torch.nn.utils.weight_norm, dim=0: w = g * v / ||v|| (norm over dims != 0)."""
    if (pre + "weight") in sd:
        return sd[pre + "weight"]
    v, g = sd[pre + "weight_v"], sd[pre + "weight_g"]
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def hifigan_forward(sd: SD, mel: Tensor, h: dict) -> Tensor:
    """HifiGanGenerator.forward, modules/hifigan.py:126-143 (resblock '1': ResBlock1 :51-58).
    mel [B,80,T] → wav [B,1,T*prod(upsample_rates)]."""
    assert str(h["resblock"]) == "1"
    ks, ds = h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]
    x = F.conv1d(mel, wn_weight(sd, "conv_pre."), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, wn_weight(sd, f"ups.{i}."), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (kk, dd) in enumerate(zip(ks, ds)):
            pre = f"resblocks.{i * len(ks) + j}."
            r = x
            for c in range(len(dd)):
                xt = F.leaky_relu(r, 0.1)
                xt = F.conv1d(xt, wn_weight(sd, f"{pre}convs1.{c}."), sd[f"{pre}convs1.{c}.bias"],
                              dilation=dd[c], padding=(kk * dd[c] - dd[c]) // 2)
                xt = F.leaky_relu(xt, 0.1)
                xt = F.conv1d(xt, wn_weight(sd, f"{pre}convs2.{c}."), sd[f"{pre}convs2.{c}.bias"],
                              padding=(kk - 1) // 2)
                r = xt + r
            xs = r if xs is None else xs + r
        x = xs / len(ks)
    x = F.leaky_relu(x)  # default slope 0.01 — modules/hifigan.py:139
    x = F.conv1d(x, wn_weight(sd, "conv_post."), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)
