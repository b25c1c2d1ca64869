"""Training step of the AccompBand DiT on B200 (BASELINE configs[4]).

What the reference does per step (LightningModule training_step -> `CFM.p_losses`,
ldm/models/diffusion/cfm1_audio.py:38-63): draw t ~ U{0..999} and noise x0, build the flow-matching pair
  x_t = (t/1000) x1 + (1 - (1 - 1e-4) t/1000) x0,   u_t = x1 - (1 - 1e-4) x0,
run the DiT in TRAIN mode — soft Gumbel-softmax routing (`hard = not self.training`, vocal2music_moe.py:144-151:
every one of the 12 expert FFNs of a block runs on every token, weighted by the soft gate), temperature decay
(:146-147) and `loss_w` decay (:427-429) — and minimise  mean((v - u_t)^2) + lb_loss.

B200 formulation: every matrix product of the forward AND the backward pass — all linear layers, the k=5 convs
(as unfold + GEMM), the attention score / context products — is a vb_gemm launch on the tcgen05 tensor cores with
fp32 operands consumed as TF32 (kind::tf32, fp32 accumulate): dX = dY W and dW = dY^T X are the same K-major
GEMM with transposed operand copies. Norms, softmaxes, activations and the Gumbel arithmetic are elementwise
PyTorch autograd ops on the device (memory-bound glue); PyTorch also provides autograd's tape, the AdamW state
and `DistributedDataParallel` (NCCL all-reduce of the gradients, overlapped with the backward pass) — plumbing.
The optimiser update itself is one fused kernel (vb_adamw).

Parity: tests/test_train_gpu.py compares loss and parameter gradients with autograd through the CPU oracle
(oracle/accompband_ref.py, `training=True`) on identical weights, inputs and Gumbel noise.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops

F32 = torch.float32


# ----------------------------------------------------------------------------------------------------
# tensor-core matmuls with autograd
# ----------------------------------------------------------------------------------------------------
def _gemm_nt(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[M, K] x [N, K]^T -> [M, N] fp32 on vb_gemm (TF32 operands)."""
    M, N = a.shape[0], b.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=F32)
    ops.gemm(a, b, bias=bias, out_f32=out, out_ld=N)
    return out


def _c(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    if t.data_ptr() % 16:  # TMA needs 16-byte aligned bases
        t = t.clone()
    return t


def _kmajor(t: torch.Tensor) -> torch.Tensor:
    """A GEMM operand [.., rows, K] as vb_gemm wants it: K contiguous, K % 4 == 0 (zero padded), 16-byte aligned. A
    transposed view (the dW / dX factors of the backward pass) goes through the tiled transpose kernel
    (vb_transpose_f32, both sides coalesced) instead of a generic strided copy."""
    if t.dim() in (2, 3) and t.stride(-1) != 1 and t.stride(-2) == 1 and t.shape[-1] > 1:
        base = t.transpose(-1, -2)
        if base.data_ptr() % 16 == 0 and (t.dim() == 2 or base.stride(0) >= base.shape[1] * base.stride(1)):
            return ops.transpose_f32(base)
    return _c(_pad_k(t))


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with forward and both backward products on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = _c(x.reshape(-1, x.shape[-1]).float())
        wc = _c(w.float())
        ctx.save_for_backward(x2, wc)
        ctx.has_bias, ctx.shape = b is not None, x.shape
        y = _gemm_nt(x2, wc, None if b is None else _c(b.float()))
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = _c(dy.reshape(-1, dy.shape[-1]).float())
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm_nt(_kmajor(dy2), _kmajor(w.t())).view(ctx.shape)         # dX = dY W   (B operand = W^T, K-major in N)
        if ctx.needs_input_grad[1]:
            dw = _gemm_nt(_kmajor(dy2.t()), _kmajor(x2.t()))                    # dW = dY^T X (reduction over the tokens)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db


def linear(x, w, b=None):
    return _LinearFn.apply(x, w, b)


def conv1d_same(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """Conv1d with "same" zero padding over [B, C, T] as unfold + GEMM (K = C * k); returns [B, T, Cout]."""
    B, C, T = x.shape
    Cout, _, k = w.shape
    if k == 1:
        return linear(x.transpose(1, 2), w[:, :, 0], b)
    xp = F.pad(x, (k // 2, k // 2))
    cols = xp.unfold(2, k, 1).permute(0, 2, 1, 3).reshape(B, T, C * k)   # [B, T, C*k], (c, tap) order = w.view(Cout, C*k)
    return linear(cols, w.reshape(Cout, C * k), b)


def _pad_k(t: torch.Tensor) -> torch.Tensor:
    """Row strides must be multiples of 16 bytes (4 fp32): pad the reduction dim with zeros when needed."""
    k = t.shape[-1]
    return t if k % 4 == 0 else F.pad(t, (0, 4 - k % 4))


class _BmmNT(torch.autograd.Function):
    """Batched C[g] = A[g] B[g]^T on vb_gemm (groups in grid.z): the attention score and context products."""

    @staticmethod
    def _run(a, b):
        a, b = _kmajor(a), _kmajor(b)
        G, M, K = a.shape
        N = b.shape[1]
        out = torch.empty(G, M, N, device=a.device, dtype=F32)
        ops.gemm(a, b, M=M, N=N, K=K, G1=G, a_c2=(1, 0), b_c2=(1, 0, 0), out_f32=out, out_ld=N, out_g=(M * N, 0))
        return out

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.float(), b.float()
        ctx.save_for_backward(a, b)
        return _BmmNT._run(a, b)

    @staticmethod
    def backward(ctx, dc):
        a, b = ctx.saved_tensors
        dc = dc.float()
        da = _BmmNT._run(dc, b.transpose(1, 2)) if ctx.needs_input_grad[0] else None                  # dA = dC B
        db = _BmmNT._run(dc.transpose(1, 2), a.transpose(1, 2)) if ctx.needs_input_grad[1] else None  # dB = dC^T A
        return da, db


def bmm_nt(a, b):
    return _BmmNT.apply(a, b)


# ----------------------------------------------------------------------------------------------------
# the DiT in train mode (reference ldm/modules/diffusionmodules/vocal2music_moe.py, flag_large_dit_moe.py)
# ----------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps):                      # flag_large_dit_moe.py:52-77
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight


def modulate(x, shift, scale):                     # flag_large_dit_moe.py:80-81
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _rope(x, ang):                                 # flag_large_dit_moe.py:237-269, x [B, T, H, hd]
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    cis = torch.polar(torch.ones_like(ang), ang)[None, :, None, :]
    return torch.view_as_real(xc * cis).flatten(3)


def _attend(q, k, v):
    """softmax(q k^T / sqrt(d)) v for [B, H, Tq, d] x [B, H, Tk, d]; both products on the tensor cores."""
    B, H, Tq, d = q.shape
    Tk = k.shape[2]
    s = bmm_nt(q.reshape(B * H, Tq, d), k.reshape(B * H, Tk, d)) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    o = bmm_nt(p, v.reshape(B * H, Tk, d).transpose(1, 2))   # P [Tq, Tk] x (V^T [d, Tk])^T
    return o.view(B, H, Tq, d)


def _attention(P, pre, x, y, ang, H):              # flag_large_dit_moe.py:323-406
    B, T, D = x.shape
    hd = D // H
    # q|k|v (and the text k|v) as one stacked projection each: one forward GEMM, one dX, one dW
    qkv = linear(x, torch.cat([P[pre + "wq.weight"], P[pre + "wk.weight"], P[pre + "wv.weight"]], 0)).view(B, T, 3, H, hd)
    q = _rope(qkv[:, :, 0], ang).permute(0, 2, 1, 3)
    k = _rope(qkv[:, :, 1], ang).permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    out = _attend(q, k, v)
    ykv = linear(y, torch.cat([P[pre + "wk_y.weight"], P[pre + "wv_y.weight"]], 0)).view(B, -1, 2, H, hd)
    yk, yv = ykv[:, :, 0].permute(0, 2, 1, 3), ykv[:, :, 1].permute(0, 2, 1, 3)
    out = out + _attend(q, yk, yv) * P[pre + "gate"].tanh().view(1, -1, 1, 1)
    return linear(out.permute(0, 2, 1, 3).reshape(B, T, D), P[pre + "wo.weight"])


def _ffn(P, pre, x):                               # flag_large_dit_moe.py:480-485
    return linear(F.silu(linear(x, P[pre + "w1.weight"])) * linear(x, P[pre + "w3.weight"]), P[pre + "w2.weight"])


def _gumbel_soft(logits, tau, expo):               # vocal2music_moe.py:81-93, hard=False
    return ((logits - expo.log()) / tau).softmax(dim=-1)


def _moe(P, pre, x, time, caption, acoustic, E, tau, expo, H=8):   # vocal2music_moe.py:117-185, training branch
    B, T, D = x.shape
    hd = D // H
    w, b = P[pre + "cross_attention.in_proj_weight"], P[pre + "cross_attention.in_proj_bias"]
    L = caption.shape[1]
    q = linear(x, w[:D], b[:D]).view(B, T, H, hd).permute(0, 2, 1, 3)
    kv = linear(caption, w[D:], b[D:]).view(B, L, 2, H, hd)
    k, v = kv[:, :, 0].permute(0, 2, 1, 3), kv[:, :, 1].permute(0, 2, 1, 3)
    cap_feat = linear(_attend(q, k, v).permute(0, 2, 1, 3).reshape(B, T, D),
                      P[pre + "cross_attention.out_proj.weight"], P[pre + "cross_attention.out_proj.bias"])
    xf, cap_flat, ac_flat = x.reshape(-1, D), cap_feat.reshape(-1, D), acoustic.reshape(-1, D)
    e0, e1, e2 = expo
    hl = linear(time, P[pre + "high_level_gating_network.weight"], P[pre + "high_level_gating_network.bias"])
    hl_p = _gumbel_soft(hl.repeat_interleave(T, dim=0), 1.0, e0)
    cap_mask, ac_mask = hl_p[:, 0:1], hl_p[:, 1:2]
    cap_p = _gumbel_soft(linear(cap_flat, P[pre + "caption_gating_network.weight"], P[pre + "caption_gating_network.bias"]), tau, e1)
    ac_p = _gumbel_soft(linear(ac_flat, P[pre + "acoustic_gating_network.weight"], P[pre + "acoustic_gating_network.bias"]), tau, e2)
    # The 2E routed experts run densely on every token in train mode (:157-165). Stacked: ONE GEMM for all w1|w3
    # (N = 2 * 2E * Fh), and — since sum_e p_e (h_e W2_e^T) = [p_e h_e]_e [W2_e]_e^T — ONE GEMM with K = 2E * Fh for all
    # w2 with the soft gates applied to the hidden activations: 2 large launches (and 2 x 2 in the backward pass,
    # sharing one transposed copy of x) instead of 24 small ones whose dW products (K = all tokens) each filled
    # a sixth of the machine.
    names = [f"{pre}caption_experts.{i}." for i in range(E)] + [f"{pre}acoustic_experts.{i}." for i in range(E)]
    Fh = P[names[0] + "w1.weight"].shape[0]
    w13 = torch.cat([P[n + "w1.weight"] for n in names] + [P[n + "w3.weight"] for n in names], 0)   # [2 * 2E * Fh, D]
    h13 = linear(xf, w13)
    gate_w = torch.cat([cap_p * cap_mask, ac_p * ac_mask], dim=1)                                   # [N, 2E]
    hid = (F.silu(h13[:, : 2 * E * Fh]) * h13[:, 2 * E * Fh:]).view(-1, 2 * E, Fh) * gate_w.unsqueeze(-1)
    w2 = torch.cat([P[n + "w2.weight"] for n in names], 1)                                          # [D, 2E * Fh]
    yb = linear(hid.reshape(-1, 2 * E * Fh), w2).view(B, T, D)
    band = D // E
    # frequency experts: expert i sees only its band of y and keeps only its band of the output (:171-178) =
    # batched GEMMs over the band slices of the weights (the masked-out columns / rows get exactly zero gradient)
    fn = [f"{pre}freq_experts.{i}." for i in range(E)]
    f13 = torch.stack([torch.cat([P[n + "w1.weight"][:, band * i: band * (i + 1)], P[n + "w3.weight"][:, band * i: band * (i + 1)]], 0)
                       for i, n in enumerate(fn)], 0)                                               # [E, 2 Fh, band]
    yband = yb.reshape(-1, E, band).permute(1, 0, 2)                                                # [E, N, band]
    hf13 = bmm_nt(yband, f13)                                                                       # [E, N, 2 Fh]
    hf = F.silu(hf13[..., :Fh]) * hf13[..., Fh:]
    f2 = torch.stack([P[n + "w2.weight"][band * i: band * (i + 1)] for i, n in enumerate(fn)], 0)   # [E, band, Fh]
    z = bmm_nt(hf, f2).permute(1, 0, 2).reshape(B, T, D)                                            # [E, N, band] -> [B, T, D]
    all_p = torch.cat([cap_p, ac_p], dim=1)        # load-balancing loss, :95-115
    all_m = torch.cat([cap_mask.repeat(1, E), ac_mask.repeat(1, E)], dim=1)
    usage = (all_p * all_m).sum(dim=0) / (all_m.sum() + 1e-10)
    return z, torch.mean(usage * torch.log(usage + 1e-10))


def dit_train_forward(P: Dict[str, torch.Tensor], x, t, context, expo, *, depth: int, n_heads: int, num_experts: int = 4,
                      temperature: float = 2.0, loss_w: float = 1.0, norm_eps: float = 1e-5):
    """TxtFlagLargeDiT.forward in train mode (vocal2music_moe.py:375-434). P: parameter dict with the reference's
    state_dict keys; expo[k] = the three Exp(1) draws of block k. Returns (v [B, C, T], lb_loss * loss_w')."""
    midi, beats = context["c_concat"]["midi"], context["c_concat"]["beats"]
    D = P["proj_in.weight"].shape[0]

    def stem(emb, proj, idx):                      # :336-350,388-391
        e = F.embedding(idx.squeeze(1), P[emb]).transpose(1, 2)
        h = conv1d_same(e, P[proj + "0.weight"], P[proj + "0.bias"])
        return F.avg_pool1d(F.leaky_relu(h, 0.01).transpose(1, 2), 2).transpose(1, 2)

    acoustic = stem("midi_embedding.weight", "midi_proj.", midi) + stem("beats_embedding.weight", "beats_proj.", beats)
    acoustic = linear(acoustic, P["final_proj.weight"][:, :, 0], P["final_proj.bias"])
    cap = context["c_crossattn"]
    hcap = linear(F.gelu(linear(cap, P["c_embedder.mlp.0.weight"], P["c_embedder.mlp.0.bias"])),
                  P["c_embedder.mlp.2.weight"], P["c_embedder.mlp.2.bias"])
    caption = F.layer_norm(hcap, (D,), P["c_embedder.mlp.3.weight"], P["c_embedder.mlp.3.bias"], 1e-5)
    pool = caption.sum(dim=1) / caption.shape[1]
    cap_emb = linear(F.layer_norm(pool, (D,), P["cap_embedder.0.weight"], P["cap_embedder.0.bias"], 1e-5),
                     P["cap_embedder.1.weight"], P["cap_embedder.1.bias"])
    h = conv1d_same(x, P["proj_in.weight"], P["proj_in.bias"])
    if abs(h.shape[1] - acoustic.shape[1]) <= 2:   # :397-401
        if h.shape[1] > acoustic.shape[1]:
            acoustic = torch.cat([acoustic, acoustic[:, -1:, :].repeat(1, h.shape[1] - acoustic.shape[1], 1)], dim=1)
        else:
            acoustic = acoustic[:, : h.shape[1], :]
    half = 128                                     # TimestepEmbedder, flag_large_dit_moe.py:109-133
    freqs = torch.exp(-math.log(10000.0) * torch.arange(0, half, dtype=F32, device=x.device) / half)
    args = t[:, None].float() * freqs[None]
    t_emb = linear(F.silu(linear(torch.cat([torch.cos(args), torch.sin(args)], dim=-1),
                                 P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"])),
                   P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"])
    h = acoustic + h
    adaln_input = t_emb + cap_emb
    hd = D // n_heads
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=x.device)[: hd // 2].float() / hd))
    ang = torch.outer(torch.arange(h.shape[1], device=x.device, dtype=F32), inv)
    loss = 0
    for k in range(depth):                         # TransformerBlock.forward, :215-266
        pre = f"blocks.{k}."
        mod = linear(F.silu(adaln_input), P[pre + "adaLN_modulation.1.weight"], P[pre + "adaLN_modulation.1.bias"])
        sh_msa, sc_msa, g_msa, sh_mlp, sc_mlp, g_mlp = mod.chunk(6, dim=1)
        y = rms_norm(caption, P[pre + "attention_y_norm.weight"], norm_eps)
        h = h + g_msa.unsqueeze(1) * _attention(P, pre + "attention.",
                                                modulate(rms_norm(h, P[pre + "attention_norm.weight"], norm_eps), sh_msa, sc_msa),
                                                y, ang, n_heads)
        ffn_in = modulate(rms_norm(h, P[pre + "ffn_norm.weight"], norm_eps), sh_mlp, sc_mlp)
        out, lb = _moe(P, pre + "feed_forward.", ffn_in, t_emb, caption, acoustic, num_experts, temperature, expo[k], n_heads)
        h = h + g_mlp.unsqueeze(1) * out
        loss = loss + lb
    loss = loss / depth
    if loss_w > 0.01:                              # :427-429
        loss_w = loss_w * 0.9999
    loss = loss * loss_w
    mod = linear(F.silu(adaln_input), P["final_layer.adaLN_modulation.1.weight"], P["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    o = modulate(F.layer_norm(h, (D,), None, None, 1e-6), shift, scale)
    v = linear(o, P["final_layer.linear.weight"], P["final_layer.linear.bias"]).transpose(1, 2)
    return v, loss


# ----------------------------------------------------------------------------------------------------
# flow-matching loss and the optimiser step
# ----------------------------------------------------------------------------------------------------
def cfm_p_losses(forward_fn, x1: torch.Tensor, t: torch.Tensor, x0: torch.Tensor, *, sigma_min: float = 1e-4,
                 num_timesteps: int = 1000, l_simple_weight: float = 1.0):
    """CFM.p_losses, cfm1_audio.py:38-63. forward_fn(x_noisy, t) -> (v, lb_loss)."""
    ut = x1 - (1 - sigma_min) * x0
    tt = t.view(-1, 1, 1).float() / num_timesteps
    x_noisy = tt * x1 + (1.0 - (1 - sigma_min) * tt) * x0
    v, lb = forward_fn(x_noisy, t)
    loss_simple = ((v - ut) ** 2).mean(dim=(1, 2))
    return l_simple_weight * loss_simple.mean() + lb, {"loss_simple": loss_simple.mean().detach(), "lb_loss": lb.detach()}


class FusedAdamW:
    """AdamW with decoupled weight decay; the update of every parameter tensor is one vb_adamw launch."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.m = [torch.zeros_like(p, dtype=F32) for p in self.params]
        self.v = [torch.zeros_like(p, dtype=F32) for p in self.params]
        self.t = 0

    @torch.no_grad()
    def step(self):
        self.t += 1
        c1, c2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is not None:
                ops.adamw(p.data, p.grad, m, v, lr=self.lr, b1=self.b1, b2=self.b2, eps=self.eps, wd=self.wd, c1=c1, c2=c2)

    def zero_grad(self):
        for p in self.params:
            p.grad = None


def draw_router_noise(depth: int, N: int, E: int, device, generator=None) -> List[Tuple[torch.Tensor, ...]]:
    """The Exp(1) draws of one forward in the reference's order (vocal2music_moe.py:134,150,151)."""
    return [tuple(torch.empty(N, k, device=device, dtype=F32).exponential_(generator=generator) for k in (2, E, E))
            for _ in range(depth)]
