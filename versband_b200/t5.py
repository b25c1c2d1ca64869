"""FLAN-T5 (T5 v1.1) encoder on B200 — the conditioning step in front of the hot path:
`FrozenTextVocalEmbedder.forward` of the reference (ldm/modules/encoders/modules.py:194-233) runs
`T5EncoderModel(input_ids).last_hidden_state` over max_length = 80 tokens with NO attention mask (padding
tokens attend and are attended, :215-218), twice per clip (caption / unconditional caption).

Restated from the T5 architecture as implemented by the reference's dependency (transformers `T5Stack`, encoder
only, version unpinned in requirements.txt — parity is pinned against the live `transformers` package installed
in this image, tests/test_t5_gpu.py):
  x = shared[ids]; per block: x += o( softmax(q k^T + rel_bias) v )  with q,k,v = W(rms(x)·w), no 1/sqrt(d) scale,
  rel_bias[h,i,j] = table[bucket(j - i), h] (bidirectional buckets, block 0's table shared by all blocks);
  x += wo( gelu_new(wi_0 rms(x)) * wi_1 rms(x) ); out = rms(x) · w_final.          RMS eps = layer_norm_epsilon.
Every matmul is a vb_gemm launch on fp16 operands (fp32 accumulate / residual / norms / softmax).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import ops
from .dit import F32, H16


def relative_position_bucket(rel: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """Bidirectional T5 bucketing of relative positions (memory - query)."""
    nb = num_buckets // 2
    ret = (rel > 0).long() * nb
    n = rel.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float().clamp_min(1) / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


class T5EncoderEngine:
    """Packed weights of a T5 v1.1 encoder (HF `T5EncoderModel.state_dict()` keys) + the launch sequence."""

    def __init__(self, sd: Dict[str, torch.Tensor], *, num_layers: int, num_heads: int, d_kv: int = 64,
                 num_buckets: int = 32, max_distance: int = 128, eps: float = 1e-6, device="cuda"):
        self.dev = torch.device(device)
        self.L, self.H, self.hd, self.eps = num_layers, num_heads, d_kv, eps
        self.nb, self.maxd = num_buckets, max_distance
        g = lambda k: sd[k].detach().to(self.dev, F32).contiguous()  # noqa: E731
        h = lambda t: t.to(H16).contiguous()  # noqa: E731
        emb_key = "shared.weight" if "shared.weight" in sd else "encoder.embed_tokens.weight"
        self.emb = g(emb_key)
        self.D = self.emb.shape[1]
        assert d_kv in (64, 96) and (num_heads * d_kv) % 64 == 0 and self.D % 64 == 0
        self.blocks = []
        for i in range(num_layers):
            p = f"encoder.block.{i}.layer."
            a = p + "0.SelfAttention."
            wi0, wi1 = g(p + "1.DenseReluDense.wi_0.weight"), g(p + "1.DenseReluDense.wi_1.weight")
            self.blocks.append(dict(
                ln1=g(p + "0.layer_norm.weight"), ln2=g(p + "1.layer_norm.weight"),
                wqkv=h(torch.cat([g(a + "q.weight"), g(a + "k.weight"), g(a + "v.weight")], 0)),
                wo=h(g(a + "o.weight")),
                wi=h(torch.stack([wi0, wi1], 1).reshape(2 * wi0.shape[0], self.D)),  # (wi_0, wi_1) row pairs: GEGLU epilogue
                wout=h(g(p + "1.DenseReluDense.wo.weight"))))
        self.rel_table = g("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight")  # [buckets, H]
        self.ln_f = g("encoder.final_layer_norm.weight")
        self.d_ff = self.blocks[0]["wout"].shape[1]
        self._bias: Dict[tuple, torch.Tensor] = {}

    def _pos_bias(self, B: int, T: int, Tp: int) -> torch.Tensor:
        key = (B, T)
        if key not in self._bias:
            ctx = torch.arange(T, device=self.dev)
            bucket = relative_position_bucket(ctx[None, :] - ctx[:, None], self.nb, self.maxd)  # [T(query), T(key)]
            bias = torch.zeros(B, self.H, T, Tp, device=self.dev, dtype=F32)
            bias[..., :T] = self.rel_table[bucket].permute(2, 0, 1)[None]
            self._bias[key] = bias.contiguous()
        return self._bias[key]

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        """input_ids int64 [B, T] -> last_hidden_state fp32 [B, T, D]."""
        dev, D, H, hd = self.dev, self.D, self.H, self.hd
        B, T = input_ids.shape
        N, Tp, I = B * T, (T + 7) // 8 * 8, H * hd
        x = torch.empty(N, D, device=dev, dtype=F32)
        ops.embed_rows(input_ids.to(dev).contiguous(), self.emb, x)
        xn = torch.empty(N, D, device=dev, dtype=H16)
        q = torch.empty(B, H, T, hd, device=dev, dtype=H16)
        k = torch.empty_like(q)
        vt = torch.zeros(B, H, hd, Tp, device=dev, dtype=H16)
        S = torch.empty(B, H, T, Tp, device=dev, dtype=F32)
        P = torch.empty(B, H, T, Tp, device=dev, dtype=H16)
        o = torch.empty(N, I, device=dev, dtype=H16)
        hff = torch.empty(N, self.d_ff, device=dev, dtype=H16)
        bias = self._pos_bias(B, T, Tp)
        for blk in self.blocks:
            ops.norm_mod(x, kind=0, w=blk["ln1"], eps=self.eps, out_fp16=xn)
            ops.gemm(xn, blk["wqkv"], mode=ops.EPI_QKV_ROPE, T=T, heads=H, head_dim=hd, Tp=Tp, q_out=q, k_out=k, vt_out=vt)
            # scores = q k^T + relative position bias (no scaling in T5), batched over (head, batch)
            ops.gemm(q, k.view(B * H, T, hd), M=T, N=T, K=hd, G1=H, G2=B, a_c2=(1, 0), a_c3=(0, 1), b_c2=(1, H, 0),
                     residual=bias, out_f32=S, out_ld=Tp, out_g=(T * Tp, H * T * Tp))
            ops.softmax_rows(S, Tp, P, Tp, B * H * T, T)
            # o[b*T + t, h*hd : (h+1)*hd] = P[b,h] v[b,h]
            ops.gemm(P[..., :T], vt.view(B * H, hd, Tp)[..., :T], M=T, N=hd, K=T, G1=H, G2=B, a_c2=(1, 0), a_c3=(0, 1),
                     b_c2=(1, H, 0),
                     out_fp16=o, outb_ld=I, outb_g=(hd, T * I))
            ops.gemm(o, blk["wo"], residual=x, out_f32=x, out_ld=D)
            ops.norm_mod(x, kind=0, w=blk["ln2"], eps=self.eps, out_fp16=xn)
            ops.gemm(xn, blk["wi"], swiglu=2, out_fp16=hff, outb_ld=self.d_ff)  # gelu_new(wi_0 x) * (wi_1 x)
            ops.gemm(hff, blk["wout"], residual=x, out_f32=x, out_ld=D)
        out = torch.empty(B, T, D, device=dev, dtype=F32)
        ops.norm_mod(x, kind=0, w=self.ln_f, eps=self.eps, out_f32=out.view(N, D))
        return out
