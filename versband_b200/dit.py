"""AccompBand DiT on B200: weight packing, step-invariant conditioning cache and the per-ODE-step
forward, expressed as a sequence of libversband_sm100 kernel launches (versband_b200.ops).

Mirrors `TxtFlagLargeDiT.forward` (reference ldm/modules/diffusionmodules/vocal2music_moe.py:375-434)
with the restructuring SURVEY.md §0.4/§0.8 shows to be value-equivalent:
  * routed experts evaluated sparsely (hard top-1 ⇒ non-selected weights are exactly 0, :87-89),
  * frequency experts restricted to their 192-channel band (:171-178),
  * everything that depends only on the conditioning computed once per clip (`prepare`).
All tensors live in HBM as: residual stream fp32 [N, D]; GEMM operands fp16; N = B*T tokens.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops

H16 = torch.float16
F32 = torch.float32


def _h16(t: torch.Tensor) -> torch.Tensor:
    return t.to(H16).contiguous()


def tf32_round(t: torch.Tensor) -> torch.Tensor:
    """Round fp32 to the TF32 grid (10-bit mantissa, round-to-nearest) so that the tensor core's
    operand truncation is exact."""
    ti = t.contiguous().view(torch.int32)
    return ((ti + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_conv_weight(w: torch.Tensor, cin_pad: Optional[int] = None, dtype=H16) -> torch.Tensor:
    """Conv1d weight [Cout, Cin, k] -> implicit-GEMM B operand [Cout, k * kpt*kq] (tap-major);
    kq = 64 (fp16) or 32 (fp32/tf32) elements = one 128-byte k-block."""
    cout, cin, k = w.shape
    kq = 64 if dtype == H16 else 32
    kp = (cin + kq - 1) // kq * kq if cin_pad is None else cin_pad
    out = torch.zeros(cout, k, kp, device=w.device, dtype=dtype)
    src = w.permute(0, 2, 1)
    out[:, :, :cin] = src.to(H16) if dtype == H16 else tf32_round(src.float())
    return out.reshape(cout, k * kp).contiguous()


@dataclass
class DiTCond:
    """Step-invariant tensors of one batch of clips for one CFG branch."""
    B: int
    T: int
    L: int
    acoustic: torch.Tensor            # fp32 [N, D]   (midi+beats stem, shared by both CFG branches)
    ac_logits: torch.Tensor           # fp32 [N, depth*E]
    caption: torch.Tensor             # fp32 [B, L, D]  (c_embedder output)
    cap_emb: torch.Tensor             # fp32 [B, D]
    # per block, fp16: text cross-attention K [B,H,L,hd] and V^T [B,H,hd,Lp]; same for the MoE MHA
    k_y: List[torch.Tensor] = field(default_factory=list)
    vt_y: List[torch.Tensor] = field(default_factory=list)
    # per block, the MoE caption cross-attention folded into its gate (include/versband.h: VB_EPI_ROUTE):
    # wk_m [B, H*96, D] operand dtype (scale * K_h W_q,h), bs_m fp32 [B, H, 96], vg_m fp32 [B, H, 96, E]
    wk_m: List[torch.Tensor] = field(default_factory=list)
    bs_m: List[torch.Tensor] = field(default_factory=list)
    vg_m: List[torch.Tensor] = field(default_factory=list)

    _LISTS = ("k_y", "vt_y", "wk_m", "bs_m", "vg_m")

    def tensors(self) -> List[torch.Tensor]:
        return [self.acoustic, self.ac_logits, self.caption, self.cap_emb,
                *[t for name in self._LISTS for t in getattr(self, name)]]

    @staticmethod
    def cat(conds: "List[DiTCond]") -> "DiTCond":
        """Stack conditioning branches along the batch axis (cond first, then uncond)."""
        c0 = conds[0]
        D = c0.acoustic.shape[-1]
        out = DiTCond(B=sum(c.B for c in conds), T=c0.T, L=c0.L,
                      acoustic=torch.cat([c.acoustic for c in conds], 0),
                      ac_logits=torch.cat([c.ac_logits for c in conds], 0),
                      caption=torch.cat([c.caption for c in conds], 0),
                      cap_emb=torch.cat([c.cap_emb for c in conds], 0))
        for name in DiTCond._LISTS:
            for k in range(len(c0.k_y)):
                getattr(out, name).append(torch.cat([getattr(c, name)[k] for c in conds], 0))
        return out

    def split(self, parts: int) -> "List[DiTCond]":
        """The inverse of `cat`: `parts` conditioning objects over equal slices of the batch (views, no copies)."""
        assert self.B % parts == 0
        b = self.B // parts
        out = []
        for i in range(parts):
            rows = slice(i * b * self.T, (i + 1) * b * self.T)
            c = DiTCond(B=b, T=self.T, L=self.L, acoustic=self.acoustic[rows], ac_logits=self.ac_logits[rows],
                        caption=self.caption[i * b:(i + 1) * b], cap_emb=self.cap_emb[i * b:(i + 1) * b])
            for name in DiTCond._LISTS:
                for t in getattr(self, name):
                    getattr(c, name).append(t[i * b:(i + 1) * b])
            out.append(c)
        return out

    def copy_from(self, other: "DiTCond") -> None:
        for dst, src in zip(self.tensors(), other.tensors()):
            dst.copy_(src)


class DiTEngine:
    """Holds the packed weights of one TxtFlagLargeImprovedDiTV2 and runs it on the GPU."""

    def __init__(self, sd: Dict[str, torch.Tensor], *, in_channels=20, hidden_size=768, depth=4,
                 num_heads=8, num_experts=4, ori_dim=1024, max_len=1500, norm_eps=1e-5,
                 temperature=2.0, device="cuda", precision: str = "fp16"):
        """precision: "fp16" (default) — IEEE fp16 GEMM operands (UMMA kind::f16, saturating
        conversion), fp32 accumulation / residual stream / norms / softmax / router: the 11-bit significand of
        TF32 at half the operand bytes, DiT output within 1e-3 (rel. L2) of the fp32 reference;
        "tf32" — fp32 activations/weights multiplied as TF32 (kind::tf32) in every linear layer (same accuracy,
        twice the operand traffic; for checkpoints whose normalised activations exceed the fp16 range)."""
        assert precision in ("fp16", "tf32")
        self.precision = precision
        self.adt = H16 if precision == "fp16" else F32      # dtype of GEMM operands
        self.kq = 64 if precision == "fp16" else 32        # elements per 128-byte k-block
        self.C, self.D, self.depth, self.H, self.E = in_channels, hidden_size, depth, num_heads, num_experts
        self.hd = hidden_size // num_heads
        self.LK = 96  # key slots per head of the folded caption cross-attention (VB_EPI_ROUTE tile width)
        self.ori_dim, self.max_len, self.eps = ori_dim, max_len, norm_eps
        self.temperature = temperature
        self.dev = torch.device(device)
        assert self.hd % 32 == 0 and hidden_size % (64 * 1) == 0
        self._ws: Dict[Tuple[int, int], dict] = {}
        # MoE expert GEMMs as one persistent chained launch (vb_gemm_chain). Off by default: measured on B200
        # (tools/bench_chain.py) the 4 separately tiled launches take 38 us per block, the chain 54 us — the
        # dependency stalls of a statically scheduled tile list cost more than the 3 launch boundaries they remove
        self.use_chain = False
        self.exact_stem = True  # acoustic stem as exact table gathers + split-precision final_proj (vb_stem_gather)
        self.chain_bn = 128
        # router + dispatch in one launch over fixed-capacity buckets (vb_route_dispatch). Off by default: measured
        # 487 vs 473 us per evaluation (tools/bench_eval.py) — the last-CTA finalisation serialises ~3 us that the
        # two-launch version overlaps under PDL
        self.fused_dispatch = False
        self.pack(sd)

    # ------------------------------------------------------------------ packing
    def pack(self, sd: Dict[str, torch.Tensor]) -> None:
        dev, D, E = self.dev, self.D, self.E
        g = lambda k: sd[k].detach().to(dev, F32).contiguous()  # noqa: E731
        adt = self.adt
        _op = (lambda t: t.to(H16).contiguous()) if adt == H16 else (lambda t: tf32_round(t.float()))  # noqa: E731
        pack_conv = lambda w_: pack_conv_weight(w_, dtype=adt)  # noqa: E731
        w: dict = {}
        # stem
        w["midi_emb"], w["beats_emb"] = g("midi_embedding.weight"), g("beats_embedding.weight")
        for name in ("midi_proj", "beats_proj"):
            w[name + "_w"] = pack_conv(g(name + ".0.weight"))
            w[name + "_b"] = g(name + ".0.bias")
        w["final_proj_w"] = _op(g("final_proj.weight")[:, :, 0])
        w["final_proj_b"] = g("final_proj.bias")
        # exact-table stem (vb_stem_gather): tab[tap][v] = W[:, :, tap] . emb[v] in fp64, and final_proj as the
        # split operand [W_hi | W_lo | W_hi] matching the activations' [hi | hi | lo]
        for name in ("midi", "beats"):
            cw = sd[name + "_proj.0.weight"].detach().to(dev, torch.float64)          # [D, D, k]
            em = sd[name + "_embedding.weight"].detach().to(dev, torch.float64)       # [V, D]
            w[name + "_tab"] = torch.einsum("oik,vi->kvo", cw, em).float().contiguous()  # [k, V, D]
        fp = g("final_proj.weight")[:, :, 0]
        fp_hi = _op(fp)
        fp_lo = _op(fp - fp_hi.float())
        w["final_proj_w3"] = torch.cat([fp_hi, fp_lo, fp_hi], 1).contiguous()
        w["proj_in_w"] = pack_conv(g("proj_in.weight"))
        w["proj_in_b"] = g("proj_in.bias")
        self.proj_in_k = sd["proj_in.weight"].shape[-1]
        self.stem_k = sd["midi_proj.0.weight"].shape[-1]
        # embedders
        for i in (0, 2):
            w[f"t_w{i}"], w[f"t_b{i}"] = _op(g(f"t_embedder.mlp.{i}.weight")), g(f"t_embedder.mlp.{i}.bias")
            w[f"c_w{i}"], w[f"c_b{i}"] = _op(g(f"c_embedder.mlp.{i}.weight")), g(f"c_embedder.mlp.{i}.bias")
        w["c_ln_w"], w["c_ln_b"] = g("c_embedder.mlp.3.weight"), g("c_embedder.mlp.3.bias")
        w["cap_ln_w"], w["cap_ln_b"] = g("cap_embedder.0.weight"), g("cap_embedder.0.bias")
        w["cap_w"], w["cap_b"] = _op(g("cap_embedder.1.weight")), g("cap_embedder.1.bias")
        half = 128
        w["t_freqs"] = torch.exp(-math.log(10000.0) * torch.arange(0, half, dtype=F32) / half).to(dev)
        # adaLN of all blocks + final layer as one [depth*6D + 2D, D] operand
        ada_w = [g(f"blocks.{k}.adaLN_modulation.1.weight") for k in range(self.depth)]
        ada_b = [g(f"blocks.{k}.adaLN_modulation.1.bias") for k in range(self.depth)]
        ada_w.append(g("final_layer.adaLN_modulation.1.weight"))
        ada_b.append(g("final_layer.adaLN_modulation.1.bias"))
        w["ada_w"], w["ada_b"] = _op(torch.cat(ada_w, 0)), torch.cat(ada_b, 0).contiguous()
        self.mod_ld = w["ada_w"].shape[0]
        # gating networks stacked over blocks
        w["hl_w"] = torch.cat([g(f"blocks.{k}.feed_forward.high_level_gating_network.weight") for k in range(self.depth)], 0).contiguous()
        w["hl_b"] = torch.cat([g(f"blocks.{k}.feed_forward.high_level_gating_network.bias") for k in range(self.depth)], 0).contiguous()
        w["acg_w"] = torch.cat([g(f"blocks.{k}.feed_forward.acoustic_gating_network.weight") for k in range(self.depth)], 0).contiguous()
        w["acg_b"] = torch.cat([g(f"blocks.{k}.feed_forward.acoustic_gating_network.bias") for k in range(self.depth)], 0).contiguous()
        band = D // E
        blocks = []
        for k in range(self.depth):
            p = f"blocks.{k}."
            a, f = p + "attention.", p + "feed_forward."
            b: dict = {}
            b["wqkv"] = _op(torch.cat([g(a + "wq.weight"), g(a + "wk.weight"), g(a + "wv.weight")], 0))
            b["wo"] = _op(g(a + "wo.weight"))
            b["wkvy"] = _op(torch.cat([g(a + "wk_y.weight"), g(a + "wv_y.weight")], 0))
            b["gate"] = g(a + "gate")
            b["attn_norm"], b["ffn_norm"], b["y_norm"] = g(p + "attention_norm.weight"), g(p + "ffn_norm.weight"), g(p + "attention_y_norm.weight")
            ipw, ipb = g(f + "cross_attention.in_proj_weight"), g(f + "cross_attention.in_proj_bias")
            b["mkv_w"], b["mkv_b"] = _op(ipw[D:]), ipb[D:].contiguous()
            # caption gate folded through the MHA out-projection (its only consumer, vocal2music_moe.py:122,141):
            # logits = Wg (Wo ctx + bo) + bg = (Wg Wo) ctx + (Wg bo + bg), all in fp32
            wo_m, bo_m = g(f + "cross_attention.out_proj.weight"), g(f + "cross_attention.out_proj.bias")
            wg, bg = g(f + "caption_gating_network.weight"), g(f + "caption_gating_network.bias")
            b["capg_w"] = (wg.double() @ wo_m.double()).float().contiguous()
            b["capg_b"] = (wg.double() @ bo_m.double() + bg.double()).float().contiguous()
            # ... and the attention itself folded into one GEMM per step: scores = xm . (scale K_h W_q,h)^T.
            # mq_wT[h] = (scale W_q,h)^T  [D_in, hd (padded to the k-block)]: B operand of the per-clip WK GEMM
            sc = 1.0 / math.sqrt(self.hd)
            kpad = (self.hd + self.kq - 1) // self.kq * self.kq
            wqT = torch.zeros(self.H, D, kpad, device=dev)
            wqT[:, :, : self.hd] = (ipw[:D] * sc).view(self.H, self.hd, D).permute(0, 2, 1)
            b["mq_wT"] = _op(wqT)
            b["mq_bs"] = (ipb[:D] * sc).contiguous()
            w13, w2 = [], []
            for grp in ("caption_experts", "acoustic_experts"):
                for e in range(E):
                    w1, w3 = g(f"{f}{grp}.{e}.w1.weight"), g(f"{f}{grp}.{e}.w3.weight")
                    w13.append(torch.stack([w1, w3], 1).reshape(2 * w1.shape[0], D))
                    w2.append(g(f"{f}{grp}.{e}.w2.weight"))
            b["w13"], b["w2"] = _op(torch.stack(w13, 0)), _op(torch.stack(w2, 0))  # [2E,2Fh,D], [2E,D,Fh]
            f13, f2 = [], []
            for e in range(E):
                w1, w3 = g(f"{f}freq_experts.{e}.w1.weight"), g(f"{f}freq_experts.{e}.w3.weight")
                sl = slice(band * e, band * (e + 1))
                f13.append(torch.stack([w1[:, sl], w3[:, sl]], 1).reshape(2 * w1.shape[0], band))
                f2.append(g(f"{f}freq_experts.{e}.w2.weight")[sl, :])
            f13w = torch.stack(f13, 0)
            b["f13"], b["f2"] = _op(torch.cat([f13w, f13w], -1)), _op(torch.stack(f2, 0))  # [E,2Fh,2*band], [E,band,Fh]
            blocks.append(b)
        self.Fh = blocks[0]["w2"].shape[-1]
        self.band = band
        w["blocks"] = blocks
        fl_w = torch.zeros(32, D, device=dev)
        fl_w[: self.C] = g("final_layer.linear.weight")
        fl_b = torch.zeros(32, device=dev)
        fl_b[: self.C] = g("final_layer.linear.bias")
        w["final_w"], w["final_b"] = _op(fl_w), fl_b
        # RoPE table [max_len, hd/2, (cos, sin)] computed exactly like the reference (vocal2music_moe.py:464-475)
        freqs = 1.0 / (10000.0 ** (torch.arange(0, self.hd, 2)[: self.hd // 2].float() / self.hd))
        ang = torch.outer(torch.arange(self.max_len, dtype=F32), freqs).float()
        cis = torch.polar(torch.ones_like(ang), ang)
        w["rope"] = torch.stack([cis.real, cis.imag], -1).contiguous().to(dev)
        self.w = w

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, B: int, T: int, slot: int = 0) -> dict:
        """Per-(B, T) scratch tensors; `slot` distinguishes evaluations that run CONCURRENTLY (the sampler's two
        CFG branches on two streams)."""
        key = (B, T, slot)
        if key in self._ws:
            return self._ws[key]
        dev, D, H, hd, E, Fh = self.dev, self.D, self.H, self.hd, self.E, self.Fh
        N = B * T
        Tp = (T + 7) // 8 * 8
        # fixed-capacity expert buckets (vb_route_dispatch): bucket k owns rows [k*cap, (k+1)*cap)
        cap = (N + 127) // 128 * 128
        max_rows = 2 * E * cap if self.fused_dispatch else ((2 * N + 127) // 128 + 2 * E) * 128
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)  # noqa: E731
        AD = self.adt
        ws = dict(
            Tp=Tp, max_rows=max_rows, cap=cap,
            tile_row0=z(max_rows // 128, dt=torch.int32), done_ticket=z(self.depth, dt=torch.int32),
            x_cl=z(B, T, 32, dt=AD), x=z(N, D), h=z(N, D),
            temb_sin=z(B, 256, dt=AD), temb_h=z(B, D, dt=AD), temb=z(B, D), ada_in=z(B, D, dt=AD),
            mod=z(B, self.mod_ld), hl=z(B, 2 * self.depth),
            xa=z(N, D, dt=AD), q=z(B, H, T, hd, dt=H16), k=z(B, H, T, hd, dt=H16), vt=z(B, H, hd, Tp, dt=H16),
            oa=z(N, D, dt=AD),
            xm=z(N, D, dt=AD), cap_part=z(N, self.H, E),
            idx_c=z(N, dt=torch.int32), idx_a=z(N, dt=torch.int32), scale_c=z(N), scale_a=z(N),
            cap_logits=z(N, E), bucket_count=z(self.depth, 2 * E, dt=torch.int32), usage=z(self.depth, 2 * E), bucket_start=z(2 * E, dt=torch.int32),
            slot_c=z(N, dt=torch.int32), slot_a=z(N, dt=torch.int32),
            row_map=z(max_rows, dt=torch.int32), row_scale=z(max_rows), tile_group=z(max_rows // 128, dt=torch.int32),
            m_tiles=z(1, dt=torch.int32), lb=z(self.depth),
            xs=z(max_rows, D, dt=AD), hs=z(max_rows, Fh, dt=AD), y2=z(2 * N, D, dt=AD),
            hf=z(E, N, Fh, dt=AD), xf=z(N, D, dt=AD), v=z(N, 32),
        )
        # dependency counters of the chained MoE launch, one zeroed row per block (see vb_gemm_chain)
        ws["chain_len"] = 8 + 2 * (max_rows // 128 + 1) + 2 * (E * ((N + 127) // 128) + 1)
        ws["chain_cnt"] = torch.zeros(self.depth, ws["chain_len"], device=dev, dtype=torch.int32)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ step-invariant part
    def _prep_desc(self):
        """The conditioning-side weights as a `vb_dit_prep_desc` (include/versband.h) for vb_dit_prepare."""
        if getattr(self, "_c_prep", None) is not None:
            return self._c_prep
        from ._lib import DitPrepDesc
        w, P = self.w, ops._ptr
        p = DitPrepDesc()
        p.ori_dim, p.stem_k = self.ori_dim, self.stem_k
        p.Vm, p.Vb = w["midi_tab"].shape[1], w["beats_tab"].shape[1]
        for n in ("midi_tab", "beats_tab", "midi_proj_b", "beats_proj_b", "final_proj_w3", "final_proj_b", "acg_w", "acg_b",
                  "c_w0", "c_w2", "cap_w", "c_b0", "c_b2", "c_ln_w", "c_ln_b", "cap_ln_w", "cap_ln_b", "cap_b"):
            setattr(p, n, P(w[n]))
        for blk, b in zip(p.blocks, w["blocks"]):
            for n in ("y_norm", "wkvy", "mkv_w", "mkv_b", "mq_wT", "mq_bs", "capg_w"):
                setattr(blk, n, P(b[n]))
        self._c_prep = p
        return p

    @torch.no_grad()
    def _prepare_c(self, midi: torch.Tensor, beats: torch.Tensor, caption: torch.Tensor, T: int) -> DiTCond:
        """ONE call of the model-level C entry point vb_dit_prepare: the acoustic stem of the Bm clips is computed
        once and shared by the B = groups x Bm caption rows (CFG pair: B = 2 Bm, cond rows first)."""
        import ctypes as C
        from . import _lib
        dev, D, H, hd, E = self.dev, self.D, self.H, self.hd, self.E
        Bm, _, Tm = midi.shape
        B, L, _ = caption.shape
        if L > self.LK:
            raise ValueError(f"caption length {L} > {self.LK}: the folded caption gate (VB_EPI_ROUTE) holds {self.LK} key "
                             "slots per head (the reference pads captions to 80 tokens, encoders/modules.py:88-89)")
        assert B % Bm == 0 and abs(T - Tm // 2) <= 2, "latent / conditioning length mismatch"
        midi, beats = midi.to(dev).contiguous(), beats.to(dev).contiguous()
        caption = caption.to(dev, F32).contiguous()
        Lp = (L + 7) // 8 * 8
        e = lambda *s, dt=F32: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        cond = DiTCond(B=B, T=T, L=L, acoustic=e(B * T, D), ac_logits=e(B * T, self.depth * E), caption=e(B, L, D),
                       cap_emb=e(B, D))
        for _ in range(self.depth):
            cond.k_y.append(e(B, H, L, hd, dt=H16))
            cond.vt_y.append(e(B, H, hd, Lp, dt=H16))
            cond.wk_m.append(e(B, H * self.LK, D, dt=self.adt))
            cond.bs_m.append(e(B, H, self.LK))
            cond.vg_m.append(e(B, H, self.LK, E))
        d, p, cs = self._desc(), self._prep_desc(), self._cond_struct(cond)
        need = int(_lib.lib().vb_dit_prepare_workspace_bytes(C.byref(d), C.byref(p), Bm, Tm, B, L, T))
        if need < 0:
            _lib.check(-1, "vb_dit_prepare_workspace_bytes")
        ws = torch.empty(need, device=dev, dtype=torch.uint8)
        P = ops._ptr
        _lib.check(_lib.lib().vb_dit_prepare(C.byref(d), C.byref(p), P(midi), P(beats), Bm, Tm, P(caption), B, L, C.byref(cs),
                                             P(cond.caption), P(ws), need, ops._stream()), "vb_dit_prepare")
        return cond

    @torch.no_grad()
    def prepare_acoustic(self, midi: torch.Tensor, beats: torch.Tensor, T: int):
        """midi/beats int64 [B,1,Tm] -> acoustic fp32 [B*T, D] and the acoustic gate logits
        (vocal2music_moe.py:388-393,397-401,142)."""
        w, D, dev = self.w, self.D, self.dev
        B, _, Tm = midi.shape
        Tpool = Tm // 2
        ac = torch.empty(B, Tpool, D, device=dev, dtype=F32)
        if self.exact_stem:
            # the two k=5 convolutions over embedding rows as table gathers (exact to fp32 rounding), pooled and
            # written as the split operand of final_proj: 2 launches instead of 6, 18.7 GFLOP less per clip
            pooled3 = torch.empty(B * Tpool, 3 * D, device=dev, dtype=self.adt)
            ops.stem_gather(midi.to(dev).contiguous(), beats.to(dev).contiguous(), w["midi_tab"], w["beats_tab"],
                            w["midi_proj_b"], w["beats_proj_b"], k=self.stem_k, Tpool=Tpool, slope=0.01, out=pooled3)
            ops.gemm(pooled3, w["final_proj_w3"], bias=w["final_proj_b"], out_f32=ac, out_ld=D)
        else:
            emb = torch.empty(B, Tm, D, device=dev, dtype=self.adt)
            conv = [torch.empty(B, Tm, D, device=dev, dtype=F32) for _ in range(2)]
            for i, (idx, tab, name) in enumerate(((midi, w["midi_emb"], "midi_proj"), (beats, w["beats_emb"], "beats_proj"))):
                ops.embed_rows(idx.to(dev).contiguous(), tab, emb)
                ops.gemm(emb, w[name + "_w"], M=Tm, G1=B, kpt=D // self.kq, dil=1, pad=self.stem_k // 2, a_c2=(1, 0),
                         bias=w[name + "_b"], act=ops.ACT_LRELU, act_slope=0.01,
                         out_f32=conv[i], out_ld=D, out_g=(Tm * D, 0))
            pooled = torch.empty(B, Tpool, D, device=dev, dtype=self.adt)
            ops.pool2_sum(conv[0], conv[1], B=B, Tin=Tm, Tout=Tpool, D=D, out_fp16=pooled)
            ops.gemm(pooled.view(B * Tpool, D), w["final_proj_w"], bias=w["final_proj_b"], out_f32=ac, out_ld=D)
        if Tpool != T:  # length fix-up, vocal2music_moe.py:397-401
            assert abs(T - Tpool) <= 2, "latent / conditioning length mismatch"
            if T > Tpool:
                ac = torch.cat([ac, ac[:, -1:, :].repeat(1, T - Tpool, 1)], 1).contiguous()
            else:
                ac = ac[:, :T].contiguous()
        ac = ac.view(B * T, D)
        ac_logits = torch.empty(B * T, self.depth * self.E, device=dev, dtype=F32)
        ops.linear_small(ac, w["acg_w"], w["acg_b"], ac_logits)
        return ac, ac_logits

    @torch.no_grad()
    def prepare(self, midi, beats, caption: torch.Tensor, T: int, acoustic=None) -> DiTCond:
        """Everything in TxtFlagLargeDiT.forward that does not depend on (x, t): one vb_dit_prepare call
        (`prepare_py` = the same launches from Python; also used when a precomputed `acoustic` is passed)."""
        if acoustic is None and self.exact_stem and ops.TIMELINE is None:
            return self._prepare_c(midi, beats, caption, T)
        return self.prepare_py(midi, beats, caption, T, acoustic=acoustic)

    @torch.no_grad()
    def prepare_py(self, midi, beats, caption: torch.Tensor, T: int, acoustic=None) -> DiTCond:
        """The launch sequence of `prepare`, op by op from Python."""
        w, D, dev = self.w, self.D, self.dev
        B, L, _ = caption.shape
        if L > self.LK:
            raise ValueError(f"caption length {L} > {self.LK}: the folded caption gate (VB_EPI_ROUTE) holds {self.LK} key "
                             "slots per head (the reference pads captions to 80 tokens, encoders/modules.py:88-89)")
        if acoustic is None:
            acoustic = self.prepare_acoustic(midi, beats, T)
        ac, ac_logits = acoustic
        cap16 = torch.empty(B * L, self.ori_dim, device=dev, dtype=self.adt)
        ops.ew(caption.to(dev, F32).contiguous().view(B * L, -1), out_fp16=cap16)
        h1 = torch.empty(B * L, D, device=dev, dtype=self.adt)
        ops.gemm(cap16, w["c_w0"], bias=w["c_b0"], act=ops.ACT_GELU, out_fp16=h1, outb_ld=D)
        h2 = torch.empty(B * L, D, device=dev, dtype=F32)
        ops.gemm(h1, w["c_w2"], bias=w["c_b2"], out_f32=h2, out_ld=D)
        cap = torch.empty(B, L, D, device=dev, dtype=F32)
        if self.adt == H16:
            cap_bf = torch.empty(B * L, D, device=dev, dtype=H16)
            ops.norm_mod(h2, kind=1, w=w["c_ln_w"], b=w["c_ln_b"], eps=1e-5, out_f32=cap, out_fp16=cap_bf)
        else:  # fp32 operands: the fp32 caption is the GEMM operand itself
            ops.norm_mod(h2, kind=1, w=w["c_ln_w"], b=w["c_ln_b"], eps=1e-5, out_f32=cap)
            cap_bf = cap.view(B * L, D)
        pool = torch.empty(B, D, device=dev, dtype=F32)
        ops.mean_tokens(cap, pool)
        pool_n = torch.empty(B, D, device=dev, dtype=self.adt)
        ops.norm_mod(pool, kind=1, w=w["cap_ln_w"], b=w["cap_ln_b"], eps=1e-5, out_fp16=pool_n)
        cap_emb = torch.empty(B, D, device=dev, dtype=F32)
        ops.gemm(pool_n, w["cap_w"], bias=w["cap_b"], out_f32=cap_emb, out_ld=D)
        cond = DiTCond(B=B, T=T, L=L, acoustic=ac, ac_logits=ac_logits, caption=cap, cap_emb=cap_emb)
        y16 = torch.empty(B * L, D, device=dev, dtype=self.adt)
        H, hd = self.H, self.hd
        Lp = (L + 7) // 8 * 8
        for k in range(self.depth):
            b = w["blocks"][k]
            # K / V^T of the two short-KV attentions, written head-major by the k|v GEMM epilogue
            ops.norm_mod(cap.view(B * L, D), kind=0, w=b["y_norm"], eps=self.eps, out_fp16=y16)
            kk = torch.empty(B, H, L, hd, device=dev, dtype=H16)
            vt = torch.zeros(B, H, hd, Lp, device=dev, dtype=H16)
            ops.gemm(y16, b["wkvy"], mode=ops.EPI_QKV_ROPE, qkv_sec0=1, T=L, heads=H, head_dim=hd, Tp=Lp,
                     k_out=kk, vt_out=vt)
            cond.k_y.append(kk)
            cond.vt_y.append(vt)
            # MoE caption cross-attention (vocal2music_moe.py:117-131), folded: fp32 k|v projection, then
            # WK[b,h] = K_h (scale W_q,h) as one batched GEMM, score bias and V.(W_gate W_o) by vb_route_prepare
            kvf = torch.empty(B * L, 2 * D, device=dev, dtype=F32)
            kvo = kvf if self.adt == F32 else torch.empty(B * L, 2 * D, device=dev, dtype=H16)
            ops.gemm(cap_bf, b["mkv_w"], bias=b["mkv_b"], out_f32=kvf, out_ld=2 * D,
                     out_fp16=None if self.adt == F32 else kvo, outb_ld=2 * D)
            wk = torch.zeros(B, H * self.LK, D, device=dev, dtype=self.adt)
            ops.gemm(kvo[:, :D].view(B, L, H, hd).permute(0, 2, 1, 3), b["mq_wT"], M=L, N=D, G1=H, G2=B,
                     a_c2=(1, 0), a_c3=(0, 1), b_c2=(1, 0, 0), out_fp16=wk, outb_ld=D,
                     outb_g=(self.LK * D, H * self.LK * D))
            bs = torch.empty(B, H, self.LK, device=dev, dtype=F32)
            vg = torch.empty(B, H, self.LK, self.E, device=dev, dtype=F32)
            ops.route_prepare(kvf, B, L, H, hd, b["mq_bs"], b["capg_w"], bs, vg)
            cond.wk_m.append(wk)
            cond.bs_m.append(bs)
            cond.vg_m.append(vg)
        return cond

    @torch.no_grad()
    def prepare_pair(self, midi, beats, caption: torch.Tensor, caption_uncond: torch.Tensor, T: int) -> DiTCond:
        """Conditioning of the (cond, uncond) CFG pair as ONE batch of 2B clips (cond rows first): the
        acoustic stem is computed once and shared, the caption path runs once on the stacked captions."""
        caps = torch.cat([caption.to(self.dev, F32), caption_uncond.to(self.dev, F32)], 0)
        if self.exact_stem and ops.TIMELINE is None:
            return self._prepare_c(midi, beats, caps, T)
        ac, ac_logits = self.prepare_acoustic(midi, beats, T)
        return self.prepare_py(None, None, caps, T, acoustic=(torch.cat([ac, ac], 0), torch.cat([ac_logits, ac_logits], 0)))

    # ------------------------------------------------------------------ time conditioning
    @torch.no_grad()
    def time_conditioning(self, t: torch.Tensor, cap_emb: torch.Tensor, mod_out: torch.Tensor, hl_out: torch.Tensor):
        """t int64 [R], cap_emb fp32 [R, D] -> mod_out [R, mod_ld] (adaLN vectors of every block + final
        layer) and hl_out [R, 2*depth] (group-gate logits). Rows are independent (R = batch, or
        steps x batch when the sampler precomputes the whole schedule)."""
        w, D, dev = self.w, self.D, self.dev
        R = t.numel()
        sin = torch.empty(R, 256, device=dev, dtype=self.adt)
        h = torch.empty(R, D, device=dev, dtype=self.adt)
        temb = torch.empty(R, D, device=dev, dtype=F32)
        ada = torch.empty(R, D, device=dev, dtype=self.adt)
        ops.timestep_embed(t, w["t_freqs"], out_fp16=sin)
        ops.gemm(sin, w["t_w0"], bias=w["t_b0"], act=ops.ACT_SILU, out_fp16=h, outb_ld=D)
        ops.gemm(h, w["t_w2"], bias=w["t_b2"], out_f32=temb, out_ld=D)
        ops.ew(temb, cap_emb, act=ops.ACT_SILU, out_fp16=ada)
        ops.gemm(ada, w["ada_w"], bias=w["ada_b"], out_f32=mod_out, out_ld=self.mod_ld)
        ops.linear_small(temb, w["hl_w"], w["hl_b"], hl_out)


    def _router_and_dispatch(self, ws, b, cond, hl_t, k, e0, e1, e2, N, T, E, max_rows, want_aux):
        """Two-launch variant (router, then dispatch over back-to-back padded buckets)."""
        ops.router(None, None, b["capg_b"], cond.ac_logits[:, k * E:], self.depth * E,
                   hl_t[:, 2 * k:], 2 * self.depth, e0, e1, e2, float(self.temperature), N, T,
                   ws["idx_c"], ws["idx_a"], ws["scale_c"], ws["scale_a"],
                   ws["cap_logits"] if want_aux else None, ws["bucket_count"][k], ws["usage"][k],
                   ws["slot_c"], ws["slot_a"], cap_partial=ws["cap_part"])
        ops.mark("router")
        if self.depth > 1:
            kn = (k + 1) % self.depth  # the dispatch of block k zeroes the counters of block k+1
            ops.moe_dispatch(ws["idx_c"], ws["idx_a"], ws["slot_c"], ws["slot_a"], ws["scale_c"], ws["scale_a"],
                             ws["bucket_count"][k], ws["usage"][k], ws["bucket_count"][kn], ws["usage"][kn], N, E,
                             ws["xm"], ws["xs"], ws["row_map"], ws["row_scale"], ws["tile_group"], ws["m_tiles"],
                             max_rows, ws["lb"][k:])
        else:  # single block: counters must be re-zeroed by a separate launch
            ops.moe_finalize(ws["bucket_count"][0], ws["usage"][0], E, ws["bucket_start"], ws["row_map"],
                             ws["row_scale"], ws["tile_group"], ws["m_tiles"], max_rows, ws["lb"][k:])
            ops.moe_scatter(ws["idx_c"], ws["idx_a"], ws["slot_c"], ws["slot_a"], ws["scale_c"], ws["scale_a"],
                            ws["bucket_start"], N, E, ws["xm"], ws["xs"], ws["row_map"], ws["row_scale"])
        ops.mark("moe_dispatch")

    # ------------------------------------------------------------------ one DiT evaluation
    def _desc(self):
        """The packed model as a `vb_dit_desc` (include/versband.h) for vb_dit_forward."""
        if getattr(self, "_c_desc", None) is not None:
            return self._c_desc
        from ._lib import DIT_MAX_DEPTH, DitDesc
        w, P = self.w, ops._ptr
        assert self.depth <= DIT_MAX_DEPTH
        d = DitDesc()
        d.C, d.D, d.depth, d.H, d.E, d.Fh = self.C, self.D, self.depth, self.H, self.E, self.Fh
        d.max_len, d.proj_in_k, d.operand_f32 = self.max_len, self.proj_in_k, int(self.adt == F32)
        d.eps, d.temperature, d.mod_ld = self.eps, float(self.temperature), self.mod_ld
        for n in ("proj_in_w", "proj_in_b", "t_w0", "t_w2", "ada_w", "t_b0", "t_b2", "ada_b", "t_freqs", "hl_w", "hl_b",
                  "final_w", "final_b", "rope"):
            setattr(d, n, P(w[n]))
        for blk, b in zip(d.blocks, w["blocks"]):
            for n in ("wqkv", "wo", "gate", "attn_norm", "ffn_norm", "capg_b", "w13", "w2", "f13", "f2"):
                setattr(blk, n, P(b[n]))
        self._c_desc = d
        return d

    def _cond_struct(self, cond: DiTCond):
        cs = getattr(cond, "_cstruct", None)
        if cs is None:
            from ._lib import DitCond
            P = ops._ptr
            cs = DitCond()
            cs.B, cs.T, cs.L, cs.Lp = cond.B, cond.T, cond.L, cond.vt_y[0].shape[-1]
            cs.acoustic, cs.ac_logits, cs.cap_emb = P(cond.acoustic), P(cond.ac_logits), P(cond.cap_emb)
            for name in DiTCond._LISTS:
                arr = getattr(cs, name)
                for k, t in enumerate(getattr(cond, name)):
                    assert t.is_contiguous()
                    arr[k] = t.data_ptr()
            cond._cstruct = cs
        return cs

    @torch.no_grad()
    def forward(self, x: torch.Tensor, t: torch.Tensor, cond: DiTCond,
                expo: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                out_v: Optional[torch.Tensor] = None, want_aux: bool = False, pre=None, x_cl_ready: bool = False,
                chain_cnt: Optional[torch.Tensor] = None, ws_slot: int = 0, x_cl: Optional[torch.Tensor] = None):
        """x fp32 [B, C, T], t int64 [B]; expo[k] = Exp(1) draws ([N,2],[N,E],[N,E]) of block k.
        Returns v as the channels-last GEMM output [N, 32] (first C columns valid) unless `out_v`
        ([B, C, T]) is given, plus lb_loss per block (device tensor).

        ONE call of the model-level C entry point vb_dit_forward (csrc/vb_model.cu issues the ~52 launches);
        `forward_py` is the same sequence issued from Python — used when per-block intermediates are wanted
        (`want_aux`), for the experimental launch variants (use_chain / fused_dispatch) and the op timeline."""
        if want_aux or self.use_chain or self.fused_dispatch or ops.TIMELINE is not None:
            return self.forward_py(x, t, cond, expo, out_v=out_v, want_aux=want_aux, pre=pre, x_cl_ready=x_cl_ready,
                                   chain_cnt=chain_cnt, ws_slot=ws_slot, x_cl=x_cl)
        import ctypes as C
        from . import _lib
        from ._lib import DitIO
        Bx, Cc, T = x.shape
        B = cond.B
        assert B % Bx == 0 and T == cond.T and T <= self.max_len
        ws = self._workspace(B, T, ws_slot)
        d = self._desc()
        if "cws" not in ws:
            need = int(_lib.lib().vb_dit_workspace_bytes(C.byref(d), B, T))
            if need < 0:
                _lib.check(-1, "vb_dit_workspace_bytes")
            ws["cws"] = torch.zeros(need, device=self.dev, dtype=torch.uint8)  # zero-filled once (MoE bucket counters)
        P = ops._ptr
        io = DitIO()
        x = x.contiguous()
        t = t.contiguous()
        io.x, io.Bx, io.x_cl_ready = P(x), Bx, int(x_cl_ready)
        io.x_cl = P(ws["x_cl"] if x_cl is None else x_cl)
        io.t = P(t)
        if pre is not None:
            assert pre[0].is_contiguous() and pre[1].is_contiguous()
            io.mod, io.hl = P(pre[0]), P(pre[1])
        for k, (e0, e1, e2) in enumerate(expo):
            io.e0[k], io.e1[k], io.e2[k] = e0.data_ptr(), e1.data_ptr(), e2.data_ptr()
        io.v, io.lb = P(ws["v"]), P(ws["lb"])
        if out_v is not None:
            assert out_v.is_contiguous()
            io.out_v = P(out_v)
        _lib.check(_lib.lib().vb_dit_forward(C.byref(d), C.byref(self._cond_struct(cond)), C.byref(io), P(ws["cws"]),
                                             ws["cws"].numel(), ops._stream()), "vb_dit_forward")
        return ws["v"], ws["lb"]

    @torch.no_grad()
    def forward_py(self, x: torch.Tensor, t: torch.Tensor, cond: DiTCond,
                   expo: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                   out_v: Optional[torch.Tensor] = None, want_aux: bool = False, pre=None, x_cl_ready: bool = False,
                   chain_cnt: Optional[torch.Tensor] = None, ws_slot: int = 0, x_cl: Optional[torch.Tensor] = None):
        """The launch sequence of `forward`, op by op from Python (same kernels, same arguments)."""
        w, D, H, hd, E, Fh = self.w, self.D, self.H, self.hd, self.E, self.Fh
        Bx, C, T = x.shape
        B = cond.B
        assert B % Bx == 0 and T == cond.T and T <= self.max_len
        N, L = B * T, cond.L
        ws = self._workspace(B, T, ws_slot)
        if x_cl is not None:  # caller-owned channels-last copy of the latent (shared between concurrent branches)
            ws = dict(ws, x_cl=x_cl)
        Tp, max_rows = ws["Tp"], ws["max_rows"]
        scale = 1.0 / math.sqrt(hd)
        aux = []
        # dependency counters of the chained launches: the caller's pre-zeroed [depth, chain_len] slice (the
        # sampler zeroes all steps with one memset) or this engine's own, zeroed here
        cnt = chain_cnt
        if cnt is None and self.use_chain:
            cnt = ws["chain_cnt"]
            cnt.zero_()

        # stem: proj_in conv (k5) + bias, fused "+ acoustic"  (vocal2music_moe.py:395,415)
        # (x_cl_ready: the sampler's Euler-update launch already refreshed the channels-last copies)
        for r in range(0 if x_cl_ready else B // Bx):  # CFG batching: the same latent feeds every conditioning branch
            ops.nct_to_ntc(x.contiguous(), ws["x_cl"][r * Bx:(r + 1) * Bx])
            ops.mark("nct_to_ntc")
        ops.gemm(ws["x_cl"][:, :, :C], w["proj_in_w"], M=T, G1=B, kpt=1, dil=1, pad=self.proj_in_k // 2,
                 a_c2=(1, 0), bias=w["proj_in_b"], residual=cond.acoustic, out_f32=ws["x"], out_ld=D,
                 out_g=(T * D, 0))
        ops.mark("gemm:proj_in_w")
        # timestep embedding MLP, adaLN input and all modulation vectors (:406,416,242,290): depend only on
        # (t, caption) — the sampler precomputes them for all 24 steps at once (`time_conditioning`)
        if pre is None:
            self.time_conditioning(t.contiguous(), cond.cap_emb, ws["mod"], ws["hl"])
            mod_t, hl_t = ws["mod"], ws["hl"]
        else:
            mod_t, hl_t = pre
        mod, mld = mod_t, self.mod_ld

        xcur, hbuf = ws["x"], ws["h"]
        for k in range(self.depth):
            b = w["blocks"][k]
            m0 = k * 6 * D
            sh_msa, sc_msa, g_msa = mod[:, m0:], mod[:, m0 + D:], mod[:, m0 + 2 * D:]
            sh_mlp, sc_mlp, g_mlp = mod[:, m0 + 3 * D:], mod[:, m0 + 4 * D:], mod[:, m0 + 5 * D:]
            # ---- attention branch (vocal2music_moe.py:245-250; flag_large_dit_moe.py:323-406)
            ops.norm_mod(xcur, kind=0, w=b["attn_norm"], shift=sh_msa, scale=sc_msa, mod_ld=mld,
                         rows_per_batch=T, eps=self.eps, out_fp16=ws["xa"])
            ops.mark("norm_mod")
            ops.gemm(ws["xa"], b["wqkv"], mode=ops.EPI_QKV_ROPE, rope_cs=w["rope"], T=T, heads=H, head_dim=hd,
                     Tp=Tp, q_out=ws["q"], k_out=ws["k"], vt_out=ws["vt"])
            ops.mark("gemm:wqkv")
            # self-attention + gated text cross-attention in one tcgen05 flash-attention launch
            ops.attn_fused(ws["q"], (hd, T * hd, H * T * hd), B=B, T=T, H=H, hd=hd, scale=scale,
                           k_self=ws["k"], vt_self=ws["vt"], Tp=Tp, k_text=cond.k_y[k], vt_text=cond.vt_y[k],
                           L=L, Lp=cond.vt_y[k].shape[-1], gate=b["gate"], out_fp16=ws["oa"])
            ops.mark("attn_fused")
            ops.gemm(ws["oa"], b["wo"], gate=g_msa, gate_rows=T, gate_ld=mld, residual=xcur,
                     out_f32=hbuf, out_ld=D)
            ops.mark("gemm:wo")
            # ---- Band-MoE branch (vocal2music_moe.py:251-257, 117-185)
            ops.norm_mod(hbuf, kind=0, w=b["ffn_norm"], shift=sh_mlp, scale=sc_mlp, mod_ld=mld,
                         rows_per_batch=T, eps=self.eps, out_fp16=ws["xm"])
            ops.mark("norm_mod")
            # caption gate: per-head softmax(xm . WK^T + bias) . VG inside one GEMM epilogue
            ops.gemm(ws["xm"].view(B, T, D), cond.wk_m[k], M=T, G1=B, a_c2=(1, 0), b_c2=(1, 0, 0), mode=ops.EPI_ROUTE,
                     bias=cond.bs_m[k], bias_g1=H * self.LK, route_vg=cond.vg_m[k], route_vg_g1=H * self.LK * E,
                     out_f32=ws["cap_part"], out_ld=H * E)
            ops.mark("gemm:route")
            e0, e1, e2 = expo[k]
            if self.fused_dispatch:
                ops.route_dispatch(ws["cap_part"], b["capg_b"], cond.ac_logits[:, k * E:], self.depth * E,
                                   hl_t[:, 2 * k:], 2 * self.depth, e0, e1, e2, float(self.temperature), N, T, E,
                                   ws["xm"], ws["xs"], ws["cap"], ws["idx_c"], ws["idx_a"], ws["bucket_count"][k],
                                   ws["usage"][k], ws["done_ticket"][k:], ws["row_map"], ws["row_scale"],
                                   ws["tile_group"], ws["tile_row0"], ws["m_tiles"], lb_out=ws["lb"][k:],
                                   cap_logits_out=ws["cap_logits"] if want_aux else None)
                ops.mark("route_dispatch")
                tile_row0 = ws["tile_row0"]
            else:
                tile_row0 = None
                self._router_and_dispatch(ws, b, cond, hl_t, k, e0, e1, e2, N, T, E, max_rows, want_aux)
            # routed experts (w1|w3 -> SwiGLU -> w2, scattered back to token rows) and frequency experts
            # (band-sliced w1|w3 -> SwiGLU -> w2 + adaLN gate + residual): four dependent GEMM problems in ONE
            # persistent launch (vb_gemm_chain) — or four launches with use_chain = False.
            # y = y_caption + y_acoustic (vocal2music_moe.py:166) is never materialised: the frequency experts'
            # first GEMM reads both halves of y2 through the conv addressing (2 "taps" N rows apart, the band's
            # weights repeated along K), so the sum happens in the fp32 accumulator.
            yb = ws["y2"].view(2 * N, E, self.band).permute(1, 0, 2)  # [E, 2N, band] view: per-band K slice
            probs = [
                (ws["xs"], b["w13"], dict(M=max_rows, b_c2=(0, 0, 1), tile_group=ws["tile_group"], tile_row0=tile_row0,
                                          m_tiles_used=ws["m_tiles"], swiglu=True, out_fp16=ws["hs"], outb_ld=Fh)),
                (ws["hs"], b["w2"], dict(M=max_rows, b_c2=(0, 0, 1), tile_group=ws["tile_group"], tile_row0=tile_row0,
                                         m_tiles_used=ws["m_tiles"], row_map=ws["row_map"], rows_out=2 * N,
                                         row_scale=ws["row_scale"], out_fp16=ws["y2"], outb_ld=D)),
                (yb, b["f13"], dict(M=N, G1=E, kpt=self.band // self.kq, dil=N, pad=0, a_c2=(1, 0), b_c2=(1, 0, 0),
                                    swiglu=True, out_fp16=ws["hf"], outb_ld=Fh, outb_g=(N * Fh, 0))),
                (ws["hf"], b["f2"], dict(M=N, G1=E, a_c2=(1, 0), b_c2=(1, 0, 0),
                                         gate=g_mlp, gate_rows=T, gate_ld=mld, gate_g1=self.band, residual=hbuf,
                                         out_f32=xcur, out_ld=D, out_g=(self.band, 0))),
            ]
            if self.use_chain:
                ops.gemm_chain([ops.gemm_args(a_, b_, **kw) for a_, b_, kw in probs], [0, 1, 2, 1], cnt[k],
                               bn=self.chain_bn)
                ops.mark("gemm:moe_chain")
            else:
                for (a_, b_, kw), name in zip(probs, ("w13", "w2", "f13", "f2")):
                    ops.gemm(a_, b_, **kw)
                    ops.mark("gemm:" + name)
            if want_aux:
                aux.append(dict(idx_c=ws["idx_c"].clone(), idx_a=ws["idx_a"].clone(),
                                cap_logits=ws["cap_logits"].clone(), h=hbuf.clone(), x=xcur.clone()))
        # ---- final layer (vocal2music_moe.py:268-291,431)
        f0 = self.depth * 6 * D
        ops.norm_mod(xcur, kind=1, shift=mod[:, f0:], scale=mod[:, f0 + D:], mod_ld=mld, rows_per_batch=T,
                     eps=1e-6, out_fp16=ws["xf"])
        ops.mark("norm_mod")
        # all 32 (zero-padded) output columns are computed so the chunk takes the coalesced epilogue path
        ops.gemm(ws["xf"], w["final_w"], bias=w["final_b"], out_f32=ws["v"], out_ld=32)
        ops.mark("gemm:final_w")
        if out_v is not None:
            ops.ntc_to_nct(ws["v"], out_v, ld=32)
            ops.mark("ntc_to_nct")
        return (ws["v"], ws["lb"], aux) if want_aux else (ws["v"], ws["lb"])
