"""AutoencoderKL (1-D) decoder / encoder on B200 as implicit-GEMM convolutions.

Mirrors reference ldm/models/autoencoder1d.py (Decoder1D.forward :480-513, Encoder1D.forward
:384-409, ResnetBlock1D :211-231, AttnBlock1D :253-274, Upsample1D :287-291, Downsample1D
:306-313).  Activations are channels-last fp32 [B, T, C] in HBM; every conv is one vb_gemm launch
whose A operand is the fp16 GroupNorm+swish output (TMA out-of-bounds zero fill = conv padding).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .dit import H16, F32, _h16, pack_conv_weight


class VAEEngine:
    def __init__(self, sd: Dict[str, torch.Tensor], ddconfig: dict, embed_dim: int, prefix: str = "",
                 device="cuda", with_encoder: bool = True):
        self.dev = torch.device(device)
        self.dd = dict(ddconfig)
        self.embed_dim = embed_dim
        self.n_levels = len(ddconfig["ch_mult"])
        self.n_res = ddconfig["num_res_blocks"]
        self.down_layers = list(ddconfig.get("down_layers", []))
        self.up_levels = [i + 1 for i in self.down_layers]  # autoencoder1d.py:423
        self.G = 32
        self.w: Dict[str, torch.Tensor] = {}
        self.k: Dict[str, int] = {}
        self._pack(sd, prefix, with_encoder)

    # ------------------------------------------------------------------ packing
    def _pack(self, sd, prefix, with_encoder):
        dev = self.dev
        for key, t in sd.items():
            if not key.startswith(prefix):
                continue
            name = key[len(prefix):]
            if name.startswith("encoder.") and not with_encoder:
                continue
            if name.startswith("loss."):
                continue
            t = t.detach().to(dev, F32)
            if name.endswith(".weight") and t.dim() == 3:
                base = name[: -len(".weight")]
                if base.endswith("downsample.conv"):
                    # k3 stride-2 conv on rows [T/2, 2C]: tap0 = (j=0 | j=1), tap1 = (j=2 | 0)
                    co, ci, k = t.shape
                    assert k == 3 and (2 * ci) % 64 == 0
                    wp = torch.zeros(co, 2, 2 * ci, device=dev, dtype=H16)
                    wp[:, 0, :ci], wp[:, 0, ci:], wp[:, 1, :ci] = t[:, :, 0].to(H16), t[:, :, 1].to(H16), t[:, :, 2].to(H16)
                    self.w[base + ".w"] = wp.reshape(co, 4 * ci).contiguous()
                else:
                    self.w[base + ".w"] = pack_conv_weight(t)
                self.k[base] = t.shape[-1]
            else:
                self.w[name] = t.contiguous()
        # fused q|k|v 1x1 convs of every attention block
        for name in [n for n in list(self.w) if n.endswith(".q.w")]:
            base = name[: -len(".q.w")]
            self.w[base + ".qkv.w"] = torch.cat([self.w[base + f".{c}.w"] for c in "qkv"], 0).contiguous()
            self.w[base + ".qkv.bias"] = torch.cat([self.w[base + f".{c}.bias"] for c in "qkv"], 0).contiguous()

    # ------------------------------------------------------------------ building blocks
    def _conv(self, a_bf, name, B, T, *, out_f32=None, out_fp16=None, residual=None, Cin=None):
        """Conv1d(k, stride 1, 'same') over channels-last fp16 a_bf [B, T, Cpitch]."""
        w = self.w[name + ".w"]
        k = self.k[name]
        Cout = w.shape[0]
        kpt = w.shape[1] // k // 64
        a = a_bf if Cin is None else a_bf[:, :, :Cin]
        ops.gemm(a, w, M=T, G1=B, kpt=kpt, dil=1, pad=k // 2, a_c2=(1, 0), bias=self.w[name + ".bias"],
                 residual=residual, out_f32=out_f32, out_ld=Cout, out_g=(T * Cout, 0),
                 out_fp16=out_fp16, outb_ld=(out_fp16.shape[-1] if out_fp16 is not None else 0),
                 outb_g=((T * out_fp16.shape[-1], 0) if out_fp16 is not None else (0, 0)))
        return Cout

    def _gn(self, x, name, B, T, C, act):
        if ops.gn_fused_fits(T, C, self.G):  # one cluster launch: stats through DSMEM, x read once
            out = torch.empty(B, T, C, device=self.dev, dtype=H16)
            ops.gn_fused(x, self.w[name + ".weight"], self.w[name + ".bias"], B, T, C, self.G, 1e-6, act, out)
            return out
        mean = torch.empty(B * self.G, device=self.dev, dtype=F32)
        rstd = torch.empty_like(mean)
        out = torch.empty(B, T, C, device=self.dev, dtype=H16)
        ops.gn_stats(x, B, T, C, self.G, 1e-6, mean, rstd)
        ops.gn_apply(x, mean, rstd, self.w[name + ".weight"], self.w[name + ".bias"], B, T, C, self.G, act, out)
        return out

    def _resblock(self, x, name, B, T):
        C = x.shape[-1]
        a = self._gn(x, name + ".norm1", B, T, C, ops.ACT_SILU)
        Cout = self.w[name + ".conv1.w"].shape[0]
        h = torch.empty(B, T, Cout, device=self.dev, dtype=F32)
        self._conv(a, name + ".conv1", B, T, out_f32=h)
        a2 = self._gn(h, name + ".norm2", B, T, Cout, ops.ACT_SILU)
        res = x
        if (name + ".nin_shortcut.w") in self.w:
            xb = torch.empty(B, T, C, device=self.dev, dtype=H16)
            ops.ew(x, out_fp16=xb)
            res = torch.empty(B, T, Cout, device=self.dev, dtype=F32)
            self._conv(xb, name + ".nin_shortcut", B, T, out_f32=res)
        out = torch.empty(B, T, Cout, device=self.dev, dtype=F32)
        self._conv(a2, name + ".conv2", B, T, out_f32=out, residual=res)
        return out

    def _attn(self, x, name, B, T):
        C = x.shape[-1]
        Tp = (T + 7) // 8 * 8
        a = self._gn(x, name + ".norm", B, T, C, ops.ACT_NONE)
        q = torch.empty(B, 1, T, C, device=self.dev, dtype=H16)
        k = torch.empty_like(q)
        vt = torch.zeros(B, 1, C, Tp, device=self.dev, dtype=H16)
        ops.gemm(a.view(B * T, C), self.w[name + ".qkv.w"], mode=ops.EPI_QKV_ROPE, bias=self.w[name + ".qkv.bias"],
                 T=T, heads=1, head_dim=C, Tp=Tp, q_out=q, k_out=k, vt_out=vt)
        S = torch.empty(B, T, Tp, device=self.dev, dtype=F32)
        ops.gemm(q.view(B, T, C), k.view(B, T, C), M=T, N=T, K=C, G1=B, a_c2=(1, 0), b_c2=(1, 0, 0),
                 alpha=float(int(C) ** (-0.5)), out_f32=S, out_ld=Tp, out_g=(T * Tp, 0))
        P = torch.empty(B, T, Tp, device=self.dev, dtype=H16)
        ops.softmax_rows(S, Tp, P, Tp, B * T, T)
        o = torch.empty(B, T, C, device=self.dev, dtype=H16)
        ops.gemm(P[:, :, :T], vt.view(B, C, Tp)[:, :, :T], M=T, N=C, K=T, G1=B, a_c2=(1, 0), b_c2=(1, 0, 0),
                 out_fp16=o, outb_ld=C, outb_g=(T * C, 0))
        out = torch.empty(B, T, C, device=self.dev, dtype=F32)
        self._conv(o, name + ".proj_out", B, T, out_f32=out, residual=x)
        return out

    # ------------------------------------------------------------------ decoder
    def _decoder_layers(self):
        """(kind, name) of every decoder layer in execution order (autoencoder1d.py:480-513)."""
        d = "decoder."
        layers = [("res", d + "mid.block_1"), ("attn", d + "mid.attn_1"), ("res", d + "mid.block_2")]
        for lvl in reversed(range(self.n_levels)):
            for blk in range(self.n_res + 1):
                layers.append(("res", f"{d}up.{lvl}.block.{blk}"))
                if f"{d}up.{lvl}.attn.{blk}.norm.weight" in self.w:
                    layers.append(("attn", f"{d}up.{lvl}.attn.{blk}"))
            if lvl in self.up_levels:
                layers.append(("up", f"{d}up.{lvl}.upsample.conv"))
        return layers

    def _desc(self):
        """The decoder as a `vb_vae_desc` (include/versband.h) for vb_vae_decode."""
        if getattr(self, "_vae_desc", None) is not None:
            return self._vae_desc
        from ._lib import VAE_MAX_OPS, VaeDesc
        w, k, P = self.w, self.k, ops._ptr
        d = "decoder."
        ds = VaeDesc()
        ds.zc, ds.zq_c = self.embed_dim, w["post_quant_conv.w"].shape[0]
        ds.c0, ds.out_ch, ds.groups, ds.eps = w[d + "conv_in.w"].shape[0], w[d + "conv_out.w"].shape[0], self.G, 1e-6
        ds.pq_k, ds.in_k, ds.out_k = k["post_quant_conv"], k[d + "conv_in"], k[d + "conv_out"]
        ds.pq_w, ds.pq_b = P(w["post_quant_conv.w"]), P(w["post_quant_conv.bias"])
        ds.in_w, ds.in_b = P(w[d + "conv_in.w"]), P(w[d + "conv_in.bias"])
        ds.no_w, ds.no_b = P(w[d + "norm_out.weight"]), P(w[d + "norm_out.bias"])
        ds.out_w, ds.out_b = P(w[d + "conv_out.w"]), P(w[d + "conv_out.bias"])
        layers = self._decoder_layers()
        assert len(layers) <= VAE_MAX_OPS
        ds.n_ops = len(layers)
        for o, (kind, n) in zip(ds.ops, layers):
            self._fill_op(o, kind, n)
        self._vae_desc = ds
        return ds

    @torch.no_grad()
    def decode(self, z: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """z fp32 [B, zc, T] (multiplied by `scale` = 1/scale_factor) -> mel fp32 [B, out_ch, T*2^n]: one call of
        the model-level C entry point vb_vae_decode (csrc/vb_model.cu), which issues the launch sequence."""
        import ctypes as C
        from . import _lib
        B, zc, T = z.shape
        z = z.to(self.dev, F32).contiguous()
        ds = self._desc()
        need = _lib.lib().vb_vae_workspace_bytes(C.byref(ds), B, T)
        if need < 0:
            _lib.check(1, "vb_vae_workspace_bytes")
        ws = torch.empty(need, device=self.dev, dtype=torch.uint8)
        mel = torch.empty(B, ds.out_ch, T * 2 ** len(self.up_levels), device=self.dev, dtype=F32)
        _lib.check(_lib.lib().vb_vae_decode(C.byref(ds), ops._ptr(z), float(scale), B, T, ops._ptr(mel),
                                            ops._ptr(ws), need, ops._stream()), "vb_vae_decode")
        return mel

    @torch.no_grad()
    def decode_py(self, z: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """The same launch sequence issued op by op from Python (kept as the readable statement of decode and as
        the cross-check of the C entry point, tests/test_vae_vocoder_gpu.py)."""
        B, zc, T = z.shape
        dev = self.dev
        zp = (zc + 7) // 8 * 8
        z_cl = torch.zeros(B, T, zp, device=dev, dtype=H16)
        ops.nct_to_ntc(z.to(dev, F32).contiguous(), z_cl, scale=scale)
        zq = torch.zeros(B, T, zp, device=dev, dtype=H16)
        self._conv(z_cl, "post_quant_conv", B, T, out_fp16=zq, Cin=zc)
        d = "decoder."
        C0 = self.w[d + "conv_in.w"].shape[0]
        h = torch.empty(B, T, C0, device=dev, dtype=F32)
        self._conv(zq, d + "conv_in", B, T, out_f32=h, Cin=self.w["post_quant_conv.w"].shape[0])
        h = self._resblock(h, d + "mid.block_1", B, T)
        h = self._attn(h, d + "mid.attn_1", B, T)
        h = self._resblock(h, d + "mid.block_2", B, T)
        for lvl in reversed(range(self.n_levels)):
            for blk in range(self.n_res + 1):
                h = self._resblock(h, f"{d}up.{lvl}.block.{blk}", B, T)
                if f"{d}up.{lvl}.attn.{blk}.norm.weight" in self.w:
                    h = self._attn(h, f"{d}up.{lvl}.attn.{blk}", B, T)
            if lvl in self.up_levels:  # nearest x2 then conv k3 (autoencoder1d.py:287-291)
                C = h.shape[-1]
                up = torch.empty(B, 2 * T, C, device=dev, dtype=H16)
                ops.ew(h, out_fp16=up, rep=2)
                T = 2 * T
                h = torch.empty(B, T, C, device=dev, dtype=F32)
                self._conv(up, f"{d}up.{lvl}.upsample.conv", B, T, out_f32=h)
        C = h.shape[-1]
        a = self._gn(h, d + "norm_out", B, T, C, ops.ACT_SILU)
        Cout = self.w[d + "conv_out.w"].shape[0]
        mel_cl = torch.empty(B, T, Cout, device=dev, dtype=F32)
        self._conv(a, d + "conv_out", B, T, out_f32=mel_cl)
        mel = torch.empty(B, Cout, T, device=dev, dtype=F32)
        ops.ntc_to_nct(mel_cl, mel, ld=Cout)
        return mel

    # ------------------------------------------------------------------ encoder (config 1)
    def _fill_op(self, o, kind: str, n: str) -> None:
        """One layer of a vb_vae_desc from the packed weights under prefix `n`."""
        from ._lib import VAE_ATTN, VAE_DOWN, VAE_RES, VAE_UP
        w, k, P = self.w, self.k, ops._ptr
        if kind == "res":
            o.kind, o.cout = VAE_RES, w[n + ".conv1.w"].shape[0]
            o.cin = w[n + ".norm1.weight"].shape[0]
            o.k1, o.k2 = k[n + ".conv1"], k[n + ".conv2"]
            o.n1_w, o.n1_b, o.n2_w, o.n2_b = (P(w[n + s]) for s in (".norm1.weight", ".norm1.bias", ".norm2.weight", ".norm2.bias"))
            o.w1, o.b1, o.w2, o.b2 = P(w[n + ".conv1.w"]), P(w[n + ".conv1.bias"]), P(w[n + ".conv2.w"]), P(w[n + ".conv2.bias"])
            if (n + ".nin_shortcut.w") in w:
                o.ws, o.bs, o.ks = P(w[n + ".nin_shortcut.w"]), P(w[n + ".nin_shortcut.bias"]), k[n + ".nin_shortcut"]
        elif kind == "attn":
            o.kind = VAE_ATTN
            o.cin = o.cout = w[n + ".norm.weight"].shape[0]
            assert k[n + ".q"] == 1 and k[n + ".proj_out"] == 1
            o.k1, o.k2 = 1, 1
            o.n1_w, o.n1_b = P(w[n + ".norm.weight"]), P(w[n + ".norm.bias"])
            o.w1, o.b1, o.w2, o.b2 = P(w[n + ".qkv.w"]), P(w[n + ".qkv.bias"]), P(w[n + ".proj_out.w"]), P(w[n + ".proj_out.bias"])
        else:
            o.kind, o.cout, o.k1 = (VAE_UP if kind == "up" else VAE_DOWN), w[n + ".w"].shape[0], k[n]
            o.cin = o.cout
            o.w1, o.b1 = P(w[n + ".w"]), P(w[n + ".bias"])

    def _desc_enc(self):
        """The encoder (+ quant_conv) as a `vb_vae_desc` for vb_vae_encode (include/versband.h)."""
        if getattr(self, "_vae_desc_enc", None) is not None:
            return self._vae_desc_enc
        from ._lib import VAE_MAX_OPS, VaeDesc
        w, k, P = self.w, self.k, ops._ptr
        e = "encoder."
        ds = VaeDesc()
        ds.zc, ds.c0 = self.dd["in_channels"], w[e + "conv_in.w"].shape[0]
        ds.out_ch, ds.zq_c, ds.groups, ds.eps = w[e + "conv_out.w"].shape[0], w["quant_conv.w"].shape[0], self.G, 1e-6
        ds.in_k, ds.out_k, ds.pq_k = k[e + "conv_in"], k[e + "conv_out"], k["quant_conv"]
        ds.in_w, ds.in_b = P(w[e + "conv_in.w"]), P(w[e + "conv_in.bias"])
        ds.no_w, ds.no_b = P(w[e + "norm_out.weight"]), P(w[e + "norm_out.bias"])
        ds.out_w, ds.out_b = P(w[e + "conv_out.w"]), P(w[e + "conv_out.bias"])
        ds.pq_w, ds.pq_b = P(w["quant_conv.w"]), P(w["quant_conv.bias"])
        layers = []
        for lvl in range(self.n_levels):
            for blk in range(self.n_res):
                layers.append(("res", f"{e}down.{lvl}.block.{blk}"))
                if f"{e}down.{lvl}.attn.{blk}.norm.weight" in w:
                    layers.append(("attn", f"{e}down.{lvl}.attn.{blk}"))
            if lvl in self.down_layers:
                layers.append(("down", f"{e}down.{lvl}.downsample.conv"))
        layers += [("res", e + "mid.block_1"), ("attn", e + "mid.attn_1"), ("res", e + "mid.block_2")]
        assert len(layers) <= VAE_MAX_OPS
        ds.n_ops = len(layers)
        for o, (kind, n) in zip(ds.ops, layers):
            self._fill_op(o, kind, n)
        self._vae_desc_enc = ds
        return ds

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """mel fp32 [B, in_ch, T] -> moments fp32 [B, 2*embed, T/2^n] (mean | logvar): one call of the model-level C
        entry point vb_vae_encode (`encode_moments_py` = the same launches from Python)."""
        import ctypes as C
        from . import _lib
        B, Cin, T = x.shape
        assert T % (2 ** len(self.down_layers)) == 0, "stride-2 downsample expects an even length"
        x = x.to(self.dev, F32).contiguous()
        ds = self._desc_enc()
        need = int(_lib.lib().vb_vae_encode_workspace_bytes(C.byref(ds), B, T))
        if need < 0:
            _lib.check(-1, "vb_vae_encode_workspace_bytes")
        ws = torch.empty(need, device=self.dev, dtype=torch.uint8)
        mom = torch.empty(B, ds.zq_c, T // 2 ** len(self.down_layers), device=self.dev, dtype=F32)
        _lib.check(_lib.lib().vb_vae_encode(C.byref(ds), ops._ptr(x), B, T, ops._ptr(mom), ops._ptr(ws), need, ops._stream()),
                   "vb_vae_encode")
        return mom

    @torch.no_grad()
    def encode_moments_py(self, x: torch.Tensor) -> torch.Tensor:
        """The launch sequence of `encode_moments`, op by op from Python."""
        B, Cin, T = x.shape
        dev = self.dev
        e = "encoder."
        x_cl = torch.zeros(B, T, Cin, device=dev, dtype=H16)
        ops.nct_to_ntc(x.to(dev, F32).contiguous(), x_cl)
        C0 = self.w[e + "conv_in.w"].shape[0]
        h = torch.empty(B, T, C0, device=dev, dtype=F32)
        self._conv(x_cl, e + "conv_in", B, T, out_f32=h)
        for lvl in range(self.n_levels):
            for blk in range(self.n_res):
                h = self._resblock(h, f"{e}down.{lvl}.block.{blk}", B, T)
                if f"{e}down.{lvl}.attn.{blk}.norm.weight" in self.w:
                    h = self._attn(h, f"{e}down.{lvl}.attn.{blk}", B, T)
            if lvl in self.down_layers:  # pad (0,1) + conv k3 stride 2 (autoencoder1d.py:306-313)
                C = h.shape[-1]
                assert T % 2 == 0, "stride-2 downsample expects an even length"
                hb = torch.empty(B, T, C, device=dev, dtype=H16)
                ops.ew(h, out_fp16=hb)
                name = f"{e}down.{lvl}.downsample.conv"
                T = T // 2
                h = torch.empty(B, T, C, device=dev, dtype=F32)
                ops.gemm(hb.view(B, T, 2 * C), self.w[name + ".w"], M=T, G1=B, kpt=2 * C // 64, dil=1, pad=0,
                         a_c2=(1, 0), bias=self.w[name + ".bias"], out_f32=h, out_ld=C, out_g=(T * C, 0))
        h = self._resblock(h, e + "mid.block_1", B, T)
        h = self._attn(h, e + "mid.attn_1", B, T)
        h = self._resblock(h, e + "mid.block_2", B, T)
        C = h.shape[-1]
        a = self._gn(h, e + "norm_out", B, T, C, ops.ACT_SILU)
        Cz = self.w[e + "conv_out.w"].shape[0]
        czp = (Cz + 7) // 8 * 8
        hz = torch.zeros(B, T, czp, device=dev, dtype=H16)
        self._conv(a, e + "conv_out", B, T, out_fp16=hz)
        Cm = self.w["quant_conv.w"].shape[0]
        mom_cl = torch.empty(B, T, Cm, device=dev, dtype=F32)
        self._conv(hz, "quant_conv", B, T, out_f32=mom_cl, Cin=Cz)
        mom = torch.empty(B, Cm, T, device=dev, dtype=F32)
        ops.ntc_to_nct(mom_cl, mom, ld=Cm)
        return mom
