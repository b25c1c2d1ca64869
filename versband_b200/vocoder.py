"""HiFi-GAN generator on B200: every Conv1d / ConvTranspose1d is an implicit-GEMM vb_gemm launch
over channels-last fp16 activations; LeakyReLU, bias, residual add, MRF sum and the 1/3 mean are
fused into the producing GEMM's epilogue.

Mirrors reference vocoder/hifigan/modules/hifigan.py (HifiGanGenerator.forward :126-143,
ResBlock1.forward :51-58).  Generic over upsample rates / kernel sizes (the reference reads them
from the checkpoint's config.yaml, vocoder/hifigan/hifigan.py:11-15).  weight_norm is folded at
pack time (w = g * v / ||v||).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import ops
from .dit import H16, F32, pack_conv_weight

LRELU = 0.1


def fold_weight_norm(sd: Dict[str, torch.Tensor], pre: str) -> torch.Tensor:
    if (pre + "weight") in sd:
        return sd[pre + "weight"].float()
    v, g = sd[pre + "weight_v"].float(), sd[pre + "weight_g"].float()
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


class HifiGanEngine:
    def __init__(self, sd: Dict[str, torch.Tensor], h: dict, device="cuda"):
        assert str(h["resblock"]) == "1", "only ResBlock1 generators are implemented"
        self.dev = torch.device(device)
        self.h = dict(h)
        self.rates: List[int] = list(h["upsample_rates"])
        self.up_k: List[int] = list(h["upsample_kernel_sizes"])
        self.rb_k: List[int] = list(h["resblock_kernel_sizes"])
        self.rb_d: List[List[int]] = [list(d) for d in h["resblock_dilation_sizes"]]
        self.C0 = h["upsample_initial_channel"]
        self.w: Dict[str, torch.Tensor] = {}
        self.fuse_pairs = True  # ResBlock1 conv pairs of the C <= 64 stages in one kernel each (vb_resblock_pair)
        self._c_desc = None
        self._pack({k: v.detach().to(self.dev) for k, v in sd.items()})

    def _pack(self, sd):
        w = self.w
        w["pre.w"] = pack_conv_weight(fold_weight_norm(sd, "conv_pre."))
        w["pre.b"] = sd["conv_pre.bias"].float().contiguous()
        self.n_mel = sd["conv_pre.weight_v" if "conv_pre.weight_v" in sd else "conv_pre.weight"].shape[1]
        for i, (u, k) in enumerate(zip(self.rates, self.up_k)):
            wt = fold_weight_norm(sd, f"ups.{i}.")  # [Cin, Cout, k]
            cin, cout, _ = wt.shape
            Q = (k + u - 1) // u
            kp = (cin + 63) // 64 * 64
            wp = torch.zeros(u, cout, Q, kp, device=self.dev, dtype=H16)
            for r in range(u):
                for q in range(Q):
                    j = r + u * q
                    if j < k:  # tap q of phase r reads x[s - q]; stored reversed so time = s-(Q-1)+tap
                        wp[r, :, Q - 1 - q, :cin] = wt[:, :, j].t().to(H16)
            w[f"up{i}.w"] = wp.reshape(u, cout, Q * kp).contiguous()
            w[f"up{i}.b"] = sd[f"ups.{i}.bias"].float().contiguous()
        nk = len(self.rb_k)
        for i in range(len(self.rates)):
            for j in range(nk):
                pre = f"resblocks.{i * nk + j}."
                for c in range(len(self.rb_d[j])):
                    for which in ("convs1", "convs2"):
                        w[f"rb{i}.{j}.{which}.{c}.w"] = pack_conv_weight(fold_weight_norm(sd, f"{pre}{which}.{c}."))
                        w[f"rb{i}.{j}.{which}.{c}.b"] = sd[f"{pre}{which}.{c}.bias"].float().contiguous()
        wpost = fold_weight_norm(sd, "conv_post.")  # [1, C, k]
        assert wpost.shape[0] == 1
        w["post.w"] = wpost[0].t().contiguous()  # [k, C]
        self.post_b = float(sd["conv_post.bias"].float().item())
        self.post_k = wpost.shape[-1]

    def _conv(self, a, wname, B, T, k, dil, **epi):
        w = self.w[wname + ".w"]
        Cout = w.shape[0]
        kpt = w.shape[1] // k // 64
        ops.gemm(a, w, M=T, G1=B, kpt=kpt, dil=dil, pad=(k * dil - dil) // 2, a_c2=(1, 0),
                 bias=self.w[wname + ".b"], out_ld=Cout, out_g=(T * Cout, 0), outb_ld=Cout,
                 outb_g=(T * Cout, 0), **epi)

    # ------------------------------------------------------------------ C entry point (vb_voc_forward)
    def _desc(self):
        """The packed generator as a `vb_voc_desc` (device pointers into self.w; built once)."""
        if self._c_desc is None:
            from ._lib import VocDesc
            d, w = VocDesc(), self.w
            nk = len(self.rb_k)
            assert len(self.rates) <= 8 and nk <= 3 and all(len(x) <= 4 for x in self.rb_d)
            d.n_mel, d.c0, d.pre_k, d.n_stages, d.n_kernels, d.post_k = self.n_mel, self.C0, 7, len(self.rates), nk, self.post_k
            d.lrelu, d.post_b = LRELU, self.post_b
            ptr = lambda t: t.data_ptr()  # noqa: E731
            d.pre_w, d.pre_b, d.post_w = ptr(w["pre.w"]), ptr(w["pre.b"]), ptr(w["post.w"])
            for i, (u, k) in enumerate(zip(self.rates, self.up_k)):
                d.rates[i], d.up_k[i], d.up_cout[i] = u, k, w[f"up{i}.w"].shape[1]
                d.up_w[i], d.up_b[i] = ptr(w[f"up{i}.w"]), ptr(w[f"up{i}.b"])
                for j in range(nk):
                    for c in range(len(self.rb_d[j])):
                        d.rb_w1[i][j][c], d.rb_b1[i][j][c] = ptr(w[f"rb{i}.{j}.convs1.{c}.w"]), ptr(w[f"rb{i}.{j}.convs1.{c}.b"])
                        d.rb_w2[i][j][c], d.rb_b2[i][j][c] = ptr(w[f"rb{i}.{j}.convs2.{c}.w"]), ptr(w[f"rb{i}.{j}.convs2.{c}.b"])
            for j in range(nk):
                d.rb_k[j], d.n_dil[j] = self.rb_k[j], len(self.rb_d[j])
                for c, dd in enumerate(self.rb_d[j]):
                    d.rb_d[j][c] = dd
            self._c_desc = d
        return self._c_desc

    @torch.no_grad()
    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        """mel fp32 [B, n_mel, T] -> wav fp32 [B, T * prod(rates)]: ONE call into the C-ABI (vb_voc_forward issues
        the whole launch sequence from C++); `forward_py` below is the same sequence issued launch by launch from
        Python (kept for experiments: fuse_pairs switch, per-launch timing)."""
        import ctypes as C
        from . import _lib
        if not self.fuse_pairs:
            return self.forward_py(mel)
        B, nm, T = mel.shape
        assert nm == self.n_mel and nm % 8 == 0
        mel = mel.to(self.dev, F32).contiguous()
        hop = 1
        for u in self.rates:
            hop *= u
        d = self._desc()
        need = int(_lib.lib().vb_voc_workspace_bytes(C.byref(d), B, T))
        if need < 0:
            _lib.check(-1, "vb_voc_workspace_bytes")
        ws = torch.empty(need, device=self.dev, dtype=torch.uint8)
        wav = torch.empty(B, T * hop, device=self.dev, dtype=F32)
        _lib.check(_lib.lib().vb_voc_forward(C.byref(d), C.c_void_p(mel.data_ptr()), B, T, C.c_void_p(wav.data_ptr()),
                                             C.c_void_p(ws.data_ptr()), need, ops._stream()), "vb_voc_forward")
        return wav

    @torch.no_grad()
    def forward_py(self, mel: torch.Tensor) -> torch.Tensor:
        """mel fp32 [B, n_mel, T] -> wav fp32 [B, T * prod(rates)] (launch sequence issued from Python)."""
        dev = self.dev
        B, nm, T = mel.shape
        assert nm == self.n_mel and nm % 8 == 0
        mel_cl = torch.empty(B, T, nm, device=dev, dtype=H16)
        ops.nct_to_ntc(mel.to(dev, F32).contiguous(), mel_cl)
        C = self.C0
        xb = torch.empty(B, T, C, device=dev, dtype=H16)  # lrelu(conv_pre(mel)), hifigan.py:128-130
        self._conv(mel_cl, "pre", B, T, 7, 1, out_fp16=xb, act_out=ops.ACT_LRELU, act_out_slope=LRELU)
        nk = len(self.rb_k)
        n_stages = len(self.rates)
        for i, (u, k) in enumerate(zip(self.rates, self.up_k)):
            # ---- ConvTranspose1d as u polyphase GEMMs in one launch (hifigan.py:131)
            wup = self.w[f"up{i}.w"]
            Cout = wup.shape[1]
            Q = (k + u - 1) // u
            kp = wup.shape[2] // Q
            p = (k - u) // 2
            Tout = (T - 1) * u - 2 * p + k
            # Activations live in HBM ONLY as y = leaky_relu_0.1(x) in fp16 (2 B/element): y is directly the next
            # conv's A operand (TMA -> tensor core, no elementwise pass), and the pre-activation x that the residual
            # connections need is recovered exactly-to-rounding as x = y >= 0 ? y : 10 y in the consuming epilogue
            # (vb_gemm res16 / res16_neg). No fp32 copy of the residual stream exists any more.
            xl = torch.empty(B, Tout, Cout, device=dev, dtype=H16)
            ops.gemm(xb, wup, M=T + Q - 1, N=Cout, K=Q * kp, G1=u, G2=B, kpt=kp // 64, dil=1, pad=Q - 1,
                     a_c2=(0, 1), b_c2=(1, 0, 0), bias=self.w[f"up{i}.b"],
                     row_mul=u, row_off=-p, row_off_g1=1, rows_out=Tout,
                     out_fp16=xl, outb_ld=Cout, outb_g=(0, Tout * Cout), act_out=ops.ACT_LRELU,
                     act_out_slope=LRELU)
            T, C = Tout, Cout
            # ---- MRF: mean of nk ResBlock1 chains that all start from x (hifigan.py:132-138)
            last_stage = i == n_stages - 1
            nxt = torch.empty(B, T, C, device=dev, dtype=H16)
            t1 = torch.empty(B, T, C, device=dev, dtype=H16)
            rlb = [torch.empty(B, T, C, device=dev, dtype=H16) for _ in range(2)]
            rout = [torch.empty(B, T, C, device=dev, dtype=H16) for _ in range(nk)]  # each chain's result (activated repr.)
            assert nk <= 3, "vb_mrf_combine takes up to three resblock outputs"
            for j in range(nk):
                kk, dils = self.rb_k[j], self.rb_d[j]
                rl = xl
                for c, d in enumerate(dils):
                    dst = rout[j] if c == len(dils) - 1 else rlb[c % 2]
                    n1, n2 = f"rb{i}.{j}.convs1.{c}", f"rb{i}.{j}.convs2.{c}"
                    if self.fuse_pairs and C in (32, 64) and ops.resblock_pair_fits(C, kk, d):
                        # both convs, the activation between them and the residual add in ONE kernel (K18)
                        ops.resblock_pair(rl, dst, self.w[n1 + ".w"], self.w[n2 + ".w"], self.w[n1 + ".b"], self.w[n2 + ".b"],
                                          B=B, T=T, C=C, k=kk, dil=d, slope=LRELU)
                    else:
                        self._conv(rl, n1, B, T, kk, d, out_fp16=t1, act_out=ops.ACT_LRELU, act_out_slope=LRELU)
                        self._conv(t1, n2, B, T, kk, 1, res16=rl, res16_neg=1.0 / LRELU,
                                   out_fp16=dst, act_out=ops.ACT_LRELU, act_out_slope=LRELU)
                    rl = dst
            # mean of the chains + the next layer's activation (slope 0.1 before ups, 0.01 before conv_post :139)
            ops.mrf_combine(rout, nxt, neg=1.0 / LRELU, scale=1.0 / nk, slope_out=(0.01 if last_stage else LRELU))
            xb = nxt
        wav = torch.empty(B, T, device=dev, dtype=F32)
        ops.conv_post(xb, self.w["post.w"], self.post_b, B, T, C, self.post_k, wav)
        return wav
