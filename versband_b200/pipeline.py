"""End-to-end AccompBand inference on one B200: conditioning -> 24-step CFG flow sampler ->
VAE decode -> HiFi-GAN, i.e. the body of `gen_song` in the reference's scripts/test_final.py
(:398-421) with every network running on the versband_b200 kernels and the mel kept on the
device between the VAE and the vocoder.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .dit import DiTEngine, F32
from .sampler import FlowSampler
from .vae import VAEEngine
from .vocoder import HifiGanEngine


class AccompBandPipeline:
    def __init__(self, dit_sd: Dict[str, torch.Tensor], vae_sd: Dict[str, torch.Tensor],
                 voc_sd: Dict[str, torch.Tensor], *, dit_cfg: dict, ddconfig: dict, embed_dim: int,
                 voc_cfg: dict, scale_factor: float = 1.0, device="cuda", use_graph: bool = True,
                 n_points: int = 25, precision: str = "fp16"):
        from . import _lib
        _lib.check(_lib.lib().vb_check_device(), "vb_check_device")
        self.dev = torch.device(device)
        kw = {k: dit_cfg[k] for k in ("in_channels", "hidden_size", "depth", "num_experts", "max_len") if k in dit_cfg}
        # NB: the reference does not forward `ori_dim` (vocal2music_moe.py:489-491): c_embedder is sized by the default 1024
        self.dit = DiTEngine(dit_sd, num_heads=dit_cfg["num_heads"], device=device, precision=precision, **kw)
        self.precision = precision
        self.vae = VAEEngine(vae_sd, ddconfig, embed_dim, device=device, with_encoder=False)
        self.voc = HifiGanEngine(voc_sd, voc_cfg, device=device)
        self.sampler = FlowSampler(self.dit, n_points=n_points, use_graph=use_graph)
        self.scale_factor = float(scale_factor)
        self.use_graph = use_graph
        self._post = {}
        self._prep = {}
        self.hop = 1
        for u in voc_cfg["upsample_rates"]:
            self.hop *= u

    @torch.no_grad()
    def generate(self, midi: torch.Tensor, beats: torch.Tensor, caption: torch.Tensor,
                 caption_uncond: torch.Tensor, *, guidance: float = 3.0, x0: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None, generator=None, events: Optional[dict] = None):
        """Device tensors in, device tensors out.
        midi/beats int64 [B,1,T_mel]; caption(_uncond) fp32 [B,L,ori_dim] (FLAN-T5 states).
        `generator`: one torch.Generator for the whole batch (the reference's semantics: one stream, draws
        shaped by the batch) or a LIST of B generators = per-clip seeding: clip c then gets exactly the latent
        noise and router draws of a batch-1 run with generator[c], whatever batch or GPU it lands on
        (SURVEY.md §8e).
        Returns dict(latent [B,C,T], mel [B,80,2T], wav [B, 2T*hop])."""
        dev = self.dev
        B, _, Tm = midi.shape
        T = (Tm + 1) // 2  # scripts/test_final.py:389 — latent length = ceil(T_mel / 2)
        mark = (lambda k: events.__setitem__(k, _ev())) if events is not None else (lambda k: None)
        mark("start")
        cond2 = self._prepare(midi, beats, caption, caption_uncond, T)
        if x0 is None:
            if isinstance(generator, (list, tuple)):
                x0 = torch.cat([torch.randn(1, self.dit.C, T, device=dev, dtype=F32, generator=g) for g in generator], 0)
            else:
                x0 = torch.randn(B, self.dit.C, T, device=dev, dtype=F32, generator=generator)
        mark("sampler_start")
        z = self.sampler.sample_cfg(x0, cond2, None, guidance, noise=noise, generator=generator)
        mark("sampler_end")
        if not self.use_graph:
            mel = self.vae.decode(z, scale=1.0 / self.scale_factor)  # ddpm_audio.py:386
            mark("vae_end")
            wav = self.voc.forward(mel)
            mark("end")
            return {"latent": z, "mel": mel, "wav": wav}
        # VAE decode and vocoder replayed from CUDA graphs per (B, T): two graphs, so that the stage boundary
        # is an observable point on the stream (bench.py records an event there; one extra graph launch)
        st = self._post.get((B, T))
        if st is None:
            st = {"z": z.clone()}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up: kernel attributes, allocator pools
                self.voc.forward(self.vae.decode(st["z"], scale=1.0 / self.scale_factor))
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                st["mel"] = self.vae.decode(st["z"], scale=1.0 / self.scale_factor)
            with torch.cuda.graph(g2, pool=g1.pool()):
                st["wav"] = self.voc.forward(st["mel"])
            st["graph_vae"], st["graph_voc"] = g1, g2
            self._post[(B, T)] = st
        st["z"].copy_(z)
        st["graph_vae"].replay()
        mark("vae_end")
        st["graph_voc"].replay()
        mark("end")
        return {"latent": z, "mel": st["mel"].clone(), "wav": st["wav"].clone()}

    def _prepare(self, midi, beats, caption, caption_uncond, T):
        """Step-invariant conditioning of the CFG pair; replayed from a CUDA graph per input shape."""
        if not self.use_graph:
            return self.dit.prepare_pair(midi, beats, caption, caption_uncond, T)
        key = (tuple(midi.shape), tuple(caption.shape), T)
        st = self._prep.get(key)
        if st is None:
            st = {"in": [t.clone() for t in (midi, beats, caption.float(), caption_uncond.float())]}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.dit.prepare_pair(*st["in"], T)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["cond2"] = self.dit.prepare_pair(*st["in"], T)
            st["graph"] = g
            self._prep[key] = st
        for dst, src in zip(st["in"], (midi, beats, caption, caption_uncond)):
            dst.copy_(src)
        st["graph"].replay()
        return st["cond2"]

    @torch.no_grad()
    def generate_host(self, midi, beats, caption, caption_uncond, *, guidance=3.0, x0=None, noise=None,
                      generator=None, out_wav: Optional[torch.Tensor] = None):
        """The reference-facing call: HOST (pinned) tensors in, host waveform out; the H2D / D2H
        copies are issued on the current stream around `generate`."""
        dev = self.dev
        to = lambda t: None if t is None else t.to(dev, non_blocking=True)  # noqa: E731
        out = self.generate(to(midi), to(beats), to(caption), to(caption_uncond), guidance=guidance,
                            x0=to(x0), noise=noise, generator=generator)
        wav = out["wav"]
        if out_wav is None:
            out_wav = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
        out_wav.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_wav


def _ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e
