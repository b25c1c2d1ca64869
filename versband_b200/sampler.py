"""Flow-matching sampler: fixed-step Euler over t_span = linspace(0, 1, 25) with classifier-free
guidance — reference ldm/models/diffusion/cfm1_audio_sampler.py:87-116 and cfm1_audio.py:145-162
(`Wrapper_cfg`), torchdyn's Euler restated as in SURVEY.md §8c.

B200 specifics: the conditional and unconditional DiT evaluations of a step run as ONE batch of
2B clips (weights stream from HBM once), the CFG mix + Euler update is one kernel, the Gumbel
noise of all 48 evaluations is drawn up front with torch's generator in the reference's order,
and the whole 24-step loop is replayed from a CUDA graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .dit import DiTCond, DiTEngine, F32


def euler_schedule(n_points: int = 25) -> List[Tuple[int, float]]:
    """(t_long, dt) of each Euler step. t_long = trunc(1000 t) (cfm1_audio.py:156); dt is
    re-derived from t_span every step, all in fp32 on the host."""
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


def draw_router_noise(depth: int, n_tokens: int, E: int, device, generator=None):
    """One DiT evaluation's Exp(1) draws, shapes/order of vocal2music_moe.py:134,150,151."""
    out = []
    for _ in range(depth):
        out.append(tuple(torch.empty(n_tokens, k, device=device, dtype=F32).exponential_(generator=generator)
                         for k in (2, E, E)))
    return out


class FlowSampler:
    def __init__(self, dit: DiTEngine, n_points: int = 25, use_graph: bool = True):
        self.dit = dit
        self.sched = euler_schedule(n_points)
        self.use_graph = use_graph
        self._graphs: Dict[Tuple[int, int, int], dict] = {}

    def draw_noise(self, B: int, T: int, generator=None) -> torch.Tensor:
        """[steps, 2 (cond, uncond), depth] draws packed as one tensor [steps, depth, 2*N, 2+2E]
        (cond rows first) in the reference's consumption order: per step cond forward then
        uncond forward, per forward block 0..depth-1, per block ([N,2],[N,E],[N,E])."""
        d, E, N = self.dit.depth, self.dit.E, B * T
        buf = torch.empty(len(self.sched), d, 2 * N, 2 + 2 * E, device=self.dit.dev, dtype=F32)
        for s in range(len(self.sched)):
            for br in range(2):
                for k, (e0, e1, e2) in enumerate(draw_router_noise(d, N, E, self.dit.dev, generator)):
                    rows = slice(br * N, (br + 1) * N)
                    buf[s, k, rows, 0:2] = e0
                    buf[s, k, rows, 2:2 + E] = e1
                    buf[s, k, rows, 2 + E:] = e2
        return buf

    def _run(self, x: torch.Tensor, cond2: DiTCond, noise: torch.Tensor, scale: float, t_all: torch.Tensor):
        d, E = self.dit.depth, self.dit.E
        N2 = cond2.B * cond2.T
        N = N2 // 2
        for s, (_, dt) in enumerate(self.sched):
            # noise rows are stored [2N, 2+2E]; the kernels want dense [2N,2],[2N,E],[2N,E]
            expo = [(self._e0[s, k], self._e1[s, k], self._e2[s, k]) for k in range(d)]
            v, _ = self.dit.forward(x, t_all[s], cond2, expo)
            ops.ntc_to_nct(v[:N], x, e_u=v[N:], ld=v.shape[-1], s=scale, dt=dt)

    @torch.no_grad()
    def sample_cfg(self, x0: torch.Tensor, cond: DiTCond, uncond: DiTCond, scale: float,
                   noise: Optional[torch.Tensor] = None, generator=None) -> torch.Tensor:
        """x0 fp32 [B, C, T] -> x1 [B, C, T]. `noise` from draw_noise(); drawn here if None."""
        dev = self.dit.dev
        B, C, T = x0.shape
        if noise is None:
            noise = self.draw_noise(B, T, generator)
        E = self.dit.E
        key = (B, T, cond.L)
        st = self._graphs.get(key)
        if st is None:
            st = dict(x=torch.empty(B, C, T, device=dev, dtype=F32),
                      cond2=DiTCond.cat([cond, uncond]),
                      e0=torch.empty(*noise.shape[:3], 2, device=dev, dtype=F32),
                      e1=torch.empty(*noise.shape[:3], E, device=dev, dtype=F32),
                      e2=torch.empty(*noise.shape[:3], E, device=dev, dtype=F32),
                      t_all=torch.tensor([[t] * (2 * B) for t, _ in self.sched], device=dev, dtype=torch.int64),
                      scale=None, graph=None)
            self._graphs[key] = st
        else:
            st["cond2"].copy_from(DiTCond.cat([cond, uncond]))
        st["x"].copy_(x0)
        st["e0"].copy_(noise[..., 0:2])
        st["e1"].copy_(noise[..., 2:2 + E])
        st["e2"].copy_(noise[..., 2 + E:])
        self._e0, self._e1, self._e2 = st["e0"], st["e1"], st["e2"]
        if not self.use_graph:
            self._run(st["x"], st["cond2"], noise, scale, st["t_all"])
        else:
            if st["graph"] is None or st["scale"] != scale:
                # warm-up outside capture (sets kernel attributes, sizes workspaces), then capture
                self.dit._workspace(2 * B, T)
                xsave = st["x"].clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    v, _ = self.dit.forward(st["x"], st["t_all"][0], st["cond2"],
                                            [(st["e0"][0, k], st["e1"][0, k], st["e2"][0, k]) for k in range(self.dit.depth)])
                torch.cuda.current_stream().wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run(st["x"], st["cond2"], noise, scale, st["t_all"])
                st["graph"], st["scale"] = g, scale
                st["x"].copy_(xsave)
            st["graph"].replay()
        return st["x"].clone()
