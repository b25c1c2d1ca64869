"""Flow-matching sampler: fixed-step Euler over t_span = linspace(0, 1, 25) with classifier-free
guidance — reference ldm/models/diffusion/cfm1_audio_sampler.py:87-116 and cfm1_audio.py:145-162
(`Wrapper_cfg`), torchdyn's Euler restated as in SURVEY.md §8c.

B200 specifics: the conditional and unconditional DiT evaluations of a step run as ONE batch of
2B clips (weights stream from HBM once), the CFG mix + Euler update is one kernel, and the whole
24-step loop — including the router's Exp(1) draws, taken from torch's Philox generator with the
reference's shapes and order — is replayed from a CUDA graph.
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


class FlowSampler:
    def __init__(self, dit: DiTEngine, n_points: int = 25, use_graph: bool = True):
        self.dit = dit
        self.sched = euler_schedule(n_points)
        self.use_graph = use_graph
        # cond / uncond evaluations as two concurrent launch chains (_run_two_streams). Off: measured 12.0 ms vs 10.5 ms
        # for the batched form — twice the launches, and the half-size kernels are not proportionally shorter
        self.concurrent_cfg = False
        self._state: Dict[tuple, dict] = {}
        self._native_noise: Optional[bool] = None  # None = not verified yet

    # ---------------------------------------------------------------- native Philox draws
    def native_noise_ok(self) -> bool:
        """One-time self-check: vb_router_noise must reproduce torch's `exponential_()` stream
        bit for bit (same seed/offset) and torch must advance its offset by 4 per call; otherwise the
        sampler keeps drawing with torch inside the graph."""
        if self._native_noise is None:
            try:
                dev, d, E = self.dit.dev, self.dit.depth, self.dit.E
                gen = torch.Generator(device=dev)
                gen.manual_seed(0x5EED1234)
                ok = True
                for N in (37, 1500):
                    st = dict(N=N, e0=torch.zeros(2, d, 2 * N, 2, device=dev), e1=torch.zeros(2, d, 2 * N, E, device=dev),
                              e2=torch.zeros(2, d, 2 * N, E, device=dev))
                    off0 = gen.get_offset()
                    for s_ in range(2):
                        self._draw_step(st, s_, gen)
                    calls = 2 * 2 * d * 3
                    ok = ok and (gen.get_offset() - off0 == 4 * calls)
                    mine = [torch.zeros_like(st[k]) for k in ("e0", "e1", "e2")]
                    ops.router_noise(gen.initial_seed(), off0, N, E, d, 2, *mine)
                    ok = ok and all(torch.equal(a, st[k]) for a, k in zip(mine, ("e0", "e1", "e2")))
                self._native_noise = bool(ok)
            except Exception:
                self._native_noise = False
        return self._native_noise

    def _draw_native(self, st: dict, generator) -> bool:
        """All Exp(1) draws of a run in one launch, consuming `generator` (or the default generator of the
        engine's device — the CURRENT device when the engine was built with a bare "cuda") exactly like the
        reference's sequence of `exponential_()` calls. Returns False (nothing consumed) if the library
        rejects the shape, in which case the caller draws with torch."""
        if generator is not None:
            gen = generator
        else:
            idx = self.dit.dev.index if self.dit.dev.index is not None else torch.cuda.current_device()
            gen = torch.cuda.default_generators[idx]
        off = gen.get_offset()
        S, d = len(self.sched), self.dit.depth
        try:
            ops.router_noise(gen.initial_seed(), off, st["N"], self.dit.E, d, S, st["e0"], st["e1"], st["e2"])
        except Exception:
            return False
        gen.set_offset(off + 4 * S * 2 * d * 3)
        return True

    def _draw_per_clip(self, st: dict, B: int, T: int, generators) -> None:
        """Router draws of a batch of B clips from B generators: each clip's [steps, depth, 2T, k] block is
        drawn as a batch-1 run (cond rows then uncond rows) and placed at the clip's rows of the batched
        layout (cond rows of all clips first, then uncond rows)."""
        S, d, E, dev = len(self.sched), self.dit.depth, self.dit.E, self.dit.dev
        native = self.native_noise_ok()
        for c, gen in enumerate(generators):
            one = dict(N=T, e0=torch.empty(S, d, 2 * T, 2, device=dev, dtype=F32),
                       e1=torch.empty(S, d, 2 * T, E, device=dev, dtype=F32),
                       e2=torch.empty(S, d, 2 * T, E, device=dev, dtype=F32))
            if not (native and self._draw_native(one, gen)):
                for s_ in range(S):
                    self._draw_step(one, s_, gen)
            for name in ("e0", "e1", "e2"):
                st[name][:, :, c * T:(c + 1) * T] = one[name][:, :, :T]
                st[name][:, :, (B + c) * T:(B + c + 1) * T] = one[name][:, :, T:]

    # ---------------------------------------------------------------- noise
    def _draw_step(self, st: dict, s: int, generator) -> None:
        """Exp(1) draws of Euler step s into the static buffers, in the reference's consumption
        order: cond evaluation then uncond evaluation; per evaluation block 0..depth-1; per block
        [N,2] (group gate), [N,E] (caption gate), [N,E] (acoustic gate) — vocal2music_moe.py:134,150,151.
        Each draw is its own `exponential_` call on a contiguous [N,k] view, which consumes the
        Philox stream exactly like the reference's `torch.empty_like(logits).exponential_()`."""
        N = st["N"]
        for br in range(2):
            rows = slice(br * N, (br + 1) * N)
            for k in range(self.dit.depth):
                st["e0"][s, k, rows].exponential_(generator=generator)
                st["e1"][s, k, rows].exponential_(generator=generator)
                st["e2"][s, k, rows].exponential_(generator=generator)

    def draw_noise(self, B: int, T: int, generator=None) -> torch.Tensor:
        """All draws of one sampling run as a tensor [steps, depth, 2N, 2+2E] (cond rows first)."""
        d, E, N = self.dit.depth, self.dit.E, B * T
        st = dict(N=N, e0=torch.empty(len(self.sched), d, 2 * N, 2, device=self.dit.dev, dtype=F32),
                  e1=torch.empty(len(self.sched), d, 2 * N, E, device=self.dit.dev, dtype=F32),
                  e2=torch.empty(len(self.sched), d, 2 * N, E, device=self.dit.dev, dtype=F32))
        for s in range(len(self.sched)):
            self._draw_step(st, s, generator)
        return torch.cat([st["e0"], st["e1"], st["e2"]], -1)

    # ---------------------------------------------------------------- loop body
    def _c_state(self, st: dict) -> None:
        """Workspace of vb_sample_cfg, allocated and zero-filled ONCE (outside any graph capture: it carries the MoE
        bucket counters from one evaluation to the next)."""
        if "sws" in st:
            return
        import ctypes as C
        from . import _lib
        dit, cond2 = self.dit, st["cond2"]
        S = len(self.sched)
        d = dit._desc()
        need = int(_lib.lib().vb_sample_workspace_bytes(C.byref(d), cond2.B // 2, cond2.T, S))
        if need < 0:
            _lib.check(-1, "vb_sample_workspace_bytes")
        st["sws"] = torch.zeros(need, device=dit.dev, dtype=torch.uint8)
        st["dt_host"] = (C.c_float * S)(*[dt for _, dt in self.sched])
        st["lb"] = torch.zeros(dit.depth, device=dit.dev, dtype=F32)

    def _run_c(self, st: dict, scale: float) -> None:
        """The whole loop as ONE call of the model-level C entry point vb_sample_cfg (csrc/vb_model.cu)."""
        import ctypes as C
        from . import _lib
        from ._lib import SampleArgs
        dit, cond2 = self.dit, st["cond2"]
        B, S = cond2.B // 2, len(self.sched)
        d = dit._desc()
        self._c_state(st)
        P = ops._ptr
        a = SampleArgs()
        a.n_steps, a.Bc, a.guidance = S, B, float(scale)
        a.t_all, a.dt = P(st["t_all"]), st["dt_host"]
        a.e0, a.e1, a.e2, a.x, a.lb = P(st["e0"]), P(st["e1"]), P(st["e2"]), P(st["x"]), P(st["lb"])
        _lib.check(_lib.lib().vb_sample_cfg(C.byref(d), C.byref(dit._cond_struct(cond2)), C.byref(a), P(st["sws"]),
                                            st["sws"].numel(), ops._stream()), "vb_sample_cfg")

    def _run(self, st: dict, scale: float, rng: bool, generator) -> None:
        dit = self.dit
        if not (rng or self.concurrent_cfg or dit.use_chain or dit.fused_dispatch or ops.TIMELINE is not None):
            return self._run_c(st, scale)
        return self._run_py(st, scale, rng, generator)

    def _run_py(self, st: dict, scale: float, rng: bool, generator) -> None:
        """The loop issued from Python (torch RNG inside the loop, or one of the experimental launch variants)."""
        d = self.dit.depth
        x, cond2, N = st["x"], st["cond2"], st["N"]
        # adaLN / gate vectors of all steps in one batch of S*2B rows (they only depend on t and caption)
        S, B2 = st["t_all"].shape
        self.dit.time_conditioning(st["t_all"].reshape(-1), cond2.cap_emb.repeat(S, 1), st["mod_all"].view(S * B2, -1),
                                   st["hl_all"].view(S * B2, -1))
        st["chain_cnt"].zero_()  # dependency counters of every chained launch of the run (one memset)
        if self.concurrent_cfg and not rng:
            return self._run_two_streams(st, scale)
        for s, (_, dt) in enumerate(self.sched):
            if rng:
                self._draw_step(st, s, generator)
            expo = [(st["e0"][s, k], st["e1"][s, k], st["e2"][s, k]) for k in range(d)]
            v, _ = self.dit.forward(x, st["t_all"][s], cond2, expo, pre=(st["mod_all"][s], st["hl_all"][s]),
                                    x_cl_ready=s > 0, chain_cnt=st["chain_cnt"][s])
            # CFG + Euler update; the same launch rewrites the DiT's channels-last copies of x for the next step
            ops.ntc_to_nct(v[:N], x, e_u=v[N:], ld=v.shape[-1], s=scale, dt=dt,
                           x_cl=self.dit._workspace(cond2.B, cond2.T)["x_cl"])

    def _run_two_streams(self, st: dict, scale: float) -> None:
        """The conditional and the unconditional evaluation of a step as TWO CONCURRENT launch chains (two streams,
        i.e. two parallel branches of the captured graph) instead of one batch of 2B clips. At batch 1 every kernel
        of an evaluation is a ~10 us, latency-bound launch that fills at most ~2/3 of the SMs; a chain of half the
        rows needs half the CTAs, so the two chains run side by side and each one's launch / pipeline-fill /
        epilogue latencies hide under the other's mainloops. They join at the CFG mix + Euler update."""
        d, dit = self.dit.depth, self.dit
        x, cond2, N = st["x"], st["cond2"], st["N"]
        B = cond2.B // 2
        halves = cond2.split(2)
        if "x_cl2" not in st:
            st["x_cl2"] = torch.zeros(2 * B, cond2.T, 32, device=dit.dev, dtype=dit.adt)
            st["side"] = torch.cuda.Stream()
            st["v2"] = [None, None]
        x_cl, side = st["x_cl2"], st["side"]
        main = torch.cuda.current_stream()
        for s, (_, dt) in enumerate(self.sched):
            if s == 0:
                for r in range(2):
                    ops.nct_to_ntc(x.contiguous(), x_cl[r * B:(r + 1) * B])
            side.wait_stream(main)
            vs = []
            for br, stream in ((0, main), (1, side)):
                with torch.cuda.stream(stream):
                    rows = slice(br * N, (br + 1) * N)
                    expo = [(st["e0"][s, k, rows], st["e1"][s, k, rows], st["e2"][s, k, rows]) for k in range(d)]
                    v, _ = dit.forward(x, st["t_all"][s, br * B:(br + 1) * B], halves[br], expo,
                                       pre=(st["mod_all"][s, br * B:(br + 1) * B], st["hl_all"][s, br * B:(br + 1) * B]),
                                       x_cl_ready=True, chain_cnt=None, ws_slot=1 + br,
                                       x_cl=x_cl[br * B:(br + 1) * B])
                    vs.append(v)
            main.wait_stream(side)
            ops.ntc_to_nct(vs[0], x, e_u=vs[1], ld=vs[0].shape[-1], s=scale, dt=dt, x_cl=x_cl)

    @torch.no_grad()
    def sample_cfg(self, x0: torch.Tensor, cond: DiTCond, uncond: Optional[DiTCond], scale: float,
                   noise: Optional[torch.Tensor] = None, generator=None) -> torch.Tensor:
        """x0 fp32 [B, C, T] -> x1 [B, C, T].  With `noise` (from draw_noise) the run is a pure
        function of its arguments; otherwise the draws come from `generator` (default CUDA
        generator if None).  `uncond=None` means `cond` already is the stacked (cond, uncond) pair of
        2B clips (DiTEngine.prepare_pair); a pair object that stays the same between calls is used in
        place (no copies), which is how the pipeline feeds the sampler graph from its prepare graph."""
        dev, d, E = self.dit.dev, self.dit.depth, self.dit.E
        B, C, T = x0.shape
        N = B * T
        rng = noise is None
        key = (B, T, cond.L)
        pair = cond if uncond is None else None
        if pair is not None:
            assert pair.B == 2 * B
        st = self._state.get(key)
        if st is None:
            S = len(self.sched)
            st = dict(N=N, x=torch.empty(B, C, T, device=dev, dtype=F32),
                      cond2=pair if pair is not None else DiTCond.cat([cond, uncond]),
                      e0=torch.empty(S, d, 2 * N, 2, device=dev, dtype=F32),
                      e1=torch.empty(S, d, 2 * N, E, device=dev, dtype=F32),
                      e2=torch.empty(S, d, 2 * N, E, device=dev, dtype=F32),
                      t_all=torch.tensor([[t] * (2 * B) for t, _ in self.sched], device=dev, dtype=torch.int64),
                      mod_all=torch.empty(S, 2 * B, self.dit.mod_ld, device=dev, dtype=F32),
                      hl_all=torch.empty(S, 2 * B, 2 * d, device=dev, dtype=F32),
                      chain_cnt=torch.zeros(S, d, self.dit._workspace(2 * B, T)["chain_len"], device=dev, dtype=torch.int32),
                      graphs={})
            self._state[key] = st
        elif pair is None:
            st["cond2"].copy_from(DiTCond.cat([cond, uncond]))
        elif pair is not st["cond2"]:
            st["cond2"].copy_from(pair)
        st["x"].copy_(x0)
        if not torch.cuda.is_current_stream_capturing():
            self._c_state(st)
        if rng and isinstance(generator, (list, tuple)):
            # per-clip seeding: clip c consumes generator[c] exactly as a batch-1 run would
            assert len(generator) == B
            self._draw_per_clip(st, B, T, generator)
            rng, noise, generator = False, None, None
        elif rng and self.native_noise_ok():
            # all 576 Exp(1) draws in one launch, bit-identical to torch's stream; the graph below then
            # runs without RNG nodes
            if self._draw_native(st, generator):
                rng = False
                noise = None
        elif not rng:
            st["e0"].copy_(noise[..., 0:2])
            st["e1"].copy_(noise[..., 2:2 + E])
            st["e2"].copy_(noise[..., 2 + E:])
        if isinstance(generator, (list, tuple)):  # (native self-check failed and per-clip draws not taken)
            generator = generator[0]
        if not self.use_graph:
            self._run(st, scale, rng, generator)
            return st["x"].clone()
        gkey = (rng, float(scale), id(generator) if rng else 0)
        g = st["graphs"].get(gkey)
        if g is None:
            # warm-up outside capture (kernel attributes, workspaces, tile plans of the shapes actually launched), then capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            if self.concurrent_cfg and not rng:
                with torch.cuda.stream(side):  # one complete eager run on the same two-stream path (x restored after)
                    x_saved = st["x"].clone()
                    self._run(st, scale, False, None)
                    st["x"].copy_(x_saved)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
            else:
                saved0 = [st[k][0].clone() for k in ("e0", "e1", "e2")]
                with torch.cuda.stream(side):
                    for name in ("e0", "e1", "e2"):  # any positive value works for the warm-up
                        st[name][0].fill_(1.0)
                    self.dit.forward(st["x"], st["t_all"][0], st["cond2"],
                                     [(st["e0"][0, k], st["e1"][0, k], st["e2"][0, k]) for k in range(d)])
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                if not rng and noise is not None:
                    st["e0"].copy_(noise[..., 0:2]); st["e1"].copy_(noise[..., 2:2 + E]); st["e2"].copy_(noise[..., 2 + E:])
                elif not rng:
                    st["e0"][0].copy_(saved0[0]); st["e1"][0].copy_(saved0[1]); st["e2"][0].copy_(saved0[2])
            g = torch.cuda.CUDAGraph()
            if rng and generator is not None:
                g.register_generator_state(generator)
            with torch.cuda.graph(g):
                self._run(st, scale, rng, generator)
            st["graphs"][gkey] = g
        g.replay()
        return st["x"].clone()
