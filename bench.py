#!/usr/bin/env python
"""Benchmark of the AccompBand inference hot path (BASELINE.json metric: 24 kHz audio-seconds per
wall-second, plus DiT latent-frame evaluations per second).

  python bench.py --gpus N --steps K --warmup W                 # this repo (CUDA, C-ABI kernels)
  python bench.py --impl reference --gpus N --steps K --warmup W # the reference algorithm on host cores

One "step" = one batch of synthetic 20 s clips through conditioning -> 24-step CFG flow sampler
(48 DiT evaluations) -> VAE decode -> HiFi-GAN. Workload at N GPUs: `--clips` clips per GPU
(default 1 = BASELINE configs[1], "25-step sampler, batch 1, 20 s clip, 1xB200"), weak scaling,
no data-path collective (clips are independent). Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 20.0
T_MEL = 1500
T_LAT = 750
N_EVALS = 48  # 24 Euler steps x (cond, uncond)
# Algorithmic FLOPs per 20 s clip: 48 x 51.0 + 22 (sampler) + 160.3 (VAE) + 848.5 (HiFi-GAN)
FLOP_SAMPLER = 48 * 51.0e9 + 22e9
FLOP_ATTN = 48 * 4 * 2 * (0.864e9 + 0.092e9)  # self + text attention core (GMAC per block) -> 0.367 TFLOP
FLOP_VAE = 160.3e9
FLOP_VOC = 848.5e9
FLOP_CLIP = FLOP_SAMPLER + FLOP_VAE + FLOP_VOC
BYTES_VOC_16 = 3.18e9  # HiFi-GAN per-layer compulsory bytes at 16-bit storage


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def synth_weights():
    from oracle import synth
    return (synth.synth_state_dict(synth.load_manifest("dit"), seed=1234),
            synth.synth_state_dict(synth.load_manifest("vae"), seed=2345),
            synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN))


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_step(sds, euler_steps: int = 24, seed: int = 0):
    """One clip on the CPU through the oracle port of the reference algorithm: `euler_steps` Euler steps
    (default: all 24; 2 DiT evaluations each, dense experts as in the reference) + full VAE decode + full
    HiFi-GAN. With fewer than 24 steps the sampler time is scaled up and the result is marked as such by the
    caller. Returns (seconds per clip, breakdown)."""
    from oracle import accompband_ref as R, synth
    torch.set_grad_enabled(False)
    x, midi, beats, cap_c, cap_u = synth.dit_inputs(1, T_LAT, seed=seed)
    ctx_c = {"c_crossattn": cap_c, "c_concat": {"midi": midi, "beats": beats}}
    ctx_u = {"c_crossattn": cap_u, "c_concat": {"midi": midi, "beats": beats}}
    sched = R.euler_schedule(25)[:euler_steps]
    t0 = time.perf_counter()
    for t_long, dt in sched:
        t = torch.full((1,), t_long, dtype=torch.long)
        e_c, _ = R.dit_forward(sds[0], x, t, ctx_c, depth=4, n_heads=8)
        e_u, _ = R.dit_forward(sds[0], x, t, ctx_u, depth=4, n_heads=8)
        x = x + dt * (e_u + 3.0 * (e_c - e_u))
    t1 = time.perf_counter()
    mel = R.vae_decode(sds[1], x, synth.VAE_DDCONFIG)
    t2 = time.perf_counter()
    wav = R.hifigan_forward(sds[2], mel, synth.HIFIGAN_CFG)
    t3 = time.perf_counter()
    assert wav.shape[-1] == T_MEL * 320
    sampler = (t1 - t0) * (24.0 / euler_steps)
    return sampler + (t2 - t1) + (t3 - t2), {"sampler_s": sampler, "vae_s": t2 - t1, "vocoder_s": t3 - t2,
                                             "euler_steps_run": euler_steps}


def pick_cpu_threads(sds) -> int:
    """Thread count for the CPU arm: the fastest of {8,16,32,64,all usable cores} on a short DiT
    evaluation (T_lat=96). os.cpu_count() threads is often far from the best on a shared host."""
    from oracle import accompband_ref as R, synth
    torch.set_grad_enabled(False)
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    x, midi, beats, cap_c, _ = synth.dit_inputs(1, 96, seed=0)
    ctx = {"c_crossattn": cap_c, "c_concat": {"midi": midi, "beats": beats}}
    best, best_t = 1, float("inf")
    for n in sorted({min(c, usable) for c in (8, 16, 32, 64, usable)}):
        torch.set_num_threads(n)
        R.dit_forward(sds[0], x, torch.tensor([1]), ctx, depth=4, n_heads=8)
        t0 = time.perf_counter()
        R.dit_forward(sds[0], x, torch.tensor([1]), ctx, depth=4, n_heads=8)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def run_reference(args):
    """Reference arm: every step is ONE COMPLETE clip (24 Euler steps x 2 dense DiT evaluations + VAE decode +
    HiFi-GAN) through the oracle port on the host cores — nothing extrapolated. The number of steps actually
    run is bounded by a wall-clock budget so the arm ends within a few minutes; `steps` reports what ran."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sds = synth_weights()
    cores = pick_cpu_threads(sds)
    if args.warmup > 0:
        cpu_reference_step(sds, euler_steps=1)  # warm-up: thread pools, allocator
    times, brks = [], []
    t_budget = 120.0
    t_run = time.time()
    for i in range(max(args.steps, 1)):
        s, brk = cpu_reference_step(sds, euler_steps=24, seed=i)
        times.append(s)
        brks.append(brk)
        if time.time() - t_run + s > t_budget:  # the next clip would overrun the budget
            break
    sec = sum(times) / len(times)
    brk = {k: sum(b[k] for b in brks) / len(brks) for k in ("sampler_s", "vae_s", "vocoder_s")}
    value = CLIP_SECONDS / sec
    sample = (f"{len(times)} complete clip(s): 24 Euler steps (48 dense DiT evaluations) + VAE decode + HiFi-GAN each, "
              "fp32 torch CPU oracle port of the reference algorithm, nothing extrapolated")
    line = {
        "impl": "reference", "metric": "audio_sec_per_wall_sec", "value": value,
        "unit": "x real-time (24 kHz audio-s / wall-s)", "n_gpus": args.gpus, "steps": len(times),
        "steps_requested": args.steps,
        "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "AccompBand 25-point (24-step) CFG flow sampler + VAE decode + HiFi-GAN, 1 clip of 20 s, host CPU",
                   "clips_per_gpu": 1, "guidance": 3.0, "t_lat": T_LAT},
        "dit_latent_frame_evals_per_s": T_LAT * N_EVALS / brk["sampler_s"],
        "cpu_baseline": {"value": value, "unit": "x real-time", "cores": cores, "kind": "port", "sample": sample,
                         "breakdown_s": brk},
        "e2e": {"value": value, "unit": "x real-time (24 kHz audio-s / wall-s)", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region. Primary source: NVML polled in-process every
    ~5 ms (the timed region of a fast run is shorter than one `nvidia-smi -lms` period); the recipe's
    `nvidia-smi --query-gpu=... -lms` loop runs beside it (started before the warm-up) and is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.nv, self.h, self.samples, self.mask, self.max_mhz = None, None, [], 0, None
        self._stop = threading.Event()
        self._in_region = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                pr = torch.cuda.get_device_properties(index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._in_region.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.005)

    def start(self):
        """call before the warm-up steps"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        if self.nv:
            threading.Thread(target=self._poll, daemon=True).start()

    def region(self, on: bool):
        (self._in_region.set if on else self._in_region.clear)()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        self._stop.set()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        smi = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if self.samples:
            sm = sorted(self.samples)
            reasons |= {n for b, n in self.BITS.items() if self.mask & b}
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz or (max(mx) if mx else None),
                    "reasons": sorted(reasons), "samples": len(sm), "source": "NVML, 5 ms poll inside the timed region",
                    "nvidia_smi_samples_warmup_and_timed": len(smi)}
        if not smi:
            return {"sm_mhz": None, "sm_max_mhz": max(mx) if mx else None, "reasons": ["no clock samples"], "samples": 0}
        return {"sm_mhz": smi[len(smi) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(smi), "source": "nvidia-smi -lms 100 over warm-up + timed region"}


def run_ours(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — versband_b200 has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from oracle import synth
    from versband_b200 import _lib
    from versband_b200.pipeline import AccompBandPipeline

    sds = synth_weights()
    pipe = AccompBandPipeline(*sds, dit_cfg=synth.DIT_CFG, ddconfig=synth.VAE_DDCONFIG,
                              embed_dim=synth.VAE_EMBED_DIM, voc_cfg=synth.HIFIGAN_CFG, device=dev,
                              precision=args.precision)
    B = args.clips
    # per-clip seeding (not per rank): clip c of the job is the same whatever the GPU count
    x0, midi, beats, cap_c, cap_u = [], [], [], [], []
    for c in range(B):
        gid = rank * B + c
        a = synth.dit_inputs(1, T_LAT, seed=1000 + gid)
        for lst, t in zip((x0, midi, beats, cap_c, cap_u), a):
            lst.append(t)
    host = [torch.cat(l, 0).pin_memory() for l in (midi, beats, cap_c, cap_u, x0)]
    devt = [t.to(dev) for t in host]
    # per-clip seeding: clip `gid` of the job draws its router noise from its own generator,
    # so its waveform is the same on any GPU count / batch composition
    gen = [torch.Generator(device=dev).manual_seed(4242 + rank * B + c) for c in range(B)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device(events=None):
        return pipe.generate(devt[0], devt[1], devt[2], devt[3], x0=devt[4], generator=gen, events=events)

    wav_host = torch.empty(B, T_MEL * 320, dtype=torch.float32, pin_memory=True)

    if use_dist and rank == 0:
        # single-controller e2e: rank 0 owns the host buffers of ALL clips of the job
        allc = [[], [], [], [], []]
        for gid in range(B * world):
            for lst, t in zip(allc, synth.dit_inputs(1, T_LAT, seed=1000 + gid)):
                lst.append(t)
        host_all = [torch.cat(allc[i], 0).pin_memory() for i in (1, 2, 3, 4, 0)]
        wav_all = torch.empty(B * world, T_MEL * 320, dtype=torch.float32, pin_memory=True)

    # N > 1, one node: the controller's host buffers are shared page-locked memory and every GPU moves its own shard
    # over its own PCIe link (versband_b200/dist.py:HostExchange); set up collectively, NCCL scatter/gather otherwise
    hx = None
    if use_dist and args.e2e_exchange == "host":
        try:
            from versband_b200.dist import HostExchange
            specs = [(tuple(t.shape[1:]), t.dtype) for t in host_all] if rank == 0 else None
            hx = HostExchange(B * world, specs, ((T_MEL * 320,), torch.float32) if rank == 0 else None, device=dev)
            if rank == 0:
                for dst_, src_ in zip(hx.inputs, host_all):
                    dst_.copy_(src_)
        except RuntimeError as e:
            hx = None
            if rank == 0:
                print(f"[bench] HostExchange unavailable ({e}); e2e leg uses NCCL scatter/gather", file=sys.stderr)

    def step_host():
        if not use_dist:
            return pipe.generate_host(host[0], host[1], host[2], host[3], x0=host[4], generator=gen, out_wav=wav_host)
        if hx is not None:
            return hx.step(lambda loc: pipe.generate(loc[0], loc[1], loc[2], loc[3], x0=loc[4], generator=gen)["wav"])
        # N > 1: H2D on rank 0 -> NCCL scatter of the clips -> generate -> NCCL gather -> D2H on rank 0
        from versband_b200.dist import run_data_parallel
        ins = [t.to(dev, non_blocking=True) for t in host_all] if rank == 0 else None
        wav = run_data_parallel(
            lambda loc: pipe.generate(loc[0], loc[1], loc[2], loc[3], x0=loc[4], generator=gen)["wav"],
            ins, B * world, device=dev)
        if rank == 0:
            wav_all.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
        step_host()
    # ---- timed region A: inputs resident in HBM (value) ----
    barrier()
    clocks.region(True)
    tot_ms, samp_ms, vae_ms, voc_ms = 0.0, 0.0, 0.0, 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        ev = {}
        step_device(ev)
        torch.cuda.synchronize()
        tot_ms += ev["start"].elapsed_time(ev["end"])
        samp_ms += ev["sampler_start"].elapsed_time(ev["sampler_end"])
        vae_ms += ev["sampler_end"].elapsed_time(ev["vae_end"])
        voc_ms += ev["vae_end"].elapsed_time(ev["end"])
    barrier()
    # ---- timed region B: host buffers in/out through the public call (e2e) ----
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_host()
        b.record()
        torch.cuda.synchronize()
        e2e_ms += a.elapsed_time(b)
    barrier()
    clocks.region(False)
    clk = clocks.stop()
    # fingerprint of job clip 0 (rank 0, first clip): per-clip seeding makes it independent of the GPU count
    clip0_sha1 = None
    if rank == 0:
        import hashlib
        gen0 = [torch.Generator(device=dev).manual_seed(4242 + c) for c in range(B)]
        w0 = pipe.generate(devt[0], devt[1], devt[2], devt[3], x0=devt[4], generator=gen0)["wav"][0]
        clip0_sha1 = hashlib.sha1(w0.cpu().numpy().tobytes()).hexdigest()
    t = torch.tensor([tot_ms, samp_ms, vae_ms, voc_ms, e2e_ms], device=dev, dtype=torch.float64) / args.steps
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
    ms_step, ms_samp, ms_vae, ms_voc, ms_e2e = [float(v) for v in t]

    # ---- BASELINE configs[3] ("batch 32 over 8 GPUs" = 4 clips per GPU): same pipeline, 4 clips per rank ----
    cfg4 = None
    if not args.no_config4 and B != 4:
        B4 = 4
        lst = [[], [], [], [], []]
        for c in range(B4):
            for l_, t_ in zip(lst, synth.dit_inputs(1, T_LAT, seed=5000 + rank * B4 + c)):
                l_.append(t_)
        d4 = [torch.cat(l_, 0).to(dev) for l_ in (lst[1], lst[2], lst[3], lst[4], lst[0])]
        gen4 = [torch.Generator(device=dev).manual_seed(9000 + rank * B4 + c) for c in range(B4)]
        for _ in range(3):
            pipe.generate(d4[0], d4[1], d4[2], d4[3], x0=d4[4], generator=gen4)
        barrier()
        ms4 = 0.0
        n4 = max(3, args.steps // 2)
        for _ in range(n4):
            flush.fill_(1)
            ev = {}
            pipe.generate(d4[0], d4[1], d4[2], d4[3], x0=d4[4], generator=gen4, events=ev)
            torch.cuda.synchronize()
            ms4 += ev["start"].elapsed_time(ev["end"])
        barrier()
        t4 = torch.tensor([ms4 / n4], device=dev, dtype=torch.float64)
        if use_dist:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        cfg4 = {"workload": "BASELINE configs[3]: 4 clips per GPU (batch 32 at 8 GPUs), same pipeline, device-timed, max over ranks",
                "clips_per_gpu": B4, "global_clips": B4 * world, "steps": n4, "ms_per_step": float(t4[0]),
                "value": B4 * world * CLIP_SECONDS / (float(t4[0]) * 1e-3), "unit": "x real-time (24 kHz audio-s / wall-s)"}
        del d4

    # ---- roofline (rank 0): every time below is a CUDA-event interval of the TIMED (graph-replay) region ----
    roof, launches, fam = None, None, None
    if rank == 0:
        import json as _json
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = _json.load(f)
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
        peak_bw = peaks.get("hbm_gbs") or 6650.0
        peak_src = ("measured (MEASURED_PEAKS.json: sustained 16-bit dense tensor TFLOP/s, STREAM-copy GB/s)" if peaks
                    else "fallback 1.4 PF sustained / 6.65 TB/s")
        # Stage = a contiguous run of this repo's kernels bracketed by two events on the launching stream
        # (pipeline.generate(events=...)): sampler graph | VAE-decode graph | vocoder graph. The algorithmic
        # work per stage is SURVEY.md §8d's per-clip figure x clips. No eager-pass share is used any more.
        stages = {
            "dit_sampler": {"kernels": "vb_gemm_kernel / vb_gemm2_kernel (tcgen05 GEMMs) + attn_fused_kernel + norm/router/dispatch",
                            "bound": "tensor", "ms": ms_samp, "algorithmic_flop": FLOP_SAMPLER * B,
                            "flop_gemm": (FLOP_SAMPLER - FLOP_ATTN) * B, "flop_attention": FLOP_ATTN * B,
                            "achieved": FLOP_SAMPLER * B / (ms_samp * 1e-3) / 1e12, "unit": "TFLOP/s", "peak": peak_tf},
            "vae_decode": {"kernels": "vb_gemm2_kernel implicit-GEMM convs + GroupNorm", "bound": "tensor", "ms": ms_vae,
                           "algorithmic_flop": FLOP_VAE * B, "achieved": FLOP_VAE * B / (ms_vae * 1e-3) / 1e12,
                           "unit": "TFLOP/s", "peak": peak_tf},
            "vocoder": {"kernels": "HiFi-GAN convs (vb_gemm*/vb_conv_halo/fused resblock) + conv_post", "bound": "hbm", "ms": ms_voc,
                        "algorithmic_bytes": BYTES_VOC_16 * B, "achieved": BYTES_VOC_16 * B / (ms_voc * 1e-3) / 1e9,
                        "unit": "GB/s", "peak": peak_bw,
                        "tensor_achieved_tflops": FLOP_VOC * B / (ms_voc * 1e-3) / 1e12,
                        "tensor_frac": FLOP_VOC * B / (ms_voc * 1e-3) / 1e12 / peak_tf},
        }
        for st_ in stages.values():
            st_["frac"] = st_["achieved"] / st_["peak"]
        # launch count of one step (an eager pass; the timed region replays the same launches from graphs)
        from versband_b200.sampler import FlowSampler
        saved = pipe.sampler
        pipe.sampler = FlowSampler(pipe.dit, use_graph=False)
        pipe.use_graph = False
        step_device()
        n0 = _lib.launch_count()
        _lib.prof_enable(True)
        step_device()
        fam = _lib.prof_report()
        _lib.prof_enable(False)
        launches = _lib.launch_count() - n0
        pipe.sampler = saved
        pipe.use_graph = True
        traffic, traffic_src = None, None
        for name in ("r2_traffic_step.json", "r1_traffic_eager_step.json"):
            try:  # dram bytes of the GEMM family over one step, from the committed ncu capture
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tj = _json.load(f)["vb_gemm_kernel"]
                traffic, traffic_src = tj["dram_bytes"] / tj["launches"], "profiles/" + name
                break
            except Exception:
                pass
        dom = stages["dit_sampler"]
        n_dit = 24 * 2  # evaluations; launches of the stage = launches - prepare - vae - vocoder (see launch lists)
        roof = {"bound": "tensor",
                "kernel": "DiT sampler stage: tcgen05 GEMM family (vb_gemm_kernel / vb_gemm2_kernel) + attn_fused_kernel, "
                          "48 evaluations replayed from one CUDA graph; the stage is 75-80 % of the step",
                "achieved": dom["achieved"], "peak": peak_tf, "unit": "TFLOP/s", "frac": dom["frac"],
                "peak_source": peak_src, "traffic": traffic,
                "traffic_note": f"mean dram read+write bytes per GEMM launch (ncu, cold), {traffic_src}",
                "method": "algorithmic FLOPs of the stage (SURVEY §8d: 48 x 51.0 + 22 GFLOP per clip, minimal top-1 routed count) / "
                          "CUDA-event time of the stage inside the timed region; every kernel of the stage counts against the "
                          "tensor roof, so this is a lower bound on the GEMM family's own fraction",
                "stages": stages, "dit_evaluations_per_step": n_dit,
                "eager_event_timed_family_ms": {k: round(v["ms"], 3) for k, v in fam.items()}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = pick_cpu_threads(sds)
        sec, brk = cpu_reference_step(sds, euler_steps=24)
        cpu = {"value": CLIP_SECONDS / sec, "unit": "x real-time", "cores": cores, "kind": "port",
               "sample": "1 complete clip: 24 Euler steps (48 dense DiT evaluations) + VAE decode + HiFi-GAN, fp32 oracle port, "
                         "nothing extrapolated",
               "breakdown_s": brk}

    if rank == 0:
        clips = B * world
        audio = clips * CLIP_SECONDS
        line = {
            "metric": "audio_sec_per_wall_sec", "value": audio / (ms_step * 1e-3),
            "unit": "x real-time (24 kHz audio-s / wall-s)", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "AccompBand 25-point (24-step) CFG flow sampler + VAE decode + HiFi-GAN, 20 s clips",
                       "clips_per_gpu": B, "global_clips": clips, "guidance": 3.0, "t_lat": T_LAT,
                       "parallelism": f"dp{world}",
                       "precision": ("IEEE fp16 GEMM operands (saturating), fp32 accumulate / residual / norms / softmax / router"
                                     if args.precision == "fp16" else
                                     "DiT linears: fp32 operands as TF32 (kind::tf32), fp16 attention core; VAE/vocoder fp16 operands"),
                       "l2": "256 MiB buffer written between timed steps (L2 flush)"},
            "dit_latent_frame_evals_per_s": clips * T_LAT * N_EVALS / (ms_samp * 1e-3),
            "stage_ms": {"sampler": ms_samp, "vae_decode": ms_vae, "vocoder": ms_voc},
            "e2e": {"value": audio / (ms_e2e * 1e-3), "unit": "x real-time (24 kHz audio-s / wall-s)",
                    "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in (host_all if use_dist else host))),
                    "d2h_bytes_per_step": int((wav_all if use_dist else wav_host).numel() * 4),
                    "path": ("shared pinned host -> per-rank H2D of its shard -> generate -> per-rank D2H -> sequence counters (no collective)"
                             if use_dist and hx is not None else
                             "pinned host -> H2D -> NCCL scatter -> generate -> NCCL gather -> D2H (rank 0)" if use_dist
                             else "pinned host -> H2D -> generate -> D2H")},
            "gpu_launches": int(launches) * args.steps * 2 if launches else None,
            "gpu_launches_per_step": int(launches) if launches else None,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk, "clip0_wav_sha1": clip0_sha1, "config4": cfg4,
        }
        print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=1, help="clips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-exchange", default="host", choices=["host", "nccl"],
                    help="N > 1 e2e leg: shared pinned host memory (each GPU moves its own shard) or NCCL scatter/gather via rank 0")
    ap.add_argument("--no-config4", action="store_true", help="skip the extra 4-clips-per-GPU measurement (BASELINE configs[3])")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "tf32"],
                    help="GEMM operand format of the DiT: fp16 (default) or fp32-as-TF32 (twice the operand bytes)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
