"""In-situ tile tuning: for every DiT GEMM signature (N, K, grouped) force each candidate tile shape
through vb_set_gemm_override, capture ONE DiT evaluation (the bench shape: 2B=2, T=750) in a CUDA
graph and time its replay. The isolated micro-benchmarks (tools/bench_bn.py) mispredict the
in-sequence cost, so the table in vb_gemm's cost model is checked against this."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from versband_b200 import _lib  # noqa: E402
from versband_b200.dit import DiTCond, DiTEngine  # noqa: E402

torch.set_grad_enabled(False)
precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
L = _lib.lib()
sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
eng = DiTEngine(sd, precision=precision)
B, T = int(os.environ.get("TUNE_B", "1")), 750
x, midi, beats, cap_c, cap_u = [t.cuda() for t in synth.dit_inputs(B, T, seed=1)]
ac = eng.prepare_acoustic(midi, beats, T)
cond2 = DiTCond.cat([eng.prepare(None, None, cap_c, T, acoustic=ac), eng.prepare(None, None, cap_u, T, acoustic=ac)])
expo = [tuple(e.cuda() for e in blk) for blk in synth.exp_noise(4, 2 * B * T, 4, seed=2)]
t = torch.tensor([500] * (2 * B), device="cuda")
REP = 4


def measure():
    eng.forward(x, t, cond2, expo)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            eng.forward(x, t, cond2, expo)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 5 / REP * 1e3)
    return best


D, H3, Fh = eng.D, 3 * eng.D, None
sigs = {}
for k, blk in enumerate(eng.w["blocks"][:1]):
    for name in ("wqkv", "wo", "w13", "w2", "f13", "f2"):
        wt = blk[name]
        N, K = int(wt.shape[-2]), int(wt.shape[-1])
        grouped = 1 if name in ("w13", "w2") else 0
        sigs.setdefault((N, K, grouped), []).append(name)

base = measure()
print(f"precision={precision} baseline (cost model): {base:.1f} us per DiT evaluation", flush=True)
for (N, K, grouped), names in sigs.items():
    res = []
    for cta2 in (0, 1):
        if cta2 and grouped:
            continue
        for bn in (64, 96, 128, 192, 256):
            if cta2 and bn < 128:
                continue
            if bn > N:
                continue
            L.vb_set_gemm_override(N, K, grouped, bn, cta2)
            try:
                us = measure()
            except Exception as e:  # noqa: BLE001
                us = float("nan")
                print("   failed:", e)
            res.append((us - base, bn, cta2))
            L.vb_set_gemm_override(0, 0, 0, 0, 0)
    res.sort()
    line = "  ".join(f"bn{bn}{'x2' if c2 else ''}:{d:+.1f}" for d, bn, c2 in res)
    print(f"{'/'.join(names):10s} N={N:5d} K={K:5d} grouped={grouped}: {line}", flush=True)
