"""Intra-kernel timeline of every GEMM of one DiT evaluation, replayed from a CUDA graph: the 1-CTA
GEMM kernel stamps %globaltimer at its phase boundaries (vb_set_gemm_trace). Prints, per GEMM call of
block 0, the median over CTAs of each phase and the kernel's span (first entry -> last exit)."""
import ctypes
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
L.vb_set_gemm2(0)
sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
eng = DiTEngine(sd, precision=precision)
B, T = 1, 750
x, midi, beats, cap_c, cap_u = [t.cuda() for t in synth.dit_inputs(B, T, seed=1)]
ac = eng.prepare_acoustic(midi, beats, T)
cond2 = DiTCond.cat([eng.prepare(None, None, cap_c, T, acoustic=ac), eng.prepare(None, None, cap_u, T, acoustic=ac)])
expo = [tuple(e.cuda() for e in blk) for blk in synth.exp_noise(4, 2 * B * T, 4, seed=2)]
t = torch.tensor([500] * (2 * B), device="cuda")
eng.forward(x, t, cond2, expo)
torch.cuda.synchronize()
MAXC = 64
buf = torch.zeros(MAXC, 256, 8, dtype=torch.int64, device="cuda")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    L.vb_set_gemm_trace(ctypes.c_void_p(buf.data_ptr()), MAXC)
    eng.forward(x, t, cond2, expo)
    L.vb_set_gemm_trace(None, 0)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
rec = buf.cpu()
names = ["proj_in", "t_w0", "t_w2", "ada"] + [f"b{k}.{n}" for k in range(4) for n in ("wqkv", "wo", "route", "w13", "w2", "f13", "f2")] + ["final"]
print("times in us; entry->pro = prologue, ->data = first operands, ->mma = first tile mainloop issue, "
      "->acc = accumulator ready, ->epi = first tile epilogue, ->exit; span = first entry to last exit over CTAs")
prev_end = None
for i in range(MAXC):
    r = rec[i]
    live = r[:, 0] > 0
    if not live.any():
        break
    r = r[live].double()
    t0 = r[:, 0].min()
    span = (r[:, 6].max() - t0) / 1e3
    med = lambda a: float(a.median()) / 1e3  # noqa: E731
    gap = (float(t0) - prev_end) / 1e3 if prev_end is not None else 0.0
    prev_end = float(r[:, 6].max())
    nm = names[i] if i < len(names) else f"#{i}"
    print(f"{nm:10s} ctas={int(live.sum()):3d} gap_before={gap:6.2f} span={span:6.2f} | start_skew={med(r[:,0]-t0):5.2f} pro={med(r[:,1]-r[:,0]):5.2f} "
          f"data={med(r[:,2]-r[:,1]):5.2f} mma={med(r[:,3]-r[:,2]):5.2f} acc={med(r[:,4]-r[:,3]):5.2f} "
          f"epi={med(r[:,5]-r[:,4]):5.2f} rest={med(r[:,6]-r[:,5]):5.2f} cta_total={med(r[:,6]-r[:,0]):5.2f}")
